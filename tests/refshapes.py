"""State-dict key order / shapes of the reference models, derived from the product modules (which must mirror
them exactly: the golden fixtures' weight checksums fail otherwise).  CPU only."""
import functools


@functools.lru_cache(maxsize=None)
def _build(name):
    from multipitch_architectures_b200.libdl import nn_models as M
    from tests.weights import MODEL_SPECS
    spec = MODEL_SPECS[name]
    return getattr(M, spec['cls'])(**spec['kw'])


def build_model(name, **extra):
    from multipitch_architectures_b200.libdl import nn_models as M
    from tests.weights import MODEL_SPECS
    spec = MODEL_SPECS[name]
    return getattr(M, spec['cls'])(**spec['kw'], **extra)


def reference_state_shapes(name):
    return _build(name).state_dict()
