"""B200-native hot path of christofw/multipitch_architectures: HCQT + patch-wise multi-pitch networks on hand-written sm_100a kernels
(libmpa.so, C ABI in include/mpa.h) behind the reference's own `libdl` Python API (the mirror lives in `.libdl`)."""


def install_as_libdl(name='libdl', precision=None):
    """Make `import libdl...` resolve to this package's mirror, so that a reference experiment script or notebook runs with its import
    lines untouched (`from libdl.nn_models import deep_cnn_segm_sigmoid`, `from libdl.data_loaders import dataset_context`,
    `from libdl.metrics import calculate_eval_measures`, `from libdl.data_preprocessing import compute_efficient_hcqt`).
    precision: default `precision` of model classes constructed WITHOUT the keyword (the reference scripts never pass it):
    'fp32' (exact CUDA-core path), 'fp16' / 'bf16' (tcgen05, 16-bit operands) or 'fp16x3' (tcgen05, split precision)."""
    import importlib
    import pkgutil
    import sys
    from . import libdl as pkg
    if precision is not None:
        set_default_precision(precision)
    sys.modules[name] = pkg
    for m in pkgutil.walk_packages(pkg.__path__, pkg.__name__ + '.'):
        sys.modules[name + m.name[len(pkg.__name__):]] = importlib.import_module(m.name)
    return pkg


def set_default_precision(precision):
    """Precision of models constructed without an explicit `precision=` (initially 'fp32', or the MPA_PRECISION environment variable)."""
    from .libdl.nn_models import _exec
    if precision not in _exec.PRECISIONS:
        raise ValueError(f'precision must be one of {_exec.PRECISIONS}')
    _exec.DEFAULT_PRECISION = precision
