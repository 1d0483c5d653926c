"""`-m gpu`: parity on the REALISTIC weight set (trained models whose outputs span 0..1 with a real decision structure, F-measure 0.90-0.96
against the labels) over a whole 30 s clip = 1,292 stride-1 patches, full-size models, against the outputs of the UNMODIFIED reference
classes (tests/golden/realistic_golden.npz).  Gates: the north star's 1e-3 for fp32, fp16x3 AND plain fp16; the looser stated bound for
bf16; thresholded activity / TP, FP, FN / P, R, F identical (compared unconditionally; flips are counted and must be zero)."""
import numpy as np
import pytest
import torch

from tests import realistic as R
from tests.refshapes import build_model

pytestmark = pytest.mark.gpu

TOL = {'fp32': 1e-3, 'fp16x3': 1e-3, 'fp16': 1e-3, 'bf16': 2e-2}


@pytest.fixture(scope='module')
def hcqt():
    """Oracle HCQT of the held-out clip (the input the reference goldens were made from), checked against the probes stored with them."""
    from oracle import hcqt_oracle as HO
    from tests import synth
    y = synth.synth_clip(**R.CLIP)
    f, _, _ = HO.compute_efficient_hcqt(y, **R.HCQT_KW)
    g = R.golden()
    assert f.shape[1] == int(g['n_frames'][0])
    assert np.abs(f[::37, ::101, :] - g['hcqt_probe']).max() <= 1e-6 * g['hcqt_probe'].max()
    return torch.from_numpy(np.ascontiguousarray(np.transpose(f, (2, 1, 0)).astype(np.float32))).cuda()


def _model(name, prec):
    m = build_model(name, precision=prec)
    m.load_state_dict(R.state_dict(name))          # strict: the reference's key names, shapes and dtypes
    return m.cuda().eval()


def _predict(m, name, prec, hcqt):
    from multipitch_architectures_b200.engine import CnnStreamEngine, predict_patchwise
    with torch.no_grad():
        if name != 'unet_m' and prec != 'fp32':
            return CnnStreamEngine(m).predict_hcqt(hcqt).cpu().numpy()          # the benchmarked path
        return predict_patchwise(m, hcqt, batch=50).cpu().numpy()               # the reference's loop: batches of 50 patches


@pytest.mark.parametrize('prec', ['fp32', 'fp16x3', 'fp16', 'bf16'])
@pytest.mark.parametrize('name', R.MODELS)
def test_realistic_weights_full_clip_vs_reference(hcqt, name, prec):
    got = _predict(_model(name, prec), name, prec, hcqt)
    c = R.compare(got, name)
    print(f"{name} {prec}: max|diff| = {c['max_abs']:.2e} over {c['frames']} frames (outputs span {c['out_span'][0]:.1e}..{c['out_span'][1]:.3f}), "
          f"flips {c['flips']} (margin {c['flip_margin']:.1e}), TP/FP/FN {c['counts']} vs {c['counts_ref']}, P/R/F {tuple(round(v, 4) for v in c['prf'])}")
    assert got.shape == (int(R.golden()['n_frames'][0]), 72)
    assert c['max_abs'] < TOL[prec]
    assert c['flip_margin'] < TOL[prec]            # a flipped cell must sit within the tolerance of the threshold
    assert c['prf_equal_3dec']                     # unconditional: P/R/F to 3 decimals
    if prec != 'bf16':
        assert c['flips'] == 0 and c['counts'] == c['counts_ref']


def test_realistic_module_forward_equals_engine(hcqt):
    """model(batch) on materialised patches (the drop-in call) and the streaming engine give the same activations (fp16x3: to rounding)."""
    from multipitch_architectures_b200.engine import CnnStreamEngine, predict_patchwise
    m = _model('drcnn', 'fp16x3')
    with torch.no_grad():
        a = CnnStreamEngine(m).predict_hcqt(hcqt, 0, 200).cpu().numpy()
        b = predict_patchwise(m, hcqt[:, :237], batch=50).cpu().numpy()[:200]
    assert np.abs(a[:163] - b[:163]).max() < 1e-5      # frames whose 75-frame context lies inside the 237-frame excerpt
