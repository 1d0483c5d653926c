"""`-m gpu`: edge cases of the path — empty and single-frame inputs, recordings shorter than a patch, very short audio, ragged chunking,
degenerate evaluation inputs, and loud failure on what the kernels cannot take."""
import numpy as np
import pytest
import torch

from oracle import hcqt_oracle as Q
from oracle import host_oracle as HO
from oracle import nn_oracle as NO
from tests.refshapes import build_model
from tests.weights import fill_state_dict

pytestmark = pytest.mark.gpu


def _engine(chunk):
    from multipitch_architectures_b200.engine import CnnStreamEngine
    m = build_model('drcnn_tiny', precision='fp16')
    sd = fill_state_dict(m.state_dict(), 21)
    m.load_state_dict(sd)
    return CnnStreamEngine(m.cuda().eval(), chunk=chunk), sd


@pytest.mark.parametrize('N,chunk', [(1, 8), (3, 2), (74, 64), (75, 37), (129, 64)])
def test_engine_on_short_and_ragged_recordings(N, chunk):
    """Recordings shorter than one 75-frame patch (all context is zero padding) and lengths that leave a ragged last chunk."""
    eng, sd = _engine(chunk)
    rng = np.random.default_rng(N)
    h = (np.abs(rng.normal(0, 0.05, size=(6, N, 216))) * rng.uniform(0.3, 2.0, size=(1, N, 1))).astype(np.float32)
    got = eng.predict_hcqt(torch.from_numpy(h).cuda()).cpu().numpy()
    ip, _ = HO.pad_for_inference(h, np.zeros((N, 72)))
    X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(N)]))
    with torch.no_grad():
        ref = NO.cnn_forward(sd, X, residual=True).reshape(N, 72).numpy()
    assert got.shape == (N, 72) and np.abs(got - ref).max() < 1e-2


def test_engine_empty_range_and_bad_shapes():
    eng, _ = _engine(16)
    h = torch.rand(6, 20, 216, device='cuda')
    assert eng.predict_hcqt(h, lo=7, hi=7).numel() == 0
    with pytest.raises(ValueError):
        eng.predict_hcqt(h, lo=5, hi=3)
    with pytest.raises(ValueError):
        eng.predict_hcqt(torch.rand(5, 20, 216, device='cuda'))
    with pytest.raises(ValueError):
        eng.predict_hcqt(torch.rand(6, 20, 215, device='cuda'))


def test_hcqt_on_very_short_audio():
    """A clip of a few hundred milliseconds: fewer samples than the longest filter (reflect padding wraps several times), and the
    loud failure below the tuning estimator's minimum."""
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_efficient_hcqt
    kw = dict(fs=22050, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)
    for n in (4000, 1537, 1030):
        y = Q.synth_clip(n, seconds=0.5)[:n]
        f, _, hop = compute_efficient_hcqt(y, **kw)
        ref, _, _ = Q.compute_efficient_hcqt(y, **kw)
        assert f.shape == ref.shape == (216, n // 512 + 1, 6)
        assert np.abs(f - ref).max() < 2e-4 * max(ref.max(), 1e-6)
    with pytest.raises(_lib.MpaError):
        compute_efficient_hcqt(np.zeros(512, np.float32), **kw)


def test_hcqt_of_silence_is_zero_and_tuning_defaults():
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_efficient_hcqt, estimate_tuning
    y = np.zeros(22050, np.float32)
    f, _, _ = compute_efficient_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    assert f.shape == (216, 44, 6) and (f == 0).all()
    assert estimate_tuning(y, 22050, 36) == 0.0 == Q.estimate_tuning(y, bins_per_octave=36)


def test_eval_measures_degenerate_inputs():
    from multipitch_architectures_b200.libdl.metrics import calculate_eval_measures, calculate_mpe_measures_mireval
    names = ['precision', 'recall', 'f_measure', 'cosine_sim', 'binary_crossentropy', 'euclidean_distance', 'binary_accuracy', 'soft_accuracy',
             'accum_energy']
    for targ, pred in ((np.zeros((1, 72)), np.zeros((1, 72), np.float32)),                 # one silent frame, silent estimate
                       (np.ones((3, 72)), np.ones((3, 72), np.float32)),                    # everything active
                       (np.eye(72)[:5], np.zeros((5, 72), np.float32))):                    # nothing detected
        d = calculate_eval_measures(targ, pred, names, threshold=0.4)
        for n in names:
            want = HO.eval_measure(targ, pred.astype(np.float64), n, 0.4)
            assert abs(d[n] - want) < 1e-12, (n, d[n], want)
        got, want = calculate_mpe_measures_mireval(targ, pred, 0.4), HO.mpe_scores(targ, pred, 0.4)
        assert all(abs(got[k] - want[k]) < 1e-12 for k in want)


def test_models_reject_what_they_cannot_run():
    from multipitch_architectures_b200 import _lib
    m = build_model('cnn_xs').cuda().eval()
    with pytest.raises(_lib.MpaError):
        m(torch.rand(1, 6, 75, 216))                       # CPU tensor: there is no CPU path
    with pytest.raises(ValueError):
        m(torch.rand(1, 5, 75, 216, device='cuda'))
    u = build_model('unet_tiny').cuda().eval()
    with pytest.raises(ValueError):
        u(torch.rand(1, 6, 40, 216, device='cuda'))       # fewer than 75 frames
    # longer inputs are legal for the CNN family: [B,6,T,216] -> [B,1,T-74,72] (fully convolutional in time, as the reference modules)
    x = torch.rand(2, 6, 80, 216, device='cuda')
    with torch.no_grad():
        y = m(x)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    ref = NO.cnn_forward(sd, x.cpu())
    assert y.shape == (2, 1, 6, 72) and (y.cpu() - ref).abs().max() < 1e-3
