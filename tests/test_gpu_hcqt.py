"""`-m gpu`: HCQT path (H0-H3) through the C ABI against the NumPy oracle (parity unpinned vs real librosa, see
oracle/hcqt_oracle.py) and the committed golden produced with the reference's own wrapper code."""
import numpy as np
import pytest
import torch

from oracle import hcqt_oracle as Q

pytestmark = pytest.mark.gpu

KW = dict(fs=22050, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)
REL_TOL = 2e-4       # of the clip's largest HCQT magnitude (fp32 FFT + fp32 contraction vs the complex64 oracle)


def test_decimator_matches_oracle():
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.data_preprocessing import _filterbank as FB
    for n in (4001, 10000, 65):
        y = np.random.default_rng(n).standard_normal(n).astype(np.float32)
        ref = Q.resample_2to1(y)
        yd = torch.from_numpy(y).cuda()
        out = torch.empty((n + 1) // 2, dtype=torch.float32, device='cuda')
        _lib.call('decimate2_f32', yd, out, torch.from_numpy(FB.kaiser_fast_half_taps()).cuda(), _lib.i64(n), _lib.stream_ptr())
        assert out.numel() == len(ref)
        assert np.abs(out.cpu().numpy() - ref).max() < 1e-6


@pytest.mark.parametrize('seed', [0, 1, 2, 3, 7])
def test_estimate_tuning_matches_oracle(seed):
    from multipitch_architectures_b200.libdl.data_preprocessing import estimate_tuning
    y = Q.synth_clip(seed, seconds=6.0)
    assert abs(estimate_tuning(y, 22050, 36) - Q.estimate_tuning(y, bins_per_octave=36)) < 1e-9


def test_hcqt_matches_reference_wrapper_golden(host_golden):
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_efficient_hcqt
    y = Q.synth_clip(int(host_golden['hcqt_clip_seed'][0]), seconds=2.0)
    f, fs_h, hop = compute_efficient_hcqt(y, **KW)
    gold = host_golden['hcqt_2s']
    assert f.dtype == np.float64 and f.shape == gold.shape == (216, len(y) // 512 + 1, 6)
    assert hop == 512 and fs_h == 22050 / 512
    err = np.abs(f - gold).max()
    print('hcqt 2 s clip: max|diff| =', err, 'of max', gold.max())
    assert err < REL_TOL * gold.max()


@pytest.mark.parametrize('seed,seconds', [(11, 5.0), (12, 3.3)])
def test_hcqt_matches_oracle_fresh_clips(seed, seconds):
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_efficient_hcqt
    y = Q.synth_clip(seed, seconds=seconds)
    f, _, _ = compute_efficient_hcqt(y, **KW)
    ref, _, _ = Q.compute_efficient_hcqt(y, **KW)
    assert f.shape == ref.shape
    assert np.abs(f - ref).max() < REL_TOL * ref.max()


def test_hcqt_full_size_properties():
    """BASELINE config-1 size (30 s = 661,500 samples -> 1,292 frames): size-independent properties."""
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_efficient_hcqt
    y = Q.synth_clip(5, seconds=30.0)
    f, fs_h, hop = compute_efficient_hcqt(y, **KW)
    assert f.shape == (216, 1292, 6) and np.isfinite(f).all() and (f >= 0).all()
    assert np.array_equal(f[72:, :, 1], f[:144, :, 4])          # h=4 is h=1 shifted by two octaves (same CQT)
    assert np.array_equal(f[36:, :, 0], f[:180, :, 1])          # h=1 is the sub-harmonic shifted by one octave
    # linearity in the input (tuning index held by scaling invariance of estimate_tuning)
    g, _, _ = compute_efficient_hcqt(0.5 * y, **KW)
    assert np.abs(g - 0.5 * f).max() < 1e-5 * f.max()
    # pure tone: peak at 3*(p-24)+1 in the fundamental channel
    t = np.arange(3 * 22050) / 22050
    tone = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
    h, _, _ = compute_efficient_hcqt(tone, **KW)
    assert abs(int(np.argmax(h[:, h.shape[1] // 2, 1])) - (3 * (69 - 24) + 1)) <= 1


def test_hopsize_and_annotation_host_functions(host_golden):
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_hopsize_cqt, compute_annotation_array_nooverlap
    for target, noct, hop, fs in host_golden['hopsize']:
        assert compute_hopsize_cqt(target, 22050, int(noct)) == (int(hop), fs)
    ev = host_golden['annot2_events']
    A = compute_annotation_array_nooverlap(ev.copy(), np.zeros((216, 200, 6)), 22050 / 512, annot_type='pitch')
    assert np.array_equal(np.packbits(A.astype(np.uint8)), host_golden['annot2'])
