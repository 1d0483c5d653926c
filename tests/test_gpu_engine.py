"""`-m gpu`: the streaming inference engine (gather-from-frames, tcgen05 stack) and the reference-shaped patch loop
against the oracle's patch-wise evaluation; thresholded activity and P/R/F to 3 decimals."""
import numpy as np
import pytest
import torch

from oracle import host_oracle as HO
from oracle import nn_oracle as NO
from tests.refshapes import build_model
from tests.weights import fill_state_dict

pytestmark = pytest.mark.gpu


def oracle_patchwise(sd, hcqt, residual):
    C, N, F = hcqt.shape
    ip, _ = HO.pad_for_inference(hcqt, np.zeros((N, 72)))
    X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(N)]))
    with torch.no_grad():
        return NO.cnn_forward(sd, X, residual=residual).reshape(N, 72).numpy()


def fake_hcqt(N, seed):
    rng = np.random.default_rng(seed)
    h = np.abs(rng.normal(0, 0.05, size=(6, N, 216))) * (1 + np.sin(np.arange(216) / 7.0) ** 2)[None, None, :]
    h *= rng.uniform(0.3, 2.0, size=(1, N, 1))
    return h.astype(np.float32)


TOL = {'fp16': 1e-2, 'bf16': 6e-2}


@pytest.mark.parametrize('prec', ['fp16', 'bf16'])
@pytest.mark.parametrize('name,N,chunk', [('drcnn_tiny', 90, 64), ('cnn_xs', 40, 592)])
def test_stream_engine_matches_oracle_patchwise(name, N, chunk, prec):
    from multipitch_architectures_b200.engine import CnnStreamEngine
    m = build_model(name, precision=prec)
    sd = fill_state_dict(m.state_dict(), 21)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    h = fake_hcqt(N, 3)
    got = CnnStreamEngine(m, chunk=chunk).predict_hcqt(torch.from_numpy(h).cuda()).cpu().numpy()
    ref = oracle_patchwise(sd, h, getattr(m, 'residual', False))
    err = np.abs(got - ref).max()
    print(f'{name}: streaming {prec} engine vs oracle max|diff| = {err:.2e}')
    assert got.shape == (N, 72) and err < TOL[prec]
    targ = np.random.default_rng(1).uniform(size=(N, 72)) < 0.3
    p_ref, p_got = HO.eval_prf(targ, ref, 0.4), HO.eval_prf(targ, got, 0.4)
    flips = (got >= 0.4) != (ref >= 0.4)
    near = np.abs(ref - 0.4) < TOL[prec]
    print(f'{name} {prec}: {int(flips.sum())} threshold flips of {flips.size} cells ({int(near.sum())} cells within the {TOL[prec]} bound of 0.4); '
          f'P/R/F {tuple(round(v, 4) for v in p_got[:3])} vs {tuple(round(v, 4) for v in p_ref[:3])}')
    assert not (flips & ~near).any()                 # a decision may only change where the reference sits within the format's bound of 0.4
    # P/R/F: every flip moves one of TP / FP / FN by one, so the measures may move by at most flips / (smallest denominator); these are
    # the ADVERSARIAL-gain random weights (the format's worst case) — the 3-decimal identity is asserted unconditionally on the
    # trained weights in tests/test_gpu_realistic.py and, for the split-precision mode, on these weights in tests/test_gpu_x3.py
    denom = max(1, min(p_ref[3] + p_ref[4], p_ref[3] + p_ref[5]) - int(flips.sum()))
    assert all(abs(a - b) <= 2.0 * flips.sum() / denom + 1e-12 for a, b in zip(p_ref[:3], p_got[:3]))


def test_patchwise_loop_fp32_matches_oracle_and_prf():
    from multipitch_architectures_b200.engine import predict_patchwise
    m = build_model('drcnn_tiny')
    sd = fill_state_dict(m.state_dict(), 22)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    h = fake_hcqt(70, 4)
    got = predict_patchwise(m, torch.from_numpy(h).cuda(), batch=50).cpu().numpy()
    ref = oracle_patchwise(sd, h, True)
    assert np.abs(got - ref).max() < 1e-3
    targ = np.random.default_rng(2).uniform(size=ref.shape) < 0.3
    flips = (got >= 0.4) != (ref >= 0.4)
    assert not flips.any(), f'{int(flips.sum())} thresholded cells differ (closest reference value to 0.4: {np.abs(ref - 0.4).min():.2e})'
    assert HO.eval_prf(targ, got, 0.4) == HO.eval_prf(targ, ref, 0.4)          # unconditional: identical counts, hence identical P/R/F


def test_dataset_context_batch_kernel(host_golden):
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    inp, tg = host_golden['ds_in'], host_golden['ds_tg']
    ip, tp = HO.pad_for_inference(inp, tg)
    ds = dataset_context(torch.from_numpy(ip).cuda(), torch.from_numpy(tp).cuda(), {'context': 75, 'stride': 1, 'compression': 10})
    X, y = ds.batch(0, 40)
    for j, i in enumerate(host_golden['ds_idx']):
        assert np.abs(X[int(i)].cpu().numpy() - host_golden['ds_X'][j]).max() < 1e-6
        assert np.array_equal(y[int(i)].cpu().numpy(), host_golden['ds_y'][j])
    # the per-item Dataset protocol on HOST tensors (what the reference scripts construct): items come back from the same kernels
    dh = dataset_context(torch.from_numpy(ip.astype(np.float64)), torch.from_numpy(tp.astype(np.float64)), {'context': 75, 'stride': 1, 'compression': 10})
    assert len(dh) == int(host_golden['ds_len'][0])
    for j, i in enumerate(host_golden['ds_idx']):
        Xi, yi = dh[int(i)]
        assert Xi.is_cuda and Xi.dtype == torch.float32 and tuple(Xi.shape) == (6, 75, 216) and tuple(yi.shape) == (1, 1, 72)
        assert np.abs(Xi.cpu().numpy() - host_golden['ds_X'][j]).max() < 1e-6 and np.array_equal(yi.cpu().numpy(), host_golden['ds_y'][j])
    d3 = dataset_context(torch.from_numpy(ip.astype(np.float64)), torch.from_numpy(tp.astype(np.float64)), {'context': 75, 'stride': 3, 'compression': None})
    X5, y5 = d3[5]
    assert abs(float(X5.double().sum()) - float(host_golden['ds3_X5_sum'][0])) < 1e-3 and np.array_equal(y5.cpu().numpy(), host_golden['ds3_y5'])


def test_audio_to_activations_end_to_end():
    """audio -> HCQT -> DRCNN(tiny) on the GPU vs the oracle chain on the same clip (fp16 bound)."""
    from oracle import hcqt_oracle as Q
    from multipitch_architectures_b200.engine import CnnStreamEngine
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import get_plan, C1_HZ
    m = build_model('drcnn_tiny', precision='fp16')
    sd = fill_state_dict(m.state_dict(), 23)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    y = Q.synth_clip(8, seconds=2.5)
    plan = get_plan(22050, float(C1_HZ / 2 ** (2 / 72)), 512, 36, 6, 5, 1, 'cuda')
    act, tun = CnnStreamEngine(m, chunk=64).predict_audio(torch.from_numpy(y).cuda(), plan)
    f, _, _ = Q.compute_efficient_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    ref = oracle_patchwise(sd, np.transpose(f, (2, 1, 0)).astype(np.float32), True)
    assert abs((-0.5 + 0.01 * int(tun.item())) - Q.estimate_tuning(y, bins_per_octave=36)) < 1e-9
    assert np.abs(act.cpu().numpy() - ref).max() < TOL['fp16']


def test_frame_range_sharding_matches_unsharded_engine():
    """Two emulated ranks (sequentially on one GPU): halo'd frame ranges reproduce the unsharded result bit for bit."""
    from multipitch_architectures_b200.engine import CnnStreamEngine
    from multipitch_architectures_b200.parallel import predict_sharded
    m = build_model('drcnn_tiny', precision='fp16')
    m.load_state_dict(fill_state_dict(m.state_dict(), 24))
    m = m.cuda().eval()
    eng = CnnStreamEngine(m, chunk=48)
    h = torch.from_numpy(fake_hcqt(131, 5)).cuda()
    full = eng.predict_hcqt(h)
    parts = [predict_sharded(eng.predict_hcqt, h, world=3, rank=r, gather=False) for r in range(3)]
    assert torch.equal(torch.cat(parts, 0), full)


@pytest.mark.parametrize('name,N,chunk,prec', [('drcnn_tiny', 90, 64, 'fp16'), ('drcnn_tiny', 131, 40, 'bf16'), ('dcnn_tiny', 60, 64, 'fp16'),
                                               ('drcnn', 100, 37, 'fp16'), ('drcnn', 80, 80, 'bf16'),
                                               ('drcnn', 1292, 646, 'fp16')])      # BASELINE configs[0] at full size: one 30 s clip
def test_fused_deduplicated_schedule_equals_plain_schedule(name, N, chunk, prec):
    """Fused conv+pool+residual kernel with frame-shared interior rows (mpa_conv_tc_pool_f16) vs the plain per-patch
    sequence conv_tc -> pool_time_res.
    (a) same main loop (tile weights): the same MMA sequence per output row and the same 16-bit roundings -> bit for bit;
    (b) ring main loop (un-duplicated weight pieces, K walked group-major): de-duplicated vs fully per-patch -> bit for bit
        (this is what makes sharing rows across patches legal), and vs (a) within fp32 summation-order noise."""
    from multipitch_architectures_b200.engine import CnnStreamEngine
    m = build_model(name, precision=prec)
    m.load_state_dict(fill_state_dict(m.state_dict(), 27))
    m = m.cuda().eval()
    h = torch.from_numpy(fake_hcqt(N, 6)).cuda()
    plain = CnnStreamEngine(m, chunk=chunk, fused=False).predict_hcqt(h)
    tiles = CnnStreamEngine(m, chunk=chunk, fused=True, ring=False, split_head=False)
    assert tiles.fused
    fused = tiles.predict_hcqt(h)
    assert fused.shape == plain.shape == (N, 72)
    assert torch.equal(fused, plain), (fused - plain).abs().max().item()
    # default engine: the last block hands its output to the head phase-split and conv2 runs as a 3x1 convolution over 3*C0 channels:
    # another K order in conv2 (fp32 summation-order noise vs (a)), but still bit-identical between the de-duplicated and the
    # fully per-patch schedule
    split_eng = CnnStreamEngine(m, chunk=chunk, fused=True)
    assert split_eng.split == 3
    sp = split_eng.predict_hcqt(h)
    assert (sp - plain).abs().max().item() < (2e-3 if prec == 'fp16' else 2e-2)
    assert torch.equal(sp, CnnStreamEngine(m, chunk=chunk, fused=True, dedup=False).predict_hcqt(h))
    ring_eng = CnnStreamEngine(m, chunk=chunk, fused=True, ring=True)
    ring = ring_eng.predict_hcqt(h)
    ring_full = CnnStreamEngine(m, chunk=chunk, fused=True, ring=True, dedup=False).predict_hcqt(h)
    assert torch.equal(ring, ring_full), (ring - ring_full).abs().max().item()
    assert (ring - plain).abs().max().item() < (2e-3 if prec == 'fp16' else 2e-2)
    # a sub-range of frames (what a rank of the sharded run evaluates)
    part = ring_eng.predict_hcqt(h, lo=11, hi=N - 7)
    assert torch.equal(part, ring[11:N - 7])
