"""`-m gpu`: the split-precision tensor-core mode (precision='fp16x3', MPA_FMT_F16X3): every value is an fp16 (hi, lo) pair and every
product three tcgen05 passes (W_hi x_hi + W_lo x_hi + W_hi x_lo) into one fp32 accumulator.  It is the tensor-core mode that meets the
north star's 1e-3 on ANY weights: the tests below hold it to 1e-3 on the adversarial goldens of the REFERENCE classes (where
plain fp16 reads 1.5e-2) and require identical thresholded activity / P/R/F."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import host_oracle as HO
from oracle import nn_oracle as NO
from tests.refshapes import build_model
from tests.weights import fill_state_dict, synth_patches

pytestmark = pytest.mark.gpu

TOL_X3 = 1e-3          # the north star's bound.  Observed: 1e-5 .. 1e-4 on trained weights, 4.2e-4 worst on the adversarial DRCNN golden
                       # (plain fp16: 1.5e-2).  What is left is the tensor core's fp32 accumulator, which truncates when it aligns addends
                       # (~3e-5 relative over the 563 x 3 chained MMAs of a 40->40 15x15 layer), not the operands (~2^-21)


def rnd(*shape, seed=0, scale=1.0):
    return torch.from_numpy((np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32))


@pytest.fixture(scope='module')
def ops():
    from multipitch_architectures_b200 import ops as O
    return O


def test_x3_planes_roundtrip_pool_and_resampling(ops):
    fmt = ops.FMT_F16X3
    x = rnd(3, 40, 20, 216, seed=1)
    xc = ops.nchw_to_cp8(x.cuda(), fmt=fmt)
    assert xc.buf.shape == (3, 10, 22, 224, 8) and xc.buf.dtype == torch.float16          # 5 hi planes, then 5 lo planes
    back = ops.cp8_to_nchw(xc).cpu()
    assert (back - x).abs().max() <= 2.0 ** -20 * x.abs().max()
    hi = xc.buf[:, :5, 1:-1, 8:].float().cpu()
    assert torch.equal(hi.permute(0, 1, 4, 2, 3).reshape(3, 40, 20, 216)[:, :, :, :216], x.half().float())   # hi plane = fp16(x)
    assert float(xc.buf[:, :, 0].abs().max()) == 0 and float(xc.buf[:, :, :, :8].abs().max()) == 0           # borders stay zero
    r = rnd(3, 40, 20, 216, seed=2)
    rc = ops.nchw_to_cp8(r.cuda(), fmt=fmt)
    got = ops.cp8_to_nchw(ops.pool3_res_cp8(xc, rc)).cpu()
    ref = F.max_pool2d(back, (3, 1), (1, 1), (1, 0)) + ops.cp8_to_nchw(rc).cpu()
    assert (got - ref).abs().max() <= 2.0 ** -20 * ref.abs().max()
    got13 = ops.cp8_to_nchw(ops.pool_time_res_cp8(xc, 13)).cpu()
    assert torch.equal(got13, F.max_pool2d(back, (13, 1), (1, 1), (6, 0)))
    # MaxPool2d(2,2) and the bilinear up-sampler between two level geometries
    pooled = ops.maxpool2x2_cp8(xc, ops.CP8(3, 40, 10, 108, fmt=fmt, device='cuda'))
    assert torch.equal(ops.cp8_to_nchw(pooled).cpu(), F.max_pool2d(back, (2, 2)))
    cat = ops.CP8(3, 80, 21, 217, fmt=fmt, device='cuda')
    ops.upsample2x_cp8(pooled, cat.channels(40, 40))
    up = F.pad(F.interpolate(F.max_pool2d(back, (2, 2)), scale_factor=2, mode='bilinear', align_corners=True), (0, 1, 0, 1))
    got_up = ops.cp8_to_nchw(cat.channels(40, 40)).cpu()
    assert (got_up - up).abs().max() < 2e-6 * up.abs().max()
    assert float(ops.cp8_to_nchw(cat.channels(0, 40)).abs().max()) == 0                    # the skip half was not touched


@pytest.mark.parametrize('cfg', [
    (2, 8, 40, 6, 24, 1, 1), (2, 16, 40, 7, 24, 3, 3), (3, 24, 40, 9, 40, 3, 3), (3, 6, 40, 20, 216, 15, 15),
    (2, 40, 40, 75, 216, 15, 15), (2, 24, 24, 30, 216, 15, 15), (2, 64, 128, 9, 27, 5, 5), (2, 8, 16, 37, 108, 15, 15),
    (5, 32, 8, 18, 54, 9, 9), (2, 16, 128, 20, 216, 15, 15),
])
def test_conv_tc_x3_matches_fp64(ops, cfg):
    """Three-pass tcgen05 convolution vs an fp64 convolution of the UNROUNDED fp32 operands."""
    B, Cin, Cout, T, Fq, KH, KW = cfg
    fmt = ops.FMT_F16X3
    x, w, b = rnd(B, Cin, T, Fq, seed=4), rnd(Cout, Cin, KH, KW, seed=5, scale=(Cin * KH * KW) ** -0.5), rnd(Cout, seed=6, scale=0.1)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), padding=(KH // 2, KW // 2)), 0.3)
    xc = ops.nchw_to_cp8(x.cuda(), fmt=fmt)
    yc = ops.conv_tc(xc, ops.conv_tc_pack(w, 'cuda', fmt), b.cuda(), Cout, (KH, KW), ops.ACT_LRELU, 0.3)
    got = ops.cp8_to_nchw(yc).cpu().double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f'{cfg}: fp16x3 conv max|diff|/max = {err:.2e}')
    assert err < 6e-5          # operands carry ~2^-21; the rest is the truncating fp32 accumulation of up to 563 x 3 chained MMAs
    assert float(yc.buf[:, :, 0].abs().max()) == 0 and float(yc.buf[:, :, :, :8].abs().max()) == 0


def test_conv_tc_x3_tiny_and_huge_weights(ops):
    """The per-channel power-of-two weight scale keeps W_lo out of the fp16 subnormals (weights ~1e-5) and out of overflow (~1e3).
    (Activations are fp16 pairs without a scale: values below ~1e-3 keep an absolute error floor of 3e-8, values above 65504 overflow.)"""
    fmt = ops.FMT_F16X3
    b = torch.zeros(16)
    for wscale, xscale in ((1e-5, 1e3), (1e3, 1e-1)):
        x = rnd(2, 16, 12, 40, seed=1, scale=xscale)
        w = rnd(16, 16, 3, 3, seed=2, scale=wscale)
        w[3] *= 1e-3
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=(1, 1))
        yc = ops.conv_tc(ops.nchw_to_cp8(x.cuda(), fmt=fmt), ops.conv_tc_pack(w, 'cuda', fmt), b.cuda(), 16, (3, 3), ops.ACT_NONE, 0.0)
        got = ops.cp8_to_nchw(yc).cpu().double()
        for co in (0, 5):
            err = (got[:, co] - ref[:, co]).abs().max().item() / ref[:, co].abs().max().item()
            assert err < 2e-5, (wscale, co, err)


def _load(name, seed, **extra):
    m = build_model(name, **extra)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed, scheme='adversarial'))
    return m.cuda().eval()


@pytest.mark.parametrize('name', ['cnn_xs', 'drcnn_tiny', 'dcnn_tiny', 'drcnn', 'unet_tiny', 'unet_m', 'punet_tiny', 'saunet_tiny'])
def test_x3_models_meet_1e3_on_the_adversarial_reference_goldens(nn_golden, name):
    """Reference goldens with outputs spanning 0..1 (gain 2.4x the default, logits to +-18): plain fp16 is gated at 1.5e-2 there."""
    tag = f'{name}__eval'
    B, seed, _ = nn_golden[tag + '__meta']
    B, seed = int(B), int(seed)
    m = _load(name, seed, precision='fp16x3')
    with torch.no_grad():
        y = m(synth_patches(B, seed).cuda())
    if isinstance(y, tuple):
        assert np.abs(y[1].cpu().numpy() - nn_golden[tag + '__n']).max() < 1e-3
        y = y[0]
    ref = nn_golden[tag + '__y']
    err = np.abs(y.cpu().numpy() - ref).max()
    print(f'{tag}: fp16x3 max|diff| vs reference golden = {err:.2e} (outputs span {ref.min():.3f}..{ref.max():.3f})')
    assert err < TOL_X3
    assert np.array_equal(y.cpu().numpy() >= 0.4, ref >= 0.4) or np.abs(ref - 0.4)[(y.cpu().numpy() >= 0.4) != (ref >= 0.4)].max() < TOL_X3


def _oracle_patchwise(sd, hcqt, residual):
    C, N, Fq = hcqt.shape
    ip, _ = HO.pad_for_inference(hcqt, np.zeros((N, 72)))
    X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(N)]))
    with torch.no_grad():
        return NO.cnn_forward(sd, X, residual=residual).reshape(N, 72).numpy()


def _fake_hcqt(N, seed):
    rng = np.random.default_rng(seed)
    h = np.abs(rng.normal(0, 0.05, size=(6, N, 216))) * (1 + np.sin(np.arange(216) / 7.0) ** 2)[None, None, :]
    h *= rng.uniform(0.3, 2.0, size=(1, N, 1))
    return h.astype(np.float32)


@pytest.mark.parametrize('name,N,chunk', [('drcnn_tiny', 90, 64), ('dcnn_tiny', 131, 40), ('cnn_xs', 40, 592)])
def test_x3_stream_engine_matches_oracle_and_prf(name, N, chunk):
    """Fused, de-duplicated streaming schedule in split precision (incl. the zero-padded 20 -> 24 channel CNN:XS) vs the oracle's
    patch-wise evaluation; the thresholded matrix and P/R/F are compared unconditionally."""
    from multipitch_architectures_b200.engine import CnnStreamEngine
    m = build_model(name, precision='fp16x3')
    sd = fill_state_dict(m.state_dict(), 21)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    h = _fake_hcqt(N, 3)
    eng = CnnStreamEngine(m, chunk=chunk)
    got = eng.predict_hcqt(torch.from_numpy(h).cuda()).cpu().numpy()
    ref = _oracle_patchwise(sd, h, getattr(m, 'residual', False))
    err = np.abs(got - ref).max()
    print(f'{name}: streaming fp16x3 engine vs oracle max|diff| = {err:.2e}')
    assert got.shape == (N, 72) and err < TOL_X3
    flips = (got >= 0.4) != (ref >= 0.4)
    assert not flips.any(), f'{int(flips.sum())} thresholded cells differ (closest reference value to 0.4: {np.abs(ref - 0.4).min():.2e})'
    targ = np.random.default_rng(1).uniform(size=(N, 72)) < 0.3
    assert HO.eval_prf(targ, got, 0.4) == HO.eval_prf(targ, ref, 0.4)          # unconditional: identical counts, hence identical P/R/F
    # the un-deduplicated schedule (every row per patch) gives the same numbers: the same MMA sequence produces each row
    got2 = CnnStreamEngine(m, chunk=chunk, dedup=False).predict_hcqt(torch.from_numpy(h).cuda()).cpu().numpy()
    assert np.array_equal(got, got2)
    # and the per-batch module forward (materialised patches, separate pool kernels) agrees to rounding
    ip, _ = HO.pad_for_inference(h, np.zeros((N, 72)))
    X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(min(N, 16))]))
    with torch.no_grad():
        y = m(X.cuda()).reshape(-1, 72).cpu().numpy()
    assert np.abs(y - ref[:len(y)]).max() < TOL_X3
