"""`-m gpu`: kernel-level parity through the C ABI against torch fp32 references of the same op (the oracle's
building blocks), on seeded inputs.  Tolerances are stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    from multipitch_architectures_b200 import ops as o
    from multipitch_architectures_b200 import _lib
    assert _lib.lib().mpa_device_check() == 0, _lib.last_error()
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def test_layernorm_cf(ops):
    from oracle import nn_oracle as NO
    x = rnd(3, 6, 20, 216, seed=1).abs()
    w, b = 1 + 0.1 * rnd(6, 216, seed=2), 0.1 * rnd(6, 216, seed=3)
    got = ops.layernorm_cf(x.cuda(), w.cuda(), b.cuda()).cpu()
    assert (got - NO.layernorm_cf(x, w, b)).abs().max() < 2e-5
    got = ops.layernorm_cf(x.cuda(), w.cuda(), b.cuda(), gamma_log=10.0).cpu()
    assert (got - NO.layernorm_cf(torch.log(1 + 10 * x), w, b)).abs().max() < 5e-5


@pytest.mark.parametrize('cfg', [
    # B, Cin, Cout, H, W, KH, KW, sh, sw, ph, pw
    (2, 6, 20, 75, 216, 15, 15, 1, 1, 7, 7),
    (2, 8, 8, 37, 108, 15, 15, 1, 1, 7, 7),
    (3, 16, 32, 18, 54, 9, 9, 1, 1, 4, 4),
    (2, 40, 40, 75, 216, 3, 3, 1, 3, 1, 0),
    (2, 40, 30, 75, 72, 75, 1, 1, 1, 0, 0),
    (2, 20, 10, 100, 72, 75, 1, 1, 1, 0, 0),
    (4, 30, 10, 1, 72, 1, 1, 1, 1, 0, 0),
    (4, 10, 1, 1, 72, 1, 1, 1, 1, 0, 0),
    (2, 64, 32, 4, 13, 2, 5, 1, 1, 0, 0),
    (2, 32, 24, 2, 3, 2, 3, 1, 1, 0, 0),
    (1, 3, 17, 5, 7, 3, 3, 1, 1, 1, 1),
])
def test_conv2d_f32(ops, cfg):
    B, Cin, Cout, H, W, KH, KW, sh, sw, ph, pw = cfg
    x, w, b = rnd(B, Cin, H, W, seed=4), rnd(Cout, Cin, KH, KW, seed=5, scale=(Cin * KH * KW) ** -0.5), rnd(Cout, seed=6, scale=0.1)
    ref = F.leaky_relu(F.conv2d(x, w, b, stride=(sh, sw), padding=(ph, pw)), 0.3)
    got = ops.conv2d(x.cuda(), ops.pack_conv_weight(w.cuda()), b.cuda(), Cout, (KH, KW), (sh, sw), (ph, pw), ops.ACT_LRELU, 0.3).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max() < 2e-5 * max(1.0, ref.abs().max().item())      # fp32 FMA vs fp32 reference


def test_conv2d_f32_concat_scale_shift_sigmoid(ops):
    x1, x2 = rnd(2, 5, 9, 27, seed=1), rnd(2, 7, 9, 27, seed=2)
    w, b = rnd(9, 12, 5, 5, seed=3, scale=0.06), rnd(9, seed=4, scale=0.1)
    sc, sf = 1 + 0.2 * rnd(9, seed=5), 0.1 * rnd(9, seed=6)
    ref = torch.relu(F.conv2d(torch.cat([x1, x2], 1), w, b, padding=2) * sc[None, :, None, None] + sf[None, :, None, None])
    got = ops.conv2d(x1.cuda(), ops.pack_conv_weight(w.cuda()), b.cuda(), 9, (5, 5), (1, 1), (2, 2), ops.ACT_RELU, 0.0,
                     scale=sc.cuda(), shift=sf.cuda(), x2=x2.cuda()).cpu()
    assert (got - ref).abs().max() < 2e-5
    ref = torch.sigmoid(F.conv2d(x1, w[:, :5], b, padding=2))
    got = ops.conv2d(x1.cuda(), ops.pack_conv_weight(w[:, :5].contiguous().cuda()), b.cuda(), 9, (5, 5), (1, 1), (2, 2), ops.ACT_SIGMOID).cpu()
    assert (got - ref).abs().max() < 1e-6


def test_pools_upsample_bn_bce(ops):
    from oracle import nn_oracle as NO
    x = rnd(2, 5, 75, 72, seed=1)
    r = rnd(2, 5, 75, 72, seed=2)
    assert torch.equal(ops.maxpool_time(x.cuda(), 13).cpu(), NO.maxpool_t(x, 13))
    assert torch.equal(ops.maxpool_time(x.cuda(), 3, res=r.cuda()).cpu(), NO.maxpool_t(x, 3) + r)
    y = rnd(2, 3, 75, 216, seed=3)
    assert torch.equal(ops.maxpool2d(y.cuda(), (2, 2), (2, 2)).cpu(), F.max_pool2d(y, (2, 2)))
    z = rnd(2, 4, 3, 9, seed=4)
    assert torch.equal(ops.maxpool2d(z.cuda(), (2, 5), (1, 2)).cpu(), F.max_pool2d(z, (2, 5), (1, 2)))
    for (hl, wl, hs, ws) in ((4, 13, 9, 27), (9, 27, 18, 54), (18, 54, 37, 108), (37, 108, 75, 216)):
        low, skip = rnd(2, 3, hl, wl, seed=5), rnd(2, 2, hs, ws, seed=6)
        ref = NO.upconcat(low, skip)
        ref2 = torch.cat([skip, F.pad(F.interpolate(low, scale_factor=2, mode='bilinear', align_corners=True),
                                      [0, ws - 2 * wl, 0, hs - 2 * hl])], 1)
        assert (ref - ref2).abs().max() < 1e-6
        got = ops.upsample2x_concat(low.cuda(), skip.cuda()).cpu()
        assert (got - ref2).abs().max() < 2e-6
    a = rnd(3, 7, 9, 27, seed=7) * 2 + 0.5
    st = ops.bn_stats(a.cuda())
    assert (st[:7].cpu() - a.mean((0, 2, 3))).abs().max() < 1e-5
    assert (st[7:].cpu() - a.var((0, 2, 3), unbiased=False)).abs().max() < 1e-4
    w, b = 1 + 0.1 * rnd(7, seed=8), rnd(7, seed=9)
    ref = torch.relu(F.batch_norm(a, None, None, w, b, training=True))
    assert (ops.bn_apply(a.cuda(), st, w.cuda(), b.cuda(), act=ops.ACT_RELU).cpu() - ref).abs().max() < 2e-5
    yp = torch.tensor([0.0, 1.0, 0.5, 0.2, 0.999999, 1e-30]).repeat(12)
    yt = torch.tensor([1.0, 0.0, 1.0, 0.0, 1.0, 0.0]).repeat(12)
    loss, grad = ops.bce_fwd_bwd(yp.cuda(), yt.cuda())
    ypr = yp.clone().requires_grad_(True)
    lref = torch.nn.BCELoss()(ypr, yt)
    lref.backward()
    assert abs(loss.item() - lref.item()) < 1e-4 * lref.item()
    assert (grad.cpu() - ypr.grad).abs().max() <= 1e-5 * ypr.grad.abs().max()


FMTS = [('fp16', torch.float16), ('bf16', torch.bfloat16)]


@pytest.mark.parametrize('prec,dt', FMTS)
def test_cp8_roundtrip_and_pool(ops, prec, dt):
    fmt = ops.fmt_of(prec)
    x = rnd(3, 40, 75, 216, seed=1)
    xc = ops.nchw_to_cp8(x.cuda(), fmt=fmt)
    assert xc.buf.shape == (3, 5, 77, 224, 8) and xc.buf.dtype == dt
    back = ops.cp8_to_nchw(xc).cpu()
    xb = x.to(dt).float()
    assert torch.equal(back, xb)
    # borders stay zero
    assert float(xc.buf[:, :, 0].abs().max()) == 0 and float(xc.buf[:, :, :, :8].abs().max()) == 0
    r = rnd(3, 40, 75, 216, seed=2)
    rc = ops.nchw_to_cp8(r.cuda(), fmt=fmt)
    got = ops.cp8_to_nchw(ops.pool3_res_cp8(xc, rc)).cpu()
    ref = (F.max_pool2d(xb, (3, 1), (1, 1), (1, 0)) + r.to(dt).float()).to(dt).float()
    assert torch.equal(got, ref)


@pytest.mark.parametrize('prec,dt', FMTS)
@pytest.mark.parametrize('cfg', [
    (2, 8, 40, 6, 24, 1, 1), (2, 16, 40, 7, 24, 3, 3), (3, 24, 40, 9, 40, 3, 3), (3, 6, 40, 20, 216, 15, 15),
    (3, 40, 40, 75, 216, 15, 15), (2, 20, 20, 75, 216, 15, 15), (2, 64, 128, 9, 27, 5, 5), (1, 40, 40, 75, 216, 15, 15),
    (2, 8, 16, 37, 108, 15, 15), (5, 32, 8, 18, 54, 9, 9),
    # KH x 1 filters with many input chunks: row-merged operand rows (R = 256 // pitch rows per MMA) and chunk-group activation stages
    (2, 384, 104, 75, 72, 3, 1), (3, 48, 128, 12, 40, 3, 1), (2, 136, 80, 10, 72, 1, 1), (2, 392, 128, 14, 72, 3, 1), (1, 16, 128, 7, 100, 3, 1),
    # ... and with J > 1 row blocks (Cout <= 64): merged rows J apart, R * J output rows per unit (the DRCNN head's 120 -> 40 conv2)
    (2, 120, 40, 75, 72, 3, 1), (3, 24, 16, 20, 40, 3, 1), (2, 64, 56, 11, 72, 3, 1), (5, 8, 32, 9, 24, 1, 1),
])
def test_conv_tc(ops, cfg, prec, dt):
    """tcgen05 path vs an fp64 convolution of the SAME 16-bit-rounded operands: only fp32 accumulation order and the
    final rounding of the output to the storage format (half an ulp) may differ."""
    B, Cin, Cout, T, Fq, KH, KW = cfg
    fmt = ops.fmt_of(prec)
    x, w, b = rnd(B, Cin, T, Fq, seed=4), rnd(Cout, Cin, KH, KW, seed=5, scale=(Cin * KH * KW) ** -0.5), rnd(Cout, seed=6, scale=0.1)
    xr, wr = x.to(dt).double(), w.to(dt).double()
    ref = F.leaky_relu(F.conv2d(xr, wr, b.double(), padding=(KH // 2, KW // 2)), 0.3).float()
    xc = ops.nchw_to_cp8(x.cuda(), fmt=fmt)
    yc = ops.conv_tc(xc, ops.conv_tc_pack(w, 'cuda', fmt), b.cuda(), Cout, (KH, KW), ops.ACT_LRELU, 0.3)
    got = ops.cp8_to_nchw(yc).cpu()
    ulp = 2.0 ** -8 if prec == 'bf16' else 2.0 ** -11
    assert (got - ref).abs().max() < ulp * ref.abs().max().item() + 1e-4
    assert float(yc.buf[:, :, 0].abs().max()) == 0 and float(yc.buf[:, :, :, :8].abs().max()) == 0     # borders untouched


@pytest.mark.parametrize('prec,dt', FMTS)
def test_conv_tc_subsampled_output_pool13_and_head_tail(ops, prec, dt):
    """conv2 of the head: 3x3 stride (1,3) pad (1,0) == stride-1 'same' conv sampled at columns 1,4,7,... written as
    compact 16-bit planes; then maxpool(13,1) on those planes and the fused conv3/conv4 tail, each against torch."""
    fmt = ops.fmt_of(prec)
    B, C0, C1, C2, C3, T, Fq = 3, 40, 40, 30, 10, 75, 216
    x, w2, b2 = rnd(B, C0, T, Fq, seed=1), rnd(C1, C0, 3, 3, seed=2, scale=(C0 * 9) ** -0.5), rnd(C1, seed=3, scale=0.1)
    xr, wr = x.to(dt).double(), w2.to(dt).double()
    ref2 = F.leaky_relu(F.conv2d(xr, wr, b2.double(), stride=(1, 3), padding=(1, 0)), 0.3).float()
    xc = ops.nchw_to_cp8(x.cuda(), fmt=fmt)
    y2 = ops.conv_tc(xc, ops.conv_tc_pack(w2, 'cuda', fmt), b2.cuda(), C1, (3, 3), ops.ACT_LRELU, 0.3, subsample=(3, 1))
    assert tuple(y2.buf.shape)[1:] == (5, T, 72, 8) and y2.compact
    got2 = ops.cp8_to_nchw(y2).cpu()
    ulp = 2.0 ** -8 if prec == 'bf16' else 2.0 ** -11
    assert (got2 - ref2).abs().max() < ulp * ref2.abs().max().item() + 1e-4
    p13 = ops.pool_time_res_cp8(y2, 13)
    assert torch.equal(ops.cp8_to_nchw(p13).cpu(), F.max_pool2d(got2, (13, 1), (1, 1), (6, 0)))
    w3, b3 = rnd(C2, C1, T, 1, seed=4, scale=(C1 * T) ** -0.5), rnd(C2, seed=5, scale=0.1)
    w40, b40 = rnd(C3, C2, 1, 1, seed=6, scale=C2 ** -0.5), rnd(C3, seed=7, scale=0.1)
    w43, b43 = rnd(1, C3, 1, 1, seed=8, scale=C3 ** -0.5), rnd(1, seed=9, scale=0.1)
    y = ops.cp8_to_nchw(p13).cpu()
    ref = torch.sigmoid(F.conv2d(F.leaky_relu(F.conv2d(F.leaky_relu(F.conv2d(y, w3, b3), 0.3), w40, b40), 0.3), w43, b43))
    got = ops.head_tail(p13, w3.cuda(), b3.cuda(), w40.cuda(), b40.cuda(), w43.cuda(), b43.cuda(), 0.3)
    assert (got.cpu() - ref.reshape(B, 72)).abs().max() < 1e-5


@pytest.mark.parametrize('prec,dt', FMTS)
@pytest.mark.parametrize('B,C1,C2,C3,Fo', [(3, 30, 10, 1, 72), (2, 20, 10, 1, 72), (5, 40, 16, 12, 24), (1, 64, 3, 2, 9)])
def test_fused_pool13_conv3_tail(ops, prec, dt, B, C1, C2, C3, Fo):
    """head.cu head_pool_conv3_tail_kernel: MaxPool((13,1)) + conv3 (75x1) + LReLU + conv4.0 + LReLU + conv4.3 + sigmoid in one launch on
    compact 16-bit planes, against torch on the same 16-bit values (basic_cnns.py:180-195) and against the three-launch sequence."""
    from multipitch_architectures_b200.libdl.nn_models import _exec
    import torch.nn as nn
    fmt, T = ops.fmt_of(prec), 75
    y = (rnd(B, C1, T, Fo, seed=1) * 2).round() / 2 + 0.1 * rnd(B, C1, T, Fo, seed=2)
    y = y.to(dt).float()
    torch.manual_seed(3)
    conv3, c40, c43 = nn.Conv2d(C1, C2, (75, 1)), nn.Conv2d(C2, C3, (1, 1)), nn.Conv2d(C3, 1, (1, 1))
    with torch.no_grad():
        ref = torch.sigmoid(c43(F.leaky_relu(c40(F.leaky_relu(conv3(F.max_pool2d(y, (13, 1), (1, 1), (6, 0))), 0.3)), 0.3))).reshape(B, Fo)
    C1p = (C1 + 7) // 8 * 8
    yc = ops.compact_cp8(B, C1p, T, Fo, 'cuda', fmt)
    yc.buf.zero_()
    ops.nchw_to_cp8(y.cuda(), out=yc.channels(0, C1))
    conv3, c40, c43 = conv3.cuda(), c40.cuda(), c43.cuda()
    w3p = ops.pack_conv3_rows(conv3.weight, C1p)
    got = ops.head_pool_conv3_tail(yc, w3p, conv3.bias, c40.weight, c40.bias, c43.weight, c43.bias, 0.3).cpu()
    assert (got - ref).abs().max() < 2e-5
    # the sequence it replaces: pool kernel, tcgen05 conv3 (16-bit weights and hidden activations), tail kernel
    cache = _exec.ParamCache()
    pc = ops.pool_time_res_cp8(yc, 13)
    C2p = (C2 + 7) // 8 * 8
    hc = ops.compact_cp8(B, C2p, 1, Fo, 'cuda', fmt)
    for wp, b, c0, c in _exec._folded_tc(cache, 'conv3', conv3, None, fmt, 'cuda', cin_pad=C1p, J=1):
        ops.conv_tc(pc, wp, b, c, (75, 1), ops.ACT_LRELU, 0.3, out=hc.channels(c0, c), J=1, rows=(37, 1))
    seq = ops.head_tail2(hc, _exec._pad_cols(c40.weight.reshape(C3, -1), C2p), c40.bias, c43.weight, c43.bias, 0.3).cpu().reshape(B, Fo)
    assert (got - seq).abs().max() < (2e-2 if prec == 'bf16' else 3e-3)


@pytest.mark.parametrize('C1,C2,C3,T', [(40, 30, 10, 75), (100, 80, 50, 75), (180, 150, 100, 75), (20, 10, 1, 90)])
def test_conv3_on_tensor_cores_and_general_tail(ops, C1, C2, C3, T):
    """conv3 (75x1 VALID) as the row-windowed 'same' convolution on compact planes (KW == 1, pitch 72 -> MMA N 80), with
    channel padding to multiples of 8 and co-blocks of 128, then the general conv4 tail; against torch fp64/fp32."""
    from multipitch_architectures_b200.libdl.nn_models import _exec
    import torch.nn as nn
    B, Fo = 3, 72
    y = rnd(B, C1, T, Fo, seed=1)
    conv3, c40, c43 = nn.Conv2d(C1, C2, (75, 1)), nn.Conv2d(C2, C3, (1, 1)), nn.Conv2d(C3, 1, (1, 1))
    yr = y.half().double()
    h_ref = F.leaky_relu(F.conv2d(yr, conv3.weight.detach().half().double(), conv3.bias.detach().double()), 0.3).float()
    C1p, C2p = (C1 + 7) // 8 * 8, (C2 + 7) // 8 * 8
    yc = ops.compact_cp8(B, C1p, T, Fo, 'cuda', ops.FMT_F16)
    yc.buf.zero_()
    ops.nchw_to_cp8(y.cuda(), out=yc.channels(0, C1))
    conv3, c40, c43 = conv3.cuda(), c40.cuda(), c43.cuda()
    cache = _exec.ParamCache()
    n_rows = T - 74
    hc = ops.compact_cp8(B, C2p, n_rows, Fo, 'cuda', ops.FMT_F16)
    for wp, b, c0, c in _exec._folded_tc(cache, 'conv3', conv3, None, ops.FMT_F16, 'cuda', cin_pad=C1p, J=1):
        ops.conv_tc(yc, wp, b, c, (75, 1), ops.ACT_LRELU, 0.3, out=hc.channels(c0, c), J=1, rows=(37, n_rows))
    h = ops.cp8_to_nchw(hc).cpu()
    assert tuple(h.shape) == (B, C2p, n_rows, Fo)
    assert (h[:, :C2] - h_ref).abs().max() < 2.0 ** -11 * h_ref.abs().max().item() + 1e-4
    assert float(h[:, C2:].abs().max() if C2p > C2 else 0.0) == 0.0
    o = ops.head_tail2(hc, _exec._pad_cols(c40.weight.reshape(C3, -1), C2p), c40.bias, c43.weight, c43.bias, 0.3).cpu()
    hq = h[:, :C2]
    ref = torch.sigmoid(F.conv2d(F.leaky_relu(F.conv2d(hq, c40.weight.cpu(), c40.bias.cpu()), 0.3), c43.weight.cpu(), c43.bias.cpu()))
    assert (o - ref.reshape(B, n_rows, Fo)).abs().max() < 1e-5


def test_encoder_layer(ops):
    from oracle import nn_oracle as NO
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict
    m = build_model('saunet_tiny')
    sd = fill_state_dict(m.state_dict(), 3)
    m.load_state_dict(sd)
    m = m.cuda()
    x5 = rnd(5, 32, 4, 13, seed=9)
    with torch.no_grad():
        got = m.attention1.run(x5.cuda()).cpu()
        ref = NO.encoder_layer(x5, sd, 'attention1', 8, True)
    assert (got - ref).abs().max() < 5e-5


@pytest.mark.parametrize('B,E,mlp', [(330, 128, 512), (400, 64, 256), (12, 128, 512), (50, 128, 8192), (50, 64, 256)])
def test_encoder_layer_tensor_core_path_with_layernorm_epilogue(ops, B, E, mlp):
    """transformer_enc_layer (unet_cnns.py:131-159) on the tensor-core inference path: from 64 token tiles on (B * 52 >= 16384 tokens) the second
    add & LayerNorm runs in the epilogue of the MLP's second product (3 launches per layer); smaller batches keep the split-K product + add&LN
    kernel.  Both against the fp32 sequence of the same layer."""
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.nn_models.unet_cnns import transformer_enc_layer
    torch.manual_seed(5)
    layer = transformer_enc_layer(embed_dim=E, num_heads=8, mlp_dim=mlp, p_dropout=0.2, pos_encoding='sinusoidal').cuda().eval()
    x = rnd(B, E, 4, 13, seed=31).cuda()
    with torch.no_grad():
        ref = layer.run(x)
        n0 = _lib.launch_count()
        got = layer.run(x, ops.FMT_F16)
        n1 = _lib.launch_count()
        got2 = layer.run(x, ops.FMT_F16)
        launches = _lib.launch_count() - n1
    assert (got - got2).abs().max().item() < 1e-4          # (split-K slices meet in fp32 atomics: the order varies from run to run)
    err = (got - ref).abs().max().item()
    print(f'B={B} E={E}: fp16 tensor-core layer vs fp32 max|d| {err:.2e}; launches per forward {launches} (first call incl. operand preparation {n1 - n0})')
    assert err < 4e-3 * max(1.0, ref.abs().max().item())
    if B <= 64 and E >= 64:                      # the fused attention half (<= 64 items per position) + two MLP products
        assert launches <= 3


@pytest.mark.parametrize('B,Cin,H,W,Cout', [(25, 80, 75, 72, 50), (256, 20, 75, 72, 10), (3, 7, 75, 72, 5), (2, 100, 9, 13, 80),
                                            # thin layers (Cout <= 4, W % 4 == 0): the HBM-bound matrix-vector kernels
                                            (256, 10, 75, 72, 1), (5, 10, 75, 72, 1), (7, 6, 20, 216, 3), (40, 3, 11, 8, 4), (4, 1, 1, 72, 1),
                                            (3, 5, 9, 70, 2)])
def test_conv_rows_gemm_kernels_match_torch(B, Cin, H, W, Cout):
    """conv3 (75x1 VALID, one output row) forward / data gradient / weight gradient as GEMMs vs torch fp32 conv2d + autograd."""
    import torch.nn.functional as F
    from multipitch_architectures_b200 import _lib
    g = torch.Generator().manual_seed(B * 7 + Cin)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, H, 1, generator=g) / (Cin * H) ** 0.5
    b = torch.randn(Cout, generator=g)
    gy = torch.randn(B, Cout, 1, W, generator=g)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    pre = F.conv2d(xr, wr, b)
    pre.backward(gy)
    ref = torch.where(pre >= 0, pre, 0.3 * pre).detach()
    xc, wc, bc, gc = x.cuda(), w.cuda(), b.cuda(), gy.cuda()
    out = torch.empty(B, Cout, 1, W, device='cuda')
    ws_bytes = _lib.lib().mpa_conv_rows_fwd_workspace(B, Cin, H, W, Cout)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device='cuda')
    _lib.call('conv_rows_fwd_f32', xc, wc, bc, out, B, Cin, H, W, Cout, 1, 0.3, ws, _lib.usize(ws_bytes), _lib.stream_ptr())
    out2 = torch.empty_like(out)
    _lib.call('conv_rows_fwd_f32', xc, wc, bc, out2, B, Cin, H, W, Cout, 1, 0.3, ws, _lib.usize(ws_bytes), _lib.stream_ptr())
    assert torch.equal(out, out2)          # split-K forward: slices added in order, bit-reproducible
    assert (out.cpu() - ref).abs().max() < 2e-5
    gi = torch.empty_like(xc)
    _lib.call('conv_rows_dgrad_f32', gc, wc, gi, B, Cin, H, W, Cout, _lib.stream_ptr())
    assert (gi.cpu() - xr.grad).abs().max() < 2e-5
    gw = torch.full_like(wc, 7.0)           # must be overwritten, not accumulated
    _lib.call('conv_rows_wgrad_f32', xc, gc, gw, B, Cin, H, W, Cout, _lib.stream_ptr())
    assert (gw.cpu() - wr.grad).abs().max() < 2e-4 * max(1.0, wr.grad.abs().max().item())


@pytest.mark.parametrize('B,E,S,H,prec', [(50, 128, 52, 8, 'fp16'), (25, 128, 52, 8, 'bf16'), (64, 64, 12, 8, 'fp16'), (3, 128, 5, 4, 'fp16'), (1, 96, 7, 6, 'fp16')])
def test_fused_attention_block_tc_matches_the_separate_fp32_stages(ops, B, E, S, H, prec):
    """mpa_enc_attn_block_tc (gather + PE, q/k/v and out-projection on tcgen05, batch-axis softmax, residual, LayerNorm1 in ONE launch)
    vs the five separate fp32 launches it replaces; the difference is the 16-bit rounding of the two GEMMs' operands."""
    from multipitch_architectures_b200 import _lib
    fmt = ops.fmt_of(prec)
    x, pe = rnd(B, E, S, seed=1).cuda(), rnd(S, E, seed=2, scale=0.5).cuda()
    w_qkv, b_qkv = rnd(3 * E, E, seed=3, scale=E ** -0.5).cuda(), rnd(3 * E, seed=4, scale=0.1).cuda()
    w_proj, b_proj = rnd(E, E, seed=5, scale=E ** -0.5).cuda(), rnd(E, seed=6, scale=0.1).cuda()
    ln_w, ln_b = (1 + 0.1 * rnd(E, seed=7)).cuda(), (0.1 * rnd(E, seed=8)).cuda()
    M, sp = B * S, _lib.stream_ptr
    f32 = lambda *sh: torch.empty(*sh, dtype=torch.float32, device='cuda')
    tok, qkv, att, proj, ref = f32(M, E), f32(M, 3 * E), f32(M, E), f32(M, E), f32(M, E)
    _lib.call('enc_gather_f32', x, pe, tok, B, E, S, sp())
    _lib.call('gemm_nt_f32', tok, w_qkv, b_qkv, qkv, M, 3 * E, E, 0, sp())
    _lib.call('batch_axis_attention_f32', qkv, att, B, S, E, H, sp())
    _lib.call('gemm_nt_f32', att, w_proj, b_proj, proj, M, E, E, 0, sp())
    _lib.call('add_layernorm_tok_f32', tok, proj, ln_w, ln_b, ref, None, _lib.i64(M), E, S, 1e-5, sp())
    h1 = f32(M, E)
    h1c = torch.zeros(_lib.lib().mpa_gemm_tc_chunked_bytes(M, E, 256), dtype=torch.uint8, device='cuda')
    _lib.call('enc_attn_block_tc', x, pe, ops.gemm_tc_chunks(w_qkv, 128, fmt), b_qkv, ops.gemm_tc_chunks(w_proj, 128, fmt), b_proj, ln_w, ln_b,
              h1, h1c, B, E, S, H, 1e-5, fmt, sp())
    err = (h1 - ref).abs().max().item()
    print(f'fused attention block B={B} E={E} {prec}: max|diff| vs fp32 stages = {err:.2e}')
    assert err < (4e-2 if prec == 'bf16' else 6e-3)
    # the 16-bit operand copy of h1 equals what the converter makes of it
    assert torch.equal(h1c, ops.gemm_tc_chunks(h1, 256, fmt)[:h1c.numel()]) or (E % 64 != 0)
