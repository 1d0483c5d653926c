"""`-m gpu`: the CP8-resident training stages of the U-Net family (train_unet_cp8.cu; reference modules libdl/nn_models/unet_cnns.py:30-104):
each kernel against torch (autograd) on the same bf16-rounded values, and the whole CP8-resident tape against the fp32-NCHW tape with
converters (identical convolutions, element-wise stages in fp32) and against the fp32 path."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.weights import fill_state_dict, synth_patches, synth_targets

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def bf(x):
    return x.to(torch.bfloat16).float()


def _planes(x, fmt=None):
    from multipitch_architectures_b200 import ops
    return ops.nchw_to_cp8(x.cuda().contiguous(), fmt=ops.FMT_BF16 if fmt is None else fmt)


@pytest.mark.parametrize('shape', [(3, 16, 9, 27), (2, 8, 75, 216), (25, 32, 4, 13)])
def test_bn_stats_apply_and_running_statistics(shape):
    from multipitch_architectures_b200 import ops
    B, C, T, Fq = shape
    x = bf(rnd(*shape, seed=1) * 1.7 + 3.0 * rnd(1, C, 1, 1, seed=2))          # |mean| > std in some channels
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.3 * rnd(C, seed=3))
        bn.bias.copy_(0.2 * rnd(C, seed=4))
    ref_bn = torch.nn.BatchNorm2d(C)
    ref_bn.load_state_dict(bn.state_dict())
    ref = torch.relu(ref_bn.train()(x))
    bn = bn.cuda().train()
    xc = _planes(x)
    stats = ops.bn_stats_cp8(xc, bn)
    assert (stats[:C].cpu() - x.mean((0, 2, 3))).abs().max() < 2e-5 * (1 + x.abs().max())
    assert ((stats[C:].cpu() - x.var((0, 2, 3), unbiased=False)).abs() / x.var((0, 2, 3), unbiased=False)).max() < 1e-4
    assert (bn.running_mean.cpu() - ref_bn.running_mean).abs().max() < 1e-5 and (bn.running_var.cpu() - ref_bn.running_var).abs().max() < 1e-4
    assert int(bn.num_batches_tracked) == 1
    out = ops.bn_relu_apply_cp8(xc, stats, bn, xc.like())
    got = ops.cp8_to_nchw(out).cpu()
    assert (got - bf(ref.detach())).abs().max() <= 2 ** -7 * ref.abs().max()                      # one bf16 rounding
    assert (out.buf[:, :, 0].float().abs().max() == 0) and (out.buf[:, :, :, :8].float().abs().max() == 0)   # borders stay zero


@pytest.mark.parametrize('shape', [(3, 16, 9, 27), (2, 8, 37, 108)])
def test_bn_relu_backward_on_planes(shape):
    from multipitch_architectures_b200 import ops
    B, C, T, Fq = shape
    x = bf(rnd(*shape, seed=5) + 0.5).requires_grad_(True)
    w, b = (1 + 0.2 * rnd(C, seed=6)).requires_grad_(True), (0.1 * rnd(C, seed=7)).requires_grad_(True)
    y = torch.relu(F.batch_norm(x, None, None, w, b, training=True, eps=1e-5))
    g = bf(rnd(*shape, seed=8))
    y.backward(g)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(w)
        bn.bias.copy_(b)
    xc, gc = _planes(x.detach()), _planes(g)
    stats = ops.bn_stats_cp8(xc, bn)
    dy = xc.like()
    gw, gb, gcb = torch.empty(C).cuda(), torch.empty(C).cuda(), torch.full((C,), 7.0).cuda()
    ops.bn_relu_bwd_cp8(gc, xc, stats, bn, dy, gw, gb, gcb)
    ref = x.grad
    assert (ops.cp8_to_nchw(dy).cpu() - ref).abs().max() < 2 ** -7 * ref.abs().max() + 1e-6
    assert (gw.cpu() - w.grad).abs().max() < 1e-4 * (1 + w.grad.abs().max()) and (gb.cpu() - b.grad).abs().max() < 1e-4 * (1 + b.grad.abs().max())
    assert gcb.abs().max().item() < 1e-3 * g.abs().sum().item() / C           # sum of dy: zero up to rounding (overwritten, not accumulated)
    # a gradient that arrives as a channel view of a wider (concat) buffer
    wide = _planes(torch.cat([g, rnd(B, 8, T, Fq, seed=9)], 1))
    dy2 = xc.like()
    ops.bn_relu_bwd_cp8(wide.channels(0, C), xc, stats, bn, dy2, gw, gb, None)
    assert torch.equal(dy2.buf, dy.buf)


@pytest.mark.parametrize('shape', [(2, 8, 9, 27), (2, 16, 75, 216), (3, 8, 18, 54)])
def test_maxpool2x2_backward_with_skip_gradient(shape):
    from multipitch_architectures_b200 import ops
    B, C, T, Fq = shape
    x = bf(rnd(*shape, seed=10)).requires_grad_(True)
    x.data[0, 0, 0:2, 0:2] = 1.25                                               # a tie: the first element of the window takes the gradient
    y = F.max_pool2d(x, 2)
    g = bf(rnd(*y.shape, seed=11))
    y.backward(g)
    add = bf(rnd(*shape, seed=12))
    xc = _planes(x.detach())
    pooled = ops.CP8(B, C, T // 2, Fq // 2, device='cuda', fmt=ops.FMT_BF16)
    ops.maxpool2x2_cp8(xc, pooled)
    assert torch.equal(ops.cp8_to_nchw(pooled).cpu(), y.detach())
    out = ops.maxpool2x2_bwd_cp8(xc, _planes(g), _planes(add), xc.like())
    assert torch.equal(ops.cp8_to_nchw(out).cpu(), bf(x.grad + add))
    out = ops.maxpool2x2_bwd_cp8(xc, _planes(g), None, xc.like())
    assert torch.equal(ops.cp8_to_nchw(out).cpu(), x.grad)


@pytest.mark.parametrize('lo,sk', [((2, 8, 4, 13), (9, 27)), ((2, 16, 9, 27), (18, 54)), ((1, 8, 37, 108), (75, 216))])
def test_upsample2x_backward_is_the_adjoint_of_the_forward(lo, sk):
    from multipitch_architectures_b200 import ops
    B, C, Tl, Fl = lo
    Ts, Fs = sk
    low = bf(rnd(*lo, seed=13)).requires_grad_(True)
    up = F.interpolate(low, scale_factor=2, mode='bilinear', align_corners=True)
    dT, dF = Ts - up.shape[2], Fs - up.shape[3]
    up = F.pad(up, [dF // 2, dF - dF // 2, dT // 2, dT - dT // 2])
    g = bf(rnd(B, C, Ts, Fs, seed=14))
    up.backward(g)
    cat_g = _planes(torch.cat([rnd(B, 8, Ts, Fs, seed=15), g], 1))                # the up-sampled part follows 8 skip channels
    g_low = ops.upsample2x_bwd_cp8(cat_g.channels(8, C), ops.CP8(B, C, Tl, Fl, device='cuda', fmt=ops.FMT_BF16))
    ref = low.grad
    assert (ops.cp8_to_nchw(g_low).cpu() - ref).abs().max() < 2 ** -7 * ref.abs().max() + 1e-6
    # forward kernel against torch as well (same geometry)
    cat = _planes(torch.zeros(B, 8 + C, Ts, Fs))
    ops.upsample2x_cp8(_planes(low.detach()), cat.channels(8, C))
    assert (ops.cp8_to_nchw(cat)[:, 8:].cpu() - up.detach()).abs().max() < 2 ** -7 * up.abs().max() + 1e-6


def _model(kind, precision):
    from multipitch_architectures_b200.libdl import nn_models as M
    kw = dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=8, precision=precision)
    if kind == 'unet':
        return M.simple_u_net_largekernels(**kw)
    return M.simple_u_net_doubleselfattn(**kw, embed_dim=64, num_heads=8, mlp_dim=128, pos_encoding='sinusoidal')


def _run(kind, precision, cp8_tape, seed=43, B=3):
    from multipitch_architectures_b200 import training_unet
    m = _model(kind, precision)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed, scheme='torch_default'))
    for mod in m.modules():
        if hasattr(mod, 'p_dropout'):
            mod.p_dropout = 0.0
    m = m.cuda().train()
    x, t = synth_patches(B, seed).cuda(), synth_targets(B, seed).cuda()
    old = training_unet.CP8_TAPE
    training_unet.CP8_TAPE = cp8_tape
    try:
        assert training_unet.cp8_tape_eligible(m, x) == (cp8_tape and precision == 'bf16')
        y = m(x)
        loss = torch.nn.BCELoss(reduction='mean')(y, t)
        loss.backward()
    finally:
        training_unet.CP8_TAPE = old
    stats = {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items() if 'running_' in k or 'num_batches' in k}
    return loss.item(), y.detach().cpu().numpy(), {k: p.grad.cpu().numpy() for k, p in m.named_parameters()}, stats


def _cos(ga, gb):
    dots = n1 = n2 = 0.0
    worst = (1.0, '')
    gnorm = sum(float((g ** 2).sum()) for g in gb.values()) ** 0.5
    for k, g in gb.items():
        d = ga[k]
        assert np.isfinite(d).all(), k
        dd, gg, dg = float((d * d).sum()), float((g * g).sum()), float((d * g).sum())
        dots += dg; n1 += dd; n2 += gg
        if gg ** 0.5 > 1e-2 * gnorm:
            worst = min(worst, (dg / max(dd ** 0.5 * gg ** 0.5, 1e-30), k))
    return dots / (n1 ** 0.5 * n2 ** 0.5), worst


@pytest.mark.parametrize('kind', ['unet', 'saunet'])
def test_cp8_resident_tape_equals_the_converter_tape_and_tracks_fp32(kind):
    """Same tensor-core convolutions, same bf16 operand values; the element-wise stages see bf16-rounded activations / gradients instead
    of fp32 ones (one extra rounding at the max-pool, up-sampling and BatchNorm-backward outputs).  The two bf16 tapes must agree much more
    closely with each other than either does with the fp32 path."""
    l_new, y_new, g_new, s_new = _run(kind, 'bf16', True)
    l_old, y_old, g_old, s_old = _run(kind, 'bf16', False)
    l_32, y_32, g_32, s_32 = _run(kind, 'fp32', False)
    cos_no, worst_no = _cos(g_new, g_old)
    cos_n32, worst_n32 = _cos(g_new, g_32)
    cos_o32, worst_o32 = _cos(g_old, g_32)
    print(f'{kind}: loss cp8 {l_new:.6f} converter {l_old:.6f} fp32 {l_32:.6f}; cosine cp8/converter {cos_no:.5f} (worst {worst_no[1]} {worst_no[0]:.4f}), '
          f'cp8/fp32 {cos_n32:.5f} ({worst_n32[0]:.4f}), converter/fp32 {cos_o32:.5f} ({worst_o32[0]:.4f})')
    assert abs(l_new - l_old) < 2e-3 * abs(l_old) and abs(l_new - l_32) < 1e-2 * abs(l_32)
    assert np.abs(y_new - y_old).max() < 1e-2 and np.abs(y_new - y_32).max() < 2e-2
    assert cos_no > 0.98 and worst_no[0] > 0.90
    assert cos_n32 > cos_o32 - 0.02 and worst_n32[0] > worst_o32[0] - 0.05
    for k in s_new:                                                           # BatchNorm running statistics / num_batches_tracked
        assert np.abs(s_new[k].astype(np.float64) - s_32[k]).max() < 2e-2 * (1 + np.abs(s_32[k]).max()), k


def test_cp8_tape_forward_without_grad_equals_the_recorded_forward():
    m = _model('saunet', 'bf16')
    m.load_state_dict(fill_state_dict(m.state_dict(), 5, scheme='torch_default'))
    m = m.cuda().train()
    x = synth_patches(3, 5).cuda()
    m._train_calls = 0
    y1 = m(x)
    m._train_calls = 0
    with torch.no_grad():
        y2 = m(x)
    y2 = y2[0] if isinstance(y2, tuple) else y2
    assert torch.equal(y1.detach(), y2)


@pytest.mark.parametrize('B,Cin,Cout,W', [(25, 80, 50, 72), (3, 10, 16, 72), (5, 24, 11, 40)])
def test_full_height_convolution_on_the_tensor_cores(B, Cin, Cout, W):
    """conv3 (75 x 1, VALID, one output row; unet_cnns.py:380-385) forward / weight gradient / data gradient as tcgen05 GEMMs whose
    operands and results are addressed inside the NCHW tensors, against torch on the bf16-rounded operands."""
    from multipitch_architectures_b200 import ops, training
    H = 75
    conv = torch.nn.Conv2d(Cin, Cout, (H, 1)).cuda()
    x = rnd(B, Cin, H, W, seed=20).cuda()
    g = rnd(B, Cout, 1, W, seed=21).cuda()
    xr, wr, gr = bf(x).requires_grad_(True), bf(conv.weight.detach()).requires_grad_(True), bf(g)
    ref = F.leaky_relu(F.conv2d(xr, wr, conv.bias), 0.3)
    y = training._rows_tc_forward(conv, x, ops.ACT_LRELU, 0.3)
    assert (y - ref.detach()).abs().max() < 2e-3 * ref.abs().max()
    F.conv2d(xr, wr, conv.bias).backward(gr)
    gw, gb = torch.empty_like(conv.weight), torch.empty_like(conv.bias)
    gx = training._rows_tc_backward(conv, x, g, gw, gb, True)
    assert (gw - wr.grad).abs().max() < 2e-3 * wr.grad.abs().max()
    assert (gx - xr.grad).abs().max() < 2e-3 * xr.grad.abs().max()
    assert (gb - g.sum((0, 2, 3))).abs().max() < 1e-4 * g.abs().sum() / Cout


@pytest.mark.parametrize('kind', ['unet', 'saunet', 'cnn'])
def test_phase_split_conv2_equals_the_stride_emulation(kind):
    """Head conv2 (3x3, stride (1,3); basic_cnns.py:391, unet_cnns.py:376) in training: phase-split hand-over (producer writes bin f to phase
    f % 3, conv2 = stride-1 3x1 convolution over 3 * C0p channels, data gradient back in phase-split planes) against the stride-1 3x3
    convolution with sub-sampled output / zero-inserted gradient.  Same bf16 operand values, different accumulation order."""
    from multipitch_architectures_b200 import training
    from multipitch_architectures_b200.libdl import nn_models as M
    res = {}
    for split in (True, False):
        if kind == 'cnn':
            m = M.basic_cnn_segm_sigmoid(n_chan_input=6, n_chan_layers=[20, 20, 10, 1], n_bins_in=216, n_bins_out=72, precision='bf16')
        else:
            m = _model(kind, 'bf16')
        m.load_state_dict(fill_state_dict(m.state_dict(), 77, scheme='torch_default'))
        m = m.cuda().train()
        x, t = synth_patches(4, 77).cuda(), synth_targets(4, 77).cuda()
        from multipitch_architectures_b200 import training_unet
        calls = []
        orig = training._tc_s3_forward_split

        def spy(*a, **k):
            calls.append(1)
            return orig(*a, **k)
        old = training.S3_SPLIT
        training.S3_SPLIT = split
        training._tc_s3_forward_split = training_unet._tc_s3_forward_split = spy
        try:
            m._train_calls = 0
            y = m(x)
            loss = torch.nn.BCELoss(reduction='mean')(y, t)
            loss.backward()
        finally:
            training.S3_SPLIT = old
            training._tc_s3_forward_split = training_unet._tc_s3_forward_split = orig
        assert len(calls) == (1 if split else 0)             # the phase-split form really ran (or really did not)
        res[split] = (loss.item(), y.detach().cpu().numpy(), {k: p.grad.cpu().numpy() for k, p in m.named_parameters()})
    (la, ya, ga), (lb, yb, gb) = res[True], res[False]
    cos, worst = _cos(ga, gb)
    print(f'{kind}: loss split {la:.6f} emulated {lb:.6f}; max|dy| {np.abs(ya - yb).max():.2e}; gradient cosine {cos:.6f}, worst {worst[1]} {worst[0]:.5f}')
    assert abs(la - lb) < 1e-3 * abs(lb) and np.abs(ya - yb).max() < 5e-3
    assert cos > 0.995 and worst[0] > 0.98
    assert np.abs(ga['conv2.0.weight'] - gb['conv2.0.weight']).max() < 2e-2 * np.abs(gb['conv2.0.weight']).max()
