"""`-m gpu`: tcgen05 weight-gradient kernel (mpa_conv_wgrad_tc) against torch autograd of nn.Conv2d on the same 16-bit-rounded
operands (fp32 accumulate on both sides: only the summation order differs), and the bias-gradient reduction."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize('fmt', ['bf16', 'fp16'])
@pytest.mark.parametrize('cfg', [
    # B, Cin, Cout, T, F, K
    (3, 6, 20, 75, 216, 15), (2, 40, 40, 75, 216, 15), (2, 16, 8, 37, 108, 15), (2, 64, 32, 18, 54, 9), (2, 128, 128, 9, 27, 5),
    (3, 256, 128, 4, 13, 3), (2, 32, 200, 9, 27, 3), (5, 8, 16, 75, 216, 15), (1, 6, 40, 75, 216, 15),
])
def test_wgrad_tc_matches_autograd(cfg, fmt):
    from multipitch_architectures_b200 import ops
    B, Cin, Cout, T, Fq, K = cfg
    dt = torch.bfloat16 if fmt == 'bf16' else torch.float16
    x = rnd(B, Cin, T, Fq, seed=1).to(dt).float()
    g = rnd(B, Cout, T, Fq, seed=2, scale=0.05).to(dt).float()
    w = torch.zeros(Cout, Cin, K, K, requires_grad=True)
    F.conv2d(x, w, padding=K // 2).backward(g)
    f = ops.fmt_of(fmt)
    pitch = (Fq + 8 + 15) // 16 * 16
    xc = ops.nchw_to_cp8(x.cuda(), pitch=pitch, fmt=f)
    gc = ops.nchw_to_cp8(g.cuda(), pitch=pitch, fmt=f)
    gw = torch.full((Cout, Cin, K, K), 7.0, device='cuda')
    ops.conv_wgrad_tc(xc, gc, gw, (K, K))
    ref = w.grad
    err = (gw.cpu() - ref).abs().max().item()
    print(f'{cfg} {fmt}: max|diff| = {err:.3e} of max|gw| = {ref.abs().max().item():.3e}')
    assert err <= 2e-4 * max(1.0, ref.abs().max().item())
    gb = ops.channel_sum(g.cuda())
    assert (gb.cpu() - g.sum(dim=(0, 2, 3))).abs().max() < 1e-3


@pytest.mark.parametrize('cfg', [
    # B, Cin, Cout, T, F, K   (shapes of the U-Net levels, incl. channel counts that are not multiples of 8 and > 128)
    (3, 6, 4, 75, 216, 15), (3, 4, 4, 75, 216, 15), (3, 4, 8, 37, 108, 15), (3, 8, 16, 18, 54, 9), (3, 16, 32, 9, 27, 5),
    (3, 32, 32, 4, 13, 3), (3, 64, 16, 9, 27, 3), (3, 32, 8, 18, 54, 5), (3, 16, 4, 37, 108, 9), (3, 8, 4, 75, 216, 15),
    (2, 40, 40, 75, 216, 15), (2, 256, 192, 9, 27, 3), (2, 24, 136, 18, 54, 9),
])
def test_tc_conv_forward_backward_matches_fp32_conv(cfg):
    """TcConv (bf16 tensor-core forward / dgrad / wgrad used by the training paths) against torch's fp32 convolution on
    bf16-rounded operands: the only differences left are the 16-bit rounding of outputs and the summation order."""
    import torch.nn as nn
    from multipitch_architectures_b200.training import TcConv
    from multipitch_architectures_b200 import ops
    B, Cin, Cout, T, Fq, K = cfg
    conv = nn.Conv2d(Cin, Cout, (K, K), padding=K // 2)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.to(torch.bfloat16).float())
    x = rnd(B, Cin, T, Fq, seed=3).to(torch.bfloat16).float().requires_grad_(True)
    y = conv(x)
    g = rnd(*y.shape, seed=4, scale=1e-4).to(torch.bfloat16).float()
    y.backward(g)
    cc = nn.Conv2d(Cin, Cout, (K, K), padding=K // 2).cuda()
    cc.load_state_dict(conv.state_dict())
    yt, xc = TcConv.forward('t', cc, x.detach().cuda(), ops.ACT_NONE, 0.0)
    gw, gb = torch.empty_like(cc.weight), torch.empty_like(cc.bias)
    gx = TcConv.backward('t', cc, xc, g.cuda(), gw, gb, True)
    rel = lambda a, b: ((a.cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    e = (rel(yt, y.detach()), rel(gx, x.grad), rel(gw, conv.weight.grad), rel(gb, conv.bias.grad))
    print(f'{cfg}: rel err y {e[0]:.2e} gx {e[1]:.2e} gw {e[2]:.2e} gb {e[3]:.2e}')
    assert e[0] < 8e-3 and e[1] < 8e-3 and e[2] < 2e-3 and e[3] < 1e-4


@pytest.mark.parametrize('fmt', ['fp16', 'bf16'])
@pytest.mark.parametrize('M,N,K,relu', [(1300, 8192, 128, True), (1300, 128, 8192, False), (52, 64, 32, True), (2600, 384, 128, False),
                                        (300, 200, 72, False), (257, 129, 8200, False)])
def test_gemm_tc_matches_linear(M, N, K, relu, fmt):
    """mpa_gemm_tc_f16 (the encoder MLP's nn.Linear layers on tcgen05) against torch on the same 16-bit-rounded operands."""
    from multipitch_architectures_b200 import ops
    dt = torch.bfloat16 if fmt == 'bf16' else torch.float16
    x = rnd(M, K, seed=5).to(dt).float()
    w = (rnd(N, K, seed=6) / K ** 0.5).to(dt).float()
    b = rnd(N, seed=7)
    ref = x @ w.T + b
    if relu:
        ref = torch.relu(ref)
    f = ops.fmt_of(fmt)
    y = ops.gemm_tc(x.cuda(), ops.gemm_tc_chunks(w.cuda(), 128, f), b.cuda(), N, relu, f).cpu()
    err = (y - ref).abs().max().item()
    print(f'gemm_tc {M}x{N}x{K} {fmt}: max|diff| = {err:.2e}')
    assert err < 2e-4 * max(1.0, ref.abs().max().item())
