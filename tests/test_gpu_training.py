"""`-m gpu`: training path (configuration 2) — backward kernels against torch autograd of the same op,
whole-model loss/gradients against the REFERENCE goldens (loss.backward() on the reference modules), and the fused
TrainStep (BCE + backward + AdamW kernels) against torch.optim.AdamW driving the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.refshapes import build_model
from tests.weights import fill_state_dict, synth_patches, synth_targets

pytestmark = pytest.mark.gpu


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize('cfg', [
    # B, Cin, Cout, H, W, KH, KW, sh, sw, ph, pw
    (3, 6, 20, 30, 216, 15, 15, 1, 1, 7, 7), (2, 20, 20, 75, 216, 3, 3, 1, 3, 1, 0), (4, 20, 10, 75, 72, 75, 1, 1, 1, 0, 0),
    (4, 10, 1, 1, 72, 1, 1, 1, 1, 0, 0), (2, 8, 8, 20, 50, 15, 15, 1, 1, 7, 7), (2, 40, 40, 12, 40, 15, 15, 1, 1, 7, 7),
])
def test_conv_dgrad_wgrad(cfg):
    import torch.nn as nn
    from multipitch_architectures_b200 import training as TR
    B, Cin, Cout, H, W, KH, KW, sh, sw, ph, pw = cfg
    conv = nn.Conv2d(Cin, Cout, (KH, KW), stride=(sh, sw), padding=(ph, pw))
    x = rnd(B, Cin, H, W, seed=1).requires_grad_(True)
    y = conv(x)
    g = rnd(*y.shape, seed=2)
    y.backward(g)
    convc = nn.Conv2d(Cin, Cout, (KH, KW), stride=(sh, sw), padding=(ph, pw)).cuda()
    convc.load_state_dict(conv.state_dict())
    gi = TR._dgrad(convc, g.cuda(), x.shape).cpu()
    assert (gi - x.grad).abs().max() < 2e-5 * max(1.0, x.grad.abs().max().item())
    gw, gb = torch.empty_like(convc.weight), torch.empty_like(convc.bias)
    TR._wgrad(convc, x.detach().cuda(), g.cuda(), gw, gb)
    assert (gw.cpu() - conv.weight.grad).abs().max() < 1e-4 * max(1.0, conv.weight.grad.abs().max().item())
    assert (gb.cpu() - conv.bias.grad).abs().max() < 1e-4 * max(1.0, conv.bias.grad.abs().max().item())


@pytest.mark.parametrize('k', [3, 13])
def test_pool_act_bwd_and_dropout(k):
    from multipitch_architectures_b200 import training as TR, ops
    y = rnd(2, 5, 75, 40, seed=3).requires_grad_(True)
    a = F.leaky_relu(y, 0.3)
    p = F.max_pool2d(a, (k, 1), (1, 1), (k // 2, 0))
    g = rnd(*p.shape, seed=4)
    p.backward(g)
    got = TR._pool_bwd(a.detach().cuda(), g.cuda(), k, ops.ACT_LRELU, 0.3).cpu()
    assert (got - y.grad).abs().max() < 1e-5      # up to k windows add into one row in arbitrary order
    x = torch.ones(100003).cuda()
    d = TR._dropout(x, 0.2, 7, 3)
    keep = (d != 0).float().mean().item()
    assert abs(keep - 0.8) < 0.01 and abs(d.max().item() - 1.25) < 1e-6
    assert torch.equal(d, TR._dropout(x, 0.2, 7, 3)) and not torch.equal(d, TR._dropout(x, 0.2, 7, 4))


@pytest.mark.parametrize('name', ['cnn_xs', 'drcnn_tiny'])
def test_model_loss_and_grads_match_reference_golden(nn_golden, name):
    tag = f'{name}__eval'
    B, seed, _ = nn_golden[tag + '__meta']
    B, seed = int(B), int(seed)
    m = build_model(name)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed))
    m.p_dropout = 0.0
    m = m.cuda().train()
    x, t = synth_patches(B, seed).cuda(), synth_targets(B, seed).cuda()
    y = m(x)                                                    # autograd path (CnnTrainFunction)
    loss = torch.nn.BCELoss(reduction='mean')(y, t)
    loss.backward()
    assert abs(loss.item() - float(nn_golden[tag + '__loss'][0])) < 1e-5
    for k, p in m.named_parameters():
        g = nn_golden[tag + '__grad__' + k]
        d = np.abs(p.grad.cpu().numpy() - g)
        # fp32 summation-order noise of 48600-term signed sums + max-pool near-tie flips (see test_oracle_pinned.py)
        assert d.max() <= 5e-3 * np.abs(g).max() and d.mean() <= 1e-3 * np.abs(g).max(), k


def test_fused_train_step_matches_torch_adamw_on_oracle():
    from oracle import nn_oracle as NO
    from multipitch_architectures_b200.training import TrainStep
    m = build_model('cnn_xs')
    sd0 = fill_state_dict(m.state_dict(), 31)
    m.load_state_dict(sd0)
    m.p_dropout = 0.0
    m = m.cuda().train()
    step = TrainStep(m, lr=1e-3, weight_decay=0.01)
    ref = {k: v.clone().requires_grad_(True) for k, v in sd0.items()}
    opt = torch.optim.AdamW(list(ref.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    losses = []
    for it in range(3):
        x, t = synth_patches(4, 40 + it), synth_targets(4, 40 + it)
        l_gpu = step(x.cuda(), t.cuda()).item()
        opt.zero_grad()
        l_ref = NO.bce_mean(NO.cnn_forward(ref, x), t)
        l_ref.backward()
        opt.step()
        losses.append((l_gpu, l_ref.item()))
        assert abs(l_gpu - l_ref.item()) < 2e-4 * max(1.0, abs(l_ref.item())), losses
    for k, p in m.named_parameters():
        # Adam's first steps move every weight by ~lr regardless of gradient scale; compare the UPDATE, not the value
        upd_ref = (ref[k].detach() - sd0[k])
        upd_gpu = (p.detach().cpu() - sd0[k])
        # elements whose gradient is at fp32-noise level at some step get a random +-lr there (sign of m/sqrt(v)), so
        # bound the mean and the share of outliers rather than the maximum
        d = (upd_gpu - upd_ref).abs()
        assert d.mean() < 0.02 * 3e-3 and (d > 0.3e-3).float().mean() < 0.02, k
    # inference after training sees the updated weights (packed-operand cache invalidated)
    m.eval()
    x = synth_patches(2, 50)
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        yr = NO.cnn_forward({k: v.detach() for k, v in ref.items()}, x)
    assert (y - yr).abs().max() < 1e-3


def test_cnn_graph_replayed_step_equals_eager_step():
    """TrainStep(graph=True): steps 3.. replay one captured forward+backward CUDA graph with dropout on; tiny learning rate so that the
    loss of step k depends on the data and the dropout masks of step k only."""
    from multipitch_architectures_b200.training import TrainStep
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict, synth_patches, synth_targets
    res = {}
    for mode in (False, True):
        m = build_model('drcnn_tiny', precision='bf16')
        m.load_state_dict(fill_state_dict(m.state_dict(), 5, scheme='torch_default'))
        m = m.cuda().train()
        assert m.p_dropout > 0
        step = TrainStep(m, lr=1e-6, graph=mode)
        losses = []
        for i in range(6):
            losses.append(float(step(synth_patches(6, 70 + i).cuda(), synth_targets(6, 70 + i).cuda()).item()))
        res[mode] = losses
        if mode:
            assert step.replays == 4 and step.launches_per_replay > 20
    print('eager', res[False], 'graph', res[True])
    assert all(abs(a - b) <= 2e-4 * max(1.0, abs(a)) for a, b in zip(res[False], res[True]))
    assert len(set(round(v, 4) for v in res[True])) == 6


def test_fused_pool_dropout_kernels_equal_the_separate_kernels():
    """MaxPool -> Dropout (-> + residual) in one kernel and its backward with the mask re-drawn on the fly: bit-identical to
    maxpool_time -> dropout -> add and dropout -> maxpool_time_bwd (same Philox convention), through a whole training step."""
    from multipitch_architectures_b200 import training as TR
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict, synth_patches, synth_targets
    res = {}
    for fuse in (True, False):
        TR.FUSE_POOL_DROPOUT = fuse
        try:
            m = build_model('drcnn_tiny')
            m.load_state_dict(fill_state_dict(m.state_dict(), 9, scheme='torch_default'))
            m = m.cuda().train()
            assert m.p_dropout > 0
            x, t = synth_patches(4, 81).cuda(), synth_targets(4, 81).cuda()
            m.dropout_seed, m._train_calls = 1234, 0
            y = m(x)
            loss = torch.nn.BCELoss()(y, t)
            loss.backward()
            res[fuse] = (y.detach().clone(), {k: p.grad.clone() for k, p in m.named_parameters()})
        finally:
            TR.FUSE_POOL_DROPOUT = True
    assert torch.equal(res[True][0], res[False][0])
    for k in res[True][1]:
        a, b = res[True][1][k], res[False][1][k]
        assert (a - b).abs().max().item() <= 1e-6 * max(1.0, b.abs().max().item()), k       # weight gradients meet in fp32 atomics
