"""Executors shared by the model classes: parameter -> packed-operand caches and the kernel sequences.

The nn.Module classes in this package own parameters with exactly the reference's state_dict names; their
`forward` never calls a torch operator on activations — every stage below is a libmpa kernel
(include/mpa.h).  torch is used to hold memory and to do one-off parameter preparation (weight packing,
BatchNorm folding)."""
import math

import torch

from ... import ops
from ..._lib import MpaError

PRECISIONS = ('fp32', 'fp16', 'bf16')     # fp16 / bf16: tcgen05 tensor-core path (16-bit operands, fp32 accumulate)


class ParamCache:
    """Derived operands (packed weights, folded BN) keyed by the versions of the parameters they came from."""

    def __init__(self):
        self._d = {}

    def get(self, key, params, build):
        sig = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        hit = self._d.get(key)
        if hit is None or hit[0] != sig:
            with torch.no_grad():
                hit = (sig, build())
            self._d[key] = hit
        return hit[1]


def _check_input(x, n_chan, n_bins, min_T=1):
    if not isinstance(x, torch.Tensor) or x.dim() != 4:
        raise ValueError('expected input of shape [B, C, T, F]')
    if not x.is_cuda:
        raise MpaError('this model runs on CUDA (sm_100a) only: move the input and the model to a B200 '
                       '(there is no CPU path)')
    if x.shape[1] != n_chan or x.shape[3] != n_bins:
        raise ValueError(f'expected [B,{n_chan},T,{n_bins}], got {tuple(x.shape)}')
    return x.contiguous().float()


def conv_f32(cache, name, conv, x, act=ops.ACT_NONE, act_param=0.0, bn=None, bn_train=False, x2=None):
    """nn.Conv2d (+ BatchNorm2d + activation) through mpa_conv2d_f32."""
    w = conv.weight
    Cout, Cin, KH, KW = w.shape
    wp = cache.get(name + ':w32', [w], lambda: ops.pack_conv_weight(w))
    stride, padding = tuple(conv.stride), tuple(conv.padding)
    if bn is None:
        return ops.conv2d(x, wp, conv.bias, Cout, (KH, KW), stride, padding, act, act_param, x2=x2)
    if not bn_train:
        def fold():
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            return s.contiguous(), (bn.bias - bn.running_mean * s).contiguous()
        scale, shift = cache.get(name + ':bnfold', [bn.weight, bn.bias, bn.running_mean, bn.running_var], fold)
        return ops.conv2d(x, wp, conv.bias, Cout, (KH, KW), stride, padding, act, act_param, scale=scale, shift=shift, x2=x2)
    # train mode: batch statistics (biased variance), running stats updated like nn.BatchNorm2d
    y = ops.conv2d(x, wp, conv.bias, Cout, (KH, KW), stride, padding, ops.ACT_NONE, 0.0, x2=x2)
    stats = ops.bn_stats(y)
    if bn.track_running_stats and bn.training:
        with torch.no_grad():
            n = y.numel() // Cout
            m = bn.momentum if bn.momentum is not None else 0.1
            bn.running_mean.mul_(1 - m).add_(stats[:Cout], alpha=m)
            bn.running_var.mul_(1 - m).add_(stats[Cout:] * (n / max(n - 1, 1)), alpha=m)
            bn.num_batches_tracked += 1
    return ops.bn_apply(y, stats, bn.weight, bn.bias, bn.eps, act, act_param)


def head_f32(cache, model, x, a):
    """conv2 (3x3, stride (1,3)) -> LReLU -> maxpool(13,1) -> conv3 (75x1) -> LReLU -> 1x1 -> LReLU -> 1xk -> sigmoid
    (basic_cnns.py:389-408; identical in every model)."""
    y = conv_f32(cache, 'conv2', model.conv2[0], x, ops.ACT_LRELU, a)
    y = ops.maxpool_time(y, 13)
    y = conv_f32(cache, 'conv3', model.conv3[0], y, ops.ACT_LRELU, a)
    y = conv_f32(cache, 'conv4.0', model.conv4[0], y, ops.ACT_LRELU, a)
    return conv_f32(cache, 'conv4.3', model.conv4[3], y, ops.ACT_SIGMOID)


def head_tc(cache, model, zc, a):
    """Head on a CP8 activation: conv2 (3x3, stride (1,3)) runs on the tensor cores as the stride-1 3x3 convolution
    whose epilogue keeps columns 1, 4, 7, ... (exactly the strided outputs), then maxpool(13,1) and the fused tail."""
    conv2, conv3, c40, c43 = model.conv2[0], model.conv3[0], model.conv4[0], model.conv4[3]
    ok2 = (tuple(conv2.kernel_size) == (3, 3) and tuple(conv2.stride) == (1, 3) and tuple(conv2.padding) == (1, 0)
           and conv2.weight.shape[0] <= 128 and zc.F % 3 == 0)
    if not ok2:
        return head_f32(cache, model, ops.cp8_to_nchw(zc), a)
    w2 = conv2.weight
    wp = cache.get(f'conv2:wtc{zc.fmt}', [w2], lambda: ops.conv_tc_pack(w2, zc.buf.device, zc.fmt))
    yc = ops.conv_tc(zc, wp, conv2.bias, w2.shape[0], (3, 3), ops.ACT_LRELU, a, subsample=(3, 1))
    yc = ops.pool_time_res_cp8(yc, 13)
    fused = (conv3.kernel_size[0] == zc.T and conv3.kernel_size[1] == 1 and tuple(c43.kernel_size) == (1, 1)
             and conv3.weight.shape[0] <= 32 and c40.weight.shape[0] <= 16 and c43.weight.shape[0] == 1 and yc.F <= 256)
    if fused:
        o = ops.head_tail(yc, conv3.weight, conv3.bias, c40.weight, c40.bias, c43.weight, c43.bias, a)
        return o.reshape(o.shape[0], 1, 1, o.shape[1])
    y = ops.cp8_to_nchw(yc)
    y = conv_f32(cache, 'conv3', conv3, y, ops.ACT_LRELU, a)
    y = conv_f32(cache, 'conv4.0', c40, y, ops.ACT_LRELU, a)
    return conv_f32(cache, 'conv4.3', c43, y, ops.ACT_SIGMOID)


def cnn_blocks(model):
    return [('conv1', model.conv1[0])] + [(f'prefilt_list.{i}', m[0]) for i, m in enumerate(getattr(model, 'prefilt_list', []))]


def tc_eligible(model, F):
    return (model.precision in ('fp16', 'bf16') and F + 8 <= 256
            and all(c.weight.shape[0] <= 128 and c.kernel_size[0] % 2 == 1 and c.kernel_size[1] % 2 == 1 for _, c in cnn_blocks(model)))


def cnn_forward(model, x):
    """basic_cnn_segm_sigmoid / deep_cnn_segm_sigmoid inference forward."""
    cache, a = model._cache, model.a_lrelu
    blocks = cnn_blocks(model)
    residual = getattr(model, 'residual', False)
    z = ops.layernorm_cf(x, model.layernorm.weight, model.layernorm.bias, model.layernorm.eps)
    if not tc_eligible(model, x.shape[3]):
        for i, (name, conv) in enumerate(blocks):
            y = conv_f32(cache, name, conv, z, ops.ACT_LRELU, a)
            z = ops.maxpool_time(y, 3, res=z if (residual and i > 0) else None)
        return head_f32(cache, model, z, a)
    fmt = ops.fmt_of(model.precision)
    zc = ops.nchw_to_cp8(z, fmt=fmt)
    for i, (name, conv) in enumerate(blocks):
        w = conv.weight
        wp = cache.get(f'{name}:wtc{fmt}', [w], lambda: ops.conv_tc_pack(w, x.device, fmt))
        yc = ops.conv_tc(zc, wp, conv.bias, w.shape[0], tuple(conv.kernel_size), ops.ACT_LRELU, a)
        zc = ops.pool3_res_cp8(yc, res=zc if (residual and i > 0) else None)
    return head_tc(cache, model, zc, a)


def double_conv_f32(cache, name, dc, x, train, x2=None):
    seq = dc.double_conv
    y = conv_f32(cache, name + '.0', seq[0], x, ops.ACT_RELU, 0.0, bn=seq[1], bn_train=train, x2=x2)
    y = conv_f32(cache, name + '.4', seq[4], y, ops.ACT_RELU, 0.0, bn=seq[5], bn_train=train)
    return y


def unet_trunk_f32(model, x, train):
    cache = model._cache
    z = ops.layernorm_cf(x, model.layernorm.weight, model.layernorm.bias, model.layernorm.eps)
    x1 = double_conv_f32(cache, 'inc', model.inc, z, train)
    x2 = double_conv_f32(cache, 'down1', model.down1[1], ops.maxpool2d(x1, (2, 2), (2, 2)), train)
    x3 = double_conv_f32(cache, 'down2', model.down2[1], ops.maxpool2d(x2, (2, 2), (2, 2)), train)
    x4 = double_conv_f32(cache, 'down3', model.down3[1], ops.maxpool2d(x3, (2, 2), (2, 2)), train)
    x5 = double_conv_f32(cache, 'down4', model.down4[1], ops.maxpool2d(x4, (2, 2), (2, 2)), train)
    return x1, x2, x3, x4, x5


def unet_up_f32(model, x5, skips, train):
    cache = model._cache
    x1, x2, x3, x4 = skips
    u = double_conv_f32(cache, 'upconv1', model.upconv1, ops.upsample2x_concat(x5, x4), train)
    u = double_conv_f32(cache, 'upconv2', model.upconv2, ops.upsample2x_concat(u, x3), train)
    u = double_conv_f32(cache, 'upconv3', model.upconv3, ops.upsample2x_concat(u, x2), train)
    u = double_conv_f32(cache, 'upconv4', model.upconv4, ops.upsample2x_concat(u, x1), train)
    return u


def sinusoidal_pe(n, E, device):
    position = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, E, 2, dtype=torch.float32) * (-math.log(10000.0) / E))
    pe = torch.zeros(n, E)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(device)
