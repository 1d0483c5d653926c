"""Executors shared by the model classes: parameter -> packed-operand caches and the kernel sequences.

The nn.Module classes in this package own parameters with exactly the reference's state_dict names; their
`forward` never calls a torch operator on activations — every stage below is a libmpa kernel
(include/mpa.h).  torch is used to hold memory and to do one-off parameter preparation (weight packing,
BatchNorm folding)."""
import logging
import math
import os
import warnings

import torch

from ... import ops
from ... import _lib
from ..._lib import MpaError

# fp16 / bf16: tcgen05 tensor-core path (16-bit operands and storage, fp32 accumulate); fp16x3: the same kernels in split precision
# (hi/lo fp16 pairs, three MMA passes per product: fp32-class results at a third of the fp16 rate); fp32: exact CUDA-core path
PRECISIONS = ('fp32', 'fp16', 'bf16', 'fp16x3')
TC_PRECISIONS = ('fp16', 'bf16', 'fp16x3')
# precision of a model constructed without the keyword (the reference's constructors have none): see install_as_libdl / set_default_precision
DEFAULT_PRECISION = os.environ.get('MPA_PRECISION', 'fp32')
log = logging.getLogger('multipitch_architectures_b200')
_warned = set()


def note_path(model, what):
    """Say ONCE per (model class, reason) which kernel path a tensor-core model actually takes when it cannot take the tcgen05 one."""
    key = (type(model).__name__, what)
    if key not in _warned:
        _warned.add(key)
        warnings.warn(f'{type(model).__name__}(precision={model.precision!r}): {what}', RuntimeWarning, stacklevel=3)
        log.warning('%s(precision=%r): %s', type(model).__name__, model.precision, what)


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


def invalidate_caches(model):
    """Clear every ParamCache of the module tree.  Needed after an in-place update through raw pointers (the fused AdamW kernel), which
    changes neither data_ptr nor _version of a parameter."""
    for m in model.modules():
        c = getattr(m, '_cache', None)
        if isinstance(c, ParamCache):
            c._d.clear()


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
    stride, padding = tuple(conv.stride), tuple(conv.padding)
    if bn is None and x2 is None and (KH, KW) == (x.shape[2], 1) and stride == (1, 1) and padding == (0, 0):
        # full-height VALID convolution (conv3 on a 75-frame patch): GEMM kernel instead of 16-row output tiles
        B, _, H, W = x.shape
        out = torch.empty(B, Cout, 1, W, dtype=torch.float32, device=x.device)
        ws_bytes = _lib.lib().mpa_conv_rows_fwd_workspace(B, Cin, H, W, Cout)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        _lib.call('conv_rows_fwd_f32', x, w.detach().contiguous(), conv.bias, out, B, Cin, H, W, Cout, act, float(act_param), ws,
                  _lib.usize(ws_bytes), _lib.stream_ptr())
        return out
    wp = cache.get(name + ':w32', [w], lambda: ops.pack_conv_weight(w))
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


def head_split_factor(model, F):
    """3 when the head's conv2 is the reference's 3x3 / stride (1,3) / pad (1,0) binning convolution and F is a multiple of 3: its producer
    may then write phase-split planes (conv_tc out_mode 2) and conv2 runs as a stride-1 3x1 convolution over 3*Cin channels of width F/3."""
    conv2 = model.conv2[0]
    ok = tuple(conv2.kernel_size) == (3, 3) and tuple(conv2.stride) == (1, 3) and tuple(conv2.padding) == (1, 0) and F % 3 == 0
    return 3 if ok else None


def _split_conv2_J(C0p, C1p):
    """Output rows per work unit for the phase-split conv2: 3 when the whole weight set ((3 + J - 1) K rows of one 8-MMA stage each) then
    stays resident in the kernel's 5 weight stages (CNN family: 3*C0p <= 128 channels); else 1, which lets the kernel merge R = 3 image
    rows into one operand row of N = 240 columns (every streamed weight tile then serves 3 output rows; J > 1 would disable that)."""
    return 3 if (3 * C0p // 8 + 1) // 2 <= 8 and 3 * C1p <= 128 else 1


def _split_conv2_blocks(cache, conv2, C0p, fmt, dev, J=0):
    """conv2's weights re-laid for the phase-split input: w'[co][ph*C0p + ci][kh][0] = w[co][ci][kh][ph]."""
    def build():
        w = conv2.weight.detach().float()
        Cout, C0 = w.shape[0], w.shape[1]
        w2 = torch.zeros(Cout, 3 * C0p, 3, 1, dtype=w.dtype, device=w.device)
        for ph in range(3):
            w2[:, ph * C0p:ph * C0p + C0, :, 0] = w[:, :, :, ph]
        import types
        return types.SimpleNamespace(weight=w2, bias=conv2.bias, kernel_size=(3, 1))
    shim = cache.get(f'conv2:splitw:{C0p}', [conv2.weight, conv2.bias], build)
    return _folded_tc(cache, f'conv2split{C0p}', shim, None, fmt, dev, J=J)


def head_tc_supported(model, T, Fin):
    """The head shape the tensor-core sequence covers: the reference's 3x3 / stride (1,3) binning conv2, a full-height odd conv3 and 1x1
    convolutions in conv4 (n_bins_out == n_bins_in // 3, as in every experiment script)."""
    conv2, conv3, c40, c43 = model.conv2[0], model.conv3[0], model.conv4[0], model.conv4[3]
    KH3 = conv3.kernel_size[0]
    return (tuple(conv2.kernel_size) == (3, 3) and tuple(conv2.stride) == (1, 3) and tuple(conv2.padding) == (1, 0) and Fin % 3 == 0
            and conv3.kernel_size[1] == 1 and (KH3 & 1) and tuple(conv3.padding) == (0, 0) and T >= KH3
            and tuple(c40.kernel_size) == (1, 1) and tuple(c43.kernel_size) == (1, 1) and c43.weight.shape[0] == 1)


def head_tc(cache, model, zc, a, split=None, out=None, pool=None):
    """Head on a CP8 activation, all on the tensor cores:
      conv2 (3x3, stride (1,3), pad (1,0)) = the stride-1 3x3 convolution whose epilogue keeps columns 1, 4, 7, ...;
      maxpool(13,1); conv3 (75x1, VALID) = the 'same' 75x1 convolution restricted to the rows whose window fits;
      conv4.0 / conv4.3 / sigmoid in one small kernel.  Channel counts are padded to multiples of 8 with zero weights."""
    conv2, conv3, c40, c43 = model.conv2[0], model.conv3[0], model.conv4[0], model.conv4[3]
    KH3 = conv3.kernel_size[0]
    Fin = zc.F * split if split else zc.F
    if not head_tc_supported(model, zc.T, Fin):
        if split:
            raise MpaError('head_tc: phase-split input needs the tensor-core head')
        y = head_f32(cache, model, ops.cp8_to_nchw(zc), a)
        if out is not None:
            out.copy_(y.reshape(out.shape))
        return y
    dev, fmt = zc.buf.device, zc.fmt
    C1p = (conv2.weight.shape[0] + 7) // 8 * 8
    yc = ops.compact_cp8(zc.B, C1p, zc.T, Fin // 3, dev, fmt, pool, 'head:y')
    if split:
        # producer wrote phase-split planes: conv2 = stride-1 3x1 convolution over 3*C0p channels of width F/3 (3x fewer MMA columns)
        J2 = _split_conv2_J(zc.C // 3, C1p)
        for wp, b, c0, c in _split_conv2_blocks(cache, conv2, zc.C // 3, fmt, dev, J=J2):
            ops.conv_tc(zc, wp, b, c, (3, 1), ops.ACT_LRELU, a, subsample=(1, 0), out=yc.channels(c0, c), J=J2)
    else:
        cin_pad = zc.C if zc.C != conv2.weight.shape[1] else None          # the producer padded its channels to whole chunks
        for wp, b, c0, c in _folded_tc(cache, 'conv2', conv2, None, fmt, dev, cin_pad=cin_pad):
            ops.conv_tc(zc, wp, b, c, (3, 3), ops.ACT_LRELU, a, subsample=(3, 1), out=yc.channels(c0, c))
    n_rows = zc.T - KH3 + 1
    if (FUSED_THIN_TAIL and fmt in (ops.FMT_F16, ops.FMT_BF16) and KH3 == 75 and n_rows == 1 and conv3.weight.shape[0] <= 16 and C1p <= 64
            and c40.weight.shape[0] <= 64):
        # CNN family (conv3 far too thin for the tensor cores): pool13 + conv3 + conv4.* in one pass over the conv2 output
        w3p = cache.get(f'conv3:rows{C1p}', [conv3.weight], lambda: ops.pack_conv3_rows(conv3.weight, C1p))
        o = ops.head_pool_conv3_tail(yc, w3p, conv3.bias, c40.weight, c40.bias, c43.weight, c43.bias, a, out=out)
        return o.reshape(zc.B, 1, 1, yc.F)
    yc = ops.pool_time_res_cp8(yc, 13, out=ops.compact_cp8(zc.B, C1p, zc.T, Fin // 3, dev, fmt, pool, 'head:p') if pool is not None else None)
    C2p = (conv3.weight.shape[0] + 7) // 8 * 8
    hc = ops.compact_cp8(zc.B, C2p, n_rows, yc.F, dev, fmt, pool, 'head:h')
    for wp, b, c0, c in _folded_tc(cache, 'conv3', conv3, None, fmt, dev, cin_pad=C1p, J=1):
        ops.conv_tc(yc, wp, b, c, (KH3, 1), ops.ACT_LRELU, a, out=hc.channels(c0, c), J=1, rows=(KH3 // 2, n_rows))
    w40 = cache.get(f'conv4.0:padded{C2p}', [c40.weight], lambda: _pad_cols(c40.weight.detach().reshape(c40.weight.shape[0], -1), C2p).contiguous())
    o = ops.head_tail2(hc, w40, c40.bias, c43.weight, c43.bias, a, out=out)
    return o.reshape(zc.B, 1, n_rows, yc.F)


def _pad_cols(w, n):
    if w.shape[1] == n:
        return w
    out = torch.zeros(w.shape[0], n, dtype=w.dtype, device=w.device)
    out[:, :w.shape[1]] = w.detach()
    return out


def cnn_blocks(model):
    return [('conv1', model.conv1[0])] + [(f'prefilt_list.{i}', m[0]) for i, m in enumerate(getattr(model, 'prefilt_list', []))]


def tc_eligible(model, F):
    return (model.precision in TC_PRECISIONS and F + 8 <= 256
            and all(c.weight.shape[0] <= 128 and c.kernel_size[0] % 2 == 1 and c.kernel_size[1] % 2 == 1 for _, c in cnn_blocks(model)))


def cnn_forward(model, x):
    """basic_cnn_segm_sigmoid / deep_cnn_segm_sigmoid inference forward."""
    cache, a = model._cache, model.a_lrelu
    blocks = cnn_blocks(model)
    residual = getattr(model, 'residual', False)
    z = ops.layernorm_cf(x, model.layernorm.weight, model.layernorm.bias, model.layernorm.eps)
    if not tc_eligible(model, x.shape[3]):
        if model.precision in TC_PRECISIONS:
            note_path(model, 'this configuration runs on the fp32 CUDA-core kernels, not on the tensor cores (the tcgen05 path needs <= 128 '
                             'channels per block, odd kernel sizes and n_bins_in <= 248)')
        for i, (name, conv) in enumerate(blocks):
            y = conv_f32(cache, name, conv, z, ops.ACT_LRELU, a)
            z = ops.maxpool_time(y, 3, res=z if (residual and i > 0) else None)
        return head_f32(cache, model, z, a)
    fmt = ops.fmt_of(model.precision)
    zc = ops.nchw_to_cp8(z, fmt=fmt)
    for i, (name, conv) in enumerate(blocks):
        w = conv.weight
        if fmt == ops.FMT_F16X3:
            # split precision stores whole channel chunks: pad the block width with zero filters (LeakyReLU / pool / residual keep zeros)
            cout = (w.shape[0] + 7) // 8 * 8
            wp, bias = block_operands(cache, name, conv, fmt, x.device, w.shape[1] if i == 0 else zc.C, cout)
        else:
            cout, bias = w.shape[0], conv.bias
            wp = cache.get(f'{name}:wtc{fmt}', [w], lambda: ops.conv_tc_pack(w, x.device, fmt))
        yc = ops.conv_tc(zc, wp, bias, cout, tuple(conv.kernel_size), ops.ACT_LRELU, a)
        zc = ops.pool3_res_cp8(yc, res=zc if (residual and i > 0) else None)
    return head_tc(cache, model, zc, a)


def block_operands(cache, name, conv, fmt, dev, cin_pad, cout_pad, ring=False):
    """(packed weights, bias) of one CNN block with its channel counts zero-padded to (cin_pad, cout_pad)."""
    def build():
        w, b = conv.weight.detach().float(), conv.bias.detach().float()
        Cout, Cin = w.shape[0], w.shape[1]
        if (Cout, Cin) != (cout_pad, cin_pad):
            w2 = torch.zeros(cout_pad, cin_pad, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
            w2[:Cout, :Cin] = w
            b2 = torch.zeros(cout_pad, dtype=b.dtype, device=b.device)
            b2[:Cout] = b
            w, b = w2, b2
        return ops.conv_tc_pack(w, dev, fmt, ring=ring), b.contiguous().to(dev)
    return cache.get(f'{name}:wtcpad{fmt}:{cin_pad}:{cout_pad}:{int(ring)}', [conv.weight, conv.bias], build)


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


# ------------------------------------------------------------------------------------------------------------------
# U-Net family on the tcgen05 path (eval mode: BatchNorm folded into the packed weights / bias)
LEVEL_PF = 8
FUSED_THIN_TAIL = True        # CNN-family head: pool13 + conv3 + conv4.* in one launch (False: pool kernel, tcgen05 conv3, tail kernel)


def level_geometry(T0, F0, n_levels=5):
    """(T, F, pitch) of every U-Net level: MaxPool(2,2) floor sizes; pitch = multiple of 16 >= F + 8 (zero gap >= 7)."""
    geo, T, F = [], T0, F0
    for _ in range(n_levels):
        geo.append((T, F, (F + LEVEL_PF + 15) // 16 * 16))
        T, F = T // 2, F // 2
    return geo


def unet_tc_eligible(model, x):
    if model.precision not in TC_PRECISIONS or model.training:
        return False
    chans = [model.inc.double_conv[0].weight.shape[0]] + [getattr(model, f'down{i}')[1].double_conv[0].weight.shape[0] for i in (1, 2, 3, 4)]
    ups = [getattr(model, f'upconv{i}').double_conv[4].weight.shape[0] for i in (1, 2, 3)]
    ok = all(c % 8 == 0 for c in chans + ups) and x.shape[3] + LEVEL_PF <= 256 and x.shape[2] >= 32
    if not ok:
        note_path(model, 'this configuration runs on the fp32 CUDA-core kernels, not on the tensor cores (the tcgen05 U-Net path needs every '
                         f'level width to be a multiple of 8 channels: {chans + ups}, n_bins_in <= 248 and >= 32 frames)')
    return ok


def _folded_tc(cache, name, conv, bn, fmt, dev, cin_pad=None, J=0):
    """Packed operand(s) of conv (+ eval BatchNorm folded): list of (packed weights, bias, cout0, cout) per <=128 block.
    The output channels are padded to a multiple of 8 (zero weights / bias) so that every block ends on a chunk
    boundary; `cin_pad` appends zero input channels (when the producer padded its outputs the same way)."""
    params = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])

    def build():
        w, b = conv.weight.detach().float(), conv.bias.detach().float()
        if bn is not None:
            s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            w = w * s[:, None, None, None]
            b = (b - bn.running_mean) * s + bn.bias
        Cout, Cin = w.shape[0], w.shape[1]
        Cp = (Cout + 7) // 8 * 8
        Ci = Cin if cin_pad is None else cin_pad
        if Cp != Cout or Ci != Cin:
            w2 = torch.zeros(Cp, Ci, w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
            w2[:Cout, :Cin] = w
            b2 = torch.zeros(Cp, dtype=b.dtype, device=b.device)
            b2[:Cout] = b
            w, b = w2, b2
        out = []
        for c0 in range(0, Cp, 128):
            c1 = min(Cp, c0 + 128)
            out.append((ops.conv_tc_pack(w[c0:c1].contiguous(), dev, fmt, J), b[c0:c1].contiguous(), c0, c1 - c0))
        return out
    return cache.get(f'{name}:tcfold{fmt}:{cin_pad}:{J}', params, build)


def conv_bn_relu_tc(cache, name, conv, bn, src, dst, act=ops.ACT_RELU, a=0.0, split=None):
    """split = s: `dst` is an ops.split_cp8 buffer (s phase sets); co-block c0 starts at chunk c0/8 of every phase set.
    src may be a frame window (ops.frame_window): patch b = rows [b, b+T) of one shared frame-major plane."""
    k = tuple(conv.kernel_size)
    stream = dict(n_patches=src.B, patch_stride_rows=1, T=src.T) if getattr(src, 'streaming', False) else {}
    for wp, b, c0, c in _folded_tc(cache, name, conv, bn, src.fmt, src.buf.device):
        ops.conv_tc(src, wp, b, c, k, act, a, out=dst.channels(c0, c), split=split, **stream)
    return dst


def conv_even_valid_tc(cache, name, conv, src, act, a):
    """A VALID convolution with an EVEN kernel height (the PUnet's convP.0, (2,5) on the 4x13 bottleneck) on the tcgen05 kernel: it
    equals the odd (KH+1) x KW 'same' convolution whose first filter row is zero, restricted to output rows [0, T-KH] and columns
    [KW//2, F-1-KW//2].  src: CP8 -> fp32 NCHW [B, Cout, T-KH+1, F-KW+1].  (The fp32 direct kernel spent 5.3 ms per 400 patches on
    these 7 GFLOP: its 16-row output tiles hold 3 real rows.)"""
    Cout, Cin, KH, KW = conv.weight.shape
    assert KH % 2 == 0 and KW % 2 == 1 and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (0, 0)
    fmt, dev = src.fmt, src.buf.device

    def build():
        import types
        w = conv.weight.detach().float()
        w3 = torch.zeros(Cout, Cin, KH + 1, KW, dtype=w.dtype, device=w.device)
        w3[:, :, 1:, :] = w
        return types.SimpleNamespace(weight=w3, bias=conv.bias, kernel_size=(KH + 1, KW))
    shim = cache.get(f'{name}:oddw', [conv.weight, conv.bias], build)
    rows = src.T - KH + 1
    out = ops.CP8(src.B, (Cout + 7) // 8 * 8, rows, src.F, src.pitch, src.pf, src.pt, dev, fmt=fmt)
    for wp, b, c0, c in _folded_tc(cache, name + ':odd', shim, None, fmt, dev):
        ops.conv_tc(src, wp, b, c, (KH + 1, KW), act, a, out=out.channels(c0, c), rows=(0, rows))
    # un-pad: columns [KW//2, F-1-KW//2] of the 'same' result, read by the converter through a shifted left pad
    Fo = src.F - KW + 1
    y = torch.empty(src.B, Cout, rows, Fo, dtype=torch.float32, device=dev)
    _lib.call('cp8_to_nchw', out.ptr(), y, src.B, Cout, rows, Fo, out.pitch, out.pf + KW // 2, out.pt, fmt, out.ncs, _lib.stream_ptr())
    return y


def double_conv_tc(cache, name, dc, src, dst, scratch, split=None):
    seq = dc.double_conv
    conv_bn_relu_tc(cache, name + '.0', seq[0], seq[1], src, scratch)
    return conv_bn_relu_tc(cache, name + '.4', seq[4], seq[5], scratch, dst, split=split)


def unet_forward_tc(model, x, frames=None):
    """simple_u_net_largekernels / _doubleselfattn / _polyphony_classif_softmax, eval mode, tensor-core path.
    Skip connections are written straight into the first chunks of the decoder's concat buffers; the bilinear
    up-sampler fills the remaining chunks, so torch.cat never happens."""
    cache, a = model._cache, model.a_lrelu
    fmt = ops.fmt_of(model.precision)
    if frames is not None:
        plane, i0, B = frames
        C, T, F, dev = model.n_chan_input, 75, model.n_bins_in, plane.device
    else:
        B, C, T, F = x.shape
        dev = x.device
    geo = level_geometry(T, F)
    c = [model.inc.double_conv[4].weight.shape[0]] + [getattr(model, f'down{i}')[1].double_conv[4].weight.shape[0] for i in (1, 2, 3, 4)]
    up_out = [getattr(model, f'upconv{i}').double_conv[4].weight.shape[0] for i in (1, 2, 3, 4)]

    # Activation planes are allocated (and zeroed) once per (batch, shape) and reused by later forwards: kernels write real pixels only,
    # so the zero borders survive, and re-zeroing ~25 buffers per forward cost 9 % of the Unet:M step.  The n-th request of a forward
    # always gets the n-th buffer, so two live buffers never alias.
    pools = model.__dict__.setdefault('_plane_pools', {})
    key = (B, T, F, fmt, str(dev))
    if key not in pools:
        while len(pools) >= 3:                          # e.g. the full batch and the ragged last batch of a recording
            pools.pop(next(iter(pools)))
        pools[key] = {}
    pool = pools[key]
    counter = [0]

    def buf(level, ch):
        Tl, Fl, P = geo[level]
        counter[0] += 1
        k = (counter[0], level, ch)
        if k not in pool:
            pool[k] = ops.CP8(B, ch, Tl, Fl, P, LEVEL_PF, 1, dev, fmt=fmt)
        return pool[k]
    if frames is not None:
        z = ops.frame_window(plane, i0, B, C, T, F, geo[0][2], LEVEL_PF, 1, fmt)      # LayerNorm was evaluated once per frame
    else:
        z = ops.nchw_to_cp8(ops.layernorm_cf(x, model.layernorm.weight, model.layernorm.bias, model.layernorm.eps), out=buf(0, C))
    # concat buffers of the decoder: level 3 (x4|up x5), level 2 (x3|up u1), level 1 (x2|up u2), level 0 (x1|up u3)
    cat = {3: buf(3, c[3] + c[4]), 2: buf(2, c[2] + up_out[0]), 1: buf(1, c[1] + up_out[1]), 0: buf(0, c[0] + up_out[2])}
    skips = {lv: cat[lv].channels(0, c[lv]) for lv in (0, 1, 2, 3)}
    # encoder
    double_conv_tc(cache, 'inc', model.inc, z, skips[0], buf(0, model.inc.double_conv[0].weight.shape[0]))
    cur = skips[0]
    for lv in (1, 2, 3, 4):
        dc = getattr(model, f'down{lv}')[1]
        pooled = ops.maxpool2x2_cp8(cur, buf(lv, c[lv - 1]))
        dst = skips[lv] if lv < 4 else buf(4, c[4])
        cur = double_conv_tc(cache, f'down{lv}', dc, pooled, dst, buf(lv, dc.double_conv[0].weight.shape[0]))
    x5 = cur
    afmt = None if fmt == ops.FMT_F16X3 else fmt        # split precision: the encoder layers run their fp32 sequence
    if hasattr(model, 'attention1'):
        t5 = model.attention2.run(model.attention1.run(ops.cp8_to_nchw(x5), afmt), afmt)
        x5 = ops.nchw_to_cp8(t5, pitch=geo[4][2], pf=LEVEL_PF, pt=1, fmt=fmt)
    if getattr(model, 'lstm_depth', 0) > 0:
        # BLUnet: BLSTM over time at the bottleneck (fp32 kernels; < 1 % of the model's FLOPs)
        x5 = ops.nchw_to_cp8(model.lstm5.run(ops.cp8_to_nchw(x5)), pitch=geo[4][2], pf=LEVEL_PF, pt=1, fmt=fmt)
    if getattr(model, 'lstm_depth', 0) > 1:
        ops.nchw_to_cp8(model.lstm4.run(ops.cp8_to_nchw(skips[3])), out=skips[3], fmt=fmt)
    if hasattr(model, 'attention3'):
        # SAUSnet: the lowest skip connection passes two encoder layers too (x5 above was computed from the un-attended x4)
        t4 = model.attention4.run(model.attention3.run(ops.cp8_to_nchw(skips[3]), afmt), afmt)
        ops.nchw_to_cp8(t4, out=skips[3], fmt=fmt)
    # decoder
    low = x5
    split = head_split_factor(model, F) if up_out[3] % 8 == 0 else None
    for i, lv in enumerate((3, 2, 1, 0)):
        dc = getattr(model, f'upconv{i + 1}')
        ops.upsample2x_cp8(low, cat[lv].channels(c[lv], low.C))
        if lv == 0 and split:
            # the last convolution feeds the head's stride-(1,3) conv2: write its output phase-split
            counter[0] += 1
            k = (counter[0], 'split', up_out[i])
            if k not in pool:
                pool[k] = ops.split_cp8(B, up_out[i], T, F, split, dev, fmt)
            low = double_conv_tc(cache, f'upconv{i + 1}', dc, cat[lv], pool[k], buf(lv, dc.double_conv[0].weight.shape[0]), split=split)
        else:
            low = double_conv_tc(cache, f'upconv{i + 1}', dc, cat[lv], buf(lv, up_out[i]), buf(lv, dc.double_conv[0].weight.shape[0]))
    y = head_tc(cache, model, low, a, split=split, pool=pool)
    return y, x5


_PE_TABLES = {}


def sinusoidal_pe(n, E, device):
    """The reference's sinusoidal table (unet_cnns.py:118-125), built once per (length, width, device): a constant, and a host->device
    copy per call would also be illegal inside a CUDA-graph capture of the training step."""
    key = (int(n), int(E), str(device))
    if key not in _PE_TABLES:
        _PE_TABLES[key] = _sinusoidal_pe(n, E, device)
    return _PE_TABLES[key]


def _sinusoidal_pe(n, E, device):
    position = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, E, 2, dtype=torch.float32) * (-math.log(10000.0) / E))
    pe = torch.zeros(n, E)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.to(device)
