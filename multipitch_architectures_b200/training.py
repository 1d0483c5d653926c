"""Training path of the CNN family (basic_cnn_segm_sigmoid / deep_cnn_segm_sigmoid): forward with saved activations,
hand-written backward, fused BCE and AdamW — all libmpa kernels (fp32).

Reference loop being replaced (experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py:315-329):
    y_pred = model(X); loss = BCELoss(y_pred, y); optimizer.zero_grad(); loss.backward(); optimizer.step()

Two ways in:
  * drop-in: `model.train(); y = model(x); loss = torch.nn.BCELoss()(y, t); loss.backward(); torch.optim.AdamW(...).step()`
    — `model.forward` routes through `CnnTrainFunction` (a torch.autograd.Function whose backward is the code below);
  * fused: `TrainStep(model, lr=...)(x, t)` — BCE forward/backward kernel, backward, (optional) gradient all-reduce over
    NCCL on ONE flat buffer, fused AdamW; no torch operator touches an activation or a gradient.
With `precision='bf16'` the convolutions run on the tensor cores (TcConv) and, for the models without the residual path, activations and
gradients stay in the 16-bit CP8 planes between the block convolutions (`_cp8_resident`: pool + dropout forward / backward and the bias
gradients are train_cp8.cu kernels; bit-identical to crossing the nchw<->CP8 converters around the fp32 kernels).
Dropout uses a Philox stream (seed, per-site offset); the reference's torch RNG stream cannot be reproduced, so parity
tests run with p_dropout = 0."""
import torch

from . import _lib, ops
from .libdl.nn_models import _exec
from ._lib import call, stream_ptr


# Device-resident step counter (int64[1]) while a training step is being captured into a CUDA graph: dropout offsets are then formed on
# the device as step * 256 + site instead of being baked into the kernel arguments.
_step_dev = None
_step_mul = 256        # offset = step * _step_mul + site (256 sites per step in the U-Net tape, 64 in the CNN path)


def _dropout(x, p, seed, offset):
    if p <= 0.0:
        return x
    out = torch.empty_like(x)
    if _step_dev is not None:
        call('dropout_dev_f32', x, out, _lib.i64(x.numel()), float(p), ctypes_u64(seed), ctypes_u64(offset), _step_dev, ctypes_u64(_step_mul),
             stream_ptr())
        return out
    call('dropout_f32', x, out, _lib.i64(x.numel()), float(p), ctypes_u64(seed), ctypes_u64(offset), stream_ptr())
    return out


def ctypes_u64(v):
    import ctypes
    return ctypes.c_ulonglong(int(v) & 0xFFFFFFFFFFFFFFFF)


def _is_rows_conv(conv, H):
    """Full-height VALID convolution with one output row (conv3, 75x1 on a 75-frame patch): served by the GEMM kernels of conv_rows.cu."""
    return (tuple(conv.kernel_size) == (H, 1) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (0, 0)
            and tuple(conv.dilation) == (1, 1))


def _conv_fwd(conv, x, act, a):
    w = conv.weight
    if _is_rows_conv(conv, x.shape[2]):
        B, Cin, H, W = x.shape
        out = torch.empty(B, w.shape[0], 1, W, dtype=torch.float32, device=x.device)
        ws_bytes = _lib.lib().mpa_conv_rows_fwd_workspace(B, Cin, H, W, w.shape[0])
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
        call('conv_rows_fwd_f32', x, w.detach().contiguous(), conv.bias, out, B, Cin, H, W, w.shape[0], act, float(a), ws, _lib.usize(ws_bytes),
             stream_ptr())
        return out
    wp = ops.pack_conv_weight(w)
    return ops.conv2d(x, wp, conv.bias, w.shape[0], tuple(w.shape[2:]), tuple(conv.stride), tuple(conv.padding), act, a)


def _dgrad(conv, g, in_shape):
    """Gradient wrt the conv input.  Stride-1 layers: forward kernel with swapped/flipped weights; else gather kernel."""
    w = conv.weight
    Cout, Cin, KH, KW = w.shape
    B, _, H, W = in_shape
    if tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (KH // 2, KW // 2) and KH % 2 == 1 and KW % 2 == 1:
        # 'same' convolutions only: for a VALID one (conv3, 75x1) the flipped full correlation multiplies 75x more zeros than data
        wt = ops.pack_conv_weight(w.detach().permute(1, 0, 2, 3).flip(2, 3).contiguous())
        return ops.conv2d(g, wt, None, Cin, (KH, KW), (1, 1), (KH - 1 - conv.padding[0], KW - 1 - conv.padding[1]))
    gi = torch.empty(in_shape, dtype=torch.float32, device=g.device)
    if _is_rows_conv(conv, H):
        call('conv_rows_dgrad_f32', g, w.detach().contiguous(), gi, B, Cin, H, W, Cout, stream_ptr())
        return gi
    call('conv2d_dgrad_f32', g, w.detach().contiguous(), gi, B, Cin, H, W, Cout, KH, KW, conv.stride[0], conv.stride[1],
         conv.padding[0], conv.padding[1], stream_ptr())
    return gi


def _wgrad(conv, x, g, gw, gb):
    Cout, Cin, KH, KW = conv.weight.shape
    B, _, H, W = x.shape
    if _is_rows_conv(conv, H):
        call('conv_rows_wgrad_f32', x, g, gw, B, Cin, H, W, Cout, stream_ptr())
        ops.channel_sum(g, out=gb)
        return
    call('conv2d_wgrad_f32', x, g, gw, gb, B, Cin, H, W, Cout, KH, KW, conv.stride[0], conv.stride[1], conv.padding[0],
         conv.padding[1], stream_ptr())


WGRAD_SIDE = True          # test knob: False = weight gradients on the main stream, in front of the data gradient of their layer
PACK_MULTI = True          # test knob: False = one weight-packing launch per convolution and direction
ROWS_TC = True             # test knob: False = conv3 of the bf16 models on the fp32 strided-GEMM kernels (conv_rows.cu)


def _rows_tc_eligible(model, conv, x):
    """The head's full-height 75x1 convolution of a bf16 model with more than 10 output channels: forward and both gradients as tcgen05
    GEMMs.  Thinner ones (CNN:XS 20 -> 10, DRCNN 30 -> 10) are HBM-bound matrix-vector products and stay on the conv_rows_thin_* kernels (the
    tensor-core form was measured at batch 256: 3.39 -> 3.48 ms per CNN:XS step — two operand-chunking passes over x outweigh the GEMMs)."""
    return (ROWS_TC and model is not None and getattr(model, 'precision', 'fp32') == 'bf16' and _is_rows_conv(conv, x.shape[2])
            and conv.weight.shape[0] > 10 and x.shape[3] % 8 == 0 and conv.bias is not None)


def _rows_tc_forward(conv, x, act, a):
    """Y[b][co][w] = act(sum_k W[co][k] x[b][k][w] + bias[co]), k = (ci, frame): tokens (b, w) x K on the tensor cores (bf16 operands, fp32
    accumulate, K slices in atomics), then bias + activation in place."""
    fm = ops.FMT_BF16
    B, Cin, H, W = x.shape
    Cout, K, Mt = conv.weight.shape[0], Cin * H, B * W
    xf = ops.strided_chunks(x, Mt, K, 256, fm, (W, K * W, 1), (K, 0, W))
    wc = ops.gemm_tc_chunks(conv.weight.detach().reshape(Cout, K), 128, fm)
    y = torch.zeros(B, Cout, 1, W, dtype=torch.float32, device=x.device)
    ops.gemm_tc_ex(xf, wc, None, Mt, Cout, K, False, fm, y=y, y_m=(W, Cout * W, 1), y_n=(0, 0, W), y_zeroed=True)
    call('bias_act_f32', y, conv.bias, y, B, Cout, W, act, float(a), stream_ptr())
    return y


def _rows_tc_backward(conv, x, g, gw, gb, need_dx):
    """g: gradient wrt the convolution output (activation already differentiated away).  Weight gradient [Cout][K] = g^T x over the B*W
    tokens, data gradient [b][k][w] = W^T g written straight into NCHW."""
    fm = ops.FMT_BF16
    B, Cin, H, W = x.shape
    Cout, K, Mt = conv.weight.shape[0], Cin * H, B * W
    def wgrad():
        gt = ops.strided_chunks(g, Cout, Mt, 256, fm, (Cout, 0, W), (W, Cout * W, 1))
        xt = ops.strided_chunks(x, K, Mt, 128, fm, (K, 0, W), (W, K * W, 1))
        ops.gemm_tc_ex(gt, xt, None, Cout, K, Mt, False, fm, y=gw.view(Cout, K))
    TcConv.wgrad_async(g.device, wgrad, keep=(g, x))
    ops.channel_sum(g, out=gb)
    if not need_dx:
        return None
    gf = ops.strided_chunks(g, Mt, Cout, 128, fm, (W, Cout * W, 1), (Cout, 0, W))
    wt = ops.gemm_tc_chunks(conv.weight.detach().reshape(Cout, K), 256, fm, True)
    gi = torch.empty(B, Cin, H, W, dtype=torch.float32, device=x.device)
    ops.gemm_tc_ex(wt, gf, None, K, Mt, Cout, False, fm, y=gi, y_m=(0, 0, W), y_n=(W, K * W, 1))
    return gi


def _add(a, b):
    out = torch.empty_like(a)
    call('add_f32', a, b, out, _lib.i64(a.numel()), stream_ptr())
    return out


def _act_bwd(out, g, act, a=0.0):
    gi = torch.empty_like(g)
    call('act_bwd_f32', out, g, gi, _lib.i64(g.numel()), act, float(a), stream_ptr())
    return gi


FUSE_POOL_DROPOUT = True      # test knob: False = separate pool / dropout (/ add) kernels; the results are identical


def _step_args():
    return _step_dev, ctypes_u64(_step_mul if _step_dev is not None else 0)


def _pool_dropout(act_t, k, res, p, seed, offset):
    """dropout(maxpool_time(act_t, k)) (+ res): one kernel when dropout is on (same mask as _dropout(., p, seed, offset))."""
    B, C, T, F = act_t.shape
    if p > 0.0 and FUSE_POOL_DROPOUT and F % 4 == 0:
        out = torch.empty_like(act_t)
        sd, sm = _step_args()
        call('maxpool_time_dropout_f32', act_t, res, out, B, C, T, F, k, float(p), ctypes_u64(seed), ctypes_u64(offset), sd, sm, stream_ptr())
        return out
    d = _dropout(ops.maxpool_time(act_t, k), p, seed, offset)
    return _add(d, res) if res is not None else d


def _pool_bwd_dropout(a_in, g_out, k, act, a, p, seed, offset):
    """Gradient wrt the pool input of dropout(maxpool_time(.)): the dropout mask is applied to g_out on the fly."""
    if p > 0.0 and FUSE_POOL_DROPOUT and k in (3, 13) and a_in.shape[2] > k // 2:
        B, C, T, F = a_in.shape
        ga = torch.empty_like(a_in)
        sd, sm = _step_args()
        call('maxpool_time_bwd_dropout_f32', a_in, g_out, ga, B, C, T, F, k, act, float(a), float(p), ctypes_u64(seed), ctypes_u64(offset), sd, sm,
             stream_ptr())
        return ga
    return _pool_bwd(a_in, _dropout(g_out, p, seed, offset), k, act, a)


def _pool_bwd(a_in, g_pool, k, act, a):
    B, C, T, F = a_in.shape
    ga = torch.empty_like(a_in)
    call('maxpool_time_bwd_f32', a_in, g_pool, ga, B, C, T, F, k, act, float(a), stream_ptr())
    return ga


class TcConv:
    """bf16 tensor-core forward / backward of a stride-1 'same' nn.Conv2d inside the fp32 NCHW training graph:
    forward = mpa_conv_tc_f16 on CP8 planes, data gradient = the same kernel with the transposed / flipped weights, weight gradient =
    mpa_conv_wgrad_tc, bias gradient = a channel reduction.  Activations and gradients cross the boundary through nchw<->CP8 converters
    (16-bit operands, fp32 accumulation; the optimiser state and every element-wise stage stay fp32)."""
    PF, PT = 8, 1
    _pool = {}

    @staticmethod
    def eligible(model, conv, F):
        KH, KW = conv.kernel_size
        return (getattr(model, 'precision', 'fp32') == 'bf16' and tuple(conv.stride) == (1, 1) and KH % 2 == 1 and KW % 2 == 1 and 3 <= KW <= 15 and KH >= 3
                and tuple(conv.padding) == (KH // 2, KW // 2) and F + TcConv.PF + 15 < 272 and conv.bias is not None)

    _scope = None           # owner of the pooled buffers (a TrainStep); None = allocate per call

    @classmethod
    def scope(cls, owner, refresh=True):
        """Context manager: inside it `_buf` hands out buffers pooled under `owner` (a fused train step, which runs forward and backward
        atomically, so reusing the SAVED activation planes across steps is safe).  Outside any scope — the autograd bridge, where a
        second forward may run before the first backward (gradient accumulation, two losses, two same-shape models) — every call gets
        fresh planes, because the planes are what the backward reads."""
        import contextlib

        @contextlib.contextmanager
        def cm():
            prev, cls._scope = cls._scope, id(owner)
            plan = cls._plans.get(id(owner))
            if plan is None:
                plan = cls._plans[id(owner)] = ops.PackPlan()
            prev_plan, ops._PACK_PLAN = ops._PACK_PLAN, plan
            if refresh:
                plan.fresh = False
                if PACK_MULTI:
                    plan.refresh()       # every packed operand of the step in one launch (from the second step on)
            else:
                plan.fresh = plan.table is not None and PACK_MULTI      # second half of a step: the operands were refreshed by the first
            try:
                yield
                cls.join()
                if PACK_MULTI:
                    plan.build()
            finally:
                cls._scope, ops._PACK_PLAN = prev, prev_plan
                plan.fresh = False
        return cm()

    _plans = {}

    @classmethod
    def release(cls, owner):
        """Drop every pooled buffer of `owner`."""
        for k in [k for k in cls._pool if k[0] == id(owner)]:
            del cls._pool[k]
        cls._plans.pop(id(owner), None)

    @classmethod
    def _buf(cls, tag, B, C, T, F, dev, fmt):
        """CP8 scratch reused across the steps of one owner (zero gap columns are never written, so they stay zero)."""
        pitch = (F + cls.PF + 15) // 16 * 16
        if cls._scope is None:
            return ops.CP8(B, C, T, F, pitch, cls.PF, cls.PT, dev, fmt=fmt)
        key = (cls._scope, tag, B, C, T, F, str(dev), fmt)
        b = cls._pool.get(key)
        if b is None:
            b = ops.CP8(B, C, T, F, pitch, cls.PF, cls.PT, dev, fmt=fmt)
            cls._pool[key] = b
        return b

    # Weight gradients on a side stream: inside a fused train step (scope set) the weight gradient of a convolution is independent of
    # everything that follows it in the backward (it reads the saved input planes and the dy planes of its own layer — pooled per layer,
    # not reused within the step — and writes its own slice of the flat gradient buffer), so it is issued on a second stream and joined
    # at the end of the backward: the small, latency-bound layers of the deep U-Net levels then run side by side with the data-gradient
    # chain instead of in front of it.  Under CUDA-graph capture the fork / join become parallel branches of the graph.
    _side_streams, _side_dirty = {}, False

    _keep = []

    @classmethod
    def wgrad_async(cls, dev, fn, keep=()):
        """keep: tensors allocated on the main stream that `fn` reads — held until the join, so that the allocator cannot hand their memory
        to later main-stream work while the side stream still reads it."""
        if not WGRAD_SIDE or cls._scope is None:
            return fn()
        cls._keep.extend(keep)
        side = cls._side_streams.get(str(dev))
        if side is None:
            side = cls._side_streams[str(dev)] = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        cls._side_dirty = True

    @classmethod
    def join(cls):
        """The current stream waits for the weight gradients issued so far."""
        if cls._side_dirty:
            for side in cls._side_streams.values():
                torch.cuda.current_stream().wait_stream(side)
            cls._side_dirty = False
        cls._keep.clear()

    @classmethod
    def _raw(cls, tag, nbytes, zero, dev):
        """Byte scratch pooled like `_buf` (zero: zero-filled when it is created; the users never dirty the parts that must stay zero)."""
        if cls._scope is None:
            return (torch.zeros if zero else torch.empty)(nbytes, dtype=torch.uint8, device=dev)
        key = (cls._scope, tag, nbytes, str(dev), bool(zero))
        b = cls._pool.get(key)
        if b is None:
            b = cls._pool[key] = (torch.zeros if zero else torch.empty)(nbytes, dtype=torch.uint8, device=dev)
        return b

    @staticmethod
    def _pad8(v, n):
        """bias (or zeros) padded to n entries: the padded output channels of a block carry zero weights and zero bias."""
        if v.numel() == n:
            return v.detach()
        out = torch.zeros(n, dtype=torch.float32, device=v.device)
        out[:v.numel()] = v.detach()
        return out

    _zeros = {}

    @classmethod
    def _zero_bias(cls, dev):
        """128 fp32 zeros (the bias operand of a data-gradient convolution), allocated once per device."""
        z = cls._zeros.get(str(dev))
        if z is None:
            z = cls._zeros[str(dev)] = torch.zeros(128, dtype=torch.float32, device=dev)
        return z

    @classmethod
    def forward(cls, tag, conv, x, act, a):
        fmt = ops.FMT_BF16
        B, Cin, T, F = x.shape
        xc = ops.nchw_to_cp8(x, out=cls._buf(tag + ':x', B, Cin, T, F, x.device, fmt), fmt=fmt)
        return ops.cp8_to_nchw(cls.forward_cp8(tag, conv, xc, act, a)), xc

    @classmethod
    def forward_cp8(cls, tag, conv, xc, act, a):
        """CP8 in -> CP8 out (pooled buffer `tag:y`)."""
        fmt = ops.FMT_BF16
        B, Cin, T, F = xc.B, xc.C, xc.T, xc.F
        Cout, _, KH, KW = conv.weight.shape
        yc = cls._buf(tag + ':y', B, Cout, T, F, xc.buf.device, fmt)
        for c0 in range(0, Cout, 128):
            c = min(128, Cout - c0)
            cp = (c + 7) // 8 * 8                        # whole chunks: the coalesced epilogue needs Cout % 8 == 0
            wp = ops.conv_tc_pack_dev(conv.weight, Cin, cp, (KH, KW), fmt, False, Cout, c0)
            ops.conv_tc(xc, wp, cls._pad8(conv.bias[c0:c0 + c], cp), cp, (KH, KW), act, a, out=yc.channels(c0, cp))
        return yc

    @classmethod
    def backward(cls, tag, conv, xc, g, gw, gb, need_dx):
        """g: fp32 gradient wrt the convolution output (before the activation has been differentiated away by the caller)."""
        fmt = ops.FMT_BF16
        B, Cout, T, F = g.shape
        gc = ops.nchw_to_cp8(g, out=cls._buf(tag + ':g', B, Cout, T, F, g.device, fmt), fmt=fmt)
        ops.channel_sum(g, out=gb)
        gxc = cls.backward_cp8(tag, conv, xc, gc, gw, None, need_dx)
        return ops.cp8_to_nchw(gxc) if need_dx else None

    @classmethod
    def backward_cp8(cls, tag, conv, xc, gc, gw, gb, need_dx):
        """gc: CP8 gradient wrt the convolution output; -> CP8 gradient wrt its input (pooled buffer `tag:gx`).  gb (optional): bias
        gradient summed from the 16-bit planes."""
        fmt = ops.FMT_BF16
        B, Cout, T, F = gc.B, gc.C, gc.T, gc.F
        _, Cin, KH, KW = conv.weight.shape
        dev = gc.buf.device
        cls.wgrad_async(dev, lambda: ops.conv_wgrad_tc(xc, gc, gw, (KH, KW)))
        if gb is not None:
            ops.channel_sum_cp8(gc, out=gb)
        if not need_dx:
            return None
        gxc = cls._buf(tag + ':gx', B, Cin, T, F, dev, fmt)
        zb = cls._zero_bias(dev)
        for c0 in range(0, Cin, 128):
            c = min(128, Cin - c0)
            cp = (c + 7) // 8 * 8
            wp = ops.conv_tc_pack_dev(conv.weight, Cout, cp, (KH, KW), fmt, True, Cin, c0)
            ops.conv_tc(gc, wp, zb[:cp], cp, (KH, KW), ops.ACT_NONE, 0.0, out=gxc.channels(c0, cp))
        return gxc


def _tc_s3_eligible(model, conv, F):
    """The head's 3x3 / stride (1,3) / pad (1,0) convolution (basic_cnns.py:391) = the stride-1 'same' 3x3 convolution sampled at
    columns 1, 4, 7, ... ; its backward is the stride-1 backward of the zero-inserted output gradient."""
    return (getattr(model, 'precision', 'fp32') == 'bf16' and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 3)
            and tuple(conv.padding) == (1, 0) and F % 3 == 0 and F + TcConv.PF + 15 < 272 and conv.bias is not None)


def _tc_s3_forward(tag, conv, x, act, a):
    """x: fp32 NCHW or the CP8 planes a preceding CP8-resident block left behind."""
    fmt = ops.FMT_BF16
    if isinstance(x, ops.CP8):
        xc = x
    else:
        B, Cin, T, F = x.shape
        xc = ops.nchw_to_cp8(x, out=TcConv._buf(tag + ':x', B, Cin, T, F, x.device, fmt), fmt=fmt)
    B, Cin, T, F = xc.B, xc.C, xc.T, xc.F
    Cout = conv.weight.shape[0]
    yc = ops.compact_cp8(B, Cout, T, F // 3, xc.buf.device, fmt)
    for c0 in range(0, Cout, 128):
        c = min(128, Cout - c0)
        cp = (c + 7) // 8 * 8
        wp = ops.conv_tc_pack_dev(conv.weight, Cin, cp, (3, 3), fmt, False, Cout, c0)
        ops.conv_tc(xc, wp, TcConv._pad8(conv.bias[c0:c0 + c], cp), cp, (3, 3), act, a, subsample=(3, 1), out=yc.channels(c0, cp))
    return ops.cp8_to_nchw(yc), xc


def _tc_s3_backward(tag, conv, xc, g, gw, gb, keep_cp8=False):
    fmt = ops.FMT_BF16
    B, Cout, T, Fo = g.shape
    Cin, F = conv.weight.shape[1], xc.F
    gc = TcConv._buf(tag + ':g3', B, Cout, T, F, g.device, fmt)
    call('nchw_to_cp8_strided', g, gc.ptr(), B, Cout, T, Fo, F, 3, 1, gc.pitch, gc.pf, gc.pt, fmt, gc.ncs, stream_ptr())
    TcConv.wgrad_async(g.device, lambda: ops.conv_wgrad_tc(xc, gc, gw, (3, 3)))
    ops.channel_sum(g, out=gb)
    gxc = TcConv._buf(tag + ':gx', B, Cin, T, F, g.device, fmt)
    zb = TcConv._zero_bias(g.device)
    for c0 in range(0, Cin, 128):
        c = min(128, Cin - c0)
        cp = (c + 7) // 8 * 8
        wp = ops.conv_tc_pack_dev(conv.weight, Cout, cp, (3, 3), fmt, True, Cin, c0)
        ops.conv_tc(gc, wp, zb[:cp], cp, (3, 3), ops.ACT_NONE, 0.0, out=gxc.channels(c0, cp))
    return gxc if keep_cp8 else ops.cp8_to_nchw(gxc)


# Phase-split form of the head's stride-(1,3) conv2 in TRAINING (the inference path uses it): test knob, False = the stride-1 3x3 convolution on
# full-width / zero-inserted planes.  It pays only together with wgrad_tc_kernel's wide-N mode (KH x 1 filters: the input chunks of a group are
# the N groups of one MMA); with one N = 16 MMA per chunk the weight gradient alone lost more than the forward gained (SAUnet:L 298 -> 534 us).
S3_SPLIT = True


def _s3_split_eligible(model, conv, F):
    """CP8-resident training paths: the producer of conv2's input writes phase-split planes (bin f -> phase f % 3, column f / 3), conv2 runs
    as the stride-1 3x1 convolution over 3 * C0p channels of width F / 3 — 3x fewer MMAs forward, no zero-inserted gradient planes backward,
    its data gradient comes out phase-split for the producer's backward."""
    return S3_SPLIT and _tc_s3_eligible(model, conv, F) and (F // 3 + TcConv.PF + 15) // 16 * 16 <= 256


def _split_buf(tag, B, C, T, F, s, dev, fmt):
    """Phase-split planes (ops.split_cp8) pooled like TcConv._buf."""
    if TcConv._scope is None:
        return ops.split_cp8(B, C, T, F, s, dev, fmt)
    key = (TcConv._scope, tag, 'split', B, C, T, F, s, str(dev), fmt)
    b = TcConv._pool.get(key)
    if b is None:
        b = TcConv._pool[key] = ops.split_cp8(B, C, T, F, s, dev, fmt)
    return b


def _tc_s3_forward_split(tag, conv, xs, act, a):
    """xs: phase-split planes of conv2's input (3 * C0p channels, width F / 3) -> (fp32 NCHW activation [B, Cout, T, F/3], xs)."""
    fmt = xs.fmt
    Cout, C0 = conv.weight.shape[0], conv.weight.shape[1]
    C0p, C1p = xs.C // 3, (Cout + 7) // 8 * 8
    J = _exec._split_conv2_J(C0p, C1p)
    yc = ops.compact_cp8(xs.B, Cout, xs.T, xs.F, xs.buf.device, fmt)
    for c0 in range(0, Cout, 128):
        c = min(128, Cout - c0)
        cp = (c + 7) // 8 * 8
        wp = ops.conv_tc_pack_dev(conv.weight, 3 * C0p, cp, (3, 1), fmt, False, Cout, c0, J=J, split=3, C0=C0)
        ops.conv_tc(xs, wp, TcConv._pad8(conv.bias[c0:c0 + c], cp), cp, (3, 1), act, a, subsample=(1, 0), out=yc.channels(c0, cp), J=J)
    return ops.cp8_to_nchw(yc), xs


def _tc_s3_backward_split(tag, conv, xs, g, gw, gb):
    """g: fp32 gradient wrt conv2's output [B, Cout, T, F/3] -> phase-split planes of the gradient wrt its input; gw / gb overwritten."""
    fmt = xs.fmt
    B, Cout, T, Fo = g.shape
    C0, C0p, dev = conv.weight.shape[1], xs.C // 3, g.device
    gc = ops.nchw_to_cp8(g, out=TcConv._buf(tag + ':g3s', B, Cout, T, Fo, dev, fmt), fmt=fmt)
    gwp = TcConv._raw(tag + ':gwsplit', Cout * 3 * C0p * 3 * 4, False, dev).view(torch.float32).view(Cout, 3 * C0p, 3, 1)
    ops.conv_wgrad_tc(xs, gc, gwp, (3, 1))
    gw.copy_(gwp.view(Cout, 3, C0p, 3)[:, :, :C0, :].permute(0, 2, 3, 1))          # gw[co][ci][kh][ph] = gw'[co][ph * C0p + ci][kh]
    ops.channel_sum(g, out=gb)
    gxs = _split_buf(tag + ':gxs', B, C0p, T, 3 * Fo, 3, dev, fmt)
    zb = TcConv._zero_bias(dev)
    for c0 in range(0, 3 * C0p, 128):
        cp = min(128, 3 * C0p - c0)
        wp = ops.conv_tc_pack_dev(conv.weight, Cout, cp, (3, 1), fmt, True, 3 * C0p, c0, split=3, C0=C0)
        ops.conv_tc(gc, wp, zb[:cp], cp, (3, 1), ops.ACT_NONE, 0.0, out=gxs.channels(c0, cp))
    return gxs


CP8_RESIDENT = True          # test knob: False = every block crosses nchw<->CP8 converters and pools in fp32 NCHW (identical results)
LN_PIXEL = True          # LayerNorm -> CP8 and its parameter gradient on the pixel-per-thread kernels (False: row kernels)


def _cp8_resident(model, blocks, F):
    """bf16 mode, no residual path: activations and gradients stay in the CP8 planes between the convolutions of the blocks
    (pool + dropout forward / backward and the bias gradients run on the planes, train_cp8.cu)."""
    return (CP8_RESIDENT and not getattr(model, 'residual', False) and F % 4 == 0 and all(TcConv.eligible(model, c, F) for _, c in blocks)
            and _tc_s3_eligible(model, model.conv2[0], F))


def _cnn_train_forward_cp8(model, blocks, x, sv, site, drop):
    a, p, seed = model.a_lrelu, sv['p'], sv['seed']
    fmt = ops.FMT_BF16
    B, C0, T, F = x.shape
    z = x           # (device of the buffers below)
    ln = model.layernorm
    # LayerNorm writes the first convolution's input planes directly (no fp32 copy of the normalised patch, no converter pass)
    # pixel-per-thread LayerNorm kernels (train_cp8.cu): (mean, rstd) of every row saved for the parameter gradient.  LN_PIXEL = False keeps
    # the row kernels, whose fp32 sums run in the order of the fp32 NCHW path (bit-for-bit tape comparisons in the tests)
    sv['ln_stats'] = torch.empty(B * T, 2, dtype=torch.float32, device=x.device) if (F <= 256 and LN_PIXEL) else None
    xc = ops.layernorm_cf_cp8(x, ln.weight, ln.bias, ln.eps, TcConv._buf(blocks[0][0] + ':x', B, C0, T, F, x.device, fmt), stats=sv['ln_stats'],
                              pixel=LN_PIXEL)
    sd, sm = _step_args()
    split = 3 if _s3_split_eligible(model, model.conv2[0], F) else 0
    for bi, (name, conv) in enumerate(blocks):
        yc = TcConv.forward_cp8(name, conv, xc, ops.ACT_LRELU, a)
        site[0] += 1
        sp = split if bi == len(blocks) - 1 else 0          # the last block hands over to conv2 in phase-split planes
        zc = _split_buf(name + ':zs', B, (yc.C + 7) // 8 * 8, T, F, sp, z.device, fmt) if sp else TcConv._buf(name + ':z', B, yc.C, T, F, z.device, fmt)
        call('pool3_dropout_split_cp8', yc.ptr(), zc.ptr(), B, yc.C, T, F, yc.pitch, yc.pf, yc.pt, fmt, float(p), ctypes_u64(seed),
             ctypes_u64(site[0]), sd, sm, sp, zc.pitch if sp else 0, stream_ptr())
        sv['blocks'].append((None, yc, xc))
        xc = zc
    sv['split'] = split
    if split:
        a2, x2c = _tc_s3_forward_split('conv2', model.conv2[0], xc, ops.ACT_LRELU, a)
    else:
        a2, x2c = _tc_s3_forward('conv2', model.conv2[0], xc, ops.ACT_LRELU, a)
    site[0] += 1
    d2 = _pool_dropout(a2, 13, None, p, seed, site[0])
    a3 = _rows_tc_forward(model.conv3[0], d2, ops.ACT_LRELU, a) if _rows_tc_eligible(model, model.conv3[0], d2) else \
        _conv_fwd(model.conv3[0], d2, ops.ACT_LRELU, a)
    d3 = drop(a3)
    a4 = _conv_fwd(model.conv4[0], d3, ops.ACT_LRELU, a)
    d4 = drop(a4)
    y = _conv_fwd(model.conv4[3], d4, ops.ACT_SIGMOID, 0.0)
    sv.update(z_head=None, a2=a2, d2=d2, a3=a3, d3=d3, a4=a4, d4=d4, y=y, x2c=x2c, cp8=True)
    return y, sv


def cnn_train_forward(model, x, seed=0, step=0):
    """-> (y_pred [B,1,T-74,72], saved).  Train mode: dropout active when model.p_dropout > 0."""
    a, p = model.a_lrelu, (model.p_dropout if model.training else 0.0)
    blocks = _exec.cnn_blocks(model)
    residual = getattr(model, 'residual', False)
    site = [step * 64]

    def drop(t):
        site[0] += 1
        return _dropout(t, p, seed, site[0])
    sv = {'x': x, 'blocks': [], 'p': p, 'seed': seed, 'site0': step * 64}
    if _cp8_resident(model, blocks, x.shape[3]) and x.shape[1] <= 8:
        return _cnn_train_forward_cp8(model, blocks, x, sv, site, drop)
    z = ops.layernorm_cf(x, model.layernorm.weight, model.layernorm.bias, model.layernorm.eps)
    for i, (name, conv) in enumerate(blocks):
        if TcConv.eligible(model, conv, z.shape[3]):
            act, xc = TcConv.forward(name, conv, z, ops.ACT_LRELU, a)
        else:
            act, xc = _conv_fwd(conv, z, ops.ACT_LRELU, a), None
        site[0] += 1
        sv['blocks'].append((z, act, xc))
        z = _pool_dropout(act, 3, z if (residual and i > 0) else None, p, seed, site[0])
    if _tc_s3_eligible(model, model.conv2[0], z.shape[3]):
        a2, x2c = _tc_s3_forward('conv2', model.conv2[0], z, ops.ACT_LRELU, a)
    else:
        a2, x2c = _conv_fwd(model.conv2[0], z, ops.ACT_LRELU, a), None
    site[0] += 1
    d2 = _pool_dropout(a2, 13, None, p, seed, site[0])
    a3 = _rows_tc_forward(model.conv3[0], d2, ops.ACT_LRELU, a) if _rows_tc_eligible(model, model.conv3[0], d2) else \
        _conv_fwd(model.conv3[0], d2, ops.ACT_LRELU, a)
    d3 = drop(a3)
    a4 = _conv_fwd(model.conv4[0], d3, ops.ACT_LRELU, a)
    d4 = drop(a4)
    y = _conv_fwd(model.conv4[3], d4, ops.ACT_SIGMOID, 0.0)
    sv.update(z_head=z, a2=a2, d2=d2, a3=a3, d3=d3, a4=a4, d4=d4, y=y, x2c=x2c)
    return y, sv


def cnn_train_backward(model, sv, g_y, grads):
    """grads: dict parameter-name -> preallocated fp32 tensor (overwritten)."""
    a, p, seed = model.a_lrelu, sv['p'], sv['seed']
    blocks = _exec.cnn_blocks(model)
    residual = getattr(model, 'residual', False)
    n_sites = len(blocks) + 3
    site = [sv['site0'] + n_sites + 1]

    def drop_bwd(g):
        site[0] -= 1
        return _dropout(g, p, seed, site[0])
    c43, c40, c3, c2 = model.conv4[3], model.conv4[0], model.conv3[0], model.conv2[0]
    g = _act_bwd(sv['y'], g_y.contiguous(), ops.ACT_SIGMOID)
    _wgrad(c43, sv['d4'], g, grads['conv4.3.weight'], grads['conv4.3.bias'])
    g = drop_bwd(_dgrad(c43, g, sv['d4'].shape))
    g = _act_bwd(sv['a4'], g, ops.ACT_LRELU, a)
    _wgrad(c40, sv['d3'], g, grads['conv4.0.weight'], grads['conv4.0.bias'])
    g = drop_bwd(_dgrad(c40, g, sv['d3'].shape))
    g = _act_bwd(sv['a3'], g, ops.ACT_LRELU, a)
    if _rows_tc_eligible(model, c3, sv['d2']):
        g_d2 = _rows_tc_backward(c3, sv['d2'], g, grads['conv3.0.weight'], grads['conv3.0.bias'], True)
    else:
        _wgrad(c3, sv['d2'], g, grads['conv3.0.weight'], grads['conv3.0.bias'])
        g_d2 = _dgrad(c3, g, sv['d2'].shape)
    site[0] -= 1
    g = _pool_bwd_dropout(sv['a2'], g_d2, 13, ops.ACT_LRELU, a, p, seed, site[0])
    if sv.get('cp8'):
        fmt = ops.FMT_BF16
        split = sv.get('split', 0)
        if split:
            gzc = _tc_s3_backward_split('conv2', c2, sv['x2c'], g, grads['conv2.0.weight'], grads['conv2.0.bias'])
        else:
            gzc = _tc_s3_backward('conv2', c2, sv['x2c'], g, grads['conv2.0.weight'], grads['conv2.0.bias'], keep_cp8=True)
        sd, sm = _step_args()
        for i in range(len(blocks) - 1, -1, -1):
            name, conv = blocks[i]
            _, yc, xc = sv['blocks'][i]
            site[0] -= 1
            gc = TcConv._buf(name + ':g', yc.B, yc.C, yc.T, yc.F, yc.buf.device, fmt)
            sp = split if i == len(blocks) - 1 else 0       # conv2's data gradient arrives in phase-split planes
            call('pool3_bwd_dropout_split_cp8', yc.ptr(), gzc.ptr(), gc.ptr(), yc.B, yc.C, yc.T, yc.F, yc.pitch, yc.pf, yc.pt, fmt, ops.ACT_LRELU,
                 float(a), float(p), ctypes_u64(seed), ctypes_u64(site[0]), sd, sm, sp, gzc.pitch if sp else 0, stream_ptr())
            gzc = TcConv.backward_cp8(name, conv, xc, gc, grads[f'{name}.0.weight'], grads[f'{name}.0.bias'], need_dx=True)
        ops.layernorm_cf_param_grad_cp8(sv['x'], gzc, grads['layernorm.weight'], grads['layernorm.bias'], model.layernorm.eps,
                                        stats=sv.get('ln_stats'))
        return grads
    elif sv.get('x2c') is not None:
        g_z = _tc_s3_backward('conv2', c2, sv['x2c'], g, grads['conv2.0.weight'], grads['conv2.0.bias'])
    else:
        _wgrad(c2, sv['z_head'], g, grads['conv2.0.weight'], grads['conv2.0.bias'])
        g_z = _dgrad(c2, g, sv['z_head'].shape)
    for i in range(len(blocks) - 1, -1, -1) if not sv.get('cp8') else ():
        name, conv = blocks[i]
        z_in, act, xc = sv['blocks'][i]
        site[0] -= 1
        g_conv = _pool_bwd_dropout(act, g_z, 3, ops.ACT_LRELU, a, p, seed, site[0])
        if xc is not None:
            g_in = TcConv.backward(name, conv, xc, g_conv, grads[f'{name}.0.weight'], grads[f'{name}.0.bias'], need_dx=True)
        else:
            _wgrad(conv, z_in, g_conv, grads[f'{name}.0.weight'], grads[f'{name}.0.bias'])
            g_in = _dgrad(conv, g_conv, z_in.shape)
        g_z = _add(g_in, g_z) if (residual and i > 0) else g_in
    x = sv['x']
    B, C, T, F = x.shape
    call('layernorm_cf_param_grad_f32', x, g_z, grads['layernorm.weight'], grads['layernorm.bias'], B, C, T, F,
         float(model.layernorm.eps), 0.0, stream_ptr())
    return grads


class CnnTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, seed, step, *params):
        with torch.no_grad():
            y, sv = cnn_train_forward(model, x, seed, step)
        ctx.model, ctx.sv = model, sv
        return y

    @staticmethod
    def backward(ctx, g_y):
        model = ctx.model
        named = list(model.named_parameters())
        with torch.no_grad():
            grads = {n: torch.empty_like(p) for n, p in named}
            cnn_train_backward(model, ctx.sv, g_y, grads)
        return (None, None, None, None) + tuple(grads[n] for n, _ in named)


def cnn_forward_train(model, x):
    """Entry used by model.forward when autograd is recording."""
    model._train_calls = getattr(model, '_train_calls', 0) + 1
    seed = getattr(model, 'dropout_seed', 0x5EED)
    return CnnTrainFunction.apply(model, x, seed, model._train_calls, *[p for _, p in model.named_parameters()])


class TrainStep:
    """Fused training step on flat parameter / gradient / moment buffers (one all-reduce, one AdamW sweep)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, seed=0x5EED, process_group=None, graph=False):
        """graph=True: from the third step on, forward + loss + backward replay ONE captured CUDA graph (fixed input shape; dropout offsets
        come from a device-side step counter, so a replayed step equals the eager step of the same number); all-reduce and AdamW stay eager."""
        self.model, self.lr, self.betas, self.eps, self.wd, self.seed = model, lr, betas, eps, weight_decay, seed
        self.group = process_group
        named = list(model.named_parameters())
        dev = named[0][1].device
        n = sum(p.numel() for _, p in named)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads, off = {}, 0
        with torch.no_grad():
            for name, p in named:
                k = p.numel()
                self.flat_p[off:off + k] = p.reshape(-1)
                p.data = self.flat_p[off:off + k].view_as(p)          # parameters become views of the flat buffer
                self.grads[name] = self.flat_g[off:off + k].view_as(p)
                off += k
        self.step_count = 0
        self.use_graph, self._graph = bool(graph), None

    def _forward_backward(self, x, target, tape_step):
        with TcConv.scope(self):
            y, sv = cnn_train_forward(self.model, x, self.seed, tape_step)
            loss, g_y = ops.bce_fwd_bwd(y, target.contiguous())
            cnn_train_backward(self.model, sv, g_y, self.grads)
        return loss

    def release(self):
        """Free the pooled activation planes of this step (and the captured graph that points into them)."""
        self._graph = None
        TcConv.release(self)

    def _capture(self, x, target):
        global _step_dev, _step_mul
        self._xs, self._ts = x.clone(), target.contiguous().clone()
        self._step_dev = torch.zeros(1, dtype=torch.int64, device=x.device)
        self._graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self._graph):
            self._step_dev += 1
            _step_dev, _step_mul = self._step_dev, 64
            try:
                self._loss = self._forward_backward(self._xs, self._ts, 0)
            finally:
                _step_dev, _step_mul = None, 256
        self.launches_per_replay = _lib.launch_count() - n0
        self.replays = 0
        self._step_dev.fill_(self.step_count - 1)

    def __call__(self, x, target):
        import torch.distributed as dist
        self.step_count += 1
        with torch.no_grad():
            if self.use_graph and self.step_count > 2:
                if self._graph is None:
                    self._capture(x, target)
                self._xs.copy_(x)
                self._ts.copy_(target)
                self._graph.replay()
                self.replays += 1
                loss = self._loss
            else:
                loss = self._forward_backward(x, target, self.step_count)
            scale = 1.0
            if self.group is not None or (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
                ev = getattr(self, 'comm_events', None)
                if ev is not None:                    # bench.py: CUDA events around the collective (time of the all-reduce per step)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                dist.all_reduce(self.flat_g, group=self.group)
                if ev is not None:
                    e1.record()
                    ev.append((e0, e1))
                scale = 1.0 / dist.get_world_size(self.group)
            call('adamw_f32', self.flat_p, self.flat_g, self.m, self.v, _lib.i64(self.flat_p.numel()), float(self.lr), float(self.betas[0]),
                 float(self.betas[1]), float(self.eps), float(self.wd), self.step_count, float(scale), stream_ptr())
        _exec.invalidate_caches(self.model)   # packed operands derived from the old parameter values are stale
        return loss
