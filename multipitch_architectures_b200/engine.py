"""Streaming patch-wise inference: raw audio (or an HCQT) in, [n_frames, 72] pitch activations out.

This is the B200 formulation of the reference's test loop
(/root/reference/experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py:413-436): pad 37/38 zero frames,
cut stride-1 patches of 75 frames, log-compress, run the network on every patch, keep the centre frame.  The
patch-wise semantics are preserved exactly (zero padding at every patch edge reaches the centre frame), but
patches are never materialised on the input side: LayerNorm(+log compression) is evaluated once per FRAME and the
first convolution gathers its 75-row windows straight out of the frame-major plane (75x fewer input bytes)."""
import numpy as np
import torch

from . import _lib, ops
from .libdl.nn_models import _exec
from .libdl.nn_models.basic_cnns import basic_cnn_segm_sigmoid, deep_cnn_segm_sigmoid

CONTEXT = 75
HALF = CONTEXT // 2


class CnnStreamEngine:
    """DRCNN / DCNN / CNN inference over a whole recording with the tcgen05 convolution stack (model.precision
    'fp16' or 'bf16')."""

    def __init__(self, model, chunk=646, compression=10.0):
        if not isinstance(model, (basic_cnn_segm_sigmoid, deep_cnn_segm_sigmoid)):
            raise TypeError('CnnStreamEngine serves the CNN / DCNN / DRCNN family')
        self.model, self.chunk, self.compression = model, int(chunk), float(compression)
        self.dev = next(model.parameters()).device
        if self.dev.type != 'cuda':
            raise _lib.MpaError('the engine needs the model on a CUDA (sm_100a) device')
        if not _exec.tc_eligible(model, model.n_bins_in):
            raise _lib.MpaError("CnnStreamEngine needs precision 'fp16' or 'bf16' and <=128 channels per layer; "
                                "use predict_patchwise for the fp32 path")
        self.fmt = ops.fmt_of(model.precision)
        self.blocks = _exec.cnn_blocks(model)
        self.residual = getattr(model, 'residual', False)
        self.F = model.n_bins_in
        self.C0 = self.blocks[0][1].weight.shape[0]
        self.pitch, self.pf, self.pt = (self.F + 8 + 15) // 16 * 16, 8, 1
        self._bufs = None
        self.timers = None          # optional: list collecting (tag, start_event, end_event)

    # -- buffers are allocated once (their zero borders are never written)
    def _buffers(self):
        if self._bufs is None:
            mk = lambda: ops.CP8(self.chunk, self.C0, CONTEXT, self.F, self.pitch, self.pf, self.pt, self.dev, fmt=self.fmt)
            self._bufs = (mk(), mk(), mk())
        return self._bufs

    def _timed(self, tag, fn):
        if self.timers is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        self.timers.append((tag, e0, e1))
        return r

    def predict_hcqt(self, hcqt, lo=0, hi=None):
        """hcqt: [C, N, F] fp32 CUDA, linear magnitudes (the reference's .npy transposed (2,1,0)) -> [hi-lo, n_out] fp32
        for the patches centred on frames lo..hi-1 (default: all N); frames beyond the array's ends are zero padding."""
        m, cache, a = self.model, self.model._cache, self.model.a_lrelu
        C, N, F = hcqt.shape
        if C != m.n_chan_input or F != self.F:
            raise ValueError(f'expected [{m.n_chan_input}, N, {self.F}], got {tuple(hcqt.shape)}')
        hcqt = hcqt.contiguous().float()
        lead, trail = self.pt + HALF, HALF + self.pt + 1
        rows = lead + N + trail
        plane = torch.zeros(rows, self.pitch, 8, dtype=ops._FMT_DTYPE[self.fmt], device=self.dev)
        self._timed('layernorm_frames', lambda: _lib.call(
            'layernorm_frames', hcqt, m.layernorm.weight, m.layernorm.bias, None, plane, C, N, F, lead, trail, self.pitch, self.pf,
            float(m.layernorm.eps), self.compression, self.fmt, _lib.stream_ptr()))
        ya, za, zb = self._buffers()
        outs = []
        hi = N if hi is None else hi
        if not (0 <= lo <= hi <= N):
            raise ValueError('bad frame range')
        for i0 in range(lo, hi, self.chunk):
            n = min(self.chunk, hi - i0)
            z_prev = None
            for li, (name, conv) in enumerate(self.blocks):
                w = conv.weight
                wp = cache.get(f'{name}:wtc{self.fmt}', [w], lambda: ops.conv_tc_pack(w, self.dev, self.fmt))
                if li == 0:
                    src = ops.CP8.__new__(ops.CP8)
                    src.B, src.C, src.T, src.F, src.pitch, src.pf, src.pt, src.NC, src.fmt = n, C, CONTEXT, F, self.pitch, self.pf, self.pt, 1, self.fmt
                    src.ncs, src.chunk0 = 1, 0
                    src.buf = plane[i0:]
                    self._timed('conv_tc_first', lambda: ops.conv_tc(src, wp, conv.bias, self.C0, tuple(conv.kernel_size), ops.ACT_LRELU, a,
                                                                    out=ya.first(n), n_patches=n, patch_stride_rows=1, T=CONTEXT))
                    cur = za
                    self._timed('pool3', lambda: ops.pool3_res_cp8(ya.first(n), None, out=cur.first(n)))
                else:
                    self._timed('conv_tc', lambda: ops.conv_tc(z_prev.first(n), wp, conv.bias, self.C0, tuple(conv.kernel_size), ops.ACT_LRELU,
                                                              a, out=ya.first(n), n_patches=n))
                    cur = zb if z_prev is za else za
                    self._timed('pool3', lambda: ops.pool3_res_cp8(ya.first(n), z_prev.first(n) if self.residual else None,
                                                                  out=cur.first(n)))
                z_prev = cur
            y = self._timed('head', lambda: _exec.head_tc(cache, m, z_prev.first(n), a))
            outs.append(y.reshape(n, -1))
        if not outs:
            return torch.empty(0, 0, dtype=torch.float32, device=self.dev)
        return outs[0] if len(outs) == 1 else torch.cat(outs, 0)

    def predict_audio(self, y, plan):
        """y: 1-D float32 CUDA audio; plan: HCQTPlan.  -> ([N, 72] activations, tuning index tensor)."""
        hcqt, tun = self._timed('hcqt', lambda: plan.run(y))
        return self.predict_hcqt(hcqt), tun


def predict_patchwise(model, hcqt, batch=50):
    """Reference-shaped loop for any model (U-Nets included): materialised stride-1 patches in batches of `batch`
    consecutive frames (the batch composition matters for the SAUnet's batch-axis attention)."""
    C, N, F = hcqt.shape
    dev = hcqt.device
    padded = torch.zeros(C, N + CONTEXT, F, dtype=torch.float32, device=dev)
    padded[:, HALF:HALF + N] = hcqt
    outs, npreds = [], []
    for i0 in range(0, N, batch):
        n = min(batch, N - i0)
        x = torch.empty(n, C, CONTEXT, F, dtype=torch.float32, device=dev)
        _lib.call('gather_patches_f32', padded, x, C, N + CONTEXT, F, i0, n, CONTEXT, 1, 10.0, _lib.stream_ptr())
        with torch.no_grad():
            y = model(x)
        if isinstance(y, tuple):
            npreds.append(y[1].reshape(n, -1))
            y = y[0]
        outs.append(y.reshape(n, -1))
    out = torch.cat(outs, 0)
    return (out, torch.cat(npreds, 0)) if npreds else out
