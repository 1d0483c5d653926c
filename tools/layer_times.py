#!/usr/bin/env python
"""Per-launch timing of the tcgen05 inference path of a U-Net-family model: CUDA events around every libmpa call of one forward.

    python tools/layer_times.py unet_m [batch] [precision]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multipitch_architectures_b200 import _lib, ops              # noqa: E402
from tests.refshapes import build_model                          # noqa: E402
from tests.weights import fill_state_dict, synth_patches         # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'unet_m'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 646
    prec = sys.argv[3] if len(sys.argv) > 3 else 'fp16'
    m = build_model(name, precision=prec)
    m.load_state_dict(fill_state_dict(m.state_dict(), 0))
    m = m.cuda().eval()
    x = synth_patches(8, 0).cuda().repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous()
    with torch.no_grad():
        for _ in range(3):
            m(x)
    torch.cuda.synchronize()
    rec = []
    orig = _lib.call

    def timed_call(fn, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(fn, *args)
        e1.record()
        desc = ''
        if fn == 'conv_tc_f16':
            # (in, w, bias, out, mode, stride, offset, n, Cin, Cout, T, F, KH, KW, pitch, ...)
            n, Cin, Cout, T, F, KH, KW = args[7:14]
            J, row0, n_rows = args[20:23]
            rows = n_rows if n_rows else T
            desc = f'n={n} {Cin}->{Cout} {KH}x{KW} T={T} F={F} mode={args[4]} J={J} rows={rows}'
            flops = 2.0 * n * Cin * Cout * KH * KW * rows * F
        else:
            flops = 0.0
        rec.append((fn, desc, e0, e1, flops))
    _lib.call = timed_call
    ops.call = timed_call
    from multipitch_architectures_b200.libdl.nn_models import _exec
    with torch.no_grad():
        m(x)
    torch.cuda.synchronize()
    _lib.call = orig
    ops.call = orig
    tot = sum(a.elapsed_time(b) for _, _, a, b, _ in rec)
    print(f'{name} B={B} {prec}: {len(rec)} calls, {tot:.3f} ms (event-to-event, includes launch gaps)')
    for fn, desc, a, b, fl in rec:
        ms = a.elapsed_time(b)
        print(f'{ms:8.3f} ms {ms / tot * 100:5.1f}%  {fn:22s} {desc}' + (f'  {fl / ms / 1e9:7.1f} TFLOP/s' if fl else ''))
    agg = collections.Counter()
    for fn, desc, a, b, fl in rec:
        agg[fn] += a.elapsed_time(b)
    print({k: round(v, 3) for k, v in agg.most_common()})


if __name__ == '__main__':
    main()
