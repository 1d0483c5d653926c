#!/usr/bin/env python
"""Time one conv_tc configuration: python tools/time_conv.py B Cin Cout T F KH KW [J] [subsample] [prec]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multipitch_architectures_b200 import ops        # noqa: E402

B, Cin, Cout, T, F, KH, KW = [int(v) for v in sys.argv[1:8]]
J = int(sys.argv[8]) if len(sys.argv) > 8 else 0
sub = int(sys.argv[9]) if len(sys.argv) > 9 else 0
prec = sys.argv[10] if len(sys.argv) > 10 else 'fp16'
fmt = ops.fmt_of(prec)
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.standard_normal((min(B, 8), Cin, T, F)).astype(np.float32)).cuda()
w = torch.from_numpy((rng.standard_normal((Cout, Cin, KH, KW)) * (Cin * KH * KW) ** -0.5).astype(np.float32))
b = torch.zeros(Cout).cuda()
xc = ops.CP8(B, Cin, T, F, fmt=fmt, device='cuda')
ops.nchw_to_cp8(x, out=xc.first(min(B, 8)), fmt=fmt)
wp = ops.conv_tc_pack(w, 'cuda', fmt, J)
out = ops.compact_cp8(B, Cout, T, F, 'cuda', fmt) if sub else None
run = lambda: ops.conv_tc(xc, wp, b, Cout, (KH, KW), ops.ACT_LRELU, 0.3, J=J, out=out, subsample=(1, 0) if sub else None)
for _ in range(3):
    y = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * B * Cin * Cout * KH * KW * T * F
print(f'{sys.argv[1:]}: {ms:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s')
