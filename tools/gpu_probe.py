"""GPU-side diagnostics (not a test): runs the tcgen05 convolution on progressively harder shapes, each in its
own subprocess with a timeout (a device trap poisons the CUDA context), and prints max-abs errors against a torch
fp32 reference on fp16-rounded operands.  Usage on the GPU box:  python tools/gpu_probe.py [case ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (B, Cin, Cout, T, F, KH, KW)
    'k1_c8':      (2, 8, 40, 6, 24, 1, 1),
    'k1x3_c8':    (2, 8, 40, 6, 24, 1, 3),
    'k3x1_c8':    (2, 8, 40, 7, 24, 3, 1),
    'k3_c16':     (2, 16, 40, 7, 24, 3, 3),
    'k3_c24':     (3, 24, 40, 9, 40, 3, 3),
    'k15_c6':     (3, 6, 40, 20, 216, 15, 15),
    'k15_c40':    (3, 40, 40, 75, 216, 15, 15),
    'k15_c20o20': (2, 20, 20, 75, 216, 15, 15),
    'k5_c64o128': (2, 64, 128, 9, 27, 5, 5),
    'perf_k15_c40': (592, 40, 40, 75, 216, 15, 15),
    'perf_k15_c6': (592, 6, 40, 75, 216, 15, 15),
    'perf_k15_c20': (592, 20, 20, 75, 216, 15, 15),
}


def run_case(name):
    import torch
    import torch.nn.functional as F
    from multipitch_architectures_b200 import ops
    B, Cin, Cout, T, Fq, KH, KW = CASES[name]
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, T, Fq, generator=g)
    w = torch.randn(Cout, Cin, KH, KW, generator=g) / (Cin * KH * KW) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    xr = x.half().float()
    wr = w.half().float()
    nref = min(B, 3)
    ref = F.leaky_relu(F.conv2d(xr[:nref].double(), wr.double(), b.double(), padding=(KH // 2, KW // 2)), 0.3).float()
    dev = 'cuda'
    xc = ops.nchw_to_cp8(x.to(dev))
    back = ops.cp8_to_nchw(xc).cpu()
    wp = ops.conv_tc_pack(w, dev)
    yc = ops.conv_tc(xc, wp, b.to(dev), Cout, (KH, KW), ops.ACT_LRELU, 0.3)
    torch.cuda.synchronize()
    y = ops.cp8_to_nchw(yc).cpu()[:nref]
    back = back[:nref]
    xr = xr[:nref]
    err = (y - ref).abs()
    res = dict(case=name, roundtrip=float((back - xr).abs().max()), max_err=float(err.max()), mean_err=float(err.mean()),
               ref_absmax=float(ref.abs().max()), f16_eps_bound=float(ref.abs().max()) * 2 ** -11)
    # where are the errors? per output row / per channel maxima help localise descriptor mistakes
    res['err_by_row'] = [round(float(v), 4) for v in err.amax(dim=(0, 1, 3))[:12]]
    res['err_by_cout'] = [round(float(v), 4) for v in err.amax(dim=(0, 2, 3))[:8]]
    res['err_by_col'] = [round(float(v), 4) for v in err.amax(dim=(0, 1, 2))[:12]]
    # timing
    if res['max_err'] < 0.05:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            ops.conv_tc(xc, wp, b.to(dev), Cout, (KH, KW), ops.ACT_LRELU, 0.3, out=yc)
        ev0.record()
        for _ in range(5):
            ops.conv_tc(xc, wp, b.to(dev), Cout, (KH, KW), ops.ACT_LRELU, 0.3, out=yc)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 5
        res['ms'] = ms
        res['tflops'] = 2.0 * B * Cout * Cin * KH * KW * T * Fq / ms / 1e9
    print('PROBE ' + json.dumps(res), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--one':
        run_case(sys.argv[2])
        sys.exit(0)
    names = sys.argv[1:] or list(CASES)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'probe.log'), 'a') as log:
        for n in names:
            try:
                r = subprocess.run([sys.executable, __file__, '--one', n], capture_output=True, text=True, timeout=120)
                out = r.stdout + ('\nSTDERR: ' + r.stderr[-1500:] if r.returncode != 0 else '')
            except subprocess.TimeoutExpired:
                out = f'PROBE {{"case": "{n}", "timeout": true}}'
            print(out.strip())
            log.write(out.strip() + '\n')
