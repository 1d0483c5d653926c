"""Secondary measurements (not the bench.py contract): patch-wise inference throughput of every BASELINE model through
the public module forward (batch of 50 materialised patches, as the reference's test loop), CUDA events.
    python tools/bench_models.py [model ...]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tests.refshapes import build_model  # noqa: E402
from tests.weights import fill_state_dict, synth_patches  # noqa: E402

GFLOP = {'cnn_xs': 0.916, 'drcnn': 48.574, 'unet_m': 12.144, 'punet': 81.777, 'saunet_l': 29.110}

if __name__ == '__main__':
    names = sys.argv[1:] or ['cnn_xs', 'drcnn', 'unet_m', 'punet', 'saunet_l']
    for name in names:
        for prec in (('fp32', 'fp16') if name in ('cnn_xs', 'drcnn', 'unet_m', 'punet', 'saunet_l') else ('fp32',)):
            try:
                m = build_model(name, precision=prec)
            except Exception as e:
                print(json.dumps({'model': name, 'precision': prec, 'error': str(e)[:200]}))
                continue
            m.load_state_dict(fill_state_dict(m.state_dict(), 0))
            m = m.cuda().eval()
            B = 50
            x = synth_patches(B, 1).cuda()
            with torch.no_grad():
                for _ in range(2):
                    m(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    m(x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            pps = B / (ms / 1e3)
            print(json.dumps({'model': name, 'precision': prec, 'batch': B, 'ms_per_batch': round(ms, 3), 'patches_per_s': round(pps, 1),
                              'audio_s_per_s': round(pps / 43.06640625, 2), 'tflops': round(pps * GFLOP[name] / 1e3, 2)}), flush=True)
