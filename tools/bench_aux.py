"""HBM-bound auxiliary kernels of the path (patch cut + augmentation, evaluation sums, .npy feature loader, HCQT variants): CUDA-event
timings against the measured copy bandwidth (MEASURED_PEAKS.json).  One JSON line per kernel.

    python tools/bench_aux.py [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    from multipitch_architectures_b200.libdl.metrics import eval_sums
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import get_plan, C1_HZ
    from tests import synth as Q                 # synthetic clip generator (workload data)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    bw = peaks.get('hbm_gbs', 6541.5)
    dev = 'cuda'
    rng = np.random.default_rng(0)
    out = []

    def report(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out.append({'kernel': name, 'ms': round(ms, 4), 'algorithmic_bytes': int(nbytes), 'achieved_gbs': round(gbs, 1), 'peak_gbs': bw,
                    'frac': round(gbs / bw, 3), 'note': note})
        print(json.dumps(out[-1]))

    # patch cut (+ augmentation): 2048 frames of HCQT, 1024 shuffled patches per launch (larger than L2: 398 MB written)
    N, n = 2048 + 75, 1024
    inp = torch.from_numpy(np.abs(rng.normal(0, 0.05, size=(6, N, 216))).astype(np.float32)).to(dev)
    tg = torch.from_numpy((rng.uniform(size=(N, 72)) < 0.05).astype(np.float32)).to(dev)
    idx = rng.permutation(N - 75)[:n]
    for tag, params in (('gather (no augmentation)', {}),
                        ('EQ + noise + tuning + transposition', {'aug:randomeq': 20, 'aug:noisestd': 1e-4, 'aug:tuning': True, 'aug:transpsemitones': 5})):
        ds = dataset_context(inp, tg, dict({'context': 75, 'stride': 1, 'compression': 10}, **params))
        dec = ds.draw(n) if ds.augmenting else {}
        ms = timed(lambda: ds.gather(idx, decisions=dec), args.reps)
        report('augment_patches_kernel: ' + tag, ms, n * 6 * 75 * 216 * 4 + 6 * N * 216 * 4,
               f'{n} shuffled 6x75x216 patches per launch; bytes = patches written + the HCQT tensor read once (re-reads are L2 hits); '
               'includes the host->device copy of the index / decision arrays and the target kernel')
    # evaluation sums over 200k frames
    Nf = 200000
    t = torch.from_numpy((rng.uniform(size=(Nf, 72)) < 0.04).astype(np.float32)).to(dev)
    p = torch.from_numpy(rng.uniform(size=(Nf, 72)).astype(np.float32) ** 3).to(dev)
    ms = timed(lambda: eval_sums(t, p, 0.4), args.reps)
    report('eval_frame_stats_kernel + eval_reduce_kernel', ms, Nf * 72 * 8, f'{Nf} frames x 72 bins, float64 arithmetic, incl. the 128-byte D2H of the sums')
    # .npy feature loader
    Nn = 40000
    src = torch.from_numpy(np.abs(rng.normal(size=(216, Nn, 6)))).to(dev)
    dst = torch.empty(6, Nn + 75, 216, dtype=torch.float32, device=dev)
    ms = timed(lambda: _lib.call('hcqt_npy_to_frames_f64', src, dst, 216, Nn, 6, 37, 38, _lib.stream_ptr()), args.reps)
    report('hcqt_npy_to_frames_kernel', ms, 216 * Nn * 6 * 12, f'float64 [216,{Nn},6] -> fp32 [6,{Nn}+75,216]')
    # HCQT of a 30 s clip, both variants
    y = torch.from_numpy(Q.synth_clip(0, seconds=30.0)).to(dev)
    fmin = float(C1_HZ / 2 ** (2 / 72))
    for tag, hop, eff in (('compute_efficient_hcqt (3 shared CQTs, hop 512)', 512, True), ('compute_hcqt (6 CQTs, hop 448, early down-sampling)', 448, False)):
        plan = get_plan(22050, fmin, hop, 36, 6, 5, 1, dev, eff)
        ms = timed(lambda: plan.run(y), args.reps)
        frames = plan.n_frames(y.numel())
        report('HCQT 30 s clip: ' + tag, ms, y.numel() * 4 + 6 * frames * 216 * 4,
               f'{frames} frames; tuning estimate + decimator chain + fused FFT/CQT levels; latency-bound (tens of small launches)')


if __name__ == '__main__':
    main()
