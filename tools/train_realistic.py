#!/usr/bin/env python
"""Produce the REALISTIC weight set of the parity tests: train CNN:XS, the DRCNN and Unet:M for a few hundred steps
with this package's own loop.fit on synthetic labelled audio (tests/synth.py: polyphonic harmonic tones with known notes), on the GPU.

    python tools/train_realistic.py --out gpurun_out/realistic [--steps 500] [--models cnn_xs drcnn unet_m]

Outputs per model: <out>/<name>.pt (fp32 state_dict with the reference's key names) and <out>/<name>.json (loss history + a first look
at the precision modes on a held-out 30 s clip: max |fp16 - fp32|, |bf16 - fp32|, threshold flips, P/R/F against the labels).
The committed goldens are then made from these weights by the UNMODIFIED reference classes on the CPU:
`python tests/golden/make_golden.py realistic` (needs /root/reference; the weights go to tests/golden/realistic_weights.npz)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import synth                                            # noqa: E402
from tests.weights import MODEL_SPECS                              # noqa: E402

FPS = 22050 / 512
TRAIN_PARAMS = {'context': 75, 'stride': 3, 'compression': 10, 'aug:transpsemitones': 5, 'aug:randomeq': 20, 'aug:noisestd': 1e-4,
                'aug:tuning': True}
TEST_SEED, TEST_SECONDS = 777, 30.0


def hcqt_and_labels(plan, seed, seconds, dev):
    y, notes = synth.synth_clip_labeled(seed, seconds)
    h, _ = plan.run(torch.from_numpy(y).to(dev))
    return h.contiguous(), synth.piano_roll(notes, h.shape[1])


def prf(t, p, thr=0.4):
    e = p >= thr
    tp, fp, fn = float((e & (t > 0)).sum()), float((e & (t == 0)).sum()), float((~e & (t > 0)).sum())
    P = tp / (tp + fp) if tp + fp else 0.0
    R = tp / (tp + fn) if tp + fn else 0.0
    return P, R, (2 * P * R / (P + R) if P + R else 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='gpurun_out/realistic')
    ap.add_argument('--steps', type=int, default=500)
    ap.add_argument('--clips', type=int, default=24)
    ap.add_argument('--models', nargs='*', default=['cnn_xs', 'drcnn', 'unet_m'])
    ap.add_argument('--train-precision', default='bf16')
    ap.add_argument('--lr', type=float, default=0.0, help='0 = the reference scripts\' values (1e-3; DRCNN 2e-4)')
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    dev = torch.device('cuda', 0)
    torch.cuda.set_device(dev)
    from multipitch_architectures_b200.engine import CnnStreamEngine, predict_patchwise
    from multipitch_architectures_b200.libdl import nn_models as M
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import C1_HZ, get_plan
    from multipitch_architectures_b200.loop import fit
    plan = get_plan(22050, float(C1_HZ / 2 ** ((3 - 1) / (2 * 36))), 512, 36, 6, 5, 1, str(dev))
    torch.manual_seed(0)
    train = []
    for i in range(a.clips):
        h, roll = hcqt_and_labels(plan, 5000 + i, 20.0, dev)
        train.append(dataset_context(h, torch.from_numpy(roll).to(dev), dict(TRAIN_PARAMS)))
    h_test, roll_test = hcqt_and_labels(plan, TEST_SEED, TEST_SECONDS, dev)
    n_items = sum(len(d) for d in train)
    print(f'{a.clips} training clips, {n_items} patches per epoch; test clip {tuple(h_test.shape)}', flush=True)
    for name in a.models:
        spec = MODEL_SPECS[name]
        torch.manual_seed(1)
        model = getattr(M, spec['cls'])(**spec['kw'], precision=a.train_precision).to(dev)
        batches_per_epoch = -(-n_items // 25)
        epochs = max(1, -(-a.steps // batches_per_epoch))
        t0 = time.time()
        lr = a.lr if a.lr > 0 else (2e-4 if name == 'drcnn' else 1e-3)          # RETRAIN4_exp128c...py: DRCNN trains at 2e-4
        hist = fit(model, train, None, batch_size=25, lr=lr, weight_decay=0.01, max_epochs=epochs,
                   max_batches_per_epoch=min(batches_per_epoch, a.steps), scheduler=False, early=False, seed=0)
        torch.cuda.synchronize()
        t_train = time.time() - t0
        sd = {k: v.detach().float().cpu().clone() if v.is_floating_point() else v.detach().cpu().clone() for k, v in model.state_dict().items()}
        torch.save(sd, os.path.join(a.out, name + '.pt'))
        # first look at the precision modes on the held-out clip (the committed parity numbers come from the reference goldens)
        model.eval()
        outs = {}
        with torch.no_grad():
            for prec in ('fp32', 'fp16', 'bf16'):
                model.precision = prec
                if prec != 'fp32' and spec['cls'].startswith(('basic_cnn', 'deep_cnn')):
                    y = CnnStreamEngine(model).predict_hcqt(h_test)
                else:
                    y = predict_patchwise(model, h_test, batch=50)
                    y = y[0] if isinstance(y, tuple) else y
                outs[prec] = y.float().cpu().numpy()
        ref = outs['fp32']
        rec = {'model': name, 'steps': len(hist) * min(batches_per_epoch, a.steps), 'train_seconds': t_train, 'history': hist,
               'train_precision': a.train_precision, 'lr': lr, 'out_min': float(ref.min()), 'out_max': float(ref.max()),
               'active_frac_ref': float((ref >= 0.4).mean()), 'prf_fp32_vs_labels': prf(roll_test, ref)}
        for prec in ('fp16', 'bf16'):
            d = np.abs(outs[prec] - ref)
            rec[prec] = {'max_abs_vs_fp32': float(d.max()), 'mean_abs': float(d.mean()), 'flips_at_0.4': int(((outs[prec] >= 0.4) != (ref >= 0.4)).sum()),
                         'prf_vs_labels': prf(roll_test, outs[prec])}
        json.dump(rec, open(os.path.join(a.out, name + '.json'), 'w'), indent=1)
        print(json.dumps({k: v for k, v in rec.items() if k != 'history'}), flush=True)
        print('loss first/last epoch:', hist[0]['train_loss'], hist[-1]['train_loss'], flush=True)


if __name__ == '__main__':
    main()
