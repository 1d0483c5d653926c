"""Shared loader of the REALISTIC weight set (tests/golden/realistic_weights.npz: CNN:XS, DRCNN and Unet:M trained by tools/train_realistic.py
with this repo's own loop.fit; tests/golden/realistic_golden.npz: the outputs of the UNMODIFIED reference classes on those weights over a
held-out 30 s clip, made by tests/golden/make_golden.py realistic).  Used by the parity tests and by bench.py's `parity` block: data only."""
import functools
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODELS = ('cnn_xs', 'drcnn', 'unet_m')
CLIP = dict(seed=777, seconds=30.0)
HCQT_KW = dict(fs=22050, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)
THRESHOLD = 0.4


@functools.lru_cache(maxsize=None)
def golden():
    return np.load(os.path.join(GOLDEN_DIR, 'realistic_golden.npz'))


@functools.lru_cache(maxsize=None)
def _weights():
    return np.load(os.path.join(GOLDEN_DIR, 'realistic_weights.npz'))


def state_dict(name):
    w = _weights()
    pre = name + '/'
    return {k[len(pre):]: torch.from_numpy(w[k]) for k in w.files if k.startswith(pre)}


def labels():
    g = golden()
    n = int(g['n_frames'][0])
    return np.unpackbits(g['roll'])[:n * 72].reshape(n, 72).astype(np.float32)


def prf_counts(targ, pred, thr=THRESHOLD):
    est = pred >= thr
    tp = int((est & (targ > 0)).sum())
    return tp, int(est.sum()) - tp, int((targ > 0).sum()) - tp


def prf(counts):
    tp, fp, fn = counts
    P = tp / (tp + fp) if tp + fp else 0.0
    R = tp / (tp + fn) if tp + fn else 0.0
    return P, R, (2 * P * R / (P + R) if P + R else 0.0)


def compare(pred, name):
    """-> dict(max_abs, flips, counts, counts_ref, prf_equal_3dec) of activations `pred` [N,72] against the reference golden."""
    g = golden()
    ref = g[name + '__y']
    n = min(len(pred), len(ref))
    pred, ref, lab = np.asarray(pred[:n], dtype=np.float32), ref[:n], labels()[:n]
    flips = (pred >= THRESHOLD) != (ref >= THRESHOLD)
    c, cr = prf_counts(lab, pred), prf_counts(lab, ref)
    return {'max_abs': float(np.abs(pred - ref).max()), 'flips': int(flips.sum()),
            'flip_margin': float(np.abs(ref - THRESHOLD)[flips].max()) if flips.any() else 0.0,
            'counts': c, 'counts_ref': cr, 'prf': prf(c), 'prf_ref': prf(cr),
            'prf_equal_3dec': all(round(a, 3) == round(b, 3) for a, b in zip(prf(c), prf(cr))), 'frames': n,
            'out_span': (float(ref.min()), float(ref.max()))}
