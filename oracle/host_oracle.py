"""ORACLE (test infrastructure only — never imported by the product path).

Plain NumPy / pure-Python restatements of the integer / indexing pieces of the hot path.
All of these are PINNED: tests/golden/make_golden.py runs the reference's own functions
(importable here with a stub shim for matplotlib / IPython / librosa) and the fixtures under
tests/golden/ hold their outputs.

  * annotation rasteriser ........ /root/reference/libdl/data_preprocessing/hcqt.py:205-272
  * patch index math + log-compr . /root/reference/libdl/data_loaders/hcqt_datasets.py:63-75,105-106
  * inference padding ............ /root/reference/experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py:413-423
  * thresholded P/R/F ............ /root/reference/libdl/metrics/eval_metrics.py:50-62 ->
                                   /root/reference/libfmp/c5/c5s2_chord_rec_template.py:238-261
  * the other evaluation measures  /root/reference/libdl/metrics/eval_metrics.py:64-110 (pinned: reference function run with stubs,
                                   tests/golden/ext_golden.npz); mir_eval multipitch scores (:158-189) restated from mir_eval's
                                   published definitions, PARITY UNPINNED (mir_eval is not installed)
  * training-time augmentations .. /root/reference/libdl/data_loaders/hcqt_datasets.py:77-139, deterministic in the random decisions
                                   (pinned: the reference __getitem__ run with its torch draws recorded, ext_golden.npz)
"""
import numpy as np


def annotation_array_nooverlap(note_events, n_frames, fs_hcqt, annot_type='pitch', shorten=1.0):
    """note_events [n,>=3] (start_sec, end_sec, pitch, ...) -> binary [H, n_frames] float64."""
    height = {'pitch_class': 12, 'pitch': 128, 'instruments': 1}[annot_type]
    ev = np.array(note_events, dtype=np.float64, copy=True)
    if shorten != 1.0:
        ev[:, 1] = ev[:, 0] + shorten * (ev[:, 1] - ev[:, 0])
    start = np.floor(ev[:, 0] * fs_hcqt).astype(np.int64)
    end = np.floor(ev[:, 1] * fs_hcqt).astype(np.int64)
    vanishing = np.nonzero((end - start) < 1)[0]
    # every frame index that is the end of a vanishing event pushes later-starting / same-ending events by one
    for v in np.unique(end[vanishing]):
        start[start == v] += 1
        end[end == v] += 1
    start[vanishing] -= 1
    still = np.nonzero((end - start) < 1)[0]
    start[still] -= 1
    assert np.all(end - start >= 1), 'still events of length<1 after correction!'
    out = np.zeros((height, n_frames))
    for s, e, p in zip(start, end, ev[:, 2]):
        if annot_type == 'pitch_class':
            row = int(np.mod(p, 12))
        elif annot_type == 'pitch':
            row = int(p)
        else:
            row = 0
        out[row, s:e] = 1
    return out


def context_len(n_frames_padded, context=75, stride=1):
    return (n_frames_padded - context) // stride


def context_item(inputs, targets, index, context=75, stride=1, compression=10.0):
    """inputs [C, N, F], targets [N, P] -> (X [C,context,F] f32, y [1,1,P] f32)."""
    half = context // 2
    c = index * stride + half
    X = np.asarray(inputs[:, c - half:c + half + 1, :], dtype=np.float32)
    y = np.asarray(targets[c, :], dtype=np.float32)[None, None, :]
    if compression is not None:
        X = np.log(1 + np.float32(compression) * X).astype(np.float32)
    return X, y


def pad_for_inference(inputs, targets, context=75):
    """exp126a:420-421: pad (context//2, context//2+1) frames of zeros -> len == n_frames."""
    half = context // 2
    return (np.pad(inputs, ((0, 0), (half, half + 1), (0, 0))),
            np.pad(targets, ((half, half + 1), (0, 0))))


def eval_prf(targ, pred, threshold=0.4):
    est = pred >= threshold
    TP = int(np.sum(np.logical_and(targ, est)))
    FP = int(np.sum(est > 0)) - TP
    FN = int(np.sum(targ > 0)) - TP
    P = R = F = 0
    if TP > 0:
        P = TP / (TP + FP)
        R = TP / (TP + FN)
        F = 2 * P * R / (P + R)
    return P, R, F, TP, FP, FN


# ----------------------------------------------------------------------------- evaluation measures (eval_metrics.py:64-110)
def eval_measure(targ, pred, measure, threshold=0.5):
    eps = np.finfo(float).eps
    targ, pred = np.asarray(targ, dtype=np.float64), np.asarray(pred, dtype=np.float64)
    est = pred >= threshold
    if measure in ('precision', 'recall', 'f_measure'):
        return eval_prf(targ, pred, threshold)[('precision', 'recall', 'f_measure').index(measure)]
    if measure == 'cosine_sim':
        def l2(X):
            out = np.empty_like(X)
            for n in range(X.shape[0]):
                s = np.sqrt(np.sum(X[n] ** 2))
                out[n] = X[n] / s if s > 1e-10 else np.ones(X.shape[1]) / np.sqrt(X.shape[1])
            return out
        return np.sum(l2(targ) * l2(pred)) / targ.shape[0]
    if measure == 'binary_crossentropy':
        return -np.mean(targ * np.log2(pred + eps) + (1 - targ) * np.log2(1 - pred + eps))
    if measure == 'euclidean_distance':
        return np.mean(np.sqrt(np.sum((targ - pred) ** 2, axis=1)))
    if measure == 'binary_accuracy':
        return np.mean(est == targ)
    if measure == 'soft_accuracy':
        return np.mean(targ * pred + (1 - targ) * (1 - pred))
    if measure == 'accum_energy':
        return np.mean(np.sum(targ * pred, axis=1) / (np.sum(targ, axis=1) + eps))
    raise ValueError(measure)


def mpe_scores(targ, pred, threshold=0.5, min_pitch=24):
    """mir_eval.multipitch.evaluate for pitches on a semitone grid, frame by frame with Python sets / multisets (PARITY UNPINNED)."""
    est = np.asarray(pred) >= threshold
    tot = dict(tp=0, ctp=0, ne=0, nr=0, nmin=0, nmax=0, miss=0, fa=0)
    for k in range(targ.shape[0]):
        r = [min_pitch + int(q) for q in np.nonzero(targ[k])[0]]
        e = [min_pitch + int(q) for q in np.nonzero(est[k])[0]]
        tot['tp'] += len(set(r) & set(e))
        tot['ctp'] += sum(min([x % 12 for x in r].count(c), [x % 12 for x in e].count(c)) for c in range(12))
        tot['ne'] += len(e)
        tot['nr'] += len(r)
        tot['nmin'] += min(len(r), len(e))
        tot['nmax'] += max(len(r), len(e))
        tot['miss'] += max(0, len(r) - len(e))
        tot['fa'] += max(0, len(e) - len(r))
    out = {}
    for prefix, tp in (('', tot['tp']), ('Chroma ', tot['ctp'])):
        d = tot['nr'] if tot['nr'] > 0 else 1.0
        out[prefix + 'Precision'] = tp / tot['ne'] if tot['ne'] else 0.0
        out[prefix + 'Recall'] = tp / tot['nr'] if tot['nr'] else 0.0
        out[prefix + 'Accuracy'] = tp / (tot['ne'] + tot['nr'] - tp) if (tot['ne'] + tot['nr'] - tp) else 0.0
        out[prefix + 'Substitution Error'] = (tot['nmin'] - tp) / d
        out[prefix + 'Miss Error'] = tot['miss'] / d
        out[prefix + 'False Alarm Error'] = tot['fa'] / d
        out[prefix + 'Total Error'] = (tot['nmax'] - tp) / d
    return out


# ----------------------------------------------------------------------------- augmentations (hcqt_datasets.py:77-139)
def augment_item(X, y, compression=10.0, eq=None, tune2=0, transp=0, noise=None, fill_tune=None, fill_transp=None):
    """One patch through the reference's augmentation chain with the random draws given explicitly.
    X [C,T,F] float32 (un-compressed patch), y [1,1,P]; eq = (alpha, beta) or None; tune2 = tuning shift in half bins (-2..2);
    transp in semitones; noise [C,T,F] = the additive Gaussian sample (or None); fill_* = the N(0,1e-4) samples whose absolute
    values are written into the bins vacated by the tuning shift / transposition (None -> zeros)."""
    X = np.array(X, dtype=np.float32, copy=True)
    y = np.array(y, dtype=np.float32, copy=True)
    C, T, F = X.shape
    if eq is not None:
        alpha, beta = eq
        a = np.float32(2e-6) * np.float32(alpha)
        for c in range(C):
            off = -36 if c == 0 else int(36 * np.log2(c))
            d = (np.arange(F) - (beta - off)) ** 2
            X[c] = (np.float32(1) - a * d.astype(np.float32)) * X[c]
    if noise is not None:
        X = np.abs(X + noise.astype(np.float32))
    if compression is not None:
        X = np.log(np.float32(1) + np.float32(compression) * X).astype(np.float32)
    if tune2:
        Xt = X.copy()
        if tune2 == 1:
            Xt[:, :, 1:] = (X[:, :, :-1] + X[:, :, 1:]) / np.float32(2)
        elif tune2 == -1:
            Xt[:, :, :-1] = (X[:, :, :-1] + X[:, :, 1:]) / np.float32(2)
        else:
            Xt = np.roll(X, tune2 // 2, axis=-1)
        if tune2 > 0:
            Xt[:, :, :1] = 0 if fill_tune is None else np.abs(fill_tune)
        else:
            Xt[:, :, -1:] = 0 if fill_tune is None else np.abs(fill_tune)
        X = Xt
    if transp:
        Xr = np.roll(X, 3 * transp, axis=-1)
        yr = np.roll(y, transp, axis=-1)
        if transp > 0:
            Xr[:, :, :3 * transp] = 0 if fill_transp is None else np.abs(fill_transp)
            yr[:, :, :transp] = 0
        else:
            Xr[:, :, 3 * transp:] = 0 if fill_transp is None else np.abs(fill_transp)
            yr[:, :, transp:] = 0
        if y.shape[-1] == 12:
            yr = np.roll(y, transp, axis=-1)
        X, y = Xr, yr
    return X, y
