"""Evaluation measures with the reference's call signatures (/root/reference/libdl/metrics/eval_metrics.py:8-189).

`calculate_single_measure` / `calculate_eval_measures` / `calculate_mpe_measures_mireval` accept what the reference accepts
(host ndarrays `[n_frames, n_bins]`) and, additionally, CUDA tensors, so the `[N, 72]` activations the inference engine leaves
in HBM can be scored without the device->host copy the reference's test loop makes (exp126a...py:433-450).  Every per-frame
term is evaluated by ONE kernel in float64 (mpa_eval_sums_f32: thresholded TP / est / ref counts, libfmp's L2-normalised
cosine similarity, cross-entropy in bits, Euclidean distance, binary / soft accuracy, accumulated energy, and the frame-level
match counts behind mir_eval.multipitch.evaluate on a semitone grid); 16 doubles come back.  ROC-AUC and average precision
rank the whole flattened array: they are sort-based scalar statistics and are computed from one host argsort (restating
sklearn.metrics.roc_auc_score / average_precision_score).  Values are taken as float32 (what the networks emit; binary
targets are exact).  mir_eval itself is absent offline: `calculate_mpe_measures_mireval` restates its published definitions
for pitches on a semitone grid (a 0.5-semitone matching window makes the bipartite matching a set intersection, resp. a
per-pitch-class min of counts for the chroma variants) — parity unpinned for that one function."""
import numpy as np
import torch

from ... import _lib

_EPS = np.finfo(float).eps


def compute_eval_measures(I_ref, I_est):
    """libfmp.c5.compute_eval_measures (libfmp/c5/c5s2_chord_rec_template.py:238-261): P, R, F, TP, FP, FN of two binary matrices,
    counted by the same device kernel as everything else (mpa_eval_sums_f32; the counts are exact integers in float64)."""
    assert tuple(I_ref.shape) == tuple(I_est.shape), 'Dimension of input matrices must agree'
    ref = I_ref if isinstance(I_ref, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(I_ref))
    est = I_est if isinstance(I_est, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(I_est))
    ref2, est2 = ref.reshape(ref.shape[0], -1).float(), (est.reshape(est.shape[0], -1) > 0).float()
    s = eval_sums(ref2, est2, 0.5)
    TP, FP, FN = int(s[0]), int(s[1] - s[0]), int(s[2] - s[0])
    P = R = F = 0
    if TP > 0:
        P = TP / (TP + FP)
        R = TP / (TP + FN)
        F = 2 * P * R / (P + R)
    return P, R, F, TP, FP, FN


def _to_device(a):
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    if not t.is_cuda:
        t = t.to('cuda')
    return t.to(torch.float32).contiguous()


def eval_sums(targets, predictions, threshold=0.5, min_pitch=24):
    """-> the 16 float64 sums of mpa_eval_sums_f32 (see include/mpa.h) as a host ndarray."""
    targ, pred = _to_device(targets), _to_device(predictions)
    assert targ.shape == pred.shape and targ.dim() == 2, 'Error: Targets and predictions have different shape!'
    n, p = targ.shape
    ws_bytes = _lib.lib().mpa_eval_workspace(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=targ.device)
    out = torch.empty(16, dtype=torch.float64, device=targ.device)
    import ctypes
    _lib.call('eval_sums_f32', targ, pred, n, p, ctypes.c_double(float(threshold)), int(min_pitch), out, ws, _lib.usize(ws_bytes), _lib.stream_ptr())
    return out.cpu().numpy()


def _prf(s):
    TP, n_est, n_ref = s[0], s[1], s[2]
    P = R = F = 0
    if TP > 0:
        P = TP / n_est
        R = TP / n_ref
        F = 2 * P * R / (P + R)
    return P, R, F


def _ranking(targets, predictions):
    """Distinct-threshold cumulative counts (sklearn.metrics._ranking._binary_clf_curve): scores descending."""
    y = np.asarray(targets.cpu() if isinstance(targets, torch.Tensor) else targets).ravel() == 1
    sc = np.asarray(predictions.cpu() if isinstance(predictions, torch.Tensor) else predictions, dtype=np.float64).ravel()
    order = np.argsort(sc, kind='mergesort')[::-1]
    y, sc = y[order], sc[order]
    idx = np.r_[np.flatnonzero(np.diff(sc)), y.size - 1]
    tps = np.cumsum(y, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps


def roc_auc(targets, predictions):
    fps, tps = _ranking(targets, predictions)
    if tps[-1] <= 0 or fps[-1] <= 0:
        raise ValueError('Only one class present in y_true. ROC AUC score is not defined in that case.')
    fpr, tpr = np.r_[0.0, fps] / fps[-1], np.r_[0.0, tps] / tps[-1]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))


def average_precision(targets, predictions):
    fps, tps = _ranking(targets, predictions)
    precision = tps / (tps + fps)
    recall = tps / tps[-1]
    return float(np.sum(np.diff(np.r_[0.0, recall]) * precision))


def calculate_single_measure(targets, predictions, measure, threshold=0.5, save_roc_plot=False, path_output='roc.pdf', _sums=None):
    assert tuple(targets.shape) == tuple(predictions.shape), 'Error: Targets and predictions have different shape!'
    if np.mod(targets.shape[1], 12) != 0:
        print('WARNING: Shape of input is ' + str(tuple(targets.shape)) +
              ', expect features (bins) as second dimension. Please make sure that size is correct!')
    n, p = targets.shape
    if measure in ('roc_auc_measure', 'average_precision_score'):
        if save_roc_plot:
            raise NotImplementedError('ROC plots (matplotlib) are outside the hot-path scope')
        return roc_auc(targets, predictions) if measure == 'roc_auc_measure' else average_precision(targets, predictions)
    s = _sums if _sums is not None else eval_sums(targets, predictions, threshold)
    if measure == 'precision':
        return _prf(s)[0]
    if measure == 'recall':
        return _prf(s)[1]
    if measure == 'f_measure':
        return _prf(s)[2]
    if measure == 'cosine_sim':
        return s[3] / n
    if measure == 'binary_crossentropy':
        return -s[4] / (n * p)
    if measure == 'euclidean_distance':
        return s[5] / n
    if measure == 'binary_accuracy':
        return s[6] / (n * p)
    if measure == 'soft_accuracy':
        return s[7] / (n * p)
    if measure == 'accum_energy':
        return s[8] / n
    assert False, 'ERROR: Evaluation measure ' + str(measure) + ' not implemented!'


def calculate_eval_measures(targets, predictions, measures=('precision', 'recall', 'f_measure'), threshold=0.5, save_roc_plot=False,
                            path_output='roc.pdf'):
    sums = None
    if any(m not in ('roc_auc_measure', 'average_precision_score') for m in measures):
        sums = eval_sums(targets, predictions, threshold)
    return {m: calculate_single_measure(targets, predictions, m, threshold, save_roc_plot, path_output, _sums=sums) for m in measures}


def calculate_mpe_measures_mireval(targets, predictions, threshold=0.5, min_pitch=24):
    """The 14 frame-level multi-pitch scores of mir_eval.multipitch.evaluate for semitone-grid pitches (eval_metrics.py:158-189)."""
    s = eval_sums(targets, predictions, threshold, min_pitch)
    tp, n_est, n_ref, ctp = s[0], s[1], s[14], s[9]
    n_min, n_max, miss, fa = s[10], s[11], s[12], s[13]

    def scores(tp_, prefix):
        prec = tp_ / n_est if n_est > 0 else 0.0
        rec = tp_ / n_ref if n_ref > 0 else 0.0
        acc = tp_ / (n_est + n_ref - tp_) if (n_est + n_ref - tp_) > 0 else 0.0
        d = n_ref if n_ref > 0 else 1.0
        return {prefix + 'Precision': prec, prefix + 'Recall': rec, prefix + 'Accuracy': acc,
                prefix + 'Substitution Error': (n_min - tp_) / d, prefix + 'Miss Error': miss / d,
                prefix + 'False Alarm Error': fa / d, prefix + 'Total Error': (n_max - tp_) / d}

    out = scores(tp, '')
    out.update(scores(ctp, 'Chroma '))
    return out
