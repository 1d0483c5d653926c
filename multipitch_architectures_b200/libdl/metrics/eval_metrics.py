"""Thresholded frame-level precision / recall / F-measure with the reference's call signature
(/root/reference/libdl/metrics/eval_metrics.py:8-62, which delegates to libfmp.c5.compute_eval_measures,
/root/reference/libfmp/c5/c5s2_chord_rec_template.py:238-261).  Integer counting on the host."""
import numpy as np


def compute_eval_measures(I_ref, I_est):
    assert I_ref.shape == I_est.shape, 'Dimension of input matrices must agree'
    TP = np.sum(np.logical_and(I_ref, I_est))
    FP = np.sum(I_est > 0, axis=None) - TP
    FN = np.sum(I_ref > 0, axis=None) - TP
    P = R = F = 0
    if TP > 0:
        P = TP / (TP + FP)
        R = TP / (TP + FN)
        F = 2 * P * R / (P + R)
    return P, R, F, TP, FP, FN


def calculate_single_measure(targ, pred, measure, threshold=0.5):
    assert targ.shape == pred.shape, 'Error: Targets and predictions have different shape!'
    P, R, F, TP, FP, FN = compute_eval_measures(targ, pred >= threshold)
    if measure == 'precision':
        return P
    if measure == 'recall':
        return R
    if measure == 'f_measure':
        return F
    raise NotImplementedError(f"measure '{measure}' is outside the hot-path scope (SURVEY.md 8f row 4)")


def calculate_eval_measures(targ, pred, measures=('precision', 'recall', 'f_measure'), threshold=0.5, save_roc_plot=False, path_output='', plot_title=''):
    return {m: calculate_single_measure(targ, pred, m, threshold) for m in measures}
