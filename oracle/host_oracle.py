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
