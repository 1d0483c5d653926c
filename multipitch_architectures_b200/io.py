"""On-disk formats and the file-level steps either side of the hot path.

  * `load_audio`       — the `librosa.load(path, sr=22050)` of notebook 01 cell 3 / notebook 02: WAV decode, channel mean, and
                         the power-of-two kaiser_best down-sampling (44.1 -> 22.05 kHz) on the device (mpa_decimate_gain_f32).
  * `load_hcqt_npy`    — the reference's feature files: float64 `[216, N, 6]` -> `[6, lead+N+trail, 216]` fp32 in HBM, transposed,
                         cast and zero-padded by one kernel (exp126a_musicnet_cnn_basic.py:255-258, 413-420).
  * `load_pitch_npy`   — annotation files `[128, N]` -> `[N, n_out]` targets, `targets[:, min_pitch:min_pitch+n_out]` (:414-416).
  * `evaluate_file` / `write_results_csv` — the per-file body of the reference test loop (:404-458): predict, save `[N, 72]`
                         predictions as .npy, score on the device, one CSV row per file + FILEWISE / FRAMEWISE means (:463-508).
No arithmetic on features or activations happens on the host."""
import csv
import os
import wave

import numpy as np
import torch

from . import _lib
from .engine import CnnStreamEngine, predict_patchwise, CONTEXT, HALF
from .libdl.metrics import calculate_eval_measures, calculate_mpe_measures_mireval

EVAL_MEASURES = ['precision', 'recall', 'f_measure', 'cosine_sim', 'binary_crossentropy', 'euclidean_distance', 'binary_accuracy',
                 'soft_accuracy', 'accum_energy', 'roc_auc_measure', 'average_precision_score']     # exp126a...py:144-146


def kaiser_best_half_taps(factor=2):
    """|j| = 0 .. 64*factor-1 taps of resampy's kaiser_best window (64 zero crossings, roll-off 0.9475937167399596,
    Kaiser beta 14.769656459379492) walked at a factor:1 ratio — librosa.load's default res_type in librosa 0.8."""
    num_zeros, beta, rolloff = 64, 14.769656459379492, 0.9475937167399596
    t = np.arange(0, factor * num_zeros) / float(factor)
    taper = np.i0(beta * np.sqrt(1.0 - (t / num_zeros) ** 2)) / np.i0(beta)
    return (rolloff * np.sinc(rolloff * t) * taper / float(factor)).astype(np.float32)


def read_wav(path):
    """PCM WAV -> (float32 [n, channels] in [-1, 1), sample rate); 16-bit (MusicNet, the shipped example) / 8 / 24 / 32-bit."""
    with wave.open(path, 'rb') as w:
        sr, nch, width, n = w.getframerate(), w.getnchannels(), w.getsampwidth(), w.getnframes()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype='<i2').astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype='<i4').astype(np.float32) / 2147483648.0
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = (((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8) >> 8).astype(np.float32) / 8388608.0
    elif width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError(f'unsupported sample width {width}')
    return x.reshape(-1, nch), sr


def load_audio(path, sr=22050, device='cuda'):
    """-> (float32 CUDA tensor [n], sr): mono mix (librosa.to_mono = channel mean) then kaiser_best resampling when the file's
    rate is a power-of-two multiple of `sr`, resampy's general table walk otherwise (no rescaling, as librosa.load)."""
    x, sr_native = read_wav(path)
    y = torch.from_numpy(np.ascontiguousarray(x.mean(axis=1, dtype=np.float32) if x.shape[1] > 1 else x[:, 0])).to(device)
    if sr is None or sr_native == sr:
        return y, sr_native
    factor = sr_native // sr
    if factor * sr != sr_native or factor & (factor - 1):
        return resample(y, sr_native, sr), sr
    taps = torch.from_numpy(kaiser_best_half_taps(factor)).to(device)
    import ctypes
    out = torch.empty(-(-y.numel() // factor), dtype=torch.float32, device=device)
    _lib.call('decimate_gain_f32', y, out, taps, taps.numel(), factor, ctypes.c_double(1.0), _lib.i64(y.numel()), _lib.stream_ptr())
    return out, sr


_WINDOWS = {}


def resampy_window(filt='kaiser_best', precision=9):
    """resampy's interpolation window (sinc_window): 64 zero crossings x 2^9 table steps for kaiser_best -> float64 [32769]."""
    key = (filt, precision)
    if key not in _WINDOWS:
        num_zeros, beta, rolloff = {'kaiser_best': (64, 14.769656459379492, 0.9475937167399596), 'kaiser_fast': (16, 8.555504641634386, 0.85)}[filt]
        n = (2 ** precision) * num_zeros
        t = np.linspace(0, num_zeros, num=n + 1, endpoint=True)
        taper = np.i0(beta * np.sqrt(np.clip(1.0 - (t / num_zeros) ** 2, 0.0, None))) / np.i0(beta)
        _WINDOWS[key] = (rolloff * np.sinc(rolloff * t) * taper, 2 ** precision)
    return _WINDOWS[key]


def resample(y, sr_orig, sr_new, res_type='kaiser_best', scale=False):
    """librosa.resample for any ratio on the device (mpa_resample_f32: resampy's table walk, one thread per output sample).
    y: 1-D float32 CUDA tensor -> float32 CUDA tensor of ceil(n * sr_new / sr_orig) samples."""
    import ctypes
    ratio = float(sr_new) / float(sr_orig)
    win, num_table = resampy_window(res_type)
    if ratio < 1:
        win = win * ratio
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    dev = y.device
    n_out = int(np.ceil(y.numel() * ratio))
    out = torch.empty(n_out, dtype=torch.float32, device=dev)
    gain = 1.0 / np.sqrt(ratio) if scale else 1.0
    # resampy's time register: 1/ratio accumulated sequentially in float64 (np.cumsum adds in order) — see the kernel's comment
    n_real = int(y.numel() * ratio)
    times = np.zeros(max(n_real, 1), dtype=np.float64)
    if n_real > 1:
        times[1:] = np.cumsum(np.full(n_real - 1, 1.0 / ratio, dtype=np.float64))
    _lib.call('resample_f32', y.contiguous(), out, torch.from_numpy(win).to(dev), torch.from_numpy(delta).to(dev), torch.from_numpy(times).to(dev),
              len(win), num_table,
              ctypes.c_double(ratio), ctypes.c_double(gain), _lib.i64(y.numel()), _lib.i64(n_out), _lib.stream_ptr())
    return out


def load_hcqt_npy(path, lead=0, trail=0, device='cuda'):
    """Reference feature file (float64 [F, N, C]) -> fp32 CUDA [C, lead+N+trail, F]; lead = 37, trail = 38 is the inference padding."""
    a = np.load(path)
    if a.ndim != 3 or a.dtype != np.float64:
        raise ValueError(f'{path}: expected a float64 [bins, frames, harmonics] array, got {a.dtype} {a.shape}')
    F, N, C = a.shape
    src = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    out = torch.empty(C, lead + N + trail, F, dtype=torch.float32, device=device)
    _lib.call('hcqt_npy_to_frames_f64', src, out, F, N, C, lead, trail, _lib.stream_ptr())
    return out


def load_pitch_npy(path, min_pitch=24, n_out=72, device='cuda'):
    """Reference annotation file ([128, N] or [12, N]) -> fp32 CUDA [N, n_out] (exp126a...py:414-416)."""
    t = np.load(path).T
    if n_out != 12:
        t = t[:, min_pitch:min_pitch + n_out]
    return torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32)).to(device)


def predict_hcqt(model, hcqt, engine=None, batch=50):
    """[C, N, F] fp32 CUDA (un-padded) -> [N, n_out] activations: the streaming engine for the CNN family on the tensor-core
    formats, the reference-shaped batches of `batch` consecutive patches otherwise."""
    if engine is not None:
        return engine.predict_hcqt(hcqt)
    out = predict_patchwise(model, hcqt, batch=batch)
    return out[0] if isinstance(out, tuple) else out


def evaluate_file(model, path_hcqt, path_annot, dir_predictions=None, eval_measures=EVAL_MEASURES, eval_thresh=0.4, min_pitch=24,
                  n_out=72, max_frames=None, engine=None, batch=50):
    """One iteration of the reference's per-file test loop -> ordered dict row (Filename, measures, mir_eval-style scores)."""
    hcqt = load_hcqt_npy(path_hcqt)
    targ = load_pitch_npy(path_annot, min_pitch, n_out)
    if max_frames is not None:                      # test_subset 1: the first 3920 frames (:420-422)
        hcqt, targ = hcqt[:, :max_frames].contiguous(), targ[:max_frames].contiguous()
    pred = predict_hcqt(model, hcqt, engine=engine, batch=batch)
    assert tuple(pred.shape) == tuple(targ.shape), \
        'Shape mismatch! Target shape: ' + str(tuple(targ.shape)) + ', Pred. shape: ' + str(tuple(pred.shape))
    fn = os.path.basename(path_hcqt)
    if dir_predictions is not None:
        os.makedirs(dir_predictions, exist_ok=True)
        np.save(os.path.join(dir_predictions, fn[:-4] + '.npy'), pred.cpu().numpy().astype(np.float64))
    row = {'Filename': fn}
    row.update(calculate_eval_measures(targ, pred, measures=eval_measures, threshold=eval_thresh))
    row.update(calculate_mpe_measures_mireval(targ, pred, threshold=eval_thresh, min_pitch=min_pitch))
    row['_kframes'] = targ.shape[0] / 1000
    return row


def write_results_csv(rows, path_output):
    """Per-file rows + 'FILEWISE MEAN' + 'FRAMEWISE MEAN' (frame-count weighted) in the reference's DataFrame.to_csv layout."""
    keys = [k for k in rows[0] if k not in ('Filename', '_kframes')]
    vals = np.array([[r[k] for k in keys] for r in rows], dtype=float)
    kf = np.array([r['_kframes'] for r in rows], dtype=float)
    table = [[r['Filename']] + list(v) for r, v in zip(rows, vals)]
    table.append(['FILEWISE MEAN'] + list(vals.sum(0) / len(rows)))
    table.append(['FRAMEWISE MEAN'] + list((vals * kf[:, None]).sum(0) / kf.sum()))
    with open(path_output, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([''] + ['Filename'] + keys)
        for i, r in enumerate(table):
            w.writerow([i] + r)
    return table


class HostPrefetcher:
    """Double-buffered host -> device feed for the training loop (the reference copies every batch synchronously on the compute stream,
    exp126a...py:317-319).  `batches`: an iterable of tuples of PINNED host tensors; iteration yields the same tuples as CUDA tensors
    while the NEXT batch is already being copied on a side stream.  A yielded batch stays valid until the one after the next is requested
    (two device buffers per tensor, reused)."""

    def __init__(self, batches, device='cuda'):
        self.it, self.dev = iter(batches), torch.device(device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.slots, self.k, self.next = [None, None], 0, None
        self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.next = None
            return
        slot = self.slots[self.k]
        if slot is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot[0], host)):
            slot = ([torch.empty(h.shape, dtype=h.dtype, device=self.dev) for h in host], torch.cuda.Event())
            self.slots[self.k] = slot
        bufs, ev = slot
        self.stream.wait_stream(torch.cuda.current_stream(self.dev))     # the consumer of this slot's previous contents has been enqueued
        with torch.cuda.stream(self.stream):
            for d, h in zip(bufs, host):
                d.copy_(h, non_blocking=True)
            ev.record(self.stream)
        self.next = (bufs, ev)
        self.k ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.next is None:
            raise StopIteration
        bufs, ev = self.next
        torch.cuda.current_stream(self.dev).wait_event(ev)
        self._issue()
        return tuple(bufs)
