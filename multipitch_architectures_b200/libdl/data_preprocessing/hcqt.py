"""HCQT feature extraction with the reference's API (/root/reference/libdl/data_preprocessing/hcqt.py:9-164,
205-272) on a B200: audio in, `[n_bins, n_frames, n_harmonics]` float64 out; all signal arithmetic runs in the
libmpa CUDA kernels (decimator chain, fused FFT + constant-Q contraction, tuning estimate)."""
import functools

import numpy as np
import torch

from ... import _lib
from . import _filterbank as FB

C1_HZ = 32.70319566257483      # librosa.note_to_hz('C1'), the reference's default fmin


def compute_hopsize_cqt(fs_cqt_target, fs=22050, num_octaves=7):
    """hcqt.py:9-30 — CQT hop size (multiple of 2^(num_octaves-1)) approximating a target frame rate."""
    factor = 2 ** (num_octaves - 1)
    n = np.round((fs / fs_cqt_target) / factor)
    hopsize_cqt = int(max(1, factor * n))
    return hopsize_cqt, fs / hopsize_cqt


def _harmonic_plan(num_harmonics, num_subharmonics):
    """Which base CQT serves each (sub)harmonic: harmonics related by a power of two share one CQT (hcqt.py:129-148)."""
    list_h = [1.0 / (s + 1) for s in range(num_subharmonics, 0, -1)] + [float(h) for h in range(1, num_harmonics + 1)]
    base = []
    for h in list_h:
        for b in base:
            if float(np.log2(h / b)).is_integer():
                base.append(b)
                break
        else:
            base.append(h)
    return list_h, base


class HCQTPlan:
    """Device-resident filter tables + launch schedule for one HCQT configuration."""

    def __init__(self, fs, fmin, hop, bins_per_octave, num_octaves, num_harmonics, num_subharmonics, device):
        self.fs, self.hop, self.bpo, self.num_octaves = fs, hop, bins_per_octave, num_octaves
        self.H = num_harmonics + num_subharmonics
        self.n_bins = bins_per_octave * num_octaves
        self.device = torch.device(device)
        list_h, base = _harmonic_plan(num_harmonics, num_subharmonics)
        tabs = FB.build_tables(fs, hop, fmin, bins_per_octave, num_octaves, list_h, base)
        self.levels = []
        for (lv, n_fft), t in sorted(tabs.items()):
            self.levels.append(dict(
                level=lv, n_fft=n_fft, n_rows=t['dest'].shape[0], n_dest=t['dest'].shape[1],
                basis=torch.from_numpy(np.ascontiguousarray(t['basis']).view(np.float32)).to(self.device),
                start=torch.from_numpy(t['start']).to(self.device), scale=torch.from_numpy(t['scale']).to(self.device),
                dest=torch.from_numpy(t['dest']).to(self.device)))
        self.max_level = max(l['level'] for l in self.levels)
        self.taps = torch.from_numpy(FB.kaiser_fast_half_taps()).to(self.device)
        n = np.arange(2048)
        self.hann2048 = torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * n / 2048)).astype(np.float32)).to(self.device)
        self.tunings = FB.tuning_values()

    def run(self, y, tuning_idx=None):
        """y: 1-D float32 CUDA tensor -> (hcqt [H, n_frames, n_bins] fp32 CUDA, tuning_idx int32[1] CUDA)."""
        if y.dim() != 1 or not y.is_cuda or y.dtype != torch.float32:
            raise _lib.MpaError('HCQTPlan.run expects a 1-D float32 CUDA tensor')
        y = y.contiguous()
        n = y.numel()
        st = _lib.stream_ptr()
        n_frames = n // self.hop + 1
        if tuning_idx is None:
            tuning_idx = torch.empty(1, dtype=torch.int32, device=y.device)
            ws_bytes = _lib.lib().mpa_tuning_workspace(n // 512 + 1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device)
            _lib.call('estimate_tuning_f32', y, _lib.i64(n), self.hann2048, float(self.fs), self.bpo, tuning_idx, ws,
                      _lib.usize(ws_bytes), st)
        sig = [y]
        for _ in range(self.max_level):
            prev = sig[-1]
            nxt = torch.empty((prev.numel() + 1) // 2, dtype=torch.float32, device=y.device)
            _lib.call('decimate2_f32', prev, nxt, self.taps, _lib.i64(prev.numel()), st)
            sig.append(nxt)
        out = torch.empty(self.H, n_frames, self.n_bins, dtype=torch.float32, device=y.device)
        for L in self.levels:
            s = sig[L['level']]
            _lib.call('cqt_level_f32', s, _lib.i64(s.numel()), L['n_fft'], self.hop >> L['level'], n_frames, L['basis'], L['start'],
                      L['scale'], L['n_rows'], FB.BAND, tuning_idx, L['dest'], L['n_dest'], out, n_frames, self.n_bins, st)
        return out, tuning_idx


@functools.lru_cache(maxsize=8)
def get_plan(fs, fmin, hop, bins_per_octave, num_octaves, num_harmonics, num_subharmonics, device):
    return HCQTPlan(fs, fmin, hop, bins_per_octave, num_octaves, num_harmonics, num_subharmonics, device)


def compute_efficient_hcqt(f_audio, fs=22050, fmin=C1_HZ, fs_hcqt_target=91, bins_per_octave=60, num_octaves=6,
                           num_harmonics=5, num_subharmonics=1, center_bins=True, device='cuda'):
    """Drop-in for the reference function of the same name (hcqt.py:89-164).

    Returns (f_hcqt float64 ndarray [n_bins, n_frames, H], fs_hcqt, hopsize_cqt).  Harmonics related by powers of two
    share one CQT, exactly like the reference; channel order: sub-harmonics (lowest first), then h = 1..num_harmonics."""
    num_octaves_eff = num_octaves + int(np.ceil(np.log2(num_subharmonics + 1) + np.log2(num_harmonics)))
    hopsize_cqt, _ = compute_hopsize_cqt(fs_hcqt_target, fs=fs, num_octaves=num_octaves_eff)
    fs_hcqt = fs / hopsize_cqt
    assert np.mod(bins_per_octave, 12) == 0, 'Error: bins_per_octave no multiple of 12'
    bins_per_semitone = int(bins_per_octave / 12)
    if center_bins:
        fmin = fmin / 2 ** ((bins_per_semitone - 1) / (2 * bins_per_octave))
    plan = get_plan(fs, float(fmin), hopsize_cqt, bins_per_octave, num_octaves, num_harmonics, num_subharmonics, str(device))
    y = torch.from_numpy(np.ascontiguousarray(f_audio, dtype=np.float32)).to(plan.device)
    out, _ = plan.run(y)
    f_hcqt = out.permute(2, 1, 0).contiguous().cpu().numpy().astype(np.float64)
    return f_hcqt, fs_hcqt, hopsize_cqt


def estimate_tuning(f_audio, fs=22050, bins_per_octave=36, device='cuda'):
    """The whole-file tuning scalar the reference obtains from librosa.estimate_tuning (hcqt.py:122)."""
    y = torch.from_numpy(np.ascontiguousarray(f_audio, dtype=np.float32)).to(device)
    n = y.numel()
    idx = torch.empty(1, dtype=torch.int32, device=y.device)
    ws_bytes = _lib.lib().mpa_tuning_workspace(n // 512 + 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=y.device)
    k = np.arange(2048)
    hann = torch.from_numpy((0.5 - 0.5 * np.cos(2 * np.pi * k / 2048)).astype(np.float32)).to(y.device)
    _lib.call('estimate_tuning_f32', y, _lib.i64(n), hann, float(fs), bins_per_octave, idx, ws, _lib.usize(ws_bytes), _lib.stream_ptr())
    return float(FB.tuning_values()[int(idx.item())])


def compute_annotation_array_nooverlap(note_events, f_hcqt, fs_hcqt, annot_type='pitch_class', shorten=1.0):
    """Note list -> binary piano-roll with collision / zero-length fixes (hcqt.py:205-272).  Integer frame arithmetic on
    the host (1956 events for the shipped example): every event lasts >= 1 frame and events that would vanish push the
    events starting / ending on the same frame one frame later."""
    if annot_type == 'pitch_class':
        height = 12
    elif annot_type == 'pitch':
        height = 128
    elif annot_type == 'instruments':
        height = 1
    else:
        assert False, ['annotation type ' + str(annot_type) + ' not valid!']
    n_frames = f_hcqt.shape[1]
    annot = np.zeros((height, n_frames))
    if shorten != 1.0:
        note_events[:, 1] = note_events[:, 0] + shorten * (note_events[:, 1] - note_events[:, 0])
    frames = np.floor(np.asarray(note_events, dtype=np.float64)[:, :2] * fs_hcqt).astype(np.int64)
    on, off = frames[:, 0].copy(), frames[:, 1].copy()
    gone = np.flatnonzero(off - on < 1)
    for v in np.unique(off[gone]):
        on[on == v] += 1
        off[off == v] += 1
    on[gone] -= 1
    on[np.flatnonzero(off - on < 1)] -= 1
    assert not np.any(off - on < 1), 'still events of length<1 after correction!'
    pitches = np.asarray(note_events)[:, 2]
    for a, b, p in zip(on, off, pitches):
        row = int(np.mod(p, 12)) if annot_type == 'pitch_class' else (int(p) if annot_type == 'pitch' else 0)
        annot[row, a:b] = 1
    return annot
