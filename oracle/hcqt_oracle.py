"""ORACLE (test infrastructure only — never imported by the product path).

NumPy/SciPy CPU restatement of the HCQT feature path
`libdl/data_preprocessing/hcqt.py:89-164` (compute_efficient_hcqt) of the reference.

PARITY UNPINNED for the arithmetic core: the reference delegates all arithmetic to
`librosa==0.8.*` (environment.yml:31) + `resampy`, neither of which is vendored
under /root/reference nor installable here (no network), and the example HCQT
array is a missing blob (/root/reference/.MISSING_LARGE_BLOBS:2).  What follows is
a restatement of the *published* librosa-0.8 algorithm (librosa/core/constantq.py
`cqt`→`vqt(gamma=0)`, `__cqt_filter_fft`, `__cqt_response`, `__trim_stack`;
librosa/filters.py `constant_q`; librosa/util `sparsify_rows`, `pad_center`,
`normalize`; librosa/core/pitch.py `estimate_tuning`, `piptrack`, `pitch_tuning`;
resampy `resample` with the `kaiser_fast` window: 16 zero-crossings, Kaiser
beta=8.555504641634386, roll-off 0.85) anchored on the reference's own call
sites (hcqt.py:122,157-158).  The reference's own Python around those calls
(hop size, harmonic bookkeeping, slicing) IS pinned: tests/golden/make_golden.py
runs the reference functions with this module's `cqt`/`estimate_tuning`
substituted for librosa and the results must agree bit for bit.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this file.
"""
import numpy as np
import scipy.signal
import scipy.special

C1_HZ = 32.70319566257483          # librosa.note_to_hz('C1')
BW_FASTEST = 0.85                  # librosa.core.audio.BW_FASTEST
HANN_ENBW = 1.50018310546875       # librosa.filters.WINDOW_BANDWIDTHS['hann']


# ----------------------------------------------------------------------------- hop size
def compute_hopsize_cqt(fs_cqt_target, fs=22050, num_octaves=7):
    """hcqt.py:9-30."""
    factor = 2 ** (num_octaves - 1)
    n = np.round((fs / fs_cqt_target) / factor)
    hop = int(max(1, factor * n))
    return hop, fs / hop


# ----------------------------------------------------------------------------- resampy 2^c:1
RESAMPY_FILTERS = {                     # resampy's published window parameters: (zero crossings, Kaiser beta, roll-off)
    'kaiser_fast': (16, 8.555504641634386, 0.85),
    'kaiser_best': (64, 14.769656459379492, 0.9475937167399596),
}


def _kaiser_fast_half(factor=2, filt='kaiser_fast'):
    """Taps h[|j|], |j| = 0 .. 16*factor-1, of resampy's kaiser_fast filter evaluated at a factor:1 ratio
    (64*factor taps for kaiser_best, the default of librosa.load).

    resampy builds interp_win = rolloff*sinc(rolloff*t)*kaiser(t), t in [0,num_zeros] at 2^precision = 512
    steps per zero crossing and, for sample_ratio = 1/factor, walks it with index_step = 512/factor, i.e.
    t = j/factor; interp_win is pre-multiplied by sample_ratio.  Its wing loops stop at
    (len(interp_win) - offset) // index_step = 16*factor (left wing, incl. the centre tap) and 16*factor-1
    (right wing) taps, i.e. |j| <= 16*factor-1: the table's end point t = 16 is never read."""
    num_zeros, beta, rolloff = RESAMPY_FILTERS[filt]
    j = np.arange(0, num_zeros * factor, dtype=np.float64)
    t = j / float(factor)
    # right half of scipy.signal.kaiser(2n+1, beta) sampled at t/num_zeros
    taper = scipy.special.i0(beta * np.sqrt(np.clip(1.0 - (t / num_zeros) ** 2, 0.0, None))) / scipy.special.i0(beta)
    return rolloff * np.sinc(rolloff * t) * taper / float(factor)


def _kaiser_fast_halfband():
    return _kaiser_fast_half(2)


def resample_pow2(y, factor, filt='kaiser_fast', scale=True):
    """librosa.resample(y, sr, sr/factor, res_type=filt, scale=scale) for 1-D float32 y, factor = 2^c
    (librosa.load resamples with kaiser_best and scale=False).

    resampy output length floor(n/factor); librosa fixes the length to ceil(n/factor) (zero pad) and
    divides by sqrt(ratio) = multiplies by sqrt(factor).  Output sample t sits at input sample factor*t;
    taps outside [0,n) are skipped (no padding).  resampy accumulates in the dtype of y."""
    y = np.asarray(y)
    n = y.shape[0]
    n_out = n // factor
    half = _kaiser_fast_half(factor, filt)
    L = half.shape[0] - 1
    ypad = np.concatenate([np.zeros(L, y.dtype), y, np.zeros(L + factor, y.dtype)]).astype(np.float64)
    taps = np.concatenate([half[::-1], half[1:]])           # j = -L..L
    # out[t] = sum_j taps[j] * y[factor*t + j]
    full = np.convolve(ypad, taps[::-1], mode='valid')      # full[m] = sum_j taps[j]*ypad[m + j + L] -> centre m
    out = full[0:factor * n_out:factor]
    out = (out * (np.sqrt(float(factor)) if scale else 1.0)).astype(y.dtype)
    n_fix = int(np.ceil(n / float(factor)))
    if n_fix > n_out:
        out = np.concatenate([out, np.zeros(n_fix - n_out, y.dtype)])
    return out


def resampy_window(filt='kaiser_best', precision=9):
    """resampy.filters.sinc_window for the published kaiser_fast / kaiser_best parameters -> (interp_win float64, 2^precision)."""
    num_zeros, beta, rolloff = RESAMPY_FILTERS[filt]
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = scipy.signal.windows.kaiser(2 * n + 1, beta)[n:]
    return sinc_win * taper, num_bits


def resample_general(y, sr_orig, sr_new, filt='kaiser_best', scale=False):
    """librosa.resample(y, sr_orig, sr_new, res_type=filt, scale=scale) for any ratio: resampy's table walk (resample_f) vectorised over
    the output samples (one pass per tap); float64 accumulation, output length fixed to ceil(n * ratio)."""
    y = np.asarray(y)
    ratio = float(sr_new) / float(sr_orig)
    interp_win, num_table = resampy_window(filt)
    if ratio < 1:
        interp_win = interp_win * ratio
    delta = np.zeros_like(interp_win)
    delta[:-1] = np.diff(interp_win)
    sc = min(1.0, ratio)
    index_step = int(sc * num_table)
    n_in = y.shape[0]
    n_real = int(n_in * ratio)
    # resampy's time register is ACCUMULATED (time_register += 1/ratio per output sample); np.cumsum adds sequentially, so this is
    # the same float64 sequence.  It matters: index_step is truncated to an integer, which makes the output discontinuous where the
    # register crosses an integer.
    t = np.zeros(n_real, dtype=np.float64)
    if n_real > 1:
        t[1:] = np.cumsum(np.full(n_real - 1, 1.0 / ratio, dtype=np.float64))
    n = t.astype(np.int64)
    out = np.zeros(n_real, dtype=np.float64)
    nwin = len(interp_win)
    yd = y.astype(np.float64)
    for wing in (0, 1):
        frac = sc * (t - n) if wing == 0 else sc - sc * (t - n)
        index_frac = frac * num_table
        offset = index_frac.astype(np.int64)
        eta = index_frac - offset
        lim = (nwin - offset) // index_step
        cnt = np.minimum(n + 1, lim) if wing == 0 else np.minimum(n_in - n - 1, lim)
        for i in range(int(cnt.max()) if len(cnt) else 0):
            m = cnt > i
            idx = offset[m] + i * index_step
            src = n[m] - i if wing == 0 else n[m] + i + 1
            out[m] += (interp_win[idx] + eta[m] * delta[idx]) * yd[src]
    if scale:
        out /= np.sqrt(ratio)
    n_fix = int(np.ceil(n_in * ratio))
    res = np.zeros(n_fix, dtype=y.dtype)
    res[:min(n_fix, n_real)] = out[:n_fix].astype(y.dtype)
    return res


def resample_2to1(y):
    return resample_pow2(y, 2)


# ----------------------------------------------------------------------------- filter bank
def constant_q_lengths(sr, fmin, n_bins, bins_per_octave):
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    Q = 1.0 / alpha
    freqs = fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)
    return Q * sr / freqs, freqs


def cqt_filter_fft(sr, fmin, n_bins, bins_per_octave, sparsity=0.01):
    """librosa __cqt_filter_fft + filters.constant_q (window='hann', filter_scale=1, norm=1, pad_fft)."""
    lengths, freqs = constant_q_lengths(sr, fmin, n_bins, bins_per_octave)
    max_len = int(2.0 ** np.ceil(np.log2(lengths.max())))
    n_fft = max_len
    basis = np.zeros((n_bins, n_fft), dtype=np.complex64)
    for i, (ilen, freq) in enumerate(zip(lengths, freqs)):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(t * 1j * 2 * np.pi * freq / sr)
        sig = sig * scipy.signal.get_window('hann', len(sig), fftbins=True)
        sig = sig / np.sum(np.abs(sig))
        lpad = int((n_fft - len(sig)) // 2)
        basis[i, lpad:lpad + len(sig)] = sig
    basis = (basis * (lengths[:, None] / float(n_fft))).astype(np.complex64)
    fft_basis = np.fft.fft(basis.astype(np.complex128), n=n_fft, axis=1)[:, :n_fft // 2 + 1]
    # util.sparsify_rows(quantile=sparsity)
    mags = np.abs(fft_basis)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cum = np.cumsum(mag_sort / norms, axis=1)
    thr_idx = np.argmin(cum < sparsity, axis=1)
    out = np.zeros_like(fft_basis, dtype=np.complex64)
    for i, j in enumerate(thr_idx):
        keep = mags[i] >= mag_sort[i, j]
        out[i, keep] = fft_basis[i, keep]
    return out, n_fft, lengths


def stft_ones(y, n_fft, hop):
    """librosa.stft(window='ones', center=True, pad_mode='reflect') -> complex64 [n_fft/2+1, 1+len(y)//hop]."""
    ypad = np.pad(y, n_fft // 2, mode='reflect')
    n_frames = 1 + (len(ypad) - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = ypad[idx]                                      # [n_frames, n_fft]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex64)


def cqt(y, sr=22050, hop_length=512, fmin=None, n_bins=84, bins_per_octave=12, tuning=0.0,
        sparsity=0.01, return_plan=False):
    """librosa.cqt (0.8.x: vqt with gamma=0), defaults filter_scale=1, norm=1, window='hann',
    scale=True, pad_mode='reflect', res_type=None.  y float32 -> complex64 [n_bins, 1+len(y)//hop]."""
    y = np.asarray(y, dtype=np.float32)
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    alpha = 2.0 ** (1.0 / bins_per_octave) - 1.0
    fmin = fmin * 2.0 ** (tuning / bins_per_octave)
    freqs = fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)
    top = freqs[-bins_per_octave:]
    fmin_t, fmax_t = np.min(top), np.max(top)
    Q = 1.0 / alpha
    filter_cutoff = fmax_t * (1 + 0.5 * HANN_ENBW / Q)
    nyquist = sr / 2.0
    res_fast = filter_cutoff < BW_FASTEST * nyquist
    # __early_downsample: count = min(max(0, ceil(log2(BW_FASTEST*nyq/cutoff))-1-1), max(0, num_twos(hop)-n_oct+1))
    d1 = max(0, int(np.ceil(np.log2(BW_FASTEST * nyquist / filter_cutoff)) - 1) - 1)
    num_twos = 0
    h = hop_length
    while h % 2 == 0 and h > 0:
        num_twos += 1
        h //= 2
    d2 = max(0, num_twos - n_octaves + 1)
    early = min(d1, d2)
    if early > 0 and res_fast:
        # __early_downsample: ONE resample by 2^early (not a chain of 2:1 steps); scale=True keeps the energy, so
        # the filter lengths of the final normalisation are those at the reduced rate
        y = resample_pow2(y, 2 ** early)
        hop_length //= 2 ** early
        sr = sr / float(2 ** early)
        num_twos -= early
    plan = []                                               # (level, n_fft, fmin_oct, scale)
    resp = []
    if not res_fast:
        fb, n_fft, _ = cqt_filter_fft(sr, fmin_t, n_filters, bins_per_octave, sparsity)
        resp.append(fb.dot(stft_ones(y, n_fft, hop_length)))
        plan.append((0, n_fft, fmin_t, 1.0))
        fmin_t /= 2
        fmax_t /= 2
        n_octaves -= 1
    assert num_twos >= n_octaves - 1
    my_y, my_sr, my_hop = y, float(sr), hop_length
    for i in range(n_octaves):
        if i > 0:
            my_y = resample_2to1(my_y)
            my_sr /= 2.0
            my_hop //= 2
        fb, n_fft, _ = cqt_filter_fft(my_sr, fmin_t * 2.0 ** -i, n_filters, bins_per_octave, sparsity)
        fb = fb * np.float32(np.sqrt(2 ** i))
        resp.append(fb.dot(stft_ones(my_y, n_fft, my_hop)))
        plan.append((i, n_fft, fmin_t * 2.0 ** -i, float(np.sqrt(2 ** i))))
    # __trim_stack
    max_col = min(c.shape[-1] for c in resp)
    C = np.empty((n_bins, max_col), dtype=np.complex64)
    end = n_bins
    for c in resp:
        n_oct = c.shape[0]
        if end < n_oct:
            C[:end] = c[-end:, :max_col]
        else:
            C[end - n_oct:end] = c[:, :max_col]
        end -= n_oct
    lengths, _ = constant_q_lengths(sr, fmin, n_bins, bins_per_octave)
    C /= np.sqrt(lengths[:, None]).astype(np.float32)
    if return_plan:
        return C, plan
    return C


# ----------------------------------------------------------------------------- tuning
def stft_hann(y, n_fft=2048, hop=512):
    ypad = np.pad(np.asarray(y, np.float32), n_fft // 2, mode='reflect')
    n_frames = 1 + (len(ypad) - n_fft) // hop
    win = scipy.signal.get_window('hann', n_fft, fftbins=True).astype(np.float32)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    return np.fft.rfft(ypad[idx] * win[None, :], axis=1).T.astype(np.complex64)


def piptrack(y, sr=22050, n_fft=2048, hop=512, fmin=150.0, fmax=4000.0, threshold=0.1):
    S = np.abs(stft_hann(y, n_fft, hop))                    # float32 [1025, n_frames]
    fmax = min(fmax, sr / 2.0)
    fft_freqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    avg = 0.5 * (S[2:] - S[:-2])
    shift = 2 * S[1:-1] - S[2:] - S[:-2]
    shift = avg / (shift + (np.abs(shift) < np.finfo(np.float32).tiny))
    avg = np.pad(avg, ([1, 1], [0, 0]), mode='constant')
    shift = np.pad(shift, ([1, 1], [0, 0]), mode='constant')
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = ((fmin <= fft_freqs) & (fft_freqs < fmax)).reshape((-1, 1))
    ref_value = threshold * np.max(S, axis=0)
    X = S * (S > ref_value)
    Xp = np.pad(X, ([1, 1], [0, 0]), mode='edge')
    localmax = (X > Xp[:-2]) & (X >= Xp[2:])
    idx = np.argwhere(freq_mask & localmax)
    pitches[idx[:, 0], idx[:, 1]] = (idx[:, 0] + shift[idx[:, 0], idx[:, 1]]) * float(sr) / n_fft
    mags[idx[:, 0], idx[:, 1]] = S[idx[:, 0], idx[:, 1]] + dskew[idx[:, 0], idx[:, 1]]
    return pitches, mags


def pitch_tuning(frequencies, resolution=0.01, bins_per_octave=12):
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        return 0.0
    octs = np.log2(frequencies / np.float32(440.0 / 16))
    residual = np.mod(bins_per_octave * octs, 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return tuning[np.argmax(counts)]


def estimate_tuning(y, sr=22050, n_fft=2048, resolution=0.01, bins_per_octave=12):
    pitch, mag = piptrack(y, sr=sr, n_fft=n_fft)
    pitch_mask = pitch > 0
    threshold = np.median(mag[pitch_mask]) if pitch_mask.any() else 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], resolution=resolution,
                        bins_per_octave=bins_per_octave)


# ----------------------------------------------------------------------------- HCQT
def harmonic_plan(num_harmonics=5, num_subharmonics=1):
    """hcqt.py:129-148: which base CQT serves each (sub)harmonic.  Returns list_harmonics, base_harmonics."""
    list_h = [1.0 / (s + 1) for s in range(num_subharmonics, 0, -1)] + [float(h) for h in range(1, num_harmonics + 1)]
    base = [0.0] * len(list_h)
    base[0] = 1.0 / (num_subharmonics + 1)
    for n in range(1, len(list_h)):
        for b in base:
            if b == 0.0:
                base[n] = list_h[n]
                break
            if np.mod(np.log2(list_h[n] / b), 1) == 0:
                base[n] = b
                break
    return list_h, base


def compute_efficient_hcqt(f_audio, fs=22050, fmin=C1_HZ, fs_hcqt_target=91, bins_per_octave=60,
                           num_octaves=6, num_harmonics=5, num_subharmonics=1, center_bins=True,
                           tuning_est=None):
    """hcqt.py:89-164 -> (f_hcqt float64 [n_bins, n_frames, H], fs_hcqt, hopsize)."""
    f_audio = np.asarray(f_audio, dtype=np.float32)
    n_oct_eff = num_octaves + int(np.ceil(np.log2(num_subharmonics + 1) + np.log2(num_harmonics)))
    hop, _ = compute_hopsize_cqt(fs_hcqt_target, fs=fs, num_octaves=n_oct_eff)
    fs_hcqt = fs / hop
    assert bins_per_octave % 12 == 0, 'Error: bins_per_octave no multiple of 12'
    bps = bins_per_octave // 12
    if center_bins:
        fmin = fmin / 2 ** ((bps - 1) / (2 * bins_per_octave))
    if tuning_est is None:
        tuning_est = estimate_tuning(f_audio, sr=fs, bins_per_octave=bins_per_octave)
    fmin_tuned = fmin * 2 ** (tuning_est / bins_per_octave)
    n_frames = int(np.floor(f_audio.shape[0] / hop)) + 1
    n_bins = bins_per_octave * num_octaves
    H = num_harmonics + num_subharmonics
    f_hcqt = np.zeros((n_bins, n_frames, H))
    list_h, base = harmonic_plan(num_harmonics, num_subharmonics)
    for b in sorted(set(base)):
        members = [i for i in range(H) if base[i] == b]
        add_oct = int(np.ceil(np.log2(list_h[max(members)] / b)))
        C = cqt(f_audio, sr=fs, hop_length=hop, fmin=fmin_tuned * b,
                n_bins=(num_octaves + add_oct) * bins_per_octave, bins_per_octave=bins_per_octave, tuning=0.0)
        for i in members:
            factor = int(np.log2(list_h[i] / b))
            f_hcqt[:, :, i] = np.abs(C[factor * bins_per_octave:(factor + num_octaves) * bins_per_octave, :])
    return f_hcqt, fs_hcqt, hop


def compute_hcqt(f_audio, fs=22050, fmin=C1_HZ, fs_hcqt_target=91, bins_per_octave=60, num_octaves=6,
                 num_harmonics=5, num_subharmonics=1, center_bins=True, tuning_est=None):
    """hcqt.py:34-85: one individual CQT per (sub)harmonic -> (f_hcqt float64 [n_bins, n_frames, H], fs_hcqt, hopsize).
    The hop size follows from num_octaves alone (hcqt.py:54), so the low harmonics run into librosa's early
    down-sampling and the top ones into its full-rate top octave."""
    f_audio = np.asarray(f_audio, dtype=np.float32)
    hop, _ = compute_hopsize_cqt(fs_hcqt_target, fs=fs, num_octaves=num_octaves)
    fs_hcqt = fs / hop
    n_bins = num_octaves * bins_per_octave
    assert bins_per_octave % 12 == 0, 'Error: bins_per_octave no multiple of 12'
    bps = bins_per_octave // 12
    if center_bins:
        fmin = fmin / 2 ** ((bps - 1) / (2 * bins_per_octave))
    if tuning_est is None:
        tuning_est = estimate_tuning(f_audio, sr=fs, bins_per_octave=bins_per_octave)
    fmin_tuned = fmin * 2 ** (tuning_est / bins_per_octave)
    kw = dict(sr=fs, hop_length=hop, n_bins=n_bins, bins_per_octave=bins_per_octave, tuning=0.0)
    C1 = cqt(f_audio, fmin=fmin_tuned, **kw)
    f_hcqt = np.zeros((n_bins, C1.shape[1], num_harmonics + num_subharmonics))
    f_hcqt[:, :, num_subharmonics] = np.abs(C1)
    for n_ha in range(2, num_harmonics + 1):
        f_hcqt[:, :, num_subharmonics + n_ha - 1] = np.abs(cqt(f_audio, fmin=n_ha * fmin_tuned, **kw))
    for n_hs in range(1, num_subharmonics + 1):
        f_hcqt[:, :, num_subharmonics - n_hs] = np.abs(cqt(f_audio, fmin=fmin_tuned / (n_hs + 1), **kw))
    return f_hcqt, fs_hcqt, hop


# ----------------------------------------------------------------------------- synthetic audio
from tests.synth import synth_clip  # noqa: E402,F401  (the generator is workload data, shared with bench.py; it lives outside oracle/)
