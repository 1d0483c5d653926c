"""`-m "not gpu"`: host-side mirror of the reference interface (dataset index math, metrics, hop size, annotation
rasteriser, harmonic plan, filter tables) against the reference-generated goldens and the oracle."""
import numpy as np
import pytest
import torch

from oracle import hcqt_oracle as Q
from oracle import host_oracle as HO


def test_dataset_context_index_math_matches_reference_golden(host_golden):
    """Host logic only (integer index math, H4): lengths and the frames each item covers; the item VALUES are produced by the CUDA kernels
    and are checked against the same golden in tests/test_gpu_engine.py.  Without a GPU an item request fails loudly."""
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    inp, tg = host_golden['ds_in'].astype(np.float64), host_golden['ds_tg'].astype(np.float64)
    ip, tp = HO.pad_for_inference(inp, tg)
    ds = dataset_context(torch.from_numpy(ip), torch.from_numpy(tp), {'context': 75, 'stride': 1, 'compression': 10})
    assert len(ds) == int(host_golden['ds_len'][0])
    for j, i in enumerate(host_golden['ds_idx']):
        a, b, c = ds.patch_frames(int(i))
        assert (b - a, c) == (75, int(i) + 37)
        assert np.array_equal(tp[c][None, None, :].astype(np.float32), host_golden['ds_y'][j])
        assert np.abs(np.log(1 + np.float32(10) * ip[:, a:b].astype(np.float32)) - host_golden['ds_X'][j]).max() < 1e-6
    ds3 = dataset_context(torch.from_numpy(ip), torch.from_numpy(tp), {'context': 75, 'stride': 3, 'compression': None})
    assert len(ds3) == int(host_golden['ds3_len'][0])
    assert ds3.patch_frames(5) == (15, 90, 52)
    assert np.array_equal(tp[52][None, None, :].astype(np.float32), host_golden['ds3_y5'])
    if not torch.cuda.is_available():
        with pytest.raises(_lib.MpaError):
            ds[0]


def test_hopsize_annotation_plan(host_golden):
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import (compute_hopsize_cqt, compute_annotation_array_nooverlap,
                                                                            _harmonic_plan)
    for target, noct, hop, fs in host_golden['hopsize']:
        assert compute_hopsize_cqt(target, 22050, int(noct)) == (int(hop), fs)
    import os
    ev = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'annot_2382_events.npy'))
    fs = 22050 / 512
    n_frames = int(np.floor(ev[:, 1].max() * fs)) + 5
    A = compute_annotation_array_nooverlap(ev.copy(), np.zeros((216, n_frames, 6)), fs, annot_type='pitch')
    assert np.array_equal(np.argwhere(A > 0).astype(np.int32), host_golden['annot_nnz'])
    assert _harmonic_plan(5, 1) == Q.harmonic_plan(5, 1)
    assert _harmonic_plan(3, 2)[1] == Q.harmonic_plan(3, 2)[1]


def test_filter_tables_match_oracle_filterbank():
    from multipitch_architectures_b200.libdl.data_preprocessing import _filterbank as FB
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import _harmonic_plan
    fmin = Q.C1_HZ / 2 ** (2 / 72)
    lh, base = _harmonic_plan(5, 1)
    tabs = FB.build_tables(22050, 512, fmin, 36, 6, lh, base)
    assert sorted(tabs) == [((0, 0), 256)] + [((0, i), 512) for i in range(9)]
    for ti in (0, 37, 50, 99):
        tun = FB.tuning_values()[ti]
        fb, n_fft, _ = Q.cqt_filter_fft(22050 / 4, fmin * 2 ** (tun / 36) * 0.5 * 2 ** (288 / 36) / 4, 36, 36)
        t = tabs[((0, 2), 512)]
        for r in range(36):
            s = t['start'][ti, r]
            dense = np.zeros(257, np.complex64)
            dense[s:s + 32] = t['basis'][ti, r][:257 - s]
            assert np.array_equal(dense, fb[r])
    assert np.array_equal(FB.kaiser_fast_half_taps(), Q._kaiser_fast_halfband().astype(np.float32))
    # every output cell (channel, bin) is written by exactly one row
    cover = np.zeros((6, 216), int)
    for t in tabs.values():
        for code in t['dest'].ravel():
            if code >= 0:
                cover[code >> 16, code & 0xffff] += 1
    assert (cover == 1).all()


def _resampy_table_walk(x, factor):
    """The table-walking form of resampy's resample_f for sample_ratio = 1/factor (kaiser_fast: 16 zero crossings, 512 table
    steps per crossing), written out literally as a second, independent statement of the decimator."""
    import scipy.signal
    num_zeros, precision, rolloff, beta = 16, 9, 0.85, 8.555504641634386
    num_table = 2 ** precision
    n_tab = num_table * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n_tab + 1, endpoint=True))
    interp_win = sinc_win * scipy.signal.windows.kaiser(2 * n_tab + 1, beta)[n_tab:]
    ratio = 1.0 / factor
    interp_win = interp_win * ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, ratio)
    time_increment = 1.0 / ratio
    index_step = int(scale * num_table)
    n_orig, n_out = len(x), int(len(x) * ratio)
    y = np.zeros(n_out, dtype=x.dtype)
    nwin = len(interp_win)
    time_register = 0.0
    for t in range(n_out):
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        for i in range(i_max):
            w = interp_win[offset + i * index_step] + eta * interp_delta[offset + i * index_step]
            y[t] += w * x[n - i]
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        for k in range(k_max):
            w = interp_win[offset + k * index_step] + eta * interp_delta[offset + k * index_step]
            y[t] += w * x[n + k + 1]
        time_register += time_increment
    return y


@pytest.mark.parametrize('factor', [2, 4])
def test_oracle_decimator_matches_table_walk(factor):
    x = np.random.default_rng(factor).standard_normal(700 + factor + 1)
    walk = _resampy_table_walk(x, factor) * np.sqrt(factor)
    got = Q.resample_pow2(x, factor)
    assert len(got) == -(-len(x) // factor) and len(walk) == len(x) // factor
    assert np.abs(got[:len(walk)] - walk).max() < 1e-12
    assert (got[len(walk):] == 0).all()


def test_compute_hcqt_schedule_uses_early_downsampling_and_top_octave():
    """compute_hcqt (hcqt.py:34-85) at the paper's 36 bins/octave: hop 448; the low harmonics trigger librosa's one-shot early
    down-sampling, h = 5 its full-rate top octave; the product tables follow the oracle's schedule."""
    from multipitch_architectures_b200.libdl.data_preprocessing import _filterbank as FB
    fmin = Q.C1_HZ / 2 ** (2 / 72)
    lh = [0.5, 1.0, 2.0, 3.0, 4.0, 5.0]
    early = [FB.cqt_schedule(22050, 448, fmin * h, 216, 36)[0] for h in lh]
    assert early == [1, 1, 0, 0, 0, 0]
    assert FB.cqt_schedule(22050, 448, fmin * 5, 216, 36)[1][0][1] == 'top'
    assert [FB.cqt_schedule(22050, 256, Q.C1_HZ / 2 ** (4 / 120) * h, 360, 60)[0] for h in (0.5, 1.0)] == [2, 1]
    tabs = FB.build_tables(22050, 448, fmin, 36, 6, lh, lh)
    cover = np.zeros((6, 216), int)
    for t in tabs.values():
        for code in t['dest'].ravel():
            if code >= 0:
                cover[code >> 16, code & 0xffff] += 1
    assert (cover == 1).all()
    assert np.array_equal(FB.kaiser_fast_half_taps(4), Q._kaiser_fast_half(4).astype(np.float32))
    # frame counts follow the shortest octave response of the oracle's cqt
    for n in (20000, 20000 + 447):
        y = np.zeros(n, np.float32)
        y[::5] = 1.0
        C = Q.cqt(y, sr=22050, hop_length=448, fmin=fmin, n_bins=216, bins_per_octave=36)
        assert C.shape[1] == FB.cqt_frames(22050, 448, fmin, 216, 36, n)


def test_oracle_general_resampler_matches_literal_table_walk():
    """Q.resample_general (vectorised over outputs) vs the literal per-output loops of resampy's resample_f for a non-trivial ratio."""
    import scipy.signal
    ratio_from, ratio_to = 48000, 22050
    ratio = ratio_to / ratio_from
    x = np.random.default_rng(1).standard_normal(900)
    win, num_table = Q.resampy_window('kaiser_best')
    win = win * ratio
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    scale = min(1.0, ratio)
    index_step = int(scale * num_table)
    n_out = int(len(x) * ratio)
    y = np.zeros(n_out)
    time_register = 0.0
    for t in range(n_out):
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        for i in range(min(n + 1, (len(win) - offset) // index_step)):
            y[t] += (win[offset + i * index_step] + eta * delta[offset + i * index_step]) * x[n - i]
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        for k in range(min(len(x) - n - 1, (len(win) - offset) // index_step)):
            y[t] += (win[offset + k * index_step] + eta * delta[offset + k * index_step]) * x[n + k + 1]
        time_register += 1.0 / ratio
    got = Q.resample_general(x, ratio_from, ratio_to, 'kaiser_best', False)
    assert len(got) == int(np.ceil(len(x) * ratio)) and np.abs(got[:n_out] - y).max() < 1e-9
    # identical to the dedicated power-of-two oracle where both apply
    xf = x.astype(np.float32)
    assert np.array_equal(Q.resample_general(xf, 44100, 22050, 'kaiser_fast', True), Q.resample_pow2(xf, 2, 'kaiser_fast', True))


def test_early_stopping_matches_reference_golden():
    import os
    from multipitch_architectures_b200.libdl.metrics import early_stopping
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ext_golden.npz'))
    cfg = [('min', 1e-5, 12, False), ('max', 0.01, 3, False), ('min', 2.0, 4, True), ('max', 1.0, 2, True), ('min', 0, 0, False)]
    for ci, (mode, md, pat, pct) in enumerate(cfg):
        for si in range(2):
            es = early_stopping(mode=mode, min_delta=md, patience=pat, percentage=pct)
            got = [int(bool(es.step(float(v)))) for v in g['es_seqs'][si]]
            assert got == list(g['es_flags'][ci, si]), (ci, si)
    with pytest.raises(ValueError):
        early_stopping(mode='median')


@pytest.mark.parametrize('width', [1, 2, 3, 4])
def test_read_wav_decodes_every_pcm_width(tmp_path, width):
    """Host-side WAV decode of io.load_audio (integer PCM -> float32 in [-1, 1)); resampling itself is a kernel (GPU tests)."""
    import wave
    from multipitch_architectures_b200.io import read_wav
    rng = np.random.default_rng(width)
    n, nch = 257, 2
    full = 2 ** (8 * width - 1)
    ints = rng.integers(-full, full, size=(n, nch))
    ints[0] = (-full, full - 1)
    if width == 1:
        raw = (ints + 128).astype(np.uint8).tobytes()
    elif width == 3:
        u = ints.astype(np.int64) & 0xFFFFFF
        raw = np.stack([u & 255, (u >> 8) & 255, (u >> 16) & 255], -1).astype(np.uint8).tobytes()
    else:
        raw = ints.astype('<i2' if width == 2 else '<i4').tobytes()
    path = str(tmp_path / f'w{width}.wav')
    with wave.open(path, 'wb') as w:
        w.setnchannels(nch)
        w.setsampwidth(width)
        w.setframerate(44100)
        w.writeframes(raw)
    x, sr = read_wav(path)
    assert sr == 44100 and x.shape == (n, nch) and x.dtype == np.float32
    assert np.array_equal(x, (ints / float(full)).astype(np.float32))


def test_results_csv_layout(tmp_path):
    import csv
    from multipitch_architectures_b200.io import write_results_csv
    rows = [{'Filename': 'a.npy', 'precision': 0.5, 'recall': 0.25, '_kframes': 1.0}, {'Filename': 'b.npy', 'precision': 1.0, 'recall': 0.75, '_kframes': 3.0}]
    table = write_results_csv(rows, str(tmp_path / 'r.csv'))
    lines = list(csv.reader(open(str(tmp_path / 'r.csv'))))
    assert lines[0] == ['', 'Filename', 'precision', 'recall']
    assert lines[3][1] == 'FILEWISE MEAN' and float(lines[3][2]) == 0.75 and float(lines[3][3]) == 0.5
    assert lines[4][1] == 'FRAMEWISE MEAN' and float(lines[4][2]) == 0.875 and float(lines[4][3]) == 0.625
    assert table[3][0] == 'FRAMEWISE MEAN'


def test_reduce_lr_on_plateau_matches_torch():
    """loop.ReduceLROnPlateau (drives the `lr` attribute of the fused train steps) vs torch.optim.lr_scheduler.ReduceLROnPlateau with the
    scripts' settings (mode min, factor 0.5, patience 5, threshold 1e-4 rel, min_lr 1e-6, exp126a...py:118-130)."""
    import types
    from multipitch_architectures_b200.loop import ReduceLROnPlateau
    rng = np.random.default_rng(0)
    for trial in range(4):
        seq = np.concatenate([np.linspace(1.0, 0.6, 15), 0.6 + 0.01 * rng.standard_normal(70)]) * (1 + 0.002 * rng.standard_normal(85))
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=1e-3)
        ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode='min', factor=0.5, patience=5 - trial, threshold=1e-4, threshold_mode='rel',
                                                         cooldown=trial % 2, eps=1e-8, min_lr=1e-6)
        holder = types.SimpleNamespace(lr=1e-3)
        mine = ReduceLROnPlateau(holder, factor=0.5, patience=5 - trial, threshold=1e-4, cooldown=trial % 2, min_lr=1e-6, eps=1e-8)
        for v in seq:
            ref.step(float(v))
            mine.step(float(v))
            assert holder.lr == opt.param_groups[0]['lr']
        assert holder.lr < 1e-3


def test_patch_sampler_covers_every_patch_once_per_epoch():
    """ConcatDataset + DataLoader(shuffle=True) semantics on stand-in datasets: every (file, patch) pair exactly once per epoch, batches
    straddle files, items land in the batch positions the sampler assigned."""
    from multipitch_architectures_b200.loop import PatchSampler

    class Fake:
        def __init__(self, n, tag):
            self.inputs, self.targets, self.context, self.n, self.tag = torch.zeros(2, n + 5, 4), torch.zeros(n + 5, 3), 5, n, tag

        def __len__(self):
            return self.n

        def gather(self, idx, out=None):
            for j, i in enumerate(idx):
                out[0][j] = 1000 * self.tag + int(i)
                out[1][j] = self.tag
    sets = [Fake(7, 1), Fake(12, 2), Fake(3, 3)]
    smp = PatchSampler(sets, batch_size=5, shuffle=True, seed=1)
    assert len(smp) == 5
    for _ in range(2):
        seen = []
        for X, y in smp.epoch():
            assert X.shape[1:] == (2, 5, 4) and y.shape[1:] == (1, 1, 3) and X.shape[0] in (5, 2)
            ids = X[:, 0, 0, 0].long().tolist()
            assert [int(v) // 1000 for v in ids] == y[:, 0, 0, 0].long().tolist()
            seen += ids
        assert sorted(seen) == sorted([1000 + i for i in range(7)] + [2000 + i for i in range(12)] + [3000 + i for i in range(3)])
    assert len(PatchSampler(sets, 5, max_batches=2)) == 2


def test_patch_sampler_ranks_partition_the_epoch():
    """Data-parallel sampling: with the same seed the ranks' items are disjoint, cover every patch, and every rank runs the same number
    of batches (the gradient all-reduce needs matching step counts)."""
    from multipitch_architectures_b200.loop import PatchSampler

    class Fake:
        def __init__(self, n, tag):
            self.inputs, self.targets, self.context, self.n, self.tag = torch.zeros(1, n + 3, 2), torch.zeros(n + 3, 1), 3, n, tag

        def __len__(self):
            return self.n

        def gather(self, idx, out=None):
            for j, i in enumerate(idx):
                out[0][j] = 100 * self.tag + int(i)
    for world in (2, 3, 8):
        per_rank = []
        for r in range(world):
            smp = PatchSampler([Fake(11, 1), Fake(20, 2)], batch_size=4, shuffle=True, seed=7, rank=r, world=world)
            items = [int(v) for X, _ in smp.epoch() for v in X[:, 0, 0, 0]]
            per_rank.append(items)
            assert len(smp) == -(-(-(-31 // world)) // 4)
        assert len({len(p) for p in per_rank}) == 1
        flat = [v for p in per_rank for v in p]
        assert set(flat) == set([100 + i for i in range(11)] + [200 + i for i in range(20)])
        assert len(flat) - len(set(flat)) == (-(-31 // world)) * world - 31          # only the wrap-around padding repeats


def test_cp8_resident_training_path_eligibility():
    """Host decision only: which CNN models keep activations / gradients in the CP8 planes between the block convolutions of a bf16
    training step (training._cp8_resident) — tensor-core mode, no residual path, whole bin quads per row."""
    from multipitch_architectures_b200 import training as TR
    from multipitch_architectures_b200.libdl.nn_models import _exec
    from tests.refshapes import build_model
    for name, prec, want in [('cnn_xs', 'bf16', True), ('dcnn_tiny', 'bf16', True), ('drcnn_tiny', 'bf16', False), ('cnn_xs', 'fp32', False)]:
        m = build_model(name, precision=prec)
        assert TR._cp8_resident(m, _exec.cnn_blocks(m), 216) is want, (name, prec)
    m = build_model('cnn_xs', precision='bf16')
    assert TR._cp8_resident(m, _exec.cnn_blocks(m), 218) is False          # rows must be whole bin quads (Philox draws serve 4 bins)
    TR.CP8_RESIDENT = False
    try:
        assert TR._cp8_resident(m, _exec.cnn_blocks(m), 216) is False
    finally:
        TR.CP8_RESIDENT = True


def test_conv3_row_packing_layout_and_pool13_doubling_table():
    """Host side of the fused CNN head (head.cu head_pool_conv3_tail_kernel): the conv3 operand layout [chunk][frame][channel][C2P] written by
    ops.pack_conv3_rows, and the doubling-table recurrence the pool13 kernels unroll (pairs -> fours -> eights -> two overlapping eights with
    delay lines of length 3 / 5 / 6) against max_pool1d((13), 1, 6) — the index arithmetic of maxpool13_bwd_table_kernel / pool13_table_cp8_kernel."""
    from multipitch_architectures_b200 import ops
    g = torch.Generator().manual_seed(0)
    for C2, C1, C1p in [(10, 20, 24), (16, 40, 40), (3, 5, 8)]:
        w = torch.randn(C2, C1, 75, 1, generator=g)
        wp = ops.pack_conv3_rows(w, C1p)
        C2P = 12 if C2 <= 12 else 16
        assert tuple(wp.shape) == (C1p // 8, 75, 8, C2P) and wp.is_contiguous()
        for ck, t, c8, co in [(0, 0, 0, 0), (C1p // 8 - 1, 74, 7, C2P - 1), (1 % (C1p // 8), 37, 3, C2 - 1), (0, 5, 4, 1)]:
            ci = ck * 8 + c8
            want = float(w[co, ci, t, 0]) if (co < C2 and ci < C1) else 0.0
            assert float(wp[ck, t, c8, co]) == want
    # the recurrence with the kernels' ring slots
    T, H = 75, 6
    x = torch.randn(T, generator=g).round()        # plateaus
    ninf = float('-inf')
    a2, a4, a8 = [ninf] * 3, [ninf] * 5, [ninf] * 6
    i2, i4, i8 = [0] * 3, [0] * 5, [0] * 6
    xprev, out, arg = ninf, [], []

    def first_max(v, i, bv, bi):
        return (bv, bi) if bv > v else (v, i)
    for p in range(T + H):
        xp = float(x[p]) if p < T else ninf
        a2[(p + 2) % 3], i2[(p + 2) % 3] = first_max(xprev, p - 1, xp, p)
        a4[(p + 2) % 5], i4[(p + 2) % 5] = first_max(a2[p % 3], i2[p % 3], a2[(p + 2) % 3], i2[(p + 2) % 3])
        a8[(p + 5) % 6], i8[(p + 5) % 6] = first_max(a4[(p + 3) % 5], i4[(p + 3) % 5], a4[(p + 2) % 5], i4[(p + 2) % 5])
        xprev = xp
        if p >= H:
            v, i = first_max(a8[p % 6], i8[p % 6], a8[(p + 5) % 6], i8[(p + 5) % 6])
            out.append(v)
            arg.append(i)
    want, idx = torch.nn.functional.max_pool1d(x.view(1, 1, T), 13, 1, 6, return_indices=True)
    assert torch.equal(torch.tensor(out), want.view(-1))
    # first maximum of every window (what the backward routes the gradient to)
    first = [min(j for j in range(max(0, t - 6), min(T, t + 7)) if x[j] == want.view(-1)[t]) for t in range(T)]
    assert arg == first
