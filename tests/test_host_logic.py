"""`-m "not gpu"`: host-side mirror of the reference interface (dataset index math, metrics, hop size, annotation
rasteriser, harmonic plan, filter tables) against the reference-generated goldens and the oracle."""
import numpy as np
import torch

from oracle import hcqt_oracle as Q
from oracle import host_oracle as HO


def test_dataset_context_matches_reference_golden(host_golden):
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    inp, tg = host_golden['ds_in'].astype(np.float64), host_golden['ds_tg'].astype(np.float64)
    ip, tp = HO.pad_for_inference(inp, tg)
    ds = dataset_context(torch.from_numpy(ip), torch.from_numpy(tp), {'context': 75, 'stride': 1, 'compression': 10})
    assert len(ds) == int(host_golden['ds_len'][0])
    for j, i in enumerate(host_golden['ds_idx']):
        X, y = ds[int(i)]
        assert X.dtype == torch.float32 and tuple(X.shape) == (6, 75, 216) and tuple(y.shape) == (1, 1, 72)
        assert np.abs(X.numpy() - host_golden['ds_X'][j]).max() < 1e-6
        assert np.array_equal(y.numpy(), host_golden['ds_y'][j])
    ds3 = dataset_context(torch.from_numpy(ip), torch.from_numpy(tp), {'context': 75, 'stride': 3, 'compression': None})
    assert len(ds3) == int(host_golden['ds3_len'][0])
    assert np.array_equal(ds3[5][1].numpy(), host_golden['ds3_y5'])


def test_metrics_match_reference_golden(host_golden):
    from multipitch_architectures_b200.libdl.metrics import calculate_eval_measures, compute_eval_measures
    got = compute_eval_measures(host_golden['prf_targ'], host_golden['prf_pred'] >= 0.4)
    assert np.allclose(np.array(got, dtype=np.float64), host_golden['prf'], atol=1e-12, rtol=0)
    d = calculate_eval_measures(host_golden['prf_targ'], host_golden['prf_pred'], threshold=0.4)
    assert abs(d['f_measure'] - host_golden['prf'][2]) < 1e-12


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
    assert sorted(tabs) == [(0, 256)] + [(i, 512) for i in range(9)]
    for ti in (0, 37, 50, 99):
        tun = FB.tuning_values()[ti]
        fb, n_fft, _ = Q.cqt_filter_fft(22050 / 4, fmin * 2 ** (tun / 36) * 0.5 * 2 ** (288 / 36) / 4, 36, 36)
        t = tabs[(2, 512)]
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
