"""`-m gpu`: rows added after the first pass, through the C ABI against the reference-generated goldens (tests/golden/ext_golden.npz)
and the oracle: compute_hcqt (early down-sampling + full-rate top octave), the training-time augmentations fused with the
patch cut (8f row 1) and the evaluation measures (8f row 4)."""
import numpy as np
import pytest
import torch

from oracle import hcqt_oracle as Q
from oracle import host_oracle as HO

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4       # of the clip's largest HCQT magnitude, as in test_gpu_hcqt.py


@pytest.mark.parametrize('factor', [2, 4, 8])
def test_pow2_decimator_matches_oracle(factor):
    from multipitch_architectures_b200 import _lib
    from multipitch_architectures_b200.libdl.data_preprocessing import _filterbank as FB
    for n in (4001, 10000, 16 * factor + 3):
        y = np.random.default_rng(n).standard_normal(n).astype(np.float32)
        ref = Q.resample_pow2(y, factor)
        taps = torch.from_numpy(FB.kaiser_fast_half_taps(factor)).cuda()
        out = torch.empty(-(-n // factor), dtype=torch.float32, device='cuda')
        _lib.call('decimate_f32', torch.from_numpy(y).cuda(), out, taps, taps.numel(), factor, _lib.i64(n), _lib.stream_ptr())
        assert out.numel() == len(ref)
        assert np.abs(out.cpu().numpy() - ref).max() < 1e-6


def test_compute_hcqt_matches_reference_wrapper_golden(ext_golden):
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_hcqt
    y = Q.synth_clip(3, seconds=2.0)
    f, fs_h, hop = compute_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    gold = ext_golden['hcqt_std_2s']
    assert f.dtype == np.float64 and f.shape == gold.shape and hop == 448 and fs_h == 22050 / 448
    err = np.abs(f - gold).max()
    print('compute_hcqt 2 s clip: max|diff| =', err, 'of max', gold.max())
    assert err < REL_TOL * gold.max()
    # the reference's default arguments: 60 bins per octave, hop 256, 4:1 one-shot early down-sampling for the sub-harmonic
    g, _, hop60 = compute_hcqt(y[:22050])
    gold60 = ext_golden['hcqt_std60_1s']
    assert hop60 == 256 and g.shape == gold60.shape
    assert np.abs(g - gold60).max() < REL_TOL * gold60.max()


@pytest.mark.parametrize('seed,seconds', [(21, 3.0), (22, 1.7)])
def test_compute_hcqt_matches_oracle_fresh_clips(seed, seconds):
    from multipitch_architectures_b200.libdl.data_preprocessing import compute_hcqt
    y = Q.synth_clip(seed, seconds=seconds)
    f, _, _ = compute_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    ref, _, _ = Q.compute_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    assert f.shape == ref.shape
    assert np.abs(f - ref).max() < REL_TOL * ref.max()


# ----------------------------------------------------------------------------- evaluation measures
def test_eval_measures_match_reference_golden(ext_golden):
    from multipitch_architectures_b200.libdl.metrics import calculate_eval_measures, calculate_single_measure
    targ, pred = ext_golden['ev_targ'], ext_golden['ev_pred']
    names = [str(n) for n in ext_golden['ev_names']]
    for thr, key in ((0.4, 'ev_values_04'), (0.7, 'ev_values_07')):
        d = calculate_eval_measures(targ.astype(np.float64), pred.astype(np.float64), names, threshold=thr)
        for n, want in zip(names, ext_golden[key]):
            assert abs(d[n] - want) < 1e-12 * max(1.0, abs(want)), (n, d[n], want)
    # CUDA-resident inputs (what the engine leaves in HBM) give the same numbers; single-measure entry point
    tc, pc = torch.from_numpy(targ).cuda(), torch.from_numpy(pred).cuda()
    assert calculate_single_measure(tc, pc, 'f_measure', 0.4) == calculate_eval_measures(targ, pred, ['f_measure'], 0.4)['f_measure']
    with pytest.raises(AssertionError):
        calculate_single_measure(tc, pc, 'no_such_measure')
    with pytest.raises(AssertionError):
        calculate_single_measure(tc, pc[:-1], 'precision')


def test_prf_matches_host_golden_and_large_random(host_golden):
    from multipitch_architectures_b200.libdl.metrics import calculate_eval_measures, compute_eval_measures
    got = compute_eval_measures(host_golden['prf_targ'], host_golden['prf_pred'] >= 0.4)       # the libfmp-shaped entry point
    assert np.allclose(np.array(got, dtype=np.float64), host_golden['prf'], atol=1e-12, rtol=0)
    d = calculate_eval_measures(host_golden['prf_targ'], host_golden['prf_pred'], threshold=0.4)
    assert abs(d['precision'] - host_golden['prf'][0]) < 1e-12 and abs(d['recall'] - host_golden['prf'][1]) < 1e-12
    assert abs(d['f_measure'] - host_golden['prf'][2]) < 1e-12
    # BASELINE config-1 size and beyond: 20,000 frames, counts must be exact
    rng = np.random.default_rng(3)
    targ = (rng.uniform(size=(20000, 72)) < 0.04).astype(np.float32)
    pred = rng.uniform(size=(20000, 72)).astype(np.float32) ** 3
    from multipitch_architectures_b200.libdl.metrics import eval_sums
    s = eval_sums(targ, pred, 0.4)
    P, R, F, TP, FP, FN = HO.eval_prf(targ, pred.astype(np.float64), 0.4)
    assert (s[0], s[1] - s[0], s[2] - s[0]) == (TP, FP, FN)
    for name in ('cosine_sim', 'binary_crossentropy', 'euclidean_distance', 'binary_accuracy', 'soft_accuracy', 'accum_energy'):
        from multipitch_architectures_b200.libdl.metrics import calculate_single_measure
        got = calculate_single_measure(targ, pred, name, 0.4)
        assert abs(got - HO.eval_measure(targ, pred, name, 0.4)) < 1e-11, name


def test_mpe_scores_match_oracle():
    from multipitch_architectures_b200.libdl.metrics import calculate_mpe_measures_mireval
    rng = np.random.default_rng(5)
    targ = (rng.uniform(size=(700, 72)) < 0.06).astype(np.float32)
    pred = np.clip(0.6 * np.roll(targ, 12, axis=1) + 0.7 * targ * rng.uniform(size=targ.shape) + rng.uniform(size=targ.shape) ** 5, 0, 1)
    pred = pred.astype(np.float32)
    targ[3] = 0
    got = calculate_mpe_measures_mireval(targ, pred, threshold=0.4, min_pitch=24)
    want = HO.mpe_scores(targ, pred, 0.4, 24)
    assert set(got) == set(want) and len(got) == 14
    for k in want:
        assert abs(got[k] - want[k]) < 1e-12, k


def test_auc_and_ap_match_sklearn():
    sk = pytest.importorskip('sklearn.metrics')
    from multipitch_architectures_b200.libdl.metrics import roc_auc, average_precision
    rng = np.random.default_rng(8)
    targ = (rng.uniform(size=(300, 72)) < 0.1).astype(np.float64)
    pred = np.round(np.clip(0.5 * targ + rng.uniform(size=targ.shape) * 0.7, 0, 1), 2).astype(np.float32)     # many ties
    assert abs(roc_auc(targ, pred) - sk.roc_auc_score(targ.flatten(), pred.flatten())) < 1e-12
    assert abs(average_precision(targ, pred) - sk.average_precision_score(targ.flatten(), pred.flatten())) < 1e-12


# ----------------------------------------------------------------------------- augmentations
def _dataset(ext_golden, params, pitch_class=False):
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    inp = torch.from_numpy(ext_golden['aug_in']).cuda()
    tg = ext_golden['aug_tg'][:, :12] if pitch_class else ext_golden['aug_tg']
    p = dict({'context': 75, 'stride': 1, 'compression': 10}, **params)
    return dataset_context(inp, torch.from_numpy(np.ascontiguousarray(tg)).cuda(), p)


def test_augmentation_deterministic_part_matches_reference_golden(ext_golden):
    """Every case of the reference run without additive noise: same decisions in, same patch out (bins filled with fresh noise
    excepted: those must be small non-negative values), same rolled targets."""
    rows = [0, 37, 74]
    for ci, params in ((0, {'aug:randomeq': 20, 'aug:tuning': True, 'aug:transpsemitones': 5}),
                       (2, {'aug:transpsemitones': 2, 'targettype': 'pitch_class'})):
        ds = _dataset(ext_golden, params, pitch_class=ci == 2)
        cases = [c for c in ext_golden['aug_cases'] if c[0] == ci]
        idx = [int(c[1]) for c in cases]
        dec = {'transp': np.array([c[5] for c in cases])}
        if ci == 0:
            dec.update(eq_alpha=np.array([c[2] for c in cases]), eq_beta=np.array([c[3] for c in cases]), tune2=np.array([c[4] for c in cases]))
        X, y = ds.gather(idx, decisions=dec)
        X, y = X.cpu().numpy(), y.cpu().numpy()
        assert X.shape == (len(idx), 6, 75, 216) and y.shape == (len(idx), 1, 1, 12 if ci == 2 else 72)
        for j, (_, i, alpha, beta, tune2, transp) in enumerate(cases):
            tag = 'aug%d_%d' % (ci, i)
            gold = ext_golden[tag + '_X']
            filled = np.zeros(216, bool)
            if transp > 0:
                filled[:3 * transp] = True
            elif transp < 0:
                filled[3 * transp:] = True
            if tune2 > 0:
                filled[(0 + 3 * transp) % 216] = True
            elif tune2 < 0:
                filled[(215 + 3 * transp) % 216] = True
            got = X[j][:, rows, :]
            assert np.abs(got - gold)[:, :, ~filled].max() < 1e-6, tag
            if filled.any():
                fv = got[:, :, filled]
                assert (fv >= 0).all() and fv.max() < 1e-3 and fv.mean() > 2e-5, tag      # |N(0, 1e-4)|: mean 8e-5
            assert np.array_equal(y[j], ext_golden[tag + '_y']), tag


def test_augmentation_noise_statistics_and_reproducibility(ext_golden):
    ds = _dataset(ext_golden, {'aug:noisestd': 1e-2})
    ds0 = _dataset(ext_golden, {})
    ds.compression = ds0.compression = None
    idx = list(range(0, 50))
    X, _ = ds.gather(idx, noise_offset=7)
    X2, _ = ds.gather(idx, noise_offset=7)
    X3, _ = ds.gather(idx, noise_offset=8)
    C, _ = ds0.gather(idx)
    assert torch.equal(X, X2) and not torch.equal(X, X3)
    big = C > 0.08                      # |x + n| = x + n where x >> std
    d = (X - C)[big].double()
    assert big.sum() > 1e5
    assert abs(d.mean().item()) < 3e-4 and abs(d.std().item() - 1e-2) < 2e-4
    k = ((d / 1e-2) ** 4).mean().item()
    assert abs(k - 3.0) < 0.15          # Gaussian kurtosis
    assert (X >= 0).all()


def test_augmentation_decision_distributions(ext_golden):
    ds = _dataset(ext_golden, {'aug:randomeq': 20, 'aug:tuning': True, 'aug:transpsemitones': 5})
    ds.generator = torch.Generator().manual_seed(0)
    d = ds.draw(3000)
    assert set(np.unique(d['tune2'])) == {-2, -1, 0, 1, 2} and set(np.unique(d['transp'])) == set(range(-5, 6))
    assert abs(np.mean(d['tune2'] == 0) - 0.2) < 0.03 and abs(np.mean(d['transp'] == 5) - 1 / 11) < 0.02
    assert d['eq_alpha'].min() >= 1 and d['eq_alpha'].max() <= 20 and d['eq_beta'].min() >= 0 and d['eq_beta'].max() < 216
    # the reference's rejection rule: the EQ curve never goes negative on any harmonic
    f = np.arange(216)
    for a, b in zip(d['eq_alpha'][:300], d['eq_beta'][:300]):
        for c in range(6):
            off = -36 if c == 0 else int(36 * np.log2(c))
            assert (1 - 2e-6 * a * (f - (b - off)) ** 2).min() >= -1e-6


def test_gather_without_augmentation_equals_batch_and_reference_items(ext_golden):
    ds = _dataset(ext_golden, {})
    Xb, yb = ds.batch(5, 20)
    Xg, yg = ds.gather(range(5, 25))
    assert torch.equal(Xb, Xg) and torch.equal(yb, yg)
    perm = [30, 2, 17, 54, 0]
    Xp, yp = ds.gather(perm)
    inp, tg = ext_golden['aug_in'], ext_golden['aug_tg']
    for j, i in enumerate(perm):
        Xr, yr = HO.context_item(inp, tg, i)
        assert np.abs(Xp[j].cpu().numpy() - Xr).max() < 1e-6 and np.array_equal(yp[j].cpu().numpy(), yr)
    Xi, yi = ds[17]
    assert torch.equal(Xi, Xp[2]) and torch.equal(yi, yp[2])
    with pytest.raises(IndexError):
        ds.gather([len(ds)])


# ----------------------------------------------------------------------------- on-disk formats
def test_hcqt_npy_loader_is_transpose_pad_cast(tmp_path):
    from multipitch_architectures_b200 import io
    rng = np.random.default_rng(1)
    for (F, N, C) in ((216, 131, 6), (216, 32, 6), (60, 7, 3)):
        a = np.abs(rng.normal(size=(F, N, C)))
        path = str(tmp_path / f'h_{F}_{N}_{C}.npy')
        np.save(path, a)
        want = np.pad(np.transpose(a, (2, 1, 0)), ((0, 0), (37, 38), (0, 0))).astype(np.float32)
        got = io.load_hcqt_npy(path, lead=37, trail=38).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want)
        assert np.array_equal(io.load_hcqt_npy(path).cpu().numpy(), np.transpose(a, (2, 1, 0)).astype(np.float32))
    roll = (rng.uniform(size=(128, 50)) < 0.05).astype(np.float64)
    np.save(str(tmp_path / 'p.npy'), roll)
    assert np.array_equal(io.load_pitch_npy(str(tmp_path / 'p.npy')).cpu().numpy(), roll.T[:, 24:96].astype(np.float32))


def test_load_audio_wav_downsampling_matches_oracle(tmp_path):
    import wave
    from multipitch_architectures_b200 import io
    rng = np.random.default_rng(2)
    n = 44100 + 17
    t = np.arange(n) / 44100
    st = np.stack([0.4 * np.sin(2 * np.pi * 440 * t) + 0.05 * rng.standard_normal(n), 0.3 * np.sin(2 * np.pi * 3000 * t)], 1)
    pcm = np.round(st * 32767).astype('<i2')
    path = str(tmp_path / 'a.wav')
    with wave.open(path, 'wb') as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(pcm.tobytes())
    y, sr = io.load_audio(path, sr=22050)
    mono = (pcm.astype(np.float32) / 32768.0).mean(axis=1, dtype=np.float32)
    want = Q.resample_pow2(mono, 2, filt='kaiser_best', scale=False)
    assert sr == 22050 and y.numel() == len(want) == (n + 1) // 2
    assert np.abs(y.cpu().numpy() - want).max() < 1e-6
    y0, sr0 = io.load_audio(path, sr=None)
    assert sr0 == 44100 and np.array_equal(y0.cpu().numpy(), mono)
    # a rate that is not a power-of-two multiple: resampy's general table walk (44.1 kHz -> 16 kHz, ratio 160/441)
    y16, sr16 = io.load_audio(path, sr=16000)
    want16 = Q.resample_general(mono, 44100, 16000, 'kaiser_best', False)
    assert sr16 == 16000 and y16.numel() == len(want16) and np.abs(y16.cpu().numpy() - want16).max() < 1e-6


@pytest.mark.parametrize('sr_from,sr_to,filt,scale', [(48000, 22050, 'kaiser_best', False), (16000, 22050, 'kaiser_best', False),
                                                      (22050, 11025, 'kaiser_fast', True), (32000, 22050, 'kaiser_fast', True)])
def test_general_resampler_matches_oracle(sr_from, sr_to, filt, scale):
    from multipitch_architectures_b200 import io
    n = 20000 + 13
    x = (0.5 * np.sin(2 * np.pi * 700 * np.arange(n) / sr_from) + 0.1 * np.random.default_rng(sr_from).standard_normal(n)).astype(np.float32)
    want = Q.resample_general(x, sr_from, sr_to, filt, scale)
    got = io.resample(torch.from_numpy(x).cuda(), sr_from, sr_to, filt, scale).cpu().numpy()
    assert got.shape == want.shape and np.abs(got - want).max() < 2e-6
    # a pure in-band tone survives with its amplitude (up to librosa's sqrt(ratio) energy rescaling when scale=True)
    tone = (0.5 * np.sin(2 * np.pi * 700 * np.arange(n) / sr_from)).astype(np.float32)
    r = io.resample(torch.from_numpy(tone).cuda(), sr_from, sr_to, filt, scale).cpu().numpy()
    mid = r[len(r) // 4: 3 * len(r) // 4]
    gain = 1.0 / np.sqrt(sr_to / sr_from) if scale else 1.0
    assert abs(np.abs(mid).max() / gain - 0.5) < 0.01


def test_evaluate_file_and_results_csv(tmp_path):
    """The reference's per-file test loop (exp126a...py:404-458) over .npy files: predictions saved as float64 [N, 72], every measure
    equal to the oracle's evaluation of the oracle's patch-wise predictions (to the fp32 path's tolerance), CSV layout."""
    import csv
    from multipitch_architectures_b200 import io
    from oracle import nn_oracle as NO
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict
    m = build_model('cnn_xs')
    sd = fill_state_dict(m.state_dict(), 31)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    rng = np.random.default_rng(4)
    rows = []
    for k, N in enumerate((60, 45)):
        h = np.abs(rng.normal(0, 0.05, size=(216, N, 6))) * rng.uniform(0.3, 2.0, size=(1, N, 1))
        roll = (rng.uniform(size=(128, N)) < 0.05).astype(np.float64)
        ph, pa = str(tmp_path / f'f{k}_hcqt.npy'), str(tmp_path / f'f{k}_pitch.npy')
        np.save(ph, h)
        np.save(pa, roll)
        row = io.evaluate_file(m, ph, pa, dir_predictions=str(tmp_path / 'pred'))
        rows.append(row)
        saved = np.load(str(tmp_path / 'pred' / f'f{k}_hcqt.npy'))
        assert saved.dtype == np.float64 and saved.shape == (N, 72)
        hc = np.transpose(h, (2, 1, 0))
        ip, _ = HO.pad_for_inference(hc, np.zeros((N, 72)))
        X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(N)]))
        with torch.no_grad():
            ref = NO.cnn_forward(sd, X, residual=False).reshape(N, 72).numpy()
        assert np.abs(saved - ref).max() < 1e-3
        targ = roll.T[:, 24:96]
        for name in ('cosine_sim', 'binary_crossentropy', 'euclidean_distance', 'soft_accuracy', 'accum_energy'):
            assert abs(row[name] - HO.eval_measure(targ, ref.astype(np.float64), name, 0.4)) < 2e-3, name
        assert abs(row['f_measure'] - HO.eval_measure(targ, saved, 'f_measure', 0.4)) < 1e-12
        assert abs(row['Accuracy'] - HO.mpe_scores(targ, saved, 0.4)['Accuracy']) < 1e-12
    table = io.write_results_csv(rows, str(tmp_path / 'res.csv'))
    with open(str(tmp_path / 'res.csv')) as f:
        lines = list(csv.reader(f))
    assert lines[0][:3] == ['', 'Filename', 'precision'] and len(lines[0]) == 2 + 11 + 14
    assert [l[1] for l in lines[1:]] == ['f0_hcqt.npy', 'f1_hcqt.npy', 'FILEWISE MEAN', 'FRAMEWISE MEAN']
    fm = (60 * rows[0]['cosine_sim'] + 45 * rows[1]['cosine_sim']) / 105
    assert abs(float(lines[4][2 + 3]) - fm) < 1e-12 and abs(table[2][1 + 3] - (rows[0]['cosine_sim'] + rows[1]['cosine_sim']) / 2) < 1e-12


# ----------------------------------------------------------------------------- BLUnet
@pytest.mark.parametrize('B,T,I,H', [(5, 4, 416, 208), (50, 4, 832, 416), (3, 9, 40, 12), (33, 1, 64, 32)])
def test_lstm_layer_matches_oracle(B, T, I, H):
    from multipitch_architectures_b200 import _lib
    from oracle import nn_oracle as NO
    rng = np.random.default_rng(B * 1000 + T)
    sd = {}
    for sfx in ('', '_reverse'):
        sd['l.blstm.weight_ih_l0' + sfx] = torch.from_numpy((rng.uniform(-1, 1, size=(4 * H, I)) / np.sqrt(H)).astype(np.float32))
        sd['l.blstm.weight_hh_l0' + sfx] = torch.from_numpy((rng.uniform(-1, 1, size=(4 * H, H)) / np.sqrt(H)).astype(np.float32))
        sd['l.blstm.bias_ih_l0' + sfx] = torch.from_numpy((rng.uniform(-1, 1, size=4 * H) / np.sqrt(H)).astype(np.float32))
        sd['l.blstm.bias_hh_l0' + sfx] = torch.from_numpy((rng.uniform(-1, 1, size=4 * H) / np.sqrt(H)).astype(np.float32))
    x = torch.from_numpy(rng.standard_normal((B, T, I)).astype(np.float32))
    # the oracle takes NCHW [B, C, T, F] with features (c, f): use C = I, F = 1
    ref = NO.blstm_layer(x.permute(0, 2, 1)[:, :, :, None].contiguous(), sd, 'l')          # [B, 2H, T, 1]
    ref = ref[:, :, :, 0].permute(0, 2, 1)
    st = lambda k: torch.stack([sd[f'l.blstm.{k}_l0'], sd[f'l.blstm.{k}_l0_reverse']]).contiguous().cuda()
    out = torch.empty(B, T, 2 * H, dtype=torch.float32, device='cuda')
    ws_bytes = _lib.lib().mpa_lstm_layer_workspace(B, T, H, 2)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device='cuda')
    _lib.call('lstm_layer_f32', x.cuda(), st('weight_ih'), st('weight_hh'), st('bias_ih'), st('bias_hh'), out, B, T, I, H, 2, ws,
              _lib.usize(ws_bytes), _lib.stream_ptr())
    assert np.abs(out.cpu().numpy() - ref.numpy()).max() < 2e-5


@pytest.mark.parametrize('name,B,seed,scheme,prec,tol', [
    ('blunet_tiny', 3, 41, 'adversarial', 'fp32', 1e-3), ('blunet_d', 2, 42, 'torch_default', 'fp32', 1e-3),
    ('blunet_d', 2, 42, 'torch_default', 'fp16', 1e-3), ('blunet_tiny', 3, 41, 'adversarial', 'fp16', 2e-2)])
def test_blunet_matches_reference_golden(ext_golden, name, B, seed, scheme, prec, tol):
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict, synth_patches
    m = build_model(name, precision=prec)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed, scheme=scheme))
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(synth_patches(B, seed).cuda())
    err = np.abs(y.cpu().numpy() - ext_golden[name + '__y']).max()
    print(f'{name} {prec}: max|diff| vs reference = {err:.2e}')
    assert y.shape == (B, 1, 1, 72) and err < tol


def test_blunet_training_loss_and_gradients_match_reference_golden(ext_golden):
    """u_net_blstm_varlayers in train mode: loss.backward() through the tape incl. the BLSTM backward through time, against
    loss.backward() on the unmodified reference module (gradients of the large LSTM matrices are pinned on every 5th element)."""
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict, synth_patches, synth_targets
    name = 'blunet_s32'
    B, seed = [int(v) for v in ext_golden[name + '__train__meta']]
    m = build_model(name)
    m.load_state_dict(fill_state_dict(m.state_dict(), seed))
    for mod in m.modules():
        if hasattr(mod, 'p_dropout'):
            mod.p_dropout = 0.0
    m = m.cuda().train()
    x, t = synth_patches(B, seed).cuda(), synth_targets(B, seed).cuda()
    y = m(x)
    loss = torch.nn.BCELoss(reduction='mean')(y, t)
    assert np.abs(y.detach().cpu().numpy() - ext_golden[name + '__train__y']).max() < 1e-3
    loss.backward()
    assert abs(loss.item() - float(ext_golden[name + '__train__loss'][0])) < 2e-5
    gmax = max(np.abs(ext_golden[name + '__train__grad__' + k]).max() for k, _ in m.named_parameters())
    worst = 0.0
    for k, p in m.named_parameters():
        g = ext_golden[name + '__train__grad__' + k]
        mine = p.grad.cpu().numpy().reshape(-1)
        mine = mine[::5] if mine.size > 20000 else mine
        d = np.abs(mine - g)
        scale = max(np.abs(g).max(), 1e-3 * gmax)
        worst = max(worst, d.max() / scale)
        assert d.max() <= 0.25 * scale and (d.mean() <= 1e-2 * scale or d.size < 64), (k, d.max() / scale, d.mean() / scale)
    print(f'{name} train: worst relative gradient deviation {worst:.2e}')


@pytest.mark.parametrize('B,T,I,H', [(5, 4, 208, 104), (25, 4, 832, 416), (3, 6, 40, 12)])
def test_lstm_layer_backward_matches_torch_autograd(B, T, I, H):
    from multipitch_architectures_b200 import _lib
    torch.manual_seed(B + T)
    lstm = torch.nn.LSTM(input_size=I, hidden_size=H, num_layers=1, batch_first=True, bidirectional=True)
    x = torch.randn(B, T, I, requires_grad=True)
    gy = torch.randn(B, T, 2 * H)
    out_ref, _ = lstm(x)
    out_ref.backward(gy)
    st = lambda k: torch.stack([getattr(lstm, k + '_l0').detach(), getattr(lstm, k + '_l0_reverse').detach()]).contiguous().cuda()
    w_ih, w_hh, b_ih, b_hh = st('weight_ih'), st('weight_hh'), st('bias_ih'), st('bias_hh')
    f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device='cuda')
    xc = x.detach().cuda()
    out, gates, c_all, ws = f32(B, T, 2 * H), f32(2, B * T, 4 * H), f32(2, B, T, H), f32(2 * B * H)
    _lib.call('lstm_layer_train_f32', xc, w_ih, w_hh, b_ih, b_hh, out, gates, c_all, B, T, I, H, 2, ws, _lib.usize(ws.numel() * 4), _lib.stream_ptr())
    assert (out.cpu() - out_ref.detach()).abs().max() < 2e-5
    wsb_bytes = _lib.lib().mpa_lstm_layer_bwd_workspace(B, T, H, 2)
    wsb = torch.empty(wsb_bytes, dtype=torch.uint8, device='cuda')
    g_x, g_wih, g_whh, g_b = f32(B, T, I), f32(2, 4 * H, I), f32(2, 4 * H, H), f32(2, 4 * H)
    _lib.call('lstm_layer_bwd_f32', xc, w_ih, w_hh, gates, c_all, out, gy.cuda(), g_x, g_wih, g_whh, g_b, B, T, I, H, 2, wsb, _lib.usize(wsb_bytes),
              _lib.stream_ptr())
    assert (g_x.cpu() - x.grad).abs().max() < 5e-5
    for d, sfx in enumerate(('', '_reverse')):
        for mine, k in ((g_wih, 'weight_ih'), (g_whh, 'weight_hh'), (g_b, 'bias_ih'), (g_b, 'bias_hh')):
            ref = getattr(lstm, k + '_l0' + sfx).grad
            assert (mine[d].cpu() - ref).abs().max() < 2e-4 * max(1.0, ref.abs().max().item()), (k, sfx)


def test_host_prefetcher_delivers_batches_in_order():
    from multipitch_architectures_b200.io import HostPrefetcher
    rng = np.random.default_rng(0)
    host = [(torch.from_numpy(rng.standard_normal((4, 6, 75, 216)).astype(np.float32)).pin_memory(),
             torch.from_numpy(rng.uniform(size=(4, 1, 1, 72)).astype(np.float32)).pin_memory()) for _ in range(5)]
    seen = 0
    for k, (x, y) in enumerate(HostPrefetcher(host)):
        assert x.is_cuda and torch.equal(x.cpu(), host[k][0]) and torch.equal(y.cpu(), host[k][1])
        (x * 2).sum()        # consumer work on the compute stream
        seen += 1
    assert seen == 5


# ----------------------------------------------------------------------------- the scripts' training loop on top of the pieces
def _toy_sets(n_files, n_frames, seed, params):
    """HCQT-like inputs whose fundamental channel carries the labels: learnable in a few dozen steps."""
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    rng = np.random.default_rng(seed)
    sets = []
    for _ in range(n_files):
        roll = (rng.uniform(size=(n_frames, 72)) < 0.06).astype(np.float32)
        roll = np.maximum.reduce([np.roll(roll, s, axis=0) for s in range(6)])          # notes last 6 frames
        x = 0.02 * np.abs(rng.standard_normal((6, n_frames, 216))).astype(np.float32)
        x[1, :, 1::3] += 0.8 * roll
        x[2, :, 37::3][:, :60] += 0.4 * roll[:, :60]
        sets.append(dataset_context(torch.from_numpy(x).cuda(), torch.from_numpy(roll).cuda(), dict({'context': 75, 'stride': 1, 'compression': 10}, **params)))
    return sets


@pytest.mark.parametrize('name,graph', [('cnn_xs', True), ('unet_tiny', False)])
def test_fit_loop_trains_validates_schedules_and_checkpoints(tmp_path, name, graph):
    from multipitch_architectures_b200.loop import fit
    from tests.refshapes import build_model
    torch.manual_seed(0)
    m = build_model(name, precision='bf16').cuda()
    aug = {'aug:randomeq': 20, 'aug:noisestd': 1e-4, 'aug:tuning': True, 'aug:transpsemitones': 5}
    train, val = _toy_sets(3, 160, 1, aug), _toy_sets(1, 120, 2, {})
    path = str(tmp_path / 'best.pt')
    logs = []
    hist = fit(m, train, val, batch_size=16, val_batch_size=25, lr=2e-3, max_epochs=4, max_batches_per_epoch=12, seed=3, save_path=path, graph=graph,
               scheduler=dict(factor=0.5, patience=0, threshold=0.5), early=dict(patience=3, min_delta=1e-5), log=logs.append)
    assert 1 <= len(hist) <= 4 and len(logs) == len(hist) and all(np.isfinite(h['train_loss']) and np.isfinite(h['val_loss']) for h in hist)
    assert hist[-1]['train_loss'] < hist[0]['train_loss']                     # it learns
    assert hist[-1]['lr'] < 2e-3                                              # a 50 % improvement threshold forces the scheduler to act
    sd = torch.load(path)
    assert set(sd) == set(m.state_dict()) and all(torch.isfinite(v.float()).all() for v in sd.values())
    m2 = build_model(name, precision='bf16')
    m2.load_state_dict(sd)


def test_example_predict_wav_runs_the_notebook_flow(tmp_path):
    """examples/predict_wav.py (notebook 02 of the reference): WAV -> 22.05 kHz -> HCQT -> DRCNN -> [n_frames, 72]."""
    import importlib.util
    import os
    import wave
    spec = importlib.util.spec_from_file_location('predict_wav', os.path.join(os.path.dirname(os.path.dirname(__file__)), 'examples', 'predict_wav.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    y = Q.synth_clip(4, seconds=1.5, sr=44100)
    path = str(tmp_path / 'clip.wav')
    with wave.open(path, 'wb') as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(np.round(y * 32767).astype('<i2').tobytes())
    torch.manual_seed(0)
    act, roll, fs_hcqt = mod.predict(path)
    n22 = (len(y) + 1) // 2
    assert act.shape == (n22 // 512 + 1, 72) and roll.shape == act.shape and roll.dtype == torch.bool
    assert abs(fs_hcqt - 22050 / 512) < 1e-12 and torch.isfinite(act).all() and (act >= 0).all() and (act <= 1).all()
