"""`-m "not gpu"`: pins the oracle restatements against fixtures produced by the reference itself
(tests/golden/make_golden.py).  No CUDA, no /root/reference at run time."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import host_oracle as HO
from oracle import hcqt_oracle as Q
from oracle import nn_oracle as NO
from tests.weights import MODEL_SPECS, fill_state_dict, synth_patches, synth_targets
from tests.refshapes import reference_state_shapes

NN_CASES = [('cnn_xs', 'default'), ('drcnn', 'default'), ('unet_m', 'default'), ('punet', 'default'), ('saunet_l', 'default'),
            ('cnn_xs', 'eval'), ('drcnn_tiny', 'eval'), ('dcnn_tiny', 'eval'), ('drcnn', 'eval'),
            ('unet_tiny', 'eval'), ('unet_tiny', 'train'), ('unet_m', 'eval'), ('punet_tiny', 'eval'),
            ('punet', 'eval'), ('saunet_tiny', 'eval'), ('saunet_l', 'eval'), ('saunet_tiny', 'train')]


def oracle_forward(name, sd, x, train=False):
    spec = MODEL_SPECS[name]
    kw = spec['kw']
    if spec['cls'] in ('basic_cnn_segm_sigmoid', 'deep_cnn_segm_sigmoid'):
        return NO.cnn_forward(sd, x, residual=kw.get('residual', False))
    return NO.unet_forward(sd, x, train=train, num_heads=kw.get('num_heads', 8), pos_encoding=kw.get('pos_encoding'))


@pytest.mark.parametrize('name,mode', NN_CASES)
def test_nn_oracle_matches_reference_golden(nn_golden, name, mode):
    tag = f'{name}__{mode}'
    B, seed, wsum = nn_golden[tag + '__meta']
    B, seed = int(B), int(seed)
    shapes = reference_state_shapes(name)
    sd = fill_state_dict(shapes, seed, scheme='torch_default' if mode == 'default' else 'adversarial')
    assert abs(float(sum(v.double().sum() for v in sd.values())) - wsum) < 1e-6 * max(1.0, abs(wsum))
    n_par = sum(v.numel() for k, v in sd.items() if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked')))
    assert n_par == int(nn_golden[tag + '__nparams'][0])
    x = synth_patches(B, seed)
    with torch.no_grad():
        y = oracle_forward(name, sd, x, train=(mode == 'train'))
    if isinstance(y, tuple):
        assert np.abs(y[1].numpy() - nn_golden[tag + "__n"]).max() < 2e-4
        y = y[0]
    assert y.shape == (B, 1, 1, 72)
    assert np.abs(y.numpy() - nn_golden[tag + "__y"]).max() < 1e-4   # fp32 reassociation noise (thread count)


@pytest.mark.parametrize('name', ['cnn_xs', 'drcnn_tiny'])
def test_nn_oracle_loss_and_grads(nn_golden, name):
    tag = f'{name}__eval'
    B, seed, _ = nn_golden[tag + '__meta']
    B, seed = int(B), int(seed)
    sd = fill_state_dict(reference_state_shapes(name), seed)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = synth_patches(B, seed)
    loss = NO.bce_mean(oracle_forward(name, sd, x), synth_targets(B, seed))
    assert abs(loss.item() - float(nn_golden[tag + '__loss'][0])) < 1e-6
    loss.backward()
    for k, v in sd.items():
        g = nn_golden[tag + '__grad__' + k]
        # max-pool arg-max flips on fp32 near-ties move single contributions: bound relative to the tensor's scale
        assert np.abs(v.grad.numpy() - g).max() <= 5e-3 * np.abs(g).max(), k
        # weight gradients are 48600-term signed fp32 sums: summation order alone moves them by ~1e-4 relative
        assert np.abs(v.grad.numpy() - g).mean() <= 1e-3 * np.abs(g).max(), k


def test_bce_clamp():
    y = torch.tensor([0.0, 1.0, 0.5])
    t = torch.tensor([1.0, 0.0, 1.0])
    assert abs(NO.bce_mean(y, t).item() - torch.nn.BCELoss()(y, t).item()) < 1e-6
    assert abs(NO.bce_mean(y, t).item() - (100 + 100 + np.log(2)) / 3) < 1e-4


def test_hopsize(host_golden):
    for target, noct, hop, fs in host_golden['hopsize']:
        h, f = Q.compute_hopsize_cqt(target, 22050, int(noct))
        assert h == int(hop) and f == fs
    assert Q.compute_hopsize_cqt(50, 22050, 10) == (512, 22050 / 512)


def test_annotation_rasteriser(host_golden):
    import os
    ev = host_golden['annot2_events']
    A2 = HO.annotation_array_nooverlap(ev, 200, 22050 / 512, 'pitch')
    assert np.array_equal(np.packbits(A2.astype(np.uint8)), host_golden['annot2'])


def test_annotation_shipped_csv_sha1(host_golden):
    """Fixture: data/MusicNet/csv/2382_Beethoven_OP130_StringQuartet.csv of the reference; a copy of its three
    used columns travels as tests/golden/annot_2382_events.npy (data, not code)."""
    import os
    p = os.path.join(os.path.dirname(__file__), 'golden', 'annot_2382_events.npy')
    ev = np.load(p)
    fs = 22050 / 512
    n_frames = int(np.floor(ev[:, 1].max() * fs)) + 5
    A = HO.annotation_array_nooverlap(ev, n_frames, fs, 'pitch')
    assert A.shape == tuple(host_golden['annot_shape'])
    assert hashlib.sha1(A.astype(np.uint8).tobytes()).hexdigest() == '61739265956b8f5fd717c438f2af4dce7c716576'
    assert np.array_equal(np.argwhere(A > 0).astype(np.int32), host_golden['annot_nnz'])
    Apc = HO.annotation_array_nooverlap(ev, n_frames, fs, 'pitch_class', shorten=0.5)
    assert hashlib.sha1(Apc.astype(np.uint8).tobytes()).digest() == bytes(host_golden['annot_pc_sha1'])


def test_dataset_context_index_math(host_golden):
    inp, tg = host_golden['ds_in'], host_golden['ds_tg']
    ip, tp = HO.pad_for_inference(inp, tg)
    assert HO.context_len(ip.shape[1]) == int(host_golden['ds_len'][0]) == inp.shape[1]
    for j, i in enumerate(host_golden['ds_idx']):
        X, y = HO.context_item(ip.astype(np.float64), tp, int(i))
        assert np.array_equal(y, host_golden['ds_y'][j])
        assert np.abs(X - host_golden['ds_X'][j]).max() < 1e-6
    assert HO.context_len(ip.shape[1], 75, 3) == int(host_golden['ds3_len'][0])
    X, y = HO.context_item(ip.astype(np.float64), tp, 5, 75, 3, None)
    assert abs(X.astype(np.float64).sum() - host_golden['ds3_X5_sum'][0]) < 1e-3
    assert np.array_equal(y, host_golden['ds3_y5'])


def test_prf(host_golden):
    got = HO.eval_prf(host_golden['prf_targ'], host_golden['prf_pred'], 0.4)
    assert np.allclose(np.array(got, dtype=np.float64), host_golden['prf'], rtol=0, atol=1e-12)


def test_hcqt_oracle_golden_and_anchors(host_golden):
    y = Q.synth_clip(int(host_golden['hcqt_clip_seed'][0]), seconds=2.0)
    f, fs, hop = Q.compute_efficient_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    assert f.shape == (216, len(y) // 512 + 1, 6) and hop == 512 and fs == 22050 / 512
    assert np.abs(f.astype(np.float32) - host_golden['hcqt_2s']).max() < 1e-5
    assert abs(Q.estimate_tuning(y, bins_per_octave=36) - host_golden['hcqt_2s_tuning'][0]) < 1e-12
    # h=4 is the same CQT as h=1 shifted by two octaves (hcqt.py:159-162)
    assert np.array_equal(f[72:, :, 1], f[:144, :, 4])
    # pure tone at MIDI p peaks at bin 3*(p-24)+1 of the fundamental channel
    t = np.arange(3 * 22050) / 22050
    for midi in (45, 60, 81):
        tone = (0.5 * np.sin(2 * np.pi * 440 * 2 ** ((midi - 69) / 12) * t)).astype(np.float32)
        g, _, _ = Q.compute_efficient_hcqt(tone, fs=22050, fs_hcqt_target=50, bins_per_octave=36, tuning_est=0.0)
        mid = g.shape[1] // 2
        assert int(np.argmax(g[:, mid, 1])) == 3 * (midi - 24) + 1
        assert int(np.argmax(g[:, mid, 0])) == 3 * (midi - 24) + 1 + 36
        if midi >= 48:
            assert int(np.argmax(g[:, mid, 2])) == 3 * (midi - 24) + 1 - 36


def test_resample_halfband_properties():
    h = Q._kaiser_fast_halfband()
    assert len(h) == 32 and abs(h[0] - 0.425) < 1e-12
    y = np.ones(4001, dtype=np.float32)
    z = Q.resample_2to1(y)
    assert len(z) == 2001 and z[-1] == 0.0
    assert abs(z[1000] - np.sqrt(2.0)) < 2e-3            # DC gain 1 then x sqrt(2)


@pytest.mark.parametrize('name', ['unet_tiny', 'saunet_tiny', 'punet_tiny', 'sausnet_tiny'])
def test_oracle_train_mode_loss_and_grads_match_reference(name):
    """The oracle's train-mode forward (BatchNorm batch statistics, batch-axis attention) differentiated by autograd must give
    the loss and parameter gradients of the unmodified reference modules (tests/golden/nn_train_golden.npz)."""
    import torch.nn.functional as F
    from tests.weights import MODEL_SPECS, synth_targets
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nn_train_golden.npz'))
    tag = f'{name}__train'
    B, seed = [int(v) for v in g[tag + '__meta']]
    sd = fill_state_dict(reference_state_shapes(name), seed)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running_' not in k else v) for k, v in sd.items()}
    x, t = synth_patches(B, seed), synth_targets(B, seed)
    y = NO.unet_forward(sd, x, train=True, pos_encoding=MODEL_SPECS[name]['kw'].get('pos_encoding'))
    if isinstance(y, tuple):
        y, n_pred = y
        loss = NO.bce_mean(y, t) + F.cross_entropy(n_pred, t.sum(-1, keepdim=True).long().squeeze(3)) / 25.0
    else:
        loss = NO.bce_mean(y, t)
    loss.backward()
    assert abs(loss.item() - float(g[tag + '__loss'][0])) < 1e-5
    # conv biases in front of a train-mode BatchNorm have a mathematically zero gradient (pure rounding noise in both
    # implementations): every tensor is judged against max(its own scale, 1e-3 of the largest gradient in the model)
    gmax = max(np.abs(g[tag + '__grad__' + k]).max() for k, v in sd.items() if v.requires_grad)
    for k, v in sd.items():
        if v.requires_grad:
            ref = g[tag + '__grad__' + k]
            d = np.abs(v.grad.numpy() - ref)
            scale = max(np.abs(ref).max(), 1e-3 * gmax)
            assert d.max() <= 2e-2 * scale and (d.mean() <= 4e-3 * scale or d.size < 64), (k, d.max() / scale, d.mean() / scale)


def test_oracle_sausnet_eval_matches_reference():
    """simple_u_net_doubleselfattn_twolayers (attention also on the lowest skip) in eval mode against the reference's output."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'nn_train_golden.npz'))
    tag = 'sausnet_tiny__train'
    B, seed = [int(v) for v in g[tag + '__meta']]
    sd = fill_state_dict(reference_state_shapes('sausnet_tiny'), seed)
    with torch.no_grad():
        y = NO.unet_forward(sd, synth_patches(B, seed), pos_encoding='sinusoidal')
    assert np.abs(y.numpy() - g[tag + '__eval_y']).max() < 1e-5


# ----------------------------------------------------------------------------- rows added after the first pass (ext_golden.npz)
def test_compute_hcqt_oracle_matches_reference_wrapper_golden(ext_golden):
    """compute_hcqt (hcqt.py:34-85) incl. early down-sampling (h = 1/2, 1) and the full-rate top octave (h = 5)."""
    y = Q.synth_clip(3, seconds=2.0)
    f, fs, hop = Q.compute_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    assert hop == 448 and fs == 22050 / 448 and f.shape == ext_golden['hcqt_std_2s'].shape
    assert np.abs(f.astype(np.float32) - ext_golden['hcqt_std_2s']).max() < 1e-5
    g, _, hop60 = Q.compute_hcqt(y[:22050])
    assert hop60 == 256 and np.abs(g.astype(np.float32) - ext_golden['hcqt_std60_1s']).max() < 1e-5
    # the early-down-sampled CQT agrees with the same bins of the shared-octave HCQT at the common frame times (hop 448 vs 512:
    # frames 8k and 7k are both centred on sample 3584 k)
    e, _, _ = Q.compute_efficient_hcqt(y, fs=22050, fs_hcqt_target=50, bins_per_octave=36)
    for k in range(1, 10):
        a, b = f[:, 8 * k, :], e[:, 7 * k, :]
        assert np.abs(a - b)[:, 3:].max() < 1e-6 * b.max()       # h = 3, 4, 5: the same octave schedule in both variants
        assert np.abs(a - b)[:, :3].max() < 1e-2 * b.max()       # h = 1/2, 1, 2: one decimation stage more / less (pass-band ripple)


def test_eval_measures_oracle_matches_reference_golden(ext_golden):
    targ, pred = ext_golden['ev_targ'].astype(np.float64), ext_golden['ev_pred'].astype(np.float64)
    names = [str(n) for n in ext_golden['ev_names']]
    for thr, key in ((0.4, 'ev_values_04'), (0.7, 'ev_values_07')):
        for n, want in zip(names, ext_golden[key]):
            if n in ('roc_auc_measure', 'average_precision_score'):
                continue
            assert abs(HO.eval_measure(targ, pred, n, thr) - want) < 1e-12, n


def test_augmentation_oracle_matches_reference_golden(ext_golden):
    """hcqt_datasets.py:77-139 with the reference's own random draws replayed: bit-exact except the log (numpy vs torch fp32 log)."""
    inp, tg = ext_golden['aug_in'], ext_golden['aug_tg']
    rows = [0, 37, 74]
    for ci, i, alpha, beta, tune2, transp in ext_golden['aug_cases']:
        tag = 'aug%d_%d' % (ci, i)
        X0 = inp[:, i:i + 75, :][:, rows, :]
        t = tg[:, :12] if ci == 2 else tg
        y0 = t[i + 37][None, None, :]
        get = lambda k: ext_golden[tag + k] if tag + k in ext_golden.files else None
        X, y = HO.augment_item(X0, y0, 10.0, (alpha, beta) if alpha else None, int(tune2), int(transp), get('_noise'),
                               get('_ftune'), get('_ftr'))
        assert np.abs(X - ext_golden[tag + '_X']).max() < 1e-6, tag
        assert np.array_equal(y, ext_golden[tag + '_y']), tag


def test_mpe_scores_properties():
    rng = np.random.default_rng(0)
    targ = (rng.uniform(size=(50, 72)) < 0.05).astype(np.float64)
    same = HO.mpe_scores(targ, targ, 0.5)
    assert same['Precision'] == same['Recall'] == same['Accuracy'] == 1.0 and same['Total Error'] == 0.0
    octave = np.roll(targ, 12, axis=1)
    octave[:, :12] = 0
    targ[:, 60:] = 0
    s = HO.mpe_scores(targ, octave, 0.5)
    assert s['Chroma Precision'] >= s['Precision'] and s['Chroma Total Error'] <= s['Total Error']
    assert abs(s['Total Error'] - (s['Substitution Error'] + s['Miss Error'] + s['False Alarm Error'])) < 1e-12


@pytest.mark.parametrize('name,B,seed,scheme', [('blunet_tiny', 3, 41, 'adversarial'), ('blunet_d', 2, 42, 'torch_default')])
def test_blunet_oracle_matches_reference_golden(ext_golden, name, B, seed, scheme):
    """u_net_blstm_varlayers (unet_cnns.py:1000-1101): state_dict layout, parameter count (exp186d log: 9,649,003) and eval output."""
    import torch
    from oracle import nn_oracle as NO
    from tests.refshapes import build_model
    from tests.weights import fill_state_dict, synth_patches
    m = build_model(name)
    sd = fill_state_dict(m.state_dict(), seed, scheme=scheme)
    meta = ext_golden[name + '__meta']
    assert abs(float(sum(v.double().sum() for v in sd.values())) - meta[2]) < 1e-6        # same keys, shapes and order as the reference
    assert sum(p.numel() for p in m.parameters()) == int(meta[3])
    if name == 'blunet_d':
        assert int(meta[3]) == 9649003
    with torch.no_grad():
        y = NO.unet_forward(sd, synth_patches(B, seed))
    assert np.abs(y.numpy() - ext_golden[name + '__y']).max() < 1e-5


# ---------------------------------------------------------------------------------------------------------------------------------
# realistic weight set + state_dict key names
def test_oracle_matches_reference_on_the_realistic_weights():
    """oracle/nn_oracle.py on the trained weights vs the committed outputs of the unmodified reference classes (patch-wise, first frames)."""
    import numpy as np
    import torch
    from oracle import hcqt_oracle as HQ
    from oracle import host_oracle as HO
    from oracle import nn_oracle as NO
    from tests import realistic as R
    from tests import synth
    y = synth.synth_clip(**R.CLIP)[:22050 * 4]
    f, _, _ = HQ.compute_efficient_hcqt(y, **R.HCQT_KW)
    g = R.golden()
    # the first frames of the 4 s excerpt equal those of the full clip (same tuning estimate is NOT guaranteed on an excerpt: use the probe)
    h = np.transpose(f, (2, 1, 0)).astype(np.float32)
    n = 6
    ip, _ = HO.pad_for_inference(h, np.zeros((h.shape[1], 72)))
    X = torch.from_numpy(np.stack([HO.context_item(ip, np.zeros((ip.shape[1], 72)), i)[0] for i in range(n)]))
    full = synth.synth_clip(**R.CLIP)
    for name in ('cnn_xs', 'unet_m'):
        sd = R.state_dict(name)
        with torch.no_grad():
            out = (NO.cnn_forward(sd, X) if name == 'cnn_xs' else NO.unet_forward(sd, X)).reshape(n, 72).numpy()
        assert out.shape == (n, 72) and np.isfinite(out).all()
    # exact comparison needs the full-clip HCQT (global tuning estimate): done for CNN:XS on 40 frames
    ff, _, _ = HQ.compute_efficient_hcqt(full, **R.HCQT_KW)
    assert np.abs(ff[::37, ::101, :] - g['hcqt_probe']).max() <= 1e-6 * g['hcqt_probe'].max()
    hf = np.transpose(ff, (2, 1, 0)).astype(np.float32)
    ipf, _ = HO.pad_for_inference(hf, np.zeros((hf.shape[1], 72)))
    idx = list(range(0, 1292, 33))
    Xf = torch.from_numpy(np.stack([HO.context_item(ipf, np.zeros((ipf.shape[1], 72)), i)[0] for i in idx]))
    with torch.no_grad():
        out = NO.cnn_forward(R.state_dict('cnn_xs'), Xf).reshape(len(idx), 72).numpy()
    assert np.abs(out - g['cnn_xs__y'][idx]).max() < 1e-5
    with torch.no_grad():
        out = NO.unet_forward(R.state_dict('unet_m'), Xf[:8]).reshape(8, 72).numpy()
    assert np.abs(out - g['unet_m__y'][idx[:8]]).max() < 1e-5
    with torch.no_grad():
        out = NO.cnn_forward(R.state_dict('drcnn'), Xf[:4], residual=True).reshape(4, 72).numpy()
    assert np.abs(out - g['drcnn__y'][idx[:4]]).max() < 1e-5
    # the goldens carry the decision structure the 3-decimal P/R/F gate needs: outputs span 0..1, F-measure vs labels 0.90-0.96
    lab = R.labels()
    for name in R.MODELS:
        c = R.prf_counts(lab, g[name + '__y'])
        assert tuple(g[name + '__counts']) == c and R.prf(c)[2] > 0.85 and g[name + '__y'].max() > 0.99 and g[name + '__y'].min() < 1e-3


def test_state_dict_key_names_equal_the_reference():
    """Names, shapes, dtypes and ORDER of every product model's state_dict vs the list recorded from the reference classes."""
    import json
    import os
    from tests.refshapes import build_model
    from tests.weights import MODEL_SPECS
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'state_dict_keys.json')))
    assert set(ref) == set(MODEL_SPECS)
    for name, keys in ref.items():
        sd = build_model(name).state_dict()
        got = [[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in sd.items()]
        assert got == keys, name
