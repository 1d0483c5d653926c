"""Generates the committed golden fixtures by running the UNMODIFIED reference (imported by path from
/root/reference, never copied) in the build container.  Re-run:  python tests/golden/make_golden.py

Outputs (tests/golden/*.npz) are what the `-m "not gpu"` suite pins the oracle against and what the
`-m gpu` suite compares the CUDA path with.  /root/reference is NOT needed to run the tests."""
import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from tests.weights import MODEL_SPECS, fill_state_dict, synth_patches, synth_targets  # noqa: E402
from oracle import hcqt_oracle as HO  # noqa: E402


def load_reference_hcqt():
    """hcqt.py imports matplotlib / IPython / librosa at module scope; none is installed.  Stub them; the
    librosa stub routes cqt / estimate_tuning to the oracle restatement so the reference's own bookkeeping
    (hop size, harmonic plan, slicing, annotation rasteriser) runs unmodified."""
    for name in ('matplotlib', 'matplotlib.pyplot', 'IPython', 'IPython.display'):
        sys.modules.setdefault(name, types.ModuleType(name))
    lib = types.ModuleType('librosa')
    lib.note_to_hz = lambda n: HO.C1_HZ
    lib.cqt = lambda y, sr, hop_length, fmin, n_bins, bins_per_octave, tuning: HO.cqt(
        y, sr=sr, hop_length=hop_length, fmin=fmin, n_bins=n_bins, bins_per_octave=bins_per_octave, tuning=tuning)
    lib.estimate_tuning = lambda y, bins_per_octave: HO.estimate_tuning(y, bins_per_octave=bins_per_octave)
    sys.modules['librosa'] = lib
    spec = importlib.util.spec_from_file_location('ref_hcqt', os.path.join(REF, 'libdl/data_preprocessing/hcqt.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference_model(name):
    import libdl.nn_models as M
    spec = MODEL_SPECS[name]
    kw = dict(spec['kw'])
    if kw.get('pos_encoding') == 'sinusoidal':
        # unet_cnns.py:121 hard-wires device="cuda:0"; drop the device kwarg while constructing on CPU
        z = torch.zeros
        torch.zeros = lambda *a, **k: z(*a, **{kk: vv for kk, vv in k.items() if kk != 'device'})
        try:
            m = getattr(M, spec['cls'])(**kw)
        finally:
            torch.zeros = z
    else:
        m = getattr(M, spec['cls'])(**kw)
    return m


def nn_goldens():
    torch.set_num_threads(8)
    cases = [  # (model, batch, seed, mode); mode 'default' = eval with the reference's default-init weight distribution
        ('cnn_xs', 4, 111, 'default'), ('drcnn', 2, 114, 'default'), ('unet_m', 2, 116, 'default'), ('punet', 2, 118, 'default'),
        ('saunet_l', 4, 120, 'default'),
        ('cnn_xs', 4, 11, 'eval'), ('drcnn_tiny', 3, 12, 'eval'), ('dcnn_tiny', 3, 13, 'eval'),
        ('drcnn', 2, 14, 'eval'), ('unet_tiny', 3, 15, 'eval'), ('unet_tiny', 3, 15, 'train'),
        ('unet_m', 2, 16, 'eval'), ('punet_tiny', 3, 17, 'eval'), ('punet', 2, 18, 'eval'),
        ('saunet_tiny', 5, 19, 'eval'), ('saunet_l', 4, 20, 'eval'), ('saunet_tiny', 5, 19, 'train'),
    ]
    out = {}
    for name, B, seed, mode in cases:
        m = build_reference_model(name)
        sd = fill_state_dict(m.state_dict(), seed, scheme='torch_default' if mode == 'default' else 'adversarial')
        m.load_state_dict(sd)
        for mod in m.modules():                                # dropout off: RNG-free parity
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train(mode == 'train')
        x = synth_patches(B, seed)
        with torch.no_grad():
            y = m(x)
        tag = f'{name}__{mode}'
        if isinstance(y, tuple):
            out[tag + '__y'] = y[0].numpy()
            out[tag + '__n'] = y[1].numpy()
        else:
            out[tag + '__y'] = y.numpy()
        out[tag + '__meta'] = np.array([B, seed, float(sum(v.double().sum() for v in sd.values()))])
        n_par = sum(p.numel() for p in m.parameters())
        out[tag + '__nparams'] = np.array([n_par])
        print(tag, 'params', n_par, 'y', out[tag + '__y'].reshape(-1)[:3])
        if mode == 'eval' and name in ('cnn_xs', 'drcnn_tiny'):
            # loss + gradients for the training path (BCELoss mean, exp126a:87,323)
            m.train(True)
            yt = synth_targets(B, seed)
            m.zero_grad()
            loss = torch.nn.BCELoss(reduction='mean')(m(x), yt)
            loss.backward()
            out[tag + '__loss'] = np.array([loss.item()])
            for k, p in m.named_parameters():
                out[tag + '__grad__' + k] = p.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'nn_golden.npz'), **out)


def host_goldens():
    ref = load_reference_hcqt()
    out = {}
    # H0 hop sizes
    hs = []
    for target, noct in ((50, 10), (91, 6), (43.0, 8), (100, 7), (25, 10)):
        hop, fs = ref.compute_hopsize_cqt(target, 22050, noct)
        hs.append((target, noct, hop, fs))
    out['hopsize'] = np.array(hs)
    # H6 annotation rasteriser on the shipped CSV (notebook 01 cell 7 usage)
    csv = np.loadtxt(os.path.join(REF, 'data/MusicNet/csv/2382_Beethoven_OP130_StringQuartet.csv'),
                     delimiter=',', skiprows=1, usecols=(0, 1, 3))
    ev = csv.copy()
    ev[:, :2] /= 44100.0
    fs_hcqt = 22050 / 512
    n_frames = int(np.floor(ev[:, 1].max() * fs_hcqt)) + 5
    A = ref.compute_annotation_array_nooverlap(ev.copy(), np.zeros((216, n_frames, 6)), fs_hcqt, annot_type='pitch')
    out['annot_shape'] = np.array(A.shape)
    out['annot_sha1'] = np.frombuffer(hashlib.sha1(A.astype(np.uint8).tobytes()).digest(), dtype=np.uint8)
    out['annot_nnz'] = np.argwhere(A > 0).astype(np.int32)
    Apc = ref.compute_annotation_array_nooverlap(ev.copy(), np.zeros((216, n_frames, 6)), fs_hcqt,
                                                 annot_type='pitch_class', shorten=0.5)
    out['annot_pc_sha1'] = np.frombuffer(hashlib.sha1(Apc.astype(np.uint8).tobytes()).digest(), dtype=np.uint8)
    # synthetic dense case with many vanishing / colliding events
    rng = np.random.default_rng(5)
    st = np.sort(rng.uniform(0, 4.0, size=300))
    ev2 = np.stack([st, st + rng.choice([0.001, 0.01, 0.03, 0.2], size=300), rng.integers(30, 90, size=300)], 1)
    A2 = ref.compute_annotation_array_nooverlap(ev2.copy(), np.zeros((216, 200, 6)), fs_hcqt, annot_type='pitch')
    out['annot2_events'] = ev2
    out['annot2'] = np.packbits(A2.astype(np.uint8))
    # H4/H5 dataset_context
    from libdl.data_loaders import dataset_context
    inp = rng.uniform(0, 1, size=(6, 40, 216)) ** 4
    tg = (rng.uniform(size=(40, 72)) < 0.05).astype(np.float64)
    half = 75 // 2
    inp_c = torch.from_numpy(np.pad(inp, ((0, 0), (half, half + 1), (0, 0))))
    tg_c = torch.from_numpy(np.pad(tg, ((half, half + 1), (0, 0))))
    ds = dataset_context(inp_c, tg_c, {'context': 75, 'stride': 1, 'compression': 10})
    out['ds_in'] = inp.astype(np.float32)
    out['ds_tg'] = tg.astype(np.float32)
    out['ds_len'] = np.array([len(ds)])
    idx = [0, 1, 17, 39]
    out['ds_idx'] = np.array(idx)
    out['ds_X'] = np.stack([ds[i][0].numpy() for i in idx])
    out['ds_y'] = np.stack([ds[i][1].numpy() for i in idx])
    ds3 = dataset_context(inp_c, tg_c, {'context': 75, 'stride': 3, 'compression': None})
    out['ds3_len'] = np.array([len(ds3)])
    out['ds3_X5_sum'] = np.array([ds3[5][0].double().sum().item()])
    out['ds3_y5'] = ds3[5][1].numpy()
    # N12 P/R/F through the reference eval function (libfmp import needs the same stubs)
    spec = importlib.util.spec_from_file_location('ref_c5', os.path.join(REF, 'libfmp/c5/c5s2_chord_rec_template.py'))
    sys.modules.setdefault('numba', __import__('numba'))
    for name in ('libfmp', 'libfmp.b', 'libfmp.c3', 'libfmp.c4', 'matplotlib.colors'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].colors = sys.modules['matplotlib.colors']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['libfmp.b'].MultiplePlot = object
    sys.modules['libfmp.b'].plot_matrix = None
    sys.modules['libfmp.b'].plot_segments = None
    sys.modules['libfmp.b'].read_csv = None
    try:
        c5 = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(c5)
        pred = rng.uniform(size=(500, 72))
        targ = (rng.uniform(size=(500, 72)) < 0.3)
        out['prf_pred'] = pred.astype(np.float32)
        out['prf_targ'] = targ
        out['prf'] = np.array(c5.compute_eval_measures(targ, pred.astype(np.float32) >= 0.4), dtype=np.float64)
    except Exception as e:                                     # pragma: no cover
        print('libfmp.c5 not importable even with stubs:', e)
    # H1-H3: the reference wrapper (with the oracle standing in for librosa) vs the oracle's own wrapper
    y = HO.synth_clip(3, seconds=2.0)
    f_ref, fs_h, hop = ref.compute_efficient_hcqt(y, fs=22050, fmin=HO.C1_HZ, fs_hcqt_target=50, bins_per_octave=36,
                                                  num_octaves=6, num_harmonics=5, num_subharmonics=1)
    f_or, fs_o, hop_o = HO.compute_efficient_hcqt(y, fs=22050, fmin=HO.C1_HZ, fs_hcqt_target=50, bins_per_octave=36,
                                                  num_octaves=6, num_harmonics=5, num_subharmonics=1)
    assert np.array_equal(f_ref, f_or) and hop == hop_o and fs_h == fs_o, 'oracle wrapper != reference wrapper'
    out['hcqt_clip_seed'] = np.array([3])
    out['hcqt_2s'] = f_ref.astype(np.float32)
    out['hcqt_2s_tuning'] = np.array([HO.estimate_tuning(y, bins_per_octave=36)])
    np.savez_compressed(os.path.join(HERE, 'host_golden.npz'), **out)
    print('annot sha1', bytes(out['annot_sha1']).hex(), 'nnz', len(out['annot_nnz']), 'shape', A.shape)


def train_goldens():
    """Loss + parameter gradients of the REFERENCE U-Net-family modules in train mode (BatchNorm batch statistics, dropout p=0):
    BCELoss(mean) and, for the PUnet, + CrossEntropyLoss(n_pred, sum(labels).long())/25 (RETRAIN4_exp195f...rerun1.py:343-346)."""
    torch.set_num_threads(8)
    out = {}
    for name, B, seed in (('unet_tiny', 3, 15), ('saunet_tiny', 5, 19), ('punet_tiny', 3, 17), ('sausnet_tiny', 5, 23)):
        m = build_reference_model(name)
        sd = fill_state_dict(m.state_dict(), seed)
        m.load_state_dict(sd)
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
        m.train(True)
        x, yt = synth_patches(B, seed), synth_targets(B, seed)
        m.zero_grad()
        y = m(x)
        tag = f'{name}__train'
        if isinstance(y, tuple):
            y, n_pred = y
            n_target = torch.sum(yt, dim=-1, keepdims=True).long().squeeze(3)
            loss = torch.nn.BCELoss(reduction='mean')(y, yt) + torch.nn.CrossEntropyLoss(reduction='mean')(n_pred, n_target) / 25.0
            out[tag + '__n'] = n_pred.detach().numpy()
        else:
            loss = torch.nn.BCELoss(reduction='mean')(y, yt)
        loss.backward()
        out[tag + '__y'] = y.detach().numpy()
        out[tag + '__loss'] = np.array([loss.item()])
        out[tag + '__meta'] = np.array([B, seed])
        for k, p in m.named_parameters():
            out[tag + '__grad__' + k] = p.grad.numpy().copy()
        for k, v in m.state_dict().items():
            if 'running_' in k:
                out[tag + '__stat__' + k] = v.numpy().copy()
        m.load_state_dict(sd)                               # (the train-mode pass above moved the BatchNorm running statistics)
        m.eval()
        with torch.no_grad():
            ye = m(x)
        out[tag + '__eval_y'] = (ye[0] if isinstance(ye, tuple) else ye).numpy()
        print(tag, 'loss', loss.item())
    np.savez_compressed(os.path.join(HERE, 'nn_train_golden.npz'), **out)


def load_reference_eval_metrics():
    """libdl/metrics/eval_metrics.py imports matplotlib, IPython, librosa, libfmp.c3/c5, libdl.data_preprocessing, sklearn and
    mir_eval at module scope.  sklearn is installed; libfmp.c3 / c5 are loaded from the reference's own files; the rest is stubbed."""
    load_reference_hcqt()
    for name in ('matplotlib.colors', 'libfmp', 'libfmp.b', 'libfmp.c4', 'mir_eval', 'mir_eval.multipitch'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].colors = sys.modules['matplotlib.colors']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mir_eval'].multipitch = sys.modules['mir_eval.multipitch']
    for attr in ('MultiplePlot', 'plot_matrix', 'plot_segments', 'read_csv'):
        setattr(sys.modules['libfmp.b'], attr, None)
    sys.modules['libfmp.b'].MultiplePlot = object
    for sub, path in (('c3', 'libfmp/c3/c3s1_post_processing.py'), ('c5', 'libfmp/c5/c5s2_chord_rec_template.py')):
        spec = importlib.util.spec_from_file_location('libfmp.' + sub, os.path.join(REF, path))
        mod = importlib.util.module_from_spec(spec)
        sys.modules['libfmp.' + sub] = mod
        setattr(sys.modules['libfmp'], sub, mod)
        spec.loader.exec_module(mod)
    dp = types.ModuleType('libdl.data_preprocessing')
    saved = sys.modules.get('libdl.data_preprocessing')
    sys.modules['libdl.data_preprocessing'] = dp
    try:
        spec = importlib.util.spec_from_file_location('ref_eval_metrics', os.path.join(REF, 'libdl/metrics/eval_metrics.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is not None:
            sys.modules['libdl.data_preprocessing'] = saved
        else:
            del sys.modules['libdl.data_preprocessing']
    return mod


AUG_ROWS = [0, 37, 74]


def ext_goldens():
    """Goldens of the rows added after the first pass (SURVEY.md 8b compute_hcqt; 8f rows 1 and 4) -> ext_golden.npz."""
    out = {}
    ref = load_reference_hcqt()
    # compute_hcqt (hcqt.py:34-85): the reference wrapper with the oracle standing in for librosa vs the oracle's own wrapper;
    # hop 448 at the paper's resolution -> early down-sampling for h = 1/2, 1 and a full-rate top octave for h = 5
    y = HO.synth_clip(3, seconds=2.0)
    kw = dict(fs=22050, fmin=HO.C1_HZ, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)
    f_ref, fs_h, hop = ref.compute_hcqt(y, **kw)
    f_or, fs_o, hop_o = HO.compute_hcqt(y, **kw)
    assert np.array_equal(f_ref, f_or) and hop == hop_o == 448 and fs_h == fs_o, 'oracle compute_hcqt != reference wrapper'
    out['hcqt_std_2s'] = f_ref.astype(np.float32)
    # the reference's default arguments (Bittner et al.: 60 bins per octave, hop 256): early down-sampling by 4 for the sub-harmonic
    kw60 = dict(fs=22050, fmin=HO.C1_HZ)
    f_ref, fs_h, hop = ref.compute_hcqt(y[:22050], **kw60)
    f_or, _, hop_o = HO.compute_hcqt(y[:22050], **kw60)
    assert np.array_equal(f_ref, f_or) and hop == hop_o == 256
    out['hcqt_std60_1s'] = f_ref.astype(np.float32)

    # evaluation measures through the reference function
    em = load_reference_eval_metrics()
    rng = np.random.default_rng(11)
    targ = (rng.uniform(size=(400, 72)) < 0.06).astype(np.float64)
    pred = np.clip(0.75 * targ * rng.uniform(0.3, 1.3, size=targ.shape) + rng.uniform(size=targ.shape) ** 6, 0, 1).astype(np.float32)
    targ[7] = 0          # silent reference frame -> libfmp's unit-vector fallback / empty reference set
    pred[9] = 0          # silent estimate
    targ[9] = 0
    pred[11, :] = 1e-12
    names = ['precision', 'recall', 'f_measure', 'cosine_sim', 'binary_crossentropy', 'euclidean_distance', 'binary_accuracy',
             'soft_accuracy', 'accum_energy', 'roc_auc_measure', 'average_precision_score']
    out['ev_targ'] = targ.astype(np.float32)
    out['ev_pred'] = pred
    for thr in (0.4, 0.7):
        d = em.calculate_eval_measures(targ, pred.astype(np.float64), names, threshold=thr)
        out['ev_values_%02d' % int(thr * 10)] = np.array([d[k] for k in names], dtype=np.float64)
    out['ev_names'] = np.array(names)

    # augmentations: the reference __getitem__ with every torch.randint / torch.normal result recorded
    from libdl.data_loaders import dataset_context
    inp = np.abs(rng.normal(0, 0.05, size=(6, 130, 216)))
    tg = (rng.uniform(size=(130, 72)) < 0.05).astype(np.float64)
    out['aug_in'] = inp.astype(np.float32)
    out['aug_tg'] = tg.astype(np.float32)
    real_randint, real_normal = torch.randint, torch.normal
    rec = []

    def randint(*a, **k):
        r = real_randint(*a, **k)
        rec.append(('i', r.clone()))
        return r

    def normal(*a, **k):
        r = real_normal(*a, **k)
        rec.append(('n', r.clone()))
        return r

    cases = []
    for ci, (params, idxs) in enumerate((
            ({'aug:randomeq': 20, 'aug:tuning': True, 'aug:transpsemitones': 5}, list(range(0, 55, 3))),
            ({'aug:randomeq': 20, 'aug:noisestd': 1e-4, 'aug:tuning': True, 'aug:transpsemitones': 5}, [1, 20, 40]),
            ({'aug:transpsemitones': 2, 'targettype': 'pitch_class'}, [2, 9, 30]))):
        p = dict({'context': 75, 'stride': 1, 'compression': 10}, **params)
        tgt = tg[:, :12] if params.get('targettype') == 'pitch_class' else tg
        ds = dataset_context(torch.from_numpy(inp.astype(np.float32).astype(np.float64)), torch.from_numpy(tgt), p)
        torch.manual_seed(100 + ci)
        for i in idxs:
            rec.clear()
            torch.randint, torch.normal = randint, normal
            try:
                X, yy = ds[i]
            finally:
                torch.randint, torch.normal = real_randint, real_normal
            ints = [int(r) for k, r in rec if k == 'i']
            norms = [r.numpy() for k, r in rec if k == 'n']
            pos = 0
            alpha = beta = 0
            if 'aug:randomeq' in params:
                n_pairs = (len(ints) - ('aug:tuning' in params) - ('aug:transpsemitones' in params)) // 2
                alpha, beta = ints[2 * n_pairs - 2], ints[2 * n_pairs - 1]
                pos = 2 * n_pairs
            tune2 = 0
            if 'aug:tuning' in params:
                tune2 = ints[pos] if True else 0
                pos += 1
            transp = ints[pos] if 'aug:transpsemitones' in params else 0
            noise = norms.pop(0) if 'aug:noisestd' in params else None
            fill_tune = norms.pop(0) if tune2 != 0 else None
            fill_tr = norms.pop(0) if transp != 0 else None
            assert not norms
            tag = 'aug%d_%d' % (ci, i)
            out[tag + '_dec'] = np.array([i, alpha, beta, tune2, transp], dtype=np.int64)
            out[tag + '_X'] = X.numpy().astype(np.float32)[:, AUG_ROWS, :]      # the chain is row-independent: 3 of 75 rows keep the fixture small
            out[tag + '_y'] = yy.numpy().astype(np.float32)
            if noise is not None:
                out[tag + '_noise'] = noise.astype(np.float32)[:, AUG_ROWS, :]
            if fill_tune is not None:
                out[tag + '_ftune'] = fill_tune.astype(np.float32)[:, AUG_ROWS, :]
            if fill_tr is not None:
                out[tag + '_ftr'] = fill_tr.astype(np.float32)[:, AUG_ROWS, :]
            cases.append((ci, i, alpha, beta, tune2, transp))
    # BLUnet (u_net_blstm_varlayers, exp186b/d/e): eval-mode outputs of the reference class
    for name, B, seed, scheme in (('blunet_tiny', 3, 41, 'adversarial'), ('blunet_d', 2, 42, 'torch_default')):
        m = build_reference_model(name)
        sd = fill_state_dict(m.state_dict(), seed, scheme=scheme)
        m.load_state_dict(sd)
        m.eval()
        with torch.no_grad():
            yb = m(synth_patches(B, seed))
        out[name + '__y'] = yb.numpy()
        out[name + '__meta'] = np.array([B, seed, float(sum(v.double().sum() for v in sd.values())), sum(p.numel() for p in m.parameters())])
        print(name, 'params', int(out[name + '__meta'][3]), 'y', out[name + '__y'].reshape(-1)[:3])
    # BLUnet training: loss + gradients of the reference class in train mode (BatchNorm batch statistics, dropout p = 0);
    # tensors above 20k elements (the LSTM matrices) are stored as every 5th element
    name, B, seed = 'blunet_s32', 3, 43
    m = build_reference_model(name)
    sd = fill_state_dict(m.state_dict(), seed)
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train(True)
    xb, yt = synth_patches(B, seed), synth_targets(B, seed)
    m.zero_grad()
    yb = m(xb)
    loss = torch.nn.BCELoss(reduction='mean')(yb, yt)
    loss.backward()
    out[name + '__train__y'] = yb.detach().numpy()
    out[name + '__train__loss'] = np.array([loss.item()])
    out[name + '__train__meta'] = np.array([B, seed])
    for k, p in m.named_parameters():
        g = p.grad.numpy().reshape(-1)
        out[name + '__train__grad__' + k] = (g[::5] if g.size > 20000 else g).copy()
    print(name, 'train loss', loss.item())
    # early_stopping (libdl/metrics/monitoring.py): stop decisions of the reference class over noisy metric sequences
    spec = importlib.util.spec_from_file_location('ref_monitoring', os.path.join(REF, 'libdl/metrics/monitoring.py'))
    mon = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mon)
    es_cfg = [('min', 1e-5, 12, False), ('max', 0.01, 3, False), ('min', 2.0, 4, True), ('max', 1.0, 2, True), ('min', 0, 0, False)]
    seqs = np.stack([np.linspace(1.0, 0.5, 60) + 0.05 * rng.standard_normal(60) * (np.arange(60) > 15), 0.5 + 0.4 * rng.uniform(size=60)])
    seqs[1, 40] = np.nan
    flags = np.zeros((len(es_cfg), 2, 60), dtype=np.int8)
    for ci, (mode, md, pat, pct) in enumerate(es_cfg):
        for si in range(2):
            es = mon.early_stopping(mode=mode, min_delta=md, patience=pat, percentage=pct)
            for k, v in enumerate(seqs[si]):
                flags[ci, si, k] = int(bool(es.step(float(v))))
    out['es_seqs'] = seqs
    out['es_flags'] = flags
    out['aug_cases'] = np.array(cases, dtype=np.int64)
    assert set(c[4] for c in cases) == {-2, -1, 0, 1, 2}, sorted(set(c[4] for c in cases))
    assert any(c[5] > 0 for c in cases) and any(c[5] < 0 for c in cases)
    np.savez_compressed(os.path.join(HERE, 'ext_golden.npz'), **out)
    print('ext goldens:', len(out), 'arrays;', len(cases), 'augmentation cases')


REALISTIC = ('cnn_xs', 'drcnn', 'unet_m')
REALISTIC_CLIP = dict(seed=777, seconds=30.0)
HCQT_KW = dict(fs=22050, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)


def realistic_goldens(src=os.path.join(ROOT, 'gpurun_out', 'realistic')):
    """The REALISTIC weight set: state_dicts trained by tools/train_realistic.py (this repo's own loop.fit on labelled synthetic audio, on
    the GPU) are committed as tests/golden/realistic_weights.npz, and the UNMODIFIED reference classes are run on them through the
    reference's own test loop (exp126a...py:413-436: np.pad -> dataset_context(stride 1, compression 10) -> DataLoader(batch 50) ->
    model(batch)) over the whole held-out 30 s clip.  Input = the oracle HCQT of tests.synth clip 777 (regenerated by the tests)."""
    import time
    from libdl.data_loaders import dataset_context
    import libdl.nn_models  # noqa: F401
    from tests import synth
    torch.set_num_threads(os.cpu_count() or 8)
    y, notes = synth.synth_clip_labeled(**REALISTIC_CLIP)
    f, _, _ = HO.compute_efficient_hcqt(y, **HCQT_KW)                      # [216, N, 6] float64, linear magnitudes
    inputs = np.transpose(f, (2, 1, 0))
    n_frames = inputs.shape[1]
    roll = synth.piano_roll(notes, n_frames)
    out = {'hcqt_sum': np.array([f.sum(dtype=np.float64)]), 'hcqt_probe': np.asarray(f[::37, ::101, :], dtype=np.float64),
           'roll': np.packbits(roll.astype(np.uint8)), 'n_frames': np.array([n_frames])}
    weights = {}
    half = 75 // 2
    for name in REALISTIC:
        sd = torch.load(os.path.join(src, name + '.pt'))
        m = build_reference_model(name)
        m.load_state_dict(sd)                                               # strict: names, shapes, dtypes of the reference
        m.eval()
        ic = torch.from_numpy(np.pad(inputs, ((0, 0), (half, half + 1), (0, 0))))
        tc = torch.from_numpy(np.pad(roll, ((half, half + 1), (0, 0))))
        gen = torch.utils.data.DataLoader(dataset_context(ic, tc, {'context': 75, 'stride': 1, 'compression': 10}), batch_size=50, shuffle=False)
        t0 = time.time()
        preds = []
        with torch.no_grad():
            for xb, _ in gen:
                yp = m(xb)
                preds.append(torch.squeeze(torch.squeeze(yp, 2), 1).numpy())
        pred = np.concatenate(preds, 0).astype(np.float32)
        assert pred.shape == (n_frames, 72)
        out[name + '__y'] = pred
        est = pred >= 0.4
        tp = int((est & (roll > 0)).sum())
        out[name + '__counts'] = np.array([tp, int(est.sum()) - tp, int((roll > 0).sum()) - tp])
        for k, v in sd.items():
            weights[name + '/' + k] = v.numpy()
        print(f'realistic {name}: {n_frames} patches in {time.time() - t0:.1f} s; outputs span {pred.min():.3g}..{pred.max():.3g}, '
              f'{int(est.sum())} active cells, TP/FP/FN = {out[name + "__counts"].tolist()}', flush=True)
    np.savez_compressed(os.path.join(HERE, 'realistic_golden.npz'), **out)
    np.savez_compressed(os.path.join(HERE, 'realistic_weights.npz'), **weights)


def grad_sample_index(numel, k=256):
    """Deterministic sample of a flattened tensor: every element when it has <= k, else k evenly spread positions."""
    return np.arange(numel) if numel <= k else (np.arange(k, dtype=np.int64) * numel) // k


def saunet_l_train_golden():
    """FULL-SIZE SAUnet:L (BASELINE configs[4]) in train mode on the reference class: loss, outputs, BatchNorm statistics and, per
    parameter, the gradient's L2 norm plus a 256-element sample (the 8.1 M gradient values themselves would be 32 MB)."""
    torch.set_num_threads(os.cpu_count() or 8)
    name, B, seed = 'saunet_l', 4, 31
    m = build_reference_model(name)
    sd = fill_state_dict(m.state_dict(), seed, scheme='torch_default')
    m.load_state_dict(sd)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train(True)
    x, yt = synth_patches(B, seed), synth_targets(B, seed)
    m.zero_grad()
    y = m(x)
    loss = torch.nn.BCELoss(reduction='mean')(y, yt)
    loss.backward()
    tag = f'{name}__train'
    out = {tag + '__y': y.detach().numpy(), tag + '__loss': np.array([loss.item()]), tag + '__meta': np.array([B, seed])}
    for k, p in m.named_parameters():
        g = p.grad.numpy().reshape(-1)
        out[tag + '__gnorm__' + k] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum())])
        out[tag + '__gsamp__' + k] = g[grad_sample_index(g.size)].copy()
    for k, v in m.state_dict().items():
        if 'running_' in k:
            out[tag + '__stat__' + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, 'saunet_l_train_golden.npz'), **out)
    print(tag, 'loss', loss.item(), 'tensors', sum(1 for k in out if '__gnorm__' in k))


def state_dict_key_fixture():
    """Names, shapes and dtypes (in order) of the state_dict of every reference model class / size the tests use: the drop-in contract
    `model.load_state_dict(torch.load(pt))` (exp126a...py:388) is a contract on NAMES, so they are pinned as text."""
    import json
    out = {}
    for name in MODEL_SPECS:
        m = build_reference_model(name)
        out[name] = [[k, list(v.shape), str(v.dtype).replace('torch.', '')] for k, v in m.state_dict().items()]
    with open(os.path.join(HERE, 'state_dict_keys.json'), 'w') as f:
        json.dump(out, f, indent=0)
    print('state_dict key lists:', {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    if 'saunet_l_train' in sys.argv[1:]:
        saunet_l_train_golden()
        sys.exit(0)
    if 'keys' in sys.argv[1:]:
        state_dict_key_fixture()
        sys.exit(0)
    if 'realistic' in sys.argv[1:]:
        realistic_goldens()
        state_dict_key_fixture()
        sys.exit(0)
    if 'train' in sys.argv[1:]:
        train_goldens()
        sys.exit(0)
    if 'ext' in sys.argv[1:]:
        ext_goldens()
        sys.exit(0)
    host_goldens()
    nn_goldens()
    train_goldens()
    ext_goldens()
