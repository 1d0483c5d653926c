"""`-m gpu`: the drop-in claim.  After `install_as_libdl()` the import lines and the call pattern of the reference's experiment script
(experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py: model construction :219-220, checkpoint loading :388, the test loop
:404-436 and the evaluation :439-458) and of notebook 02 (cells 3-6) run as written: host ndarrays, `torch.load` -> `load_state_dict`,
`dataset_context` + `DataLoader(batch 50)`, `model(test_batch)`, `.to('cpu')`, `calculate_eval_measures`.  The results are compared
with the outputs of the UNMODIFIED reference classes on the same checkpoint (tests/golden/realistic_golden.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

from tests import realistic as R

pytestmark = pytest.mark.gpu


@pytest.fixture()
def libdl_alias():
    import multipitch_architectures_b200 as pkg
    from multipitch_architectures_b200.libdl.nn_models import _exec
    saved = {k: v for k, v in sys.modules.items() if k == 'libdl' or k.startswith('libdl.')}
    prev = _exec.DEFAULT_PRECISION
    yield pkg
    _exec.DEFAULT_PRECISION = prev
    for k in [k for k in sys.modules if k == 'libdl' or k.startswith('libdl.')]:
        del sys.modules[k]
    sys.modules.update(saved)


def _reference_test_section(model_cls_name, model_params, path_checkpoint, f_hcqt, annot, device):
    """The body of the reference's test section, import lines included (restated call for call, not copied)."""
    from libdl.data_loaders import dataset_context
    import libdl.nn_models
    from libdl.metrics import calculate_eval_measures
    test_params = {'batch_size': 50, 'shuffle': False, 'num_workers': 0}
    test_dataset_params = {'context': 75, 'stride': 1, 'compression': 10}
    half_context = test_dataset_params['context'] // 2
    num_output_bins, eval_thresh = 72, 0.4
    model = getattr(libdl.nn_models, model_cls_name)(**model_params)
    model.load_state_dict(torch.load(path_checkpoint))
    model.to(device)
    model.eval()
    inputs = np.transpose(f_hcqt, (2, 1, 0))
    targets = annot
    inputs_context = torch.from_numpy(np.pad(inputs, ((0, 0), (half_context, half_context + 1), (0, 0))))
    targets_context = torch.from_numpy(np.pad(targets, ((half_context, half_context + 1), (0, 0))))
    test_set = dataset_context(inputs_context, targets_context, test_dataset_params)
    test_generator = torch.utils.data.DataLoader(test_set, **test_params)
    pred_tot = np.zeros((0, num_output_bins))
    for test_batch, test_labels in test_generator:
        test_batch = test_batch.to(device)
        y_pred = model(test_batch)
        y_pred = y_pred.to('cpu')
        pred = torch.squeeze(torch.squeeze(y_pred, 2), 1).detach().numpy()
        pred_tot = np.append(pred_tot, pred, axis=0)
    measures = calculate_eval_measures(targets, pred_tot, measures=['precision', 'recall', 'f_measure', 'cosine_sim', 'binary_crossentropy',
                                                                    'euclidean_distance', 'binary_accuracy', 'soft_accuracy', 'accum_energy',
                                                                    'roc_auc_measure', 'average_precision_score'],
                                       threshold=eval_thresh, save_roc_plot=False)
    return pred_tot, measures


@pytest.mark.parametrize('precision', ['fp32', 'fp16', 'fp16x3'])
def test_reference_test_section_runs_unmodified_against_the_alias(libdl_alias, tmp_path, precision):
    from oracle import hcqt_oracle as HO
    from tests import synth
    libdl_alias.install_as_libdl(precision=precision)
    import libdl
    import libdl.nn_models
    import libdl.data_preprocessing
    assert libdl.nn_models.deep_cnn_segm_sigmoid.__module__.startswith('multipitch_architectures_b200.')
    # a checkpoint as the reference writes it: torch.save(model.state_dict()) of the reference class == the committed realistic weights
    ckpt = os.path.join(tmp_path, 'RETRAIN4_exp128c_drcnn.pt')
    torch.save(R.state_dict('drcnn'), ckpt)
    f_hcqt, _, _ = HO.compute_efficient_hcqt(synth.synth_clip(**R.CLIP), **R.HCQT_KW)     # float64 [216, N, 6]: the layout of the reference's .npy
    n = 400                                                                               # the first 400 frames (8 DataLoader batches)
    params = dict(n_chan_input=6, n_chan_layers=[40, 40, 30, 10], n_prefilt_layers=5, residual=True, n_bins_in=216, n_bins_out=72, a_lrelu=0.3, p_dropout=0.2)
    pred, meas = _reference_test_section('deep_cnn_segm_sigmoid', params, ckpt, f_hcqt[:, :n + 38], R.labels()[:n + 38], torch.device('cuda:0'))
    ref = R.golden()['drcnn__y']
    # frames whose right context lies inside the excerpt equal the full-clip golden
    err = np.abs(pred[:n] - ref[:n]).max()
    print(f'drop-in loop, default precision {precision}: max|diff| vs reference golden = {err:.2e}')
    assert pred.shape == (n + 38, 72) and err < 1e-3
    lab = R.labels()[:n]
    assert R.prf_counts(lab, pred[:n].astype(np.float32)) == R.prf_counts(lab, ref[:n])
    assert set(meas) >= {'precision', 'recall', 'f_measure', 'average_precision_score'} and 0.8 < meas['f_measure'] <= 1.0


def test_notebook_02_pattern_host_audio_to_pitch_activations(libdl_alias):
    """Notebook 02: audio array -> compute_efficient_hcqt (host float64 [216, N, 6]) -> model on patches -> thresholded piano roll."""
    from tests import synth
    libdl_alias.install_as_libdl(precision='fp16')
    from libdl.data_preprocessing import compute_efficient_hcqt, compute_hopsize_cqt
    from libdl.nn_models import basic_cnn_segm_sigmoid
    from libdl.data_loaders import dataset_context
    f_audio = synth.synth_clip(5, seconds=3.0)
    f_hcqt, fs_hcqt, hop = compute_efficient_hcqt(f_audio, fs=22050, fmin=32.70319566257483, fs_hcqt_target=50, bins_per_octave=36, num_octaves=6,
                                                  num_harmonics=5, num_subharmonics=1, center_bins=True)
    assert isinstance(f_hcqt, np.ndarray) and f_hcqt.dtype == np.float64 and f_hcqt.shape[0] == 216 and f_hcqt.shape[2] == 6
    assert hop == 512 and abs(fs_hcqt - 22050 / 512) < 1e-9 and compute_hopsize_cqt(50, 22050, 10)[0] == 512
    model = basic_cnn_segm_sigmoid(n_chan_input=6, n_chan_layers=[20, 20, 10, 1], n_bins_in=216, n_bins_out=72, a_lrelu=0.3, p_dropout=0.2)
    assert model.precision == 'fp16'
    model.load_state_dict(R.state_dict('cnn_xs'))
    model.to('cuda:0').eval()
    inputs = torch.from_numpy(np.pad(np.transpose(f_hcqt, (2, 1, 0)), ((0, 0), (37, 38), (0, 0))))
    ds = dataset_context(inputs, torch.zeros(inputs.shape[1], 72), {'context': 75, 'stride': 1, 'compression': 10})
    assert len(ds) == f_hcqt.shape[1]
    X = torch.stack([ds[i][0] for i in range(len(ds))])
    with torch.no_grad():
        pred = model(X.to('cuda:0')).to('cpu')
    assert tuple(pred.shape) == (f_hcqt.shape[1], 1, 1, 72) and float(pred.min()) >= 0 and float(pred.max()) <= 1
    assert (pred >= 0.4).float().mean() > 0.005          # the trained CNN:XS finds the synthetic notes
