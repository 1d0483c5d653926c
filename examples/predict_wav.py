#!/usr/bin/env python
"""Notebook 02 of the reference (02_predict_with_pretrained_model.ipynb) on the B200 path: audio file -> HCQT -> pretrained network ->
[n_frames, 72] pitch activations -> thresholded piano roll.

    python examples/predict_wav.py input.wav [--checkpoint models_pretrained/RETRAIN4_exp128c_..._rerun2.pt] [--out pred.npy]

Without --checkpoint the network keeps its random initialisation (the reference's checkpoints are not redistributed with the repo);
the state_dict layout is the reference's, so `torch.load(<reference .pt>)` loads unchanged."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def predict(path_audio, checkpoint=None, threshold=0.4, precision='fp16', device='cuda'):
    from multipitch_architectures_b200 import io
    from multipitch_architectures_b200.engine import CnnStreamEngine
    from multipitch_architectures_b200.libdl.data_preprocessing.hcqt import get_plan, compute_hopsize_cqt, C1_HZ
    from multipitch_architectures_b200.libdl.nn_models import deep_cnn_segm_sigmoid
    # notebook cell 2-3: model = deep_cnn_segm_sigmoid(n_chan_input=6, n_chan_layers=[40,40,30,10], n_prefilt_layers=5, residual=True, ...)
    model = deep_cnn_segm_sigmoid(n_chan_input=6, n_chan_layers=[40, 40, 30, 10], n_prefilt_layers=5, residual=True, n_bins_in=216,
                                  n_bins_out=72, a_lrelu=0.3, p_dropout=0.2, precision=precision)
    if checkpoint:
        model.load_state_dict(torch.load(checkpoint, map_location='cpu'))
    model = model.to(device).eval()
    # cell 5-6: f_audio, fs = librosa.load(path, sr=22050); f_hcqt, fs_hcqt, hop = compute_efficient_hcqt(f_audio, fs=22050, fmin=C1,
    #           fs_hcqt_target=50, bins_per_octave=36, num_octaves=6, num_harmonics=5, num_subharmonics=1)
    audio, fs = io.load_audio(path_audio, sr=22050, device=device)
    hop, fs_hcqt = compute_hopsize_cqt(50, fs=22050, num_octaves=10)
    plan = get_plan(22050, float(C1_HZ / 2 ** ((3 - 1) / (2 * 36))), hop, 36, 6, 5, 1, device)
    # cell 7: pad 37/38 frames, stride-1 patches of 75 frames, log(1 + 10 x), batches of 50 through the network -> [n_frames, 72]
    with torch.no_grad():
        act, _ = CnnStreamEngine(model).predict_audio(audio, plan)
    return act, (act >= threshold), fs_hcqt


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('audio')
    ap.add_argument('--checkpoint')
    ap.add_argument('--out', default=None)
    ap.add_argument('--threshold', type=float, default=0.4)
    a = ap.parse_args()
    act, roll, fs_hcqt = predict(a.audio, a.checkpoint, a.threshold)
    print(f'{act.shape[0]} frames at {fs_hcqt:.3f} Hz, {int(roll.sum())} active (frame, pitch) cells above {a.threshold}')
    if a.out:
        np.save(a.out, act.cpu().numpy().astype(np.float64))
