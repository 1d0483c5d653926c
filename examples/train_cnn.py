#!/usr/bin/env python
"""The training part of an experiment script of the reference (experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py:240-373) on the
B200 path: HCQT / pitch .npy files -> dataset_context objects resident in HBM -> loop.fit.

    python examples/train_cnn.py <dir with hcqt .npy> <dir with pitch .npy> [--epochs 100] [--out model.pt]

Under `torchrun --nproc-per-node N` the ranks train data-parallel (rank-partitioned patches, one NCCL all-reduce per step)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

TRAIN_PARAMS = {'context': 75, 'stride': 50, 'compression': 10, 'aug:transpsemitones': 5, 'aug:randomeq': 20, 'aug:noisestd': 1e-4,
                'aug:tuning': True}                                                    # exp126a...py:36-45
VAL_PARAMS = {'context': 75, 'stride': 50, 'compression': 10}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('dir_hcqt')
    ap.add_argument('dir_pitch')
    ap.add_argument('--val-prefix', nargs='*', default=['2628_', '1933_'])
    ap.add_argument('--epochs', type=int, default=100)
    ap.add_argument('--out', default='model.pt')
    a = ap.parse_args()
    import torch.distributed as dist
    if 'RANK' in os.environ:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group('nccl')
    from multipitch_architectures_b200 import io
    from multipitch_architectures_b200.libdl.data_loaders import dataset_context
    from multipitch_architectures_b200.libdl.nn_models import basic_cnn_segm_sigmoid
    from multipitch_architectures_b200.loop import fit
    train, val = [], []
    for fn in sorted(os.listdir(a.dir_hcqt)):
        if not fn.endswith('.npy'):
            continue
        x = io.load_hcqt_npy(os.path.join(a.dir_hcqt, fn))                             # [6, N, 216] fp32 in HBM
        y = io.load_pitch_npy(os.path.join(a.dir_pitch, fn), min_pitch=24, n_out=72)   # [N, 72]
        is_val = any(fn.startswith(p) for p in a.val_prefix)
        (val if is_val else train).append(dataset_context(x, y, dict(VAL_PARAMS if is_val else TRAIN_PARAMS)))
    model = basic_cnn_segm_sigmoid(n_chan_input=6, n_chan_layers=[20, 20, 10, 1], n_bins_in=216, n_bins_out=72, a_lrelu=0.3, p_dropout=0.2,
                                   precision='bf16').cuda()
    fit(model, train, val or None, batch_size=25, val_batch_size=50, lr=1e-3, weight_decay=0.01, max_epochs=a.epochs, save_path=a.out)


if __name__ == '__main__':
    main()
