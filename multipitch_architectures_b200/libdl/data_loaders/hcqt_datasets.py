"""`dataset_context` with the reference's interface (/root/reference/libdl/data_loaders/hcqt_datasets.py:10-141):
stride-`stride` patches of `context` frames with the centre-frame target and log compression.

Index arithmetic is integer-exact with the reference.  For CUDA-resident inputs whole batches are cut by the
mpa_gather_patches_f32 kernel (`batch`); `__getitem__` keeps the per-item Dataset protocol for host tensors
(what the reference's DataLoader workers do).  The training-time augmentations ('aug:*') are a later row of the
scope table and are rejected explicitly."""
import numpy as np
import torch
import torch.utils.data

from ... import _lib


class dataset_context(torch.utils.data.Dataset):
    def __init__(self, inputs, targets, params):
        unsupported = [k for k in params if k.startswith('aug:')]
        if unsupported:
            raise NotImplementedError(f'augmentations {unsupported} are not part of this hot-path implementation yet')
        self.inputs = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(inputs)
        self.targets = targets if isinstance(targets, torch.Tensor) else torch.as_tensor(targets)
        self.context = params['context']
        self.stride = params['stride']
        self.compression = params['compression']
        self.targettype = params.get('targettype', 'pitch_class')

    def __len__(self):
        return (self.inputs.size()[1] - self.context) // self.stride

    def __getitem__(self, index):
        half = self.context // 2
        centre = index * self.stride + half
        X = self.inputs[:, centre - half:centre + half + 1, :].type(torch.FloatTensor)
        y = self.targets[centre, :].type(torch.FloatTensor)[None, None, :]
        if self.compression is not None:
            X = torch.log(1 + self.compression * X)
        return X, y

    def batch(self, start, n):
        """Patches [start, start+n) in one launch for CUDA-resident inputs -> (X [n,C,context,F], y [n,1,1,P])."""
        if not self.inputs.is_cuda:
            raise _lib.MpaError('dataset_context.batch needs CUDA-resident inputs')
        C, NT, F = self.inputs.shape
        inp = self.inputs if self.inputs.dtype == torch.float32 else self.inputs.float()
        X = torch.empty(n, C, self.context, F, dtype=torch.float32, device=inp.device)
        gamma = float(self.compression) if self.compression is not None else 0.0
        _lib.call('gather_patches_f32', inp.contiguous(), X, C, NT, F, start, n, self.context, self.stride, gamma, _lib.stream_ptr())
        half = self.context // 2
        idx = torch.arange(start, start + n, device=inp.device) * self.stride + half
        y = self.targets.to(inp.device)[idx].float()[:, None, None, :]
        return X, y
