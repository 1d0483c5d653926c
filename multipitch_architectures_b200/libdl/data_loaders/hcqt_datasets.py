"""`dataset_context` with the reference's interface (/root/reference/libdl/data_loaders/hcqt_datasets.py:10-141):
stride-`stride` patches of `context` frames with the centre-frame target, log compression and the training-time
augmentations ('aug:randomeq', 'aug:noisestd', 'aug:tuning', 'aug:transpsemitones', 'aug:smooth_len').

Index arithmetic is integer-exact with the reference.  For CUDA-resident inputs whole batches are cut — and augmented — by
ONE kernel launch (`batch` / `gather`: mpa_gather_patches_f32, mpa_augment_patches_f32); the reference does this per item
in 16 DataLoader worker processes.  The random DECISIONS of the augmentations (EQ curve incl. the reference's rejection
loop, tuning step, transposition) are drawn on the host from a torch generator, the Gaussian values on the device (Philox),
so the distributions — not the random streams — match the reference.  `__getitem__` keeps the
per-item Dataset protocol on top of the same kernels (host tensors are moved to the GPU on first use; no CPU arithmetic path);
`patch_frames` is the integer index math on its own."""
import itertools

import numpy as np
import torch
import torch.utils.data

from ... import _lib


def eq_offsets(n_chan, bins_per_octave=36):
    """Bin offset of harmonic channel c relative to the fundamental: -36 for the sub-harmonic, int(36*log2(c)) for h = c
    (hcqt_datasets.py:88-92)."""
    return [int(-bins_per_octave)] + [int(bins_per_octave * np.log2(c)) for c in range(1, n_chan)]


def draw_eq(randomeq, n_chan, generator=None, n_bins=216):
    """One (alpha, beta) of the random EQ parabola, redrawn until the curve is non-negative on every harmonic (hcqt_datasets.py:80-95)."""
    f = np.arange(n_bins)
    offs = eq_offsets(n_chan)
    while True:
        alpha = int(torch.randint(1, randomeq + 1, (1,), generator=generator))
        beta = int(torch.randint(0, n_bins, (1,), generator=generator))
        a = np.float32(2e-6) * np.float32(alpha)
        if min(float((np.float32(1) - a * ((f - (beta - o)) ** 2).astype(np.float32)).min()) for o in offs) >= 0:
            return alpha, beta


_DATASET_IDS = itertools.count()


class dataset_context(torch.utils.data.Dataset):
    def __init__(self, inputs, targets, params):
        self.inputs = inputs if isinstance(inputs, torch.Tensor) else torch.as_tensor(inputs)
        self.targets = targets if isinstance(targets, torch.Tensor) else torch.as_tensor(targets)
        self.context = params['context']
        self.stride = params['stride']
        self.compression = params['compression']
        self.targettype = params.get('targettype', 'pitch_class')
        self.transposition = params.get('aug:transpsemitones')
        self.scalingfactor = params.get('aug:scalingfactor')
        self.randomeq = params.get('aug:randomeq')
        self.noisestd = params.get('aug:noisestd')
        self.tuning = params.get('aug:tuning')
        if params.get('aug:smooth_len', 0) > 1:
            # label smoothing along time, one-off host preparation (hcqt_datasets.py:56-60)
            import scipy.signal
            filt = np.expand_dims(scipy.signal.get_window(params['aug:smooth_win'], params['aug:smooth_len'] + 1)[1:], axis=1)
            t = scipy.signal.convolve(self.targets.cpu().numpy(), filt, mode='same')
            t /= np.max(t)
            self.targets = torch.from_numpy(t).to(self.targets.device)
        self.generator = None          # torch.Generator for the augmentation decisions (None = global RNG)
        # Philox stream of the additive / fill noise: every dataset object owns its own key (the reference draws independent noise per
        # item; with one shared seed, call k of file A and call k of file B would produce bit-identical noise fields)
        self._seed = (int(torch.initial_seed()) + 0x9E3779B97F4A7C15 * (1 + next(_DATASET_IDS))) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0
        self._dev = {}

    @property
    def augmenting(self):
        return bool(self.randomeq or self.noisestd or self.tuning or self.transposition)

    def __len__(self):
        return (self.inputs.size()[1] - self.context) // self.stride

    def patch_frames(self, index):
        """Integer index math of the reference's __getitem__ (hcqt_datasets.py:67-75): -> (first frame, last frame + 1, target frame)."""
        half = self.context // 2
        centre = index * self.stride + half
        return centre - half, centre + half + 1, centre

    def __getitem__(self, index):
        """(X [C,context,F], y [1,1,P]) as CUDA tensors.  Host tensors are moved to the GPU on first use: the cut, the log compression and
        the augmentations run in the libmpa kernels only — there is no CPU arithmetic path."""
        if self.scalingfactor:
            assert False, 'Scaling not implemented for dataset_context!'
        if not self.inputs.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.MpaError('dataset_context items are produced by the libmpa kernels: a CUDA (sm_100a) device is required '
                                    '(there is no CPU path)')
            self.inputs, self.targets = self.inputs.cuda(), self.targets.cuda()
        X, y = self.gather([int(index)])
        return X[0], y[0]

    def __getitems__(self, indices):
        """Batched fetch hook of torch.utils.data.DataLoader (one kernel launch per batch instead of one per item): the reference's
        `DataLoader(dataset_context(...), batch_size=50)` loop then costs one cut per batch.  Items are views of one [n,C,T,F] tensor."""
        if self.scalingfactor:
            assert False, 'Scaling not implemented for dataset_context!'
        if not self.inputs.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.MpaError('dataset_context items are produced by the libmpa kernels: a CUDA (sm_100a) device is required '
                                    '(there is no CPU path)')
            self.inputs, self.targets = self.inputs.cuda(), self.targets.cuda()
        X, y = self.gather([int(i) for i in indices])
        return [(X[i], y[i]) for i in range(X.shape[0])]

    # ------------------------------------------------------------------ device path
    def _resident(self):
        if 'inp' not in self._dev:
            dev = self.inputs.device
            self._dev['inp'] = (self.inputs if self.inputs.dtype == torch.float32 else self.inputs.float()).contiguous()
            self._dev['tg'] = self.targets.to(dev).float().contiguous()
            self._dev['offs'] = torch.tensor(eq_offsets(self.inputs.shape[0]), dtype=torch.int32, device=dev)
        return self._dev['inp'], self._dev['tg'], self._dev['offs']

    def draw(self, n):
        """Augmentation decisions for n patches -> dict of int32 host arrays (eq_alpha, eq_beta, tune2, transp); absent = off."""
        g, d = self.generator, {}
        if self.randomeq:
            ab = [draw_eq(int(self.randomeq), self.inputs.shape[0], g, self.inputs.shape[2]) for _ in range(n)]
            d['eq_alpha'] = np.array([a for a, _ in ab], dtype=np.int32)
            d['eq_beta'] = np.array([b for _, b in ab], dtype=np.int32)
        if self.tuning:
            d['tune2'] = torch.randint(-2, 3, (n,), generator=g).numpy().astype(np.int32)
        if self.transposition:
            t = int(self.transposition)
            d['transp'] = torch.randint(-t, t + 1, (n,), generator=g).numpy().astype(np.int32)
        return d

    def gather(self, indices, decisions=None, noise_offset=None, out=None):
        """Patches `indices` (any order) in one launch for CUDA-resident inputs -> (X [n,C,context,F], y [n,1,1,P]).
        `decisions`: output of `draw` (drawn here when None and augmentations are configured); `out` = (X, y) contiguous
        tensors (e.g. slices of a larger batch) to write into."""
        if not self.inputs.is_cuda:
            raise _lib.MpaError('dataset_context.gather needs CUDA-resident inputs')
        if self.scalingfactor:
            assert False, 'Scaling not implemented for dataset_context!'
        inp, tg, offs = self._resident()
        dev = inp.device
        C, NT, F = inp.shape
        idx = torch.as_tensor(indices, dtype=torch.int64).reshape(-1)
        n = idx.numel()
        if n == 0 or int(idx.min()) < 0 or int(idx.max()) >= len(self):
            raise IndexError('dataset_context.gather: patch index out of range')
        if decisions is None:
            decisions = self.draw(n) if self.augmenting else {}
        # ONE host->device copy for the per-patch integers: row 0 = start frames (int64), rows 1-2 = the four int32 decision arrays
        pack = np.zeros((3, n), dtype=np.int64)
        pack[0] = idx.numpy() * self.stride
        dec32 = pack[1:].view(np.int32).reshape(4, n)
        names = ('eq_alpha', 'eq_beta', 'tune2', 'transp')
        for r, k in enumerate(names):
            if k in decisions:
                dec32[r] = np.asarray(decisions[k], dtype=np.int32)
        pack_d = torch.from_numpy(pack).to(dev)
        start = pack_d[0]
        dec_d = pack_d[1:].view(torch.int32).view(4, n)
        dd = {k: dec_d[r] for r, k in enumerate(names) if k in decisions}
        X = out[0] if out is not None else torch.empty(n, C, self.context, F, dtype=torch.float32, device=dev)
        gamma = float(self.compression) if self.compression is not None else 0.0
        if noise_offset is None:
            noise_offset = self._calls
            self._calls += 1
        st = _lib.stream_ptr()
        _lib.call('augment_patches_f32', inp, start, X, C, NT, F, n, self.context, dd.get('eq_alpha'), dd.get('eq_beta'), offs,
                  float(self.noisestd or 0.0), gamma, dd.get('tune2'), dd.get('transp'), 3, 1e-4,
                  _lib.u64(self._seed), _lib.u64(noise_offset), st)
        P = tg.shape[1]
        y = out[1] if out is not None else torch.empty(n, 1, 1, P, dtype=torch.float32, device=dev)
        frame = start + self.context // 2
        _lib.call('augment_targets_f32', tg, frame, dd.get('transp'), y, n, P, st)
        return X, y

    def batch(self, start, n):
        """Patches [start, start+n) in one launch for CUDA-resident inputs -> (X [n,C,context,F], y [n,1,1,P])."""
        if not self.inputs.is_cuda:
            raise _lib.MpaError('dataset_context.batch needs CUDA-resident inputs')
        if self.augmenting:
            return self.gather(range(start, start + n))
        inp, tg, _ = self._resident()
        C, NT, F = inp.shape
        X = torch.empty(n, C, self.context, F, dtype=torch.float32, device=inp.device)
        gamma = float(self.compression) if self.compression is not None else 0.0
        _lib.call('gather_patches_f32', inp, X, C, NT, F, start, n, self.context, self.stride, gamma, _lib.stream_ptr())
        half = self.context // 2
        idx = torch.arange(start, start + n, device=inp.device) * self.stride + half
        y = tg[idx][:, None, None, :]
        return X, y
