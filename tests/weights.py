"""Deterministic, torch-RNG-independent weight / input generators shared by the golden generator
(tests/golden/make_golden.py, which fills the REFERENCE modules) and by the parity tests (which fill
the product modules and the oracle).  numpy PCG64 streams only."""
import numpy as np
import torch


def fill_state_dict(sd, seed, scheme='adversarial'):
    """Return a new state_dict with the same keys/shapes/dtypes as `sd`, filled from default_rng(seed) in key order.

    scheme 'adversarial' (default): conv/linear weights ~ U(+-1.4*sqrt(3/fan_in)) (2.4x the PyTorch default gain: logits reach
    +-18, outputs span 0..1), norm scales ~ U(.5,1.5), small biases — a deliberately harsh numerical test.
    scheme 'torch_default': the distribution of the reference constructors' own initialisation ("reference
    constructor default init"): weights and biases ~ U(+-1/sqrt(fan_in)), norm weight 1 / bias 0, running stats 0 / 1."""
    if scheme == 'torch_default':
        return _fill_torch_default(sd, seed)
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in sd.items():
        shape = tuple(v.shape)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.tensor(7, dtype=v.dtype)
            continue
        if k.endswith('running_var'):
            a = rng.uniform(0.5, 1.5, size=shape)
        elif k.endswith('running_mean'):
            a = rng.normal(0.0, 0.1, size=shape)
        elif 'layernorm' in k and k.endswith('weight'):
            a = 1.0 + 0.1 * rng.standard_normal(size=shape)
        elif 'layernorm' in k and k.endswith('bias'):
            a = 0.1 * rng.standard_normal(size=shape)
        elif k.endswith('weight') and len(shape) == 1:          # BatchNorm scale
            a = rng.uniform(0.5, 1.5, size=shape)
        elif k.endswith('weight'):
            fan_in = int(np.prod(shape[1:]))
            a = rng.uniform(-1.0, 1.0, size=shape) * (1.4 * np.sqrt(3.0 / fan_in))
        else:                                                   # biases
            a = rng.uniform(-0.1, 0.1, size=shape)
        out[k] = torch.from_numpy(np.asarray(a, dtype=np.float32)).to(v.dtype)
    return out


def _fill_torch_default(sd, seed):
    rng = np.random.default_rng(seed)
    out, last_fan_in = {}, 1
    for k, v in sd.items():
        shape = tuple(v.shape)
        if k.endswith('num_batches_tracked'):
            out[k] = torch.tensor(0, dtype=v.dtype)
            continue
        if k.endswith('running_var'):
            a = np.ones(shape)
        elif k.endswith('running_mean'):
            a = np.zeros(shape)
        elif ('layernorm' in k) or (k.endswith('weight') and len(shape) == 1) or (k.endswith('bias') and _is_norm_bias(k, sd)):
            a = np.ones(shape) if k.endswith('weight') else np.zeros(shape)
        elif k.endswith('weight'):
            last_fan_in = int(np.prod(shape[1:]))
            a = rng.uniform(-1.0, 1.0, size=shape) / np.sqrt(last_fan_in)
        else:
            a = rng.uniform(-1.0, 1.0, size=shape) / np.sqrt(last_fan_in)
        out[k] = torch.from_numpy(np.asarray(a, dtype=np.float32)).to(v.dtype)
    return out


def _is_norm_bias(k, sd):
    w = k[:-4] + 'weight'
    return w in sd and sd[w].dim() == 1 and (k[:-4] + 'running_mean') in sd


def synth_patches(B, seed, T=75, C=6, F=216):
    """log1p(10*|N(0,.05^2)|) patches with a little structure so rows differ."""
    rng = np.random.default_rng(seed)
    x = np.abs(rng.normal(0.0, 0.05, size=(B, C, T, F)))
    x *= (1.0 + np.sin(np.arange(F) / 9.0)[None, None, None, :] ** 2) * (0.5 + rng.uniform(size=(B, 1, T, 1)))
    return torch.from_numpy(np.log1p(10.0 * x).astype(np.float32))


def synth_targets(B, seed, P=72):
    rng = np.random.default_rng(seed + 1000)
    return torch.from_numpy((rng.uniform(size=(B, 1, 1, P)) < 0.04).astype(np.float32))


# The five BASELINE configs and reduced variants for fast tests.
MODEL_SPECS = {
    'cnn_xs':     dict(cls='basic_cnn_segm_sigmoid', kw=dict(n_chan_input=6, n_chan_layers=[20, 20, 10, 1], n_bins_in=216, n_bins_out=72)),
    'drcnn':      dict(cls='deep_cnn_segm_sigmoid', kw=dict(n_chan_input=6, n_chan_layers=[40, 40, 30, 10], n_prefilt_layers=5, residual=True, n_bins_in=216, n_bins_out=72)),
    'drcnn_tiny': dict(cls='deep_cnn_segm_sigmoid', kw=dict(n_chan_input=6, n_chan_layers=[8, 8, 6, 4], n_prefilt_layers=3, residual=True, n_bins_in=216, n_bins_out=72)),
    'dcnn_tiny':  dict(cls='deep_cnn_segm_sigmoid', kw=dict(n_chan_input=6, n_chan_layers=[8, 8, 6, 4], n_prefilt_layers=2, residual=False, n_bins_in=216, n_bins_out=72)),
    'unet_m':     dict(cls='simple_u_net_largekernels', kw=dict(n_chan_input=6, n_chan_layers=[128, 100, 80, 50], n_bins_in=216, n_bins_out=72, scalefac=8)),
    'unet_tiny':  dict(cls='simple_u_net_largekernels', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=16)),
    'punet':      dict(cls='simple_u_net_polyphony_classif_softmax', kw=dict(n_chan_input=6, n_chan_layers=[128, 180, 150, 100], n_bins_in=216, n_bins_out=72, scalefac=2, num_polyphony_steps=24)),
    'punet_tiny': dict(cls='simple_u_net_polyphony_classif_softmax', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=16, num_polyphony_steps=24)),
    'saunet_l':   dict(cls='simple_u_net_doubleselfattn', kw=dict(n_chan_input=6, n_chan_layers=[128, 80, 50, 30], n_bins_in=216, n_bins_out=72, scalefac=4, embed_dim=128, num_heads=8, mlp_dim=8192, pos_encoding='sinusoidal')),
    'sausnet_tiny': dict(cls='simple_u_net_doubleselfattn_twolayers', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=16, embed_dim=32, num_heads=8, mlp_dim=64, pos_encoding='sinusoidal')),
    'sausnet_s8': dict(cls='simple_u_net_doubleselfattn_twolayers', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=8, embed_dim=64, num_heads=8, mlp_dim=128, pos_encoding='sinusoidal')),
    'blunet_tiny': dict(cls='u_net_blstm_varlayers', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=16, embed_dim=416, hidden_size=208, lstm_depth=1, lstm_number=2)),
    'blunet_s32': dict(cls='u_net_blstm_varlayers', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=32, embed_dim=208, hidden_size=104, lstm_depth=1, lstm_number=2)),
    'blunet_d':   dict(cls='u_net_blstm_varlayers', kw=dict(n_chan_input=6, n_chan_layers=[128, 80, 50, 30], n_bins_in=216, n_bins_out=72, scalefac=8, embed_dim=832, hidden_size=416, lstm_depth=1, lstm_number=2)),
    'saunet_tiny': dict(cls='simple_u_net_doubleselfattn', kw=dict(n_chan_input=6, n_chan_layers=[16, 10, 8, 5], n_bins_in=216, n_bins_out=72, scalefac=16, embed_dim=32, num_heads=8, mlp_dim=64, pos_encoding='sinusoidal')),
}
