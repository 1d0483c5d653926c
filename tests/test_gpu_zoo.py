"""`-m gpu`: every model size the reference's experiment scripts instantiate (Exp1 / Exp2: CNN:XS..L, DCNN, DRCNN, Unet:S..XL,
SAUnet, SAUSnet, PUnet, BLUnet — SURVEY.md 8f row 2) runs through the product classes and agrees with the oracle on
default-initialised weights: fp32 path <= 1e-3 (the north-star gate), fp16 tensor-core path (or its CUDA-core stand-in when a layer is
wider than the tensor-core kernels take) <= 1e-3 as well."""
import numpy as np
import pytest
import torch

from oracle import nn_oracle as NO
from tests.weights import fill_state_dict, synth_patches

pytestmark = pytest.mark.gpu

COMMON = dict(n_chan_input=6, n_bins_in=216, n_bins_out=72)
SA = dict(num_heads=8, pos_encoding='sinusoidal')
ZOO = [  # (script, class, kwargs)
    ('exp126b', 'basic_cnn_segm_sigmoid', dict(n_chan_layers=[100, 100, 50, 10])),
    ('exp126c', 'basic_cnn_segm_sigmoid', dict(n_chan_layers=[250, 150, 100, 100])),
    ('exp126d', 'basic_cnn_segm_sigmoid', dict(n_chan_layers=[280, 180, 120, 100])),
    ('exp127a', 'deep_cnn_segm_sigmoid', dict(n_chan_layers=[20, 20, 10, 1], n_prefilt_layers=5)),
    ('exp127c', 'deep_cnn_segm_sigmoid', dict(n_chan_layers=[70, 70, 50, 10], n_prefilt_layers=5)),
    ('exp128a', 'deep_cnn_segm_sigmoid', dict(n_chan_layers=[20, 20, 10, 1], n_prefilt_layers=5, residual=True)),
    ('exp128c', 'deep_cnn_segm_sigmoid', dict(n_chan_layers=[70, 70, 50, 10], n_prefilt_layers=5, residual=True)),
    ('exp160d2', 'simple_u_net_largekernels', dict(n_chan_layers=[64, 30, 20, 10], scalefac=8)),
    ('exp160e3', 'simple_u_net_largekernels', dict(n_chan_layers=[128, 150, 100, 80], scalefac=4)),
    ('exp160f', 'simple_u_net_largekernels', dict(n_chan_layers=[128, 180, 150, 100], scalefac=2)),
    ('exp180b', 'simple_u_net_doubleselfattn', dict(n_chan_layers=[64, 30, 20, 10], scalefac=8, embed_dim=64, mlp_dim=1024, **SA)),
    ('exp180e', 'simple_u_net_doubleselfattn', dict(n_chan_layers=[128, 200, 150, 150], scalefac=2, embed_dim=256, mlp_dim=8192, **SA)),
    ('exp180f', 'simple_u_net_doubleselfattn', dict(n_chan_layers=[128, 200, 150, 150], scalefac=4, embed_dim=128, mlp_dim=8192, **SA)),
    ('exp181b', 'simple_u_net_doubleselfattn_twolayers', dict(n_chan_layers=[64, 30, 20, 10], scalefac=8, embed_dim=64, mlp_dim=512, **SA)),
    ('exp181d', 'simple_u_net_doubleselfattn_twolayers', dict(n_chan_layers=[128, 80, 50, 30], scalefac=4, embed_dim=128, mlp_dim=4096, **SA)),
    ('exp195e3', 'simple_u_net_polyphony_classif_softmax', dict(n_chan_layers=[128, 150, 100, 80], scalefac=4, num_polyphony_steps=24)),
    ('exp195g', 'simple_u_net_polyphony_classif_softmax', dict(n_chan_layers=[128, 100, 80, 50], scalefac=8, num_polyphony_steps=24)),
    ('exp186b', 'u_net_blstm_varlayers', dict(n_chan_layers=[64, 30, 20, 10], scalefac=16, embed_dim=416, hidden_size=208, lstm_depth=1, lstm_number=1)),
    ('exp186e', 'u_net_blstm_varlayers', dict(n_chan_layers=[128, 200, 150, 150], scalefac=4, embed_dim=1664, hidden_size=832, lstm_depth=1, lstm_number=1)),
]


@pytest.mark.parametrize('script,cls,kw', ZOO, ids=[z[0] for z in ZOO])
def test_paper_model_sizes_match_oracle(script, cls, kw):
    from multipitch_architectures_b200.libdl import nn_models as M
    B, seed = 2, 300 + len(script)
    x = synth_patches(B, seed)
    ref = None
    for prec in ('fp32', 'fp16'):
        m = getattr(M, cls)(**COMMON, **kw, precision=prec)
        sd = fill_state_dict(m.state_dict(), seed, scheme='torch_default')
        m.load_state_dict(sd)
        if ref is None:
            torch.set_num_threads(max(1, torch.get_num_threads()))
            with torch.no_grad():
                if 'u_net' in cls:
                    ref = NO.unet_forward(sd, x, pos_encoding=kw.get('pos_encoding'))
                else:
                    ref = NO.cnn_forward(sd, x, residual=kw.get('residual', False))
            ref = [r.numpy() for r in (ref if isinstance(ref, tuple) else (ref,))]
        m = m.cuda().eval()
        with torch.no_grad():
            y = m(x.cuda())
        y = [t.cpu().numpy() for t in (y if isinstance(y, tuple) else (y,))]
        assert len(y) == len(ref)
        errs = [float(np.abs(a - b).max()) for a, b in zip(y, ref)]
        print(f'{script} {cls} {prec}: max|diff| = {errs}')
        assert y[0].shape == (B, 1, 1, 72) and errs[0] < 1e-3
        if len(errs) > 1:
            assert errs[1] < (1e-3 if prec == 'fp32' else 2e-2)        # DoP logits (not probabilities)
