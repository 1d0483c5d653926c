"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of the patch-wise network forward of the reference
`libdl.nn_models` classes, written functionally over a `state_dict` so that it
shares no module code with either the reference or the CUDA product.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this file.

Pinned against the reference itself: `tests/golden/make_golden.py` imports
`/root/reference/libdl/nn_models` in the build container, runs the reference
modules on seeded weights/inputs and commits the outputs under `tests/golden/`;
`tests/test_oracle_nn.py` checks this restatement against those fixtures.

Reference sites restated here (all relative to /root/reference):
  * LayerNorm over (chan, bin) per frame ........ libdl/nn_models/basic_cnns.py:371,411
  * CNN / DCNN / DRCNN forward .................. libdl/nn_models/basic_cnns.py:133-195, 342-423
  * double_conv (Conv+BN+ReLU x2, opt. 1x1 res) . libdl/nn_models/unet_cnns.py:30-82
  * bilinear x2 upsample + pad + concat ......... libdl/nn_models/unet_cnns.py:85-104
  * transformer_enc_layer (batch-axis MHA) ...... libdl/nn_models/unet_cnns.py:107-159
  * Unet / SAUnet / SAUSnet / PUnet forward ..... libdl/nn_models/unet_cnns.py:333-407, 496-575, 670-754, 2251-2335
  * blstm_temporal_enc_layer / BLUnet forward ... libdl/nn_models/unet_cnns.py:220-243, 1000-1101
  * BCELoss(mean) with -100 log clamp ........... experiments/Exp1_SectionIV-B/exp126a_musicnet_cnn_basic.py:87
"""
import math

import torch
import torch.nn.functional as F

EPS_LN = 1e-5
EPS_BN = 1e-5


def layernorm_cf(x, w, b):
    """x [B,C,T,F]; normalise over (C,F) for every (b,t); affine w,b [C,F]."""
    mu = x.mean(dim=(1, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(1, 3), keepdim=True)
    xn = (x - mu) / torch.sqrt(var + EPS_LN)
    return xn * w[None, :, None, :] + b[None, :, None, :]


def lrelu(x, a):
    return torch.where(x >= 0, x, a * x)


def maxpool_t(x, k):
    """max over a centred window of k frames (time axis = dim 2), -inf padding."""
    return F.max_pool2d(x, kernel_size=(k, 1), stride=(1, 1), padding=(k // 2, 0))


def prefilt_block(x, w, b, a):
    kh, kw = w.shape[2], w.shape[3]
    y = F.conv2d(x, w, b, padding=(kh // 2, kw // 2))
    return maxpool_t(lrelu(y, a), 3)


def head(x, sd, a):
    """conv2 (3x3 stride (1,3)) -> maxpool13 -> conv3 (75x1) -> 1x1 -> 1xk -> sigmoid."""
    y = F.conv2d(x, sd['conv2.0.weight'], sd['conv2.0.bias'], stride=(1, 3), padding=(1, 0))
    y = maxpool_t(lrelu(y, a), 13)
    y = lrelu(F.conv2d(y, sd['conv3.0.weight'], sd['conv3.0.bias']), a)
    y = lrelu(F.conv2d(y, sd['conv4.0.weight'], sd['conv4.0.bias']), a)
    y = F.conv2d(y, sd['conv4.3.weight'], sd['conv4.3.bias'])
    return torch.sigmoid(y)


def cnn_forward(sd, x, a_lrelu=0.3, residual=False):
    """basic_cnn_segm_sigmoid / deep_cnn_segm_sigmoid (eval mode: dropout = identity)."""
    z = layernorm_cf(x, sd['layernorm.weight'], sd['layernorm.bias'])
    z = prefilt_block(z, sd['conv1.0.weight'], sd['conv1.0.bias'], a_lrelu)
    p = 0
    while f'prefilt_list.{p}.0.weight' in sd:
        z_new = prefilt_block(z, sd[f'prefilt_list.{p}.0.weight'], sd[f'prefilt_list.{p}.0.bias'], a_lrelu)
        z = z_new + z if residual else z_new
        p += 1
    return head(z, sd, a_lrelu)


def batchnorm(x, sd, key, train):
    w, b = sd[key + '.weight'], sd[key + '.bias']
    if train:
        mu = x.mean(dim=(0, 2, 3))
        var = ((x - mu[None, :, None, None]) ** 2).mean(dim=(0, 2, 3))
    else:
        mu, var = sd[key + '.running_mean'], sd[key + '.running_var']
    xn = (x - mu[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + EPS_BN)
    return xn * w[None, :, None, None] + b[None, :, None, None]


def double_conv(x, sd, pre, train=False):
    w0 = sd[pre + '.double_conv.0.weight']
    k = w0.shape[2]
    y = F.conv2d(x, w0, sd[pre + '.double_conv.0.bias'], padding=k // 2)
    y = torch.relu(batchnorm(y, sd, pre + '.double_conv.1', train))
    y = F.conv2d(y, sd[pre + '.double_conv.4.weight'], sd[pre + '.double_conv.4.bias'], padding=k // 2)
    y = torch.relu(batchnorm(y, sd, pre + '.double_conv.5', train))
    if pre + '.resize.weight' in sd:
        y = y + F.conv2d(x, sd[pre + '.resize.weight'], sd[pre + '.resize.bias'])
    return y


def maxpool2x2(x):
    return F.max_pool2d(x, (2, 2))


def upsample2_bilinear_ac(x):
    """x2 bilinear, align_corners=True, written out explicitly."""
    B, C, H, W = x.shape
    Ho, Wo = 2 * H, 2 * W

    def taps(n_in, n_out):
        if n_in == 1:
            z = torch.zeros(n_out, dtype=torch.long)
            return z, z, torch.zeros(n_out)
        pos = torch.arange(n_out, dtype=torch.float32) * ((n_in - 1) / (n_out - 1))
        i0 = torch.clamp(pos.floor().long(), max=n_in - 1)
        i1 = torch.clamp(i0 + 1, max=n_in - 1)
        return i0, i1, (pos - i0.float())

    r0, r1, fr = taps(H, Ho)
    c0, c1, fc = taps(W, Wo)
    top = x[:, :, r0, :] * (1 - fr)[None, None, :, None] + x[:, :, r1, :] * fr[None, None, :, None]
    out = top[:, :, :, c0] * (1 - fc)[None, None, None, :] + top[:, :, :, c1] * fc[None, None, None, :]
    return out


def upconcat(x_low, x_skip):
    up = upsample2_bilinear_ac(x_low)
    dY = x_skip.shape[2] - up.shape[2]
    dX = x_skip.shape[3] - up.shape[3]
    up = F.pad(up, [dX // 2, dX - dX // 2, dY // 2, dY - dY // 2])
    return torch.cat([x_skip, up], dim=1)


def sinusoidal_pe(n, E):
    position = torch.arange(n, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, E, 2, dtype=torch.float32) * (-math.log(10000.0) / E))
    pe = torch.zeros(n, E)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def encoder_layer(x, sd, pre, num_heads, use_pe):
    """transformer_enc_layer: tokens = T*F positions, the *batch* axis is the attention sequence
    (nn.MultiheadAttention without batch_first receives [B, S, E] and reads dim0 as sequence)."""
    B, E, Th, Fw = x.shape
    S = Th * Fw
    t = x.reshape(B, E, S).transpose(1, 2)                  # [B,S,E]
    if use_pe:
        t = t + sinusoidal_pe(S, E)[None]
    q = t @ sd[pre + '.q_linear.weight'].T
    k = t @ sd[pre + '.k_linear.weight'].T
    v = t @ sd[pre + '.v_linear.weight'].T
    Wi, bi = sd[pre + '.attn.in_proj_weight'], sd[pre + '.attn.in_proj_bias']
    q = q @ Wi[:E].T + bi[:E]
    k = k @ Wi[E:2 * E].T + bi[E:2 * E]
    v = v @ Wi[2 * E:].T + bi[2 * E:]
    hd = E // num_heads
    # sequence axis = B, "batch" axis = S
    qh = q.reshape(B, S, num_heads, hd).permute(1, 2, 0, 3)   # [S,H,B,hd]
    kh = k.reshape(B, S, num_heads, hd).permute(1, 2, 0, 3)
    vh = v.reshape(B, S, num_heads, hd).permute(1, 2, 0, 3)
    att = torch.softmax((qh @ kh.transpose(-1, -2)) / math.sqrt(hd), dim=-1)   # [S,H,B,B]
    o = (att @ vh).permute(2, 0, 1, 3).reshape(B, S, E)
    o = o @ sd[pre + '.attn.out_proj.weight'].T + sd[pre + '.attn.out_proj.bias']
    o = o @ sd[pre + '.o_linear.weight'].T
    h1 = F.layer_norm(t + o, (E,), sd[pre + '.layernorm1.weight'], sd[pre + '.layernorm1.bias'], EPS_LN)
    m = torch.relu(h1 @ sd[pre + '.mlp.0.weight'].T + sd[pre + '.mlp.0.bias'])
    m = m @ sd[pre + '.mlp.2.weight'].T + sd[pre + '.mlp.2.bias']
    h2 = F.layer_norm(h1 + m, (E,), sd[pre + '.layernorm2.weight'], sd[pre + '.layernorm2.bias'], EPS_LN)
    return h2.transpose(1, 2).reshape(B, E, Th, Fw)


def blstm_layer(x, sd, pre):
    """blstm_temporal_enc_layer: stacked bidirectional LSTM over time, written out gate by gate (PyTorch order i, f, g, o)."""
    B, C, T, Fw = x.shape
    seq = x.permute(0, 2, 1, 3).reshape(B, T, C * Fw)            # [B,T,(c,f)]
    layer = 0
    while f'{pre}.blstm.weight_ih_l{layer}' in sd:
        outs = []
        for sfx in ('', '_reverse'):
            w_ih, w_hh = sd[f'{pre}.blstm.weight_ih_l{layer}{sfx}'], sd[f'{pre}.blstm.weight_hh_l{layer}{sfx}']
            b = sd[f'{pre}.blstm.bias_ih_l{layer}{sfx}'] + sd[f'{pre}.blstm.bias_hh_l{layer}{sfx}']
            H = w_hh.shape[1]
            h, c = torch.zeros(B, H), torch.zeros(B, H)
            hs = [None] * T
            for t in (range(T) if sfx == '' else range(T - 1, -1, -1)):
                g = seq[:, t] @ w_ih.T + h @ w_hh.T + b
                i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
                c = f * c + i * gg
                h = o * torch.tanh(c)
                hs[t] = h
            outs.append(torch.stack(hs, 1))
        seq = torch.cat(outs, 2)                                 # [B,T,2H]
        layer += 1
    return seq.reshape(B, T, -1, Fw).permute(0, 2, 1, 3).contiguous()


def unet_forward(sd, x, a_lrelu=0.3, train=False, num_heads=8, pos_encoding=None):
    """simple_u_net_largekernels / _doubleselfattn / _polyphony_classif_softmax.
    Returns y_pred, or (y_pred, n_pred) when the state_dict holds the convP head."""
    z = layernorm_cf(x, sd['layernorm.weight'], sd['layernorm.bias'])
    x1 = double_conv(z, sd, 'inc', train)
    x2 = double_conv(maxpool2x2(x1), sd, 'down1.1', train)
    x3 = double_conv(maxpool2x2(x2), sd, 'down2.1', train)
    x4 = double_conv(maxpool2x2(x3), sd, 'down3.1', train)
    x5 = double_conv(maxpool2x2(x4), sd, 'down4.1', train)
    if 'lstm5.blstm.weight_ih_l0' in sd:                         # BLUnet (unet_cnns.py:1083-1086)
        x5 = blstm_layer(x5, sd, 'lstm5')
    if 'lstm4.blstm.weight_ih_l0' in sd:
        x4 = blstm_layer(x4, sd, 'lstm4')
    if 'attention1.q_linear.weight' in sd:
        x5 = encoder_layer(x5, sd, 'attention1', num_heads, pos_encoding == 'sinusoidal')
        x5 = encoder_layer(x5, sd, 'attention2', num_heads, False)
    if 'attention3.q_linear.weight' in sd:                       # SAUSnet (unet_cnns.py:744-745): attention on the lowest skip
        x4 = encoder_layer(x4, sd, 'attention3', num_heads, pos_encoding == 'sinusoidal')
        x4 = encoder_layer(x4, sd, 'attention4', num_heads, False)
    u = double_conv(upconcat(x5, x4), sd, 'upconv1', train)
    u = double_conv(upconcat(u, x3), sd, 'upconv2', train)
    u = double_conv(upconcat(u, x2), sd, 'upconv3', train)
    u = double_conv(upconcat(u, x1), sd, 'upconv4', train)
    y = head(u, sd, a_lrelu)
    if 'convP.0.weight' in sd:
        p = lrelu(F.conv2d(x5, sd['convP.0.weight'], sd['convP.0.bias']), a_lrelu)
        p = F.max_pool2d(p, kernel_size=(2, 5), stride=(1, 2))
        p = F.conv2d(p, sd['convP.4.weight'], sd['convP.4.bias'])
        return y, p
    return y


def bce_mean(y_pred, y):
    """nn.BCELoss(reduction='mean'): each log term clamped at -100."""
    lp = torch.clamp(torch.log(y_pred), min=-100.0)
    l1p = torch.clamp(torch.log(1 - y_pred), min=-100.0)
    return -(y * lp + (1 - y) * l1p).mean()


def round_bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def round_tf32(x):
    """round-to-nearest-even to 10 explicit mantissa bits (what cvt.rna.tf32.f32 produces, ties aside)."""
    i = x.contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)
