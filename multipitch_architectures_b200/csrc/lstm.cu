// Bidirectional LSTM layer over the time axis of a U-Net level: the BLUnet's bottleneck (reference:
// libdl/nn_models/unet_cnns.py:220-243 `blstm_temporal_enc_layer`, used by `u_net_blstm_varlayers`, :1000-1101, experiments exp186b/d/e).
// The sequence is SHORT (4 frames at the bottleneck) and the layer is < 1 % of the model's FLOPs, so the formulation is
//   1. sequence layout change  NCHW [B,C,T,F] -> [B,T,C*F]                                       (lstm_seq_from_nchw_kernel)
//   2. input projections of all steps and both directions as ONE GEMM-shaped launch              (lstm_pregate_kernel)
//   3. per step, both directions in one launch: h W_hh^T for a tile of hidden units x all four gates, then the cell update
//      (sigmoid/tanh, c and h) in the same kernel — gates never reach HBM                        (lstm_step_kernel)
// fp32 CUDA cores throughout (PyTorch gate order i, f, g, o; both bias vectors added).
#include "common.cuh"

namespace mpa {

__global__ void lstm_seq_from_nchw_kernel(const float* __restrict__ x, float* __restrict__ seq, long long total, int C, int T, int F) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int c = (int)((i / F) % C);
    const int t = (int)((i / ((long long)F * C)) % T);
    const long long b = i / ((long long)F * C * T);
    seq[i] = x[((b * C + c) * T + t) * F + f];
  }
}

__global__ void lstm_seq_to_nchw_kernel(const float* __restrict__ seq, float* __restrict__ x, long long total, int C, int T, int F) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i % F);
    const int t = (int)((i / F) % T);
    const int c = (int)((i / ((long long)F * T)) % C);
    const long long b = i / ((long long)F * T * C);
    x[i] = seq[((b * T + t) * C + c) * F + f];
  }
}

// G[d][m][n] = sum_k x[m][k] * w_ih[d][n][k] + b_ih[d][n] + b_hh[d][n];  m = b*T+t, n in [0, 4H).  64 x 64 tiles, K step 16.
__global__ void __launch_bounds__(256) lstm_pregate_kernel(const float* __restrict__ x, const float* __restrict__ w_ih, const float* __restrict__ b_ih,
                                                           const float* __restrict__ b_hh, float* __restrict__ G, int M, int N, int K) {
  __shared__ float As[16][65], Ws[16][65];
  const int d = blockIdx.z, m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const float* W = w_ih + (size_t)d * N * K;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int r = i >> 4, kk = i & 15;
      As[kk][r] = (m0 + r < M && k0 + kk < K) ? x[(size_t)(m0 + r) * K + k0 + kk] : 0.f;
      Ws[kk][r] = (n0 + r < N && k0 + kk < K) ? W[(size_t)(n0 + r) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[kk][ty * 4 + i];
        w[i] = Ws[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) G[((size_t)d * M + m) * N + n] = acc[i][j] + b_ih[(size_t)d * N + n] + b_hh[(size_t)d * N + n];
    }
}

constexpr int kLstmUnits = 8, kLstmBatch = 32;

// one time step of both directions: direction d handles t = step (d = 0) or T-1-step (d = 1).
// CTA = kLstmUnits hidden units x 4 gates (32 weight rows) x kLstmBatch sequences; h_prev is read from `out` (the previous step's slot).
// c_all != nullptr (training): the activated gates overwrite their pre-activations in G and every cell state is kept in
// c_all [D][B][T][H] — what the backward pass needs.
__global__ void __launch_bounds__(256) lstm_step_kernel(float* __restrict__ G, const float* __restrict__ w_hh, float* __restrict__ cstate,
                                                        float* __restrict__ out, float* __restrict__ c_all, int B, int T, int H, int D, int step) {
  __shared__ float Wt[32][33], Ht[kLstmBatch][33], gates[32][kLstmBatch + 1];
  const int d = blockIdx.y, j0 = blockIdx.x * kLstmUnits, b0 = blockIdx.z * kLstmBatch;
  const int t = d == 0 ? step : T - 1 - step;
  const int tprev = d == 0 ? t - 1 : t + 1;
  const int r = threadIdx.x >> 3, bq = threadIdx.x & 7;      // weight row 0..31 (gate = r / 8, unit = r % 8), 4 sequences per thread
  const int grow = (r / kLstmUnits) * H + j0 + (r % kLstmUnits);
  const bool row_ok = j0 + (r % kLstmUnits) < H;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (step > 0) {
    const float* W = w_hh + (size_t)d * 4 * H * H;
    for (int k0 = 0; k0 < H; k0 += 32) {
      for (int i = threadIdx.x; i < 32 * 32; i += 256) {
        const int rr = i >> 5, kk = i & 31;
        const int gr = (rr / kLstmUnits) * H + j0 + (rr % kLstmUnits);
        Wt[rr][kk] = (j0 + (rr % kLstmUnits) < H && k0 + kk < H) ? W[(size_t)gr * H + k0 + kk] : 0.f;
        const int b = b0 + rr;
        Ht[rr][kk] = (b < B && k0 + kk < H) ? out[((size_t)b * T + tprev) * D * H + (size_t)d * H + k0 + kk] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) {
        const float w = Wt[r][kk];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, Ht[bq * 4 + i][kk], acc[i]);
      }
      __syncthreads();
    }
  }
  const int M = B * T, N = 4 * H;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + bq * 4 + i;
    gates[r][bq * 4 + i] = (row_ok && b < B) ? acc[i] + G[((size_t)d * M + (size_t)b * T + t) * N + grow] : 0.f;
  }
  __syncthreads();
  {
    const int u = threadIdx.x >> 5, bl = threadIdx.x & 31;     // 8 units x 32 sequences
    const int j = j0 + u, b = b0 + bl;
    if (j < H && b < B) {
      const float gi = 1.f / (1.f + expf(-gates[u][bl]));
      const float gf = 1.f / (1.f + expf(-gates[kLstmUnits + u][bl]));
      const float gg = tanhf(gates[2 * kLstmUnits + u][bl]);
      const float go = 1.f / (1.f + expf(-gates[3 * kLstmUnits + u][bl]));
      float* cp = cstate + ((size_t)d * B + b) * H + j;
      const float c = (step > 0 ? gf * cp[0] : 0.f) + gi * gg;
      cp[0] = c;
      out[((size_t)b * T + t) * D * H + (size_t)d * H + j] = go * tanhf(c);
      if (c_all) {
        c_all[(((size_t)d * B + b) * T + t) * H + j] = c;
        float* gs = G + ((size_t)d * M + (size_t)b * T + t) * N + j;
        gs[0] = gi; gs[H] = gf; gs[2 * (size_t)H] = gg; gs[3 * (size_t)H] = go;
      }
    }
  }
}

// One backward time step of both directions (run in the reverse of the forward order).  Thread per (d, b, j):
//   dh = g_out[b][t][d*H+j] + sum_n dgate[d][(b, t_later)][n] * w_hh[d][n][j]     (the recurrent term, absent at the first backward step)
//   dc = dc_carry + dh * o * (1 - tanh(c)^2);   di, df, dg, do = ... * sigma' / tanh';   dc_carry = dc * f
// gates [D][B*T][4H] holds the activated gates and is overwritten in place by the pre-activation gradients.
__global__ void __launch_bounds__(128) lstm_step_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_all, const float* __restrict__ w_hh,
                                                            const float* __restrict__ g_out, float* __restrict__ dc_carry, int B, int T, int H, int D,
                                                            int bstep) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y, d = blockIdx.z;
  if (j >= H) return;
  const int fstep = T - 1 - bstep;                       // forward step index being undone
  const int t = d == 0 ? fstep : T - 1 - fstep;
  const int t_later = d == 0 ? t + 1 : t - 1;            // the time index processed one forward step later
  const int t_prev = d == 0 ? t - 1 : t + 1;
  const size_t M = (size_t)B * T, N = 4 * (size_t)H;
  float dh = g_out[((size_t)b * T + t) * D * H + (size_t)d * H + j];
  if (bstep > 0) {
    const float* dg_later = gates + ((size_t)d * M + (size_t)b * T + t_later) * N;
    const float* W = w_hh + (size_t)d * N * H + j;
    float acc = 0.f;
    for (int n = 0; n < (int)N; ++n) acc = fmaf(dg_later[n], W[(size_t)n * H], acc);
    dh += acc;
  }
  float* gs = gates + ((size_t)d * M + (size_t)b * T + t) * N + j;
  const float gi = gs[0], gf = gs[H], gg = gs[2 * (size_t)H], go = gs[3 * (size_t)H];
  const float c = c_all[(((size_t)d * B + b) * T + t) * H + j];
  const float c_prev = fstep > 0 ? c_all[(((size_t)d * B + b) * T + t_prev) * H + j] : 0.f;
  const float tc = tanhf(c);
  float* dcp = dc_carry + ((size_t)d * B + b) * H + j;
  const float dc = (bstep > 0 ? dcp[0] : 0.f) + dh * go * (1.f - tc * tc);
  dcp[0] = dc * gf;
  // (no cross-thread hazard on gs: each thread owns its four entries; the dg_later rows belong to another time index)
  gs[0] = dc * gg * gi * (1.f - gi);
  gs[H] = dc * c_prev * gf * (1.f - gf);
  gs[2 * (size_t)H] = dc * gi * (1.f - gg * gg);
  gs[3 * (size_t)H] = dh * tc * go * (1.f - go);
}

// hprev[d][(b,t)][k] = h of the previous forward step of direction d (zero at the first one), from out [B][T][D*H]
__global__ void lstm_hprev_kernel(const float* __restrict__ out, float* __restrict__ hprev, long long total, int B, int T, int H, int D) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % H);
    const int t = (int)((i / H) % T);
    const int b = (int)((i / ((long long)H * T)) % B);
    const int d = (int)(i / ((long long)H * T * B));
    const int tp = d == 0 ? t - 1 : t + 1;
    hprev[i] = (tp >= 0 && tp < T) ? out[((size_t)b * T + tp) * D * H + (size_t)d * H + k] : 0.f;
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_lstm_seq_from_nchw_f32(const float* x, float* seq, int B, int C, int T, int F, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && seq && B > 0 && C > 0 && T > 0 && F > 0, "lstm_seq_from_nchw: bad argument");
  const long long total = (long long)B * C * T * F;
  lstm_seq_from_nchw_kernel<<<ceil_div(total, 256) > 2368 ? 2368 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, seq, total, C, T, F);
  MPA_CHECK_LAUNCH("lstm_seq_from_nchw");
  return MPA_OK;
}

int mpa_lstm_seq_to_nchw_f32(const float* seq, float* x, int B, int C, int T, int F, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && seq && B > 0 && C > 0 && T > 0 && F > 0, "lstm_seq_to_nchw: bad argument");
  const long long total = (long long)B * C * T * F;
  lstm_seq_to_nchw_kernel<<<ceil_div(total, 256) > 2368 ? 2368 : ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(seq, x, total, C, T, F);
  MPA_CHECK_LAUNCH("lstm_seq_to_nchw");
  return MPA_OK;
}

size_t mpa_lstm_layer_workspace(int B, int T, int H, int D) {
  return ((size_t)D * B * T * 4 * H + (size_t)D * B * H) * sizeof(float);
}

static int lstm_forward(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* out, int B, int T,
                        int I, int H, int D, float* G, float* cstate, float* c_all, cudaStream_t st) {
  const int M = B * T, N = 4 * H;
  lstm_pregate_kernel<<<dim3(ceil_div(N, 64), ceil_div(M, 64), D), 256, 0, st>>>(x, w_ih, b_ih, b_hh, G, M, N, I);
  MPA_CHECK_LAUNCH("lstm_pregate");
  for (int s = 0; s < T; ++s) {
    lstm_step_kernel<<<dim3(ceil_div(H, kLstmUnits), D, ceil_div(B, kLstmBatch)), 256, 0, st>>>(G, w_hh, cstate, out, c_all, B, T, H, D, s);
    MPA_CHECK_LAUNCH("lstm_step");
  }
  return MPA_OK;
}

/* training forward: `gates` [D][B*T][4H] and `c_all` [D][B][T][H] are kept for mpa_lstm_layer_bwd_f32; workspace >= D*B*H floats */
int mpa_lstm_layer_train_f32(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* out, float* gates,
                             float* c_all, int B, int T, int I, int H, int D, void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && out && gates && c_all && workspace, "lstm_layer_train: null argument");
  MPA_REQUIRE(B > 0 && T > 0 && I > 0 && H > 0 && (D == 1 || D == 2), "lstm_layer_train: bad shape");
  MPA_REQUIRE(ws_bytes >= (size_t)D * B * H * sizeof(float), "lstm_layer_train: workspace too small");
  return lstm_forward(x, w_ih, w_hh, b_ih, b_hh, out, B, T, I, H, D, gates, (float*)workspace, c_all, (cudaStream_t)stream);
}

size_t mpa_lstm_layer_bwd_workspace(int B, int T, int H, int D) { return ((size_t)D * B * H + (size_t)D * B * T * H) * sizeof(float); }

/* backward of one layer.  gates / c_all / out from mpa_lstm_layer_train_f32 (gates is consumed: it holds the pre-activation gradients
 * afterwards); g_out [B][T][D*H].  Outputs (overwritten): g_x [B][T][I] (may be NULL), g_w_ih [D][4H][I], g_w_hh [D][4H][H], g_b [D][4H]
 * (= the gradient of both bias vectors). */
int mpa_lstm_layer_bwd_f32(const float* x, const float* w_ih, const float* w_hh, float* gates, const float* c_all, const float* out,
                           const float* g_out, float* g_x, float* g_w_ih, float* g_w_hh, float* g_b, int B, int T, int I, int H, int D,
                           void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_ih && w_hh && gates && c_all && out && g_out && g_w_ih && g_w_hh && g_b && workspace, "lstm_layer_bwd: null argument");
  MPA_REQUIRE(B > 0 && T > 0 && I > 0 && H > 0 && (D == 1 || D == 2), "lstm_layer_bwd: bad shape");
  if (ws_bytes < mpa_lstm_layer_bwd_workspace(B, T, H, D)) {
    set_error("lstm_layer_bwd: workspace %zu < %zu bytes", ws_bytes, mpa_lstm_layer_bwd_workspace(B, T, H, D));
    return MPA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* dc_carry = (float*)workspace;
  float* hprev = dc_carry + (size_t)D * B * H;
  const int M = B * T, N = 4 * H;
  for (int s = 0; s < T; ++s) {
    lstm_step_bwd_kernel<<<dim3(ceil_div(H, 128), B, D), 128, 0, st>>>(gates, c_all, w_hh, g_out, dc_carry, B, T, H, D, s);
    MPA_CHECK_LAUNCH("lstm_step_bwd");
  }
  const long long total = (long long)D * M * H;
  lstm_hprev_kernel<<<ceil_div(total, 256) > 2368 ? 2368 : ceil_div(total, 256), 256, 0, st>>>(out, hprev, total, B, T, H, D);
  MPA_CHECK_LAUNCH("lstm_hprev");
  for (int d = 0; d < D; ++d) {
    const float* dg = gates + (size_t)d * M * N;
    int rc = mpa_gemm_f32(dg, x, g_w_ih + (size_t)d * N * I, N, I, M, 1, 0, stream);                       // dW_ih = dgates^T x
    if (rc != MPA_OK) return rc;
    rc = mpa_gemm_f32(dg, hprev + (size_t)d * M * H, g_w_hh + (size_t)d * N * H, N, H, M, 1, 0, stream);   // dW_hh = dgates^T h_prev
    if (rc != MPA_OK) return rc;
    rc = mpa_colsum_f32(dg, g_b + (size_t)d * N, M, N, stream);
    if (rc != MPA_OK) return rc;
    if (g_x) {
      rc = mpa_gemm_f32(dg, w_ih + (size_t)d * N * I, g_x, M, I, N, 0, d > 0 ? 1 : 0, stream);             // dx (+)= dgates W_ih
      if (rc != MPA_OK) return rc;
    }
  }
  return MPA_OK;
}

int mpa_lstm_layer_f32(const float* x, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, float* out, int B, int T,
                       int I, int H, int D, void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_ih && w_hh && b_ih && b_hh && out && workspace, "lstm_layer: null argument");
  MPA_REQUIRE(B > 0 && T > 0 && I > 0 && H > 0 && (D == 1 || D == 2), "lstm_layer: bad shape");
  if (ws_bytes < mpa_lstm_layer_workspace(B, T, H, D)) {
    set_error("lstm_layer: workspace %zu < %zu bytes", ws_bytes, mpa_lstm_layer_workspace(B, T, H, D));
    return MPA_ERR_WORKSPACE;
  }
  float* G = (float*)workspace;
  float* cstate = G + (size_t)D * B * T * 4 * H;
  return lstm_forward(x, w_ih, w_hh, b_ih, b_hh, out, B, T, I, H, D, G, cstate, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
