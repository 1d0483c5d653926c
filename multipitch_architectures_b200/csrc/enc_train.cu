// Training path of the attention half of transformer_enc_layer (libdl/nn_models/unet_cnns.py:107-159) in fp32:
//   forward   t = Dropout(tokens + PE);  qkv = t W_qkv^T + b;  att = batch-axis softmax attention;  u1 = t + Dropout(att W_proj^T + b_proj);
//             h1 = LayerNorm1(u1)
//   backward  the exact adjoint, given the gradient wrt h1.
// The S bottleneck positions are independent (the attention runs over the BATCH axis), so ONE CTA owns one position and keeps
// its B x E token tile, q | k | v and every intermediate in shared memory: gather + PE + dropout + projections + attention + residual +
// LayerNorm are one launch instead of 8 forward / 11 backward launches whose GEMMs had M = B*S = 1300 rows on 42 CTAs.  What needs ALL
// tokens — the weight gradients dW = g^T x (K = B*S) — stays outside as GEMMs on the tensors this kernel writes.
// The parameter-only products of the layer (the q/k/v and o/out_proj Linear pairs folded into one matrix each, forward; the chain rule
// through the fold, backward) are one launch each (enc_fold_kernel / enc_fold_bwd_kernel) instead of 5 / 9 GEMM launches of 128^3.
// Dropout masks follow dropout_kernel's convention (element index of the [B*S, E] token matrix, Philox counter i/4) so that the fused path
// equals the stage-by-stage path element for element.
#include "common.cuh"

namespace mpa {

// ---- parameter folding ----------------------------------------------------------------------------------------------------------------
// C[i][j] = sum_k A(i,k) B(k,j) on E x E matrices; 32 x 32 tile per block of (32, 8) threads, 4 rows per thread
template <bool TA, bool TB>
__device__ __forceinline__ void small_gemm_tile(const float* __restrict__ A, const float* __restrict__ Bm, int E, int i0, int j0, float (&acc)[4],
                                                float (*As)[33], float (*Bs)[33]) {
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int r = 0; r < 4; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < E; k0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int rr = ty * 4 + r;
      // As[ii][kk] = A(i0+ii, k0+kk); Bs[kk][jj] = B(k0+kk, j0+jj); the fast thread index follows the contiguous dimension
      if (TA) As[tx][rr] = A[(size_t)(k0 + rr) * E + i0 + tx];
      else As[rr][tx] = A[(size_t)(i0 + rr) * E + k0 + tx];
      if (TB) Bs[tx][rr] = Bm[(size_t)(j0 + rr) * E + k0 + tx];
      else Bs[rr][tx] = Bm[(size_t)(k0 + rr) * E + j0 + tx];
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float b = Bs[kk][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(As[ty * 4 + r][kk], b, acc[r]);
    }
    __syncthreads();
  }
}

// grid (E/32, E/32, 4).  z < 3: W_qkv block z = Wi_z . Wx_z;  z = 3: W_proj = Wo . Wout, b_proj = Wo . bout.  Both layouts are written:
// [rows][E] (the backward's operand) and the transpose [E][rows] (the forward's operand: consecutive threads read consecutive outputs).
__global__ void __launch_bounds__(256) enc_fold_kernel(const float* __restrict__ Wi, const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                       const float* __restrict__ Wv, const float* __restrict__ Wo, const float* __restrict__ Wout,
                                                       const float* __restrict__ bout, float* __restrict__ w_qkv, float* __restrict__ w_qkvT,
                                                       float* __restrict__ w_proj, float* __restrict__ w_projT, float* __restrict__ b_proj, int E) {
  __shared__ float As[32][33], Bs[32][33];
  const int z = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4];
  if (z < 3) {
    small_gemm_tile<false, false>(Wi + (size_t)z * E * E, z == 0 ? Wq : z == 1 ? Wk : Wv, E, i0, j0, acc, As, Bs);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + threadIdx.y * 4 + r, j = j0 + threadIdx.x;
      w_qkv[(size_t)(z * E + i) * E + j] = acc[r];
      w_qkvT[(size_t)j * 3 * E + z * E + i] = acc[r];
    }
  } else {
    small_gemm_tile<false, false>(Wo, Wout, E, i0, j0, acc, As, Bs);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + threadIdx.y * 4 + r, j = j0 + threadIdx.x;
      w_proj[(size_t)i * E + j] = acc[r];
      w_projT[(size_t)j * E + i] = acc[r];
    }
    if (blockIdx.x == 0) {                       // b_proj rows i0 .. i0+31: one warp-wide dot product per row
      const int warp = threadIdx.y, lane = threadIdx.x;
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + warp * 4 + r;
        float s = 0.f;
        for (int k = lane; k < E; k += 32) s = fmaf(Wo[(size_t)i * E + k], bout[k], s);
        s = warp_sum(s);
        if (lane == 0) b_proj[i] = s;
      }
    }
  }
}

// grid (E/32, E/32, 8): the chain rule through the fold (all outputs overwritten)
//   z 0-2: G in_proj_weight block z = dWf_z . Wx_z^T        z 3-5: G {q,k,v}_linear.weight = Wi_z^T . dWf_z
//   z 6:   G o_linear.weight = dWp . Wout^T + dbp bout^T    z 7:   G out_proj.weight = Wo^T . dWp,  G out_proj.bias = Wo^T dbp
__global__ void __launch_bounds__(256) enc_fold_bwd_kernel(const float* __restrict__ dWf, const float* __restrict__ dWp, const float* __restrict__ dbp,
                                                           const float* __restrict__ Wi, const float* __restrict__ Wq, const float* __restrict__ Wk,
                                                           const float* __restrict__ Wv, const float* __restrict__ Wo, const float* __restrict__ Wout,
                                                           const float* __restrict__ bout, float* __restrict__ g_Wi, float* __restrict__ g_Wq,
                                                           float* __restrict__ g_Wk, float* __restrict__ g_Wv, float* __restrict__ g_Wo,
                                                           float* __restrict__ g_Wout, float* __restrict__ g_bout, int E) {
  __shared__ float As[32][33], Bs[32][33];
  const int z = blockIdx.z, i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const size_t EE = (size_t)E * E;
  float acc[4];
  float* out;
  if (z < 3) {
    small_gemm_tile<false, true>(dWf + z * EE, z == 0 ? Wq : z == 1 ? Wk : Wv, E, i0, j0, acc, As, Bs);
    out = g_Wi + z * EE;
  } else if (z < 6) {
    small_gemm_tile<true, false>(Wi + (z - 3) * EE, dWf + (z - 3) * EE, E, i0, j0, acc, As, Bs);
    out = z == 3 ? g_Wq : z == 4 ? g_Wk : g_Wv;
  } else if (z == 6) {
    small_gemm_tile<false, true>(dWp, Wout, E, i0, j0, acc, As, Bs);
    out = g_Wo;
  } else {
    small_gemm_tile<true, false>(Wo, dWp, E, i0, j0, acc, As, Bs);
    out = g_Wout;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + threadIdx.y * 4 + r, j = j0 + threadIdx.x;
    float v = acc[r];
    if (z == 6) v = fmaf(dbp[i], bout[j], v);
    out[(size_t)i * E + j] = v;
  }
  if (z == 7 && blockIdx.x == 0 && threadIdx.y == 0) {      // g_bout[i] = sum_k Wo[k][i] dbp[k], rows i0 .. i0+31
    const int i = i0 + threadIdx.x;
    float s = 0.f;
    for (int k = 0; k < E; ++k) s = fmaf(Wo[(size_t)k * E + i], dbp[k], s);
    g_bout[i] = s;
  }
}

// ---- the per-position kernels ---------------------------------------------------------------------------------------------------------
constexpr int kBC = 32;          // batch items per register block of the q|k|v projection (B = 25: one pass over the weight column)
constexpr int kBG = 12;          // ... of the E-column projections, whose items are split over blockDim / E thread groups

// out[b][j] = sum_k in[b][k] * WT[k*ldw + j] + bias for the items b_lo <= b < b_hi and this thread's column j; in: shared [B][E].
// NB items per register block (one pass over the weight column per block).
template <int NB, class Store>
__device__ __forceinline__ void project_column(const float* __restrict__ in, int b_lo, int b_hi, int E, const float* __restrict__ WT, int ldw, int j,
                                               float bias, Store store) {
  for (int b0 = b_lo; b0 < b_hi; b0 += NB) {
    float acc[NB];
#pragma unroll
    for (int bb = 0; bb < NB; ++bb) acc[bb] = bias;
#pragma unroll 2
    for (int k = 0; k < E; k += 4) {          // 8 independent weight loads in flight: the loop is bound by the L2 latency of WT otherwise
      const float w0 = WT[(size_t)k * ldw + j], w1 = WT[(size_t)(k + 1) * ldw + j], w2 = WT[(size_t)(k + 2) * ldw + j],
                  w3 = WT[(size_t)(k + 3) * ldw + j];
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const int b = min(b0 + bb, b_hi - 1);
        const float4 v = *reinterpret_cast<const float4*>(in + (size_t)b * E + k);
        acc[bb] = fmaf(v.x, w0, acc[bb]);
        acc[bb] = fmaf(v.y, w1, acc[bb]);
        acc[bb] = fmaf(v.z, w2, acc[bb]);
        acc[bb] = fmaf(v.w, w3, acc[bb]);
      }
    }
#pragma unroll
    for (int bb = 0; bb < NB; ++bb)
      if (b0 + bb < b_hi) store(b0 + bb, acc[bb]);
  }
}
// The E-column projections use ALL threads: thread = (column tid % E, item group tid / E), each group a contiguous range of the B items
// (with E = 128 columns on 128 of the 384 threads the other 8 warps waited at the next barrier: a third of the forward kernel's time)
__device__ __forceinline__ void item_range(int B, int E, int& col, int& b_lo, int& b_hi) {
  const int groups = max(1, (int)blockDim.x / E), grp = threadIdx.x / E;
  col = threadIdx.x - grp * E;
  const int per = (B + groups - 1) / groups;
  b_lo = min(B, grp * per);
  b_hi = grp < groups ? min(B, b_lo + per) : b_lo;
}

__device__ __forceinline__ float drop_fac(const DropoutArgs& d, unsigned long long off, float scale, long long i) {
  return d.p > 0.f ? dropout_factor_at(i, d.p, scale, d.seed, off) : 1.f;
}

struct EncTrainParams {
  const float *x, *pe;                           // [B,E,S] NCHW tokens source; [S,E] or null
  const float *w_qkv, *w_qkvT, *b_qkv;           // [3E,E], [E,3E], [3E]
  const float *w_proj, *w_projT, *b_proj;        // [E,E], [E,E]^T, [E]
  const float *ln_w, *ln_b;
  float eps;
  float *t, *qkv, *att, *u1, *h1;                // saved / produced token tensors [B*S, ...]
  int B, E, S, H;
  DropoutArgs d_tok, d_att;
  // backward only
  const float* g_h1;                             // [B*S,E]
  float *g_p, *g_qkv, *g_x;                      // [B*S,E], [B*S,3E], [B,E,S]
  float *g_ln_w, *g_ln_b;                        // accumulated with atomics (pre-zeroed by the caller)
};

__global__ void __launch_bounds__(384) enc_attn_train_fwd_kernel(const EncTrainParams p) {
  extern __shared__ __align__(16) float sm[];
  const int B = p.B, E = p.E, S = p.S, H = p.H, hd = E / H, E3 = 3 * E;
  // q | k | v of one item in shared memory: [3 parts][H heads][hd + 4] — the 4-float pad per head spreads the heads over the banks
  // (lanes of a warp differ in the head: with the dense layout the key / value reads of the attention loop were 4-way, the query reads
  // 16-way bank conflicted and the attention step took ~100 us of this kernel)
  const int HS = hd + 4, PS = H * HS, RS = 3 * PS;
  float* t = sm;                       // [B][E]
  float* qkv = t + (size_t)B * E;      // [B][RS]
  float* att = qkv + (size_t)B * RS;   // [B][E]
  float* u1 = att + (size_t)B * E;     // [B][E]
  const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const unsigned long long off_t = p.d_tok.p > 0.f ? dropout_offset(p.d_tok) : 0ull, off_a = p.d_att.p > 0.f ? dropout_offset(p.d_att) : 0ull;
  const float sc_t = p.d_tok.p > 0.f ? 1.f / (1.f - p.d_tok.p) : 1.f, sc_a = p.d_att.p > 0.f ? 1.f / (1.f - p.d_att.p) : 1.f;
  // 1. gather + PE + dropout
  for (int i = tid; i < B * E; i += nt) {
    const int b = i / E, e = i - b * E;
    float v = p.x[((size_t)b * E + e) * S + s];
    if (p.pe) v += p.pe[(size_t)s * E + e];
    const long long gi = ((long long)b * S + s) * E + e;
    if (p.d_tok.p > 0.f) {
      const float f = drop_fac(p.d_tok, off_t, sc_t, gi);
      v = f != 0.f ? v * f : 0.f;
    }
    t[i] = v;
    p.t[gi] = v;
  }
  __syncthreads();
  // 2. q | k | v
  for (int j = tid; j < E3; j += nt) {
    const int part = j / E, c = j - part * E, hh = c / hd;
    const int js = part * PS + hh * HS + (c - hh * hd);
    project_column<kBC>(t, 0, B, E, p.w_qkvT, E3, j, p.b_qkv[j], [&](int b, float v) {
      qkv[(size_t)b * RS + js] = v;
      p.qkv[((size_t)b * S + s) * E3 + j] = v;
    });
  }
  __syncthreads();
  // 3. batch-axis attention: thread = (query item b1, head h); the same two-pass softmax as batch_axis_attention_kernel
  const float sc = rsqrtf((float)hd);
  for (int i = tid; i < B * H; i += nt) {
    const int b1 = i / H, h = i - b1 * H;
    float q[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) q[k] = k < hd ? qkv[(size_t)b1 * RS + h * HS + k] : 0.f;
    float mx = -INFINITY;
    for (int b2 = 0; b2 < B; ++b2) {
      const float* kr = qkv + (size_t)b2 * RS + PS + h * HS;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < hd) d = fmaf(q[k], kr[k], d);
      mx = fmaxf(mx, d * sc);
    }
    float den = 0.f, o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) o[k] = 0.f;
    for (int b2 = 0; b2 < B; ++b2) {
      const float* kr = qkv + (size_t)b2 * RS + PS + h * HS;
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < hd) d = fmaf(q[k], kr[k], d);
      const float pr = expf(d * sc - mx);
      den += pr;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < hd) o[k] = fmaf(pr, kr[PS + k], o[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < hd) {
        const float v = o[k] / den;
        att[(size_t)b1 * E + h * hd + k] = v;
        p.att[((size_t)b1 * S + s) * E + h * hd + k] = v;
      }
  }
  __syncthreads();
  // 4. out-projection + dropout + residual
  int e, b_lo, b_hi;
  item_range(B, E, e, b_lo, b_hi);
  if (b_lo < b_hi)
    project_column<kBG>(att, b_lo, b_hi, E, p.w_projT, E, e, p.b_proj[e], [&](int b, float v) {
      const long long gi = ((long long)b * S + s) * E + e;
      if (p.d_att.p > 0.f) {
        const float f = drop_fac(p.d_att, off_a, sc_a, gi);
        v = f != 0.f ? v * f : 0.f;
      }
      v = t[(size_t)b * E + e] + v;
      u1[(size_t)b * E + e] = v;
      p.u1[gi] = v;
    });
  __syncthreads();
  // 5. LayerNorm1: one warp per token
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int b = warp; b < B; b += nw) {
    float v[4], sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      v[i] = e < E ? u1[(size_t)b * E + e] : 0.f;
      sum += v[i];
    }
    const float mean = warp_sum(sum) / E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (lane + 32 * i < E) {
        const float d = v[i] - mean;
        q += d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) / E + p.eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      if (e < E) p.h1[((size_t)b * S + s) * E + e] = (v[i] - mean) * rstd * p.ln_w[e] + p.ln_b[e];
    }
  }
}

__global__ void __launch_bounds__(384) enc_attn_train_bwd_kernel(const EncTrainParams p) {
  extern __shared__ __align__(16) float sm[];
  const int B = p.B, E = p.E, S = p.S, H = p.H, hd = E / H, E3 = 3 * E;
  float* g_u1 = sm;                        // [B][E]
  float* xh = g_u1 + (size_t)B * E;        // [B][E]  LayerNorm xhat, then g_p
  float* g_att = xh + (size_t)B * E;       // [B][E]
  float* g_qkv = g_att + (size_t)B * E;    // [B][3E]
  const int HP = hd + 1;                   // odd row stride of the per-head tiles: the (b1, b2) passes read them conflict-free
  float* Qh = g_qkv + (size_t)B * E3;      // [B][HP] this head's q, k, v and dO
  float* Kh = Qh + (size_t)B * HP;
  float* Vh = Kh + (size_t)B * HP;
  float* dOh = Vh + (size_t)B * HP;
  float* P = dOh + (size_t)B * HP;         // [B][B]
  float* dS = P + (size_t)B * B;           // [B][B]
  float* rowv = dS + (size_t)B * B;        // [B]
  const int s = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const unsigned long long off_t = p.d_tok.p > 0.f ? dropout_offset(p.d_tok) : 0ull, off_a = p.d_att.p > 0.f ? dropout_offset(p.d_att) : 0ull;
  const float sc_t = p.d_tok.p > 0.f ? 1.f / (1.f - p.d_tok.p) : 1.f, sc_a = p.d_att.p > 0.f ? 1.f / (1.f - p.d_att.p) : 1.f;
  // 1. LayerNorm1 backward, one warp per token (the arithmetic of ln_bwd_kernel); xh keeps xhat, g_att (scratch) keeps g_y
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  for (int b = warp; b < B; b += nw) {
    float v[4], g[4], sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      v[i] = e < E ? p.u1[((size_t)b * S + s) * E + e] : 0.f;
      sum += v[i];
    }
    const float mean = warp_sum(sum) / E;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (lane + 32 * i < E) {
        const float d = v[i] - mean;
        q += d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) / E + p.eps);
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      g[i] = 0.f;
      if (e < E) {
        const float x_hat = (v[i] - mean) * rstd;
        const float gy = p.g_h1[((size_t)b * S + s) * E + e];
        xh[(size_t)b * E + e] = x_hat;
        g_att[(size_t)b * E + e] = gy;
        g[i] = gy * p.ln_w[e];
        a1 += g[i];
        a2 += g[i] * x_hat;
        v[i] = x_hat;
      }
    }
    a1 = warp_sum(a1) / E;
    a2 = warp_sum(a2) / E;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = lane + 32 * i;
      if (e < E) g_u1[(size_t)b * E + e] = rstd * (g[i] - a1 - v[i] * a2);
    }
  }
  __syncthreads();
  // LayerNorm parameter gradients of this position's B tokens, then g_p = dropout mask * g_u1 (xh is overwritten by g_p)
  int col, b_lo, b_hi;
  item_range(B, E, col, b_lo, b_hi);
  if (b_lo < b_hi) {
    const int e = col;
    float gw = 0.f, gb = 0.f;
    for (int b = b_lo; b < b_hi; ++b) {
      const float gy = g_att[(size_t)b * E + e];
      gw = fmaf(gy, xh[(size_t)b * E + e], gw);
      gb += gy;
    }
    atomicAdd(&p.g_ln_w[e], gw);
    atomicAdd(&p.g_ln_b[e], gb);
    for (int b = b_lo; b < b_hi; ++b) {
      const long long gi = ((long long)b * S + s) * E + e;
      float v = g_u1[(size_t)b * E + e];
      if (p.d_att.p > 0.f) {
        const float f = drop_fac(p.d_att, off_a, sc_a, gi);
        v = f != 0.f ? v * f : 0.f;
      }
      xh[(size_t)b * E + e] = v;
      p.g_p[gi] = v;
    }
  }
  __syncthreads();
  // 2. g_att = g_p W_proj   (g_att[b][k] = sum_e g_p[b][e] w_proj[e][k])
  if (b_lo < b_hi) project_column<kBG>(xh, b_lo, b_hi, E, p.w_proj, E, col, 0.f, [&](int b, float v) { g_att[(size_t)b * E + col] = v; });
  __syncthreads();
  // 3. attention backward, head by head (the passes of batch_axis_attention_bwd_kernel on this CTA's shared tiles)
  const float sc = rsqrtf((float)hd);
  for (int h = 0; h < H; ++h) {
    for (int e = tid; e < B * hd; e += nt) {
      const int b = e / hd, k = e - b * hd;
      const float* row = p.qkv + ((size_t)b * S + s) * E3 + h * hd + k;
      Qh[b * HP + k] = row[0];
      Kh[b * HP + k] = row[E];
      Vh[b * HP + k] = row[2 * E];
      dOh[b * HP + k] = g_att[(size_t)b * E + h * hd + k];
    }
    __syncthreads();
    for (int e = tid; e < B * B; e += nt) {
      const int b1 = e / B, b2 = e - b1 * B;
      float d = 0.f, dp = 0.f;
      for (int k = 0; k < hd; ++k) {
        d = fmaf(Qh[b1 * HP + k], Kh[b2 * HP + k], d);
        dp = fmaf(dOh[b1 * HP + k], Vh[b2 * HP + k], dp);
      }
      P[e] = d * sc;
      dS[e] = dp;
    }
    __syncthreads();
    for (int b1 = tid; b1 < B; b1 += nt) {
      float mx = -INFINITY;
      for (int b2 = 0; b2 < B; ++b2) mx = fmaxf(mx, P[b1 * B + b2]);
      float den = 0.f;
      for (int b2 = 0; b2 < B; ++b2) {
        const float pr = expf(P[b1 * B + b2] - mx);
        P[b1 * B + b2] = pr;
        den += pr;
      }
      float dot = 0.f;
      for (int b2 = 0; b2 < B; ++b2) {
        const float pr = P[b1 * B + b2] / den;
        P[b1 * B + b2] = pr;
        dot = fmaf(pr, dS[b1 * B + b2], dot);
      }
      rowv[b1] = dot;
    }
    __syncthreads();
    for (int e = tid; e < B * B; e += nt) dS[e] = P[e] * (dS[e] - rowv[e / B]) * sc;
    __syncthreads();
    for (int e = tid; e < B * hd; e += nt) {
      const int b = e / hd, k = e - b * hd;
      float dq = 0.f, dk = 0.f, dv = 0.f;
      for (int o = 0; o < B; ++o) {
        dq = fmaf(dS[b * B + o], Kh[o * HP + k], dq);
        dk = fmaf(dS[o * B + b], Qh[o * HP + k], dk);
        dv = fmaf(P[o * B + b], dOh[o * HP + k], dv);
      }
      float* row = g_qkv + (size_t)b * E3 + h * hd + k;
      row[0] = dq;
      row[E] = dk;
      row[2 * E] = dv;
    }
    __syncthreads();
  }
  for (int i = tid; i < B * E3; i += nt) {
    const int b = i / E3, j = i - b * E3;
    p.g_qkv[((size_t)b * S + s) * E3 + j] = g_qkv[i];
  }
  // 4. g_t = g_qkv W_qkv + g_u1 (residual), through the token dropout mask, scattered back to NCHW
  const int e = col;
  if (b_lo < b_hi)
    project_column<kBG>(g_qkv, b_lo, b_hi, E3, p.w_qkv, E, e, 0.f, [&](int b, float v) {
      v += g_u1[(size_t)b * E + e];
      if (p.d_tok.p > 0.f) {
        const float f = drop_fac(p.d_tok, off_t, sc_t, ((long long)b * S + s) * E + e);
        v = f != 0.f ? v * f : 0.f;
      }
      p.g_x[((size_t)b * E + e) * S + s] = v;
    });
}

static size_t enc_fwd_smem(int B, int E, int H) { return sizeof(float) * (size_t)B * (3 * E + 3 * (E + 4 * H)); }
static size_t enc_bwd_smem(int B, int E, int H) { return sizeof(float) * ((size_t)B * E * 6 + 4 * (size_t)B * (E / H + 1) + 2 * (size_t)B * B + B); }

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_enc_train_supported(int B, int E, int num_heads) {
  return E % 32 == 0 && E <= 128 && num_heads > 0 && E % num_heads == 0 && E / num_heads <= 16 && B >= 1 &&
         enc_bwd_smem(B, E, num_heads) <= 220 * 1024 && enc_fwd_smem(B, E, num_heads) <= 220 * 1024;
}

int mpa_enc_fold_f32(const float* in_proj_weight, const float* wq, const float* wk, const float* wv, const float* wo, const float* out_proj_weight,
                     const float* out_proj_bias, float* w_qkv, float* w_qkvT, float* w_proj, float* w_projT, float* b_proj, int E, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(in_proj_weight && wq && wk && wv && wo && out_proj_weight && out_proj_bias && w_qkv && w_qkvT && w_proj && w_projT && b_proj &&
                  E > 0 && E % 32 == 0,
              "enc_fold: bad argument (E must be a multiple of 32)");
  enc_fold_kernel<<<dim3(E / 32, E / 32, 4), dim3(32, 8), 0, (cudaStream_t)stream>>>(in_proj_weight, wq, wk, wv, wo, out_proj_weight, out_proj_bias,
                                                                                    w_qkv, w_qkvT, w_proj, w_projT, b_proj, E);
  MPA_CHECK_LAUNCH("enc_fold");
  return MPA_OK;
}

int mpa_enc_fold_bwd_f32(const float* d_w_qkv, const float* d_w_proj, const float* d_b_proj, const float* in_proj_weight, const float* wq,
                         const float* wk, const float* wv, const float* wo, const float* out_proj_weight, const float* out_proj_bias,
                         float* g_in_proj_weight, float* g_wq, float* g_wk, float* g_wv, float* g_wo, float* g_out_proj_weight,
                         float* g_out_proj_bias, int E, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(d_w_qkv && d_w_proj && d_b_proj && in_proj_weight && wq && wk && wv && wo && out_proj_weight && out_proj_bias && g_in_proj_weight &&
                  g_wq && g_wk && g_wv && g_wo && g_out_proj_weight && g_out_proj_bias && E > 0 && E % 32 == 0,
              "enc_fold_bwd: bad argument (E must be a multiple of 32)");
  enc_fold_bwd_kernel<<<dim3(E / 32, E / 32, 8), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      d_w_qkv, d_w_proj, d_b_proj, in_proj_weight, wq, wk, wv, wo, out_proj_weight, out_proj_bias, g_in_proj_weight, g_wq, g_wk, g_wv, g_wo,
      g_out_proj_weight, g_out_proj_bias, E);
  MPA_CHECK_LAUNCH("enc_fold_bwd");
  return MPA_OK;
}

static unsigned char g_enc_fwd_flags[64], g_enc_bwd_flags[64];

int mpa_enc_attn_train_fwd_f32(const float* x, const float* pe, const float* w_qkvT, const float* b_qkv, const float* w_projT, const float* b_proj,
                               const float* ln_w, const float* ln_b, float eps, float* t, float* qkv, float* att, float* u1, float* h1, int B, int E,
                               int S, int num_heads, float p_drop, unsigned long long seed, unsigned long long site_tok,
                               unsigned long long site_att, const long long* step_dev, unsigned long long step_mul, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_qkvT && b_qkv && w_projT && b_proj && ln_w && ln_b && t && qkv && att && u1 && h1 && S > 0 && p_drop >= 0.f && p_drop < 1.f,
              "enc_attn_train_fwd: bad argument");
  MPA_REQUIRE(mpa_enc_train_supported(B, E, num_heads), "enc_attn_train_fwd: unsupported shape (E %% 32 == 0, E <= 128, head dim <= 16, B*E tile in shared memory)");
  if (opt_in_max_smem(enc_attn_train_fwd_kernel, g_enc_fwd_flags) != cudaSuccess) {
    set_error("enc_attn_train_fwd: cannot opt in to large shared memory");
    return MPA_ERR_CUDA;
  }
  EncTrainParams p{};
  p.x = x; p.pe = pe; p.w_qkvT = w_qkvT; p.b_qkv = b_qkv; p.w_projT = w_projT; p.b_proj = b_proj; p.ln_w = ln_w; p.ln_b = ln_b; p.eps = eps;
  p.t = t; p.qkv = qkv; p.att = att; p.u1 = u1; p.h1 = h1; p.B = B; p.E = E; p.S = S; p.H = num_heads;
  p.d_tok = DropoutArgs{pe ? p_drop : 0.f, seed, site_tok, step_dev, step_mul};       // the reference drops the tokens only behind a positional encoding
  p.d_att = DropoutArgs{p_drop, seed, site_att, step_dev, step_mul};
  enc_attn_train_fwd_kernel<<<S, 384, enc_fwd_smem(B, E, num_heads), (cudaStream_t)stream>>>(p);
  MPA_CHECK_LAUNCH("enc_attn_train_fwd");
  return MPA_OK;
}

int mpa_enc_attn_train_bwd_f32(const float* g_h1, const float* u1, const float* qkv, const float* w_qkv, const float* w_proj, const float* ln_w,
                               float eps, float* g_p, float* g_qkv, float* g_x, float* g_ln_w, float* g_ln_b, int B, int E, int S, int num_heads,
                               int has_pe, float p_drop, unsigned long long seed, unsigned long long site_tok, unsigned long long site_att,
                               const long long* step_dev, unsigned long long step_mul, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(g_h1 && u1 && qkv && w_qkv && w_proj && ln_w && g_p && g_qkv && g_x && g_ln_w && g_ln_b && S > 0 && p_drop >= 0.f && p_drop < 1.f,
              "enc_attn_train_bwd: bad argument");
  MPA_REQUIRE(mpa_enc_train_supported(B, E, num_heads), "enc_attn_train_bwd: unsupported shape");
  if (opt_in_max_smem(enc_attn_train_bwd_kernel, g_enc_bwd_flags) != cudaSuccess) {
    set_error("enc_attn_train_bwd: cannot opt in to large shared memory");
    return MPA_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(g_ln_w, 0, sizeof(float) * E, st);
  cudaMemsetAsync(g_ln_b, 0, sizeof(float) * E, st);
  EncTrainParams p{};
  p.g_h1 = g_h1; p.u1 = const_cast<float*>(u1); p.qkv = const_cast<float*>(qkv); p.w_qkv = w_qkv; p.w_proj = w_proj; p.ln_w = ln_w; p.eps = eps;
  p.g_p = g_p; p.g_qkv = g_qkv; p.g_x = g_x; p.g_ln_w = g_ln_w; p.g_ln_b = g_ln_b; p.B = B; p.E = E; p.S = S; p.H = num_heads;
  p.d_tok = DropoutArgs{has_pe ? p_drop : 0.f, seed, site_tok, step_dev, step_mul};
  p.d_att = DropoutArgs{p_drop, seed, site_att, step_dev, step_mul};
  enc_attn_train_bwd_kernel<<<S, 384, enc_bwd_smem(B, E, num_heads), st>>>(p);
  MPA_CHECK_LAUNCH("enc_attn_train_bwd");
  return MPA_OK;
}

}  // extern "C"
