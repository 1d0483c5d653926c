// transformer_enc_layer of the SAUnet bottleneck (reference: libdl/nn_models/unet_cnns.py:107-159).
//
// Semantics to preserve: the reference hands [B, S, E] to nn.MultiheadAttention without batch_first, so the
// attention SEQUENCE axis is the batch B (<= 50) and the S = Th*Fw bottleneck positions are independent
// "batch" items.  One launch sequence: gather(+PE) -> fused QKV GEMM -> batch-axis attention -> projection GEMM
// -> add&LayerNorm -> MLP GEMM (ReLU) -> MLP GEMM -> add&LayerNorm + scatter back to NCHW.  The q/k/v Linear
// layers are folded into the MHA in-projection and o_linear into the out-projection on the host (one-off).
#include "common.cuh"

namespace mpa {

// tokens[(b*S+s)*E + e] = x[b,e,s] (+ pe[s,e])
__global__ void enc_gather_kernel(const float* __restrict__ x, const float* __restrict__ pe, float* __restrict__ tok, long long total,
                                  int E, int S) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int e = (int)(i % E);
    long long r = i / E;
    int s = (int)(r % S);
    int b = (int)(r / S);
    float v = x[((size_t)b * E + e) * S + s];
    if (pe) v += pe[(size_t)s * E + e];
    tok[i] = v;
  }
}

__device__ __forceinline__ int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }
static inline int pick_ksplit(int blocks, int K, int allow) {
  if (!allow || blocks >= 148 || K < 512) return 1;
  int ks = 296 / (blocks < 1 ? 1 : blocks);
  if (ks > K / 128) ks = K / 128;
  if (ks > 32) ks = 32;
  return ks < 1 ? 1 : ks;
}

// C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); 64x64 tile, 256 threads, 4x4 register tile, K step 16
__global__ void __launch_bounds__(256) gemm_nt_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                                                      float* __restrict__ C, int M, int N, int K, int relu) {
  __shared__ float As[16][64 + 4];
  __shared__ float Ws[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  // split-K over blockIdx.z (wide-K, small-output products: the slices are combined with atomicAdd into a pre-zeroed C)
  const int kchunk = (ceil_div_dev(K, (int)gridDim.z) + 15) / 16 * 16;
  const int k_begin = blockIdx.z * kchunk, k_end = min(K, k_begin + kchunk);
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int kk = e & 15, r = e >> 4;
      int m = m0 + r, n = n0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < k_end) ? A[(size_t)m * K + k] : 0.f;
      Ws[kk][r] = (n < N && k < k_end) ? W[(size_t)n * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = Ws[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + ((bias && blockIdx.z == 0) ? bias[n] : 0.f);
      if (gridDim.z > 1) atomicAdd(&C[(size_t)m * N + n], v);
      else C[(size_t)m * N + n] = relu ? fmaxf(v, 0.f) : v;
    }
  }
}

// one block per (s, head): sequence axis = b.  qkv [(b*S+s)][3E]; out [(b*S+s)][E]
__global__ void __launch_bounds__(64) batch_axis_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int B, int S, int E,
                                                                  int H) {
  extern __shared__ float sm[];
  const int hd = E / H;
  const int s = blockIdx.x / H, h = blockIdx.x % H;
  float* Ks = sm;                 // [B][hd]
  float* Vs = sm + (size_t)B * hd;
  for (int e = threadIdx.x; e < B * hd; e += blockDim.x) {
    int b = e / hd, d = e - b * hd;
    const float* row = qkv + ((size_t)b * S + s) * 3 * E + h * hd + d;
    Ks[e] = row[E];
    Vs[e] = row[2 * E];
  }
  __syncthreads();
  const float sc = rsqrtf((float)hd);
  for (int b1 = threadIdx.x; b1 < B; b1 += blockDim.x) {
    const float* q = qkv + ((size_t)b1 * S + s) * 3 * E + h * hd;
    float mx = -INFINITY;
    for (int b2 = 0; b2 < B; ++b2) {
      float d = 0.f;
      for (int k = 0; k < hd; ++k) d = fmaf(q[k], Ks[b2 * hd + k], d);
      mx = fmaxf(mx, d * sc);
    }
    float den = 0.f;
    float o[64];
    for (int k = 0; k < hd; ++k) o[k] = 0.f;
    for (int b2 = 0; b2 < B; ++b2) {
      float d = 0.f;
      for (int k = 0; k < hd; ++k) d = fmaf(q[k], Ks[b2 * hd + k], d);
      float pr = expf(d * sc - mx);
      den += pr;
      for (int k = 0; k < hd; ++k) o[k] = fmaf(pr, Vs[b2 * hd + k], o[k]);
    }
    float* orow = out + ((size_t)b1 * S + s) * E + h * hd;
    for (int k = 0; k < hd; ++k) orow[k] = o[k] / den;
  }
}

// y = LayerNorm(a + b) over E per token; optional scatter to NCHW [B,E,S]
__global__ void __launch_bounds__(128) add_ln_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ out_tok, float* __restrict__ out_nchw, int E,
                                                     int S, float eps) {
  __shared__ float sh[4];
  const long long tok = blockIdx.x;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int e = threadIdx.x + i * 128;
    v[i] = (e < E) ? a[tok * E + e] + b[tok * E + e] : 0.f;
    s += v[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  float mean = (sh[0] + sh[1] + sh[2] + sh[3]) / E;
  __syncthreads();
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int e = threadIdx.x + i * 128;
    if (e < E) {
      float d = v[i] - mean;
      q += d * d;
    }
  }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = q;
  __syncthreads();
  float rstd = rsqrtf((sh[0] + sh[1] + sh[2] + sh[3]) / E + eps);
  const int bb = (int)(tok / S), ss = (int)(tok % S);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int e = threadIdx.x + i * 128;
    if (e < E) {
      float r = (v[i] - mean) * rstd * w[e] + bias[e];
      if (out_tok) out_tok[tok * E + e] = r;
      if (out_nchw) out_nchw[((size_t)bb * E + e) * S + ss] = r;
    }
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace mpa

using namespace mpa;

extern "C" {

size_t mpa_encoder_layer_workspace(int B, int E, int S, int mlp_dim) {
  size_t n = (size_t)B * S;
  // tok (E), qkv (3E), att (E), proj (E), h1 (E), mlp out (E) = 8E, plus the MLP hidden layer
  return sizeof(float) * n * ((size_t)E * 8 + (size_t)mlp_dim) + 256;
}

int mpa_encoder_layer_f32(const float* x, float* out, int B, int E, int Th, int Fw, int num_heads, int mlp_dim, const float* pe,
                          const float* w_qkv, const float* b_qkv, const float* w_proj, const float* b_proj, const float* ln1_w,
                          const float* ln1_b, const float* mlp0_w, const float* mlp0_b, const float* mlp2_w, const float* mlp2_b,
                          const float* ln2_w, const float* ln2_b, float eps, void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && out && w_qkv && b_qkv && w_proj && b_proj && ln1_w && ln1_b && mlp0_w && mlp0_b && mlp2_w && mlp2_b && ln2_w && ln2_b && workspace,
              "encoder_layer: null argument");
  MPA_REQUIRE(B > 0 && E > 0 && E <= 1024 && num_heads > 0 && E % num_heads == 0 && E / num_heads <= 64 && mlp_dim > 0,
              "encoder_layer: unsupported E=%d heads=%d", E, num_heads);
  const int S = Th * Fw;
  const size_t need = mpa_encoder_layer_workspace(B, E, S, mlp_dim);
  if (ws_bytes < need) {
    set_error("encoder_layer: workspace %zu < %zu bytes", ws_bytes, need);
    return MPA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * S;
  float* tok = (float*)workspace;
  float* qkv = tok + n * E;
  float* att = qkv + n * 3 * E;
  float* proj = att + n * E;
  float* h1 = proj + n * E;
  float* mo = h1 + n * E;
  float* hid = mo + n * E;
  enc_gather_kernel<<<grid_for(n * E, 256), 256, 0, st>>>(x, pe, tok, n * E, E, S);
  MPA_CHECK_LAUNCH("enc_gather");
  gemm_nt_kernel<<<dim3(ceil_div(3 * E, 64), ceil_div(n, 64)), 256, 0, st>>>(tok, w_qkv, b_qkv, qkv, (int)n, 3 * E, E, 0);
  MPA_CHECK_LAUNCH("gemm_qkv");
  const int hd = E / num_heads;
  batch_axis_attention_kernel<<<S * num_heads, 64, 2 * (size_t)B * hd * sizeof(float), st>>>(qkv, att, B, S, E, num_heads);
  MPA_CHECK_LAUNCH("batch_axis_attention");
  gemm_nt_kernel<<<dim3(ceil_div(E, 64), ceil_div(n, 64)), 256, 0, st>>>(att, w_proj, b_proj, proj, (int)n, E, E, 0);
  MPA_CHECK_LAUNCH("gemm_proj");
  add_ln_kernel<<<(unsigned)n, 128, 0, st>>>(tok, proj, ln1_w, ln1_b, h1, nullptr, E, S, eps);
  MPA_CHECK_LAUNCH("add_ln1");
  gemm_nt_kernel<<<dim3(ceil_div(mlp_dim, 64), ceil_div(n, 64)), 256, 0, st>>>(h1, mlp0_w, mlp0_b, hid, (int)n, mlp_dim, E, 1);
  MPA_CHECK_LAUNCH("gemm_mlp0");
  {
    const int ks = pick_ksplit(ceil_div(E, 64) * ceil_div(n, 64), mlp_dim, 1);
    if (ks > 1) cudaMemsetAsync(mo, 0, sizeof(float) * (size_t)n * E, st);
    gemm_nt_kernel<<<dim3(ceil_div(E, 64), ceil_div(n, 64), ks), 256, 0, st>>>(hid, mlp2_w, mlp2_b, mo, (int)n, E, mlp_dim, 0);
  }
  MPA_CHECK_LAUNCH("gemm_mlp2");
  add_ln_kernel<<<(unsigned)n, 128, 0, st>>>(h1, mo, ln2_w, ln2_b, nullptr, out, E, S, eps);
  MPA_CHECK_LAUNCH("add_ln2");
  return MPA_OK;
}

}  // extern "C"

// ---- the same stages as separate entry points (the training path interleaves dropout and keeps every intermediate) --------
extern "C" {

int mpa_enc_gather_f32(const float* x, const float* pe, float* tok, int B, int E, int S, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && tok && B > 0 && E > 0 && S > 0, "enc_gather: bad argument");
  const long long n = (long long)B * S * E;
  enc_gather_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, pe, tok, n, E, S);
  MPA_CHECK_LAUNCH("enc_gather");
  return MPA_OK;
}

int mpa_gemm_nt_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int relu, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, "gemm_nt: bad argument");
  {
    const int ks = pick_ksplit(ceil_div(N, 64) * ceil_div(M, 64), K, !relu);
    if (ks > 1) cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, (cudaStream_t)stream);
    gemm_nt_kernel<<<dim3(ceil_div(N, 64), ceil_div(M, 64), ks), 256, 0, (cudaStream_t)stream>>>(A, W, bias, C, M, N, K, relu);
  }
  MPA_CHECK_LAUNCH("gemm_nt");
  return MPA_OK;
}

int mpa_batch_axis_attention_f32(const float* qkv, float* out, int B, int S, int E, int num_heads, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(qkv && out && B > 0 && S > 0 && E > 0 && num_heads > 0 && E % num_heads == 0 && E / num_heads <= 64, "attention: bad argument");
  const int hd = E / num_heads;
  batch_axis_attention_kernel<<<S * num_heads, 64, 2 * (size_t)B * hd * sizeof(float), (cudaStream_t)stream>>>(qkv, out, B, S, E, num_heads);
  MPA_CHECK_LAUNCH("batch_axis_attention");
  return MPA_OK;
}

int mpa_add_layernorm_tok_f32(const float* a, const float* b, const float* w, const float* bias, float* out_tok, float* out_nchw, long long n_tok,
                              int E, int S, float eps, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(a && b && w && bias && (out_tok || out_nchw) && n_tok > 0 && E > 0 && E <= 1024 && S > 0, "add_layernorm_tok: bad argument");
  add_ln_kernel<<<(unsigned)n_tok, 128, 0, (cudaStream_t)stream>>>(a, b, w, bias, out_tok, out_nchw, E, S, eps);
  MPA_CHECK_LAUNCH("add_ln");
  return MPA_OK;
}

}  // extern "C"
