// Full-height VALID convolution (the head's "time reduction" conv3: 75x1 over a 75-frame patch -> one row, basic_cnns.py:396-401,
// unet_cnns.py:380-385) forward, data gradient and weight gradient as GEMMs on fp32 CUDA cores.
//   With H == KH and KW == 1 the layer is   Y[co][(b,w)] = sum_k Wm[co][k] * X[k][(b,w)],  k = (ci, kh)  (K = Cin*75 = 1500 .. 11250),
//   and x[b] IS the row-major [K][W] matrix, so no im2col and no halo: one strided GEMM kernel serves all three products
//     forward  : M = Cout,  N = (b,w), K = Cin*H        (+ bias, activation)
//     dgrad    : M = Cin*H, N = (b,w), K = Cout
//     wgrad    : M = Cout,  N = Cin*H, K = (b,w)        (split-K, fp32 atomics)
//   Every GEMM index is a 2-level index (i1, i2) with its own strides, which is what lets (b,w) be a single GEMM dimension.
// The generic direct kernel evaluated this layer through 16-row output tiles of which one row exists (and a 90-row halo per tile):
// 7 ms of a 21 ms SAUnet:L training step; this formulation needs ~0.3 ms.
#include "common.cuh"

namespace mpa {

struct Idx2 {        // index i in [0, n): i1 = i / n2, i2 = i % n2 -> offset i1*s1 + i2*s2
  int n2;
  long long s1, s2;
  __device__ __forceinline__ long long off(int i) const { return (long long)(i / n2) * s1 + (long long)(i % n2) * s2; }
};

struct RowsGemm {
  const float* A;
  const float* B;
  float* C;
  const float* bias;     // per m or null
  int M, N, K;
  Idx2 am, ak, bk, bn, cm, cn;
  int act;
  float act_param;
  int a_k_fast, b_k_fast;   // which index is the unit-stride one (coalescing of the tile loads)
  float* partial;           // split-K forward: per-slice partial sums [ks][M][N], reduced IN ORDER by the second pass (deterministic)
};

// TM x TN output tile per CTA of 256 threads (16 x 16 threads, (TM/16) x (TN/16) outputs each), K step 16.
template <int TM, int TN>
__global__ void __launch_bounds__(256) rows_gemm_kernel(RowsGemm p) {
  constexpr int RM = TM / 16, RN = TN / 16;
  __shared__ float As[16][TM + 4];
  __shared__ float Bs[16][TN + 4];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[RM][RN] = {};
  const int kchunk = (((p.K + (int)gridDim.z - 1) / (int)gridDim.z) + 15) / 16 * 16;
  const int k_begin = blockIdx.z * kchunk, k_end = min(p.K, k_begin + kchunk);
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int e = threadIdx.x; e < TM * 16; e += 256) {
      const int kk = p.a_k_fast ? (e & 15) : (e / TM), r = p.a_k_fast ? (e >> 4) : (e % TM);
      const int m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < p.M && k < k_end) ? p.A[p.am.off(m) + p.ak.off(k)] : 0.f;
    }
    for (int e = threadIdx.x; e < TN * 16; e += 256) {
      const int kk = p.b_k_fast ? (e & 15) : (e / TN), r = p.b_k_fast ? (e >> 4) : (e % TN);
      const int n = n0 + r, k = k0 + kk;
      Bs[kk][r] = (n < p.N && k < k_end) ? p.B[p.bk.off(k) + p.bn.off(n)] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Bs[kk][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int m = m0 + ty * RM + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int n = n0 + tx * RN + j;
      if (n >= p.N) continue;
      float* c = p.C + p.cm.off(m) + p.cn.off(n);
      if (gridDim.z > 1 && p.partial) {
        p.partial[((size_t)blockIdx.z * p.M + m) * p.N + n] = acc[i][j];
      } else if (gridDim.z > 1) {
        atomicAdd(c, acc[i][j]);
      } else {
        float v = acc[i][j];
        if (p.bias) v += p.bias[m];
        *c = apply_act(v, p.act, p.act_param);
      }
    }
  }
}

// second pass of a split-K forward: C = act(sum_z partial[z] + bias[m]), slices added in order
__global__ void rows_bias_act_kernel(RowsGemm p, int ks) {
  const long long total = (long long)p.M * p.N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / p.N), n = (int)(i % p.N);
    float v = 0.f;
    for (int z = 0; z < ks; ++z) v += p.partial[(size_t)z * total + i];
    if (p.bias) v += p.bias[m];
    p.C[p.cm.off(m) + p.cn.off(n)] = apply_act(v, p.act, p.act_param);
  }
}

// Split-K factor: these GEMMs are short and wide in K (Cin*75 up to 11,250) with few output tiles; one CTA per tile would walk hundreds of
// latency-bound K steps on a mostly idle chip (measured 0.32 ms for the SAUnet:L conv3 forward), so K is cut until every SM holds
// several CTAs; the slices meet in fp32 atomics.
static int split_k_for(const RowsGemm& p, int tm, int tn) {
  const long long tiles = (long long)ceil_div(p.N, tn) * ceil_div(p.M, tm);
  int ks = 1;
  while (tiles * ks < 148 * 4 && p.K / (ks * 2) >= 128 && ks < 64) ks *= 2;
  return ks;
}

// 64 x 64 tiles (4 x 4 outputs per thread: 8 shared-memory loads per 16 FMAs) unless the GEMM has very few rows; a chip-filling number
// of CTAs comes from splitting K, not from smaller tiles (32 x 32 tiles measured shared-memory-bound: 4.4 TFLOP/s vs 7 for 64 x 64)
static void tile_for(const RowsGemm& p, int& tm, int& tn) {
  tm = tn = 64;
  if (p.M <= 16) tm = 16;
}

// c_bytes: size of the (dense) output buffer, zeroed when the K slices meet in atomics (gradients).  The forward (p.partial set by the
// caller from its workspace) keeps the slices apart and adds them in order together with bias and activation.
static int launch_rows_gemm(RowsGemm& p, size_t c_bytes, cudaStream_t st) {
  int tm, tn;
  tile_for(p, tm, tn);
  const int ks = split_k_for(p, tm, tn);
  if (ks == 1) p.partial = nullptr;
  if (ks > 1 && !p.partial) cudaMemsetAsync(p.C, 0, c_bytes, st);
  if (tm == 16)
    rows_gemm_kernel<16, 64><<<dim3(ceil_div(p.N, 64), ceil_div(p.M, 16), ks), 256, 0, st>>>(p);
  else
    rows_gemm_kernel<64, 64><<<dim3(ceil_div(p.N, 64), ceil_div(p.M, 64), ks), 256, 0, st>>>(p);
  if (ks > 1 && p.partial) {
    const long long total = (long long)p.M * p.N;
    rows_bias_act_kernel<<<ceil_div(total, 256) > 1184 ? 1184 : ceil_div(total, 256), 256, 0, st>>>(p, ks);
    count_launch();
  }
  return MPA_OK;
}

static inline Idx2 flat(long long s) { return Idx2{1 << 30, 0, s}; }

}  // namespace mpa

using namespace mpa;

extern "C" {

/* x [B][Cin][H][W], w [Cout][Cin][H][1] (state_dict layout), out [B][Cout][1][W] = act(conv + bias) */
size_t mpa_conv_rows_fwd_workspace(int B, int Cin, int H, int W, int Cout) {
  RowsGemm p{};
  p.M = Cout; p.N = B * W; p.K = Cin * H;
  int tm, tn;
  tile_for(p, tm, tn);
  const int ks = split_k_for(p, tm, tn);
  return ks > 1 ? (size_t)ks * p.M * p.N * sizeof(float) : 0;
}

int mpa_conv_rows_fwd_f32(const float* x, const float* w, const float* bias, float* out, int B, int Cin, int H, int W, int Cout, int act,
                          float act_param, void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w && out && B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv_rows_fwd: bad argument");
  const size_t need = mpa_conv_rows_fwd_workspace(B, Cin, H, W, Cout);
  if (need > 0 && (!workspace || ws_bytes < need)) {
    set_error("conv_rows_fwd: workspace %zu < %zu bytes", ws_bytes, need);
    return MPA_ERR_WORKSPACE;
  }
  RowsGemm p{};
  p.partial = need > 0 ? (float*)workspace : nullptr;
  const int K = Cin * H;
  p.A = w; p.B = x; p.C = out; p.bias = bias; p.M = Cout; p.N = B * W; p.K = K;
  p.am = flat(K); p.ak = flat(1); p.a_k_fast = 1;
  p.bk = flat(W); p.bn = Idx2{W, (long long)K * W, 1}; p.b_k_fast = 0;
  p.cm = flat(W); p.cn = Idx2{W, (long long)Cout * W, 1};
  p.act = act; p.act_param = act_param;
  launch_rows_gemm(p, sizeof(float) * (size_t)B * Cout * W, (cudaStream_t)stream);
  MPA_CHECK_LAUNCH("conv_rows_fwd");
  return MPA_OK;
}

/* g_in [B][Cin][H][W] = sum_co w[co][ci][h] * g_out[b][co][0][w] */
int mpa_conv_rows_dgrad_f32(const float* g_out, const float* w, float* g_in, int B, int Cin, int H, int W, int Cout, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(g_out && w && g_in && B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv_rows_dgrad: bad argument");
  RowsGemm p{};
  const int K = Cin * H;
  p.A = w; p.B = g_out; p.C = g_in; p.bias = nullptr; p.M = K; p.N = B * W; p.K = Cout;
  p.am = flat(1); p.ak = flat(K); p.a_k_fast = 0;
  p.bk = flat(W); p.bn = Idx2{W, (long long)Cout * W, 1}; p.b_k_fast = 0;
  p.cm = flat(W); p.cn = Idx2{W, (long long)K * W, 1};
  p.act = MPA_ACT_NONE; p.act_param = 0.f;
  launch_rows_gemm(p, sizeof(float) * (size_t)B * K * W, (cudaStream_t)stream);
  MPA_CHECK_LAUNCH("conv_rows_dgrad");
  return MPA_OK;
}

/* g_w [Cout][Cin][H] (overwritten) = sum_{b,w} g_out[b][co][0][w] * x[b][ci][h][w] */
int mpa_conv_rows_wgrad_f32(const float* x, const float* g_out, float* g_w, int B, int Cin, int H, int W, int Cout, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && g_out && g_w && B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0, "conv_rows_wgrad: bad argument");
  RowsGemm p{};
  const int K = Cin * H;
  p.A = g_out; p.B = x; p.C = g_w; p.bias = nullptr; p.M = Cout; p.N = K; p.K = B * W;
  p.am = flat(W); p.ak = Idx2{W, (long long)Cout * W, 1}; p.a_k_fast = 1;
  p.bn = flat(W); p.bk = Idx2{W, (long long)K * W, 1}; p.b_k_fast = 1;
  p.cm = flat(K); p.cn = flat(1);
  p.act = MPA_ACT_NONE; p.act_param = 0.f;
  launch_rows_gemm(p, sizeof(float) * (size_t)Cout * K, (cudaStream_t)stream);
  MPA_CHECK_LAUNCH("conv_rows_wgrad");
  return MPA_OK;
}

}  // extern "C"
