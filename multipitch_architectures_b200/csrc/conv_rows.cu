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

// ---- thin layers (Cout <= 10: the conv3 of CNN:XS 20 -> 10 and of the DRCNN 30 -> 10, the 1 x 1 convolutions of conv4): the products
// above degenerate to a few matrix-vector products on the [K][W] matrix x[b] and are HBM-bound — read (or write) the 4*B*K*W bytes once
// with 16-byte accesses instead of tiling a GEMM whose M (or K) is a fraction of a tile.
constexpr int kThinMaxCout = 10;
constexpr size_t kThinScratchFloats = 1u << 22;      // weight-gradient partial sums [slices][Cout][K][W/4]

static float* thin_scratch() {
  static float* ptr[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!ptr[dev]) {
    float* q = nullptr;
    if (cudaMalloc(&q, kThinScratchFloats * sizeof(float)) != cudaSuccess) return nullptr;
    ptr[dev] = q;
  }
  return ptr[dev];
}

static unsigned* thin_counters() {
  static unsigned* ptr[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!ptr[dev]) {
    unsigned* q = nullptr;
    if (cudaMalloc(&q, 65536 * sizeof(unsigned)) != cudaSuccess || cudaMemset(q, 0, 65536 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    ptr[dev] = q;
  }
  return ptr[dev];
}

// forward: CTA (item b, K slice ks); thread (row group g, bin quad f4) adds rows g, g+G, ... of its slice of x[b]; the G groups meet in
// shared memory in order.  One CTA per item left 1.7 CTAs per SM at batch 256 (ncu: 15 % of the issue slots, 0.97 TB/s, latency-bound);
// with KS slices the partial sums go to a scratch buffer and the slice that arrives last at the item's counter adds them IN ORDER
// (deterministic) and applies bias + activation.
template <int CO>
__global__ void __launch_bounds__(256) conv_rows_thin_fwd_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ out, int K, int W4, int act,
                                                                 float act_param, float4* __restrict__ partial, unsigned* __restrict__ counters) {
  extern __shared__ float4 red[];                 // [G][CO][W4]
  __shared__ bool last;
  const int G = 256 / W4;
  const int f4 = threadIdx.x % W4, g = threadIdx.x / W4;
  const int b = blockIdx.x, ks = blockIdx.y, KS = gridDim.y, B = gridDim.x;
  const int k0 = (int)((long long)K * ks / KS), k1 = (int)((long long)K * (ks + 1) / KS);
  if (g < G) {
    float4 acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* xb = x + (size_t)b * K * W4 + f4;
#pragma unroll 4
    for (int r = k0 + g; r < k1; r += G) {
      const float4 v = xb[(size_t)r * W4];
#pragma unroll
      for (int co = 0; co < CO; ++co) {
        const float wv = w[(size_t)co * K + r];
        acc[co].x = fmaf(v.x, wv, acc[co].x); acc[co].y = fmaf(v.y, wv, acc[co].y);
        acc[co].z = fmaf(v.z, wv, acc[co].z); acc[co].w = fmaf(v.w, wv, acc[co].w);
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) red[((size_t)g * CO + co) * W4 + f4] = acc[co];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < CO * W4; e += 256) {
    const int co = e / W4, q = e - co * W4;
    float4 t = red[(size_t)co * W4 + q];
    for (int gg = 1; gg < G; ++gg) {
      const float4 u = red[((size_t)gg * CO + co) * W4 + q];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    if (KS > 1) {
      partial[((size_t)ks * B + b) * CO * W4 + e] = t;
      continue;
    }
    const float bv = bias ? bias[co] : 0.f;
    t.x = apply_act(t.x + bv, act, act_param); t.y = apply_act(t.y + bv, act, act_param);
    t.z = apply_act(t.z + bv, act, act_param); t.w = apply_act(t.w + bv, act, act_param);
    reinterpret_cast<float4*>(out)[((size_t)b * CO + co) * W4 + q] = t;
  }
  if (KS == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&counters[b], 1u) == (unsigned)KS - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int e = threadIdx.x; e < CO * W4; e += 256) {
    const int co = e / W4;
    float4 t = __ldcg(&partial[(size_t)b * CO * W4 + e]);
    for (int s2 = 1; s2 < KS; ++s2) {
      const float4 u = __ldcg(&partial[((size_t)s2 * B + b) * CO * W4 + e]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const float bv = bias ? bias[co] : 0.f;
    t.x = apply_act(t.x + bv, act, act_param); t.y = apply_act(t.y + bv, act, act_param);
    t.z = apply_act(t.z + bv, act, act_param); t.w = apply_act(t.w + bv, act, act_param);
    reinterpret_cast<float4*>(out)[(size_t)b * CO * W4 + e] = t;
  }
  if (threadIdx.x == 0) counters[b] = 0u;
}

// data gradient: g_in[b][r][:] = sum_co w[co][r] * g_out[b][co][:]; one thread per 4 bins, write-bound
template <int CO>
__global__ void __launch_bounds__(256) conv_rows_thin_dgrad_kernel(const float4* __restrict__ go, const float* __restrict__ w, float4* __restrict__ gi,
                                                                   long long total, int K, int W4) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int f4 = (int)(i % W4);
    const long long br = i / W4;
    const int r = (int)(br % K);
    const long long b = br / K;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      const float4 gv = go[(b * CO + co) * W4 + f4];
      const float wv = w[(size_t)co * K + r];
      t.x = fmaf(gv.x, wv, t.x); t.y = fmaf(gv.y, wv, t.y); t.z = fmaf(gv.z, wv, t.z); t.w = fmaf(gv.w, wv, t.w);
    }
    gi[i] = t;
  }
}

// weight gradient, pass 1: thread j = (row r, bin quad f4) of the [K][W4] matrix adds its items of slice s; partial[s][co][j]
template <int CO>
__global__ void __launch_bounds__(256) conv_rows_thin_wgrad_kernel(const float4* __restrict__ x, const float4* __restrict__ go,
                                                                   float* __restrict__ partial, int B, int K, int W4, int S) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  const int KW4 = K * W4;
  if (j >= KW4) return;
  const int f4 = j % W4, s = blockIdx.y;
  const int b0 = (int)((long long)B * s / S), b1 = (int)((long long)B * (s + 1) / S);
  float4 acc[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) acc[co] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float4 v = x[(size_t)b * KW4 + j];
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      const float4 gv = go[((size_t)b * CO + co) * W4 + f4];
      acc[co].x = fmaf(v.x, gv.x, acc[co].x); acc[co].y = fmaf(v.y, gv.y, acc[co].y);
      acc[co].z = fmaf(v.z, gv.z, acc[co].z); acc[co].w = fmaf(v.w, gv.w, acc[co].w);
    }
  }
#pragma unroll
  for (int co = 0; co < CO; ++co) partial[((size_t)s * CO + co) * KW4 + j] = (acc[co].x + acc[co].y) + (acc[co].z + acc[co].w);
}
// pass 2: g_w[co][r] = sum over slices and bin quads; one warp per output, lane l adds terms l, l+32, ... in order, then a shuffle tree
__global__ void __launch_bounds__(128) conv_rows_thin_wgrad_final_kernel(const float* __restrict__ partial, float* __restrict__ gw, int CO, int K,
                                                                         int W4, int S) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= CO * K) return;
  const int co = i / K, r = i - co * K;
  float t = 0.f;
  for (int e = lane; e < S * W4; e += 32) {
    const int sl = e / W4, q = e - sl * W4;
    t += partial[((size_t)sl * CO + co) * K * W4 + (size_t)r * W4 + q];
  }
  t = warp_sum(t);
  if (lane == 0) gw[i] = t;
}

static inline bool thin_ok(int Cout, int W) {
  return Cout <= kThinMaxCout && W % 4 == 0 && W / 4 <= 128 && sizeof(float4) * (size_t)(256 / (W / 4)) * Cout * (W / 4) <= 48 * 1024;
}
static inline bool aligned16(const void* q) { return ((uintptr_t)q & 15) == 0; }

#define THIN_CASE(N, CALL) case N: { constexpr int C_ = N; CALL; } break;
#define THIN_DISPATCH(CO, CALL)                                                                                             \
  switch (CO) {                                                                                                             \
    THIN_CASE(1, CALL) THIN_CASE(2, CALL) THIN_CASE(3, CALL) THIN_CASE(4, CALL) THIN_CASE(5, CALL) THIN_CASE(6, CALL)        \
    THIN_CASE(7, CALL) THIN_CASE(8, CALL) THIN_CASE(9, CALL) default: { constexpr int C_ = 10; CALL; } break;               \
  }

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
  if (thin_ok(Cout, W) && aligned16(x) && aligned16(out)) {
    const int K = Cin * H, W4 = W / 4, G = 256 / W4;
    const size_t smem = sizeof(float4) * (size_t)G * Cout * W4;
    // K slices per item: about 8 CTAs per SM in total, at least ~8 rows per row group and slice, partial sums inside the scratch buffer
    int KS = 1;
    while (KS < 16 && (long long)B * KS * 2 <= 148 * 8 && K / (KS * 2) >= 8 * G && (size_t)(KS * 2) * B * Cout * W <= kThinScratchFloats && B <= 65536) KS *= 2;
    float4* partial = KS > 1 ? (float4*)thin_scratch() : nullptr;
    unsigned* counters = KS > 1 ? thin_counters() : nullptr;
    if (KS > 1 && (!partial || !counters)) KS = 1;
    THIN_DISPATCH(Cout, (conv_rows_thin_fwd_kernel<C_><<<dim3(B, KS), 256, smem, (cudaStream_t)stream>>>((const float4*)x, w, bias, out, K, W4, act,
                                                                                                       act_param, partial, counters)));
    MPA_CHECK_LAUNCH("conv_rows_fwd(thin)");
    return MPA_OK;
  }
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
  if (thin_ok(Cout, W) && aligned16(g_out) && aligned16(g_in)) {
    const int K = Cin * H, W4 = W / 4;
    const long long total = (long long)B * K * W4;
    const int grid = ceil_div(total, 256) > 148 * 16 ? 148 * 16 : ceil_div(total, 256);
    THIN_DISPATCH(Cout, (conv_rows_thin_dgrad_kernel<C_><<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)g_out, w, (float4*)g_in, total, K, W4)));
    MPA_CHECK_LAUNCH("conv_rows_dgrad(thin)");
    return MPA_OK;
  }
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
  if (thin_ok(Cout, W) && aligned16(x) && aligned16(g_out)) {
    const int K = Cin * H, W4 = W / 4;
    const int blocks = ceil_div((long long)K * W4, 256);
    int S = ceil_div(148 * 8, blocks);
    S = S > B ? B : S;
    while (S > 1 && (size_t)S * Cout * K * W4 > kThinScratchFloats) --S;
    float* scratch = (size_t)S * Cout * K * W4 <= kThinScratchFloats ? thin_scratch() : nullptr;
    if (scratch) {
      THIN_DISPATCH(Cout, (conv_rows_thin_wgrad_kernel<C_><<<dim3(blocks, S), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (const float4*)g_out,
                                                                                                         scratch, B, K, W4, S)));
      conv_rows_thin_wgrad_final_kernel<<<ceil_div((long long)Cout * K, 4), 128, 0, (cudaStream_t)stream>>>(scratch, g_w, Cout, K, W4, S);
      count_launch();
      MPA_CHECK_LAUNCH("conv_rows_wgrad(thin)");
      return MPA_OK;
    }
  }
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
