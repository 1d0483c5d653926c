// Element-wise training stages of the CNN blocks directly on the 16-bit CP8 planes the tensor-core convolutions read and write:
//   forward   z  = Dropout(MaxPool((3,1), stride 1, pad (1,0))(a))                 (basic_cnns.py:172-175, 376-377)
//   backward  ga = act'(a) * sum_t [first arg-max of window t == .] * mask * g[t]
//   bias gradient = per-channel sum of ga
// so that no nchw<->CP8 converter runs between two convolutions of a training step.  The planes hold the same 16-bit values the fp32
// NCHW kernels of elementwise.cu / backward.cu see behind the converters, every sum is formed in fp32 in the same order and rounded
// once on the store, and the dropout mask uses the SAME convention (Philox counter i/4, lane i%4 of the NCHW element index
// i = ((b*C + c)*T + t)*F + f) — outputs are bit-identical to the converter path (tests/test_gpu_training.py).
// One thread owns the 8 channels of a pixel; the 4 threads of a quad (f = 4q..4q+3) share 8 Philox draws (2 each, exchanged by shuffles).
// HBM-bound: forward reads a once (rows re-read from L1/L2) and writes z; backward reads a and g once and writes ga.
#include "common.cuh"

namespace mpa {

int channel_sum_cp8_launch(const uint4* g, float* out, int B, int C, int T, int F, int TP, int P, int pf, int pt, int ncs, int fmt,
                           cudaStream_t st);   // reduce.cu

template <int FMT>
__device__ __forceinline__ void unpack8(const uint4& u, float v[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (FMT == MPA_FMT_BF16) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
    } else {
      const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
      v[2 * e] = f2.x;
      v[2 * e + 1] = f2.y;
    }
  }
}
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float v[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (FMT == MPA_FMT_BF16) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ uint32_t comp4(const uint4& r, int j) { return j == 0 ? r.x : j == 1 ? r.y : j == 2 ? r.z : r.w; }

// Dropout keep factors of channels c0..c0+7 at pixel (t, f) for the 4 lanes of a quad (lane j holds f = 4q + j; all 32 lanes of the warp
// must call).  i_quad = NCHW element index of channel c0 at (t, 4q) — a multiple of 4 because F % 4 == 0; plane = T*F.
__device__ __forceinline__ void quad_dropout8(float fac[8], long long i_quad, long long plane, int j, float p, float scale,
                                              unsigned long long seed, unsigned long long offset) {
  uint32_t v[4][2];                                   // v[r][e]: bits of channel 2*(j^r)+e at my column
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const uint4 mine = dropout_bits((i_quad + (long long)(2 * j + e) * plane) >> 2, seed, offset);   // channel 2j+e, columns 4q..4q+3
    v[0][e] = comp4(mine, j);
#pragma unroll
    for (int r = 1; r < 4; ++r) v[r][e] = __shfl_xor_sync(0xffffffffu, comp4(mine, j ^ r), r);
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int r = s ^ j;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const uint32_t bits = r == 0 ? v[0][e] : r == 1 ? v[1][e] : r == 2 ? v[2][e] : v[3][e];
      fac[2 * s + e] = dropout_factor(bits, p, scale);
    }
  }
}

constexpr int kCp8Threads = 128;

// grid (B * NCk * T, ceil(F / 128)); thread = one pixel of one chunk plane
template <int FMT>
__global__ void __launch_bounds__(kCp8Threads) pool3_dropout_cp8_kernel(const uint4* __restrict__ a, uint4* __restrict__ out, int C, int NCk, int T,
                                                                        int F, int TP, int P, int pf, int pt, DropoutArgs d) {
  const int f = blockIdx.y * kCp8Threads + threadIdx.x;
  int r = blockIdx.x;
  const int t = r % T;
  r /= T;
  const int ck = r % NCk;
  const int b = r / NCk;
  const bool ok = f < F;
  const size_t base = (((size_t)b * NCk + ck) * TP + pt + t) * P + pf + (ok ? f : 0);
  float m[8];
  if (ok) {
    uint4 c = a[base];
    __nv_bfloat162* cm = reinterpret_cast<__nv_bfloat162*>(&c);
    __half2* ch = reinterpret_cast<__half2*>(&c);
    if (t > 0) {
      const uint4 u = a[base - P];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (FMT == MPA_FMT_BF16) cm[e] = __hmax2(cm[e], reinterpret_cast<const __nv_bfloat162*>(&u)[e]);
        else ch[e] = __hmax2(ch[e], reinterpret_cast<const __half2*>(&u)[e]);
      }
    }
    if (t < T - 1) {
      const uint4 u = a[base + P];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (FMT == MPA_FMT_BF16) cm[e] = __hmax2(cm[e], reinterpret_cast<const __nv_bfloat162*>(&u)[e]);
        else ch[e] = __hmax2(ch[e], reinterpret_cast<const __half2*>(&u)[e]);
      }
    }
    if (d.p <= 0.f) {
      out[base] = c;
      return;
    }
    unpack8<FMT>(c, m);
  }
  if (d.p <= 0.f) return;
  // dropout: every lane of the warp takes part in the shuffles (F % 4 == 0: quads are entirely inside or outside the row)
  float fac[8];
  const long long plane = (long long)T * F;
  const long long i_quad = (((long long)b * C + ck * 8) * T + t) * F + (f & ~3);
  quad_dropout8(fac, i_quad, plane, f & 3, d.p, 1.f / (1.f - d.p), d.seed, dropout_offset(d));
  if (!ok) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = __fmul_rn(m[e], fac[e]);
  out[base] = pack8<FMT>(m);
}

// grid (B * NCk, ceil(F / 128), row segments); thread = one column of one chunk plane, walking down its segment of the T rows with the
// window's 3 activation rows and 3 pending gradient sums (x 8 channels) in registers — the register version of maxpool_time_bwd_col_kernel<3>
template <int FMT>
__global__ void __launch_bounds__(kCp8Threads) pool3_bwd_dropout_cp8_kernel(const uint4* __restrict__ a, const uint4* __restrict__ g_p,
                                                                            uint4* __restrict__ g_a, int C, int NCk, int T, int F, int TP, int P,
                                                                            int pf, int pt, int act, float act_param, DropoutArgs d) {
  const int f = blockIdx.y * kCp8Threads + threadIdx.x;
  const int ck = blockIdx.x % NCk, b = blockIdx.x / NCk;
  const bool ok = f < F;
  // output rows [r0, r1) of this segment; windows r0-1 .. r1 contribute to them
  const int seg = (T + (int)gridDim.z - 1) / (int)gridDim.z;
  const int r0 = blockIdx.z * seg, r1 = min(T, r0 + seg);
  if (r0 >= r1) return;
  const size_t base = (((size_t)b * NCk + ck) * TP + pt) * P + pf + (ok ? f : 0);
  const bool drop = d.p > 0.f;
  const unsigned long long d_off = drop ? dropout_offset(d) : 0ull;
  const float d_scale = drop ? 1.f / (1.f - d.p) : 1.f;
  const long long plane = (long long)T * F;
  const long long i_col = ((long long)b * C + ck * 8) * plane + (f & ~3);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  float win[3][8], acc[3][8];      // win[j] = activation row w - 1 + j of window w; acc[j] = pending gradient of that row
  bool valid[3];
  // first window of the segment: w = r0 - 1 (only its contribution to row r0 matters; it exists when r0 >= 1)
  int w = r0 >= 1 ? r0 - 1 : 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int row = w - 1 + j;
    valid[j] = row >= 0 && row < T;
    unpack8<FMT>((ok && valid[j]) ? a[base + (size_t)row * P] : zero4, win[j]);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  }
  const int w_last = min(T - 1, r1);
  for (; w <= w_last; ++w) {
    float g[8];
    unpack8<FMT>(ok ? g_p[base + (size_t)w * P] : zero4, g);
    if (drop) {
      float fac[8];
      quad_dropout8(fac, i_col + (long long)w * F, plane, f & 3, d.p, d_scale, d.seed, d_off);
#pragma unroll
      for (int e = 0; e < 8; ++e) g[e] = __fmul_rn(g[e], fac[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      // strict '>': the first maximum of the window wins (ATen semantics); rows outside [0, T) never win
      int am = valid[0] ? 0 : 1;
      float best = valid[0] ? win[0][e] : win[1][e];
      if (valid[0] && win[1][e] > best) { best = win[1][e]; am = 1; }
      if (valid[2] && win[2][e] > best) { am = 2; }
      acc[0][e] += am == 0 ? g[e] : 0.f;
      acc[1][e] += am == 1 ? g[e] : 0.f;
      acc[2][e] += am == 2 ? g[e] : 0.f;
    }
    // row w - 1 has now seen its three windows (w - 2, w - 1, w)
    const int done = w - 1;
    if (ok && done >= r0 && done < r1) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float dd = 1.f;
        if (act == MPA_ACT_LRELU) dd = win[0][e] >= 0.f ? 1.f : act_param;
        else if (act == MPA_ACT_RELU) dd = win[0][e] > 0.f ? 1.f : 0.f;
        o[e] = acc[0][e] * dd;
      }
      g_a[base + (size_t)done * P] = pack8<FMT>(o);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      win[0][e] = win[1][e]; win[1][e] = win[2][e];
      acc[0][e] = acc[1][e]; acc[1][e] = acc[2][e]; acc[2][e] = 0.f;
    }
    valid[0] = valid[1]; valid[1] = valid[2];
    const int nrow = w + 2;
    valid[2] = nrow < T;
    unpack8<FMT>((ok && valid[2]) ? a[base + (size_t)nrow * P] : zero4, win[2]);
  }
  // the last row of the tensor has no window below it: it is final after window T - 1
  if (ok && r1 == T && w == T) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float dd = 1.f;
      if (act == MPA_ACT_LRELU) dd = win[0][e] >= 0.f ? 1.f : act_param;
      else if (act == MPA_ACT_RELU) dd = win[0][e] > 0.f ? 1.f : 0.f;
      o[e] = acc[0][e] * dd;
    }
    g_a[base + (size_t)(T - 1) * P] = pack8<FMT>(o);
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_pool3_dropout_cp8(const void* a_cp8, void* out_cp8, int B, int C, int T, int F, int pitch, int pf, int pt, int fmt, float p,
                          unsigned long long seed, unsigned long long offset, const long long* step_dev, unsigned long long step_mul,
                          void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(a_cp8 && out_cp8 && B > 0 && C > 0 && T > 0 && F > 0 && F % 4 == 0 && pitch >= pf + F && pt >= 0 && p >= 0.f && p < 1.f &&
                  (fmt == MPA_FMT_BF16 || fmt == MPA_FMT_F16),
              "pool3_dropout_cp8: bad argument (F must be a multiple of 4)");
  const int NCk = (C + 7) / 8;
  const dim3 grid(B * NCk * T, ceil_div(F, kCp8Threads));
  const DropoutArgs d{p, seed, offset, step_dev, step_mul};
  if (fmt == MPA_FMT_BF16)
    pool3_dropout_cp8_kernel<MPA_FMT_BF16><<<grid, kCp8Threads, 0, (cudaStream_t)stream>>>((const uint4*)a_cp8, (uint4*)out_cp8, C, NCk, T, F,
                                                                                          T + 2 * pt, pitch, pf, pt, d);
  else
    pool3_dropout_cp8_kernel<MPA_FMT_F16><<<grid, kCp8Threads, 0, (cudaStream_t)stream>>>((const uint4*)a_cp8, (uint4*)out_cp8, C, NCk, T, F,
                                                                                         T + 2 * pt, pitch, pf, pt, d);
  MPA_CHECK_LAUNCH("pool3_dropout_cp8");
  return MPA_OK;
}

int mpa_pool3_bwd_dropout_cp8(const void* a_cp8, const void* g_out_cp8, void* g_a_cp8, int B, int C, int T, int F, int pitch, int pf, int pt,
                              int fmt, int act, float act_param, float p, unsigned long long seed, unsigned long long offset,
                              const long long* step_dev, unsigned long long step_mul, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(a_cp8 && g_out_cp8 && g_a_cp8 && B > 0 && C > 0 && T > 0 && F > 0 && F % 4 == 0 && pitch >= pf + F && pt >= 0 && p >= 0.f &&
                  p < 1.f && (fmt == MPA_FMT_BF16 || fmt == MPA_FMT_F16),
              "pool3_bwd_dropout_cp8: bad argument (F must be a multiple of 4)");
  const int NCk = (C + 7) / 8;
  // enough CTAs to fill the GPU: split the T rows into segments when there are few planes (each segment re-reads 2 halo rows)
  const int cols = ceil_div(F, kCp8Threads);
  int segs = ceil_div(148 * 16, (long long)B * NCk * cols);
  segs = segs < 1 ? 1 : (segs > ceil_div(T, 8) ? ceil_div(T, 8) : segs);
  const dim3 grid(B * NCk, cols, segs);
  const DropoutArgs d{p, seed, offset, step_dev, step_mul};
  if (fmt == MPA_FMT_BF16)
    pool3_bwd_dropout_cp8_kernel<MPA_FMT_BF16><<<grid, kCp8Threads, 0, (cudaStream_t)stream>>>(
        (const uint4*)a_cp8, (const uint4*)g_out_cp8, (uint4*)g_a_cp8, C, NCk, T, F, T + 2 * pt, pitch, pf, pt, act, act_param, d);
  else
    pool3_bwd_dropout_cp8_kernel<MPA_FMT_F16><<<grid, kCp8Threads, 0, (cudaStream_t)stream>>>(
        (const uint4*)a_cp8, (const uint4*)g_out_cp8, (uint4*)g_a_cp8, C, NCk, T, F, T + 2 * pt, pitch, pf, pt, act, act_param, d);
  MPA_CHECK_LAUNCH("pool3_bwd_dropout_cp8");
  return MPA_OK;
}

int mpa_channel_sum_cp8(const void* g_cp8, float* out, int B, int C, int T, int F, int pitch, int pf, int pt, int fmt, int ncs, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(g_cp8 && out && B > 0 && C > 0 && T > 0 && F > 0 && pitch >= pf + F && pt >= 0 && (fmt == MPA_FMT_BF16 || fmt == MPA_FMT_F16),
              "channel_sum_cp8: bad argument");
  if (ncs <= 0) ncs = (C + 7) / 8;
  int rc = channel_sum_cp8_launch((const uint4*)g_cp8, out, B, C, T, F, T + 2 * pt, pitch, pf, pt, ncs, fmt, (cudaStream_t)stream);
  if (rc != MPA_OK) return rc;
  MPA_CHECK_LAUNCH("channel_sum_cp8");
  return MPA_OK;
}

}  // extern "C"
