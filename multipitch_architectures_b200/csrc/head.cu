// Fused tail of the network head (identical in every model, e.g. basic_cnns.py:396-408):
//   conv3 (T x 1 "time reduction", C1 -> C2) + LeakyReLU -> conv4.0 (1x1, C2 -> C3) + LeakyReLU
//   -> conv4.3 (1x1, C3 -> 1) -> sigmoid
// for the patch-wise case T == kernel height (one output frame per patch).  Input is the max-pooled conv2 output in the
// compact 16-bit chunk layout [B][ceil(C1/8)][T][Fo][8] written by the tcgen05 convolution (out_mode 1); output
// [B, Fo] fp32.  One thread owns one (patch, bin) and keeps all C2 conv3 accumulators in registers, so conv4.* never
// touch memory.  Each load is one 16-byte pixel-chunk (8 channels); the conv3 weights of (chunk, 15-frame block) are
// staged in shared memory as [t][ci][co] and read as broadcast float4.
#include "common.cuh"
#include <cuda_fp16.h>

namespace mpa {

constexpr int kC2Max = 32, kC3Max = 16, kTBlk = 15;

template <int FMT>
__device__ __forceinline__ float cvt_in(uint16_t v) {
  return FMT == MPA_FMT_BF16 ? __bfloat162float(__ushort_as_bfloat16(v)) : __half2float(__ushort_as_half(v));
}

template <int FMT>
__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
  if (FMT == MPA_FMT_BF16) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __bfloat1622float2(h[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  } else {
    const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f = __half22float2(h[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) head_tail_kernel(const uint4* __restrict__ x, const float* __restrict__ w3, const float* __restrict__ b3,
                                                        const float* __restrict__ w40, const float* __restrict__ b40, const float* __restrict__ w43,
                                                        const float* __restrict__ b43, float* __restrict__ out, int B, int C1, int T, int Fo, int C2,
                                                        int C3, int pb, float a) {
  extern __shared__ float sm[];
  float* w3s = sm;                                  // [kTBlk][8][kC2Max]
  float* w40s = sm + kTBlk * 8 * kC2Max;            // [C3][kC2Max]
  float* misc = w40s + kC3Max * kC2Max;             // b3[32] b40[16] w43[16] b43[1]
  const int tid = threadIdx.x;
  const int bl = tid / Fo, f = tid - bl * Fo;
  const int b = blockIdx.x * pb + bl;
  const bool active = (bl < pb) && (b < B);
  const int NC1 = (C1 + 7) / 8;
  for (int e = tid; e < kC3Max * kC2Max; e += blockDim.x) {
    int c3 = e / kC2Max, co = e - c3 * kC2Max;
    w40s[e] = (c3 < C3 && co < C2) ? w40[c3 * C2 + co] : 0.f;
  }
  for (int e = tid; e < 65; e += blockDim.x) {
    float v = 0.f;
    if (e < 32) v = e < C2 ? b3[e] : 0.f;
    else if (e < 48) v = (e - 32) < C3 ? b40[e - 32] : 0.f;
    else if (e < 64) v = (e - 48) < C3 ? w43[e - 48] : 0.f;
    else v = b43[0];
    misc[e] = v;
  }
  float acc[kC2Max];
#pragma unroll
  for (int i = 0; i < kC2Max; ++i) acc[i] = 0.f;
  const uint4* xb = x + ((size_t)(active ? b : 0) * NC1) * T * Fo + f;
  for (int ck = 0; ck < NC1; ++ck) {
    for (int tb = 0; tb < T; tb += kTBlk) {
      const int nt = min(kTBlk, T - tb);
      __syncthreads();
      for (int e = tid; e < nt * 8 * kC2Max; e += blockDim.x) {
        const int co = e % kC2Max;
        const int ci8 = (e / kC2Max) % 8;
        const int tt = e / (kC2Max * 8);
        const int ci = ck * 8 + ci8;
        w3s[e] = (co < C2 && ci < C1) ? w3[((size_t)co * C1 + ci) * T + tb + tt] : 0.f;
      }
      __syncthreads();
      if (active) {
        const uint4* xp = xb + ((size_t)ck * T + tb) * Fo;
        for (int tt = 0; tt < nt; ++tt) {
          float v[8];
          unpack8<FMT>(xp[(size_t)tt * Fo], v);
#pragma unroll
          for (int e8 = 0; e8 < 8; ++e8) {
            const float4* wr = reinterpret_cast<const float4*>(w3s + (tt * 8 + e8) * kC2Max);
#pragma unroll
            for (int q = 0; q < kC2Max / 4; ++q) {
              const float4 w = wr[q];
              acc[4 * q + 0] = fmaf(v[e8], w.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(v[e8], w.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(v[e8], w.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(v[e8], w.w, acc[4 * q + 3]);
            }
          }
        }
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < kC2Max; ++i) {
    float h = acc[i] + misc[i];
    acc[i] = h >= 0.f ? h : a * h;
  }
  float o = misc[64];
  for (int c3 = 0; c3 < C3; ++c3) {
    float g = misc[32 + c3];
#pragma unroll
    for (int i = 0; i < kC2Max; ++i) g = fmaf(w40s[c3 * kC2Max + i], acc[i], g);
    g = g >= 0.f ? g : a * g;
    o = fmaf(misc[48 + c3], g, o);
  }
  out[(size_t)b * Fo + f] = 1.f / (1.f + expf(-o));
}

// conv4.0 (1x1, C2 -> C3) + LeakyReLU + conv4.3 (1x1, C3 -> 1) + sigmoid on the (already activated) conv3 output held in compact
// 16-bit planes [B][ceil(C2/8)][R][Fo][8]; any C2 / C3; one CTA per (patch, row), hidden vector staged in shared memory.
template <int FMT>
__global__ void __launch_bounds__(128) head_tail2_kernel(const uint16_t* __restrict__ h, const float* __restrict__ w40, const float* __restrict__ b40,
                                                         const float* __restrict__ w43, const float* __restrict__ b43, float* __restrict__ out, int R,
                                                         int Fo, int C2, int C3, float a, int x3) {
  extern __shared__ float sm[];
  float* hs = sm;                          // [C2][Fo]
  float* ws = sm + (size_t)C2 * Fo;        // [C3][C2]
  const int b = blockIdx.x / R, r = blockIdx.x % R;
  const int NC2 = (C2 + 7) / 8;
  const int ncs = x3 ? 2 * NC2 : NC2;      // split precision: hi planes, then lo planes
  for (int e = threadIdx.x; e < NC2 * Fo * 8; e += blockDim.x) {
    const int c8 = e & 7, f = (e >> 3) % Fo, ck = e / (8 * Fo);
    const int co = ck * 8 + c8;
    if (co < C2) {
      float v = cvt_in<FMT>(h[((((size_t)b * ncs + ck) * R + r) * Fo + f) * 8 + c8]);
      if (x3) v += cvt_in<FMT>(h[((((size_t)b * ncs + NC2 + ck) * R + r) * Fo + f) * 8 + c8]);
      hs[co * Fo + f] = v;
    }
  }
  for (int e = threadIdx.x; e < C3 * C2; e += blockDim.x) ws[e] = w40[e];
  __syncthreads();
  for (int f = threadIdx.x; f < Fo; f += blockDim.x) {
    float o = b43[0];
    for (int c3 = 0; c3 < C3; ++c3) {
      float g = b40[c3];
      const float* wr = ws + (size_t)c3 * C2;
      for (int co = 0; co < C2; ++co) g = fmaf(wr[co], hs[co * Fo + f], g);
      g = g >= 0.f ? g : a * g;
      o = fmaf(w43[c3], g, o);
    }
    out[((size_t)b * R + r) * Fo + f] = 1.f / (1.f + expf(-o));
  }
}


// ---- MaxPool((13,1)) + conv3 (75 x 1, C1 -> C2 <= 16) + LeakyReLU + conv4.0 + LeakyReLU + conv4.3 + sigmoid in ONE launch ----------------
// (basic_cnns.py:180-195 / :396-408 for the CNN family, whose conv3 is far too thin for the tensor cores: M = 10 of 128 lanes.)
// The three launches this replaces moved the conv2 output through HBM three times (pool: read + write, conv3 as a tcgen05 convolution with
// one real output row per 16-row tile: read) and took 155 + 328 + 21 us per 646 DRCNN patches for 2 GFLOP.  Here a warp owns one input chunk
// (8 channels) of 32 consecutive (patch, bin) columns, a lane one column: it walks down the 75 frames ONCE with the 13-frame maximum coming
// from a doubling table in packed 16-bit registers (pairs -> fours -> eights -> two overlapping eights: 16 max instructions per frame for the
// 8 channels) and feeds each pooled frame to C2P fp32 accumulators; the conv3 weights of the warp's chunk are staged per 30-frame block in
// shared memory as [frame][channel][C2P] and read as broadcast float4.  The chunk warps meet in shared memory, warp 0 finishes the tail.
constexpr int kHpT = 75, kHpBlk = 30;        // 30 = lcm of the delay-line lengths 3, 5, 6: static register indices inside a block

template <int FMT>
__device__ __forceinline__ void hmax8(uint4& c, const uint4& a, const uint4& b) {
  const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
  const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
  uint32_t* pc = reinterpret_cast<uint32_t*>(&c);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (FMT == MPA_FMT_BF16) {
      const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&pa[e]), *reinterpret_cast<const __nv_bfloat162*>(&pb[e]));
      pc[e] = *reinterpret_cast<const uint32_t*>(&m);
    } else {
      const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&pa[e]), *reinterpret_cast<const __half2*>(&pb[e]));
      pc[e] = *reinterpret_cast<const uint32_t*>(&m);
    }
  }
}

template <int FMT, int C2P>
__global__ void __launch_bounds__(256) head_pool_conv3_tail_kernel(const uint4* __restrict__ y, const float* __restrict__ w3p,
                                                                   const float* __restrict__ b3, const float* __restrict__ w40,
                                                                   const float* __restrict__ b40, const float* __restrict__ w43,
                                                                   const float* __restrict__ b43, float* __restrict__ out, long long n_cols, int NC1,
                                                                   int Fo, int C2, int C3, float a) {
  extern __shared__ __align__(16) float hp_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;      // warp = input chunk
  float* w3s = hp_sm + (size_t)warp * kHpBlk * 8 * C2P;             // this warp's [30][8][C2P]
  float* red = hp_sm + (size_t)NC1 * kHpBlk * 8 * C2P;              // [NC1][32][C2P + 1]
  const long long col_raw = blockIdx.x * 32ll + lane;
  const long long col = col_raw < n_cols ? col_raw : n_cols - 1;
  const long long b = col / Fo;
  const int f = (int)(col - b * Fo);
  const uint4* yp = y + ((size_t)(b * NC1 + warp) * kHpT) * Fo + f;
  const uint32_t ninf = FMT == MPA_FMT_BF16 ? 0xFF80FF80u : 0xFC00FC00u;
  const uint4 NINF = make_uint4(ninf, ninf, ninf, ninf);
  uint4 xprev = NINF, a2[3], a4[5], a8[6], xs[6];
#pragma unroll
  for (int i = 0; i < 3; ++i) a2[i] = NINF;
#pragma unroll
  for (int i = 0; i < 5; ++i) a4[i] = NINF;
#pragma unroll
  for (int i = 0; i < 6; ++i) a8[i] = NINF;
  float acc[C2P];
#pragma unroll
  for (int i = 0; i < C2P; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int blk = 0; blk < (kHpT + 6 + kHpBlk - 1) / kHpBlk; ++blk) {
    const int p0 = blk * kHpBlk;
    // weights of the pooled frames t = p - 6 of this block's positions p (frames outside [0, 75): zeros)
    __syncwarp();
    for (int e = lane; e < kHpBlk * 8 * C2P / 4; e += 32) {
      const int i = e / (8 * C2P / 4), t = p0 + i - 6;
      float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t >= 0 && t < kHpT) wv = reinterpret_cast<const float4*>(w3p + ((size_t)warp * kHpT + t) * 8 * C2P)[e - i * (8 * C2P / 4)];
      reinterpret_cast<float4*>(w3s)[e] = wv;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kHpBlk; ++i) {
      const int p = p0 + i;
      if (i % 6 == 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) xs[j] = p + j < kHpT ? yp[(size_t)(p + j) * Fo] : NINF;
      }
      const uint4 xp = xs[i % 6];
      // pairs [p-1,p] -> slot (p+2)%3; fours [p-3,p] -> slot (p+2)%5; eights [p-7,p] -> slot (p+5)%6 (30 = 0 mod 3, 5, 6: p and i agree)
      hmax8<FMT>(a2[(i + 2) % 3], xprev, xp);
      hmax8<FMT>(a4[(i + 2) % 5], a2[i % 3], a2[(i + 2) % 3]);
      hmax8<FMT>(a8[(i + 5) % 6], a4[(i + 3) % 5], a4[(i + 2) % 5]);
      xprev = xp;
      if (p >= 6 && p < kHpT + 6) {
        uint4 m;
        hmax8<FMT>(m, a8[i % 6], a8[(i + 5) % 6]);       // frames [p-12, p] = the window of frame t = p - 6
        float v[8];
        unpack8<FMT>(m, v);
        const float4* wr = reinterpret_cast<const float4*>(w3s + (size_t)i * 8 * C2P);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
#pragma unroll
          for (int q = 0; q < C2P / 4; ++q) {
            const float4 w = wr[c8 * (C2P / 4) + q];
            acc[4 * q + 0] = fmaf(v[c8], w.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v[c8], w.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v[c8], w.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v[c8], w.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < C2P; ++i) red[((size_t)warp * 32 + lane) * (C2P + 1) + i] = acc[i];
  __syncthreads();
  if (warp != 0 || col_raw >= n_cols) return;
  float h[C2P];
#pragma unroll
  for (int i = 0; i < C2P; ++i) {
    float sum = 0.f;
    for (int ck = 0; ck < NC1; ++ck) sum += red[((size_t)ck * 32 + lane) * (C2P + 1) + i];      // fixed order
    sum += i < C2 ? b3[i] : 0.f;
    h[i] = sum >= 0.f ? sum : a * sum;
  }
  float o = b43[0];
  for (int c3 = 0; c3 < C3; ++c3) {
    float g = b40[c3];
#pragma unroll
    for (int i = 0; i < C2P; ++i) g = fmaf(i < C2 ? w40[c3 * C2 + i] : 0.f, h[i], g);
    g = g >= 0.f ? g : a * g;
    o = fmaf(w43[c3], g, o);
  }
  out[col_raw] = 1.f / (1.f + expf(-o));
}

template <int FMT, int C2P>
static int launch_head_pool_conv3_tail(const void* y, const float* w3p, const float* b3, const float* w40, const float* b40, const float* w43,
                                       const float* b43, float* out, int B, int NC1, int Fo, int C2, int C3, float a, cudaStream_t st) {
  static unsigned char flags[64];
  if (opt_in_max_smem(head_pool_conv3_tail_kernel<FMT, C2P>, flags) != cudaSuccess) return MPA_ERR_CUDA;
  const size_t smem = ((size_t)NC1 * kHpBlk * 8 * C2P + (size_t)NC1 * 32 * (C2P + 1)) * sizeof(float);
  const long long n_cols = (long long)B * Fo;
  head_pool_conv3_tail_kernel<FMT, C2P><<<ceil_div(n_cols, 32), NC1 * 32, smem, st>>>((const uint4*)y, w3p, b3, w40, b40, w43, b43, out, n_cols, NC1, Fo,
                                                                                     C2, C3, a);
  return MPA_OK;
}

}  // namespace mpa

using namespace mpa;

extern "C" int mpa_head_tail_cp8(const void* x_cp8, const float* w3, const float* b3, const float* w40, const float* b40, const float* w43,
                                 const float* b43, float* out, int B, int C1, int T, int Fo, int C2, int C3, float a_lrelu, int fmt,
                                 void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x_cp8 && w3 && b3 && w40 && b40 && w43 && b43 && out && B > 0 && C1 > 0 && T > 0, "head_tail: bad argument");
  MPA_REQUIRE(C2 >= 1 && C2 <= kC2Max && C3 >= 1 && C3 <= kC3Max && Fo >= 1 && Fo <= 256,
              "head_tail: unsupported widths C2=%d (<=%d) C3=%d (<=%d) Fo=%d", C2, kC2Max, C3, kC3Max, Fo);
  const int pb = 256 / Fo;
  const size_t smem = ((size_t)kTBlk * 8 * kC2Max + kC3Max * kC2Max + 80) * sizeof(float);
  if (fmt == MPA_FMT_BF16)
    head_tail_kernel<MPA_FMT_BF16><<<ceil_div(B, pb), 256, smem, (cudaStream_t)stream>>>((const uint4*)x_cp8, w3, b3, w40, b40, w43, b43, out, B,
                                                                                         C1, T, Fo, C2, C3, pb, a_lrelu);
  else
    head_tail_kernel<MPA_FMT_F16><<<ceil_div(B, pb), 256, smem, (cudaStream_t)stream>>>((const uint4*)x_cp8, w3, b3, w40, b40, w43, b43, out, B,
                                                                                        C1, T, Fo, C2, C3, pb, a_lrelu);
  MPA_CHECK_LAUNCH("head_tail");
  return MPA_OK;
}

extern "C" int mpa_head_tail2_cp8(const void* h_cp8, const float* w40, const float* b40, const float* w43, const float* b43, float* out, int B,
                                  int R, int Fo, int C2, int C3, float a_lrelu, int fmt, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(h_cp8 && w40 && b40 && w43 && b43 && out && B > 0 && R > 0 && Fo > 0 && C2 > 0 && C3 > 0, "head_tail2: bad argument");
  const size_t smem = ((size_t)C2 * Fo + (size_t)C3 * C2) * sizeof(float);
  MPA_REQUIRE(smem <= 200 * 1024, "head_tail2: C2=%d C3=%d Fo=%d need %zu B of shared memory", C2, C3, Fo, smem);
  if (fmt == MPA_FMT_BF16) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_tail2_kernel<MPA_FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    head_tail2_kernel<MPA_FMT_BF16><<<B * R, 128, smem, (cudaStream_t)stream>>>((const uint16_t*)h_cp8, w40, b40, w43, b43, out, R, Fo, C2, C3, a_lrelu, 0);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(head_tail2_kernel<MPA_FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    head_tail2_kernel<MPA_FMT_F16><<<B * R, 128, smem, (cudaStream_t)stream>>>((const uint16_t*)h_cp8, w40, b40, w43, b43, out, R, Fo, C2, C3, a_lrelu,
                                                                               fmt == MPA_FMT_F16X3 ? 1 : 0);
  }
  MPA_CHECK_LAUNCH("head_tail2");
  return MPA_OK;
}

extern "C" int mpa_head_pool_conv3_tail_cp8(const void* y_cp8, const float* w3_packed, const float* b3, const float* w40, const float* b40,
                                            const float* w43, const float* b43, float* out, int B, int C1, int T, int Fo, int C2, int C3,
                                            float a_lrelu, int fmt, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(y_cp8 && w3_packed && b3 && w40 && b40 && w43 && b43 && out && B > 0 && C1 > 0 && Fo > 0, "head_pool_conv3_tail: bad argument");
  const int NC1 = (C1 + 7) / 8;
  MPA_REQUIRE(T == kHpT && NC1 <= 8 && C2 >= 1 && C2 <= 16 && C3 >= 1 && C3 <= 64 && (fmt == MPA_FMT_BF16 || fmt == MPA_FMT_F16),
              "head_pool_conv3_tail: T == 75, C1 <= 64, C2 <= 16, 16-bit formats (got T=%d C1=%d C2=%d fmt=%d)", T, C1, C2, fmt);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (C2 <= 12)
    rc = fmt == MPA_FMT_BF16 ? launch_head_pool_conv3_tail<MPA_FMT_BF16, 12>(y_cp8, w3_packed, b3, w40, b40, w43, b43, out, B, NC1, Fo, C2, C3, a_lrelu, st)
                             : launch_head_pool_conv3_tail<MPA_FMT_F16, 12>(y_cp8, w3_packed, b3, w40, b40, w43, b43, out, B, NC1, Fo, C2, C3, a_lrelu, st);
  else
    rc = fmt == MPA_FMT_BF16 ? launch_head_pool_conv3_tail<MPA_FMT_BF16, 16>(y_cp8, w3_packed, b3, w40, b40, w43, b43, out, B, NC1, Fo, C2, C3, a_lrelu, st)
                             : launch_head_pool_conv3_tail<MPA_FMT_F16, 16>(y_cp8, w3_packed, b3, w40, b40, w43, b43, out, B, NC1, Fo, C2, C3, a_lrelu, st);
  MPA_REQUIRE(rc == MPA_OK, "head_pool_conv3_tail: shared memory opt-in failed");
  MPA_CHECK_LAUNCH("head_pool_conv3_tail");
  return MPA_OK;
}
