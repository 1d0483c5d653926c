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
