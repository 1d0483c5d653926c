// Fused tail of the network head (identical in every model, e.g. basic_cnns.py:396-408):
//   conv3 (T x 1 "time reduction", C1 -> C2) + LeakyReLU -> conv4.0 (1x1, C2 -> C3) + LeakyReLU
//   -> conv4.3 (1x1, C3 -> 1) -> sigmoid
// for the patch-wise case T == kernel height (one output frame per patch).  Input is the max-pooled conv2 output
// NCHW fp32 [B, C1, T, Fo]; output [B, Fo].  One thread owns one (patch, bin) and keeps all C2 conv3 accumulators
// in registers, so conv4.* never touch memory.  Reads are coalesced along Fo; the conv3 weights of one input
// channel are staged in shared memory ([t][co], broadcast float4 reads).
#include "common.cuh"

namespace mpa {

constexpr int kC2Max = 32, kC3Max = 16;

__global__ void __launch_bounds__(256) head_tail_kernel(const float* __restrict__ x, const float* __restrict__ w3, const float* __restrict__ b3,
                                                        const float* __restrict__ w40, const float* __restrict__ b40, const float* __restrict__ w43,
                                                        const float* __restrict__ b43, float* __restrict__ out, int B, int C1, int T, int Fo, int C2,
                                                        int C3, int pb, float a) {
  extern __shared__ float sm[];
  float* w3s = sm;                          // [T][kC2Max]
  float* w40s = sm + (size_t)T * kC2Max;    // [C3][kC2Max]
  float* misc = w40s + kC3Max * kC2Max;     // b3[32] b40[16] w43[16] b43[1]
  const int tid = threadIdx.x;
  const int bl = tid / Fo, f = tid - bl * Fo;
  const int b = blockIdx.x * pb + bl;
  const bool active = (bl < pb) && (b < B);
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
  const float* xb = x + ((size_t)(active ? b : 0) * C1) * T * Fo + f;
  for (int ci = 0; ci < C1; ++ci) {
    __syncthreads();
    for (int e = tid; e < T * kC2Max; e += blockDim.x) {
      int t = e / kC2Max, co = e - t * kC2Max;
      w3s[e] = co < C2 ? w3[((size_t)co * C1 + ci) * T + t] : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* xp = xb + (size_t)ci * T * Fo;
#pragma unroll 5
      for (int t = 0; t < T; ++t) {
        const float v = xp[(size_t)t * Fo];
        const float4* wr = reinterpret_cast<const float4*>(w3s + t * kC2Max);
#pragma unroll
        for (int q = 0; q < kC2Max / 4; ++q) {
          const float4 w = wr[q];
          acc[4 * q + 0] = fmaf(v, w.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(v, w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v, w.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(v, w.w, acc[4 * q + 3]);
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

}  // namespace mpa

using namespace mpa;

extern "C" int mpa_head_tail_f32(const float* x, const float* w3, const float* b3, const float* w40, const float* b40, const float* w43,
                                 const float* b43, float* out, int B, int C1, int T, int Fo, int C2, int C3, float a_lrelu, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w3 && b3 && w40 && b40 && w43 && b43 && out && B > 0 && C1 > 0 && T > 0, "head_tail: bad argument");
  MPA_REQUIRE(C2 >= 1 && C2 <= kC2Max && C3 >= 1 && C3 <= kC3Max && Fo >= 1 && Fo <= 256,
              "head_tail: unsupported widths C2=%d (<=%d) C3=%d (<=%d) Fo=%d", C2, kC2Max, C3, kC3Max, Fo);
  const int pb = 256 / Fo;
  const size_t smem = ((size_t)T * kC2Max + kC3Max * kC2Max + 80) * sizeof(float);
  MPA_REQUIRE(smem <= 48 * 1024, "head_tail: T=%d too long", T);
  head_tail_kernel<<<ceil_div(B, pb), 256, smem, (cudaStream_t)stream>>>(x, w3, b3, w40, b40, w43, b43, out, B, C1, T, Fo, C2, C3, pb, a_lrelu);
  MPA_CHECK_LAUNCH("head_tail");
  return MPA_OK;
}
