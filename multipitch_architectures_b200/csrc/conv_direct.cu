// Generic direct convolution on fp32 CUDA cores: every kernel size / stride of the model zoo
// (15x15, 9x9, 5x5, 3x3, 3x3 stride (1,3), 75x1, 1x1, 2x5, 2x3).  This is the exact-fp32 path and the path of
// all layers that are not (yet) served by the tcgen05 implicit-GEMM kernel (conv_tc.cu).
//
// Tiling: a CTA of 256 threads owns a 16 x 32 output-pixel tile for a block of 16 output channels; each thread
// accumulates 2 pixels x 16 channels in registers.  Input channels are staged through shared memory
// CIN_STEP at a time (halo tile + the 16-channel weight slab), so every input element is read from L2 once per
// 16 output channels and every weight once per 512 pixels.
#include "common.cuh"

namespace mpa {

constexpr int TW = 32, TH = 16, COB = 16;

__global__ void __launch_bounds__(256)
conv2d_direct_kernel(const float* __restrict__ x, const float* __restrict__ x2, int Cin1, const float* __restrict__ wp,
                     const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                     float* __restrict__ out, int Cin, int H, int W, int Cout, int CoutPad, int KH, int KW, int sh, int sw,
                     int ph, int pw, int Ho, int Wo, int cin_step, int act, float act_param) {
  extern __shared__ float smem[];
  const int IH = (TH - 1) * sh + KH, IW = (TW - 1) * sw + KW;
  const int KK = KH * KW;
  float* in_s = smem;                                  // [cin_step][IH][IW]
  float* w_s = smem + (size_t)cin_step * IH * IW;      // [cin_step][KK][16]
  const int n_cob = CoutPad / COB;
  const int b = blockIdx.z / n_cob, cb = blockIdx.z % n_cob;
  const int ho0 = blockIdx.y * TH, wo0 = blockIdx.x * TW;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // ty in 0..7 -> rows ty, ty+8
  const int ih0 = ho0 * sh - ph, iw0 = wo0 * sw - pw;

  float acc0[COB], acc1[COB];
#pragma unroll
  for (int j = 0; j < COB; ++j) acc0[j] = acc1[j] = 0.f;

  const int Cin2 = Cin - Cin1;
  for (int ci0 = 0; ci0 < Cin; ci0 += cin_step) {
    const int nci = min(cin_step, Cin - ci0);
    __syncthreads();
    // stage the halo tile
    const int tile_n = nci * IH * IW;
    for (int e = threadIdx.x; e < tile_n; e += 256) {
      int iw = e % IW;
      int r = e / IW;
      int ih = r % IH;
      int s = r / IH;
      int gh = ih0 + ih, gw = iw0 + iw, ci = ci0 + s;
      float v = 0.f;
      if (gh >= 0 && gh < H && gw >= 0 && gw < W) {
        v = (ci < Cin1) ? x[(((size_t)b * Cin1 + ci) * H + gh) * W + gw]
                        : x2[(((size_t)b * Cin2 + (ci - Cin1)) * H + gh) * W + gw];
      }
      in_s[e] = v;
    }
    // stage the weight slab
    const int w_n = nci * KK * COB;
    for (int e = threadIdx.x; e < w_n; e += 256) {
      int j = e & (COB - 1);
      int r = e >> 4;
      int tap = r % KK;
      int s = r / KK;
      w_s[e] = wp[((size_t)(ci0 + s) * KK + tap) * CoutPad + cb * COB + j];
    }
    __syncthreads();
    for (int s = 0; s < nci; ++s) {
      const float* ip0 = in_s + (size_t)s * IH * IW + (ty * sh) * IW + tx * sw;
      const float* ip1 = ip0 + (8 * sh) * IW;
      const float4* wq = reinterpret_cast<const float4*>(w_s + (size_t)s * KK * COB);
      for (int kh = 0; kh < KH; ++kh) {
        for (int kw = 0; kw < KW; ++kw) {
          float v0 = ip0[kh * IW + kw], v1 = ip1[kh * IW + kw];
          const float4* wt = wq + (kh * KW + kw) * 4;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 w4 = wt[q];
            acc0[q * 4 + 0] = fmaf(v0, w4.x, acc0[q * 4 + 0]);
            acc0[q * 4 + 1] = fmaf(v0, w4.y, acc0[q * 4 + 1]);
            acc0[q * 4 + 2] = fmaf(v0, w4.z, acc0[q * 4 + 2]);
            acc0[q * 4 + 3] = fmaf(v0, w4.w, acc0[q * 4 + 3]);
            acc1[q * 4 + 0] = fmaf(v1, w4.x, acc1[q * 4 + 0]);
            acc1[q * 4 + 1] = fmaf(v1, w4.y, acc1[q * 4 + 1]);
            acc1[q * 4 + 2] = fmaf(v1, w4.z, acc1[q * 4 + 2]);
            acc1[q * 4 + 3] = fmaf(v1, w4.w, acc1[q * 4 + 3]);
          }
        }
      }
    }
  }
  const int wo = wo0 + tx;
  if (wo >= Wo) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int ho = ho0 + ty + half * 8;
    if (ho >= Ho) continue;
#pragma unroll
    for (int j = 0; j < COB; ++j) {
      const int co = cb * COB + j;
      if (co < Cout) {
        float v = half ? acc1[j] : acc0[j];
        if (bias) v += bias[co];
        if (scale) v = v * scale[co] + shift[co];
        out[(((size_t)b * Cout + co) * Ho + ho) * Wo + wo] = apply_act(v, act, act_param);
      }
    }
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" int mpa_conv2d_f32(const float* x, const float* x2, int Cin1, const float* w_packed, const float* bias,
                              const float* scale, const float* shift, float* out, int B, int Cin, int H, int W, int Cout,
                              int KH, int KW, int sh, int sw, int ph, int pw, int act, float act_param, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_packed && out && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && KH > 0 && KW > 0 && sh > 0 && sw > 0,
              "conv2d: bad argument");
  MPA_REQUIRE((scale == nullptr) == (shift == nullptr), "conv2d: scale and shift must be given together");
  if (!x2) Cin1 = Cin;
  MPA_REQUIRE(Cin1 > 0 && Cin1 <= Cin, "conv2d: bad Cin1");
  const int Ho = (H + 2 * ph - KH) / sh + 1, Wo = (W + 2 * pw - KW) / sw + 1;
  MPA_REQUIRE(Ho > 0 && Wo > 0, "conv2d: empty output (H=%d W=%d K=%dx%d)", H, W, KH, KW);
  const int CoutPad = (Cout + COB - 1) / COB * COB;
  const int IH = (TH - 1) * sh + KH, IW = (TW - 1) * sw + KW;
  const size_t per_ci = ((size_t)IH * IW + (size_t)KH * KW * COB) * sizeof(float);
  int cin_step = (int)(48 * 1024 / per_ci);
  if (cin_step > 8) cin_step = 8;
  if (cin_step > Cin) cin_step = Cin;
  size_t smem = per_ci * (cin_step < 1 ? 1 : cin_step);
  cudaStream_t st = (cudaStream_t)stream;
  if (cin_step < 1) {
    cin_step = 1;
    MPA_REQUIRE(smem <= 200 * 1024, "conv2d: kernel %dx%d stride %dx%d needs %zu B of shared memory", KH, KW, sh, sw, smem);
    cudaError_t e = cudaFuncSetAttribute(conv2d_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv2d: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPA_ERR_CUDA;
    }
  }
  dim3 grid(ceil_div(Wo, TW), ceil_div(Ho, TH), B * (CoutPad / COB));
  MPA_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "conv2d: grid too large (B*cout blocks = %u)", grid.z);
  conv2d_direct_kernel<<<grid, 256, smem, st>>>(x, x2, Cin1, w_packed, bias, scale, shift, out, Cin, H, W, Cout, CoutPad, KH,
                                                 KW, sh, sw, ph, pw, Ho, Wo, cin_step, act, act_param);
  MPA_CHECK_LAUNCH("conv2d_direct");
  return MPA_OK;
}
