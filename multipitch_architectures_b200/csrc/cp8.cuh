// 16-byte CP8 pixel (8 channels x 16 bit) <-> 8 floats, for the element-wise training kernels on the planes (train_cp8.cu, train_unet_cp8.cu)
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace mpa {

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

}  // namespace mpa
