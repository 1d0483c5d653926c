// Shared helpers for libmpa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include "../../include/mpa.h"

namespace mpa {

void set_error(const char* fmt, ...);
int check_arch();
void count_launch(int n = 1);
// reduce.cu: parallel, order-stable per-channel reductions (launch only; the caller checks the launch)
int channel_sum_launch(const float* x, float* out, int B, int C, int HW, cudaStream_t st);
int bn_bwd_sums_launch(const float* x, const float* out_act, const float* dy, const float* stats, float eps, float* sums2c, float* dw, float* db,
                       int B, int C, int HW, int relu, cudaStream_t st);
int bn_stats_launch(const float* x, float* stats, int B, int C, int HW, cudaStream_t st);

#define MPA_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      mpa::set_error(__VA_ARGS__);             \
      return MPA_ERR_ARG;                      \
    }                                          \
  } while (0)

#define MPA_CHECK_ARCH()                       \
  do {                                         \
    int _a = mpa::check_arch();                \
    if (_a != MPA_OK) return _a;               \
  } while (0)

#define MPA_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      mpa::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));     \
      return MPA_ERR_CUDA;                                                       \
    }                                                                            \
    mpa::count_launch();                                                         \
  } while (0)

__device__ __forceinline__ float apply_act(float v, int act, float p) {
  if (act == MPA_ACT_LRELU) return v >= 0.f ? v : p * v;
  if (act == MPA_ACT_RELU) return fmaxf(v, 0.f);
  if (act == MPA_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox-4x32-10 counter-based generator (dropout masks, augmentation noise): 4 x 32 random bits per (counter, key)
__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// nn.Dropout keep factor of element i of a tensor under (seed, offset): the convention of dropout_kernel (counter i/4, lane i%4)
struct DropoutArgs {
  float p;                        // 0 = no dropout
  unsigned long long seed, offset;
  const long long* step_dev;      // optional device-side step counter: offset += step_dev[0] * step_mul
  unsigned long long step_mul;
};
__device__ __forceinline__ unsigned long long dropout_offset(const DropoutArgs& d) {
  return d.step_dev ? d.offset + (unsigned long long)d.step_dev[0] * d.step_mul : d.offset;
}
__device__ __forceinline__ uint4 dropout_bits(long long i4, unsigned long long seed, unsigned long long offset) {
  return philox(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
// keep <=> (bits >> 8) * 2^-24 >= p  <=>  (bits >> 8) >= ceil(p * 2^24)  <=>  bits >= ceil(p * 2^24) << 8   (all steps exact for 0 <= p < 1;
// the threshold is loop-invariant, which leaves one integer compare and one select per element)
__device__ __forceinline__ float dropout_factor(uint32_t bits, float p, float scale) {
  const uint32_t thr = (uint32_t)ceilf(p * 16777216.f) << 8;
  return bits >= thr ? scale : 0.f;
}
__device__ __forceinline__ float dropout_factor_at(long long i, float p, float scale, unsigned long long seed, unsigned long long offset) {
  const uint4 r = dropout_bits(i >> 2, seed, offset);
  const int e = (int)(i & 3);
  return dropout_factor(e == 0 ? r.x : e == 1 ? r.y : e == 2 ? r.z : r.w, p, scale);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on the symbol); nullptr when unavailable
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = nullptr;
  static int tried = 0;
  if (!__atomic_load_n(&tried, __ATOMIC_ACQUIRE)) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    const char* off = getenv("MPA_NO_TENSOR_MAP");          // A/B switch and test hook: "1" keeps every kernel on its bulk-copy loops
    if (!(off && off[0] == '1') && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (TensorMapEncodeFn)sym;
    __atomic_store_n(&tried, 1, __ATOMIC_RELEASE);
  }
  return fn;
}

// Opt a kernel into the maximum dynamic shared memory once per (kernel, device) for the whole PROCESS.  (The attribute is a
// property of the function on the device, not of the calling thread: per-thread bookkeeping lets an autograd worker thread
// lower what the main thread believes it has set.)  `flags`: one static array per kernel.
template <typename K>
static inline cudaError_t opt_in_max_smem(K kernel, unsigned char* flags /*[64]*/) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (__atomic_load_n(&flags[dev], __ATOMIC_ACQUIRE)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) __atomic_store_n(&flags[dev], (unsigned char)1, __ATOMIC_RELEASE);
  return e;
}

}  // namespace mpa
