// Fused attention half of transformer_enc_layer (reference: libdl/nn_models/unet_cnns.py:131-153) in ONE launch, eval mode:
//   gather the bottleneck tokens (+ sinusoidal PE)  ->  folded q/k/v in-projection on tcgen05  ->  batch-axis softmax attention
//   ->  folded out_proj . o_linear on tcgen05  ->  + residual  ->  LayerNorm1
// (replaces enc_gather + gemm_nt(QKV) + batch_axis_attention + gemm_nt(proj) + add_ln: five launches and four fp32 round trips).
//
// The reference hands [B, S, E] to nn.MultiheadAttention without batch_first: the attention SEQUENCE is the batch axis (B <= 64 here),
// the S = Th*Fw bottleneck positions are independent.  One CTA therefore owns one position s and all B tokens of it:
//   D_blk[feature (128 TMEM lanes), token (N = 64 columns)] = W_blk[128 x E] . X_s^T[E x 64]     blk = q, k, v  (3 x E/16 MMAs)
// Operands are K-major SWIZZLE_NONE tiles: the weights come pre-chunked ([E/8][rows][8] 16-bit, mpa_gemm_tc_to_chunks with row tile 128:
// one 2 KB bulk copy per K chunk), the token tile is written by the CTA itself ([E/8][64][8]).  q, k, v return to shared memory as fp32
// (bias added), the B x B x heads attention runs on the CUDA cores in fp32 (two-pass softmax, the same formula as
// batch_axis_attention_kernel), its output goes back through the tensor cores for the projection, and residual + LayerNorm finish on the
// accumulator.  Outputs: h1 as fp32 tokens [B*S][E] and, optionally, as the 16-bit chunk operand of the MLP's first GEMM.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

namespace mpa {

constexpr int kEaThreads = 256;
constexpr int kEaTok = 64;                          // MMA N: tokens (= batch items) per position, zero padded
constexpr unsigned long long kEaTimeoutNs = 4000000000ull;

struct EncAttnParams {
  const float* x;            // [B][E][S] fp32 (NCHW bottleneck)
  const float* pe;           // [S][E] or null
  const uint8_t* w_qkv;      // chunks [KC][Nq][8] 16-bit, Nq = 3E padded to 128 rows
  const uint8_t* w_proj;     // chunks [KC][128][8]
  const float* b_qkv;        // [3E]
  const float* b_proj;       // [E]
  const float* ln_w;
  const float* ln_b;
  float* h1;                 // [B*S][E] fp32
  uint16_t* h1_chunks;       // optional: [KC][Mpad][8] 16-bit (X operand of the MLP's first GEMM)
  int B, E, S, H, KC, Nq_pad, Mpad, fmt;
  float eps;
  uint32_t idesc;
};

__device__ __forceinline__ uint32_t ea_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ea_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ea_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void ea_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ea_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool ea_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(ea_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void ea_wait(uint64_t* bar, uint32_t parity) {
  if (ea_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!ea_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > kEaTimeoutNs) __trap();
    }
  }
}
__device__ __forceinline__ void ea_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ea_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(ea_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ea_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ea_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ea_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ea_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ uint16_t ea_cvt16(float v, int fmt) {
  return fmt == MPA_FMT_BF16 ? __bfloat16_as_ushort(__float2bfloat16(v)) : __half_as_ushort(__float2half_rn(v));
}

__global__ void __launch_bounds__(kEaThreads, 1) enc_attn_tc_kernel(const EncAttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int E = p.E, B = p.B, S = p.S, KC = p.KC;
  const int wtile_bytes = KC * 128 * 16;                       // one 128-row weight block, all K chunks
  const int btile_bytes = KC * kEaTok * 16;                    // token operand tile
  uint8_t* wbuf = smem;                                        // [2][KC][128][16 B]
  uint8_t* bop = wbuf + 2 * wtile_bytes;                       // [KC][64][16 B]: tokens, later the attention output
  float* xs = reinterpret_cast<float*>(bop + btile_bytes);     // [64][E] gathered tokens (residual)
  float* qkv = xs + kEaTok * E;                                // [64][3E]; later y = proj + residual [64][E]
  uint64_t* bars = reinterpret_cast<uint64_t*>(qkv + kEaTok * 3 * E);
  uint64_t* wfull = bars;          // [2]
  uint64_t* wempty = bars + 2;     // [2]
  uint64_t* acc = bars + 4;        // [2]: q/k/v done, projection done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { ea_mbar_init(&wfull[i], 1); ea_mbar_init(&wempty[i], 1); ea_mbar_init(&acc[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ea_smem_u32(tmem_slot)), "r"(256u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // weight block i: 0,1,2 = q,k,v rows of the folded in-projection; 3 = folded out-projection
  auto load_block = [&](int i) {
    uint64_t* bar = &wfull[i & 1];
    uint8_t* dst = wbuf + (i & 1) * wtile_bytes;
    ea_expect_tx(bar, (uint32_t)wtile_bytes);
    const uint8_t* src = i < 3 ? p.w_qkv + (size_t)i * E * 16 : p.w_proj;          // rows [iE, iE+128) of the chunked [KC][Nq_pad] matrix
    const size_t kstride = (size_t)(i < 3 ? p.Nq_pad : 128) * 16;
    for (int kc = 0; kc < KC; ++kc) ea_bulk(dst + kc * 2048, src + kc * kstride, 2048u, bar);
  };
  if (threadIdx.x == 0) { load_block(0); load_block(1); }

  // ---- 1. gather x[:, :, s] (+ PE) -> fp32 residual copy + 16-bit token operand (rows >= B are zero)
  for (int i = threadIdx.x; i < kEaTok * E; i += kEaThreads) {
    const int e = i % E, b = i / E;
    float v = 0.f;
    if (b < B) {
      v = p.x[((size_t)b * E + e) * S + s];
      if (p.pe) v += p.pe[(size_t)s * E + e];
    }
    xs[b * E + e] = v;
    reinterpret_cast<uint16_t*>(bop)[((size_t)(e >> 3) * kEaTok + b) * 8 + (e & 7)] = ea_cvt16(v, p.fmt);
  }
  for (int i = threadIdx.x; i < (KC * 8 - E) * kEaTok; i += kEaThreads) {         // zero K padding (E < 8*KC)
    const int e = E + i / kEaTok, b = i % kEaTok;
    reinterpret_cast<uint16_t*>(bop)[((size_t)(e >> 3) * kEaTok + b) * 8 + (e & 7)] = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  constexpr uint32_t kHi = (128u >> 4) | (1u << 14);                               // SBO = 128 B, descriptor version 1
  const uint32_t kALo = (2048u >> 4) << 16;                                        // weights: next K chunk 2 KB further
  const uint32_t kBLo = ((uint32_t)(kEaTok * 16) >> 4) << 16;                      // tokens: next K chunk 1 KB further
  const uint32_t b16 = ea_smem_u32(bop) >> 4;
  auto issue_block = [&](int i, uint32_t col) {
    const uint32_t a16 = (ea_smem_u32(wbuf + (i & 1) * wtile_bytes) >> 4);
    for (int k = 0; k < KC / 2; ++k)
      ea_mma(tmem_base + col, ((uint64_t)kHi << 32) | (uint64_t)((a16 + (uint32_t)k * (2u * 2048u >> 4)) | kALo),
             ((uint64_t)kHi << 32) | (uint64_t)((b16 + (uint32_t)k * (2u * kEaTok * 16u >> 4)) | kBLo), p.idesc, k ? 1u : 0u);
  };
  // ---- 2. q, k, v on the tensor cores (one thread issues; the weight blocks stream through two buffers)
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) {
      ea_wait(&wfull[i & 1], (uint32_t)(i >> 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_block(i, (uint32_t)i * kEaTok);
      ea_commit(&wempty[i & 1]);
      if (i + 2 < 4) {
        ea_wait(&wempty[i & 1], (uint32_t)(i >> 1));
        load_block(i + 2);
      }
    }
    ea_commit(&acc[0]);
  }
  // ---- 3. accumulators -> fp32 q | k | v rows in shared memory (+ bias); warp w: lane quadrant w % 4, token half w / 4
  ea_wait(&acc[0], 0);
  __syncwarp();                                                  // lane 0 of warp 0 was the issuer: re-converge before the aligned loads
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int quad = warp & 3, half = warp >> 2;
    const int f = quad * 32 + lane;
    for (int blk = 0; blk < 3; ++blk) {
      uint32_t v[32];
      ea_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(blk * kEaTok + half * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (f < E) {
        const float bias = p.b_qkv[blk * E + f];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int tok = half * 32 + i;
          if (tok < B) qkv[(size_t)tok * 3 * E + blk * E + f] = __uint_as_float(v[i]) + bias;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---- 4. batch-axis attention in fp32: one (query token, head) per thread; output -> 16-bit operand tile (reuses the token tile)
  {
    const int hd = E / p.H;
    const float sc = rsqrtf((float)hd);
    for (int it = threadIdx.x; it < kEaTok * p.H; it += kEaThreads) {
      const int b1 = it / p.H, h = it - b1 * p.H;
      float o[64];
#pragma unroll
      for (int k = 0; k < 64; ++k) o[k] = 0.f;
      if (b1 < B) {
        const float* q = qkv + (size_t)b1 * 3 * E + h * hd;
        float mx = -INFINITY;
        for (int b2 = 0; b2 < B; ++b2) {
          const float* kk = qkv + (size_t)b2 * 3 * E + E + h * hd;
          float d = 0.f;
          for (int k = 0; k < hd; ++k) d = fmaf(q[k], kk[k], d);
          mx = fmaxf(mx, d * sc);
        }
        float den = 0.f;
        for (int b2 = 0; b2 < B; ++b2) {
          const float* kk = qkv + (size_t)b2 * 3 * E + E + h * hd;
          const float* vv = kk + E;
          float d = 0.f;
          for (int k = 0; k < hd; ++k) d = fmaf(q[k], kk[k], d);
          const float pr = expf(d * sc - mx);
          den += pr;
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (k < hd) o[k] = fmaf(pr, vv[k], o[k]);
        }
        const float inv = 1.f / den;
#pragma unroll
        for (int k = 0; k < 64; ++k)
          if (k < hd) o[k] = o[k] * inv;
      }
#pragma unroll
      for (int k = 0; k < 64; ++k)
        if (k < hd) {
          const int f = h * hd + k;
          reinterpret_cast<uint16_t*>(bop)[((size_t)(f >> 3) * kEaTok + b1) * 8 + (f & 7)] = ea_cvt16(o[k], p.fmt);
        }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  // ---- 5. projection on the tensor cores
  if (threadIdx.x == 0) {
    ea_wait(&wfull[1], 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    issue_block(3, 3u * kEaTok);
    ea_commit(&acc[1]);
  }
  ea_wait(&acc[1], 0);
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- 6. + bias + residual -> y [64][E] fp32 (over the q|k|v rows)
  float* ys = qkv;
  {
    const int quad = warp & 3, half = warp >> 2;
    const int f = quad * 32 + lane;
    uint32_t v[32];
    ea_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(3 * kEaTok + half * 32), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    __syncthreads();                                             // every thread is done reading q | k | v
    if (f < E) {
      const float bias = p.b_proj[f];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int tok = half * 32 + i;
        if (tok < B) ys[(size_t)tok * E + f] = __uint_as_float(v[i]) + bias + xs[(size_t)tok * E + f];
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---- 7. LayerNorm1 per token: 16 lanes x 8 features (E <= 128), two tokens per warp pass (uniform trip count: the half-warp
  // shuffles run under the full mask, a missing second token only skips its stores)
  {
    const int sub = lane >> 4, l16 = lane & 15;
    for (int t0 = warp * 2; t0 < B; t0 += 2 * (kEaThreads / 32)) {
      const int tok = t0 + sub;
      const bool live = tok < B;
      float v[8];
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = l16 * 8 + e;
        v[e] = (live && f < E) ? ys[(size_t)tok * E + f] : 0.f;
        sum += v[e];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
      const float mean = sum / E;
      float q = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = (l16 * 8 + e < E) ? v[e] - mean : 0.f;
        q += d * d;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o, 16);
      const float rstd = rsqrtf(q / E + p.eps);
      if (!live) continue;
      const size_t m = (size_t)tok * S + s;
      __align__(16) uint16_t c[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = l16 * 8 + e;
        float r = 0.f;
        if (f < E) {
          r = (v[e] - mean) * rstd * p.ln_w[f] + p.ln_b[f];
          p.h1[m * E + f] = r;
        }
        c[e] = ea_cvt16(r, p.fmt);
      }
      if (p.h1_chunks && l16 < KC) *reinterpret_cast<uint4*>(p.h1_chunks + ((size_t)l16 * p.Mpad + m) * 8) = *reinterpret_cast<const uint4*>(c);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u));
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" int mpa_enc_attn_block_tc(const float* x, const float* pe, const void* w_qkv_chunks, const float* b_qkv, const void* w_proj_chunks,
                                     const float* b_proj, const float* ln_w, const float* ln_b, float* h1, void* h1_chunks, int B, int E, int S,
                                     int num_heads, float eps, int fmt, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && w_qkv_chunks && b_qkv && w_proj_chunks && b_proj && ln_w && ln_b && h1, "enc_attn_block_tc: null argument");
  MPA_REQUIRE(B > 0 && B <= kEaTok && S > 0 && E >= 16 && E <= 128 && E % 16 == 0 && num_heads > 0 && E % num_heads == 0 && E / num_heads <= 64,
              "enc_attn_block_tc: needs B <= %d, E a multiple of 16 <= 128, head width <= 64 (got B=%d E=%d heads=%d)", kEaTok, B, E, num_heads);
  MPA_REQUIRE(fmt == MPA_FMT_F16 || fmt == MPA_FMT_BF16, "enc_attn_block_tc: fmt must be MPA_FMT_F16 or MPA_FMT_BF16");
  EncAttnParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.pe = pe;
  p.w_qkv = (const uint8_t*)w_qkv_chunks; p.w_proj = (const uint8_t*)w_proj_chunks;
  p.b_qkv = b_qkv; p.b_proj = b_proj; p.ln_w = ln_w; p.ln_b = ln_b;
  p.h1 = h1; p.h1_chunks = (uint16_t*)h1_chunks;
  p.B = B; p.E = E; p.S = S; p.H = num_heads;
  p.KC = (E + 63) / 64 * 8;
  // the in-projection rows [iE, iE+128) of block i must exist in the chunked matrix: mpa_gemm_tc_to_chunks pads 3E to a multiple of 128,
  // and for E < 128 block i over-reads into the next block's rows (harmless: those accumulator lanes are ignored) — 2E + 128 <= Nq_pad
  p.Nq_pad = (3 * E + 127) / 128 * 128;
  MPA_REQUIRE(2 * E + 128 <= p.Nq_pad, "enc_attn_block_tc: E=%d: the chunked in-projection has too few padded rows", E);
  p.Mpad = ((B * S) + 255) / 256 * 256;
  p.fmt = fmt; p.eps = eps;
  const uint32_t f = (fmt == MPA_FMT_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(kEaTok >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = (size_t)2 * p.KC * 2048 + (size_t)p.KC * kEaTok * 16 + sizeof(float) * kEaTok * E * 4 + 128;
  MPA_REQUIRE(smem <= 227 * 1024, "enc_attn_block_tc: needs %zu B of shared memory", smem);
  {
    static unsigned char flags[64];
    cudaError_t e = opt_in_max_smem(enc_attn_tc_kernel, flags);
    if (e != cudaSuccess) {
      set_error("enc_attn_block_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPA_ERR_CUDA;
    }
  }
  enc_attn_tc_kernel<<<S, kEaThreads, smem, (cudaStream_t)stream>>>(p);
  MPA_CHECK_LAUNCH("enc_attn_block_tc");
  return MPA_OK;
}
