// tcgen05 GEMM for the token-wise Linear layers of transformer_enc_layer (unet_cnns.py:131-157: the MLP E -> mlp_dim -> E holds
// > 95 % of the layer's FLOPs):      Y[M, N] = act(X[M, K] * W[N, K]^T + bias[N])        (nn.Linear)
//
// Both operands are pre-arranged in the channel-chunk layout the convolutions use, [K/8][rows][8] 16-bit, which IS the K-major
// SWIZZLE_NONE canonical operand layout: a tile (128 weight rows or 256 tokens) x one 8-wide K chunk is one contiguous bulk copy.
//   D[n (128 TMEM lanes = output features), m (256 columns = tokens)] += W_tile[n, k] * X_tile[k, m]
// so the fp32 epilogue writes Y row-major with 32 consecutive features per warp store (coalesced), adding bias / ReLU, or — for
// split-K (wide K, few tiles: the second MLP layer) — accumulates with atomicAdd into a pre-initialised Y.
// Warp roles as in conv_tc.cu: TMA producer, single-lane MMA issue from a warp-uniform loop, 4 epilogue warps; the accumulator is
// double buffered in TMEM (2 x 256 columns) so the epilogue of tile k overlaps the main loop of tile k+1.
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>

namespace mpa {

constexpr int kGmThreads = 192;              // producer warp, MMA warp, 4 epilogue warps
constexpr int kGmStages = 4;
constexpr int kGmChunksPerStage = 8;         // K = 64 per stage = 4 MMAs
constexpr int kGmTileN = 128, kGmTileM = 256;
constexpr int kGmWBytes = kGmChunksPerStage * kGmTileN * 16;     // 16 KB
constexpr int kGmXBytes = kGmChunksPerStage * kGmTileM * 16;     // 32 KB
constexpr unsigned long long kGmTimeoutNs = 4000000000ull;

struct GemmTcParams {
  const uint8_t* x;          // [KC][Mpad][8] 16-bit
  const uint8_t* w;          // [KC][Npad][8] 16-bit
  const float* bias;         // [N] or null
  float* y;                  // [M][N] fp32 row-major
  int M, N, KC, Mpad, Npad, relu, ksplit, n_tiles_n, n_tiles_m, n_units;
  uint32_t idesc;
  // optional 16-bit copies of the result in operand layouts (ksplit == 1 only), written by the epilogue instead of a converter launch:
  uint16_t* y_tok;           // token-chunked [y_tok_chunks][y_tok_rows (features)][8]: the operand of a product that reduces over the tokens
  int y_tok_rows, y_tok_chunks;
  uint16_t* y_feat;          // feature-chunked [ceil(Npad/8)][y_feat_rows (tokens)][8]: the X operand of the next Linear layer
  int y_feat_rows;
  const uint16_t* mask_tok;  // layout of y_tok: the result is zeroed where this tensor is <= 0 (ReLU backward)
  float* colsum;             // [N] += sum over the tokens of the (masked) result (bias gradient), atomics
  int bf16;
  // fp32 result address = y + moff(m) + noff(n), each a 2-level index (i / n2) * s1 + (i % n2) * s2 (n2 == 0: i * s2): lets a product write
  // straight into an NCHW tensor whose (item, bin) pair is one GEMM dimension (the head's 75x1 convolution and its gradients)
  int y_mn2, y_nn2;
  long long y_ms1, y_ms2, y_ns1, y_ns2;
  // residual add + LayerNorm over the N features in the epilogue (N <= 128: one feature tile; no K split): out_nchw[b][n][s] =
  // LN(result + bias + ln_res[m][n]) * ln_w + ln_b for token m = b * ln_S + s — the second add & LayerNorm of transformer_enc_layer
  const float *ln_res, *ln_w, *ln_b;
  float* ln_out;
  float ln_eps;
  int ln_S;
  unsigned* ln_counters;     // K split: arrival counter per token tile (zero on entry, reset by the last arriver)
  // a stage's 8 K chunks of each operand tile as ONE tensor-map copy (the copy unit retires ~1 bulk copy per 150 clocks whatever its size:
  // 16 copies of 2-4 KB per stage of 4 MMAs = 256 tensor clocks left the encoder products copy-issue bound at 3-17 % tensor-active)
  int tma;
  alignas(64) CUtensorMap tm_x;   // [KC][Mpad / 128][128 rows x 16 B]: box 8 x 2 x 2 KB
  alignas(64) CUtensorMap tm_w;   // [KC][Npad / 128][128 rows x 16 B]: box 8 x 1 x 2 KB
};
__device__ __forceinline__ long long gm_off(int i, int n2, long long s1, long long s2) {
  return n2 > 0 ? (long long)(i / n2) * s1 + (long long)(i % n2) * s2 : (long long)i * s2;
}
constexpr int kGmTrBytes = 32 * 80;            // per epilogue warp: 32 tokens x (32 features x 2 B + 16 B pad) transposing tile
constexpr int kGmLnPitch = 129;                // fp32 staging tile of the LayerNorm epilogue: [32 tokens][128 features + 1]
constexpr int kGmLnBytes = 32 * kGmLnPitch * 4 + 32 * 8;     // + (mean, rstd) per token

__device__ __forceinline__ uint32_t gm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gm_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void gm_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gm_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool gm_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(gm_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void gm_wait(uint64_t* bar, uint32_t parity) {
  if (gm_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!gm_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > kGmTimeoutNs) __trap();
    }
  }
}
__device__ __forceinline__ void gm_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gm_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(gm_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gm_tensor3(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   gm_smem_u32(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(gm_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool gm_elect() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void gm_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gm_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gm_ld32(uint32_t taddr, uint32_t* r) {
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

struct GmUnit {
  int n0, m0, kc0, kc1;
};
__device__ __forceinline__ GmUnit gm_decode(const GemmTcParams& p, int u) {
  GmUnit g;
  const int ks = u % p.ksplit;
  int t = u / p.ksplit;
  const int tn = t % p.n_tiles_n, tm = t / p.n_tiles_n;
  g.n0 = tn * kGmTileN;
  g.m0 = tm * kGmTileM;
  const int per = (p.KC / kGmChunksPerStage + p.ksplit - 1) / p.ksplit;      // stages per K slice (KC is a multiple of 8 chunks)
  g.kc0 = ks * per * kGmChunksPerStage;
  g.kc1 = min(p.KC, g.kc0 + per * kGmChunksPerStage);
  return g;
}

__global__ void __launch_bounds__(kGmThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* w_smem = smem;                                        // [stages][8 chunks][128 rows][16 B]
  uint8_t* x_smem = smem + kGmStages * kGmWBytes;                 // [stages][8 chunks][256 rows][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_smem + kGmStages * kGmXBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kGmStages;
  uint64_t* acc_full = empty + kGmStages;      // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* tr_smem = reinterpret_cast<uint8_t*>(bars) + 256;      // [4 epilogue warps][kGmTrBytes]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGmStages; ++i) { gm_mbar_init(&full[i], 1); gm_mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { gm_mbar_init(&acc_full[i], 1); gm_mbar_init(&acc_empty[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gm_smem_u32(tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // the whole warp walks the loops: lane 0 waits on / arms the stage barrier, then 16 lanes issue the stage's 16 bulk copies at once
    // (one thread needs ~170 clocks per copy for address arithmetic + issue: 16 copies per stage of 4 MMAs = 512 tensor-clocks made
    // the single-thread producer the limiter of this kernel)
    int st = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
      const GmUnit g = gm_decode(p, u);
      for (int kc = g.kc0; kc < g.kc1; kc += kGmChunksPerStage) {
        if (lane == 0) {
          gm_wait(&empty[st], ph ^ 1);
          gm_expect_tx(&full[st], (uint32_t)(kGmWBytes + kGmXBytes));
        }
        __syncwarp();
        if (p.tma) {
          if (lane == 0) {
            gm_tensor3(w_smem + st * kGmWBytes, &p.tm_w, 0, g.n0 / 128, kc, &full[st]);
            gm_tensor3(x_smem + st * kGmXBytes, &p.tm_x, 0, g.m0 / 128, kc, &full[st]);
          }
        } else if (lane < 2 * kGmChunksPerStage) {
          const int c = lane >> 1;
          if (lane & 1)
            gm_bulk(x_smem + st * kGmXBytes + c * (kGmTileM * 16), p.x + ((size_t)(kc + c) * p.Mpad + g.m0) * 16, kGmTileM * 16, &full[st]);
          else
            gm_bulk(w_smem + st * kGmWBytes + c * (kGmTileN * 16), p.w + ((size_t)(kc + c) * p.Npad + g.n0) * 16, kGmTileN * 16, &full[st]);
        }
        __syncwarp();
        if (++st == kGmStages) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t w16 = gm_smem_u32(w_smem) >> 4, x16 = gm_smem_u32(x_smem) >> 4;
    constexpr uint32_t kHi = (128u >> 4) | (1u << 14);                               // SBO = 128 B
    constexpr uint32_t kWLo = ((uint32_t)(kGmTileN * 16) >> 4) << 16;                // LBO = next K chunk of the weight tile
    constexpr uint32_t kXLo = ((uint32_t)(kGmTileM * 16) >> 4) << 16;
    int st = 0;
    uint32_t ph = 0, k_unit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++k_unit) {
      const GmUnit g = gm_decode(p, u);
      const uint32_t buf = k_unit & 1u;
      gm_wait(&acc_empty[buf], ((k_unit >> 1) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_u + buf * 256;
      uint32_t first = 1;
      for (int kc = g.kc0; kc < g.kc1; kc += kGmChunksPerStage) {
        gm_wait(&full[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_lo = (w16 + (uint32_t)st * (kGmWBytes >> 4)) | kWLo;
        const uint32_t b_lo = (x16 + (uint32_t)st * (kGmXBytes >> 4)) | kXLo;
        if (gm_elect()) {
          gm_mma(tmem_d, ((uint64_t)kHi << 32) | a_lo, ((uint64_t)kHi << 32) | b_lo, p.idesc, first ? 0u : 1u);
#pragma unroll
          for (int i = 1; i < kGmChunksPerStage / 2; ++i)
            gm_mma(tmem_d, ((uint64_t)kHi << 32) | (uint64_t)(a_lo + (uint32_t)i * (2 * kGmTileN * 16 >> 4)),
                   ((uint64_t)kHi << 32) | (uint64_t)(b_lo + (uint32_t)i * (2 * kGmTileM * 16 >> 4)), p.idesc, 1u);
          gm_commit(&empty[st]);
        }
        __syncwarp();
        first = 0;
        if (++st == kGmStages) { st = 0; ph ^= 1; }
      }
      if (gm_elect()) gm_commit(&acc_full[buf]);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    uint32_t k_unit = 0;
    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x, ++k_unit) {
      const GmUnit g = gm_decode(p, u);
      const uint32_t buf = k_unit & 1u;
      gm_wait(&acc_full[buf], (k_unit >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int n = g.n0 + quad * 32 + lane;
      const bool n_ok = n < p.N;
      const float b = (n_ok && p.bias && g.kc0 == 0) ? p.bias[n] : 0.f;
      const int m_hi = min(kGmTileM, p.M - g.m0);
      const bool extra = p.y_tok || p.y_feat || p.mask_tok || p.colsum || p.ln_out;
      float csum = 0.f;
      uint8_t* tr = tr_smem + quad * kGmTrBytes;
      // ReLU-backward mask of the NEXT 32-token block is requested before the current block's arithmetic: the loads were issued where
      // they were needed and every block of the tile waited a full L2 / HBM latency for them (ncu: 81 us against 27 us for the same
      // product without the mask)
      uint4 mk_nxt[4];
      auto load_mask = [&](int c0, uint4* mk) {
        const int ch = (g.m0 + c0) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          mk[q] = (p.mask_tok && n_ok && c0 < m_hi && ch + q < p.y_tok_chunks)
                      ? __ldg(reinterpret_cast<const uint4*>(p.mask_tok + ((size_t)(ch + q) * p.y_tok_rows + n) * 8))
                      : make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);       // "positive": nothing masked
      };
      if (p.mask_tok) load_mask(0, mk_nxt);
      for (int c0 = 0; c0 < (extra ? kGmTileM : m_hi); c0 += 32) {
        uint4 mk_cur[4];
        if (p.mask_tok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) mk_cur[q] = mk_nxt[q];
          load_mask(c0 + 32, mk_nxt);
        }
        if (c0 >= m_hi) {
          // token rows past M inside the padded tile: the feature-chunked copy must hold zeros there (the consumer's bulk copies read them)
          if (p.y_feat) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(p.y_feat + ((size_t)((g.n0 + quad * 32) / 8 + j) * p.y_feat_rows + g.m0 + c0 + lane) * 8) = make_uint4(0, 0, 0, 0);
          }
          continue;
        }
        uint32_t v[32];
        gm_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 256 + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!extra || (p.ln_out && p.ksplit > 1)) {
          if (n_ok) {
            float* yn = p.y + gm_off(n, p.y_nn2, p.y_ns1, p.y_ns2);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (c0 + i < m_hi) {
                float val = __uint_as_float(v[i]) + b;
                float* yp = yn + gm_off(g.m0 + c0 + i, p.y_mn2, p.y_ms1, p.y_ms2);
                if (p.ksplit > 1) atomicAdd(yp, val);
                else *yp = p.relu ? fmaxf(val, 0.f) : val;
              }
            }
          }
          continue;
        }
        // ---- epilogue with operand-layout copies (ksplit == 1)
        const int ch0 = (g.m0 + c0) >> 3;                      // first token chunk of this block
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float val = __uint_as_float(v[i]) + b;
          if (p.relu) val = fmaxf(val, 0.f);
          if (!n_ok || c0 + i >= m_hi) val = 0.f;
          v[i] = __float_as_uint(val);
        }
        if (p.ln_out && p.ksplit == 1) {
          // ---- residual + LayerNorm over the features (the four epilogue warps hold all N <= 128 features of these 32 tokens)
          float* stage = reinterpret_cast<float*>(tr_smem);                    // [32][kGmLnPitch]
          float2* tstat = reinterpret_cast<float2*>(tr_smem + 32 * kGmLnPitch * 4);
          const int nloc = quad * 32 + lane;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float val = __uint_as_float(v[i]);
            if (n_ok && c0 + i < m_hi) val += p.ln_res[(size_t)(g.m0 + c0 + i) * p.N + n];
            stage[i * kGmLnPitch + nloc] = val;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          for (int t8 = 0; t8 < 8; ++t8) {                                    // warp `quad` owns tokens quad*8 .. quad*8+7
            const int tok = quad * 8 + t8;
            float x4[4], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              x4[k] = lane + 32 * k < p.N ? stage[tok * kGmLnPitch + lane + 32 * k] : 0.f;
              sum += x4[k];
            }
            const float mean = warp_sum(sum) / p.N;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (lane + 32 * k < p.N) {
                const float d = x4[k] - mean;
                q += d * d;
              }
            const float rstd = rsqrtf(warp_sum(q) / p.N + p.ln_eps);
            if (lane == 0) tstat[tok] = make_float2(mean, rstd);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          {
            // lanes = tokens: consecutive lanes write consecutive bottleneck positions of one NCHW feature plane
            const int m = g.m0 + c0 + lane;
            const bool live = c0 + lane < m_hi;
            const float2 st = tstat[lane];
            const int bb = live ? m / p.ln_S : 0, ss = live ? m - bb * p.ln_S : 0;
            for (int e = quad * 32; e < quad * 32 + 32 && e < p.N; ++e) {
              const float r = (stage[lane * kGmLnPitch + e] - st.x) * st.y * p.ln_w[e] + p.ln_b[e];
              if (live) p.ln_out[((size_t)bb * p.N + e) * p.ln_S + ss] = r;
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          continue;
        }
        if (p.mask_tok && n_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ch0 + q < p.y_tok_chunks) {
              const uint4 mk = mk_cur[q];
              const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t h = (e & 1) ? (mw[e >> 1] >> 16) : (mw[e >> 1] & 0xFFFFu);
                // > 0 in either 16-bit format: sign bit clear and magnitude non-zero
                if ((h & 0x8000u) || (h & 0x7FFFu) == 0u) v[q * 8 + e] = 0u;
              }
            }
          }
        }
        if (p.colsum) {
#pragma unroll
          for (int i = 0; i < 32; ++i) csum += __uint_as_float(v[i]);
        }
        if (p.y && n_ok) {
          float* yn = p.y + gm_off(n, p.y_nn2, p.y_ns1, p.y_ns2);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < m_hi) yn[gm_off(g.m0 + c0 + i, p.y_mn2, p.y_ms1, p.y_ms2)] = __uint_as_float(v[i]);
        }
        uint32_t h16[16];                                      // the 32 values as 16-bit pairs (tokens 2i, 2i+1)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (p.bf16) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            h16[i] = *reinterpret_cast<const uint32_t*>(&h);
          } else {
            const __half2 h = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            h16[i] = *reinterpret_cast<const uint32_t*>(&h);
          }
        }
        if (p.y_tok && n < p.y_tok_rows) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (ch0 + q < p.y_tok_chunks)
              *reinterpret_cast<uint4*>(p.y_tok + ((size_t)(ch0 + q) * p.y_tok_rows + n) * 8) =
                  make_uint4(h16[4 * q], h16[4 * q + 1], h16[4 * q + 2], h16[4 * q + 3]);
        }
        if (p.y_feat) {
          // transpose the warp's 32 features x 32 tokens through shared memory: row = token (80-byte stride: conflict-free 16-byte reads)
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            *reinterpret_cast<uint16_t*>(tr + (2 * i) * 80 + lane * 2) = (uint16_t)(h16[i] & 0xFFFFu);
            *reinterpret_cast<uint16_t*>(tr + (2 * i + 1) * 80 + lane * 2) = (uint16_t)(h16[i] >> 16);
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 piece = *reinterpret_cast<const uint4*>(tr + lane * 80 + j * 16);
            *reinterpret_cast<uint4*>(p.y_feat + ((size_t)((g.n0 + quad * 32) / 8 + j) * p.y_feat_rows + g.m0 + c0 + lane) * 8) = piece;
          }
        }
      }
      if (p.colsum && n_ok) atomicAdd(&p.colsum[n], csum);
      if (p.ln_out && p.ksplit > 1) {
        // K split: every slice of this token tile has added its part to y with atomics; the slice that arrives LAST (all N <= 128 features
        // are one feature tile) reads the sums back from L2 and finishes residual + LayerNorm for the tile's tokens
        volatile unsigned& ln_last = *reinterpret_cast<volatile unsigned*>(tmem_slot + 2);       // inside the 256-byte barrier block (no static shared memory: the kernel opts in to all 227 KB as dynamic)
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (quad == 0 && lane == 0) ln_last = atomicAdd(&p.ln_counters[g.m0 / kGmTileM], 1u) == (unsigned)p.ksplit - 1;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (ln_last) {
          __threadfence();
          for (int tok = quad; tok < m_hi; tok += 4) {                    // one warp per token, 4 features per lane
            const int m = g.m0 + tok;
            float x4[4], sum = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = lane + 32 * k;
              x4[k] = e < p.N ? __ldcg(&p.y[(size_t)m * p.N + e]) + p.ln_res[(size_t)m * p.N + e] : 0.f;
              sum += x4[k];
            }
            const float mean = warp_sum(sum) / p.N;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (lane + 32 * k < p.N) {
                const float d = x4[k] - mean;
                q += d * d;
              }
            const float rstd = rsqrtf(warp_sum(q) / p.N + p.ln_eps);
            const int bb = m / p.ln_S, ss = m - bb * p.ln_S;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = lane + 32 * k;
              if (e < p.N) p.ln_out[((size_t)bb * p.N + e) * p.ln_S + ss] = (x4[k] - mean) * rstd * p.ln_w[e] + p.ln_b[e];
            }
          }
          if (quad == 0 && lane == 0) p.ln_counters[g.m0 / kGmTileM] = 0u;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      gm_arrive(&acc_empty[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// fp32 row-major [R][K] (or, transposed, [K][R]) -> 16-bit chunk layout [KCpad][Rpad][8] (zero padded rows / columns)
__global__ void rows_to_chunks_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, long long total, int R, int K, int Rpad, int fmt,
                                      int transposed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % Rpad);
    const int kc = (int)(i / Rpad);
    __align__(16) uint16_t v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      const float f = (r < R && k < K) ? (transposed ? x[(size_t)k * R + r] : x[(size_t)r * K + k]) : 0.f;
      v[e] = fmt == MPA_FMT_BF16 ? __bfloat16_as_ushort(__float2bfloat16(f)) : __half_as_ushort(__float2half_rn(f));
    }
    *reinterpret_cast<uint4*>(out + (size_t)i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// element (r, k) at x[(r / r_n2) * r_s1 + (r % r_n2) * r_s2 + (k / k_n2) * k_s1 + (k % k_n2) * k_s2] -> the same chunk layout: any NCHW tensor whose
// GEMM dimensions are pairs of its axes (the head's 75x1 convolution: tokens = (item, bin), K = (channel, frame))
__global__ void strided_to_chunks_kernel(const float* __restrict__ x, uint16_t* __restrict__ out, long long total, int R, int K, int Rpad, int fmt,
                                         int r_n2, long long r_s1, long long r_s2, int k_n2, long long k_s1, long long k_s2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % Rpad);
    const int kc = (int)(i / Rpad);
    __align__(16) uint16_t v[8];
    const long long ro = r < R ? (long long)(r / r_n2) * r_s1 + (long long)(r % r_n2) * r_s2 : 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kc * 8 + e;
      const float f = (r < R && k < K) ? x[ro + (long long)(k / k_n2) * k_s1 + (long long)(k % k_n2) * k_s2] : 0.f;
      v[e] = fmt == MPA_FMT_BF16 ? __bfloat16_as_ushort(__float2bfloat16(f)) : __half_as_ushort(__float2half_rn(f));
    }
    *reinterpret_cast<uint4*>(out + (size_t)i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_gemm_tc_strided_to_chunks(const float* x, void* out, int rows, int K, int row_tile, int fmt, int r_n2, long long r_s1, long long r_s2,
                                  int k_n2, long long k_s1, long long k_s2, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && out && rows > 0 && K > 0 && row_tile > 0 && row_tile % kGmTileN == 0 && r_n2 > 0 && k_n2 > 0,
              "gemm_tc_strided_to_chunks: bad argument (row_tile: a multiple of 128)");
  const int rpad = (rows + row_tile - 1) / row_tile * row_tile, kc = (K + 63) / 64 * 8;
  const long long total = (long long)kc * rpad;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  strided_to_chunks_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)out, total, rows, K, rpad, fmt, r_n2, r_s1, r_s2, k_n2, k_s1,
                                                                     k_s2);
  MPA_CHECK_LAUNCH("gemm_tc_strided_to_chunks");
  return MPA_OK;
}

size_t mpa_gemm_tc_chunked_bytes(int rows, int K, int row_tile) {
  if (rows <= 0 || K <= 0 || row_tile <= 0) return 0;
  const size_t rpad = (size_t)(rows + row_tile - 1) / row_tile * row_tile;
  const size_t kc = (size_t)(K + 63) / 64 * 8;
  return kc * rpad * 16;
}

int mpa_gemm_tc_to_chunks(const float* x, void* out, int rows, int K, int row_tile, int fmt, int transposed, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(x && out && rows > 0 && K > 0 && (row_tile == kGmTileN || row_tile == kGmTileM), "gemm_tc_to_chunks: bad argument (row_tile 128 or 256)");
  const int rpad = (rows + row_tile - 1) / row_tile * row_tile, kc = (K + 63) / 64 * 8;
  const long long total = (long long)kc * rpad;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  rows_to_chunks_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(x, (uint16_t*)out, total, rows, K, rpad, fmt, transposed);
  MPA_CHECK_LAUNCH("gemm_tc_to_chunks");
  return MPA_OK;
}

int mpa_gemm_tc_run(const mpa_gemm_tc_desc* d, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(d, "gemm_tc: null descriptor");
  const void *x_chunks = d->x_chunks, *w_chunks = d->w_chunks, *mask_tok = d->mask_tok;
  const float* bias = d->bias;
  float *y = d->y, *colsum = d->colsum;
  void *y_tok = d->y_tok, *y_feat = d->y_feat;
  const int M = d->M, N = d->N, K = d->K, relu = d->relu, fmt = d->fmt, x_rows = d->x_rows, w_rows = d->w_rows, y_tok_rows = d->y_tok_rows,
            y_tok_chunks = d->y_tok_chunks, y_feat_rows = d->y_feat_rows;
  const bool extra = y_tok || y_feat || mask_tok || colsum || d->ln_out;
  MPA_REQUIRE(!d->ln_out || (d->ln_res && d->ln_w && d->ln_b && d->ln_S > 0 && N <= kGmTileN && !y_tok && !y_feat && !mask_tok && !colsum && !relu),
              "gemm_tc: the LayerNorm epilogue needs N <= 128, its residual / weight / bias and no other epilogue output");
  MPA_REQUIRE(x_chunks && w_chunks && (y || extra) && M > 0 && N > 0 && K > 0, "gemm_tc: bad argument");
  MPA_REQUIRE(fmt == MPA_FMT_F16 || fmt == MPA_FMT_BF16, "gemm_tc: fmt must be MPA_FMT_F16 or MPA_FMT_BF16");
  MPA_REQUIRE((((uintptr_t)x_chunks | (uintptr_t)w_chunks | (uintptr_t)y_tok | (uintptr_t)y_feat | (uintptr_t)mask_tok) & 15) == 0,
              "gemm_tc: 16-byte alignment required");
  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.x = (const uint8_t*)x_chunks;
  p.w = (const uint8_t*)w_chunks;
  p.bias = bias;
  p.y = y;
  p.M = M; p.N = N;
  p.KC = (K + 63) / 64 * 8;
  p.Mpad = (M + kGmTileM - 1) / kGmTileM * kGmTileM;
  p.Npad = (N + kGmTileN - 1) / kGmTileN * kGmTileN;
  p.relu = relu;
  p.n_tiles_n = p.Npad / kGmTileN;
  p.n_tiles_m = p.Mpad / kGmTileM;
  // row strides of the operand buffers (a buffer written by another product's epilogue may be padded further than this product needs)
  MPA_REQUIRE((x_rows == 0 || x_rows >= p.Mpad) && (w_rows == 0 || w_rows >= p.Npad), "gemm_tc: operand row stride smaller than the padded tile range");
  if (x_rows > 0) p.Mpad = x_rows;
  if (w_rows > 0) p.Npad = w_rows;
  if (y_tok || mask_tok) MPA_REQUIRE(y_tok_rows >= N && y_tok_chunks > 0, "gemm_tc: y_tok / mask_tok need their row stride and chunk count");
  if (y_feat) MPA_REQUIRE(y_feat_rows >= p.n_tiles_m * kGmTileM, "gemm_tc: y_feat row stride must cover the padded token tiles");
  p.y_tok = (uint16_t*)y_tok; p.y_tok_rows = y_tok_rows; p.y_tok_chunks = y_tok_chunks;
  p.y_feat = (uint16_t*)y_feat; p.y_feat_rows = y_feat_rows;
  p.mask_tok = (const uint16_t*)mask_tok; p.colsum = colsum;
  p.bf16 = fmt == MPA_FMT_BF16;
  p.ln_res = d->ln_res; p.ln_w = d->ln_w; p.ln_b = d->ln_b; p.ln_out = d->ln_out; p.ln_eps = d->ln_eps; p.ln_S = d->ln_S;
  p.y_mn2 = d->y_mn2; p.y_ms1 = d->y_ms1; p.y_ms2 = d->y_mn2 > 0 || d->y_ms2 != 0 ? d->y_ms2 : (long long)N;
  p.y_nn2 = d->y_nn2; p.y_ns1 = d->y_ns1; p.y_ns2 = d->y_nn2 > 0 || d->y_ns2 != 0 ? d->y_ns2 : 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = p.n_tiles_n * p.n_tiles_m, stages = p.KC / kGmChunksPerStage;
  int ks = 1;
  const bool ln_only = d->ln_out && !y_tok && !y_feat && !mask_tok && !colsum;
  if (!relu && (!extra || (ln_only && y)) && tiles < sms && stages >= 8) {
    ks = sms / tiles;
    if (ks > stages / 4) ks = stages / 4;
    if (ks < 1) ks = 1;
    const int per = (stages + ks - 1) / ks;
    ks = (stages + per - 1) / per;                       // every K slice non-empty
  }
  p.ksplit = ks;
  p.n_units = tiles * ks;
  if (d->ln_out && ks > 1) {
    static unsigned* counters[64] = {nullptr};
    int cd = dev < 0 || dev >= 64 ? 0 : dev;
    if (!counters[cd]) {
      if (cudaMalloc(&counters[cd], sizeof(unsigned) * 65536) != cudaSuccess || cudaMemset(counters[cd], 0, sizeof(unsigned) * 65536) != cudaSuccess) {
        set_error("gemm_tc: counter allocation failed");
        return MPA_ERR_CUDA;
      }
    }
    MPA_REQUIRE(p.n_tiles_m <= 65536, "gemm_tc: too many token tiles for the LayerNorm epilogue");
    p.ln_counters = counters[cd];
  }
  const uint32_t f = (fmt == MPA_FMT_BF16) ? 1u : 0u;
  p.idesc = (1u << 4) | (f << 7) | (f << 10) | ((uint32_t)(kGmTileM >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // the K slices meet in atomics: a dense result is zeroed here, a strided one (y_mn2 / y_nn2 / explicit strides) by the caller
  const bool dense_y = d->y_mn2 == 0 && d->y_nn2 == 0 && d->y_ms2 == 0 && d->y_ns2 == 0;
  if (ks > 1 && dense_y) cudaMemsetAsync(y, 0, sizeof(float) * (size_t)M * N, (cudaStream_t)stream);
  if (ks > 1 && !dense_y && !d->y_zeroed) ks = 1, p.ksplit = 1, p.n_units = tiles;
  p.tma = 0;
  if (TensorMapEncodeFn enc = tensor_map_encoder()) {
    if (p.Mpad % 128 == 0 && p.Npad % 128 == 0) {
      const cuuint64_t xd[3] = {256, (cuuint64_t)(p.Mpad / 128), (cuuint64_t)p.KC}, xs[2] = {2048, (cuuint64_t)p.Mpad * 16};
      const cuuint64_t wd[3] = {256, (cuuint64_t)(p.Npad / 128), (cuuint64_t)p.KC}, ws[2] = {2048, (cuuint64_t)p.Npad * 16};
      const cuuint32_t xb[3] = {256, kGmTileM / 128, kGmChunksPerStage}, wb[3] = {256, kGmTileN / 128, kGmChunksPerStage}, es[3] = {1, 1, 1};
      const CUresult r1 = enc(&p.tm_x, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(p.x), xd, xs, xb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      const CUresult r2 = enc(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(p.w), wd, ws, wb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      p.tma = (r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS) ? 1 : 0;          // else: the bulk-copy loop
    }
  }
  const size_t smem = (size_t)kGmStages * (kGmWBytes + kGmXBytes) + 256 + (4 * kGmTrBytes > kGmLnBytes ? 4 * kGmTrBytes : kGmLnBytes);
  {
    static unsigned char flags[64];
    cudaError_t e = opt_in_max_smem(gemm_tc_kernel, flags);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return MPA_ERR_CUDA;
    }
  }
  const int grid = p.n_units < sms ? p.n_units : sms;
  gemm_tc_kernel<<<grid, kGmThreads, smem, (cudaStream_t)stream>>>(p);
  MPA_CHECK_LAUNCH("gemm_tc");
  return MPA_OK;
}

int mpa_gemm_tc_f16(const void* x_chunks, const void* w_chunks, const float* bias, float* y, int M, int N, int K, int relu, int fmt, void* stream) {
  mpa_gemm_tc_desc d;
  memset(&d, 0, sizeof(d));
  d.x_chunks = x_chunks; d.w_chunks = w_chunks; d.bias = bias; d.y = y; d.M = M; d.N = N; d.K = K; d.relu = relu; d.fmt = fmt;
  return mpa_gemm_tc_run(&d, stream);
}

}  // extern "C"
