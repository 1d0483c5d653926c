// Training-time patch augmentation on the device (reference: libdl/data_loaders/hcqt_datasets.py:67-141, where it
// runs per item in 16 DataLoader worker processes).  ONE kernel cuts the patches out of the un-patched [C][NT][F] HCQT tensor (arbitrary,
// e.g. shuffled, start frames) and applies, in the reference's order,
//   random EQ (per-harmonic parabola)  ->  additive Gaussian noise + abs  ->  log(1 + gamma x)  ->  tuning shift by -1, -1/2, 0, +1/2, +1 bin
//   ->  transposition by whole semitones (3 bins) with the vacated bins filled with |N(0, fill_std)|,
// and writes the centre-frame targets rolled by the same transposition.  The per-patch random DECISIONS (alpha, beta, tuning step,
// transposition) are drawn by the host mirror (incl. the reference's rejection loop on the EQ curve) and passed in as int arrays, so the
// deterministic part of the transform is bit-comparable with the reference; the Gaussian values come from Philox-4x32-10 keyed on the
// SOURCE element, so the two neighbours averaged by a half-bin tuning shift see the same noisy value, as in the reference.
// HBM-bound without noise: 5.2 kB read (L2-resident re-reads) + 388.8 kB written per 6x75x216 patch; with additive noise the Philox +
// Box-Muller arithmetic per element dominates.
#include "common.cuh"

namespace mpa {

__device__ __forceinline__ float gauss_from(uint32_t a, uint32_t b) {
  // Box-Muller on two 24-bit uniforms, u1 in (0,1]
  const float u1 = ((float)(a >> 8) + 1.f) * (1.f / 16777216.f);
  const float u2 = (float)(b >> 8) * (1.f / 16777216.f);
  return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

struct AugParams {
  const float* in;
  const long long* start;     // [n] first frame of each patch
  float* out;                 // [n][C][T][F]
  const int* eq_alpha;        // [n] or null; 0 = no EQ for this patch
  const int* eq_beta;         // [n]
  const int* eq_offset;       // [C] per-harmonic bin offset of the parabola's apex
  const int* tune2;           // [n] or null; tuning shift in half bins, -2..2
  const int* transp;          // [n] or null; semitones
  int C, NT, F, n, T, bins_per_semitone;
  float noise_std, gamma_log, fill_std;
  unsigned long long seed, offset;
};

// value of the patch at SOURCE bin f after EQ, noise and compression (hcqt_datasets.py:80-106)
__device__ __forceinline__ float aug_source(const AugParams& p, const float* __restrict__ row, long long elem_row, int f, float a_eq, int apex,
                                            uint2 key) {
  float v = row[f];
  if (a_eq != 0.f) {
    const int d = f - apex;
    // (1 - (2e-6*alpha)*d^2) * x with one rounding per operation, as the reference's tensor expression evaluates it
    v = __fmul_rn(__fsub_rn(1.f, __fmul_rn(a_eq, (float)(d * d))), v);
  }
  if (p.noise_std > 0.f) {
    const unsigned long long e = (unsigned long long)(elem_row + f);
    const uint4 r = philox(make_uint4((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)p.offset, (uint32_t)(p.offset >> 32) ^ 0x10000000u), key);
    v = fabsf(__fadd_rn(v, __fmul_rn(p.noise_std, gauss_from(r.x, r.y))));
  }
  if (p.gamma_log > 0.f) v = logf(__fadd_rn(1.f, __fmul_rn(p.gamma_log, v)));
  return v;
}

// one CTA per (patch, channel) plane: 256 threads sweep its T x F elements (16,200 per plane: enough work per CTA to amortise the
// per-patch parameter loads; consecutive threads write consecutive bins)
__global__ void __launch_bounds__(256) augment_patches_kernel(AugParams p) {
  const int c = blockIdx.x % p.C;
  const int b = blockIdx.x / p.C;
  const int F = p.F;
  const float* plane = p.in + ((size_t)c * p.NT + (size_t)p.start[b]) * F;
  const long long elem_plane = (long long)blockIdx.x * p.T * F;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  const int alpha = p.eq_alpha ? p.eq_alpha[b] : 0;
  const float a_eq = alpha ? __fmul_rn(2e-6f, (float)alpha) : 0.f;
  const int apex = alpha ? p.eq_beta[b] - p.eq_offset[c] : 0;
  const int ts = p.tune2 ? p.tune2[b] : 0;
  const int tr = p.transp ? p.transp[b] : 0;
  const int shift = tr * p.bins_per_semitone;
  const bool plain = (alpha == 0 && ts == 0 && tr == 0 && p.noise_std == 0.f);
  float* oplane = p.out + elem_plane;
  const int total = p.T * F;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int t = e / F, f = e - t * F;
    const float* row = plane + (size_t)t * F;
    const long long elem_row = elem_plane + (long long)t * F;
    float v;
    if (plain) {
      v = row[f];
      if (p.gamma_log > 0.f) v = logf(__fadd_rn(1.f, __fmul_rn(p.gamma_log, v)));
      oplane[e] = v;
      continue;
    }
    const bool fill_tr = (tr > 0 && f < shift) || (tr < 0 && f >= F + shift);
    int s = (f - shift) % F;
    if (s < 0) s += F;
    const bool fill_tu = (ts > 0 && s == 0) || (ts < 0 && s == F - 1);
    if (fill_tr || fill_tu) {
      const unsigned long long q = (unsigned long long)(elem_row + (fill_tr ? f : s));
      const uint4 r = philox(make_uint4((uint32_t)q, (uint32_t)(q >> 32), (uint32_t)p.offset,
                                        (uint32_t)(p.offset >> 32) ^ (fill_tr ? 0x30000000u : 0x20000000u)), key);
      v = fabsf(__fmul_rn(p.fill_std, gauss_from(r.x, r.y)));
    } else if (ts == 0) {
      v = aug_source(p, row, elem_row, s, a_eq, apex, key);
    } else if (ts == 2) {
      v = aug_source(p, row, elem_row, s - 1, a_eq, apex, key);
    } else if (ts == -2) {
      v = aug_source(p, row, elem_row, s + 1, a_eq, apex, key);
    } else {
      const int s0 = ts == 1 ? s - 1 : s;
      v = __fmul_rn(0.5f, __fadd_rn(aug_source(p, row, elem_row, s0, a_eq, apex, key), aug_source(p, row, elem_row, s0 + 1, a_eq, apex, key)));
    }
    oplane[e] = v;
  }
}

// y[b][q] = targets[frame[b]][q - tr] with the wrapped entries zeroed (pitch targets) or a plain roll (12 pitch classes)
__global__ void augment_targets_kernel(const float* __restrict__ targets, const long long* __restrict__ frame, const int* __restrict__ transp,
                                       float* __restrict__ y, int n, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * P) return;
  const int b = i / P, q = i % P;
  const int tr = transp ? transp[b] : 0;
  int s = (q - tr) % P;
  if (s < 0) s += P;
  float v = targets[(size_t)frame[b] * P + s];
  if (P != 12 && ((tr > 0 && q < tr) || (tr < 0 && q >= P + tr))) v = 0.f;
  y[i] = v;
}

}  // namespace mpa

using namespace mpa;

extern "C" {

int mpa_augment_patches_f32(const float* in, const long long* start, float* out, int C, int NT, int F, int n, int T, const int* eq_alpha,
                            const int* eq_beta, const int* eq_offset, float noise_std, float gamma_log, const int* tune2, const int* transp,
                            int bins_per_semitone, float fill_std, unsigned long long seed, unsigned long long offset, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(in && start && out && C > 0 && NT >= T && F > 1 && n > 0 && T > 0, "augment_patches: bad argument");
  MPA_REQUIRE(!eq_alpha || (eq_beta && eq_offset), "augment_patches: eq_alpha needs eq_beta and eq_offset");
  MPA_REQUIRE(noise_std >= 0.f && fill_std >= 0.f && bins_per_semitone >= 1, "augment_patches: bad noise / transposition parameter");
  AugParams p;
  p.in = in; p.start = start; p.out = out; p.eq_alpha = eq_alpha; p.eq_beta = eq_beta; p.eq_offset = eq_offset; p.tune2 = tune2;
  p.transp = transp; p.C = C; p.NT = NT; p.F = F; p.n = n; p.T = T; p.bins_per_semitone = bins_per_semitone;
  p.noise_std = noise_std; p.gamma_log = gamma_log; p.fill_std = fill_std; p.seed = seed; p.offset = offset;
  MPA_REQUIRE((long long)n * C < 2147483647LL && (long long)T * F < 2147483647LL, "augment_patches: too many planes for one launch");
  augment_patches_kernel<<<(unsigned)(n * C), 256, 0, (cudaStream_t)stream>>>(p);
  MPA_CHECK_LAUNCH("augment_patches");
  return MPA_OK;
}

int mpa_augment_targets_f32(const float* targets, const long long* frame, const int* transp, float* y, int n, int P, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(targets && frame && y && n > 0 && P > 0, "augment_targets: bad argument");
  augment_targets_kernel<<<ceil_div((long long)n * P, 256), 256, 0, (cudaStream_t)stream>>>(targets, frame, transp, y, n, P);
  MPA_CHECK_LAUNCH("augment_targets");
  return MPA_OK;
}

}  // extern "C"
