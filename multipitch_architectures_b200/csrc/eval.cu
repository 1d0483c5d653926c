// Evaluation measures on the device (reference: libdl/metrics/eval_metrics.py:8-110,158-189, which works through
// numpy / libfmp / mir_eval on the host after copying every prediction back).  One warp per frame computes, in float64 like the
// reference's numpy expressions, every per-frame term of
//   precision / recall / F (TP, est and ref counts; libfmp.c5.compute_eval_measures), cosine similarity of the L2-normalised frames
//   (libfmp.c3.normalize_feature_sequence, unit vector below 1e-10), binary cross-entropy in bits, Euclidean distance, binary and soft
//   accuracy, accumulated energy, and the frame-level multi-pitch counts behind mir_eval.multipitch.evaluate on a semitone grid
//   (matched pitches, matched pitch classes, substitution / miss / false-alarm terms);
// a second single-CTA kernel adds the frames in a fixed order (bit-reproducible).  Only the reduced 16 numbers travel to the host.
// HBM-bound: 2 x 4 B per (frame, bin) read once.
#include "common.cuh"

namespace mpa {

constexpr int kEvalStats = 16;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stats[frame][kEvalStats]:
//  0 TP  1 n_est  2 #(t > 0)  3 <t/|t|, p/|p|>  4 sum_p t*log2(p+eps) + (1-t)*log2(1-p+eps)  5 |t-p|_2  6 #(thresholded == t)
//  7 sum_p t*p + (1-t)*(1-p)  8 sum(t*p)/(sum(t)+eps)  9 matched pitch classes  10 min(n_ref,n_est)  11 max(n_ref,n_est)
//  12 max(0,n_ref-n_est)  13 max(0,n_est-n_ref)  14 n_ref = #(t != 0)  15 zero
__global__ void __launch_bounds__(128) eval_frame_stats_kernel(const float* __restrict__ targ, const float* __restrict__ pred, int N, int P,
                                                               double threshold, int min_pitch, double* __restrict__ stats) {
  __shared__ int cls[4][2][12];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 4 + w;
  if (lane < 24) cls[w][lane / 12][lane % 12] = 0;
  __syncwarp();
  if (n >= N) return;
  const double eps = 2.220446049250313e-16;      // np.finfo(float).eps
  double tp = 0, ne = 0, nr = 0, tt = 0, pp = 0, tpdot = 0, bce = 0, d2 = 0, acc = 0, soft = 0, st = 0, sp = 0;
  for (int q = lane; q < P; q += 32) {
    const float pf = pred[(size_t)n * P + q];
    const double t = (double)targ[(size_t)n * P + q], p = (double)pf;
    const bool est = p >= threshold, ref = t != 0.0;
    tp += (ref && est) ? 1.0 : 0.0;
    ne += est ? 1.0 : 0.0;
    nr += (t > 0.0) ? 1.0 : 0.0;
    tt += t * t;
    pp += p * p;
    tpdot += t * p;
    bce += t * log2(p + eps) + (1.0 - t) * log2(1.0 - p + eps);
    d2 += (t - p) * (t - p);
    acc += ((est ? 1.0 : 0.0) == t) ? 1.0 : 0.0;
    soft += t * p + (1.0 - t) * (1.0 - p);
    st += t;
    sp += p;
    const int pc = (min_pitch + q) % 12;
    if (ref) atomicAdd(&cls[w][0][pc], 1);
    if (est) atomicAdd(&cls[w][1][pc], 1);
  }
  tp = warp_sum_f64(tp); ne = warp_sum_f64(ne); nr = warp_sum_f64(nr); tt = warp_sum_f64(tt); pp = warp_sum_f64(pp);
  tpdot = warp_sum_f64(tpdot); bce = warp_sum_f64(bce); d2 = warp_sum_f64(d2); acc = warp_sum_f64(acc); soft = warp_sum_f64(soft);
  st = warp_sum_f64(st); sp = warp_sum_f64(sp);
  __syncwarp();
  if (lane != 0) return;
  // cosine term with libfmp's fallback: a frame whose L2 norm is <= 1e-10 is replaced by the unit vector ones(P)/sqrt(P)
  const double nt = sqrt(tt), np_ = sqrt(pp), u = 1.0 / sqrt((double)P);
  double cosv;
  if (nt > 1e-10 && np_ > 1e-10) cosv = tpdot / (nt * np_);
  else if (nt > 1e-10) cosv = (st / nt) * u;
  else if (np_ > 1e-10) cosv = (sp / np_) * u;
  else cosv = (double)P * u * u;
  int ctp = 0, nref_nz = 0;
  for (int c = 0; c < 12; ++c) {
    ctp += min(cls[w][0][c], cls[w][1][c]);
    nref_nz += cls[w][0][c];
  }
  double* s = stats + (size_t)n * kEvalStats;
  const double nref = (double)nref_nz;     // mir_eval's reference pitches are np.nonzero(targ[k])
  s[0] = tp; s[1] = ne; s[2] = nr; s[3] = cosv; s[4] = bce; s[5] = sqrt(d2); s[6] = acc; s[7] = soft; s[8] = tpdot / (st + eps);
  s[9] = (double)ctp; s[10] = fmin(nref, ne); s[11] = fmax(nref, ne); s[12] = fmax(0.0, nref - ne); s[13] = fmax(0.0, ne - nref);
  s[14] = nref; s[15] = 0.0;
}

// out[k] = sum_n stats[n][k], fixed order: CTA k, thread i adds frames i, i+1024, ...; then a shared-memory tree
__global__ void __launch_bounds__(1024) eval_reduce_kernel(const double* __restrict__ stats, int N, double* __restrict__ out) {
  __shared__ double sm[1024];
  const int k = blockIdx.x;
  double a = 0.0;
  for (int n = threadIdx.x; n < N; n += 1024) a += stats[(size_t)n * kEvalStats + k];
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = sm[0];
}

}  // namespace mpa

using namespace mpa;

extern "C" {

size_t mpa_eval_workspace(int n_frames) { return (size_t)(n_frames < 1 ? 1 : n_frames) * kEvalStats * sizeof(double); }

int mpa_eval_sums_f32(const float* targ, const float* pred, int n_frames, int n_bins, double threshold, int min_pitch, double* sums16,
                      void* workspace, size_t ws_bytes, void* stream) {
  MPA_CHECK_ARCH();
  MPA_REQUIRE(targ && pred && sums16 && workspace && n_frames > 0 && n_bins > 0 && min_pitch >= 0, "eval_sums: bad argument");
  if (ws_bytes < mpa_eval_workspace(n_frames)) {
    set_error("eval_sums: workspace %zu < %zu bytes", ws_bytes, mpa_eval_workspace(n_frames));
    return MPA_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  eval_frame_stats_kernel<<<ceil_div(n_frames, 4), 128, 0, st>>>(targ, pred, n_frames, n_bins, threshold, min_pitch, (double*)workspace);
  MPA_CHECK_LAUNCH("eval_frame_stats");
  eval_reduce_kernel<<<kEvalStats, 1024, 0, st>>>((const double*)workspace, n_frames, sums16);
  MPA_CHECK_LAUNCH("eval_reduce");
  return MPA_OK;
}

}  // extern "C"
