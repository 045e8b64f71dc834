// Detector validation metric on the device: the per-sample part of DC_Metrics.update
// (/root/reference/src/models/metrics.py:48-73) -- compute_l2_distance (:102-129) and compute_ratio (:75-100) -- on the
// engine's decode output (dcu_decode_gather / dcu_infer_batch: kpts rows x, y, id, cell per frame) and the label maps.
//
// Per sample: labels = cells with ids_target != dustbin, decoded like label_to_keypoints (:25-35: x = 8*col + p%8,
// y = 8*row + p//8); for every label id that was also predicted, the WORST distance between a prediction with that id and
// the label (torch.cdist + max over predictions).  l2 = sum of those / max(1, ids found), ratio = #(worst < 3 px) / #labels.
// All coordinates are integers, so each distance is a correctly rounded fp32 sqrt of an exact integer (bit-exact with
// torch.cdist); only the final sum's order differs from torch.sum.  Label ids are unique per sample in the reference (its
// metric raises on a repeated label id); a repeated label id is treated here as "worst over all pairs".
// One CTA per sample: threads stride over the label cells, predictions (a few dozen rows) are re-read from L2.
#include "common.cuh"

namespace dcu {

namespace {

constexpr int M_THREADS = 128;
constexpr int M_MAX_IDS = 64;

__global__ void __launch_bounds__(M_THREADS)
dc_metrics_kernel(MetricsParams p) {
  __shared__ unsigned int maxd_bits[M_MAX_IDS];     // worst distance per id (non-negative floats order like their bit patterns)
  __shared__ int has_label[M_MAX_IDS], matched[M_MAX_IDS];
  __shared__ int n_labels;
  const int f = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < M_MAX_IDS; i += M_THREADS) { maxd_bits[i] = 0u; has_label[i] = 0; matched[i] = 0; }
  if (tid == 0) n_labels = 0;
  __syncthreads();
  const int cells = p.h * p.w;
  const long long* lt = p.loc_target + (size_t)f * cells;
  const long long* it = p.ids_target + (size_t)f * cells;
  // decode_gather writes the full counts / offsets even when it dropped rows at max_patches: never read past the buffer
  const int K = max(0, min(p.counts[f], p.max_rows - p.offsets[f]));
  const int4* rows = reinterpret_cast<const int4*>(p.kpts) + p.offsets[f];
  for (int c = tid; c < cells; c += M_THREADS) {
    const long long id = it[c];
    if (id == p.dust_bin || id < 0 || id >= M_MAX_IDS) continue;
    atomicAdd(&n_labels, 1);
    has_label[id] = 1;
    const int pp = (int)lt[c];
    const float tx = (float)(8 * (c % p.w) + (pp % 8)), ty = (float)(8 * (c / p.w) + (pp / 8));
    for (int j = 0; j < K; ++j) {
      const int4 r = rows[j];
      if (r.z != (int)id) continue;
      const float dx = (float)r.x - tx, dy = (float)r.y - ty;
      const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      atomicMax(&maxd_bits[id], __float_as_uint(d));
      matched[id] = 1;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    int found = 0, good = 0;
    for (int i = 0; i < M_MAX_IDS; ++i) {           // ascending id, like enumerate(torch.unique(target_ids))
      if (!has_label[i] || !matched[i]) continue;
      const float d = __uint_as_float(maxd_bits[i]);
      sum = __fadd_rn(sum, d);
      ++found;
      if (d < 3.0f) ++good;                         // px_margin, metrics.py:46,97
    }
    const int nl = n_labels;
    p.valid[f] = nl > 0 ? 1 : 0;                    // no labels: the reference returns None and skips the sample
    p.l2[f] = nl > 0 ? __fdiv_rn(sum, (float)(found > 1 ? found : 1)) : 0.f;
    p.ratio[f] = nl > 0 ? __fdiv_rn((float)good, (float)nl) : 0.f;
  }
}

// Refinenet_Metrics.update (models/metrics.py:141-158): per sample the L2 distance, in heat-map pixels, between the arg-max of the
// predicted heat map and the arg-max of the target map (torch.argmax of the flattened map: first maximum).  One CTA per sample;
// every thread scans a strided slice keeping (value, lowest index), then a shared-memory tree reduction with the same tie rule.
constexpr int R_THREADS = 256;

__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__global__ void __launch_bounds__(R_THREADS)
heat_argmax_dist_kernel(const float* __restrict__ pred, const int32_t* __restrict__ pred_corners, const float* __restrict__ target,
                        int hw, int w, float* __restrict__ dist) {
  __shared__ float sv[2][R_THREADS];
  __shared__ int si[2][R_THREADS];
  const int f = blockIdx.x, tid = threadIdx.x;
  for (int which = 0; which < 2; ++which) {
    const float* src = which ? target + (size_t)f * hw : (pred ? pred + (size_t)f * hw : nullptr);
    float bv = 0.f; int bi = 0x7fffffff;
    if (src != nullptr)
      for (int i = tid; i < hw; i += R_THREADS) {
        const float v = src[i];
        if (bi == 0x7fffffff || v > bv) { bv = v; bi = i; }          // ascending i: strict '>' keeps the first maximum of this slice
      }
    sv[which][tid] = bv; si[which][tid] = bi;
  }
  __syncthreads();
  for (int s = R_THREADS / 2; s >= 1; s >>= 1) {
    if (tid < s)
      for (int which = 0; which < 2; ++which) {
        float v = sv[which][tid]; int i = si[which][tid];
        const int oi = si[which][tid + s];
        if (oi != 0x7fffffff) {
          if (i == 0x7fffffff) { v = sv[which][tid + s]; i = oi; }
          else argmax_merge(v, i, sv[which][tid + s], oi);
        }
        sv[which][tid] = v; si[which][tid] = i;
      }
    __syncthreads();
  }
  if (tid == 0) {
    int pr, pc;
    if (pred != nullptr) { pr = si[0][0] / w; pc = si[0][0] - pr * w; }
    else { pc = pred_corners[2 * f]; pr = pred_corners[2 * f + 1]; }     // (col, row) as RefineNet.infer_patches returns them
    const int tr = si[1][0] / w, tc = si[1][0] - tr * w;
    const float dr = (float)(pr - tr), dc = (float)(pc - tc);
    dist[f] = __fsqrt_rn(__fadd_rn(__fmul_rn(dr, dr), __fmul_rn(dc, dc)));
  }
}

}  // namespace

void launch_heat_argmax_dist(const float* pred, const int32_t* pred_corners, const float* target, int p, int h, int w, float* dist,
                             cudaStream_t s) {
  if (p <= 0) return;
  heat_argmax_dist_kernel<<<p, R_THREADS, 0, s>>>(pred, pred_corners, target, h * w, w, dist);
}

void launch_dc_metrics(const MetricsParams& p, cudaStream_t s) {
  if (p.n <= 0) return;
  dc_metrics_kernel<<<p.n, M_THREADS, 0, s>>>(p);
}

}  // namespace dcu
