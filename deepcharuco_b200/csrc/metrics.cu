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

// ---- utils.pixel_error (/root/reference/src/utils.py:33-52), batched -----------------------------------------------------------
// Per frame: raw corners (x, y, id: the engine's integer pixels), refined corners (x, y), labels (x, y, id as float64), and
//   d         = compute_l2_distance(raw, labels)      d_ref = compute_l2_distance(refined, labels)      d_rr = compute_l2_distance(refined, raw)
// with compute_l2_distance (utils.py:6-30): distances = zeros(len(target_ids)); for i, id in enumerate(unique(target_ids)): the largest
// float64 distance between the predictions with that id and the target(s) with that id (numpy broadcasting of (m,2) - (t,2): needs
// m == t, m == 1 or t == 1 -- anything else raises in the reference and is reported as status -1 here); ids nobody predicted keep 0.
// out[f] = {mean d, mean d_ref, mean d_rr, max d, max d_ref, max d_rr}.  status[f]: 1 evaluated; 0 skipped like the caller / the
// function do (no labels, no predictions, or a predicted id that is not among the labels: utils.py:34-35); -1 shapes numpy rejects.
// All float64 with separately rounded multiplies / adds (numpy has no FMA contraction) and numpy's pairwise summation order, so the
// numbers are bit-identical with the reference's.  One thread per frame: a frame has a few dozen corners.
constexpr int PE_MAX_IDS = 64;
constexpr int PE_MAX_ROWS = 256;

__device__ double np_pairwise_sum(const double* a, int n) {          // numpy's DOUBLE_pairwise_sum for n <= 128 (one block)
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
    return r;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, a[i]);
  return res;
}

struct PeSet { const double* x; const double* y; const int* id; int n; };    // a list of (x, y, id) rows in local memory

// distances[] (length tgt.n) as utils.compute_l2_distance; returns false on a broadcast numpy would reject
__device__ bool pe_l2(const PeSet& pred, const PeSet& tgt, double* distances) {
  for (int i = 0; i < tgt.n; ++i) distances[i] = 0.0;
  int slot = 0;
  for (int idv = 0; idv < PE_MAX_IDS; ++idv) {                      // ascending id = np.unique order
    int t = 0, m = 0;
    for (int j = 0; j < tgt.n; ++j) t += tgt.id[j] == idv;
    if (t == 0) continue;
    for (int j = 0; j < pred.n; ++j) m += pred.id[j] == idv;
    const int i = slot++;
    if (m == 0) continue;
    if (!(m == t || m == 1 || t == 1)) return false;
    const int pairs = m > t ? m : t;
    double mx = 0.0;
    int jp = -1, jt = -1;
    for (int k = 0; k < pairs; ++k) {
      if (k == 0 || m > 1) { do { ++jp; } while (pred.id[jp] != idv); }
      if (k == 0 || t > 1) { do { ++jt; } while (tgt.id[jt] != idv); }
      const double dx = __dsub_rn(pred.x[jp], tgt.x[jt]), dy = __dsub_rn(pred.y[jp], tgt.y[jt]);
      const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      if (k == 0 || d > mx) mx = d;
    }
    distances[i] = mx;
  }
  return true;
}

__global__ void pixel_error_kernel(PixelErrorParams p) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.n) return;
  double* out = p.out + 6 * (size_t)f;
  for (int k = 0; k < 6; ++k) out[k] = 0.0;
  const int off = p.offsets[f];
  const int K = max(0, min(p.counts[f], p.max_rows - off));
  const int T = p.tcounts[f];
  if (K == 0 || T == 0) { p.status[f] = 0; return; }               // `if len(label_kpts) != 0 and len(keypoints) != 0` (inference.py:154)
  if (K > PE_MAX_ROWS || T > PE_MAX_ROWS) { p.status[f] = -1; return; }
  double rx[PE_MAX_ROWS], ry[PE_MAX_ROWS], fx[PE_MAX_ROWS], fy[PE_MAX_ROWS], tx[PE_MAX_ROWS], ty[PE_MAX_ROWS], dist[PE_MAX_ROWS];
  int rid[PE_MAX_ROWS], tid[PE_MAX_ROWS];
  const double* tg = p.target + 3 * (size_t)p.toffsets[f];
  bool ok = true;
  for (int j = 0; j < T; ++j) {
    tx[j] = tg[3 * j]; ty[j] = tg[3 * j + 1];
    const double idd = tg[3 * j + 2];
    tid[j] = (int)idd;
    ok = ok && idd >= 0.0 && idd < (double)PE_MAX_IDS && (double)tid[j] == idd;
  }
  for (int j = 0; j < K; ++j) {
    const int4 r = reinterpret_cast<const int4*>(p.kpts)[off + j];
    rx[j] = (double)r.x; ry[j] = (double)r.y; rid[j] = r.z;
    fx[j] = (double)p.refined[2 * (size_t)(off + j)]; fy[j] = (double)p.refined[2 * (size_t)(off + j) + 1];
    ok = ok && r.z >= 0 && r.z < PE_MAX_IDS;
  }
  if (!ok) { p.status[f] = -1; return; }
  for (int j = 0; j < K; ++j) {                                    // set(raw ids).issubset(set(label ids)), utils.py:34
    bool found = false;
    for (int i = 0; i < T; ++i) found = found || tid[i] == rid[j];
    if (!found) { p.status[f] = 0; return; }
  }
  const PeSet raw{rx, ry, rid, K}, ref{fx, fy, rid, K}, lab{tx, ty, tid, T};
  struct { const PeSet* a; const PeSet* b; } legs[3] = {{&raw, &lab}, {&ref, &lab}, {&ref, &raw}};
  for (int leg = 0; leg < 3; ++leg) {
    if (!pe_l2(*legs[leg].a, *legs[leg].b, dist)) { p.status[f] = -1; return; }
    const int n = legs[leg].b->n;
    double mx = dist[0];
    for (int i = 1; i < n; ++i) mx = dist[i] > mx ? dist[i] : mx;
    out[leg] = __ddiv_rn(np_pairwise_sum(dist, n), (double)n);      // ndarray.mean(): pairwise add.reduce / n
    out[3 + leg] = mx;
  }
  p.status[f] = 1;
}

}  // namespace

void launch_pixel_error(const PixelErrorParams& p, cudaStream_t s) {
  if (p.n <= 0) return;
  pixel_error_kernel<<<(p.n + 31) / 32, 32, 0, s>>>(p);
}

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
