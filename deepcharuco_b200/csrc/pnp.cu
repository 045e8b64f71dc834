// Batched board pose kernel: one warp per frame over the engine's result buffers (lanes share the per-corner loops).  The algorithm (a restatement of
// cv2.solvePnP, SOLVEPNP_ITERATIVE, for the planar board) is in pnp_core.cuh.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "pnp_core.cuh"

namespace dcu {

namespace {

__global__ void pnp_batch_kernel(PnpParams q, pnp::Cam cam) {
  const int f = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (f >= q.n) return;                                   // whole warp
  pnp::Pts P;
  const int off = q.offsets[f];
  // decode_gather writes the full counts / offsets even when it dropped rows at max_patches: never read past the buffer
  P.n = max(0, min(q.counts[f], q.max_rows - off)); P.n_obj = q.n_obj; P.obj = q.obj;
  P.kp = q.kpts + 4 * (size_t)off;
  P.xy = q.refined ? q.refined + 2 * (size_t)off : nullptr;
  double rv[3], tv[3];
  const int ret = pnp::solve_frame(pnp::WarpLanes(), cam, P, rv, tv);
  if ((threadIdx.x & 31u) == 0) {
    q.ret[f] = ret;
    for (int k = 0; k < 3; ++k) { q.rvec[3 * (size_t)f + k] = rv[k]; q.tvec[3 * (size_t)f + k] = tv[k]; }
  }
}

}  // namespace

void launch_pnp_batch(const PnpParams& q, const double* camera9, const double* dist, int n_dist, cudaStream_t s) {
  if (q.n <= 0) return;
  pnp::Cam c{};
  c.fx = camera9[0]; c.cx = camera9[2]; c.fy = camera9[4]; c.cy = camera9[5];
  for (int i = 0; i < 8; ++i) c.k[i] = (dist && i < n_dist) ? dist[i] : 0.0;
  pnp_batch_kernel<<<(q.n + 1) / 2, 64, 0, s>>>(q, c);      // 2 frames (warps) per block: spreads a batch over all SMs
}

}  // namespace dcu
