// Synthetic ChArUco frames on the device: what the reference's training / validation data path does per
// sample on the host -- warp the rendered board (transformations.py:22-52 -> cv2.warpAffine / warpPerspective), paste it on a
// background (custom_aug.PasteBoard), blur, brightness, noise (transformations.py:105-114) -- as ONE kernel per batch, so that
// benchmarks and px-error evaluations are fed at engine speed.  Recipe and ranges: deepcharuco_b200/synth.py.
//
// Everything that needs a transcendental function or a linear solve (homographies, Gaussian taps) is a per-frame parameter
// computed on the host in float64 (deepcharuco_b200/synth.py: gpu_frame_params); the kernel does the per-pixel work with
//   * cv2.warpPerspective's own arithmetic for the board texture and its mask (u8, INTER_LINEAR, BORDER_CONSTANT): float64
//     coordinates evaluated per 64-column block exactly like warpPerspectiveInvoker, 1/32-pixel fixed point, 15-bit weights --
//     bit-exact with OpenCV (tests/test_gpu_synth.py compares dcu_warp_perspective_u8 with cv2 itself);
//   * counter-based random numbers (Philox4x32-10: lattice background, per-pixel noise), so a frame depends only on (seed, index);
//   * float32 operations in a fixed order with no fused multiply-adds, mirrored by oracle/synth.py bit for bit.
// Bound: HBM write of the u8 frame (76.8 kB); ~3 k integer / float operations per pixel, no global reads besides the 57 kB texture.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/deepcharuco_b200.h"
#include "common.cuh"

namespace dcu {

namespace {

constexpr int SY_R = 6;                       // blur radius (13 taps)
constexpr int SY_T = 32;                      // output tile
constexpr int SY_H = SY_T + 2 * SY_R;         // composite tile with halo
constexpr int STREAM_LATTICE = 2, STREAM_NOISE = 3;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// cv2.warpPerspective's fixed-point source coordinate of destination pixel (x, y): warpPerspectiveInvoker evaluates
// X0 = M0*xb + M1*y + M2 at the first column xb of a block of bw columns and adds M0*x1 inside the block (float64, no FMA).
__device__ __forceinline__ void warp_coord(const double* m, int x, int y, int bw, long long& X, long long& Y) {
  const int xb = (x / bw) * bw;
  const double dx = (double)xb, dy = (double)y, x1 = (double)(x - xb);
  const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(m[0], dx), __dmul_rn(m[1], dy)), m[2]);
  const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(m[3], dx), __dmul_rn(m[4], dy)), m[5]);
  const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(m[6], dx), __dmul_rn(m[7], dy)), m[8]);
  double W = __dadd_rn(W0, __dmul_rn(m[6], x1));
  W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(X0, __dmul_rn(m[0], x1)), W)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, __dmul_rn(__dadd_rn(Y0, __dmul_rn(m[3], x1)), W)));
  X = __double2ll_rn(fX); Y = __double2ll_rn(fY);       // saturate_cast<int>(double) = round half to even
}

// remapBilinear (u8, BORDER_CONSTANT 0) at fixed-point (X, Y); src == nullptr: a virtual image that is 255 inside [0,sw) x [0,sh)
__device__ __forceinline__ int warp_sample(const uint8_t* __restrict__ src, int sh, int sw, long long X, long long Y) {
  long long sxl = X >> 5, syl = Y >> 5;
  sxl = sxl < -32768 ? -32768 : (sxl > 32767 ? 32767 : sxl);       // saturate_cast<short>
  syl = syl < -32768 ? -32768 : (syl > 32767 ? 32767 : syl);
  const int sx = (int)sxl, sy = (int)syl, ax = (int)(X & 31), ay = (int)(Y & 31);
  // BilinearTab_i: (32-ay)(32-ax)*32 ... exact integers; alpha == 0 saturates to 32767 and OpenCV's fix-up puts the missing 1 on entry 3
  int w0 = (32 - ay) * (32 - ax) * 32, w1 = (32 - ay) * ax * 32, w2 = ay * (32 - ax) * 32, w3 = ay * ax * 32;
  if ((ax | ay) == 0) { w0 = 32767; w3 = 1; }
  auto tap = [&](int yy, int xx) -> int {
    if (xx < 0 || xx >= sw || yy < 0 || yy >= sh) return 0;
    return src ? (int)src[yy * sw + xx] : 255;
  };
  const int acc = tap(sy, sx) * w0 + tap(sy, sx + 1) * w1 + tap(sy + 1, sx) * w2 + tap(sy + 1, sx + 1) * w3;
  const int v = (acc + (1 << 14)) >> 15;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

__device__ __forceinline__ int reflect101(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * n - 2 - i : i;
}

__global__ void synth_lattice_kernel(const DcuSynthFrame* __restrict__ fp, int n, int first_index, uint32_t k0, uint32_t k1, int lat_cap,
                                     uint8_t* __restrict__ lat) {
  const int f = blockIdx.y;
  const DcuSynthFrame& P = fp[f];
  const int cnt = P.lat_h * P.lat_w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint32_t)i, (uint32_t)(first_index + f), STREAM_LATTICE, 0u, k0, k1, r);
    lat[(size_t)f * lat_cap + i] = (uint8_t)(r[0] >> 24);
  }
}

__global__ void __launch_bounds__(256)
synth_frame_kernel(const DcuSynthFrame* __restrict__ fp, const uint8_t* __restrict__ board, int bpx, const uint8_t* __restrict__ lat, int lat_cap,
                   int first_index, uint32_t k0, uint32_t k1, int H, int W, uint8_t* __restrict__ frames) {
  __shared__ float comp[SY_H][SY_H + 1];
  __shared__ float hb[SY_H][SY_T + 1];
  __shared__ DcuSynthFrame P;
  const int f = blockIdx.z, tid = threadIdx.x;
  if (tid < (int)(sizeof(DcuSynthFrame) / 4)) reinterpret_cast<uint32_t*>(&P)[tid] = reinterpret_cast<const uint32_t*>(&fp[f])[tid];
  __syncthreads();
  const int x0 = blockIdx.x * SY_T, y0 = blockIdx.y * SY_T;
  const uint8_t* L = lat + (size_t)f * lat_cap;
  const int S = P.lat_step, lw = P.lat_w;
  const float fS = (float)S;
  const int bw = (1024 / (H < 16 ? H : 16)) < W ? (1024 / (H < 16 ? H : 16)) : W;       // warpPerspective's block width (64 for real frames)
  // ---- composite (background + boards) on the tile + halo, at reflected coordinates ----
  for (int i = tid; i < SY_H * SY_H; i += 256) {
    const int ty = i / SY_H, tx = i - ty * SY_H;
    const int y = reflect101(y0 - SY_R + ty, H), x = reflect101(x0 - SY_R + tx, W);
    const int gy = y / S, gx = x / S;
    const float fy = __fdiv_rn((float)(y - gy * S), fS), fx = __fdiv_rn((float)(x - gx * S), fS);
    const float a = (float)L[gy * lw + gx], b = (float)L[gy * lw + gx + 1], c = (float)L[(gy + 1) * lw + gx], d = (float)L[(gy + 1) * lw + gx + 1];
    const float top = __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), fx));
    const float bot = __fadd_rn(c, __fmul_rn(__fsub_rn(d, c), fx));
    const float t = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), fy));
    float v = __fadd_rn(P.bg_lo, __fmul_rn(__fsub_rn(P.bg_hi, P.bg_lo), __fmul_rn(t, 1.0f / 255.0f)));
    for (int bi = 0; bi < P.n_boards; ++bi) {
      long long X, Y;
      warp_coord(P.hinv[bi], x, y, bw, X, Y);
      const float wv = (float)warp_sample(board, bpx, bpx, X, Y);
      const float mk = __fmul_rn((float)warp_sample(nullptr, bpx, bpx, X, Y), 1.0f / 255.0f);
      v = __fadd_rn(__fmul_rn(v, __fsub_rn(1.0f, mk)), __fmul_rn(wv, mk));
    }
    comp[ty][tx] = v;
  }
  __syncthreads();
  // ---- separable 13-tap blur: rows, then columns; taps accumulated in order from 0 ----
  for (int i = tid; i < SY_H * SY_T; i += 256) {
    const int ty = i / SY_T, tx = i - ty * SY_T;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * SY_R + 1; ++k) acc = __fadd_rn(acc, __fmul_rn(P.blur_w[k], comp[ty][tx + k]));
    hb[ty][tx] = acc;
  }
  __syncthreads();
  for (int i = tid; i < SY_T * SY_T; i += 256) {
    const int ty = i / SY_T, tx = i - ty * SY_T;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 2 * SY_R + 1; ++k) acc = __fadd_rn(acc, __fmul_rn(P.blur_w[k], hb[ty + k][tx]));
    uint32_t r[4];
    philox4x32_10((uint32_t)(y * W + x), (uint32_t)(first_index + f), STREAM_NOISE, 0u, k0, k1, r);
    const int s4 = (int)((r[0] & 255u) + ((r[0] >> 8) & 255u) + ((r[0] >> 16) & 255u) + (r[0] >> 24));      // Irwin-Hall, sigma 147.8
    const float noise = __fmul_rn((float)(s4 - 510), (float)(3.0 / 147.8005413));
    const float o = rintf(__fadd_rn(__fmul_rn(acc, P.gain), noise));
    frames[((size_t)f * H + y) * W + x] = (uint8_t)fminf(fmaxf(o, 0.f), 255.f);
  }
}

__global__ void warp_perspective_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, const double* __restrict__ minv, uint8_t* __restrict__ dst,
                                           int H, int W) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = minv[i];
  const int bh0 = H < 16 ? H : 16;
  const int bw = (1024 / bh0) < W ? (1024 / bh0) : W;
  long long X, Y;
  warp_coord(m, x, y, bw, X, Y);
  dst[(size_t)y * W + x] = (uint8_t)warp_sample(src, sh, sw, X, Y);
}

}  // namespace

cudaError_t launch_synth_frames(const DcuSynthFrame* params_dev, const uint8_t* board_dev, int board_px, uint8_t* lattice_dev, int lat_cap,
                                uint64_t seed, int first_index, int n, int H, int W, uint8_t* frames_dev, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  synth_lattice_kernel<<<dim3(8, n), 256, 0, s>>>(params_dev, n, first_index, k0, k1, lat_cap, lattice_dev);
  dim3 grid((W + SY_T - 1) / SY_T, (H + SY_T - 1) / SY_T, n);
  synth_frame_kernel<<<grid, 256, 0, s>>>(params_dev, board_dev, board_px, lattice_dev, lat_cap, first_index, k0, k1, H, W, frames_dev);
  return cudaGetLastError();
}

cudaError_t launch_warp_perspective_u8(const uint8_t* src, int sh, int sw, const double* minv_dev, uint8_t* dst, int H, int W, cudaStream_t s) {
  warp_perspective_u8_kernel<<<dim3((W + 127) / 128, H), 128, 0, s>>>(src, sh, sw, minv_dev, dst, H, W);
  return cudaGetLastError();
}

}  // namespace dcu
