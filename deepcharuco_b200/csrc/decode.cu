// Heat-map decode + corner-patch gather in ONE bandwidth-bound kernel.
//
// Per frame (one CTA): per-cell arg-max over the 65 loc and n_ids+1 ids logits (first max wins),
// dustbin filter, cell -> pixel, ordered compaction, then the 24x24 patch gather with implicit zero
// padding.  Restates, for a whole batch and without a host round trip:
//   pred_argmax          /root/reference/src/models/model_utils.py:72-78
//   label_to_keypoints   model_utils.py:108-124   (x = 8*col + p%8, y = 8*row + p//8)
//   extract_patches      model_utils.py:19-36     (pad 12 with 0.0 of the normalised image)
//   sort by id (stable)  /root/reference/src/inference.py:68-69
// Memory behaviour: the logits are read once, channel-strided so consecutive threads read consecutive
// cells of one channel plane (fully coalesced 128 B lines); patches are written as coalesced fp32 rows.
// Frames get their output rows by a chained (look-back) scan over frame index, so row numbering is
// deterministic and needs no second launch.
#include "common.cuh"

namespace dcu {

constexpr int D_THREADS = 320;   // 300 four-cell quads of a 30x40 cell grid fit one pass

size_t decode_smem_bytes(int cells) { return (size_t)cells * 4 * sizeof(uint32_t); }

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(D_THREADS)
decode_gather_kernel(DecodeParams p) {
  extern __shared__ __align__(16) uint32_t dsm[];
  const int cells = p.h * p.w;
  uint32_t* key_u = dsm;                 // unsorted keys  (id * cells + cell)
  uint32_t* pix_u = dsm + cells;         // unsorted sub-cell pixel index 0..63
  uint32_t* key_s = dsm + 2 * cells;     // sorted keys
  uint32_t* xy_s = dsm + 3 * cells;      // sorted x | y << 16
  __shared__ int s_count;
  __shared__ int s_base;
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) s_count = 0;
  __syncthreads();

  // ---- phase 1: per-cell arg-max + dustbin filter ------------------------------------------------
  // Four consecutive cells per thread, one 16-byte load per logit plane: consecutive threads read consecutive 16 B of a
  // plane (fully coalesced) and every thread keeps 8 planes x 16 B in flight, which is what it takes to cover HBM latency
  // with one CTA per frame.  Strict '>' keeps the FIRST maximum, as torch.argmax does.
  const float* loc = p.loc + (size_t)f * 65 * cells;
  const float* ids = p.ids + (size_t)f * p.n_ids1 * cells;
  if (p.loc_arg != nullptr) {
    // the arg-maxes were taken in the 1x1 head epilogues (conv_tc.cu, on the very logit values this phase would read): 2 bytes per cell
    const uint8_t* la_map = p.loc_arg + (size_t)f * cells;
    const uint8_t* ia_map = p.ids_arg + (size_t)f * cells;
    for (int c = tid; c < cells; c += D_THREADS) {
      const int la = la_map[c];
      const int id = (la == 64) ? p.dust_bin : (int)ia_map[c];      // model_utils.py:77
      if (id != p.dust_bin) {                                       // model_utils.py:111
        const int slot = atomicAdd(&s_count, 1);
        key_u[slot] = (uint32_t)id * (uint32_t)cells + (uint32_t)c;
        pix_u[slot] = (uint32_t)la;
      }
    }
  } else if ((cells & 3) == 0) {
    const int quads = cells >> 2;
    for (int q = tid; q < quads; q += D_THREADS) {
      const float4* lp = reinterpret_cast<const float4*>(loc) + q;
      float4 best = lp[0];
      int la[4] = {0, 0, 0, 0};
#pragma unroll 8
      for (int ch = 1; ch < 65; ++ch) {
        const float4 v = lp[(size_t)ch * quads];
        if (v.x > best.x) { best.x = v.x; la[0] = ch; }
        if (v.y > best.y) { best.y = v.y; la[1] = ch; }
        if (v.z > best.z) { best.z = v.z; la[2] = ch; }
        if (v.w > best.w) { best.w = v.w; la[3] = ch; }
      }
      const float4* ip = reinterpret_cast<const float4*>(ids) + q;
      float4 bi = ip[0];
      int ia[4] = {0, 0, 0, 0};
#pragma unroll 4
      for (int ch = 1; ch < p.n_ids1; ++ch) {
        const float4 v = ip[(size_t)ch * quads];
        if (v.x > bi.x) { bi.x = v.x; ia[0] = ch; }
        if (v.y > bi.y) { bi.y = v.y; ia[1] = ch; }
        if (v.z > bi.z) { bi.z = v.z; ia[2] = ch; }
        if (v.w > bi.w) { bi.w = v.w; ia[3] = ch; }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int id = (la[e] == 64) ? p.dust_bin : ia[e];     // model_utils.py:77 (loc dustbin hard-coded 64)
        if (id != p.dust_bin) {                          // model_utils.py:111
          const int slot = atomicAdd(&s_count, 1);
          key_u[slot] = (uint32_t)id * (uint32_t)cells + (uint32_t)(4 * q + e);
          pix_u[slot] = (uint32_t)la[e];
        }
      }
    }
  } else {
    for (int c = tid; c < cells; c += D_THREADS) {
      float best = loc[c];
      int la = 0;
#pragma unroll 8
      for (int ch = 1; ch < 65; ++ch) {
        const float v = loc[(size_t)ch * cells + c];
        if (v > best) { best = v; la = ch; }
      }
      float bi = ids[c];
      int ia = 0;
      for (int ch = 1; ch < p.n_ids1; ++ch) {
        const float v = ids[(size_t)ch * cells + c];
        if (v > bi) { bi = v; ia = ch; }
      }
      if (la == 64) ia = p.dust_bin;
      if (ia != p.dust_bin) {
        const int slot = atomicAdd(&s_count, 1);
        key_u[slot] = (uint32_t)ia * (uint32_t)cells + (uint32_t)c;
        pix_u[slot] = (uint32_t)la;
      }
    }
  }
  __syncthreads();
  const int K = s_count;

  // ---- phase 2: rank sort by (id, cell); keys are unique -----------------------------------------
  for (int i = tid; i < K; i += D_THREADS) {
    const uint32_t k = key_u[i];
    int rank = 0;
    for (int j = 0; j < K; ++j) rank += (key_u[j] < k) ? 1 : 0;
    const int cell = (int)(k % (uint32_t)cells);
    const int cx = cell % p.w, cy = cell / p.w;
    const uint32_t pp = pix_u[i];
    const uint32_t x = 8u * cx + (pp & 7u), y = 8u * cy + (pp >> 3);
    key_s[rank] = k;
    xy_s[rank] = x | (y << 16);
  }

  // ---- phase 3: chained scan over frames -> base row ---------------------------------------------
  if (tid == 0) {
    const unsigned long long tag_agg = ((unsigned long long)(p.epoch * 4u + 1u)) << 32;
    const unsigned long long tag_inc = ((unsigned long long)(p.epoch * 4u + 2u)) << 32;
    int excl = 0;
    if (f == 0) {
      excl = p.append ? p.total[0] : 0;
    } else {
      atomicExch(&p.scan_state[f], tag_agg | (unsigned int)K);
      int look = f - 1;
      while (true) {
        const unsigned long long st = ld_volatile_u64(&p.scan_state[look]);
        const unsigned long long tag = st & 0xffffffff00000000ull;
        if (tag == tag_inc) { excl += (int)(unsigned int)st; break; }
        if (tag == tag_agg) { excl += (int)(unsigned int)st; --look; continue; }   // look >= 1 here: frame 0 only publishes inclusive
        __nanosleep(40);
      }
    }
    __threadfence();
    atomicExch(&p.scan_state[f], tag_inc | (unsigned int)(excl + K));
    p.counts[f] = K;
    p.offsets[f] = excl;
    if (f == p.n - 1) p.total[0] = excl + K;
    s_base = excl;
  }
  __syncthreads();
  const int base = s_base;

  // ---- phase 4: keypoint rows --------------------------------------------------------------------
  for (int j = tid; j < K; j += D_THREADS) {
    const int row = base + j;
    if (row < p.max_patches) {
      const uint32_t k = key_s[j], xy = xy_s[j];
      int4 rec;
      rec.x = (int)(xy & 0xffffu); rec.y = (int)(xy >> 16);
      rec.z = (int)(k / (uint32_t)cells); rec.w = (int)(k % (uint32_t)cells);
      reinterpret_cast<int4*>(p.kpts)[row] = rec;
    }
  }

  // ---- phase 5: 24x24 patch gather (window [y-12, y+12) x [x-12, x+12), zeros outside) -----------
  if (p.patches != nullptr) {
    const uint8_t* frame = p.frames + (size_t)f * p.H * p.W;
    const int total_el = K * 576;
    for (int e = tid; e < total_el; e += D_THREADS) {
      const int j = e / 576, q = e - j * 576;
      const int row = base + j;
      if (row >= p.max_patches) break;
      const uint32_t xy = xy_s[j];
      const int py = q / 24, px = q - py * 24;
      const int gy = (int)(xy >> 16) - 12 + py, gx = (int)(xy & 0xffffu) - 12 + px;
      float v = 0.f;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) v = __ldg(p.lut + frame[(size_t)gy * p.W + gx]);
      p.patches[(size_t)row * 576 + q] = v;
    }
  }
}

void launch_decode_gather(const DecodeParams& p, cudaStream_t s) {
  if (p.n <= 0) return;
  const size_t smem = decode_smem_bytes(p.h * p.w);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(decode_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  decode_gather_kernel<<<p.n, D_THREADS, smem, s>>>(p);
}

// BGR -> gray on the device (reference: cv2.cvtColor(img, COLOR_BGR2GRAY), inference.py:40).
// OpenCV's 8-bit path is the fixed-point form  (3735*B + 19235*G + 9798*R + 2^14) >> 15 ; verified against cv2 4.13 on all
// 2^24 colours (tests/test_host_logic.py).  One thread converts 4 pixels: three 4-byte loads, one 4-byte store.
__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray, long long n_px) {
  const long long n4 = n_px >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(bgr) + i * 3;
    const uint32_t w0 = src[0], w1 = src[1], w2 = src[2];     // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    const uint32_t b0 = w0 & 255u, g0 = (w0 >> 8) & 255u, r0 = (w0 >> 16) & 255u;
    const uint32_t b1 = w0 >> 24, g1 = w1 & 255u, r1 = (w1 >> 8) & 255u;
    const uint32_t b2 = (w1 >> 16) & 255u, g2 = w1 >> 24, r2 = w2 & 255u;
    const uint32_t b3 = (w2 >> 8) & 255u, g3 = (w2 >> 16) & 255u, r3 = w2 >> 24;
    const uint32_t y0 = (3735u * b0 + 19235u * g0 + 9798u * r0 + 16384u) >> 15;
    const uint32_t y1 = (3735u * b1 + 19235u * g1 + 9798u * r1 + 16384u) >> 15;
    const uint32_t y2 = (3735u * b2 + 19235u * g2 + 9798u * r2 + 16384u) >> 15;
    const uint32_t y3 = (3735u * b3 + 19235u * g3 + 9798u * r3 + 16384u) >> 15;
    reinterpret_cast<uint32_t*>(gray)[i] = y0 | (y1 << 8) | (y2 << 16) | (y3 << 24);
  }
}
// cv2.resize(src, (W, H), interpolation=cv2.INTER_LINEAR) for uint8 images when shrinking (the reference's evaluation loop resizes the
// camera frame to the network input, inference.py:131-132), bit-exact with OpenCV's fixed-point path: 11-bit coefficients from float32
// fractions (tables built on the host exactly like cv::resize builds them), horizontal pass in int, vertical pass
//   dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// One thread per destination pixel (all channels); tab = [sx | ax0 | ax1] (W entries each) then [sy | ay0 | ay1] (H entries each).
__global__ void resize_linear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const int* __restrict__ tab, int n, int hs, int ws,
                                        int h, int w, int ch) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * h * w) return;
  const int x = (int)(i % w), y = (int)((i / w) % h);
  const long long f = i / ((long long)w * h);
  const int* tx = tab; const int* ty = tab + 3 * w;
  const int sx = tx[x], a0 = tx[w + x], a1 = tx[2 * w + x], sy = ty[y], b0 = ty[h + y], b1 = ty[2 * h + y];
  const int sx1 = min(sx + 1, ws - 1), sy1 = min(sy + 1, hs - 1);
  const uint8_t* r0 = src + ((size_t)f * hs + sy) * ws * ch;
  const uint8_t* r1 = src + ((size_t)f * hs + sy1) * ws * ch;
  uint8_t* o = dst + (size_t)i * ch;
  for (int c = 0; c < ch; ++c) {
    const int h0 = (int)r0[sx * ch + c] * a0 + (int)r0[sx1 * ch + c] * a1;
    const int h1 = (int)r1[sx * ch + c] * a0 + (int)r1[sx1 * ch + c] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

void launch_resize_linear_u8(const uint8_t* src, uint8_t* dst, const int* tab_dev, int n, int hs, int ws, int h, int w, int ch, cudaStream_t s) {
  const long long total = (long long)n * h * w;
  if (total <= 0) return;
  resize_linear_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(src, dst, tab_dev, n, hs, ws, h, w, ch);
}

void launch_bgr_to_gray(const uint8_t* bgr, uint8_t* gray, long long n_px, cudaStream_t s) {
  if (n_px <= 0) return;
  long long b = ((n_px >> 2) + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  bgr_to_gray_kernel<<<(int)b, 256, 0, s>>>(bgr, gray, n_px);
}

// stand-alone extract_patches on a normalised fp32 image (model_utils.py:19-36)
__global__ void extract_patches_kernel(const float* image, int H, int W, const int32_t* xy, int k, float* patches) {
  const int j = blockIdx.x;
  if (j >= k) return;
  const int x = xy[2 * j], y = xy[2 * j + 1];
  for (int q = threadIdx.x; q < 576; q += blockDim.x) {
    const int py = q / 24, px = q - py * 24;
    const int gy = y - 12 + py, gx = x - 12 + px;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = image[(size_t)gy * W + gx];
    patches[(size_t)j * 576 + q] = v;
  }
}
void launch_extract_patches(const float* image, int H, int W, const int32_t* xy, int k, float* patches, cudaStream_t s) {
  if (k <= 0) return;
  extract_patches_kernel<<<k, 192, 0, s>>>(image, H, W, xy, k, patches);
}

}  // namespace dcu
