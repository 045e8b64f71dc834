// fp32 CUDA-core kernels: first-layer conv (cin = 1), generic 3x3 conv + BN + ReLU (+pool / +2x up /
// +fused RefineNet head), detector 1x1 heads, layout converters.
//
// Reference semantics implemented here:
//   conv -> BatchNorm2d(eval) -> ReLU           net.py:60-77, refinenet.py:56-80
//   MaxPool2d(2,2)                              net.py:62,65,68; refinenet.py:62
//   UpsamplingNearest2d(2)                      refinenet.py:67,72,77
//   convPb 1x1 + flat arg-max (first max wins)  refinenet.py:81,111; model_utils.py:39-43
// The accumulation is plain fp32 FMA (no tensor cores): this is the strict-fp32 path and the on-GPU
// cross-check for the tcgen05 path in conv_tc.cu.
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace dcu {

// ---------------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;   // src-size 0 => 16 bytes of zeros (conv zero padding / tile overhang)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float bn_relu(float acc, float bias, float alpha, float beta) {
  // relu(fma(acc + bias, alpha, beta)): reproduces ATen's eval BatchNorm2d bit-for-bit given the same
  // conv output; conv bias is added first, as F.conv2d does.
  return fmaxf(fmaf(acc + bias, alpha, beta), 0.0f);
}

// x -> (fp16(x), fp16(x - fp16(x))) for two values at once; inputs are post-ReLU (>= 0), clamped to fp16's finite range
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  x0 = fminf(x0, 65504.f); x1 = fminf(x1, 65504.f);
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}

// monotone map float -> uint32 (larger float => larger key); used for the packed arg-max key
__device__ __forceinline__ unsigned int orderable(float v) {
  unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ---------------------------------------------------------------------------------------------------
// generic 3x3 convolution, C4 layout, 16x8 output pixels x 64 output channels per work item
// ---------------------------------------------------------------------------------------------------
constexpr int F_TH = 16, F_TW = 8, F_HH = F_TH + 2, F_HW = F_TW + 2;
constexpr int F_THREADS = 128;
constexpr int F_CB = 64;

__global__ void __launch_bounds__(F_THREADS)
conv3x3_ffma_kernel(ConvParams p, const float* __restrict__ wpk, int tiles_x, int tiles_y, int cblocks,
                    long long total_items) {
  __shared__ __align__(16) float4 in_s[2][F_HH * F_HW];
  __shared__ __align__(16) float w_s[2][9 * 4 * F_CB];

  const int tid = threadIdx.x;
  const int cg = tid & 7;       // channel sub-group: channels {4cg..4cg+3} and {32+4cg..32+4cg+3} of the 64-block
  const int r = tid >> 3;       // tile row 0..15 (8 pixels of that row per thread)
  const int groups = p.cin >> 2;
  const int cgroups_out = p.cout_total >> 2;

  for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
    int cb = (int)(item % cblocks);
    long long t = item / cblocks;
    int tx = (int)(t % tiles_x); t /= tiles_x;
    int ty = (int)(t % tiles_y);
    int img = (int)(t / tiles_y);
    const int y0 = ty * F_TH, x0 = tx * F_TW;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    auto load_chunk = [&](int g, int buf) {
      const float4* src = reinterpret_cast<const float4*>(p.in) + ((size_t)img * groups + g) * p.hin * p.win;
      for (int i = tid; i < F_HH * F_HW; i += F_THREADS) {
        int hy = i / F_HW, hx = i - hy * F_HW;
        int gy = y0 - p.pad + hy, gx = x0 - p.pad + hx;
        bool ok = (gy >= 0) && (gy < p.hin) && (gx >= 0) && (gx < p.win);
        const float4* s = ok ? (src + (size_t)gy * p.win + gx) : src;
        cp_async16(&in_s[buf][i], s, ok);
      }
      const float* wsrc = wpk + (size_t)g * 36 * p.cout_total + cb * F_CB;
      for (int i = tid; i < 36 * 16; i += F_THREADS) {
        int row = i >> 4, q = i & 15;
        cp_async16(&w_s[buf][row * F_CB + q * 4], wsrc + (size_t)row * p.cout_total + q * 4, true);
      }
    };

    load_chunk(0, 0);
    cp_async_commit();
    for (int g = 0; g < groups; ++g) {
      const int buf = g & 1;
      if (g + 1 < groups) {
        load_chunk(g + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float4* in_t = in_s[buf];
      const float* w_t = w_s[buf];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float4 iv[F_HW];
#pragma unroll
        for (int j = 0; j < F_HW; ++j) iv[j] = in_t[(r + ky) * F_HW + j];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int tap = ky * 3 + kx;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float4 wa = *reinterpret_cast<const float4*>(&w_t[(tap * 4 + c) * F_CB + cg * 4]);
            const float4 wb = *reinterpret_cast<const float4*>(&w_t[(tap * 4 + c) * F_CB + 32 + cg * 4]);
#pragma unroll
            for (int px = 0; px < 8; ++px) {
              const float4 q = iv[px + kx];
              const float v = (c == 0) ? q.x : (c == 1) ? q.y : (c == 2) ? q.z : q.w;
              acc[px][0] = fmaf(v, wa.x, acc[px][0]);
              acc[px][1] = fmaf(v, wa.y, acc[px][1]);
              acc[px][2] = fmaf(v, wa.z, acc[px][2]);
              acc[px][3] = fmaf(v, wa.w, acc[px][3]);
              acc[px][4] = fmaf(v, wb.x, acc[px][4]);
              acc[px][5] = fmaf(v, wb.y, acc[px][5]);
              acc[px][6] = fmaf(v, wb.z, acc[px][6]);
              acc[px][7] = fmaf(v, wb.w, acc[px][7]);
            }
          }
        }
      }
      __syncthreads();
    }

    // ---------------- epilogue ----------------
    const int ch0 = cb * F_CB + cg * 4;   // first quad; second quad at +32
    const int oy = y0 + r;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ch = ch0 + 32 * h;
      const float4 bi = *reinterpret_cast<const float4*>(p.bias + ch);
      const float4 al = *reinterpret_cast<const float4*>(p.alpha + ch);
      const float4 be = *reinterpret_cast<const float4*>(p.beta + ch);
#pragma unroll
      for (int px = 0; px < 8; ++px) {
        acc[px][4 * h + 0] = bn_relu(acc[px][4 * h + 0], bi.x, al.x, be.x);
        acc[px][4 * h + 1] = bn_relu(acc[px][4 * h + 1], bi.y, al.y, be.y);
        acc[px][4 * h + 2] = bn_relu(acc[px][4 * h + 2], bi.z, al.z, be.z);
        acc[px][4 * h + 3] = bn_relu(acc[px][4 * h + 3], bi.w, al.w, be.w);
      }
    }

    if (p.head_w != nullptr) {
      // fused RefineNet head: heat = convPb(relu(bn(convPa))) then flat arg-max (cout_total == 64, cb == 0)
      const float4 ha = *reinterpret_cast<const float4*>(p.head_w + cg * 4);
      const float4 hb = *reinterpret_cast<const float4*>(p.head_w + 32 + cg * 4);
      unsigned long long best = 0ull;
#pragma unroll
      for (int px = 0; px < 8; ++px) {
        float s = acc[px][0] * ha.x;
        s = fmaf(acc[px][1], ha.y, s); s = fmaf(acc[px][2], ha.z, s); s = fmaf(acc[px][3], ha.w, s);
        s = fmaf(acc[px][4], hb.x, s); s = fmaf(acc[px][5], hb.y, s); s = fmaf(acc[px][6], hb.z, s);
        s = fmaf(acc[px][7], hb.w, s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += p.head_b;
        const int ox = x0 + px;
        if (oy < p.hout && ox < p.wout) {
          const unsigned int idx = (unsigned)(oy * p.wout + ox);
          if (p.heat != nullptr && cg == 0) p.heat[(size_t)img * p.hout * p.wout + idx] = s;
          const unsigned long long key = ((unsigned long long)orderable(s) << 32) | (unsigned long long)(~idx);
          best = key > best ? key : best;
        }
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
      }
      if ((tid & 31) == 0 && best != 0ull) atomicMax(p.head_key + img, best);
      continue;
    }

    float4* outp = reinterpret_cast<float4*>(p.out);
    if (p.pool) {
      const int hp = p.hout >> 1, wp = p.wout >> 1;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int grp = (ch0 >> 2) + 8 * h;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float m[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = fmaxf(acc[2 * j][4 * h + e], acc[2 * j + 1][4 * h + e]);
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));   // row partner r^1 lives in lane^8
            m[e] = v;
          }
          const int py = oy >> 1, pxo = (x0 >> 1) + j;
          if ((r & 1) == 0 && py < hp && pxo < wp)
            outp[(((size_t)img * cgroups_out + grp) * hp + py) * wp + pxo] = make_float4(m[0], m[1], m[2], m[3]);
        }
      }
    } else if (p.ups) {
      const int hu = p.hout * 2, wu = p.wout * 2;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int grp = (ch0 >> 2) + 8 * h;
#pragma unroll
        for (int px = 0; px < 8; ++px) {
          const int ox = x0 + px;
          if (oy < p.hout && ox < p.wout) {
            const float4 v = make_float4(acc[px][4 * h], acc[px][4 * h + 1], acc[px][4 * h + 2], acc[px][4 * h + 3]);
            float4* base = outp + (((size_t)img * cgroups_out + grp) * hu + 2 * oy) * wu + 2 * ox;
            base[0] = v; base[1] = v; base[wu] = v; base[wu + 1] = v;
          }
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int grp = (ch0 >> 2) + 8 * h;
#pragma unroll
        for (int px = 0; px < 8; ++px) {
          const int ox = x0 + px;
          if (oy < p.hout && ox < p.wout)
            outp[(((size_t)img * cgroups_out + grp) * p.hout + oy) * p.wout + ox] =
                make_float4(acc[px][4 * h], acc[px][4 * h + 1], acc[px][4 * h + 2], acc[px][4 * h + 3]);
        }
      }
    }
  }
}

void launch_conv3x3_ffma(const ConvParams& p, const float* w_packed, cudaStream_t s) {
  const int tiles_x = ceil_div(p.wout, F_TW), tiles_y = ceil_div(p.hout, F_TH), cblocks = p.cout_total / F_CB;
  const long long total = (long long)p.n * tiles_x * tiles_y * cblocks;
  if (total <= 0) return;
  const long long cap = 148LL * 16;
  const int grid = (int)(total < cap ? total : cap);
  conv3x3_ffma_kernel<<<grid, F_THREADS, 0, s>>>(p, w_packed, tiles_x, tiles_y, cblocks, total);
}

// ---------------------------------------------------------------------------------------------------
// first layer: cin = 1, 64 output channels, one output pixel per thread
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
conv_first_kernel(FirstConvParams p, long long total_px) {
  __shared__ __align__(16) float w_s[9 * 64];
  __shared__ __align__(16) float bi_s[64], al_s[64], be_s[64];
  __shared__ float lut_s[256];
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) w_s[i] = p.w[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) { bi_s[i] = p.bias[i]; al_s[i] = p.alpha[i]; be_s[i] = p.beta[i]; }
  if (p.in_u8 != nullptr)
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut_s[i] = p.lut[i];
  __syncthreads();
  if (p.n_dev != nullptr) total_px = (long long)max(0, min(*p.n_dev - p.n_off, p.n)) * p.hout * p.wout;     // image count from the device

  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_px;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(idx % p.wout);
    long long t = idx / p.wout;
    const int oy = (int)(t % p.hout);
    const int img = (int)(t / p.hout);
    float v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gy = oy - p.pad + ky, gx = ox - p.pad + kx;
        float x = 0.f;   // zero padding is applied to the NORMALISED image (F.conv2d padding; net.py:23)
        if (gy >= 0 && gy < p.hin && gx >= 0 && gx < p.win) {
          const size_t o = ((size_t)img * p.hin + gy) * p.win + gx;
          x = (p.in_u8 != nullptr) ? lut_s[p.in_u8[o]] : p.in_f32[o];
        }
        v[ky * 3 + kx] = x;
      }
    float4* outp = reinterpret_cast<float4*>(p.out);
    uint4* outh = reinterpret_cast<uint4*>(p.out);
    const H2Layout lay = p.out_layout.plane ? p.out_layout : h2_standard(64, p.hout, p.wout);
    const size_t opix = (size_t)img * lay.img + (size_t)oy * lay.row + ox;
#pragma unroll 2
    for (int g8 = 0; g8 < 8; ++g8) {          // 8 output channels per iteration
      float y[8];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = g8 * 8 + half * 4;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float4 w = *reinterpret_cast<const float4*>(&w_s[tap * 64 + c]);
          a0 = fmaf(v[tap], w.x, a0); a1 = fmaf(v[tap], w.y, a1);
          a2 = fmaf(v[tap], w.z, a2); a3 = fmaf(v[tap], w.w, a3);
        }
        y[half * 4 + 0] = bn_relu(a0, bi_s[c], al_s[c], be_s[c]);
        y[half * 4 + 1] = bn_relu(a1, bi_s[c + 1], al_s[c + 1], be_s[c + 1]);
        y[half * 4 + 2] = bn_relu(a2, bi_s[c + 2], al_s[c + 2], be_s[c + 2]);
        y[half * 4 + 3] = bn_relu(a3, bi_s[c + 3], al_s[c + 3], be_s[c + 3]);
      }
      if (p.out_h2) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_h2(y[2 * e], y[2 * e + 1], h[e], l[e]);
        outh[opix + (size_t)g8 * lay.plane] = make_uint4(h[0], h[1], h[2], h[3]);
        outh[opix + (size_t)g8 * lay.plane + lay.lo] = make_uint4(l[0], l[1], l[2], l[3]);
      } else {
        outp[(((size_t)img * 16 + 2 * g8) * p.hout + oy) * p.wout + ox] = make_float4(y[0], y[1], y[2], y[3]);
        outp[(((size_t)img * 16 + 2 * g8 + 1) * p.hout + oy) * p.wout + ox] = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
}

// Row-streaming variant for wide maps (the detector's conv1a: 76.8k pixels and 19.7 MB of output per frame).  The kernel above
// re-reads its 576 weights from shared memory for every pixel (144 LDS.128) and is bound by the L1 / shared data pipe (ncu: 82 %)
// at 42 % of the HBM write rate.  Here a warp owns ONE group of 8 output channels and keeps its 72 weights + 24 BN constants in
// registers (plain register FFMA operands; constant-bank operands were tried and are slow: indexed LDC issues ~1 per 15 cycles per
// SM), lanes are 32 adjacent pixels, and the warp walks down FR_ROWS rows with a rolling 3x3 window fed from a small shared tile
// of the NORMALISED input (3 LDS per pixel instead of 144).  A block = 8 warps = all 64 channels of a 32-pixel-wide strip, so
// every store instruction writes 512 contiguous bytes of one H2 plane.  FMA order per output (taps 0..8 ascending, from 0) is the
// same as above: bit-identical results.
constexpr int FR_ROWS = 30;
template <bool kU8>
__global__ void __launch_bounds__(256, 2)
conv_first_rows_kernel(const FirstConvParams p, int strips_x, int row_blocks) {
  __shared__ float tile[(FR_ROWS + 2) * 34];
  const int lane = threadIdx.x & 31, g8 = threadIdx.x >> 5;
  int item = blockIdx.x;
  const int sx = item % strips_x; item /= strips_x;
  const int rb = item % row_blocks;
  const int img = item / row_blocks;
  const int x0 = sx * 32, y0 = rb * FR_ROWS;
  const int rows = min(FR_ROWS, p.hout - y0);
  // normalised input tile [rows + 2][34] with the layer's zero padding (applied to the NORMALISED image, net.py:23)
  {
    const size_t in_base = (size_t)img * p.hin * p.win;
    for (int i = threadIdx.x; i < (rows + 2) * 34; i += 256) {
      const int ty = i / 34, tx = i - ty * 34;
      const int gy = y0 - p.pad + ty, gx = x0 - p.pad + tx;
      float v = 0.f;
      if (gy >= 0 && gy < p.hin && gx >= 0 && gx < p.win) {
        const size_t o = in_base + (size_t)gy * p.win + gx;
        v = kU8 ? __fdiv_rn((float)p.in_u8[o] - 128.0f, 255.0f) : p.in_f32[o];      // (x - 128) / 255, model_utils.py:48-49
      }
      tile[i] = v;
    }
  }
  float w[9][8], bi[8], al[8], be[8];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.w + t * 64 + g8 * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.w + t * 64 + g8 * 8 + 4));
    w[t][0] = a.x; w[t][1] = a.y; w[t][2] = a.z; w[t][3] = a.w; w[t][4] = b.x; w[t][5] = b.y; w[t][6] = b.z; w[t][7] = b.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { bi[j] = __ldg(p.bias + g8 * 8 + j); al[j] = __ldg(p.alpha + g8 * 8 + j); be[j] = __ldg(p.beta + g8 * 8 + j); }
  __syncthreads();
  const int ox = x0 + lane;
  const H2Layout lay = p.out_layout.plane ? p.out_layout : h2_standard(64, p.hout, p.wout);
  uint4* outh = reinterpret_cast<uint4*>(p.out) + (size_t)img * lay.img + (size_t)g8 * lay.plane + (size_t)y0 * lay.row + ox;
  float v[9];
#pragma unroll
  for (int t = 0; t < 6; ++t) v[t] = tile[(t / 3) * 34 + lane + (t % 3)];
#pragma unroll 1
  for (int r = 0; r < rows; ++r) {
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) v[6 + kx] = tile[(r + 2) * 34 + lane + kx];
    float y[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = fmaf(v[t], w[t][j], y[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = bn_relu(y[j], bi[j], al[j], be[j]);
    if (ox < p.wout) {
      uint32_t h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) split_h2(y[2 * e], y[2 * e + 1], h[e], l[e]);
      uint4* o = outh + (size_t)r * lay.row;
      o[0] = make_uint4(h[0], h[1], h[2], h[3]);
      o[lay.lo] = make_uint4(l[0], l[1], l[2], l[3]);
    }
#pragma unroll
    for (int t = 0; t < 6; ++t) v[t] = v[t + 3];
  }
}

void launch_conv_first(const FirstConvParams& p, cudaStream_t s) {
  const long long total = (long long)p.n * p.hout * p.wout;
  if (total <= 0) return;
  static const bool rows_ok = [] { const char* v = getenv("DCU_FIRST_ROWS"); return !v || atoi(v) != 0; }();
  if (p.out_h2 && p.wout >= 64 && p.pad == 1 && rows_ok) {
    const int strips_x = ceil_div(p.wout, 32), row_blocks = ceil_div(p.hout, FR_ROWS);
    const long long items = (long long)p.n * strips_x * row_blocks;
    if (items < 0x7fffffffLL) {
      if (p.in_u8 != nullptr) conv_first_rows_kernel<true><<<(int)items, 256, 0, s>>>(p, strips_x, row_blocks);
      else conv_first_rows_kernel<false><<<(int)items, 256, 0, s>>>(p, strips_x, row_blocks);
      return;
    }
  }
  long long blocks = (total + 127) / 128;
  const long long cap = 148LL * 32;
  if (blocks > cap) blocks = cap;
  conv_first_kernel<<<(int)blocks, 128, 0, s>>>(p, total);
}

// ---------------------------------------------------------------------------------------------------
// detector heads: convPb (256 -> 65) on cPa and convDb (256 -> n_ids+1) on cDa, 1x1, bias, no activation
// ---------------------------------------------------------------------------------------------------
constexpr int H_CELLS = 16;
constexpr int H_MAXOUT = 12;   // outputs per thread: ceil((65 + n_ids1) / 8) <= 12  (n_ids1 <= 31)

__global__ void __launch_bounds__(128)
heads_1x1_kernel(HeadParams p, int cells_per_img, int blocks_per_img) {
  __shared__ float in_s[512][H_CELLS + 1];
  const int img = blockIdx.x / blocks_per_img;
  const int cell0 = (blockIdx.x % blocks_per_img) * H_CELLS;
  const int tid = threadIdx.x;
  if (p.in_h2) {
    // H2: [img][hi|lo][64 groups of 8][cells][8 fp16]; x = hi + lo (22 significant bits)
    const uint4* src = reinterpret_cast<const uint4*>(p.in) + (size_t)img * 2 * 64 * cells_per_img;
    for (int i = tid; i < 64 * H_CELLS; i += 128) {
      const int g = i / H_CELLS, c = i % H_CELLS;
      uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
      if (cell0 + c < cells_per_img) {
        h = src[(size_t)g * cells_per_img + cell0 + c];
        l = src[(size_t)(64 + g) * cells_per_img + cell0 + c];
      }
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
        in_s[8 * g + 2 * e][c] = a.x + b.x;
        in_s[8 * g + 2 * e + 1][c] = a.y + b.y;
      }
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>(p.in) + (size_t)img * 128 * cells_per_img;
    for (int i = tid; i < 128 * H_CELLS; i += 128) {
      const int g = i / H_CELLS, c = i % H_CELLS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cell0 + c < cells_per_img) v = src[(size_t)g * cells_per_img + cell0 + c];
      in_s[4 * g + 0][c] = v.x; in_s[4 * g + 1][c] = v.y; in_s[4 * g + 2][c] = v.z; in_s[4 * g + 3][c] = v.w;
    }
  }
  __syncthreads();
  const int cell = tid % H_CELLS, og = tid / H_CELLS;   // og 0..7
  const int n_out = 65 + p.n_ids1;
  float acc[H_MAXOUT];
  const float* wrow[H_MAXOUT];
  int base[H_MAXOUT];
#pragma unroll
  for (int j = 0; j < H_MAXOUT; ++j) {
    const int o = og + 8 * j;
    acc[j] = 0.f;
    if (o < 65) { wrow[j] = p.w_loc + (size_t)o * 256; base[j] = 0; }
    else if (o < n_out) { wrow[j] = p.w_ids + (size_t)(o - 65) * 256; base[j] = 256; }
    else { wrow[j] = p.w_loc; base[j] = 0; }
  }
  for (int k = 0; k < 256; k += 4) {
    const float xl0 = in_s[k][cell], xl1 = in_s[k + 1][cell], xl2 = in_s[k + 2][cell], xl3 = in_s[k + 3][cell];
    const float xi0 = in_s[256 + k][cell], xi1 = in_s[257 + k][cell], xi2 = in_s[258 + k][cell], xi3 = in_s[259 + k][cell];
#pragma unroll
    for (int j = 0; j < H_MAXOUT; ++j) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(wrow[j] + k));
      const bool is_ids = base[j] != 0;
      acc[j] = fmaf(is_ids ? xi0 : xl0, w.x, acc[j]);
      acc[j] = fmaf(is_ids ? xi1 : xl1, w.y, acc[j]);
      acc[j] = fmaf(is_ids ? xi2 : xl2, w.z, acc[j]);
      acc[j] = fmaf(is_ids ? xi3 : xl3, w.w, acc[j]);
    }
  }
  if (cell0 + cell < cells_per_img) {
#pragma unroll
    for (int j = 0; j < H_MAXOUT; ++j) {
      const int o = og + 8 * j;
      if (o < 65)
        p.loc[((size_t)img * 65 + o) * cells_per_img + cell0 + cell] = acc[j] + p.b_loc[o];
      else if (o < n_out)
        p.ids[((size_t)img * p.n_ids1 + (o - 65)) * cells_per_img + cell0 + cell] = acc[j] + p.b_ids[o - 65];
    }
  }
}

void launch_heads_1x1(const HeadParams& p, cudaStream_t s) {
  const int cells = p.h * p.w;
  const int bpi = ceil_div(cells, H_CELLS);
  if (p.n <= 0) return;
  heads_1x1_kernel<<<p.n * bpi, 128, 0, s>>>(p, cells, bpi);
}

// ---------------------------------------------------------------------------------------------------
// RefineNet tail and layout converters
// ---------------------------------------------------------------------------------------------------
__global__ void refine_finalize_kernel(const unsigned long long* keys, const int32_t* xy, int xy_stride, int p,
                                       int32_t* corners, float* refined, const int* n_dev, int n_off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev != nullptr) p = max(0, min(*n_dev - n_off, p));
  if (i >= p) return;
  const unsigned int idx = ~(unsigned int)(keys[i] & 0xffffffffull);
  const int col = (int)(idx & 63u), row = (int)(idx >> 6);          // speedy_bargmax2d: idx % 64, idx // 64
  if (corners != nullptr) { corners[2 * i] = col; corners[2 * i + 1] = row; }
  // corners_og = (corners - 32) / 8 + keypoints   (refinenet.py:114; exact in fp32: multiples of 1/8)
  refined[2 * i] = (float)(col - 32) / 8.0f + (float)xy[(size_t)i * xy_stride];
  refined[2 * i + 1] = (float)(row - 32) / 8.0f + (float)xy[(size_t)i * xy_stride + 1];
}

void launch_refine_finalize(const unsigned long long* keys, const int32_t* xy, int xy_stride, int p,
                            int32_t* corners, float* refined, cudaStream_t s, const int* n_dev, int n_off) {
  if (p <= 0) return;
  refine_finalize_kernel<<<ceil_div(p, 128), 128, 0, s>>>(keys, xy, xy_stride, p, corners, refined, n_dev, n_off);
}

__global__ void nchw_to_c4_kernel(const float* in, float* out, int n, int c, int h, int w) {
  const long long total = (long long)n * c * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w); long long t = i / w;
    const int y = (int)(t % h); t /= h;
    const int ch = (int)(t % c); const int img = (int)(t / c);
    out[((((size_t)img * (c >> 2) + (ch >> 2)) * h + y) * w + x) * 4 + (ch & 3)] = in[i];
  }
}
__global__ void c4_to_nchw_kernel(const float* in, float* out, int n, int c, int h, int w) {
  const long long total = (long long)n * c * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w); long long t = i / w;
    const int y = (int)(t % h); t /= h;
    const int ch = (int)(t % c); const int img = (int)(t / c);
    out[i] = in[((((size_t)img * (c >> 2) + (ch >> 2)) * h + y) * w + x) * 4 + (ch & 3)];
  }
}
__global__ void nchw_to_h2_kernel(const float* in, __half* out, int n, int c, int h, int w, int sub, H2Layout lay) {
  const long long total = (long long)n * c * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w); long long t = i / w;
    const int y = (int)(t % h); t /= h;
    const int ch = (int)(t % c); const int img = (int)(t / c);
    const float v = fminf(in[(((size_t)img * c + ch) * (h * sub) + (size_t)y * sub) * (w * sub) + (size_t)x * sub], 65504.f);
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const size_t o = (size_t)img * lay.img + (size_t)(ch >> 3) * lay.plane + (size_t)y * lay.row + x;
    out[o * 8 + (ch & 7)] = hi;
    out[(o + (size_t)lay.lo) * 8 + (ch & 7)] = lo;
  }
}
__global__ void h2_to_nchw_kernel(const __half* in, float* out, int n, int c, int h, int w, int rep, H2Layout lay) {
  const int ho = h * rep, wo = w * rep;
  const long long total = (long long)n * c * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo) / rep; long long t = i / wo;
    const int y = (int)(t % ho) / rep; t /= ho;
    const int ch = (int)(t % c); const int img = (int)(t / c);
    const size_t o = (size_t)img * lay.img + (size_t)(ch >> 3) * lay.plane + (size_t)y * lay.row + x;
    out[i] = __half2float(in[o * 8 + (ch & 7)]) + __half2float(in[(o + (size_t)lay.lo) * 8 + (ch & 7)]);
  }
}
void launch_nchw_to_h2(const float* in, void* out, int n, int c, int h, int w, cudaStream_t s, int sub, const H2Layout* lay) {
  const long long total = (long long)n * c * h * w;
  if (total <= 0) return;
  long long b = (total + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  nchw_to_h2_kernel<<<(int)b, 256, 0, s>>>(in, reinterpret_cast<__half*>(out), n, c, h, w, sub, lay ? *lay : h2_standard(c, h, w));
}
void launch_h2_to_nchw(const void* in, float* out, int n, int c, int h, int w, cudaStream_t s, int rep, const H2Layout* lay) {
  const long long total = (long long)n * c * h * w * rep * rep;
  if (total <= 0) return;
  long long b = (total + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  h2_to_nchw_kernel<<<(int)b, 256, 0, s>>>(reinterpret_cast<const __half*>(in), out, n, c, h, w, rep, lay ? *lay : h2_standard(c, h, w));
}

void launch_nchw_to_c4(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s) {
  const long long total = (long long)n * c * h * w;
  if (total <= 0) return;
  long long b = (total + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  nchw_to_c4_kernel<<<(int)b, 256, 0, s>>>(in, out, n, c, h, w);
}
void launch_c4_to_nchw(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s) {
  const long long total = (long long)n * c * h * w;
  if (total <= 0) return;
  long long b = (total + 255) / 256; if (b > 148 * 16) b = 148 * 16;
  c4_to_nchw_kernel<<<(int)b, 256, 0, s>>>(in, out, n, c, h, w);
}

}  // namespace dcu
