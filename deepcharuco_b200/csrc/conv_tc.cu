// 3x3 convolution as an implicit GEMM on the 5th-generation tensor cores (tcgen05.mma, sm_100a).
//
//   D[pixel, cout] = sum over (tap, cin)  A[pixel + tap, cin] * W[cout, cin, tap]
//
// Reference semantics: conv -> BatchNorm2d(eval) -> ReLU (-> MaxPool2d(2) | UpsamplingNearest2d(2) | convPb +
// arg-max), /root/reference/src/models/net.py:60-77 and refinenet.py:56-81.
//
// Numerics.  The reference is fp32 and its outputs are quantised by arg-maxes whose top-1/top-2 margins go
// down to ~1e-6, so a single low-precision pass is not acceptable.  Each operand is split into
// two fp16 numbers, x = x_hi + x_lo with x_hi = fp16(x), x_lo = fp16(x - x_hi): 2 x 11 significant bits, the same
// 22 bits a TF32 hi/lo split keeps, but at half the bytes per element and twice the tensor-pipe rate.  Weights
// are pre-multiplied by a per-layer power of two (exact) so that w_lo stays in fp16's normal range; the epilogue
// multiplies the accumulator back by 2^-s (exact).  Activations are post-ReLU values of magnitude <= ~1e2; an
// activation below 2^-14*2^11 only loses bits whose weight is below one fp32 ulp of the accumulator (checked
// against the fp32 oracle: same logit / heat-map error as the TF32 split, tools/emulate_split.py).
// The three products a_hi*w_hi + a_hi*w_lo + a_lo*w_hi (a_lo*w_lo <= 2^-22 relative is dropped) are issued as
// TWO kind::f16 MMAs per (tap, 16-channel chunk):  a_hi x [w_hi | w_lo]  (N' = 2*NT, hi and lo weight rows are
// adjacent in the block) and  a_lo x w_hi  (N = NT, into the right half).  Each m-tile therefore owns 2*NT TMEM
// columns: the left half accumulates the main term, the right half the two small terms; the epilogue adds them
// in fp32.  This keeps the small terms out of the main accumulator (the hardware accumulation truncates once per
// MMA) and reads a_hi from shared memory once for both weight halves.
//
// Data movement.  Activations live in HBM already split ("H2": __half [n][hi|lo][C/8][H][W][8], common.cuh); the
// producing kernel's epilogue does the split.  One 5-D TMA box {8*haloW, haloH, 2 k-groups, hi|lo} brings a
// (16*TR+2) x (8*TC+2) pixel halo of 16 input channels into shared memory as four planes of 16-byte pixels.  In
// that layout any run of 8 horizontally adjacent pixels IS a no-swizzle K-major core matrix (8 rows x 16 B),
// consecutive image rows are SBO = haloW*16 B apart and the next 8 channels are LBO = plane bytes apart -- so all
// nine taps of the convolution are nine descriptor start addresses into the same halo tile: no im2col copy, no
// conversion pass, each input element is fetched from L2 once per tile (+halo); zero padding comes from TMA
// out-of-bounds fill.  Weight blocks (pre-packed, three taps per stage) arrive by 1-D bulk copy.
// The kernel is bound by the shared-memory data pipe (operand fetches of the tensor core + fills), not by HBM or
// the tensor pipe (profiles/r1_ncu_conv3x3_tc.txt), which is why every byte of fill traffic was removed that could be.
//
// Roles (384 threads, one persistent CTA per SM).  TMEM: 512 columns = NBUF sets x MT=2 m-tiles x (main | small) x NT
// columns; NT=64 double-buffers the set (epilogue of tile i overlaps the MMAs of tile i+1), NT=128 has one set that is
// handed back to the MMA issuer per m-tile as the epilogue drains it:
//   warp 0      TMA producer for activation halo chunks                   (a_empty -> a_full)
//   warp 1      tcgen05.mma issuer: warp-uniform loop, one elected lane   (a_full, b_full -> commits)
//   warp 2      TMEM alloc/dealloc + bulk-copy producer for weight stages (b_empty -> b_full)
//   warps 4-7, 8-11  epilogue, two groups (even / odd m-tiles): tcgen05.ld -> main+small -> *2^-s -> bias/BN/ReLU ->
//               pool | upsample | head -> hi/lo split -> global           (acc_full -> acc_empty[mt])
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace dcu {

namespace {

constexpr int TC_THREADS = 384;

template <int NT>
struct TcCfg {
  static constexpr int MT = 2;                                  // 128-pixel m-tiles per CTA tile (16x16 or 32x8 pixels)
  static constexpr int NBUF = 512 / (MT * 2 * NT);              // accumulator sets in TMEM (NBUF * MT * 2*NT = 512 columns)
  static constexpr int A_STAGES = 4;
  static constexpr int TPB = 3;                                 // taps per weight stage (one bulk copy brings TPB blocks)
  static constexpr int B_STAGES = (NT == 64) ? 6 : 4;
  static constexpr int MAX_HALO_PX = 34 * 10;                   // 2x1 arrangement; 1x2 is 18 x 18 = 324
  static constexpr int A_STAGE_BYTES = 4 * MAX_HALO_PX * 16;    // hi (2 planes of 8 channels) | lo (2 planes)
  static constexpr int B_BLOCK_BYTES = 2 * (2 * NT) * 16;       // 2 k-groups x (NT hi rows + NT lo rows) x 8 fp16
  static constexpr int B_STAGE_BYTES = TPB * B_BLOCK_BYTES;
  static constexpr int PARAM_BYTES = 3 * 512 * 4;               // bias / alpha / beta for up to 512 channels
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + PARAM_BYTES + BAR_BYTES + 1024;
};

struct TcGeo {
  int tr, tc;            // m-tile arrangement: TR x TC tiles of 16 rows x 8 cols
  int halo_w, halo_h;    // 8*tc+2, 16*tr+2
  int tiles_x, tiles_y;
  int slices;            // cout_total / NT
  int w_copies;          // replicas of the packed weights in HBM (L2 hot-spot avoidance)
  long long w_copy_bytes;
  long long total_tiles;
};

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fail the launch (trap), never hang the GPU box.
// kBackoffNs > 0: sleep between polls.  Every poll is a shared-memory transaction, and with 14 of 16 warps waiting most
// of the time un-throttled polling took ~15 % of the shared-memory data pipe this kernel is bound by (ncu:
// l1tex__data_pipe_lsu_wavefronts_mem_shared, profiles/r1_conv1b_ncu.txt).  Only the MMA issuer polls without back-off.
template <int kBackoffNs = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (kBackoffNs > 0) __nanosleep(kBackoffNs);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// wait + accumulate the cycles spent waiting into *acc (profiling builds of the role loops)
template <int kBackoffNs = 0>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool timed, long long& acc) {
  if (!timed) { mbar_wait<kBackoffNs>(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait<kBackoffNs>(bar, parity);
  acc += clock64() - t0;
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 64-bit descriptors passed as two 32-bit halves: only the low word (start address field) changes between MMAs
__device__ __forceinline__ void umma_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "elect.sync _|P1, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// main and small 16-column slices in flight together, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, float* a, uint32_t tb, float* b) {
  tmem_ld16_nowait(ta, a);
  tmem_ld16_nowait(tb, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// x -> (fp16(x), fp16(x - fp16(x))) for two values; the inputs are post-ReLU (>= 0) and clamped to fp16's finite range
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  x0 = fminf(x0, 65504.f); x1 = fminf(x1, 65504.f);
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
// 16 fp32 channel values of one pixel -> two hi and two lo 16-byte H2 pixels (k-groups kg0, kg0+1)
__device__ __forceinline__ void store_h2_16(uint4* hi_plane0, size_t lo_offset, size_t kg_stride, const float* v) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_h2(v[8 * k + 2 * e], v[8 * k + 2 * e + 1], h[e], l[e]);
    hi_plane0[(size_t)k * kg_stride] = make_uint4(h[0], h[1], h[2], h[3]);
    hi_plane0[(size_t)k * kg_stride + lo_offset] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__device__ __forceinline__ unsigned int orderable(float v) {
  unsigned int b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

struct TileCoord { int img, slice, y0, x0; };
__device__ __forceinline__ TileCoord decode_tile(long long t, const TcGeo& g) {
  TileCoord c;
  c.slice = (int)(t % g.slices); t /= g.slices;
  const int tx = (int)(t % g.tiles_x); t /= g.tiles_x;
  const int ty = (int)(t % g.tiles_y);
  c.img = (int)(t / g.tiles_y);
  c.y0 = ty * 16 * g.tr;
  c.x0 = tx * 8 * g.tc;
  return c;
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <int NT, int KS>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvParams p, const float* __restrict__ w_blocks,
                  const TcGeo g) {
  using Cfg = TcCfg<NT>;
  constexpr int MT = Cfg::MT;
  constexpr int NBUF = Cfg::NBUF;
  constexpr int B_STAGES = Cfg::B_STAGES;
  constexpr int ROWS = KS, TPR = KS, TAPS = KS * KS;       // kernel rows, taps per row (= per weight stage), taps
  // NB: no integer round-trip on this pointer -- it would demote every shared-memory access below to a generic LD/ST
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* a_smem = smem_raw;
  uint8_t* b_smem = a_smem + Cfg::A_STAGES * Cfg::A_STAGE_BYTES;
  float* prm = reinterpret_cast<float*>(b_smem + B_STAGES * Cfg::B_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(prm) + Cfg::PARAM_BYTES);
  uint64_t* a_full = bars;                         // [A_STAGES] TMA landed (tx bytes)
  uint64_t* a_empty = a_full + Cfg::A_STAGES;      // [A_STAGES] MMAs reading the stage retired
  uint64_t* b_full = a_empty + Cfg::A_STAGES;      // [B_STAGES]
  uint64_t* b_empty = b_full + B_STAGES;           // [B_STAGES]
  uint64_t* acc_full = b_empty + B_STAGES;         // [NBUF]     all MMAs of the tile retired
  uint64_t* acc_empty = acc_full + NBUF;           // [NBUF][MT] m-tile drained by the epilogue (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NBUF * MT);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");         // programmatic dependent launch, see conv_tc2.cu
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (keeps role code on the uniform datapath)
  const int lane = threadIdx.x & 31;
  const int chunks = p.cin >> 4;
  const int halo_px = g.halo_w * g.halo_h;

  for (int i = threadIdx.x; i < p.cout_total; i += TC_THREADS) {
    prm[i] = p.bias[i];
    prm[512 + i] = p.alpha[i];
    prm[1024 + i] = p.beta[i];
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < NBUF; ++i) mbar_init(&acc_full[i], 1);
    for (int i = 0; i < NBUF * MT; ++i) mbar_init(&acc_empty[i], 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ================= activation producer (TMA) =================
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");        // activations = the previous kernel's output
    int st = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(t, g);
      for (int q = 0; q < chunks; ++q) {
        mbar_wait<200>(&a_empty[st], ph ^ 1u);
        mbar_expect_tx(&a_full[st], (uint32_t)halo_px * 64u);
        tma_load_5d(smem_u32(a_smem + (size_t)st * Cfg::A_STAGE_BYTES), &tmap, &a_full[st], (c.x0 - p.pad) * 8, c.y0 - p.pad,
                    (p.cin_offset >> 3) + q * 2, 0, c.img);
        if (++st == Cfg::A_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2 && lane == 0) {
    // ================= weight producer (bulk copy) =================
    int st = 0; uint32_t ph = 0;
    for (long long t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(t, g);
      // every CTA streams the same weight blocks at about the same time: read them from one of g.w_copies replicas so the
      // requests spread over more L2 slices (and both dies) instead of hammering the few slices that home one copy
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(w_blocks) + (size_t)(blockIdx.x % g.w_copies) * g.w_copy_bytes +
                            (size_t)c.slice * chunks * TAPS * Cfg::B_BLOCK_BYTES;
      for (int blk = 0; blk < chunks * TAPS; blk += TPR) {
        mbar_wait<200>(&b_empty[st], ph ^ 1u);
        mbar_expect_tx(&b_full[st], (uint32_t)(TPR * Cfg::B_BLOCK_BYTES));
        bulk_load(smem_u32(b_smem + (size_t)st * Cfg::B_STAGE_BYTES), wsrc + (size_t)blk * Cfg::B_BLOCK_BYTES,
                  (uint32_t)(TPR * Cfg::B_BLOCK_BYTES), &b_full[st]);
        if (++st == B_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer: whole warp runs the (uniform) loop, one elected lane issues =================
    constexpr uint32_t IDESC_BASE = (1u << 4) | (0u << 7) | (0u << 10) | (8u << 24);     // f32 accum, f16 x f16, K-major, M = 128, K = 16
    constexpr uint32_t IDESC_2N = IDESC_BASE | ((uint32_t)((2 * NT) >> 3) << 17);          // a_hi x [w_hi | w_lo]
    constexpr uint32_t IDESC_1N = IDESC_BASE | ((uint32_t)(NT >> 3) << 17);                // a_lo x w_hi
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    // descriptor words (cute::UMMA::SmemDescriptor): lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version(1)<<14
    const uint32_t a_desc_hi = (uint32_t)g.halo_w | (1u << 14);                 // SBO = halo_w * 16 B
    const uint32_t a_desc_lo0 = ((uint32_t)halo_px << 16);                      // LBO = plane = halo_px * 16 B
    constexpr uint32_t b_desc_hi = 8u | (1u << 14);                             // SBO = 128 B
    constexpr uint32_t b_desc_lo0 = ((uint32_t)(2 * NT) << 16);                 // LBO = 2*NT * 16 B (hi rows | lo rows)
    const uint32_t a_base0 = smem_u32(a_smem) >> 4, b_base0 = smem_u32(b_smem) >> 4;   // 16-byte units from here on
    uint32_t mt_off[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int tri = mt / g.tc, tci = mt - tri * g.tc;
      mt_off[mt] = (uint32_t)(tri * 16 * g.halo_w + tci * 8);
    }
    int sa = 0, sb = 0, buf = 0; uint32_t pha = 0, phb = 0, phc = 0;
    const bool timed = p.stats != nullptr;
    long long w_a = 0, w_b = 0, w_c = 0;
    const long long t_start = clock64();
    for (long long t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      for (int q = 0; q < chunks; ++q) {
        mbar_wait_t(&a_full[sa], pha, timed, w_a);
        tc_fence_after();
        const uint32_t a_hi = a_desc_lo0 + a_base0 + (uint32_t)sa * (Cfg::A_STAGE_BYTES >> 4);
        if (q == 0) {                            // first touch of this accumulator set in this tile: wait for the epilogue
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) mbar_wait_t(&acc_empty[buf * MT + mt], phc ^ 1u, timed, w_c);
          tc_fence_after();
        }
#pragma unroll 1
        for (int ky = 0; ky < ROWS; ++ky) {      // one weight stage (TPR taps) per kernel row
          mbar_wait_t(&b_full[sb], phb, timed, w_b);
          tc_fence_after();
          const uint32_t a_row = a_hi + (uint32_t)(ky * g.halo_w);
          const uint32_t b_row = b_desc_lo0 + b_base0 + (uint32_t)sb * (Cfg::B_STAGE_BYTES >> 4);
          // ONE elected region per kernel row (3 taps x MT m-tiles x 2 MMAs): the elect / reconverge / syncwarp sequence
          // costs ~100 cycles, which dominated when it wrapped every pair of MMAs (profiles/r1_tc_variants.txt)
          if (elect_one()) {
#pragma unroll
            for (int kx = 0; kx < TPR; ++kx) {
              const uint32_t b_blk = b_row + (uint32_t)kx * (Cfg::B_BLOCK_BYTES >> 4);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t d = tmem_u + (uint32_t)((buf * MT + mt) * 2 * NT);
                const uint32_t da_hi = a_row + (uint32_t)kx + mt_off[mt];
                const uint32_t da_lo = da_hi + 2u * (uint32_t)halo_px;            // lo planes follow the two hi planes
                umma_f16_w(d, da_hi, a_desc_hi, b_blk, b_desc_hi, IDESC_2N, (q | ky | kx) ? 1u : 0u);   // main | a_hi*w_lo
                umma_f16_w(d + NT, da_lo, a_desc_hi, b_blk, b_desc_hi, IDESC_1N, 1u);                     // + a_lo*w_hi
              }
            }
            umma_commit(&b_empty[sb]);
            if (ky == ROWS - 1) umma_commit(&a_empty[sa]);
            if (ky == ROWS - 1 && q == chunks - 1) umma_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (++sb == B_STAGES) { sb = 0; phb ^= 1u; }
        }
        if (++sa == Cfg::A_STAGES) { sa = 0; pha ^= 1u; }
      }
      if (++buf == NBUF) { buf = 0; phc ^= 1u; }
    }
    if (timed && lane == 0) {
      atomicAdd(p.stats + 0, (unsigned long long)(clock64() - t_start));   // MMA warp: total loop cycles
      atomicAdd(p.stats + 1, (unsigned long long)w_a);                     //   waiting for TMA halo chunks
      atomicAdd(p.stats + 2, (unsigned long long)w_b);                     //   waiting for weight blocks
      atomicAdd(p.stats + 3, (unsigned long long)w_c);                     //   waiting for the epilogue to drain TMEM
    }
  } else if (warp >= 4) {
    // ================= epilogue (2 groups x 128 threads; warp w reads TMEM lanes 32*(w%4) ..) =================
    constexpr int CW = 16;                 // accumulator columns (channels) per step
    const int grp = (warp >= 8) ? 1 : 0;   // group 0 drains even m-tiles, group 1 odd ones
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;          // accumulator row = pixel index inside the 16x8 m-tile
    const int prow = m >> 3, pcol = m & 7;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int c8_out = p.cout_total >> 3;   // output k-groups (8 channels each) per hi or lo half
    uint32_t phc = 0;
    int buf = 0;
    const bool timed = p.stats != nullptr;
    long long w_e = 0;
    const long long t_start = clock64();
    for (long long t = blockIdx.x; t < g.total_tiles; t += gridDim.x) {
      const TileCoord c = decode_tile(t, g);
      mbar_wait_t<200>(&acc_full[buf], phc, timed, w_e);
      tc_fence_after();
      const int ch_base = c.slice * NT;
#pragma unroll 1
      for (int mt = grp; mt < MT; mt += 2) {
        const int tri = mt / g.tc, tci = mt - tri * g.tc;
        const int oy = c.y0 + tri * 16 + prow, ox = c.x0 + tci * 8 + pcol;
        const bool inb = (oy < p.hout) && (ox < p.wout);
        float head_sum = 0.f;
        float arg_best = 0.f; int arg_idx = -1;          // 1x1 logits mode: running arg-max over the channels (first max wins)
#pragma unroll 1
        for (int cc = 0; cc < NT / CW; ++cc) {
          float v[CW];
          {
            float sm[CW];
            const uint32_t col = (uint32_t)((buf * MT + mt) * 2 * NT + cc * CW);
            tmem_ld16x2(tmem_base + lane_addr + col, v, tmem_base + lane_addr + col + NT, sm);
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = (v[j] + sm[j]) * p.wscale_inv;   // main + (a_hi*w_lo + a_lo*w_hi), undo 2^s
          }
          const int ch0 = ch_base + cc * CW;
          if (p.logits != nullptr || p.arg_out != nullptr) {
            // detector heads: logits = acc + bias, no activation, fp32 NCHW (the layout pred_argmax reads, model_utils.py:72);
            // arg_out: torch.argmax over the channels (model_utils.py:73-74) on exactly those values, strict '>' = first maximum
            if (inb) {
              const size_t plane_o = (size_t)p.hout * p.wout;
              float* o = p.logits ? p.logits + (size_t)c.img * p.n_valid * plane_o + (size_t)oy * p.wout + ox : nullptr;
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (ch0 + j < p.n_valid) {
                  const float lv = v[j] + prm[ch0 + j];
                  if (o) o[(size_t)(ch0 + j) * plane_o] = lv;
                  if (arg_idx < 0 || lv > arg_best) { arg_best = lv; arg_idx = ch0 + j; }
                }
              if (p.arg_out != nullptr && cc == NT / CW - 1)
                p.arg_out[(size_t)c.img * plane_o + (size_t)oy * p.wout + ox] = (uint8_t)arg_idx;
            }
            continue;
          }
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 bi = *reinterpret_cast<const float4*>(&prm[ch0 + j]);
            const float4 al = *reinterpret_cast<const float4*>(&prm[512 + ch0 + j]);
            const float4 be = *reinterpret_cast<const float4*>(&prm[1024 + ch0 + j]);
            v[j + 0] = fmaxf(fmaf(v[j + 0] + bi.x, al.x, be.x), 0.0f);
            v[j + 1] = fmaxf(fmaf(v[j + 1] + bi.y, al.y, be.y), 0.0f);
            v[j + 2] = fmaxf(fmaf(v[j + 2] + bi.z, al.z, be.z), 0.0f);
            v[j + 3] = fmaxf(fmaf(v[j + 3] + bi.w, al.w, be.w), 0.0f);
          }
          if (p.head_w != nullptr) {
#pragma unroll
            for (int j = 0; j < CW; ++j) head_sum = fmaf(v[j], __ldg(p.head_w + cc * CW + j), head_sum);
          } else if (p.pool) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              float x = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));   // column partner
              v[j] = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));            // row partner
            }
            const int hp = p.hout >> 1, wp = p.wout >> 1;
            if (((lane & 9) == 0) && (oy >> 1) < hp && (ox >> 1) < wp) {
              const size_t plane_o = (size_t)hp * wp;
              uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)c.img * 2 * c8_out + (ch0 >> 3)) * plane_o + (size_t)(oy >> 1) * wp + (ox >> 1);
              store_h2_16(o, (size_t)c8_out * plane_o, plane_o, v);
            }
          } else if (p.ups) {
            if (inb) {
              const int hu = p.hout * 2, wu = p.wout * 2;
              const size_t plane_o = (size_t)hu * wu;
              uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)c.img * 2 * c8_out + (ch0 >> 3)) * plane_o + (size_t)(2 * oy) * wu + 2 * ox;
              const size_t lo_off = (size_t)c8_out * plane_o;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) split_h2(v[8 * k + 2 * e], v[8 * k + 2 * e + 1], h[e], l[e]);
                const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
                uint4* ok = o + (size_t)k * plane_o;
                ok[0] = hv; ok[1] = hv; ok[wu] = hv; ok[wu + 1] = hv;
                ok[lo_off] = lv; ok[lo_off + 1] = lv; ok[lo_off + wu] = lv; ok[lo_off + wu + 1] = lv;
              }
            }
          } else {
            if (inb) {
              const size_t plane_o = (size_t)p.hout * p.wout;
              uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)c.img * 2 * c8_out + (ch0 >> 3)) * plane_o + (size_t)oy * p.wout + ox;
              store_h2_16(o, (size_t)c8_out * plane_o, plane_o, v);
            }
          }
        }
        if (p.head_w != nullptr) {
          // convPb (64 -> 1) + bias, then the flat 64x64 arg-max with lowest-index tie-break (model_utils.py:39-43)
          const float s = head_sum + p.head_b;
          unsigned long long key = 0ull;
          if (inb) {
            const unsigned int idx = (unsigned)(oy * p.wout + ox);
            if (p.heat != nullptr) p.heat[(size_t)c.img * p.hout * p.wout + idx] = s;
            key = ((unsigned long long)orderable(s) << 32) | (unsigned long long)(~idx);
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
          }
          if (lane == 0 && key != 0ull) atomicMax(p.head_key + c.img, key);
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[buf * MT + mt]);   // these columns may now be overwritten by a later tile's MMAs
      }
      if (++buf == NBUF) { buf = 0; phc ^= 1u; }
    }
    if (timed && (threadIdx.x == 128 || threadIdx.x == 256)) {
      atomicAdd(p.stats + 6, (unsigned long long)(clock64() - t_start) / 2);   // epilogue: total loop cycles (avg of 2 groups)
      atomicAdd(p.stats + 7, (unsigned long long)w_e / 2);                     //   waiting for MMAs
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// m-tile arrangement (TR x TC tiles of 16 rows x 8 cols) per CTA tile; also sizes the TMA box (engine.cu)
void tc_tile_arrangement(int nt, int hout, int wout, int* tr, int* tc) {
  (void)nt;                      // MT = 2 for both NT: 16 rows x 16 cols (1x2) or 32 rows x 8 cols (2x1)
  if (wout % 16 == 0 || wout > 40) { *tr = 1; *tc = 2; }
  else if (hout > 16) { *tr = 2; *tc = 1; }
  else { *tr = 1; *tc = 2; }
}

int tc_supported_shape(int cin, int cout) {
  if (cin % 16 != 0 || cin > 256) return 0;
  if (cout == 64) return 64;
  // Slice width (DCU_NT64: 0 = 128-channel slices on two m-tiles per CTA wherever possible, 1 = as 0 but the 64 -> 128 layers in 64s,
  // 2 = every layer in 64-channel slices, 3 = default: 128-channel slices on ONE m-tile per CTA for the plain 3x3 launches, 64-channel
  // slices for the FLAT / upsample-fused / small ones -- the engine picks per launch, see run_3x3).  A 128-channel slice on two
  // m-tiles owns all 512 TMEM columns (2 m-tiles x [main | correction] x 128): its accumulators are single-buffered and the MMA warp
  // waits for the epilogue's tcgen05.ld (64 B / cycle / SM: ~4 k cycles per tile, 14 - 17 % of the tile).  64-channel slices are
  // double-buffered (+ 1.1 % frames/s over mode 0), but their N = 64 correction MMA is operand-fetch bound (40 cycles of shared-memory
  // reads for 32 of math).  One m-tile x 128 channels is double-buffered AND has no fetch-bound MMA: - 5 ... - 9 % on the detector's
  // 128- and 512-channel layers, same box.
  static const int nt64 = [] { const char* v = getenv("DCU_NT64"); return v ? atoi(v) : 3; }();
  if (tc2_segmented(cin) && cout % 64 == 0 && cout <= 512) return 64;
  if ((nt64 == 1 || nt64 == 2) && cin == 64 && cout == 128) return 64;
  if (nt64 == 2 && cout % 64 == 0 && cout <= 512) return 64;
  if (cout % 128 == 0 && cout <= 512) return 128;
  return 0;
}

template <int NT, int KS>
static cudaError_t launch_nt(const ConvParams& p, const float* w_blocks, int n_slices, int w_copies, const CUtensorMap* tm,
                             int sm_count, cudaStream_t s) {
  using Cfg = TcCfg<NT>;
  static bool attr_done[64] = {};      // function attributes are per device: engines on several GPUs of one process each opt in
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<NT, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  TcGeo g;
  tc_tile_arrangement(NT, p.hout, p.wout, &g.tr, &g.tc);
  if (g.tr * g.tc != Cfg::MT) return cudaErrorInvalidValue;
  g.halo_w = 8 * g.tc + 2; g.halo_h = 16 * g.tr + 2;
  if (g.halo_w * g.halo_h > Cfg::MAX_HALO_PX) return cudaErrorInvalidValue;
  g.tiles_x = ceil_div(p.wout, 8 * g.tc); g.tiles_y = ceil_div(p.hout, 16 * g.tr);
  g.slices = n_slices;
  g.w_copies = w_copies < 1 ? 1 : w_copies;
  g.w_copy_bytes = (long long)n_slices * (p.cin / 16) * (KS * KS) * Cfg::B_BLOCK_BYTES;
  g.total_tiles = (long long)p.n * g.slices * g.tiles_x * g.tiles_y;
  if (g.total_tiles <= 0) return cudaSuccess;
  const int grid = (int)(g.total_tiles < sm_count ? g.total_tiles : sm_count);
  static const bool pdl = [] { const char* v = getenv("DCU_PDL"); return !v || atoi(v) != 0; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid, 1, 1); cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NT, KS>, *tm, p, w_blocks, g);
}

cudaError_t launch_conv3x3_tc(const ConvParams& p, const float* w_blocks, int n_slices, int w_copies, const void* tmap_in,
                              int sm_count, cudaStream_t s) {
  const CUtensorMap* tm = reinterpret_cast<const CUtensorMap*>(tmap_in);
  const int nt = p.cout_total / n_slices;        // output channels per CTA pass (weight block rows)
  if (p.ksize == 1) {
    if (nt == 64) return launch_nt<64, 1>(p, w_blocks, n_slices, w_copies, tm, sm_count, s);
    if (nt == 128) return launch_nt<128, 1>(p, w_blocks, n_slices, w_copies, tm, sm_count, s);
    return cudaErrorInvalidValue;
  }
  if (nt == 64) return launch_nt<64, 3>(p, w_blocks, n_slices, w_copies, tm, sm_count, s);
  if (nt == 128) return launch_nt<128, 3>(p, w_blocks, n_slices, w_copies, tm, sm_count, s);
  return cudaErrorInvalidValue;
}

}  // namespace dcu
