// CTA-pair (cta_group::2) variant of the tcgen05 convolution kernel in conv_tc.cu.
//
// Same math, layouts and epilogue as conv_tc.cu (read its header first).  What changes: two CTAs of a cluster (one TPC)
// work on two adjacent CTA tiles with ONE stream of MMAs of M = 256 issued by the leader CTA.  Each CTA supplies the
// A rows (pixels) of its own tile from its own shared memory, but only HALF of the weight rows: conv_tc.cu is bound by
// the tensor core's operand fetch from shared memory (~0.33 cycles per 32-byte operand row), and in a pair the B operand
// is fetched once per two m-tiles.  Per (tap, 16-channel chunk, m-tile) a CTA fetches 128 + NT rows for a_hi x [w_hi|w_lo]
// and 128 + NT/2 rows for a_lo x w_hi instead of 128 + 2*NT and 128 + NT.
//
// Weight blocks are packed per CTA rank (engine.cu: pack_tc_pair):
//   rank 0:  main = w_hi rows [0,NT)   | X = w_hi rows [0,NT/2)
//   rank 1:  main = w_lo rows [0,NT)   | X = w_hi rows [NT/2,NT)
// so that a_hi x [main0 ; main1] = [a_hi*w_hi | a_hi*w_lo] (N' = 2*NT) and a_lo x [X0 ; X1] = a_lo*w_hi (N = NT), with the
// same shared-memory offsets in both CTAs (one descriptor addresses both halves).
//
// Protocol (after cutlass/pipeline/sm100_pipeline.hpp, PipelineTmaUmmaAsync with a 2-SM MMA):
//   * every barrier exists in both CTAs at the same offset; "full" barriers are only waited on in the leader: both CTAs'
//     TMA loads (.cta_group::2) complete_tx on the LEADER's barrier, the leader's producer arms it for 2x the bytes;
//   * "empty" / acc_full barriers are released by tcgen05.commit.cta_group::2 ... multicast::cluster with mask 0b11, i.e. they
//     fire in both CTAs, so each CTA's producers and epilogue wait on their local copy;
//   * acc_empty lives in the leader and counts the epilogue threads of BOTH CTAs (remote mbarrier.arrive via mapa).
//
// FIRST mode (ConvParams::first_w): the detector's conv1a (1 -> 64, net.py:60) is computed INSIDE conv1b's kernel.  conv1a's output
// is 19.7 MB per frame; written by its own kernel and read back by conv1b it was 10 GB of HBM traffic per 256-frame step and a
// write-bound 1.5 ms kernel (10 % of the step).  Here six producer warps per CTA (warps 8-11, 0 and 3; the epilogue, fully overlapped
// for NT = 64, shrinks to warps 4-7) read the u8 frame window of the tile (+2 pixels), evaluate conv1a + BN + ReLU on the CUDA cores
// with the same FMA order as conv_first_kernel (bit-identical values), split to fp16 hi/lo and store the halo tile straight into
// the A stage in the layout the TMA box would have produced, 16 channels per stage.  Weights and BN constants are constant-bank
// operands (by-value parameter).  generic-proxy stores -> fence.proxy.async -> release.cluster arrive on the leader's a_full
// (count 12 = 6 warps x 2 CTAs); conv1b's zero padding = zeros for halo pixels outside the image.
//
// FLAT mode (ConvParams::flat_in, F2 tensors of common.cuh).  RefineNet's first maps are 22x22 ... 8x8 pixels: a 16x8 /
// 16x16 pixel tile of such a map is mostly empty (8x8: 25 % of the MMA rows useful).  In FLAT mode all images of the launch
// form ONE run of pixels j = k*period + y*row + x, an m-tile is 128 CONSECUTIVE pixels of that run and a tap (ky, kx) is the
// same tile shifted by (ky-pad)*row + (kx-pad) pixels -- still nine descriptor start addresses into one halo buffer (core
// matrix = 8 consecutive pixels, SBO = 128 B).  Valid convolutions compute and discard the wrap-around positions
// (x >= wout or y >= hout); same-padded maps carry one zero gutter column / row that pads every neighbour (never written:
// the epilogue only stores data positions).  Useful rows: 20x20 of 22x22 = 83 %, 18x18 of 20x20 = 81 %, 8x8 of 9x9 = 79 %.
// The halo arrives as one 4-D TMA box {16 pixels, R rows of 16 pixels, 2 k-groups, hi|lo}.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace dcu {

namespace {

constexpr int t2_threads(int cg) { return 128 + 256 * cg; }     // 4 control warps + 8 epilogue warps per channel group

// WRES (weights resident): a 64 -> 64 layer's whole weight set (4 chunks x 3 kernel rows = 12 stages, 110.6 kB per rank) stays in
// shared memory for the lifetime of the CTA instead of being re-fetched for every tile.  UP + WRES (RefineNet convPa, the longest
// kernel of the step): a work item is (tile, row phase) and needs the 4 stages of ITS row phase (4 chunks x [2 kernel rows x 4 blocks],
// 98 kB per rank); every cluster keeps one row phase for its lifetime (even clusters phase 0, odd clusters phase 1), so those stay
// resident too.  Without it the kernel asks L2 for 48 B / cycle / SM (83 kB of halo + 96 kB of weights per 3.7 k-cycle item) = 7.1 kB /
// cycle chip-wide against a measured L2 limit of ~6.3 kB / cycle: tensor pipe 66 % active (profiles/r2_ncu_step_full.txt).  The kernel is bound by the shared-memory
// data pipe (tensor-core operand fetches + fills, profiles/r1_ncu_conv_tc2.txt: 77 % + 23 %); the weight refills were 10 % of it.
// (The 66 % on convPa turned out to be its epilogue: indexed constant loads, see `ch_base` in the kernel.)
// MT_ = 1: one 128-pixel m-tile per CTA (16 x 8 pixels, FLAT: a run of 128 pixels) instead of two; not for UP (its two m-tiles are the
// column phases).  Two uses: (a) NT = 64, launches with few work items (single frames): twice the items with half the MMA chain each;
// (b) NT = 128, the default for 128- / 512-channel layers: 256 TMEM columns per accumulator set, so the sets are double-buffered like
// NT = 64's, and both MMAs of a tap are math-bound (N = 256 and 128; at NT = 64 the N = 64 correction MMA reads 5 kB of operands from
// shared memory, 40 cycles, for 32 cycles of math).  The second group of epilogue warps, which has no m-tile, takes the upper half of
// the channels (HALF in the kernel).
template <int NT, bool UP = false, bool WRES = false, int MT_ = 2>
struct Tc2Cfg {
  static constexpr int MT = MT_;
  static constexpr int STAGE_BLOCKS = UP ? 8 : 3;                  // weight blocks per stage (one kernel row; UP: the chunk's 2 x 2 taps x 2 column phases)
  static constexpr int NBUF = 512 / (MT * 2 * NT);
  static constexpr int A_STAGES = UP ? 6 : 4;                      // UP: 4 taps per chunk -> a chunk is consumed in ~900 cycles; six stages keep ~2.5 us of loads in flight
  // WRES: every stage of the layer (UP: of one row phase).  Streamed weights: as many stages as shared memory holds -- the MMA warp of
  // the streamed launches spent 27-40 % of its time waiting for weight stages (tools/tc_stats.py), i.e. ~1.4 us of loads in flight
  // did not cover the L2 latency under load
  static constexpr int B_STAGES = WRES ? (UP ? 4 : 12) : (UP ? ((NT == 64) ? 5 : 2) : ((NT == 64) ? 6 : (MT_ == 1 ? 7 : 4)));
  static constexpr int MAX_HALO_PX = UP ? 240 : 34 * 10;           // UP: 10 x 18 low-resolution pixels (FLAT: up to 15 rows of 16)
  static constexpr int A_STAGE_BYTES = 4 * MAX_HALO_PX * 16;
  static constexpr int B_MAIN_BYTES = 2 * NT * 16;                 // 2 k-groups x NT rows x 8 fp16 (this rank's half of [w_hi | w_lo])
  static constexpr int B_X_BYTES = 2 * (NT / 2) * 16;              // 2 k-groups x NT/2 rows     (this rank's half of w_hi)
  static constexpr int B_BLOCK_BYTES = B_MAIN_BYTES + B_X_BYTES;   // 48 * NT
  static constexpr int B_STAGE_BYTES = STAGE_BLOCKS * B_BLOCK_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int WIN_ELEMS = 36 * 12;                        // FIRST: input window (halo + 2) of the largest tile arrangement
  static constexpr int WIN_BYTES = 2 * WIN_ELEMS * 4 + 768 * 4;    // two windows (double-buffered) + conv1a's weight / BN table, fp32
  static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + BAR_BYTES + WIN_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "227 kB of shared memory per CTA");
};

// -DDCU_TC2_STATS: role-level cycle counters (MMA warp of every leader CTA, epilogue warp 4 of every leader CTA) for tools/tc_stats.py;
// a profiling build only -- the shipped library compiles the counters out
#ifdef DCU_TC2_STATS
#define T2_STATS(...) __VA_ARGS__
#else
#define T2_STATS(...)
#endif

struct Tc2Geo {
  int tr, tc, halo_w, halo_h, tiles_x, tiles_y, slices;
  long long tiles_per_slice, pairs_per_slice, total_pairs;
  // A-operand addressing in 16-byte pixels: tap (ky, kx) of m-tile mt starts at a_org + ky*row_step + kx + mt_off;
  // sbo = distance between consecutive groups of 8 M rows.  Standard: row_step = sbo = halo_w, a_org = 0.
  int row_step, sbo, a_org;
  int flat, tile_px, lead;     // FLAT: pixels per CTA tile (256; UP or one m-tile: 128), pixels loaded ahead of the tile start (multiple of 16)
  int seg0, segc;              // SEG: 16-channel chunks in the FIRST segment of a tile and in every later segment (>= 1).  Each drain costs
                               // 128 TMEM columns x 128 lanes of tcgen05.ld per m-tile (64 B / cycle / SM): 4-chunk layers afford 2 segments
  int epi_pipe;                // epilogue loads one column group ahead (launches with >= 4 work items per cluster)
  int slice_minor;             // work items ordered tile-major (item = pair * slices + slice): the slices of one pixel tile run at the same
                               // time on neighbouring clusters, so its halo is read from HBM once and from L2 by the other slices
  H2Layout out;                // output addressing
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in the even (leader) CTA of the pair

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// release at cluster scope: publishes this thread's (fenced) shared-memory stores to the waiter in the leader CTA
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();
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
template <int kBackoffNs = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (kBackoffNs > 0) __nanosleep(kBackoffNs);
    if (clock64() - t0 > 4000000000LL) __trap();      // a protocol bug must fail the launch, never hang the GPU box
  }
}

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

// both CTAs execute these; the transaction bytes are credited to the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3,
                                                int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma2_f16_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// commit: arrive (once) on the barrier at this offset in BOTH CTAs when all prior MMAs of the pair have retired
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(smem_u32(bar))
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

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, float* a, uint32_t tb, float* b) {
  tmem_ld16_nowait(ta, a);
  tmem_ld16_nowait(tb, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  x0 = fminf(x0, 65504.f); x1 = fminf(x1, 65504.f);
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ void split_h2_noclamp(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(hh);
  const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
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

__device__ __forceinline__ int pair_slice(long long pt, const Tc2Geo& g) {
  return g.slice_minor ? (int)(pt % g.slices) : (int)(pt / g.pairs_per_slice);
}
struct Tile2 { int img, slice, y0, x0; bool valid; };       // FLAT: x0 = first pixel of the CTA tile in the run, img / y0 unused
__device__ __forceinline__ Tile2 decode_pair_tile(long long pt, uint32_t rank, const Tc2Geo& g) {
  Tile2 c;
  c.slice = pair_slice(pt, g);
  long long t = (g.slice_minor ? pt / g.slices : pt - (long long)c.slice * g.pairs_per_slice) * 2 + rank;
  c.valid = t < g.tiles_per_slice;
  if (!c.valid) t = g.tiles_per_slice - 1;            // odd tail: the peer recomputes the last tile and discards it
  if (g.flat) { c.img = 0; c.y0 = 0; c.x0 = (int)(t * g.tile_px); return c; }
  const int tx = (int)(t % g.tiles_x); t /= g.tiles_x;
  const int ty = (int)(t % g.tiles_y);
  c.img = (int)(t / g.tiles_y);
  c.y0 = ty * 16 * g.tr;
  c.x0 = tx * 8 * g.tc;
  return c;
}

struct NoFirst {};
// SEG (two-level accumulation).  The tensor core adds each MMA's 16 exact products to the fp32 accumulator with TRUNCATION
// (tools/mma_probe.py): over a chain of 36 - 72 MMAs the losses are one-sided and add up coherently (measured max
// |dloc| 0.055 on logits of +-100 against 0.0024 for an fp32 FMA chain).  In SEG mode the chain in tensor memory is ONE 16-channel
// chunk (9 taps; 4 in UP mode): the MMA warp starts a fresh accumulator per chunk (ping-pong over the NBUF accumulator sets) and the
// epilogue warps drain every finished chunk into fp32 registers with round-to-nearest adds while the next chunk is being issued.
// Partial sums stay small (shorter chains lose less per step) and the register adds are unbiased.  NT = 64 only (64 running sums per
// epilogue thread); 128-channel layers run as 64-channel slices in this mode.
// CG (channel groups, SEG only): the NT channels of an m-tile are shared by CG sets of four epilogue warps (CG = 2: 16 epilogue warps,
// 32 running sums each).  The accumulator sets are handed back per CHUNK, so the issuer can only run NBUF - 1 chunks ahead of the
// epilogue: the per-tile tail (BN, ReLU, pooling, fp16 split, stores) has to fit into that window, and with 64 channels per thread it
// did not (measured: + 35 % on conv1b).  Half the channels per thread halves both the tail and the drains.
template <int NT, int KS, bool UP, bool WRES, bool FIRST, int MT_ = 2, bool SEG = false, int CG = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(t2_threads(CG), 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w0,
                const __grid_constant__ CUtensorMap tmap_w1, const ConvParams p, const Tc2Geo g_host, const __grid_constant__ TcBn bn,
                const __grid_constant__ typename std::conditional<FIRST, FirstWeights, NoFirst>::type fw) {
  // bn: bias / BN scale / BN shift by value = constant bank.  The epilogue's channel index is warp-uniform, so these become
  // uniform constant loads instead of shared-memory reads (the shared-memory data pipe is what bounds this kernel).
  using Cfg = Tc2Cfg<NT, UP, WRES, MT_>;
  constexpr int MT = Cfg::MT, NBUF = Cfg::NBUF, B_STAGES = Cfg::B_STAGES;
  static_assert(!SEG || (NT == 64 && !FIRST && NBUF >= 2), "SEG: 64-channel slices, double-buffered accumulators, one m-tile per epilogue group");
  static_assert(CG == 1 || (SEG && CG == 2), "channel groups exist for the two-level accumulation only");
  constexpr int NTG = NT / CG;          // channels per epilogue thread
  // one m-tile per CTA: the second group of four epilogue warps has no m-tile of its own and takes the upper half of the channels
  constexpr bool HALF = MT == 1 && !SEG && !FIRST;
  constexpr int NTE = HALF ? NT / 2 : NTG;
  // UP (input = 2x nearest upsampling of the tensor in HBM): per output phase (a, b) the 3x3 taps collapse to 2x2 taps on
  // the low-resolution tensor (see header).
  // weight stages per chunk (ROWS) and kernel rows inside one stage (KYS): 3 x 1 for a 3x3 kernel; UP: ONE stage with both rows of the
  // collapsed 2x2 kernel -- 16 MMAs per elected issue region instead of 8 (every region costs the issuing warp ~100 cycles)
  constexpr int ROWS = UP ? 1 : KS, KYS = UP ? 2 : 1, TPR = UP ? 2 : KS, SB = Cfg::STAGE_BLOCKS, TAPS = ROWS * SB;
  constexpr int ROWS_PER_BLOCK = Cfg::B_BLOCK_BYTES / 512;          // weight tensor map rows (512 B each) per block
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* a_smem = smem_raw;
  uint8_t* b_smem = a_smem + Cfg::A_STAGES * Cfg::A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_smem + B_STAGES * Cfg::B_STAGE_BYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + Cfg::A_STAGES;
  uint64_t* b_full = a_empty + Cfg::A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* acc_full = b_empty + B_STAGES;
  uint64_t* acc_empty = acc_full + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NBUF * MT);
  float* in_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);     // FIRST: input window of the tile

  // Programmatic dependent launch (launch attribute set by launch_pair): let the NEXT kernel of the stream start its prologue
  // (barrier init, TMEM allocation, tensor-map and weight prefetch) on SMs this grid leaves idle or has left; it blocks in its own
  // griddepcontrol.wait until this grid has completed and flushed.  Without the attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // Image count from device memory (ConvParams::n_dev; sync-free detector -> RefineNet hand-off): the geometry the host computed for
  // the upper bound p.n shrinks to the images that exist.  No programmatic-dependency wait is needed before this read (it would
  // cost the overlap of this kernel's prologue and weight prefetch with its predecessor's tail): the count is written by the decode
  // kernel, and RefineNet's first kernel (conv_first_kernel) is an ordinary launch that starts only after the decode has completed;
  // every kernel of this template comes later in the stream than that launch (engine.cu: refine_run).
  Tc2Geo g = g_host;
  int n_imgs = p.n;
  if (p.n_dev != nullptr) {
    n_imgs = max(0, min(*reinterpret_cast<const volatile int*>(p.n_dev) - p.n_off, p.n));
    g.tiles_per_slice = g.flat ? ((long long)n_imgs * p.in_period + g.tile_px - 1) / g.tile_px : (long long)n_imgs * g.tiles_x * g.tiles_y;
    g.pairs_per_slice = (g.tiles_per_slice + 1) / 2;
    g.total_pairs = g.pairs_per_slice * g.slices;
  }
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int chunks = p.cin >> 4;
  const int halo_px = g.halo_w * g.halo_h;
  // this cluster's work items: pt_begin, pt_begin + pt_step, ... < pt_end.  UP + WRES: the items of ONE row phase (slice-major numbering)
  long long pt_begin = cluster_id, pt_end = g.total_pairs, pt_step = n_clusters;
  if (UP && WRES) {
    const long long ph = cluster_id & 1;
    pt_begin = ph * g.pairs_per_slice + (cluster_id >> 1); pt_end = (ph + 1) * g.pairs_per_slice; pt_step = n_clusters >> 1;
  }

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::A_STAGES; ++i) { mbar_init(&a_full[i], FIRST ? 12 : 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < NBUF; ++i) mbar_init(&acc_full[i], 1);
    for (int i = 0; i < NBUF * MT; ++i) mbar_init(&acc_empty[i], 256 * (HALF ? 2 : CG));      // epilogue threads of both CTAs (leader's copy is used)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                 // barriers of BOTH CTAs initialised before anyone signals across the pair
  __syncthreads();                    // (implied by the cluster barrier; compute-sanitizer's racecheck only models the CTA barrier and
                                      //  otherwise reports tcgen05.alloc's write of tmem_slot against the reads below)
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (FIRST && (warp >= 8 || warp == 0 || warp == 3)) {
    // ================= conv1a producers (both CTAs): frame window -> conv1a + BN + ReLU -> hi/lo halo tile in the A stage =================
    if constexpr (FIRST) {
      // six producer warps: 8-11 plus the two that have no other role in this mode (0: no TMA activation loads, 3: spare)
      const int pw = warp >= 8 ? warp - 8 : (warp == 0 ? 4 : 5);
      const int ptid = pw * 32 + lane;
      constexpr int PT = 192;
      const int win_w = g.halo_w + 2, win_h = g.halo_h + 2;
      int H = p.hin, W = p.win;
      const uint8_t* f_u8 = p.first_u8;
      const float* f_f32 = p.first_f32;
      // keep these in registers: re-reading them from the parameter bank inside the tile loop missed the constant cache every time
      // (the epilogue's 6 kB BN table shares it) -- 8 % of the producers' time on one LDCU
      asm volatile("" : "+r"(H), "+r"(W), "+l"(f_u8), "+l"(f_f32));
      // warp role: cg = which 8 of a chunk's 16 channels, pg = which half of the pixels; a thread's pixels (pg*32 + lane + 64k) are
      // the same for every tile, so their (row, column) inside the halo is computed once
      const int cg = pw & 1, pg = pw >> 1;
      uint32_t hyx[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = pg * 32 + lane + 96 * k;
        const int hy = px / g.halo_w;
        hyx[k] = ((uint32_t)hy << 16) | (uint32_t)(px - hy * g.halo_w);
      }
      // conv1a's table in shared memory, channel-major per 8-channel group: [group][9 taps + bias, alpha, beta][8] -> 24 LDS.128 per
      // (chunk, warp); indexed constant-bank loads (LDC) turned out to issue at ~1 per 15 cycles per SM and starved the producers
      float* wtab = in_s + 2 * Cfg::WIN_ELEMS;
      for (int i = ptid; i < 768; i += PT) {
        const int grp8 = i / 96, r = (i - grp8 * 96) >> 3, j = i & 7, ch = grp8 * 8 + j;
        wtab[i] = r < 9 ? fw.w[r * 64 + ch] : (r == 9 ? fw.bias[ch] : (r == 10 ? fw.alpha[ch] : fw.beta[ch]));
      }
      // the tile's frame window is fetched one tile ahead into registers (global latency hidden behind the previous tile's math)
      const int win_n = win_w * win_h;
      uint32_t pre[3];                                           // raw bits (u8 value or fp32 pattern): no arithmetic on them until they are used
      auto prefetch = [&](long long pt) {
        const Tile2 c = decode_pair_tile(pt, rank, g);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = ptid + PT * k;
          const int iy = i / win_w, ix = i - iy * win_w;
          const int gy = c.y0 - 2 + iy, gx = c.x0 - 2 + ix;
          uint32_t v = 0xffffffffu;                              // marker: outside the image (a NaN pattern, never a normalised value)
          if (i < win_n && gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = ((size_t)c.img * H + gy) * W + gx;
            v = f_u8 ? (uint32_t)f_u8[o] : __float_as_uint(f_f32[o]);
          }
          pre[k] = v;
        }
      };
      asm volatile("griddepcontrol.wait;" ::: "memory");
      prefetch(pt_begin);
      int st = 0; uint32_t ph = 0, wb = 0;
      for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
        const Tile2 c = decode_pair_tile(pt, rank, g);
        float* win = in_s + wb * Cfg::WIN_ELEMS;
        wb ^= 1u;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int i = ptid + PT * k;
          if (i < win_n) {
            // conv1a's zero padding is applied to the NORMALISED image (net.py:23); (x - 128) / 255 as model_utils.py:48-49
            const uint32_t v = pre[k];
            win[i] = (v == 0xffffffffu) ? 0.f : (f_u8 ? __fdiv_rn((float)v - 128.0f, 255.0f) : __uint_as_float(v));
          }
        }
        asm volatile("bar.sync 1, 192;" ::: "memory");          // window (and, first time, the table) visible to the six producer warps
        if (pt + pt_step < pt_end) prefetch(pt + pt_step);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          // this warp's 8 channels of the chunk: weights and BN constants into registers (FFMA operands), reused for all its pixels
          float wr[9][8], cb[8], ca[8], ce[8];
          {
            const float4* tp = reinterpret_cast<const float4*>(wtab + (q * 2 + cg) * 96);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const float4 lo4 = tp[2 * t], hi4 = tp[2 * t + 1];
              wr[t][0] = lo4.x; wr[t][1] = lo4.y; wr[t][2] = lo4.z; wr[t][3] = lo4.w;
              wr[t][4] = hi4.x; wr[t][5] = hi4.y; wr[t][6] = hi4.z; wr[t][7] = hi4.w;
            }
            const float4 b0 = tp[18], b1 = tp[19], a0 = tp[20], a1 = tp[21], e0 = tp[22], e1 = tp[23];
            cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
            ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w;
            ce[0] = e0.x; ce[1] = e0.y; ce[2] = e0.z; ce[3] = e0.w; ce[4] = e1.x; ce[5] = e1.y; ce[6] = e1.z; ce[7] = e1.w;
          }
          mbar_wait<40>(&a_empty[st], ph ^ 1u);
          uint4* stage = reinterpret_cast<uint4*>(a_smem + (size_t)st * Cfg::A_STAGE_BYTES) + cg * halo_px;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int px = pg * 32 + lane + 96 * k;
            if (px >= halo_px) continue;
            const int hy = (int)(hyx[k] >> 16), hx = (int)(hyx[k] & 0xffffu);
            const int gy = c.y0 - 1 + hy, gx = c.x0 - 1 + hx;
            if (!(gy >= 0 && gy < H && gx >= 0 && gx < W)) {                 // outside the image: conv1b's own zero padding
              stage[px] = make_uint4(0u, 0u, 0u, 0u);
              stage[2 * halo_px + px] = make_uint4(0u, 0u, 0u, 0u);
              continue;
            }
            const float* ip = win + hy * win_w + hx;
            float v[9];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) v[ky * 3 + kx] = ip[ky * win_w + kx];
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t)                                                // taps ascending from 0: conv_first_kernel's order;
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = fmaf(v[t], wr[t][j], y[j]);           // 8 independent chains interleaved
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(fmaf(y[j] + cb[j], ca[j], ce[j]), 0.0f);
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split_h2_noclamp(y[2 * e], y[2 * e + 1], h[e], l[e]);   // |conv1a| < 65504 is checked on the host
            stage[px] = make_uint4(h[0], h[1], h[2], h[3]);                            // hi plane of this k-group
            stage[2 * halo_px + px] = make_uint4(l[0], l[1], l[2], l[3]);              // lo plane
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                 // generic-proxy stores -> visible to the tensor core
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&a_full[st], 0);     // release (default) after the proxy fence, as cutlass' producer_commit
          if (++st == Cfg::A_STAGES) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (!FIRST && warp == 0 && lane == 0) {
    // ================= activation producer: this CTA's halo into this CTA's smem, bytes credited to the leader =================
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");        // the activations are the previous kernel's output (weights are not: no wait there)
    int st = 0; uint32_t ph = 0;
    for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
      const Tile2 c = decode_pair_tile(pt, rank, g);
      for (int q = 0; q < chunks; ++q) {
        mbar_wait<200>(&a_empty[st], ph ^ 1u);
        if (rank == 0) mbar_expect_tx(&a_full[st], 2u * (uint32_t)halo_px * 64u);
        if (g.flat)
          tma_load_4d_2sm(smem_u32(a_smem + (size_t)st * Cfg::A_STAGE_BYTES), &tmap_a, &a_full[st], 0, (c.x0 - g.lead) >> 4,
                          (p.cin_offset >> 3) + q * 2, 0);
        else
          tma_load_5d_2sm(smem_u32(a_smem + (size_t)st * Cfg::A_STAGE_BYTES), &tmap_a, &a_full[st], (c.x0 - p.pad) * 8, c.y0 - p.pad,
                          (p.cin_offset >> 3) + q * 2, 0, c.img);
        if (++st == Cfg::A_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 2 && lane == 0) {
    // ================= weight producer: this rank's half blocks (2-D tensor map over 512-byte rows) =================
    const CUtensorMap* wm = rank ? &tmap_w1 : &tmap_w0;
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(wm)) : "memory");
    int st = 0; uint32_t ph = 0;
    for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
      if (WRES && pt != pt_begin) break;              // resident weights (one slice): loaded with the first tile only
      const int slice = pair_slice(pt, g);
      const int blk0 = slice * chunks * TAPS;
      for (int blk = 0; blk < chunks * TAPS; blk += SB) {
        mbar_wait<200>(&b_empty[st], ph ^ 1u);
        if (rank == 0) mbar_expect_tx(&b_full[st], 2u * (uint32_t)(SB * Cfg::B_BLOCK_BYTES));
        tma_load_2d_2sm(smem_u32(b_smem + (size_t)st * Cfg::B_STAGE_BYTES), wm, &b_full[st], 0, (blk0 + blk) * ROWS_PER_BLOCK);
        if (++st == B_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================= MMA issuer (leader CTA only): M = 256 spans both CTAs' m-tiles =================
    constexpr uint32_t IDESC_BASE = (1u << 4) | (0u << 7) | (0u << 10) | (16u << 24);    // f32 accum, f16 x f16, K-major, M = 256
    constexpr uint32_t IDESC_2N = IDESC_BASE | ((uint32_t)((2 * NT) >> 3) << 17);
    constexpr uint32_t IDESC_1N = IDESC_BASE | ((uint32_t)(NT >> 3) << 17);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t a_desc_hi = (uint32_t)g.sbo | (1u << 14);                    // SBO = halo_w * 16 B (FLAT: 128 B)
    const uint32_t a_desc_lo0 = ((uint32_t)halo_px << 16);                      // LBO = plane
    constexpr uint32_t b_desc_hi = 8u | (1u << 14);                             // SBO = 128 B
    constexpr uint32_t bm_desc_lo0 = ((uint32_t)NT << 16);                      // main half: NT rows per k-group
    constexpr uint32_t bx_desc_lo0 = ((uint32_t)(NT / 2) << 16);                // X half: NT/2 rows per k-group
    const uint32_t a_base0 = (smem_u32(a_smem) & 0x3FFFFu) >> 4, b_base0 = (smem_u32(b_smem) & 0x3FFFFu) >> 4;
    uint32_t mt_off[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int tri = mt / g.tc, tci = mt - tri * g.tc;
      mt_off[mt] = g.flat ? (uint32_t)(mt * 128) : (uint32_t)(tri * 16 * g.halo_w + tci * 8);
    }
    int sa = 0, sb = 0, buf = 0; uint32_t pha = 0, phb = 0, phc = 0;
    T2_STATS(long long w_a = 0, w_b = 0, w_c = 0, t_w = 0; const long long t_start = clock64();)
    for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
      const uint32_t ph_a = UP ? (uint32_t)(pair_slice(pt, g) & 1) : 0u;     // row phase of this work item
      for (int q = 0; q < chunks; ++q) {
        T2_STATS(t_w = clock64();)
        mbar_wait(&a_full[sa], pha);
        T2_STATS(w_a += clock64() - t_w;)
        tc_fence_after();
        const uint32_t a_hi = a_desc_lo0 + a_base0 + (uint32_t)sa * (Cfg::A_STAGE_BYTES >> 4);
        const bool seg_open = !SEG || q == 0 || (q >= g.seg0 && (q - g.seg0) % g.segc == 0);
        const bool seg_close = !SEG || q == chunks - 1 || (q >= g.seg0 - 1 && (q - g.seg0 + 1) % g.segc == 0);
        if (SEG ? seg_open : q == 0) {          // SEG: a fresh accumulator set per segment
#pragma unroll
          T2_STATS(t_w = clock64();)
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) mbar_wait(&acc_empty[buf * MT + mt], phc ^ 1u);
          T2_STATS(w_c += clock64() - t_w;)
          tc_fence_after();
        }
#pragma unroll 1
        for (int ky = 0; ky < ROWS; ++ky) {
          if (!WRES || pt == pt_begin) {
            T2_STATS(t_w = clock64();)
            mbar_wait(&b_full[sb], phb);
            T2_STATS(w_b += clock64() - t_w;)
            tc_fence_after();
          }
          const uint32_t b_row = b_base0 + (uint32_t)sb * (Cfg::B_STAGE_BYTES >> 4);
          if (elect_one()) {
#pragma unroll
            for (int kyy = 0; kyy < KYS; ++kyy) {
              const int kyr = ky * KYS + kyy;                     // kernel row (UP: row of the collapsed 2x2 kernel)
              const uint32_t a_row = a_hi + (uint32_t)(g.a_org + (kyr + (int)ph_a) * g.row_step);
#pragma unroll
              for (int kx = 0; kx < TPR; ++kx) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  // UP: m-tile mt is column phase b = mt of the same low-resolution tile: its own weight block, A window shifted by b
                  const uint32_t blk = UP ? (uint32_t)(kyy * 4 + kx * 2 + mt) : (uint32_t)kx;
                  const uint32_t b_main = bm_desc_lo0 + b_row + blk * (Cfg::B_BLOCK_BYTES >> 4);
                  const uint32_t b_x = bx_desc_lo0 + b_row + blk * (Cfg::B_BLOCK_BYTES >> 4) + (uint32_t)(Cfg::B_MAIN_BYTES >> 4);
                  const uint32_t d = tmem_u + (uint32_t)((buf * MT + mt) * 2 * NT);
                  const uint32_t da_hi = a_row + (uint32_t)kx + (UP ? (uint32_t)mt : mt_off[mt]);
                  const uint32_t da_lo = da_hi + 2u * (uint32_t)halo_px;
                  umma2_f16_w(d, da_hi, a_desc_hi, b_main, b_desc_hi, IDESC_2N, ((SEG ? (seg_open ? 0 : 1) : q) | kyr | kx) ? 1u : 0u);   // [a_hi*w_hi | a_hi*w_lo]
                  umma2_f16_w(d + NT, da_lo, a_desc_hi, b_x, b_desc_hi, IDESC_1N, 1u);                        // + a_lo*w_hi
                }
              }
            }
            if (!WRES) umma2_commit_mc(&b_empty[sb]);
            if (ky == ROWS - 1) umma2_commit_mc(&a_empty[sa]);
            if (ky == ROWS - 1 && (SEG ? seg_close : q == chunks - 1)) umma2_commit_mc(&acc_full[buf]);
          }
          __syncwarp();
          if (++sb == B_STAGES) { sb = 0; phb ^= 1u; }
        }
        if (++sa == Cfg::A_STAGES) { sa = 0; pha ^= 1u; }
        if (SEG && seg_close) { if (++buf == NBUF) { buf = 0; phc ^= 1u; } }
      }
      if (!SEG) { if (++buf == NBUF) { buf = 0; phc ^= 1u; } }
    }
    T2_STATS(if (lane == 0 && p.stats) {
      atomicAdd(p.stats + 0, (unsigned long long)(clock64() - t_start)); atomicAdd(p.stats + 1, (unsigned long long)w_a);
      atomicAdd(p.stats + 2, (unsigned long long)w_b); atomicAdd(p.stats + 3, (unsigned long long)w_c); })
  } else if (warp >= 4 && (!FIRST || warp < 8) && (!SEG || (((warp - 4) >> 2) & 1) < MT)) {
    // ================= epilogue (both CTAs, each drains its own 128 TMEM lanes) =================
    constexpr int CW = 16;
    const int grp = (!FIRST && ((warp - 4) & 4)) ? 1 : 0;       // m-tile of this warp.  FIRST: warps 4-7 drain both m-tiles (warps 8-11 are the conv1a producers)
    const int cg = (CG == 1) ? 0 : (warp - 4) >> 3;      // channel group: channels [cg * NTG, (cg + 1) * NTG) of the slice
    float racc[SEG ? NTG : 1];                           // SEG: this thread's running sums (one pixel x NTG channels of m-tile `grp`)
    const int q4 = warp & 3;
    const int m = q4 * 32 + lane;
    const int prow = m >> 3, pcol = m & 7;
    const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
    const int c8_out = p.cout_total >> 3;
    uint32_t phc = 0;
    int buf = 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");        // (ordered anyway through the activation loads; keeps the stores formally after the wait)
    T2_STATS(long long w_e = 0, t_w = 0; const long long t_start = clock64();)
    for (long long pt = pt_begin; pt < pt_end; pt += pt_step) {
      const Tile2 c = decode_pair_tile(pt, rank, g);
      if constexpr (SEG) {
        // drain every finished chunk of m-tile `grp` into registers (RN adds), releasing its accumulator set for the chunk after next
#pragma unroll
        for (int j = 0; j < NTG; ++j) racc[j] = 0.f;
#pragma unroll 1
        for (int q = 0; q < chunks; ++q) {          // one pass per segment
          if (!(q == chunks - 1 || (q >= g.seg0 - 1 && (q - g.seg0 + 1) % g.segc == 0))) continue;
          mbar_wait<40>(&acc_full[buf], phc);
          tc_fence_after();
#pragma unroll
          for (int cc = 0; cc < NTG / CW; ++cc) {
            float v[CW], sm[CW];
            const uint32_t col = (uint32_t)((buf * MT + grp) * 2 * NT + cg * NTG + cc * CW);
            tmem_ld16x2(tmem_base + lane_addr + col, v, tmem_base + lane_addr + col + NT, sm);
#pragma unroll
            for (int j = 0; j < CW; ++j) racc[cc * CW + j] += v[j] + sm[j];
          }
          tc_fence_before();
          mbar_arrive_cluster(&acc_empty[buf * MT + grp], 0);      // always the leader's barrier
          if (++buf == NBUF) { buf = 0; phc ^= 1u; }
        }
      } else {
        T2_STATS(t_w = clock64();)
        mbar_wait<200>(&acc_full[buf], phc);
        T2_STATS(w_e += clock64() - t_w;)
        tc_fence_after();
      }
      // resident weights = ONE channel slice: every bias / BN / head constant of the unrolled epilogue below then has a compile-time
      // offset in the constant bank and is an instruction operand.  Indexed constant loads (LDC) issue at ~1 per 8-15 cycles per SM:
      // 104 of them per warp and tile made the epilogue the pacing unit of convPa + head (MMA warp 39 % in wait_epilogue).
      const int ch_base = WRES ? 0 : (UP ? (c.slice >> 1) : c.slice) * NT;
#pragma unroll 1
      for (int mt = HALF ? 0 : grp; mt < MT; mt += ((FIRST || HALF) ? 1 : 2)) {
        int oy, ox, img = c.img;
        bool inb;
        if (g.flat) {    // pixel j of the run -> (image, y, x) by the input's period / row stride; wrap-around positions are dropped
          const int j = c.x0 + (UP ? 0 : mt * 128) + m;
          img = j / p.in_period;
          const int r = j - img * p.in_period;
          const int y = r / p.in_row, x = r - y * p.in_row;
          inb = c.valid && img < n_imgs && y < (UP ? p.hin : p.hout) && x < (UP ? p.win : p.wout);
          oy = UP ? 2 * y + (c.slice & 1) : y; ox = UP ? 2 * x + mt : x;
        } else {
          if (UP) {        // low-resolution pixel (y0 + prow, x0 + pcol), output phase (slice & 1, mt)
            oy = 2 * (c.y0 + prow) + (c.slice & 1); ox = 2 * (c.x0 + pcol) + mt;
          } else {
            const int tri = mt / g.tc, tci = mt - tri * g.tc;
            oy = c.y0 + tri * 16 + prow; ox = c.x0 + tci * 8 + pcol;
          }
          inb = c.valid && (oy < p.hout) && (ox < p.wout);
        }
        float head_sum = 0.f;
        // one group of CW = 16 output channels of this thread's pixel: bias + BN + ReLU, then one of {fp32 logits, fused 1x1 head,
        // 2x2 max-pool + store, 2x2 replicated store, plain store}
        auto process = [&](const int cc, float (&v)[CW]) {
          const int ch0 = ch_base + cg * NTG + (HALF ? grp * NTE : 0) + cc * CW;
          if (p.logits != nullptr) {
            if (inb) {
              const size_t plane_o = (size_t)p.hout * p.wout;
              float* o = p.logits + (size_t)img * p.n_valid * plane_o + (size_t)oy * p.wout + ox;
#pragma unroll
              for (int j = 0; j < CW; ++j)
                if (ch0 + j < p.n_valid) o[(size_t)(ch0 + j) * plane_o] = v[j] + bn.v[0][ch0 + j];
            }
            return;
          }
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 bi = *reinterpret_cast<const float4*>(&bn.v[0][ch0 + j]);
            const float4 al = *reinterpret_cast<const float4*>(&bn.v[1][ch0 + j]);
            const float4 be = *reinterpret_cast<const float4*>(&bn.v[2][ch0 + j]);
            v[j + 0] = fmaxf(fmaf(v[j + 0] + bi.x, al.x, be.x), 0.0f);
            v[j + 1] = fmaxf(fmaf(v[j + 1] + bi.y, al.y, be.y), 0.0f);
            v[j + 2] = fmaxf(fmaf(v[j + 2] + bi.z, al.z, be.z), 0.0f);
            v[j + 3] = fmaxf(fmaf(v[j + 3] + bi.w, al.w, be.w), 0.0f);
          }
          if (p.head_w != nullptr) {
#pragma unroll
            for (int j = 0; j < CW; ++j) head_sum = fmaf(v[j], bn.head[cc * CW + j], head_sum);
          } else if (p.pool) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
              float x = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
              v[j] = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 8));
            }
            const int hp = p.hout >> 1, wp = p.wout >> 1;
            if (c.valid && ((lane & 9) == 0) && (oy >> 1) < hp && (ox >> 1) < wp) {
              uint4* o = reinterpret_cast<uint4*>(p.out) + (size_t)img * g.out.img + (size_t)(ch0 >> 3) * g.out.plane +
                         (size_t)(oy >> 1) * g.out.row + (ox >> 1);
              store_h2_16(o, (size_t)g.out.lo, (size_t)g.out.plane, v);
            }
          } else if (p.ups) {
            if (inb) {
              const int hu = p.hout * 2, wu = p.wout * 2;
              const size_t plane_o = (size_t)hu * wu;
              uint4* o = reinterpret_cast<uint4*>(p.out) + ((size_t)img * 2 * c8_out + (ch0 >> 3)) * plane_o + (size_t)(2 * oy) * wu + 2 * ox;
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
              uint4* o = reinterpret_cast<uint4*>(p.out) + (size_t)img * g.out.img + (size_t)(ch0 >> 3) * g.out.plane +
                         (size_t)oy * g.out.row + ox;
              store_h2_16(o, (size_t)g.out.lo, (size_t)g.out.plane, v);
            }
          }
        };
        // EPI_PIPE (64-channel slices): the tcgen05.ld of column group cc + 1 is in flight while group cc goes through BN / ReLU / pool /
        // split / stores.  tcgen05.ld moves 64 B / cycle / SM (~2 k cycles for the two m-tiles of a tile); load -> wait -> compute in
        // series made the epilogue the pacing unit of the short work items (upsample-fused layers: 4 taps per chunk).  The pipelined
        // loop is unrolled four times; launches with a handful of work items per cluster (single frames: 28 dependent kernels of
        // ~10 us) take the compact loop instead -- the larger code cost them 10 % (instruction-cache misses of a cold kernel).
#ifdef DCU_NO_EPI_PIPE
        constexpr bool EPI_PIPE = false;
#else
        constexpr bool EPI_PIPE = !SEG && (NT == 64 || HALF);
#endif
        if constexpr (SEG) {
#pragma unroll
          for (int cc = 0; cc < NTG / CW; ++cc) {
            float v[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = racc[cc * CW + j] * p.wscale_inv;
            process(cc, v);
          }
        } else if (EPI_PIPE && g.epi_pipe) {
          float vb[2][CW], sb[2][CW];
          const uint32_t col0 = tmem_base + lane_addr + (uint32_t)((buf * MT + mt) * 2 * NT + (HALF ? grp * NTE : 0));
          tmem_ld16_nowait(col0, vb[0]);
          tmem_ld16_nowait(col0 + NT, sb[0]);
#pragma unroll
          for (int cc = 0; cc < NTE / CW; ++cc) {
            float v[CW];
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (cc + 1 < NTE / CW) {
              tmem_ld16_nowait(col0 + (cc + 1) * CW, vb[(cc + 1) & 1]);
              tmem_ld16_nowait(col0 + (cc + 1) * CW + NT, sb[(cc + 1) & 1]);
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = (vb[cc & 1][j] + sb[cc & 1][j]) * p.wscale_inv;
            process(cc, v);
          }
        } else {
#pragma unroll 1
          for (int cc = 0; cc < NTE / CW; ++cc) {
            float v[CW], sm[CW];
            const uint32_t col = (uint32_t)((buf * MT + mt) * 2 * NT + (HALF ? grp * NTE : 0) + cc * CW);
            tmem_ld16x2(tmem_base + lane_addr + col, v, tmem_base + lane_addr + col + NT, sm);
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = (v[j] + sm[j]) * p.wscale_inv;
            process(cc, v);
          }
        }
        if (p.head_w != nullptr) {
          const float s = head_sum + p.head_b;
          unsigned long long key = 0ull;
          if (inb) {
            const unsigned int idx = (unsigned)(oy * p.wout + ox);
            if (p.heat != nullptr) p.heat[(size_t)img * p.hout * p.wout + idx] = s;
            key = ((unsigned long long)orderable(s) << 32) | (unsigned long long)(~idx);
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
          }
          if (lane == 0 && key != 0ull) atomicMax(p.head_key + img, key);
        }
        if constexpr (!SEG) {
          tc_fence_before();
          mbar_arrive_cluster(&acc_empty[buf * MT + mt], 0);      // always the leader's barrier
        }
      }
      if constexpr (!SEG) { if (++buf == NBUF) { buf = 0; phc ^= 1u; } }
    }
    T2_STATS(if (rank == 0 && warp == 4 && lane == 0 && p.stats) {
      atomicAdd(p.stats + 6, (unsigned long long)(clock64() - t_start)); atomicAdd(p.stats + 7, (unsigned long long)w_e); })
  }

  // ---- teardown: nobody may exit (or free TMEM) while the peer can still touch this CTA's barriers / tensor memory ----
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int NT, int KS, bool UP, bool WRES = false, bool FIRST = false, int MT_ = 2, bool SEG = false, int CG = 1>
cudaError_t launch_pair(const ConvParams& p, int n_slices, const CUtensorMap* ta, const CUtensorMap* w0, const CUtensorMap* w1,
                        int sm_count, cudaStream_t s, double* issued_flops) {
  using Cfg = Tc2Cfg<NT, UP, WRES, MT_>;
  static bool attr_done[64] = {};      // function attributes are per device: engines on several GPUs of one process each opt in
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<NT, KS, UP, WRES, FIRST, MT_, SEG, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  Tc2Geo g{};
  if (UP && (p.pad != 1 || p.pool || p.ups || p.hout != 2 * p.hin || p.wout != 2 * p.win)) return cudaErrorInvalidValue;
  if (MT_ == 1 && (UP || p.head_w != nullptr)) return cudaErrorInvalidValue;   // (fused head: all channels of a pixel in one thread)
  if (CG != 1 && p.head_w != nullptr) return cudaErrorInvalidValue;      // the fused 1x1 head sums over all channels of a pixel in one thread
  if (p.flat_in) {
    // the batch as one run of pixels; CTA tile = 256 consecutive pixels (UP: 128, the two m-tiles are the column phases)
    if (p.pool || p.ups || p.head_w || p.logits || p.in_period <= 0 || p.in_row <= 0) return cudaErrorInvalidValue;
    if ((long long)p.n * p.in_period + 1024 > 0x7fffffffLL) return cudaErrorInvalidValue;
    const int back = (UP || p.pad) ? p.in_row + 1 : 0;            // pixels a tap reaches behind its output position
    g.flat = 1; g.tile_px = (UP || MT_ == 1) ? 128 : 256;
    g.lead = ceil_div(back, 16) * 16;
    g.a_org = g.lead - back; g.row_step = p.in_row; g.sbo = 8;
    g.tr = 1; g.tc = 1;
    g.halo_w = 16; g.halo_h = tc2_flat_rows(p.in_row, (UP || p.pad) ? 1 : 0, (UP || MT_ == 1) ? 1 : 0);
    g.tiles_x = 1; g.tiles_y = 1;
    g.tiles_per_slice = ((long long)p.n * p.in_period + g.tile_px - 1) / g.tile_px;
    if (UP) n_slices *= 2;
  } else if (UP) {
    // tiles of 16 x 8 LOW-resolution pixels; the two m-tiles are the column phases; slices = channel slices x 2 row phases
    g.tr = 1; g.tc = 1;
    g.halo_w = 10; g.halo_h = 18;
    g.tiles_x = ceil_div(p.win, 8); g.tiles_y = ceil_div(p.hin, 16);
    n_slices *= 2;
  } else {
    if (MT_ == 1) { g.tr = 1; g.tc = 1; } else tc_tile_arrangement(NT, p.hout, p.wout, &g.tr, &g.tc);
    g.halo_w = 8 * g.tc + 2; g.halo_h = 16 * g.tr + 2;
    g.tiles_x = ceil_div(p.wout, 8 * g.tc); g.tiles_y = ceil_div(p.hout, 16 * g.tr);
  }
  if (!g.flat) {
    g.row_step = g.halo_w; g.sbo = g.halo_w; g.a_org = 0;
    g.tiles_per_slice = (long long)p.n * g.tiles_x * g.tiles_y;
  }
  if (g.halo_w * g.halo_h > Cfg::MAX_HALO_PX) return cudaErrorInvalidValue;
  g.out = p.out_layout.plane ? p.out_layout
                             : (p.pool ? h2_standard(p.cout_total, p.hout >> 1, p.wout >> 1) : h2_standard(p.cout_total, p.hout, p.wout));
  g.slices = n_slices;
  // chunks per segment (default 1 = a drain per 16-channel chunk; measured on B200, batch 256: 1/1 -> max |dloc| 0.0088 at 15.2 k
  // frames/s, 2/1 -> 0.019 at 15.6 k, 2/2 -> 0.021 at 15.9 k, off -> 0.055 at 17.3 k)
  static const int seg_c64 = [] { const char* v = getenv("DCU_SEG_CHUNKS64"); return v ? atoi(v) : 1; }();
  static const int seg_c128 = [] { const char* v = getenv("DCU_SEG_CHUNKS128"); return v ? atoi(v) : 1; }();
  static const int seg_first = [] { const char* v = getenv("DCU_SEG_FIRST"); return v ? atoi(v) : 0; }();
  g.segc = std::max(1, std::min(p.cin <= 64 ? seg_c64 : seg_c128, p.cin / 16));
  g.seg0 = std::max(1, std::min(seg_first > 0 ? seg_first : g.segc, p.cin / 16));
  static const bool slice_minor = [] { const char* v = getenv("DCU_SLICE_MINOR"); return !v || atoi(v) != 0; }();
  g.slice_minor = (slice_minor && n_slices > 1 && !(UP && WRES)) ? 1 : 0;
  g.pairs_per_slice = (g.tiles_per_slice + 1) / 2;
  g.total_pairs = g.pairs_per_slice * g.slices;
  if (g.total_pairs <= 0) return cudaSuccess;
  // per pair work item and (16-channel chunk, tap, m-tile): one M=256 x N=2*NT x K=16 and one M=256 x N=NT x K=16 MMA
  if (issued_flops) *issued_flops = 2.0 * (double)g.total_pairs * (p.cin / 16) * (UP ? 4 : KS * KS) * Cfg::MT * 256.0 * 3.0 * NT * 16.0;
  long long clusters = g.total_pairs < sm_count / 2 ? g.total_pairs : sm_count / 2;
  if (UP && WRES) {       // one row phase per cluster: an even number of clusters, each phase with at least one item per cluster
    if (g.slices != 2) return cudaErrorInvalidValue;
    clusters = std::min<long long>(sm_count / 2, 2 * g.pairs_per_slice) & ~1LL;
    if (clusters < 2) return cudaErrorInvalidValue;
  }
  if (p.host_bn == nullptr) return cudaErrorInvalidValue;
  g.epi_pipe = (g.total_pairs >= 4 * clusters) ? 1 : 0;
  static const bool pdl = [] { const char* v = getenv("DCU_PDL"); return !v || atoi(v) != 0; }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)clusters * 2, 1, 1); cfg.blockDim = dim3(t2_threads(CG), 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  if constexpr (FIRST) {
    if (p.first_w == nullptr || (!p.first_u8 && !p.first_f32) || p.cin != 64 || p.pad != 1 || g.flat ||
        (g.halo_w + 2) * (g.halo_h + 2) > Cfg::WIN_ELEMS || (g.halo_w + 2) * (g.halo_h + 2) > 576 || g.halo_w * g.halo_h > 384)
      return cudaErrorInvalidValue;
    return cudaLaunchKernelEx(&cfg, conv_tc2_kernel<NT, KS, UP, WRES, true>, *ta, *w0, *w1, p, g, *p.host_bn, *p.first_w);
  } else {
    return cudaLaunchKernelEx(&cfg, conv_tc2_kernel<NT, KS, UP, WRES, false, MT_, SEG, CG>, *ta, *w0, *w1, p, g, *p.host_bn, NoFirst());
  }
  return cudaGetLastError();
}

}  // namespace

int tc2_block_bytes(int nt) { return 48 * nt; }
// two-level accumulation (SEG template parameter) is the default; DCU_SEG=0 keeps whole-tile accumulation chains in tensor memory
// DCU_SEG: 0 (default) = whole-tile accumulation chains in tensor memory, 1 = two-level accumulation on every layer ("strict":
// max |dloc| 0.0088 instead of 0.055 at - 12 % frames/s; every drain re-reads 128 TMEM columns per m-tile at 64 B / cycle / SM,
// which is what it costs), 2 = on the layers with >= 128 input channels only.  Chunks per segment: DCU_SEG_CHUNKS64 / _CHUNKS128.
// Read per call: tests switch it between engines of one process.
int tc2_seg_policy() {
  const char* v = getenv("DCU_SEG");
  return v ? atoi(v) : 0;
}
bool tc2_segmented(int cin) {
  const int pol = tc2_seg_policy();
  return pol == 1 || (pol == 2 && cin >= 128);
}
int tc2_flat_rows(int in_row, int pad_or_up, int tile128) {
  const int back = pad_or_up ? in_row + 1 : 0, fwd = pad_or_up ? in_row + 1 : 2 * in_row + 2;
  return ceil_div(ceil_div(back, 16) * 16 + (tile128 ? 128 : 256) + fwd, 16);
}
int tc2_stage_blocks(int up) { return up ? 8 : 3; }

cudaError_t launch_conv_tc2(const ConvParams& p, int n_slices, int up, const void* tmap_a, const void* tmap_w0, const void* tmap_w1,
                            int sm_count, cudaStream_t s, double* issued_flops) {
  const CUtensorMap* ta = reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap* w0 = reinterpret_cast<const CUtensorMap*>(tmap_w0);
  const CUtensorMap* w1 = reinterpret_cast<const CUtensorMap*>(tmap_w1);
  const int nt = p.cout_total / n_slices;
  if (p.ksize == 1) return cudaErrorInvalidValue;      // the 1x1 heads stay on the single-CTA kernel
  const bool seg = tc2_segmented(p.cin) && nt == 64 && p.first_w == nullptr;
  if (seg) {
    static const int cg = [] { const char* v = getenv("DCU_SEG_CG"); return v ? atoi(v) : 2; }();
    static const bool wres_seg = [] { const char* v = getenv("DCU_WRES"); return !v || atoi(v) != 0; }();
    const bool wres = n_slices == 1 && p.cin == 64 && wres_seg;
    if (cg == 2 && p.head_w == nullptr) {
      if (up) return launch_pair<64, 3, true, false, false, 2, true, 2>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
      if (p.mt1) return launch_pair<64, 3, false, false, false, 1, true, 2>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
      if (wres) return launch_pair<64, 3, false, true, false, 2, true, 2>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
      return launch_pair<64, 3, false, false, false, 2, true, 2>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    }
    if (up) return launch_pair<64, 3, true, false, false, 2, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    if (p.mt1) return launch_pair<64, 3, false, false, false, 1, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    if (wres) return launch_pair<64, 3, false, true, false, 2, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    return launch_pair<64, 3, false, false, false, 2, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  }
  if (up) {
    static const bool wres_up = [] { const char* v = getenv("DCU_WRES_UP"); return !v || atoi(v) != 0; }();
    if (nt == 64 && n_slices == 1 && p.cin == 64 && wres_up && !p.flat_in && sm_count >= 4)
      return launch_pair<64, 3, true, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    if (nt == 64) return launch_pair<64, 3, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    if (nt == 128) return launch_pair<128, 3, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    return cudaErrorInvalidValue;
  }
  static const bool wres_ok = [] { const char* v = getenv("DCU_WRES"); return !v || atoi(v) != 0; }();
  if (p.first_w != nullptr) {
    if (nt != 64 || n_slices != 1) return cudaErrorInvalidValue;
    return wres_ok ? launch_pair<64, 3, false, true, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops)
                   : launch_pair<64, 3, false, false, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  }
  if (p.mt1) {
    if (up) return cudaErrorInvalidValue;
    if (nt == 128) return launch_pair<128, 3, false, false, false, 1>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
    if (nt != 64) return cudaErrorInvalidValue;
    return launch_pair<64, 3, false, false, false, 1>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  }
  if (nt == 64 && n_slices == 1 && p.cin == 64 && wres_ok) return launch_pair<64, 3, false, true>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  if (nt == 64) return launch_pair<64, 3, false>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  if (nt == 128) return launch_pair<128, 3, false>(p, n_slices, ta, w0, w1, sm_count, s, issued_flops);
  return cudaErrorInvalidValue;
}

}  // namespace dcu
