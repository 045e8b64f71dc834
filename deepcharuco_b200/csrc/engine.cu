// Engine + C ABI (include/deepcharuco_b200.h).  Owns packed weights, workspace, TMA descriptors and
// the launch sequence  detector -> decode+gather -> RefineNet  that replaces the body of
// inference.infer_image (/root/reference/src/inference.py:41-60) for whole batches.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges per stage for nsys / ncu --nvtx; no-ops when no tool is attached

#include "../../include/deepcharuco_b200.h"
#include "common.cuh"

using namespace dcu;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(DCU_ERR_CUDA, std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" + \
                                    __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
  } while (0)

namespace {

struct NvtxRange {          // host-side range around the launches of one stage
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t alloc(size_t b) {
    bytes = b;
    return cudaMalloc(&p, b ? b : 16);
  }
  void release() { if (p) cudaFree(p); p = nullptr; }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One 3x3 layer with cin % 4 == 0 (everything except the two first layers and the 1x1 heads)
struct Layer3x3 {
  int cin = 0, cout = 0;      // cout = total output channels (Pa|Da merged: 512)
  int pad = 1, pool = 0, ups = 0;
  DevBuf w_ffma;              // [cin/4][9][4][cout]
  DevBuf w_tc;                // tcgen05 blocks (conv_tc.cu layout), empty if shape unsupported
  float tc_scale = 1.f;       // power-of-two weight scale baked into w_tc
  int tc_copies = 1;          // replicas of w_tc (spread the all-CTA broadcast reads over more L2 slices)
  DevBuf w_tc2[2];            // CTA-pair packing per rank (conv_tc2.cu)
  CUtensorMap wmap2[2];       // 2-D tensor maps over w_tc2 (rows of 512 B, box = one 3-tap stage)
  bool ups_in = false;        // the layer's input is the 2x nearest upsampling of the previous layer's output (refinenet.py:66,71,76)
  DevBuf w_up[2];             // CTA-pair packing of the phase-collapsed 2x2 weights (pack_tc_pair_up), when ups_in
  CUtensorMap wmap_up[2];
  float up_scale = 1.f;
  int tc_nt = 0;              // N per CTA pass for the tcgen05 kernel (64 / 128), 0 = unsupported
  // the same layer in 64-channel slices (layers whose tc_nt is 128): used for launches with fewer work items than SMs (small
  // batches / single frames), where twice as many, half as long CTA-pair tiles cut the latency of the launch
  DevBuf w_tc2_64[2], w_up_64[2];
  CUtensorMap wmap2_64[2], wmap_up_64[2];
  bool has_64 = false;
  DevBuf bias, alpha, beta;   // [cout]
  TcBn host_bn{};             // the same three vectors, passed by value to the pair kernel
};

struct FirstLayer {
  DevBuf w, bias, alpha, beta;   // [9][64], [64] x3
  FirstWeights host{};           // the same values, passed by value to the fused conv1a + conv1b kernel (conv_tc2.cu FIRST mode)
  float out_bound = 0.f;         // upper bound of |output| for inputs in [-0.51, 0.5] (the fused kernel splits to fp16 without a clamp)
  int pad = 1;
};

std::vector<float> pack_ffma(const std::vector<const float*>& ws, const std::vector<int>& couts, int cin) {
  // OIHW (possibly several tensors concatenated along O) -> [cin/4][tap][c][cout_total]
  int cout_total = 0;
  for (int c : couts) cout_total += c;
  std::vector<float> out((size_t)cin * 9 * cout_total);
  int obase = 0;
  for (size_t t = 0; t < ws.size(); ++t) {
    for (int o = 0; o < couts[t]; ++o)
      for (int i = 0; i < cin; ++i)
        for (int tap = 0; tap < 9; ++tap)
          out[(((size_t)(i >> 2) * 9 + tap) * 4 + (i & 3)) * cout_total + obase + o] =
              ws[t][((size_t)o * cin + i) * 9 + tap];
    obase += couts[t];
  }
  return out;
}

std::vector<float> concat(const std::vector<const float*>& v, const std::vector<int>& n) {
  std::vector<float> out;
  for (size_t i = 0; i < v.size(); ++i) out.insert(out.end(), v[i], v[i] + n[i]);
  return out;
}

cudaError_t upload(DevBuf& b, const std::vector<float>& h) {
  cudaError_t e = b.alloc(h.size() * sizeof(float));
  if (e != cudaSuccess) return e;
  return cudaMemcpy(b.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice);
}

// tcgen05 weight blocks (fp16 hi/lo split, conv_tc.cu).  For slice s (NT output channels), chunk q (16 input
// channels), tap t:   block[(s*chunks + q)*9 + t] = [2 k-groups of 8 channels][2*NT rows: NT hi rows then NT lo rows][8 fp16]
// so that [w_hi | w_lo] is ONE K-major operand with N' = 2*NT.  Weights are pre-multiplied by `scale` (a power of two,
// exact) so that w_lo = fp16(w*scale - w_hi) stays in fp16's normal range; the kernel's epilogue multiplies by 1/scale.
float tc_weight_scale(const std::vector<const float*>& ws, const std::vector<int>& couts, int cin, int taps = 9) {
  float mx = 0.f;
  for (size_t t = 0; t < ws.size(); ++t)
    for (size_t i = 0; i < (size_t)couts[t] * cin * taps; ++i) mx = std::max(mx, std::fabs(ws[t][i]));
  if (!(mx > 0.f)) return 1.f;
  return std::exp2(std::floor(std::log2(32768.0f / mx)));      // |w*scale| < 65504 with a 2x margin
}

std::vector<uint16_t> pack_tc(const std::vector<const float*>& ws, const std::vector<int>& couts, int cin, int nt, float scale,
                              int taps = 9, int pad_rows_to = 0) {
  int cout_total = 0;
  for (int c : couts) cout_total += c;
  const int rows_total = std::max(cout_total, pad_rows_to);     // rows beyond cout_total are zero (1x1 heads padded to NT)
  const int slices = rows_total / nt, chunks = cin / 16;
  const size_t blk = (size_t)2 * (2 * nt) * 8;      // halves per block
  std::vector<uint16_t> out((size_t)slices * chunks * taps * blk, 0);
  std::vector<const float*> row(rows_total, nullptr);
  {
    int o = 0;
    for (size_t t = 0; t < ws.size(); ++t)
      for (int k = 0; k < couts[t]; ++k) row[o++] = ws[t] + (size_t)k * cin * taps;
  }
  for (int s = 0; s < slices; ++s)
    for (int q = 0; q < chunks; ++q)
      for (int tap = 0; tap < taps; ++tap) {
        uint16_t* b = out.data() + (((size_t)s * chunks + q) * taps + tap) * blk;
        for (int kg = 0; kg < 2; ++kg)
          for (int n = 0; n < nt; ++n)
            for (int e = 0; e < 8; ++e) {
              const int ci = q * 16 + kg * 8 + e;
              if (row[s * nt + n] == nullptr) continue;
              const float w = row[s * nt + n][(size_t)ci * taps + tap] * scale;
              const __half hi = __float2half_rn(w);
              const __half lo = __float2half_rn(w - __half2float(hi));
              uint16_t hb, lb;
              std::memcpy(&hb, &hi, 2); std::memcpy(&lb, &lo, 2);
              b[((size_t)kg * 2 * nt + n) * 8 + e] = hb;
              b[((size_t)kg * 2 * nt + nt + n) * 8 + e] = lb;
            }
      }
  return out;
}

// CTA-pair packing (conv_tc2.cu): per rank r, block = [main_r: 2 k-groups x NT rows][X_r: 2 k-groups x NT/2 rows] of 8 fp16,
//   main_0 = w_hi, main_1 = w_lo, X_r = w_hi rows [r*NT/2, (r+1)*NT/2).  Same block order as pack_tc.
std::vector<uint16_t> pack_tc_pair(const std::vector<const float*>& ws, const std::vector<int>& couts, int cin, int nt, float scale,
                                   int rank) {
  int cout_total = 0;
  for (int c : couts) cout_total += c;
  const int slices = cout_total / nt, chunks = cin / 16;
  const size_t main_h = (size_t)2 * nt * 8, x_h = (size_t)2 * (nt / 2) * 8, blk = main_h + x_h;
  std::vector<uint16_t> out((size_t)slices * chunks * 9 * blk, 0);
  std::vector<const float*> row(cout_total);
  {
    int o = 0;
    for (size_t t = 0; t < ws.size(); ++t)
      for (int k = 0; k < couts[t]; ++k) row[o++] = ws[t] + (size_t)k * cin * 9;
  }
  auto split = [&](float w, uint16_t& hb, uint16_t& lb) {
    const float ws_ = w * scale;
    const __half hi = __float2half_rn(ws_);
    const __half lo = __float2half_rn(ws_ - __half2float(hi));
    std::memcpy(&hb, &hi, 2); std::memcpy(&lb, &lo, 2);
  };
  for (int s = 0; s < slices; ++s)
    for (int q = 0; q < chunks; ++q)
      for (int tap = 0; tap < 9; ++tap) {
        uint16_t* b = out.data() + (((size_t)s * chunks + q) * 9 + tap) * blk;
        for (int kg = 0; kg < 2; ++kg)
          for (int e = 0; e < 8; ++e) {
            const int ci = q * 16 + kg * 8 + e;
            for (int n = 0; n < nt; ++n) {
              uint16_t hb, lb;
              split(row[s * nt + n][(size_t)ci * 9 + tap], hb, lb);
              b[((size_t)kg * nt + n) * 8 + e] = rank == 0 ? hb : lb;
            }
            for (int n = 0; n < nt / 2; ++n) {
              uint16_t hb, lb;
              split(row[s * nt + rank * (nt / 2) + n][(size_t)ci * 9 + tap], hb, lb);
              b[main_h + ((size_t)kg * (nt / 2) + n) * 8 + e] = hb;
            }
          }
      }
  return out;
}

// Upsample-fused layers (conv_tc2.cu, UP mode).  With U[i][j] = L[i>>1][j>>1] (UpsamplingNearest2d(2), refinenet.py:66,71,76)
// the 3x3 convolution of U at output pixel (2y+a, 2x+b) touches only the 2x2 low-resolution pixels
// L[y+a-1+ky][x+b-1+kx], ky,kx in {0,1}, with the kernel rows / columns that land on the same pixel summed:
//   a = 0: ky=0 <- row 0,  ky=1 <- rows 1+2;      a = 1: ky=0 <- rows 0+1,  ky=1 <- row 2      (same for b and columns)
// (zero padding of U at -1 / 2H is zero padding of L at -1 / H).  4 MACs per output instead of 9; the sums are formed in
// double and then split into fp16 hi/lo like every other weight (22 bits).
// Blocks: index ((((cslice*2 + a)*chunks + q)*2 + ky)*2 + kx)*2 + b, each in the pack_tc_pair format for `rank`.
std::vector<double> collapse_up_weights(const float* w, int cout, int cin) {
  // -> [a][b][ky][kx][o][ci]
  std::vector<double> c((size_t)16 * cout * cin, 0.0);
  const int lo[2][2] = {{0, 1}, {0, 2}}, hi[2][2] = {{0, 2}, {1, 2}};   // [phase][k] -> kernel index range [lo, hi]
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 2; ++kx)
          for (int o = 0; o < cout; ++o)
            for (int ci = 0; ci < cin; ++ci) {
              double acc = 0.0;
              for (int r = lo[a][ky]; r <= hi[a][ky]; ++r)
                for (int cc = lo[b][kx]; cc <= hi[b][kx]; ++cc) acc += (double)w[((size_t)o * cin + ci) * 9 + r * 3 + cc];
              c[((((size_t)(a * 2 + b) * 2 + ky) * 2 + kx) * cout + o) * cin + ci] = acc;
            }
  return c;
}

std::vector<uint16_t> pack_tc_pair_up(const std::vector<double>& comb, int cout, int cin, int nt, float scale, int rank) {
  const int cslices = cout / nt, chunks = cin / 16;
  const size_t main_h = (size_t)2 * nt * 8, x_h = (size_t)2 * (nt / 2) * 8, blk = main_h + x_h;
  std::vector<uint16_t> out((size_t)cslices * 2 * chunks * 8 * blk, 0);
  auto split = [&](double w, uint16_t& hb, uint16_t& lb) {
    const double ws_ = w * (double)scale;
    const __half hi = __float2half_rn((float)ws_);
    const __half lo = __float2half_rn((float)(ws_ - (double)__half2float(hi)));
    std::memcpy(&hb, &hi, 2); std::memcpy(&lb, &lo, 2);
  };
  for (int cs = 0; cs < cslices; ++cs)
    for (int a = 0; a < 2; ++a)
      for (int q = 0; q < chunks; ++q)
        for (int ky = 0; ky < 2; ++ky)
          for (int kx = 0; kx < 2; ++kx)
            for (int b = 0; b < 2; ++b) {
              uint16_t* blkp = out.data() + ((((((size_t)cs * 2 + a) * chunks + q) * 2 + ky) * 2 + kx) * 2 + b) * blk;
              const double* cw = comb.data() + (((size_t)(a * 2 + b) * 2 + ky) * 2 + kx) * cout * cin;
              for (int kg = 0; kg < 2; ++kg)
                for (int e = 0; e < 8; ++e) {
                  const int ci = q * 16 + kg * 8 + e;
                  for (int n = 0; n < nt; ++n) {
                    uint16_t hb, lb;
                    split(cw[(size_t)(cs * nt + n) * cin + ci], hb, lb);
                    blkp[((size_t)kg * nt + n) * 8 + e] = rank == 0 ? hb : lb;
                  }
                  for (int n = 0; n < nt / 2; ++n) {
                    uint16_t hb, lb;
                    split(cw[(size_t)(cs * nt + rank * (nt / 2) + n) * cin + ci], hb, lb);
                    blkp[main_h + ((size_t)kg * (nt / 2) + n) * 8 + e] = hb;
                  }
                }
            }
  return out;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace

struct TcGeom { int tr, tc; };
static TcGeom tc_geom(int nt, int hout, int wout) {
  TcGeom g;
  tc_tile_arrangement(nt, hout, wout, &g.tr, &g.tc);   // single source of truth: conv_tc.cu
  return g;
}

struct DcuEngine {
  DcuConfig cfg{};
  int sm_count = 148;
  int conv_impl = DCU_CONV_FFMA;
  bool has_ref = false;
  bool decode_only = false;     // DCU_FLAG_DECODE_ONLY: no networks, no conv workspace (pred_to_keypoints / extract_patches helpers)
  int64_t launches = 0;

  // detector
  FirstLayer det_first;
  Layer3x3 det[8];              // 1b,2a,2b,3a,3b,4a,4b,(Pa|Da)
  DevBuf w_loc, b_loc, w_ids, b_ids;
  // the same 1x1 heads for the tcgen05 kernel: weight blocks padded to NT rows, padded bias, dummy BN vectors
  struct TcHead { DevBuf w, bias, ones; float scale = 1.f; int nt = 0, copies = 1; } tc_loc, tc_ids;
  // refinenet
  FirstLayer ref_first;
  Layer3x3 ref[10];             // 1b,2a,2b,3a,3b,4a,4b,5a,5b,Pa
  DevBuf ref_head_w; float ref_head_b = 0.f;

  DevBuf lut;                   // [256] (x-128)/255
  int mb1 = 64, mb2 = 256, rp = 4096;   // micro-batch sizes: full-res layers, low-res layers, RefineNet patches
  int rp_plain = 1024;                  // RefineNet chunk when the upsampled tensors are materialised
  DevBuf act[2];                // ping-pong activation buffers
  DevBuf c1[2];                 // conv1a outputs, double-buffered: conv1a of micro-batch i+1 (HBM-write bound, side stream)
                                // overlaps conv1b/2a/2b of micro-batch i (tensor bound, caller's stream)
  // host entry point: the frames are copied in micro-batch sized chunks on `copy`, and the first layer of micro-batch i waits for
  // chunk i only, so all but the first chunk's transfer hides behind the previous micro-batch's kernels
  cudaStream_t copy = nullptr;
  cudaEvent_t ev_copy_start = nullptr;
  std::vector<cudaEvent_t> h2d_ev;
  bool h2d_active = false;
  int h2d_base = 0;                     // chunk index of the current group's first micro-batch
  cudaStream_t side = nullptr;
  cudaEvent_t ev_start = nullptr, ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  bool overlap_first = true;
  bool f32_in_range = true;     // fp32 image inputs are normalised ((x-128)/255, |x| <= 0.51): required by the fused kernel's bound
  bool fuse_first = false;      // detector: conv1a inside conv1b's kernel (DCU_FUSE_FIRST=1).  Off by default: bit-identical and 39 MB / frame less
                                // DRAM traffic, but the CUDA-core producers pace the kernel (tensor pipe 57 % instead of 84 % active) and the step
                                // time is the same within 1 %
  bool small_slices = true;     // DCU_SMALL_SLICES=0: never switch 128-channel layers to 64-channel slices for small launches
  bool chunked_h2d = true;      // DCU_CHUNKED_H2D=0: one copy of the whole batch on the caller's stream before the first kernel
  bool fuse_up = true;          // RefineNet: fold the 2x nearest upsamplings into the consuming convolution (DCU_FUSE_UP=0: materialise)
  bool flat = true;             // RefineNet maps up to conv4a's input as F2 runs (conv_tc2.cu FLAT mode; DCU_FLAT=0: per-patch tiles)
  DevBuf flat8[3];              // 8x8 maps in 9x9 cells (conv2b / conv3a / conv3b outputs); gutters stay zero
  bool tc_pair = true;          // use the CTA-pair (cta_group::2) kernel for the 3x3 layers (DCU_TC_PAIR=0: single-CTA kernel)
  DevBuf stage2_in;             // conv2b output for mb2 frames (input of conv3a)
  DevBuf heads;                 // (Pa|Da) output for mb2 frames
  DevBuf loc, ids;              // [mb2] logits when the caller does not want them
  DevBuf loc_arg, ids_arg;      // [mb2][cells] u8 arg-max maps written by the 1x1 head epilogues in the fused pipeline
  bool arg_heads = true;        // DCU_ARG_HEADS=0: heads write fp32 logits and the decode re-reads all 82 planes
  bool arg_heads_now = false;   // set around the fused pipeline's detector + decode
  DevBuf counts, offsets, total, kpts, patches, keys, refined, scan_state, frames;
  DevBuf resize_tab; int resize_hs = 0, resize_ws = 0;      // coefficient tables of the last dcu_resize_u8 source size
  DevBuf synth_params, synth_lat, synth_m;   // dcu_synth_frames / dcu_warp_perspective_u8 scratch (grown on demand)
  DevBuf pnp_obj;               // [n_obj][2] board corner table of the last solve_pnp geometry
  int pnp_cols = 0, pnp_rows = 0; double pnp_sq = 0.0;
  DevBuf bgr;                   // [max_batch][H][W][3] staging for the BGR entry point (allocated on first use)
  unsigned int epoch = 1;
  unsigned int epoch_override = 0;      // != 0 while a small-batch graph is captured (replays reset scan_state instead)
  // Small batches (one frame per call is the reference's own benchmark loop, src/benchmark.py:38-53): ~27 launches of a few
  // microseconds each are launch-bound, so the whole fixed sequence  H2D -> detector -> decode -> RefineNet on a fixed number of
  // patch slots -> D2H  is captured once per (n, dust_bin, use_refinenet) and replayed as ONE cudaGraphLaunch.
  struct SmallGraph { cudaGraphExec_t exec = nullptr; int n = 0, dust = 0, use_ref = 0, pfix = 0, channels = 1; int64_t launches = 0; bool dead = false; };
  std::vector<SmallGraph> graphs;
  cudaStream_t gstream = nullptr;       // capture stream
  bool use_graphs = true;               // DCU_GRAPH=0: always launch kernel by kernel
  int graph_max_n = 8;
  int graph_hint = 0;                   // recent corner count per call, picks the graph's number of patch slots
  // Sync-free detector -> RefineNet hand-off: RefineNet is enqueued right behind the decode for a number of
  // 4096-patch chunks predicted from recent calls; its kernels take the true patch count from device memory, so a chunk does exactly
  // the work that exists.  The host still learns the count (an event after the 4-byte copy, not a stream synchronisation: the GPU is
  // already running RefineNet) and launches the rare chunk the prediction missed.  DCU_DEVICE_COUNT=0: read the count back first.
  bool dev_count = true;
  int patch_hint = 0;
  cudaEvent_t ev_total = nullptr;
  // optional per-launch event timing (dcu_profile_*)
  struct ProfRec { cudaEvent_t a, b; double work; int cls; int shape[5]; double issued; };   // shape: cin, cout, hout, wout, n
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t ev = ev_pool.back(); ev_pool.pop_back(); return ev; }
    cudaEvent_t ev; cudaEventCreate(&ev); return ev;
  }
  void prof_begin(int cls, double work, cudaStream_t s, int cin = 0, int cout = 0, int hout = 0, int wout = 0, int n = 0) {
    if (!profiling) return;
    ProfRec r{get_event(), get_event(), work, cls, {cin, cout, hout, wout, n}, 0.0};
    cudaEventRecord(r.a, s);
    prof.push_back(r);
  }
  void prof_end(cudaStream_t s, double issued = 0.0) {
    if (!profiling) return;
    cudaEventRecord(prof.back().b, s);
    prof.back().issued = issued;
  }
  // pinned staging (host entry point)
  uint8_t* h_frames = nullptr; int32_t* h_counts = nullptr; int32_t* h_offsets = nullptr;
  int32_t* h_kpts = nullptr; float* h_refined = nullptr; int32_t* h_total = nullptr;

  ~DcuEngine() {
    if (side) cudaStreamDestroy(side);
    if (copy) cudaStreamDestroy(copy);
    if (ev_copy_start) cudaEventDestroy(ev_copy_start);
    for (auto ev : h2d_ev) cudaEventDestroy(ev);
    for (auto& g : graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (gstream) cudaStreamDestroy(gstream);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_total) cudaEventDestroy(ev_total);
    for (int i = 0; i < 2; ++i) { if (ev_done[i]) cudaEventDestroy(ev_done[i]); if (ev_free[i]) cudaEventDestroy(ev_free[i]); }
    DevBuf* all[] = {&resize_tab, &synth_params, &synth_lat, &synth_m, &loc_arg, &ids_arg, &pnp_obj, &flat8[0], &flat8[1], &flat8[2], &bgr, &c1[0], &c1[1], &tc_loc.w, &tc_loc.bias, &tc_loc.ones, &tc_ids.w, &tc_ids.bias, &tc_ids.ones, &w_loc, &b_loc, &w_ids, &b_ids, &ref_head_w, &lut, &act[0], &act[1], &stage2_in, &heads, &loc,
                     &ids, &counts, &offsets, &total, &kpts, &patches, &keys, &refined, &scan_state, &frames};
    for (DevBuf* b : all) b->release();
    FirstLayer* fl[] = {&det_first, &ref_first};
    for (FirstLayer* f : fl) { f->w.release(); f->bias.release(); f->alpha.release(); f->beta.release(); }
    for (Layer3x3& l : det) { l.w_tc2_64[0].release(); l.w_tc2_64[1].release(); l.w_up_64[0].release(); l.w_up_64[1].release(); l.w_tc2[0].release(); l.w_tc2[1].release(); l.w_ffma.release(); l.w_tc.release(); l.bias.release(); l.alpha.release(); l.beta.release(); }
    for (Layer3x3& l : ref) { l.w_tc2_64[0].release(); l.w_tc2_64[1].release(); l.w_up_64[0].release(); l.w_up_64[1].release(); l.w_up[0].release(); l.w_up[1].release(); l.w_tc2[0].release(); l.w_tc2[1].release(); l.w_ffma.release(); l.w_tc.release(); l.bias.release(); l.alpha.release(); l.beta.release(); }
    for (auto& r : prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto ev : ev_pool) cudaEventDestroy(ev);
    if (h_frames) cudaFreeHost(h_frames);
    if (h_counts) cudaFreeHost(h_counts);
    if (h_offsets) cudaFreeHost(h_offsets);
    if (h_kpts) cudaFreeHost(h_kpts);
    if (h_refined) cudaFreeHost(h_refined);
    if (h_total) cudaFreeHost(h_total);
  }
};

namespace {

int build_first(FirstLayer& f, const DcuConvLayer& L, int pad) {
  if (L.cin != 1 || L.cout != 64 || L.ksize != 3 || !L.alpha || !L.beta)
    return fail(DCU_ERR_INVALID, "first layer must be 1->64 3x3 with BN");
  std::vector<float> w(9 * 64);
  for (int o = 0; o < 64; ++o)
    for (int t = 0; t < 9; ++t) w[t * 64 + o] = L.weight[o * 9 + t];
  f.pad = pad;
  std::memcpy(f.host.w, w.data(), sizeof(f.host.w));
  std::memcpy(f.host.bias, L.bias, 256); std::memcpy(f.host.alpha, L.alpha, 256); std::memcpy(f.host.beta, L.beta, 256);
  f.out_bound = 0.f;
  for (int o = 0; o < 64; ++o) {
    double sw = 0;
    for (int t = 0; t < 9; ++t) sw += std::fabs((double)L.weight[o * 9 + t]);
    const double b = (0.51 * sw + std::fabs((double)L.bias[o])) * std::fabs((double)L.alpha[o]) + std::fabs((double)L.beta[o]);
    f.out_bound = std::max(f.out_bound, (float)b);
  }
  CK(upload(f.w, w));
  CK(upload(f.bias, std::vector<float>(L.bias, L.bias + 64)));
  CK(upload(f.alpha, std::vector<float>(L.alpha, L.alpha + 64)));
  CK(upload(f.beta, std::vector<float>(L.beta, L.beta + 64)));
  return DCU_OK;
}

int encode_weight_map(CUtensorMap* tm, void* base, size_t bytes, int nt, int up) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int rows_per_block = tc2_block_bytes(nt) / 512;
  cuuint64_t dims[2] = {256, (cuuint64_t)(bytes / 512)};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {256, (cuuint32_t)(tc2_stage_blocks(up) * rows_per_block)};
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed: " + std::to_string((int)cr));
  return DCU_OK;
}

int build_3x3(Layer3x3& l, std::vector<const DcuConvLayer*> parts, int pad, int pool, int ups, bool ups_in = false) {
  std::vector<const float*> ws, bs, as, es;
  std::vector<int> couts;
  int cin = parts[0]->cin;
  for (const DcuConvLayer* L : parts) {
    if (L->ksize != 3 || L->cin != cin || (cin % 16) != 0 || (L->cout % 64) != 0 || !L->alpha || !L->beta)
      return fail(DCU_ERR_INVALID, "unsupported 3x3 layer shape");
    ws.push_back(L->weight); bs.push_back(L->bias); as.push_back(L->alpha); es.push_back(L->beta);
    couts.push_back(L->cout);
  }
  l.cin = cin; l.cout = 0;
  for (int c : couts) l.cout += c;
  l.pad = pad; l.pool = pool; l.ups = ups;
  CK(upload(l.w_ffma, pack_ffma(ws, couts, cin)));
  CK(upload(l.bias, concat(bs, couts)));
  CK(upload(l.alpha, concat(as, couts)));
  CK(upload(l.beta, concat(es, couts)));
  if (l.cout > 512) return fail(DCU_ERR_INVALID, "3x3 layer with more than 512 output channels");
  {
    const std::vector<float> hb = concat(bs, couts), ha = concat(as, couts), he = concat(es, couts);
    std::memcpy(l.host_bn.v[0], hb.data(), hb.size() * 4);
    std::memcpy(l.host_bn.v[1], ha.data(), ha.size() * 4);
    std::memcpy(l.host_bn.v[2], he.data(), he.size() * 4);
  }
  l.tc_nt = tc_supported_shape(cin, l.cout);
  if (l.tc_nt > 0) {
    l.tc_scale = tc_weight_scale(ws, couts, cin);
    const std::vector<uint16_t> blocks = pack_tc(ws, couts, cin, l.tc_nt, l.tc_scale);
    l.tc_copies = 8;
    if (const char* v = getenv("DCU_W_COPIES")) l.tc_copies = std::max(1, atoi(v));
    CK(l.w_tc.alloc(blocks.size() * 2 * l.tc_copies));
    for (int c = 0; c < l.tc_copies; ++c)
      CK(cudaMemcpy(l.w_tc.as<uint8_t>() + (size_t)c * blocks.size() * 2, blocks.data(), blocks.size() * 2, cudaMemcpyHostToDevice));
    int rc;
    for (int r = 0; r < 2; ++r) {
      const std::vector<uint16_t> pb = pack_tc_pair(ws, couts, cin, l.tc_nt, l.tc_scale, r);
      CK(l.w_tc2[r].alloc(pb.size() * 2));
      CK(cudaMemcpy(l.w_tc2[r].p, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
      if ((rc = encode_weight_map(&l.wmap2[r], l.w_tc2[r].p, pb.size() * 2, l.tc_nt, 0))) return rc;
    }
    if (ups_in && parts.size() == 1 && pad == 1 && !pool && !ups) {
      const std::vector<double> comb = collapse_up_weights(ws[0], l.cout, cin);
      double mx = 0.0;
      for (double v : comb) mx = std::max(mx, std::fabs(v));
      l.up_scale = mx > 0.0 ? (float)std::exp2(std::floor(std::log2(32768.0 / mx))) : 1.f;
      for (int r = 0; r < 2; ++r) {
        const std::vector<uint16_t> pb = pack_tc_pair_up(comb, l.cout, cin, l.tc_nt, l.up_scale, r);
        CK(l.w_up[r].alloc(pb.size() * 2));
        CK(cudaMemcpy(l.w_up[r].p, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
        if ((rc = encode_weight_map(&l.wmap_up[r], l.w_up[r].p, pb.size() * 2, l.tc_nt, 1))) return rc;
      }
      l.ups_in = true;
    }
    if (l.tc_nt == 128) {
      for (int r = 0; r < 2; ++r) {
        const std::vector<uint16_t> pb = pack_tc_pair(ws, couts, cin, 64, l.tc_scale, r);
        CK(l.w_tc2_64[r].alloc(pb.size() * 2));
        CK(cudaMemcpy(l.w_tc2_64[r].p, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
        if ((rc = encode_weight_map(&l.wmap2_64[r], l.w_tc2_64[r].p, pb.size() * 2, 64, 0))) return rc;
      }
      if (l.ups_in) {
        const std::vector<double> comb = collapse_up_weights(ws[0], l.cout, cin);
        for (int r = 0; r < 2; ++r) {
          const std::vector<uint16_t> pb = pack_tc_pair_up(comb, l.cout, cin, 64, l.up_scale, r);
          CK(l.w_up_64[r].alloc(pb.size() * 2));
          CK(cudaMemcpy(l.w_up_64[r].p, pb.data(), pb.size() * 2, cudaMemcpyHostToDevice));
          if ((rc = encode_weight_map(&l.wmap_up_64[r], l.w_up_64[r].p, pb.size() * 2, 64, 1))) return rc;
        }
      }
      l.has_64 = true;
    }
  }
  return DCU_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// layer runners
// ---------------------------------------------------------------------------------------------------
static int make_tmap(CUtensorMap* tm, const void* base, int n, int cin, int h, int w, int box_w, int box_h) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // H2 activations: __half [n][hi|lo][cin/8][h][w][8].  Innermost dimension = a whole image row of 16-byte pixels (w*8
  // halves) so the box row is halo_w*16 contiguous bytes; box = {8*halo_w, halo_h, 2 k-groups (16 channels), hi|lo, 1}.
  cuuint64_t dims[5] = {(cuuint64_t)w * 8, (cuuint64_t)h, (cuuint64_t)(cin / 8), 2, (cuuint64_t)n};
  const cuuint64_t plane = (cuuint64_t)w * h * 16;
  cuuint64_t strides[4] = {(cuuint64_t)w * 16, plane, plane * (cin / 8), plane * (cin / 8) * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, 2, 2, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return DCU_OK;
}

// F2 tensor (common.cuh) as a 4-D tensor map {16 pixels x 8 halves, rows of 16 pixels, C/8, hi|lo}; box = {128, box_rows, 2, 2}.
// `rows` bounds the pixels that hold data (reads beyond are zero-filled); plane_px is the allocated run length per channel group.
static int make_tmap_flat(CUtensorMap* tm, const void* base, long long px_used, int cin, long long plane_px, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  if (plane_px % 16 != 0 || px_used > plane_px) return fail(DCU_ERR_INVALID, "flat tensor map: bad plane size");
  cuuint64_t dims[4] = {128, (cuuint64_t)((px_used + 15) / 16), (cuuint64_t)(cin / 8), 2};
  cuuint64_t strides[3] = {256, (cuuint64_t)plane_px * 16, (cuuint64_t)plane_px * 16 * (cin / 8)};
  cuuint32_t box[4] = {128, (cuuint32_t)box_rows, 2, 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DCU_ERR_CUDA, "cuTensorMapEncodeTiled (flat) failed: " + std::to_string((int)r));
  return DCU_OK;
}

struct FirstIn { const uint8_t* u8; const float* f32; const FirstWeights* w; };   // conv1a fused into this layer (conv_tc2.cu FIRST mode)
struct FlatIn { int period, row; long long plane_px; };    // the layer's input is an F2 tensor (conv_tc2.cu FLAT mode)
static long long flat_plane_px(int n, int period) { return (((long long)n * period + 15) / 16) * 16; }

struct HeadFuse { const float* w = nullptr; float b = 0.f; unsigned long long* keys = nullptr; float* heat = nullptr; };
struct DevCount { const int* p; int off; };      // image count in device memory: the launch processes clamp(*p - off, 0, n) images
static unsigned long long* g_tc_stats = nullptr;   // device [8]; set by dcu_debug_tc_stats (profiling only)

// hin x win: the layer's input size as the reference sees it.  fuse_up (tcgen05 pair kernel only): the upsampling between a
// producer (l.ups) and its consumer (l.ups_in) is not materialised -- the producer stores its low-resolution output and the
// consumer runs the phase-collapsed 2x2 kernels on it (`in` is then the hin/2 x win/2 tensor).
static int run_3x3(DcuEngine* e, const Layer3x3& l, int impl, const float* in, float* out, int n, int hin, int win,
                   const HeadFuse* hf, cudaStream_t s, bool fuse_up = false, const FlatIn* fin = nullptr,
                   const H2Layout* lout = nullptr, const FirstIn* first = nullptr, const DevCount* dcnt = nullptr) {
  const bool up_in = fuse_up && l.ups_in;
  if ((fin || lout || first || dcnt) && !(impl == DCU_CONV_TCGEN05 && e->tc_pair)) return fail(DCU_ERR_INVALID, "flat layouts / conv1a fusion / device-side counts need the CTA-pair kernel");
  ConvParams p{};
  p.in = in; p.out = out; p.bias = l.bias.as<float>(); p.alpha = l.alpha.as<float>(); p.beta = l.beta.as<float>();
  p.n = n; p.cin = l.cin; p.cout_total = l.cout; p.hin = up_in ? hin / 2 : hin; p.win = up_in ? win / 2 : win;
  p.hout = hin + 2 * l.pad - 2; p.wout = win + 2 * l.pad - 2; p.pad = l.pad; p.pool = l.pool; p.ups = fuse_up ? 0 : l.ups;
  if (hf) { p.head_w = hf->w; p.head_b = hf->b; p.head_key = hf->keys; p.heat = hf->heat; }
  p.stats = g_tc_stats;
  p.wscale_inv = (impl == DCU_CONV_TCGEN05) ? 1.0f / (up_in ? l.up_scale : l.tc_scale) : 1.0f;
  p.ksize = 3;
  p.host_bn = &l.host_bn;
  if (fin) { p.flat_in = 1; p.in_period = fin->period; p.in_row = fin->row; }
  if (lout) p.out_layout = *lout;
  if (dcnt) { p.n_dev = dcnt->p; p.n_off = dcnt->off; }
  if (first) { p.first_u8 = first->u8; p.first_f32 = first->f32; p.first_w = first->w; p.in = e->act[0].as<float>(); in = p.in; }
  if (n <= 0) return DCU_OK;
  e->prof_begin(0, 2.0 * 9.0 * (l.cin + (first ? 1 : 0)) * l.cout * (double)p.hout * p.wout * n, s, l.cin, l.cout, p.hout, p.wout, n);
  double issued = 0.0;
  if (impl == DCU_CONV_TCGEN05) {
    if (l.tc_nt == 0) return fail(DCU_ERR_UNSUPPORTED, "layer shape not supported by the tcgen05 kernel");
    TcGeom g = tc_geom(l.tc_nt, p.hout, p.wout);
    // few work items (small batch): 64-channel slices -> twice the items, half the MMA chain per item; still few: one m-tile per CTA
    bool use64 = false;
    if (e->tc_pair && e->small_slices && (l.has_64 || l.tc_nt == 64)) {
      const long long px = fin ? (long long)n * fin->period : (long long)n * (up_in ? p.hin * p.win * 2 : p.hout * p.wout);
      const long long items128 = ((px + 511) / 512) * ((l.cout + 127) / 128) * (up_in ? 2 : 1);
      use64 = l.has_64 && items128 < e->sm_count / 2;
      const long long items64 = ((px + 511) / 512) * (l.cout / 64) * (up_in ? 2 : 1);
      if ((use64 || l.tc_nt == 64) && !up_in && !fin && !first && items64 < e->sm_count / 2) { p.mt1 = 1; g.tr = 1; g.tc = 1; }
    }
    // DCU_NT64=3 (default): a 128-channel layer runs as 128-channel slices on ONE m-tile per CTA (two accumulator sets, no
    // fetch-bound MMA; conv_tc.cu tc_supported_shape) where the kernel has that form, in 64-channel slices otherwise
    static const bool nt128_mt1 = [] { const char* v = getenv("DCU_NT64"); return !v || atoi(v) == 3; }();
    if (nt128_mt1 && e->tc_pair && l.tc_nt == 128 && !use64) {
      if (!up_in && !first) { p.mt1 = 1; g.tr = 1; g.tc = 1; }
      else if (l.has_64) use64 = true;
    }
    CUtensorMap tm;
    int rc = fin ? make_tmap_flat(&tm, in, (long long)n * fin->period, l.cin, fin->plane_px,
                                  tc2_flat_rows(fin->row, (up_in || l.pad) ? 1 : 0, (up_in || p.mt1) ? 1 : 0))
             : up_in ? make_tmap(&tm, in, n, l.cin, p.hin, p.win, 10, 18)
                     : make_tmap(&tm, in, n, l.cin, hin, win, 8 * g.tc + 2, 16 * g.tr + 2);
    if (rc) return rc;
    if (fuse_up && !e->tc_pair) return fail(DCU_ERR_INVALID, "upsample fusion needs the CTA-pair kernel");
    const int nt_use = use64 ? 64 : l.tc_nt;
    cudaError_t ce = e->tc_pair
                         ? (up_in ? launch_conv_tc2(p, l.cout / nt_use, 1, &tm, use64 ? &l.wmap_up_64[0] : &l.wmap_up[0],
                                                    use64 ? &l.wmap_up_64[1] : &l.wmap_up[1], e->sm_count, s, &issued)
                                  : launch_conv_tc2(p, l.cout / nt_use, 0, &tm, use64 ? &l.wmap2_64[0] : &l.wmap2[0],
                                                    use64 ? &l.wmap2_64[1] : &l.wmap2[1], e->sm_count, s, &issued))
                         : launch_conv3x3_tc(p, l.w_tc.as<float>(), l.cout / l.tc_nt, l.tc_copies, &tm, e->sm_count, s);
    if (ce != cudaSuccess) return fail(DCU_ERR_CUDA, std::string("tcgen05 conv launch: ") + cudaGetErrorString(ce));
  } else {
    launch_conv3x3_ffma(p, l.w_ffma.as<float>(), s);
  }
  e->prof_end(s, issued);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

static int run_first(DcuEngine* e, const FirstLayer& f, const uint8_t* in_u8, const float* in_f32, float* out, int n,
                     int hin, int win, int out_h2, cudaStream_t s, const H2Layout* lout = nullptr, const DevCount* dcnt = nullptr) {
  FirstConvParams p{};
  p.out_h2 = out_h2;
  if (lout) p.out_layout = *lout;
  if (dcnt) { p.n_dev = dcnt->p; p.n_off = dcnt->off; }
  p.in_u8 = in_u8; p.in_f32 = in_f32; p.lut = e->lut.as<float>(); p.out = out; p.w = f.w.as<float>();
  p.bias = f.bias.as<float>(); p.alpha = f.alpha.as<float>(); p.beta = f.beta.as<float>();
  p.n = n; p.hin = hin; p.win = win; p.pad = f.pad; p.hout = hin + 2 * f.pad - 2; p.wout = win + 2 * f.pad - 2;
  if (n <= 0) return DCU_OK;
  e->prof_begin(1, 2.0 * 9.0 * 64 * (double)p.hout * p.wout * n, s, 1, 64, p.hout, p.wout, n);
  launch_conv_first(p, s);
  e->prof_end(s);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

// detector on n <= mb2 frames: loc/ids NCHW out
static int detector_group(DcuEngine* e, const uint8_t* frames, const float* images, int n, float* loc, float* ids,
                          cudaStream_t s) {
  NvtxRange nvtx_range("dcu:detector");
  const int H = e->cfg.height, W = e->cfg.width;
  float* a0 = e->act[0].as<float>();
  float* a1 = e->act[1].as<float>();
  float* s2 = e->stage2_in.as<float>();
  const size_t s2_frame = (size_t)64 * (H / 4) * (W / 4);
  int rc;
  // conv1a is HBM-write bound (19.7 MB per frame), conv1b/2a/2b are tensor bound: run conv1a of micro-batch i+1 on a side
  // stream while the tensor-core kernels of micro-batch i run on the caller's stream (double-buffered conv1a output).
  const bool h2 = e->conv_impl == DCU_CONV_TCGEN05;
  const int n_mb = (n + e->mb1 - 1) / e->mb1;
  // conv1a computed inside conv1b's kernel (pair kernel only): no conv1a launch, no 19.7 MB / frame round trip through HBM
  const bool fused_first = e->fuse_first && h2 && e->tc_pair && e->det[0].tc_nt == 64 && e->det_first.out_bound < 60000.f &&
                           (frames != nullptr || e->f32_in_range);
  const bool overlap = e->overlap_first && n_mb > 1 && !e->profiling && !fused_first;
  auto first = [&](int i, cudaStream_t st) -> int {
    const int f0 = i * e->mb1, m = std::min(e->mb1, n - f0);
    if (e->h2d_active && cudaStreamWaitEvent(st, e->h2d_ev[e->h2d_base + i], 0) != cudaSuccess)
      return fail(DCU_ERR_CUDA, "cudaStreamWaitEvent (chunked H2D) failed");
    return run_first(e, e->det_first, frames ? frames + (size_t)f0 * H * W : nullptr,
                     images ? images + (size_t)f0 * H * W : nullptr, e->c1[i & 1].as<float>(), m, H, W, h2, st);
  };
  if (overlap) {
    CK(cudaEventRecord(e->ev_start, s));
    CK(cudaStreamWaitEvent(e->side, e->ev_start, 0));        // inputs (e.g. the H2D copy) are ordered on the caller's stream
    if ((rc = first(0, e->side))) return rc;
    CK(cudaEventRecord(e->ev_done[0], e->side));
  }
  for (int i = 0; i < n_mb; ++i) {
    const int f0 = i * e->mb1, m = std::min(e->mb1, n - f0);
    float* c1 = e->c1[i & 1].as<float>();
    if (overlap) {
      if (i + 1 < n_mb) {
        if (i >= 1) CK(cudaStreamWaitEvent(e->side, e->ev_free[(i + 1) & 1], 0));   // conv1b of micro-batch i-1 has read that buffer
        if ((rc = first(i + 1, e->side))) return rc;
        CK(cudaEventRecord(e->ev_done[(i + 1) & 1], e->side));
      }
      CK(cudaStreamWaitEvent(s, e->ev_done[i & 1], 0));
    } else if (!fused_first) {
      if ((rc = first(i, s))) return rc;                                                                         // conv1a
    }
    if (fused_first) {
      if (e->h2d_active) CK(cudaStreamWaitEvent(s, e->h2d_ev[e->h2d_base + i], 0));
      const FirstIn fi{frames ? frames + (size_t)f0 * H * W : nullptr, images ? images + (size_t)f0 * H * W : nullptr, &e->det_first.host};
      if ((rc = run_3x3(e, e->det[0], e->conv_impl, nullptr, a1, m, H, W, nullptr, s, false, nullptr, nullptr, &fi))) return rc;   // conv1a + conv1b + pool
    } else
    if ((rc = run_3x3(e, e->det[0], e->conv_impl, c1, a1, m, H, W, nullptr, s))) return rc;                     // conv1b + pool
    if (overlap) CK(cudaEventRecord(e->ev_free[i & 1], s));
    if ((rc = run_3x3(e, e->det[1], e->conv_impl, a1, a0, m, H / 2, W / 2, nullptr, s))) return rc;             // conv2a
    if ((rc = run_3x3(e, e->det[2], e->conv_impl, a0, s2 + f0 * s2_frame, m, H / 2, W / 2, nullptr, s))) return rc;  // conv2b + pool
  }
  if ((rc = run_3x3(e, e->det[3], e->conv_impl, s2, a0, n, H / 4, W / 4, nullptr, s))) return rc;               // conv3a
  if ((rc = run_3x3(e, e->det[4], e->conv_impl, a0, a1, n, H / 4, W / 4, nullptr, s))) return rc;               // conv3b + pool
  if ((rc = run_3x3(e, e->det[5], e->conv_impl, a1, a0, n, H / 8, W / 8, nullptr, s))) return rc;               // conv4a
  if ((rc = run_3x3(e, e->det[6], e->conv_impl, a0, a1, n, H / 8, W / 8, nullptr, s))) return rc;               // conv4b
  if ((rc = run_3x3(e, e->det[7], e->conv_impl, a1, e->heads.as<float>(), n, H / 8, W / 8, nullptr, s))) return rc;  // convPa | convDa
  if (e->conv_impl == DCU_CONV_TCGEN05) {
    // convPb / convDb on the tensor cores: 1x1 "logits" mode of the same kernel, reading cPa / cDa from the H2 tensor
    for (int which = 0; which < 2; ++which) {
      const DcuEngine::TcHead& hd = which ? e->tc_ids : e->tc_loc;
      ConvParams p{};
      p.in = e->heads.as<float>(); p.bias = hd.bias.as<float>(); p.alpha = hd.ones.as<float>(); p.beta = hd.ones.as<float>();
      p.n = n; p.cin = 256; p.cout_total = hd.nt; p.hin = H / 8; p.win = W / 8; p.hout = H / 8; p.wout = W / 8;
      p.ksize = 1; p.cin_offset = which ? 256 : 0; p.logits = which ? ids : loc; p.n_valid = which ? e->cfg.n_ids + 1 : 65;
      if (e->arg_heads_now) {       // fused pipeline: arg-max in the epilogue, logits never written (loc / ids are null then)
        p.logits = nullptr;
        p.arg_out = (which ? e->ids_arg : e->loc_arg).as<uint8_t>();
      }
      p.wscale_inv = 1.0f / hd.scale; p.stats = nullptr;
      int tr, tc;
      tc_tile_arrangement(hd.nt, p.hout, p.wout, &tr, &tc);
      CUtensorMap tm;
      if ((rc = make_tmap(&tm, e->heads.p, n, 512, H / 8, W / 8, 8 * tc + 2, 16 * tr + 2))) return rc;
      e->prof_begin(2, 2.0 * 256.0 * p.n_valid * (double)p.hout * p.wout * n, s);
      cudaError_t ce = launch_conv3x3_tc(p, hd.w.as<float>(), 1, hd.copies, &tm, e->sm_count, s);
      e->prof_end(s);
      if (ce != cudaSuccess) return fail(DCU_ERR_CUDA, std::string("tcgen05 1x1 head launch: ") + cudaGetErrorString(ce));
      e->launches++;
    }
    CK(cudaGetLastError());
    return DCU_OK;
  }
  HeadParams hp{};
  hp.in = e->heads.as<float>(); hp.w_loc = e->w_loc.as<float>(); hp.b_loc = e->b_loc.as<float>();
  hp.w_ids = e->w_ids.as<float>(); hp.b_ids = e->b_ids.as<float>(); hp.loc = loc; hp.ids = ids;
  hp.n = n; hp.h = H / 8; hp.w = W / 8; hp.n_ids1 = e->cfg.n_ids + 1;
  hp.in_h2 = 0;
  e->prof_begin(2, 2.0 * 256.0 * (65 + hp.n_ids1) * (double)hp.h * hp.w * n, s);
  launch_heads_1x1(hp, s);                                                                                       // convPb, convDb
  e->prof_end(s);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

static int decode_group(DcuEngine* e, const float* loc, const float* ids, const uint8_t* frames, int n, int dust_bin,
                        int append, int32_t* counts, int32_t* offsets, int32_t* total, int32_t* kpts, float* patches,
                        cudaStream_t s) {
  NvtxRange nvtx_range("dcu:decode_gather");
  DecodeParams d{};
  d.loc = loc; d.ids = ids; d.frames = frames; d.lut = e->lut.as<float>();
  if (e->arg_heads_now) { d.loc_arg = e->loc_arg.as<uint8_t>(); d.ids_arg = e->ids_arg.as<uint8_t>(); }
  d.n = n; d.H = e->cfg.height; d.W = e->cfg.width; d.h = d.H / 8; d.w = d.W / 8; d.n_ids1 = e->cfg.n_ids + 1;
  d.dust_bin = dust_bin; d.append = append; d.counts = counts; d.offsets = offsets; d.total = total; d.kpts = kpts;
  d.patches = patches; d.max_patches = e->cfg.max_patches; d.scan_state = e->scan_state.as<unsigned long long>();
  d.epoch = e->epoch_override ? e->epoch_override : e->epoch++;
  // algorithmic bytes: logits read once; + K*(2304 read + 2304 written + 16) added by the caller's K
  // with the arg-max taken in the head epilogues the kernel that is left is still reported against these bytes
  e->prof_begin(3, (double)(65 + d.n_ids1) * d.h * d.w * 4.0 * n, s);
  launch_decode_gather(d, s);
  e->prof_end(s);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

// RefineNet on p patches (any p; processed in chunks of rp)
// total_dev != nullptr (tcgen05 pair kernel with the flat / upsample-fused layouts only): p is an upper bound chosen by the host and every
// kernel takes the true patch count from device memory -- no host round trip between the decode and RefineNet
static int refine_run(DcuEngine* e, const float* patches, const int32_t* xy, int xy_stride, int p, int32_t* corners,
                      float* refined, float* heat, cudaStream_t s, const int32_t* total_dev = nullptr, int p_begin = 0) {
  NvtxRange nvtx_range("dcu:refinenet");
  float* a0 = e->act[0].as<float>();
  float* a1 = e->act[1].as<float>();
  int rc;
  // tcgen05 pair kernel: the three 2x upsamplings are never materialised (run_3x3: fuse_up)
  const bool fu = e->fuse_up && e->tc_pair && e->conv_impl == DCU_CONV_TCGEN05 && e->ref[5].ups_in && e->ref[7].ups_in &&
                  e->ref[9].ups_in;
  const int chunk = fu ? e->rp : e->rp_plain;
  if (total_dev && !(fu && e->flat && e->flat8[0].p)) return fail(DCU_ERR_INVALID, "device-side patch count needs the flat / upsample-fused RefineNet path");
  for (int p0 = p_begin; p0 < p; p0 += chunk) {
    const int m = std::min(chunk, p - p0);
    const DevCount dcv{total_dev, p0};
    const DevCount* dc = total_dev ? &dcv : nullptr;
    unsigned long long* keys = e->keys.as<unsigned long long>() + p0;
    CK(cudaMemsetAsync(keys, 0, (size_t)m * sizeof(unsigned long long), s));
    if (fu && e->flat && e->flat8[0].p) {
      // small maps as F2 runs (conv_tc2.cu FLAT mode): 22x22 and 20x20 dense, 8x8 in 9x9 cells with zero gutters (flat8[],
      // zeroed once at creation; only data positions are ever written)
      const long long pl22 = flat_plane_px(m, 484), pl20 = flat_plane_px(m, 400), pl9 = flat_plane_px(e->rp, 81);
      const H2Layout l22 = h2_flat(64, 484, 22, pl22), l20 = h2_flat(64, 400, 20, pl20), l9 = h2_flat(128, 81, 9, pl9);
      const FlatIn f22{484, 22, pl22}, f20{400, 20, pl20}, f9{81, 9, pl9};
      float* g0 = e->flat8[0].as<float>(); float* g1 = e->flat8[1].as<float>(); float* g2 = e->flat8[2].as<float>();
      if ((rc = run_first(e, e->ref_first, nullptr, patches + (size_t)p0 * 576, a0, m, 24, 24, 1, s, &l22, dc))) return rc;  // conv1a -> 22
      if ((rc = run_3x3(e, e->ref[0], e->conv_impl, a0, a1, m, 22, 22, nullptr, s, false, &f22, &l20, nullptr, dc))) return rc;     // conv1b -> 20
      if ((rc = run_3x3(e, e->ref[1], e->conv_impl, a1, a0, m, 20, 20, nullptr, s, false, &f20, nullptr, nullptr, dc))) return rc;  // conv2a -> 18
      if ((rc = run_3x3(e, e->ref[2], e->conv_impl, a0, g0, m, 18, 18, nullptr, s, false, nullptr, &l9, nullptr, dc))) return rc;   // conv2b -> 16 -> pool 8
      if ((rc = run_3x3(e, e->ref[3], e->conv_impl, g0, g1, m, 8, 8, nullptr, s, false, &f9, &l9, nullptr, dc))) return rc;         // conv3a
      if ((rc = run_3x3(e, e->ref[4], e->conv_impl, g1, g2, m, 8, 8, nullptr, s, true, &f9, &l9, nullptr, dc))) return rc;          // conv3b (-> up 16)
      if ((rc = run_3x3(e, e->ref[5], e->conv_impl, g2, a0, m, 16, 16, nullptr, s, true, &f9, nullptr, nullptr, dc))) return rc;    // conv4a
    } else {
    if ((rc = run_first(e, e->ref_first, nullptr, patches + (size_t)p0 * 576, a0, m, 24, 24, e->conv_impl == DCU_CONV_TCGEN05, s)))
      return rc;                                                                                  // conv1a -> 22
    if ((rc = run_3x3(e, e->ref[0], e->conv_impl, a0, a1, m, 22, 22, nullptr, s))) return rc;   // conv1b -> 20
    if ((rc = run_3x3(e, e->ref[1], e->conv_impl, a1, a0, m, 20, 20, nullptr, s))) return rc;   // conv2a -> 18
    if ((rc = run_3x3(e, e->ref[2], e->conv_impl, a0, a1, m, 18, 18, nullptr, s))) return rc;   // conv2b -> 16 -> pool 8
    if ((rc = run_3x3(e, e->ref[3], e->conv_impl, a1, a0, m, 8, 8, nullptr, s))) return rc;         // conv3a
    if ((rc = run_3x3(e, e->ref[4], e->conv_impl, a0, a1, m, 8, 8, nullptr, s, fu))) return rc;     // conv3b -> up 16
    if ((rc = run_3x3(e, e->ref[5], e->conv_impl, a1, a0, m, 16, 16, nullptr, s, fu))) return rc;   // conv4a
    }
    if ((rc = run_3x3(e, e->ref[6], e->conv_impl, a0, a1, m, 16, 16, nullptr, s, fu, nullptr, nullptr, nullptr, dc))) return rc;   // conv4b -> up 32
    if ((rc = run_3x3(e, e->ref[7], e->conv_impl, a1, a0, m, 32, 32, nullptr, s, fu, nullptr, nullptr, nullptr, dc))) return rc;   // conv5a
    if ((rc = run_3x3(e, e->ref[8], e->conv_impl, a0, a1, m, 32, 32, nullptr, s, fu, nullptr, nullptr, nullptr, dc))) return rc;   // conv5b -> up 64
    HeadFuse hf;
    hf.w = e->ref_head_w.as<float>(); hf.b = e->ref_head_b; hf.keys = keys;
    hf.heat = heat ? heat + (size_t)p0 * 4096 : nullptr;
    if ((rc = run_3x3(e, e->ref[9], e->conv_impl, a1, nullptr, m, 64, 64, &hf, s, fu, nullptr, nullptr, nullptr, dc))) return rc;  // convPa + convPb + arg-max
    e->prof_begin(4, 0.0, s);
    launch_refine_finalize(keys, xy + (size_t)p0 * xy_stride, xy_stride, m, corners ? corners + 2 * (size_t)p0 : nullptr,
                           refined + 2 * (size_t)p0, s, total_dev, p0);
    e->prof_end(s);
    e->launches++;
    CK(cudaGetLastError());
  }
  return DCU_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

const char* dcu_last_error(void) { return g_err.c_str(); }
const char* dcu_version(void) { return "deepcharuco_b200 0.1 (sm_100a)"; }

int dcu_create(const DcuConfig* cfg, const DcuConvLayer* D, int n_det, const DcuConvLayer* R, int n_ref,
               DcuEngine** out) {
  if (!cfg || !out) return fail(DCU_ERR_INVALID, "dcu_create: null argument");
  const bool decode_only = (cfg->reserved & DCU_FLAG_DECODE_ONLY) != 0;
  if (!decode_only && (!D || n_det != 12)) return fail(DCU_ERR_INVALID, "dcu_create: need 12 detector layers");
  if (cfg->height % 8 || cfg->width % 8 || cfg->height < 24 || cfg->width < 24)
    return fail(DCU_ERR_INVALID, "dcu_create: height/width must be multiples of 8 (>= 24)");
  if (cfg->n_ids < 1 || cfg->n_ids > 30 || cfg->max_batch < 1 || cfg->max_patches < 1)
    return fail(DCU_ERR_INVALID, "dcu_create: bad n_ids / max_batch / max_patches");
  if (cfg->width > 65535 || cfg->height > 65535) return fail(DCU_ERR_INVALID, "dcu_create: frame too large");
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(DCU_ERR_UNSUPPORTED, "deepcharuco_b200 needs an sm_100 (B200) device");
  DcuEngine* e = new DcuEngine();
  e->cfg = *cfg;
  e->sm_count = prop.multiProcessorCount;
  if (const char* v = getenv("DCU_SMS")) e->sm_count = std::max(2, std::min(atoi(v), prop.multiProcessorCount));   // experiments
  e->conv_impl = cfg->conv_impl;
  e->decode_only = decode_only;
  int rc;
#define TRY(x) do { if ((rc = (x)) != DCU_OK) { delete e; return rc; } } while (0)
#define TRYC(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { delete e; return fail(DCU_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } } while (0)
  if (decode_only) {
    // decode_gather / extract_patches on caller-owned buffers only: the normalisation table and the look-back scan cells
    std::vector<float> lut(256);
    for (int i = 0; i < 256; ++i) lut[i] = ((float)i - 128.0f) / 255.0f;
    TRYC(upload(e->lut, lut));
    TRYC(e->total.alloc(16));
    TRYC(cudaMemset(e->total.p, 0, 16));
    TRYC(e->scan_state.alloc((size_t)cfg->max_batch * 8));
    TRYC(cudaMemset(e->scan_state.p, 0, e->scan_state.bytes));
    TRYC(cudaDeviceSynchronize());
    *out = e;
    return DCU_OK;
  }
  // detector: conv1a,1b,2a,2b,3a,3b,4a,4b,Pa,Pb,Da,Db  (net.py:22-48)
  TRY(build_first(e->det_first, D[0], 1));
  const int pools[7] = {1, 0, 1, 0, 1, 0, 0};
  for (int i = 0; i < 7; ++i) TRY(build_3x3(e->det[i], {&D[1 + i]}, 1, pools[i], 0));
  TRY(build_3x3(e->det[7], {&D[8], &D[10]}, 1, 0, 0));
  if (D[9].ksize != 1 || D[9].cin != 256 || D[9].cout != 65 || D[11].ksize != 1 || D[11].cin != 256 ||
      D[11].cout != cfg->n_ids + 1 || D[8].cout != 256 || D[10].cout != 256) {
    delete e;
    return fail(DCU_ERR_INVALID, "dcu_create: head shapes do not match n_ids");
  }
  TRYC(upload(e->w_loc, std::vector<float>(D[9].weight, D[9].weight + 65 * 256)));
  TRYC(upload(e->b_loc, std::vector<float>(D[9].bias, D[9].bias + 65)));
  TRYC(upload(e->w_ids, std::vector<float>(D[11].weight, D[11].weight + (size_t)(cfg->n_ids + 1) * 256)));
  TRYC(upload(e->b_ids, std::vector<float>(D[11].bias, D[11].bias + cfg->n_ids + 1)));
  {
    auto build_head = [&](DcuEngine::TcHead& h, const DcuConvLayer& L, int nt) -> cudaError_t {
      h.nt = nt;
      h.scale = tc_weight_scale({L.weight}, {L.cout}, 256, 1);
      const std::vector<uint16_t> blocks = pack_tc({L.weight}, {L.cout}, 256, nt, h.scale, 1, nt);
      h.copies = 4;
      cudaError_t ce = h.w.alloc(blocks.size() * 2 * h.copies);
      for (int c = 0; c < h.copies && ce == cudaSuccess; ++c)
        ce = cudaMemcpy(h.w.as<uint8_t>() + (size_t)c * blocks.size() * 2, blocks.data(), blocks.size() * 2, cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) return ce;
      std::vector<float> b(nt, 0.f), ones(nt, 1.f);
      std::copy(L.bias, L.bias + L.cout, b.begin());
      if ((ce = upload(h.bias, b)) != cudaSuccess) return ce;
      return upload(h.ones, ones);
    };
    TRYC(build_head(e->tc_loc, D[9], 128));
    TRYC(build_head(e->tc_ids, D[11], cfg->n_ids + 1 <= 64 ? 64 : 128));
  }
  // refinenet: conv1a,1b,2a,2b,3a,3b,4a,4b,5a,5b,Pa,Pb  (refinenet.py:22-47)
  e->has_ref = (R != nullptr && n_ref == 12);
  if (R != nullptr && n_ref != 0 && n_ref != 12) { delete e; return fail(DCU_ERR_INVALID, "dcu_create: need 12 RefineNet layers"); }
  if (e->has_ref) {
    TRY(build_first(e->ref_first, R[0], 0));
    const int rpad[10] = {0, 0, 0, 1, 1, 1, 1, 1, 1, 1};
    const int rpool[10] = {0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
    const int rups[10] = {0, 0, 0, 0, 1, 0, 1, 0, 1, 0};
    for (int i = 0; i < 10; ++i) TRY(build_3x3(e->ref[i], {&R[1 + i]}, rpad[i], rpool[i], rups[i], i > 0 && rups[i - 1]));
    if (R[11].ksize != 1 || R[11].cin != 64 || R[11].cout != 1 || R[10].cout != 64) {
      delete e;
      return fail(DCU_ERR_INVALID, "dcu_create: RefineNet head shape");
    }
    TRYC(upload(e->ref_head_w, std::vector<float>(R[11].weight, R[11].weight + 64)));
    std::memcpy(e->ref[9].host_bn.head, R[11].weight, 64 * sizeof(float));     // pair kernel: the fused head reads them from the constant bank
    e->ref_head_b = R[11].bias[0];
  }
  // (x - 128) / 255 in fp32 with a true division, as numpy does (model_utils.py:48-49)
  {
    std::vector<float> lut(256);
    for (int i = 0; i < 256; ++i) lut[i] = ((float)i - 128.0f) / 255.0f;
    TRYC(upload(e->lut, lut));
  }
  // workspace
  const int H = cfg->height, W = cfg->width;
  const double area = (double)H * W / (320.0 * 240.0);
  // Sized for wave efficiency, not L2 residency: at the tensor-bound rate the activation traffic is a few hundred
  // GB/s, far below HBM bandwidth, while a launch with few tiles leaves most of the 148 SMs idle in its last wave.
  e->mb1 = std::max(1, (int)std::floor(64.0 / area + 1e-9));       // 64 frames at 320x240: +1.3 % over 32 (fewer launches and tails)
  e->mb2 = std::max(e->mb1, (int)std::floor(256.0 / area + 1e-9));
  e->mb1 = std::min(e->mb1, std::max(1, cfg->max_batch));
  e->mb2 = std::min(e->mb2, std::max(e->mb1, cfg->max_batch));
  e->rp = std::min(4096, std::max(64, cfg->max_patches));
  if (const char* v = getenv("DCU_MB1")) e->mb1 = std::max(1, atoi(v));
  if (const char* v = getenv("DCU_MB2")) e->mb2 = std::max(1, atoi(v));
  if (const char* v = getenv("DCU_RP")) e->rp = std::max(1, atoi(v));
  e->mb2 = std::max(e->mb2, e->mb1);
  e->mb2 = (e->mb2 / e->mb1) * e->mb1;
  const size_t det_full = (size_t)e->mb1 * 64 * H * W;                       // conv1a out (floats)
  const size_t det_low = (size_t)e->mb2 * 128 * (H / 4) * (W / 4);           // conv3a out
  // RefineNet chunk: rp patches when the upsamplings are fused (largest tensor: conv5a output, 64 x 32 x 32), at most 1024
  // when conv5b's 64 x 64 x 64 upsampled output is materialised (fp32 CUDA-core path, DCU_FUSE_UP=0)
  e->rp_plain = std::min(e->rp, 1024);
  const size_t ref_big = e->has_ref ? std::max((size_t)e->rp * 64 * 32 * 32, (size_t)e->rp_plain * 64 * 64 * 64) : 0;
  const size_t act_floats = std::max(std::max(det_full, det_low), ref_big);
  TRYC(e->c1[0].alloc(det_full * 4));
  TRYC(e->c1[1].alloc(det_full * 4));
  TRYC(cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking));
  TRYC(cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    TRYC(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
    TRYC(cudaEventCreateWithFlags(&e->ev_free[i], cudaEventDisableTiming));
  }
  if (const char* v = getenv("DCU_OVERLAP_FIRST")) e->overlap_first = atoi(v) != 0;
  if (const char* v = getenv("DCU_TC_PAIR")) e->tc_pair = atoi(v) != 0;
  if (const char* v = getenv("DCU_FUSE_UP")) e->fuse_up = atoi(v) != 0;
  if (const char* v = getenv("DCU_CHUNKED_H2D")) e->chunked_h2d = atoi(v) != 0;
  if (const char* v = getenv("DCU_SMALL_SLICES")) e->small_slices = atoi(v) != 0;
  TRYC(cudaStreamCreateWithFlags(&e->copy, cudaStreamNonBlocking));
  TRYC(cudaEventCreateWithFlags(&e->ev_copy_start, cudaEventDisableTiming));
  if (const char* v = getenv("DCU_FUSE_FIRST")) e->fuse_first = atoi(v) != 0;
  if (const char* v = getenv("DCU_DEVICE_COUNT")) e->dev_count = atoi(v) != 0;
  TRYC(cudaEventCreateWithFlags(&e->ev_total, cudaEventDisableTiming));
  if (const char* v = getenv("DCU_GRAPH")) e->use_graphs = atoi(v) != 0;
  if (const char* v = getenv("DCU_GRAPH_MAX_N")) e->graph_max_n = std::max(0, atoi(v));
  TRYC(cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking));
  if (const char* v = getenv("DCU_FLAT")) e->flat = atoi(v) != 0;
  if (e->has_ref && e->flat)
    for (int i = 0; i < 3; ++i) {
      TRYC(e->flat8[i].alloc((size_t)flat_plane_px(e->rp, 81) * 16 * 2 * 16));     // [hi|lo][128/8][plane px] x 16 B
      TRYC(cudaMemset(e->flat8[i].p, 0, e->flat8[i].bytes));
    }
  TRYC(e->act[0].alloc(act_floats * 4));
  TRYC(e->act[1].alloc(act_floats * 4));
  TRYC(e->stage2_in.alloc((size_t)e->mb2 * 64 * (H / 4) * (W / 4) * 4));
  TRYC(e->heads.alloc((size_t)e->mb2 * 512 * (H / 8) * (W / 8) * 4));
  TRYC(e->loc_arg.alloc((size_t)e->mb2 * (H / 8) * (W / 8)));
  TRYC(e->ids_arg.alloc((size_t)e->mb2 * (H / 8) * (W / 8)));
  if (const char* v = getenv("DCU_ARG_HEADS")) e->arg_heads = atoi(v) != 0;
  TRYC(e->loc.alloc((size_t)e->mb2 * 65 * (H / 8) * (W / 8) * 4));
  TRYC(e->ids.alloc((size_t)e->mb2 * (cfg->n_ids + 1) * (H / 8) * (W / 8) * 4));
  TRYC(e->counts.alloc((size_t)cfg->max_batch * 4));
  TRYC(e->offsets.alloc((size_t)cfg->max_batch * 4));
  TRYC(e->total.alloc(16));
  TRYC(e->kpts.alloc((size_t)cfg->max_patches * 16));
  TRYC(e->patches.alloc((size_t)cfg->max_patches * 576 * 4));
  TRYC(e->keys.alloc((size_t)cfg->max_patches * 8));
  TRYC(e->refined.alloc((size_t)cfg->max_patches * 8));
  TRYC(e->scan_state.alloc((size_t)std::max(cfg->max_batch, e->mb2) * 8));
  TRYC(cudaMemset(e->scan_state.p, 0, e->scan_state.bytes));
  TRYC(cudaMemset(e->total.p, 0, 16));
  TRYC(e->frames.alloc((size_t)cfg->max_batch * H * W));
  TRYC(cudaMallocHost(&e->h_frames, (size_t)cfg->max_batch * H * W * 3));      // grey or BGR frames
  TRYC(cudaMallocHost(&e->h_counts, (size_t)cfg->max_batch * 4));
  TRYC(cudaMallocHost(&e->h_offsets, (size_t)cfg->max_batch * 4));
  TRYC(cudaMallocHost(&e->h_kpts, (size_t)cfg->max_patches * 16));
  TRYC(cudaMallocHost(&e->h_refined, (size_t)cfg->max_patches * 8));
  TRYC(cudaMallocHost(&e->h_total, 16));
  TRYC(cudaDeviceSynchronize());
#undef TRY
#undef TRYC
  *out = e;
  return DCU_OK;
}

int dcu_destroy(DcuEngine* e) {
  if (!e) return DCU_OK;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  delete e;
  return DCU_OK;
}

int dcu_set_conv_impl(DcuEngine* e, int impl) {
  if (!e || (impl != DCU_CONV_FFMA && impl != DCU_CONV_TCGEN05)) return fail(DCU_ERR_INVALID, "bad conv_impl");
  if (impl != e->conv_impl) {           // captured small-batch graphs hold the old kernels
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
  }
  e->conv_impl = impl;
  return DCU_OK;
}

int64_t dcu_launch_count(const DcuEngine* e) { return e ? e->launches : 0; }

int dcu_profile_enable(DcuEngine* e, int on) {
  if (!e) return fail(DCU_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  for (auto& r : e->prof) { e->ev_pool.push_back(r.a); e->ev_pool.push_back(r.b); }
  e->prof.clear();
  e->profiling = on != 0;
  return DCU_OK;
}

int dcu_profile_read(DcuEngine* e, int cls, double* total_ms, double* total_work, int64_t* n_launches) {
  if (!e || !total_ms || !total_work || !n_launches) return fail(DCU_ERR_INVALID, "dcu_profile_read: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  double ms = 0, work = 0; int64_t n = 0;
  for (auto& r : e->prof) {
    if (r.cls != cls) continue;
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t; work += r.work; ++n;
  }
  *total_ms = ms; *total_work = work; *n_launches = n;
  return DCU_OK;
}

int dcu_profile_read_issued(DcuEngine* e, int cls, double* issued_flops) {
  if (!e || !issued_flops) return fail(DCU_ERR_INVALID, "dcu_profile_read_issued: bad argument");
  double w = 0;
  for (auto& r : e->prof)
    if (r.cls == cls) w += r.issued;
  *issued_flops = w;
  return DCU_OK;
}

int dcu_profile_records(DcuEngine* e, int cap, double* rec8, int* n_records) {
  if (!e || !n_records || (cap > 0 && !rec8)) return fail(DCU_ERR_INVALID, "dcu_profile_records: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  int i = 0;
  for (auto& r : e->prof) {
    if (i < cap) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, r.a, r.b));
      double* o = rec8 + (size_t)i * 8;
      o[0] = r.cls; o[1] = t; o[2] = r.work;
      for (int k = 0; k < 5; ++k) o[3 + k] = r.shape[k];
    }
    ++i;
  }
  *n_records = i;
  return DCU_OK;
}

double dcu_detector_flops_per_frame(const DcuEngine* e) {
  if (!e) return 0;
  const double H = e->cfg.height, W = e->cfg.width;
  double mac = 9.0 * 64 * H * W;                                                   // conv1a
  mac += 9.0 * 64 * 64 * H * W;                                                    // conv1b
  mac += 2 * 9.0 * 64 * 64 * (H / 2) * (W / 2);                                    // conv2a, 2b
  mac += 9.0 * 64 * 128 * (H / 4) * (W / 4) + 9.0 * 128 * 128 * (H / 4) * (W / 4);   // conv3a, 3b
  mac += 2 * 9.0 * 128 * 128 * (H / 8) * (W / 8);                                  // conv4a, 4b
  mac += 2 * 9.0 * 128 * 256 * (H / 8) * (W / 8);                                  // convPa, convDa
  mac += 256.0 * (65 + e->cfg.n_ids + 1) * (H / 8) * (W / 8);                      // convPb, convDb
  return 2.0 * mac;
}

double dcu_refine_flops_per_patch(const DcuEngine*) {
  double mac = 9.0 * 64 * 22 * 22 + 9.0 * 64 * 64 * 20 * 20 + 9.0 * 64 * 128 * 18 * 18 + 9.0 * 128 * 128 * 16 * 16;
  mac += 2 * 9.0 * 128 * 128 * 8 * 8 + 2 * 9.0 * 128 * 128 * 16 * 16;
  mac += 9.0 * 128 * 64 * 32 * 32 + 9.0 * 64 * 64 * 32 * 32 + 9.0 * 64 * 64 * 64 * 64 + 64.0 * 64 * 64;
  return 2.0 * mac;
}

int dcu_detector_forward(DcuEngine* e, const uint8_t* frames_dev, int n, float* loc_dev, float* ids_dev, void* stream) {
  if (!e || !frames_dev || !loc_dev || !ids_dev || n < 0) return fail(DCU_ERR_INVALID, "dcu_detector_forward: bad argument");
  if (e->decode_only) return fail(DCU_ERR_INVALID, "decode-only engine: no detector");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->cfg.height, W = e->cfg.width, cells = (H / 8) * (W / 8);
  for (int f0 = 0; f0 < n; f0 += e->mb2) {
    const int m = std::min(e->mb2, n - f0);
    int rc = detector_group(e, frames_dev + (size_t)f0 * H * W, nullptr, m, loc_dev + (size_t)f0 * 65 * cells,
                            ids_dev + (size_t)f0 * (e->cfg.n_ids + 1) * cells, s);
    if (rc) return rc;
  }
  return DCU_OK;
}

int dcu_detector_forward_f32(DcuEngine* e, const float* images_dev, int n, float* loc_dev, float* ids_dev, void* stream) {
  if (!e || !images_dev || !loc_dev || !ids_dev || n < 0) return fail(DCU_ERR_INVALID, "dcu_detector_forward_f32: bad argument");
  if (e->decode_only) return fail(DCU_ERR_INVALID, "decode-only engine: no detector");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->cfg.height, W = e->cfg.width, cells = (H / 8) * (W / 8);
  for (int f0 = 0; f0 < n; f0 += e->mb2) {
    const int m = std::min(e->mb2, n - f0);
    int rc = detector_group(e, nullptr, images_dev + (size_t)f0 * H * W, m, loc_dev + (size_t)f0 * 65 * cells,
                            ids_dev + (size_t)f0 * (e->cfg.n_ids + 1) * cells, s);
    if (rc) return rc;
  }
  return DCU_OK;
}

int dcu_extract_patches(DcuEngine* e, const float* image_dev, const int32_t* xy_dev, int k, float* patches_dev, void* stream) {
  if (!e || !image_dev || !xy_dev || !patches_dev || k < 0) return fail(DCU_ERR_INVALID, "dcu_extract_patches: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  launch_extract_patches(image_dev, e->cfg.height, e->cfg.width, xy_dev, k, patches_dev, (cudaStream_t)stream);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

int dcu_decode_gather(DcuEngine* e, const float* loc_dev, const float* ids_dev, const uint8_t* frames_dev, int n,
                      int dust_bin_ids, int append, int32_t* counts_dev, int32_t* offsets_dev, int32_t* total_dev,
                      int32_t* kpts_dev, float* patches_dev, void* stream) {
  if (!e || !loc_dev || !ids_dev || !counts_dev || !offsets_dev || !total_dev || !kpts_dev || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_decode_gather: bad argument");
  if (patches_dev && !frames_dev) return fail(DCU_ERR_INVALID, "dcu_decode_gather: patches need frames");
  if ((size_t)n * 8 > e->scan_state.bytes) return fail(DCU_ERR_INVALID, "dcu_decode_gather: n > max_batch");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = decode_smem_bytes((e->cfg.height / 8) * (e->cfg.width / 8));
  if (smem > 200 * 1024) return fail(DCU_ERR_UNSUPPORTED, "frame too large for the single-CTA decode");
  if (n == 0) { if (!append) CK(cudaMemsetAsync(total_dev, 0, 4, s)); return DCU_OK; }
  return decode_group(e, loc_dev, ids_dev, frames_dev, n, dust_bin_ids, append, counts_dev, offsets_dev, total_dev,
                      kpts_dev, patches_dev, s);
}

int dcu_refine_forward(DcuEngine* e, const float* patches_dev, const int32_t* xy_dev, int xy_stride, int p,
                       int32_t* corners_dev, float* refined_dev, float* heat_dev, void* stream) {
  if (!e || !patches_dev || !xy_dev || !refined_dev || p < 0 || xy_stride < 2)
    return fail(DCU_ERR_INVALID, "dcu_refine_forward: bad argument");
  if (!e->has_ref) return fail(DCU_ERR_INVALID, "engine was created without RefineNet weights");
  if (p > e->cfg.max_patches) return fail(DCU_ERR_CAPACITY, "p > max_patches");
  CK(cudaSetDevice(e->cfg.device));
  return refine_run(e, patches_dev, xy_dev, xy_stride, p, corners_dev, refined_dev, heat_dev, (cudaStream_t)stream);
}

int dcu_infer_batch(DcuEngine* e, const uint8_t* frames_dev, int n, int dust_bin_ids, int use_refinenet,
                    int32_t* counts_dev, int32_t* offsets_dev, int32_t* total_dev, int32_t* kpts_dev,
                    float* refined_dev, void* stream) {
  if (!e || !frames_dev || !counts_dev || !offsets_dev || !total_dev || !kpts_dev || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_infer_batch: bad argument");
  if (n > e->cfg.max_batch) return fail(DCU_ERR_INVALID, "dcu_infer_batch: n > max_batch");
  if (e->decode_only) return fail(DCU_ERR_INVALID, "decode-only engine: no detector");
  if (use_refinenet && (!e->has_ref || !refined_dev)) return fail(DCU_ERR_INVALID, "dcu_infer_batch: RefineNet not available");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->cfg.height, W = e->cfg.width;
  if (n == 0) { CK(cudaMemsetAsync(total_dev, 0, 4, s)); return DCU_OK; }
  int rc;
  for (int f0 = 0; f0 < n; f0 += e->mb2) {
    const int m = std::min(e->mb2, n - f0);
    const uint8_t* fr = frames_dev + (size_t)f0 * H * W;
    e->h2d_base = f0 / e->mb1;
    e->arg_heads_now = e->arg_heads && e->conv_impl == DCU_CONV_TCGEN05 && e->cfg.n_ids + 1 <= 64;
    rc = detector_group(e, fr, nullptr, m, e->loc.as<float>(), e->ids.as<float>(), s);
    if (rc == DCU_OK)
      rc = decode_group(e, e->loc.as<float>(), e->ids.as<float>(), fr, m, dust_bin_ids, f0 > 0, counts_dev + f0,
                        offsets_dev + f0, total_dev, kpts_dev, use_refinenet ? e->patches.as<float>() : nullptr, s);
    e->arg_heads_now = false;
    if (rc) return rc;
  }
  if (!use_refinenet) return DCU_OK;
  CK(cudaMemcpyAsync(e->h_total, total_dev, 4, cudaMemcpyDeviceToHost, s));
  // predicted hand-off: RefineNet goes into the stream before the host knows the corner count (see DcuEngine::dev_count)
  const bool can_dev_count = e->dev_count && e->patch_hint > 0 && !e->profiling && e->conv_impl == DCU_CONV_TCGEN05 && e->tc_pair &&
                             e->fuse_up && e->flat && e->flat8[0].p && e->ref[5].ups_in && e->ref[7].ups_in && e->ref[9].ups_in;
  int launched = 0;
  if (can_dev_count) {
    CK(cudaEventRecord(e->ev_total, s));
    const long long chunks = ((long long)e->patch_hint + e->rp - 1) / e->rp;
    launched = (int)std::min<long long>(e->cfg.max_patches, chunks * e->rp);
    rc = refine_run(e, e->patches.as<float>(), kpts_dev, 4, launched, nullptr, refined_dev, nullptr, s, total_dev);
    if (rc) return rc;
    CK(cudaEventSynchronize(e->ev_total));      // the count is on the host as soon as the decode has run; RefineNet is already queued behind it
  } else {
    CK(cudaStreamSynchronize(s));               // corner count first, as the reference's nonzero() does (model_utils.py:114)
  }
  const int total = std::min(e->h_total[0], e->cfg.max_patches);
  e->patch_hint = std::max(e->h_total[0], e->patch_hint - std::max(1, e->patch_hint / 16));      // recent maximum, decaying slowly
  if (total > launched) {                        // first call, or more corners than predicted: the remaining chunks with the known count
    rc = refine_run(e, e->patches.as<float>(), kpts_dev, 4, total, nullptr, refined_dev, nullptr, s, nullptr, launched);
    if (rc) return rc;
  }
  if (e->h_total[0] > e->cfg.max_patches)       // rows up to capacity are valid and refined; counts / offsets describe the full set
    return fail(DCU_ERR_CAPACITY, "corner count " + std::to_string(e->h_total[0]) + " exceeds max_patches " +
                                      std::to_string(e->cfg.max_patches));
  return DCU_OK;
}

int dcu_bgr_to_gray(DcuEngine* e, const uint8_t* bgr_dev, int n, uint8_t* gray_dev, void* stream) {
  if (!e || !bgr_dev || !gray_dev || n < 0) return fail(DCU_ERR_INVALID, "dcu_bgr_to_gray: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  launch_bgr_to_gray(bgr_dev, gray_dev, (long long)n * e->cfg.height * e->cfg.width, (cudaStream_t)stream);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

static int infer_batch_host_impl(DcuEngine* e, const uint8_t* frames_host, int n, int channels, int dust_bin_ids, int use_refinenet,
                                 int32_t* counts_host, int32_t* offsets_host, int32_t* total_host, int32_t* kpts_host,
                                 float* refined_host, void* stream);

// cv::resize's coefficient tables for INTER_LINEAR on 8-bit images (imgproc/resize.cpp): per destination index the source index and
// the two 11-bit weights, from a float32 fraction of (d + 0.5) * scale - 0.5 clamped at the borders.
static void resize_axis_table(int n_dst, int n_src, int* idx, int* a0, int* a1) {
  const double scale = (double)n_src / n_dst;
  for (int d = 0; d < n_dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int i = (int)std::floor(f);
    f -= (float)i;
    if (i < 0) { f = 0.f; i = 0; }
    if (i >= n_src - 1) { f = 0.f; i = n_src - 1; }
    auto sat = [](float v) { const long r = std::lrintf(v); return (int)std::max(-32768L, std::min(32767L, r)); };
    idx[d] = i; a0[d] = sat((1.f - f) * 2048.f); a1[d] = sat(f * 2048.f);
  }
}

int dcu_resize_u8(DcuEngine* e, const uint8_t* src_dev, int n, int src_h, int src_w, int channels, uint8_t* dst_dev, void* stream) {
  if (!e || !src_dev || !dst_dev || n < 0 || (channels != 1 && channels != 3)) return fail(DCU_ERR_INVALID, "dcu_resize_u8: bad argument");
  const int H = e->cfg.height, W = e->cfg.width;
  if (src_h < H || src_w < W)
    return fail(DCU_ERR_UNSUPPORTED, "dcu_resize_u8: only shrinking (source >= engine frame size) is bit-exact with cv2.resize; enlarge on the host");
  if (n == 0) return DCU_OK;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e->resize_hs != src_h || e->resize_ws != src_w || !e->resize_tab.p) {
    std::vector<int> tab(3 * (size_t)W + 3 * (size_t)H);
    resize_axis_table(W, src_w, tab.data(), tab.data() + W, tab.data() + 2 * W);
    resize_axis_table(H, src_h, tab.data() + 3 * W, tab.data() + 3 * W + H, tab.data() + 3 * W + 2 * H);
    CK(cudaStreamSynchronize(s));                    // an earlier launch on this stream may still read the old table
    e->resize_tab.release();
    CK(e->resize_tab.alloc(tab.size() * 4));
    CK(cudaMemcpy(e->resize_tab.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
    e->resize_hs = src_h; e->resize_ws = src_w;
  }
  launch_resize_linear_u8(src_dev, dst_dev, e->resize_tab.as<int>(), n, src_h, src_w, H, W, channels, s);
  e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

int dcu_infer_batch_host(DcuEngine* e, const uint8_t* frames_host, int n, int dust_bin_ids, int use_refinenet,
                         int32_t* counts_host, int32_t* offsets_host, int32_t* total_host, int32_t* kpts_host,
                         float* refined_host, void* stream) {
  return infer_batch_host_impl(e, frames_host, n, 1, dust_bin_ids, use_refinenet, counts_host, offsets_host, total_host, kpts_host,
                               refined_host, stream);
}

int dcu_infer_batch_host_bgr(DcuEngine* e, const uint8_t* frames_host, int n, int dust_bin_ids, int use_refinenet,
                             int32_t* counts_host, int32_t* offsets_host, int32_t* total_host, int32_t* kpts_host,
                             float* refined_host, void* stream) {
  return infer_batch_host_impl(e, frames_host, n, 3, dust_bin_ids, use_refinenet, counts_host, offsets_host, total_host, kpts_host,
                               refined_host, stream);
}

// Small-batch path: returns 1 if it did not handle the call (caller continues kernel by kernel), else a DCU_* status.
static int infer_small_graph(DcuEngine* e, const uint8_t* frames_host, int n, int channels, int dust_bin_ids, int use_refinenet,
                             int32_t* counts_host, int32_t* offsets_host, int32_t* total_host, int32_t* kpts_host,
                             float* refined_host, cudaStream_t s) {
  const int H = e->cfg.height, W = e->cfg.width;
  // patch slots: the smallest of 8 / 16 / 32 per frame that held the corner counts of the recent calls (video frames are alike);
  // a frame with more corners than slots finishes kernel by kernel below and widens the hint
  int per = 8;
  while (per < 32 && per * n < e->graph_hint) per *= 2;
  const int pfix = std::min(per * n, e->cfg.max_patches);
  DcuEngine::SmallGraph* g = nullptr;
  for (auto& x : e->graphs)
    if (x.n == n && x.dust == dust_bin_ids && x.use_ref == use_refinenet && x.pfix == pfix && x.channels == channels) g = &x;
  if (!g) {            // first call of this shape runs kernel by kernel (sets function attributes, allocates lazily)
    if (e->graphs.size() >= 64) return 1;
    DcuEngine::SmallGraph x; x.n = n; x.dust = dust_bin_ids; x.use_ref = use_refinenet; x.pfix = pfix; x.channels = channels;
    e->graphs.push_back(x);
    return 1;
  }
  if (g->dead) return 1;
  if (!g->exec) {
    cudaStream_t gs = e->gstream;
    if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); g->dead = true; return 1; }
    const int64_t l0 = e->launches;
    e->epoch_override = 0x3ffffff0u;
    int rc = DCU_OK;
    cudaError_t ce;
    if (channels == 3) {            // BGR frames: cv2.cvtColor(BGR2GRAY) (inference.py:40) on the device, inside the graph
      ce = cudaMemcpyAsync(e->bgr.p, e->h_frames, (size_t)n * H * W * 3, cudaMemcpyHostToDevice, gs);
      if (ce == cudaSuccess) {
        launch_bgr_to_gray(e->bgr.as<uint8_t>(), e->frames.as<uint8_t>(), (long long)n * H * W, gs);
        e->launches++;
        ce = cudaGetLastError();
      }
    } else {
      ce = cudaMemcpyAsync(e->frames.p, e->h_frames, (size_t)n * H * W, cudaMemcpyHostToDevice, gs);
    }
    e->arg_heads_now = e->arg_heads && e->conv_impl == DCU_CONV_TCGEN05 && e->cfg.n_ids + 1 <= 64;
    if (ce == cudaSuccess) rc = detector_group(e, e->frames.as<uint8_t>(), nullptr, n, e->loc.as<float>(), e->ids.as<float>(), gs);
    if (ce == cudaSuccess && rc == DCU_OK) ce = cudaMemsetAsync(e->scan_state.p, 0, (size_t)n * 8, gs);
    if (ce == cudaSuccess && rc == DCU_OK)
      rc = decode_group(e, e->loc.as<float>(), e->ids.as<float>(), e->frames.as<uint8_t>(), n, dust_bin_ids, 0, e->counts.as<int32_t>(),
                        e->offsets.as<int32_t>(), e->total.as<int32_t>(), e->kpts.as<int32_t>(),
                        use_refinenet ? e->patches.as<float>() : nullptr, gs);
    e->arg_heads_now = false;
    if (ce == cudaSuccess && rc == DCU_OK && use_refinenet)
      rc = refine_run(e, e->patches.as<float>(), e->kpts.as<int32_t>(), 4, pfix, nullptr, e->refined.as<float>(), nullptr, gs);
    if (ce == cudaSuccess && rc == DCU_OK) ce = cudaMemcpyAsync(e->h_total, e->total.p, 4, cudaMemcpyDeviceToHost, gs);
    if (ce == cudaSuccess && rc == DCU_OK) ce = cudaMemcpyAsync(e->h_counts, e->counts.p, (size_t)n * 4, cudaMemcpyDeviceToHost, gs);
    if (ce == cudaSuccess && rc == DCU_OK) ce = cudaMemcpyAsync(e->h_offsets, e->offsets.p, (size_t)n * 4, cudaMemcpyDeviceToHost, gs);
    if (ce == cudaSuccess && rc == DCU_OK) ce = cudaMemcpyAsync(e->h_kpts, e->kpts.p, (size_t)pfix * 16, cudaMemcpyDeviceToHost, gs);
    if (ce == cudaSuccess && rc == DCU_OK && use_refinenet)
      ce = cudaMemcpyAsync(e->h_refined, e->refined.p, (size_t)pfix * 8, cudaMemcpyDeviceToHost, gs);
    e->epoch_override = 0;
    cudaGraph_t graph = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(gs, &graph);
    g->launches = e->launches - l0;
    e->launches = l0;
    if (ce != cudaSuccess || rc != DCU_OK || ee != cudaSuccess || !graph ||
        cudaGraphInstantiate(&g->exec, graph, 0) != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      g->exec = nullptr; g->dead = true;
      return 1;
    }
    cudaGraphDestroy(graph);
  }
  std::memcpy(e->h_frames, frames_host, (size_t)n * H * W * channels);
  CK(cudaGraphLaunch(g->exec, s));
  e->launches += g->launches;
  CK(cudaStreamSynchronize(s));
  const int total = e->h_total[0];
  const int kept = std::min(total, e->cfg.max_patches);
  // hint = recent maximum: jumps up at once, decays slowly (1/16 per call)
  e->graph_hint = std::max(total, e->graph_hint - std::max(1, e->graph_hint / 16));
  if (total > g->pfix) {
    // crowded frames: more corners than the graph's patch slots -> finish kernel by kernel on the device-resident decode output
    if (use_refinenet) {
      int rc = refine_run(e, e->patches.as<float>(), e->kpts.as<int32_t>(), 4, kept, nullptr, e->refined.as<float>(), nullptr, s);
      if (rc) return rc;
      CK(cudaMemcpyAsync(e->h_refined, e->refined.p, (size_t)kept * 8, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaMemcpyAsync(e->h_kpts, e->kpts.p, (size_t)kept * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  std::memcpy(counts_host, e->h_counts, (size_t)n * 4);
  std::memcpy(offsets_host, e->h_offsets, (size_t)n * 4);
  if (kept > 0) {
    std::memcpy(kpts_host, e->h_kpts, (size_t)kept * 16);
    if (use_refinenet) std::memcpy(refined_host, e->h_refined, (size_t)kept * 8);
  }
  *total_host = total;
  if (total > e->cfg.max_patches)
    return fail(DCU_ERR_CAPACITY, "corner count " + std::to_string(total) + " exceeds max_patches " +
                                      std::to_string(e->cfg.max_patches));
  return DCU_OK;
}

static int infer_batch_host_impl(DcuEngine* e, const uint8_t* frames_host, int n, int channels, int dust_bin_ids, int use_refinenet,
                                 int32_t* counts_host, int32_t* offsets_host, int32_t* total_host, int32_t* kpts_host,
                                 float* refined_host, void* stream) {
  NvtxRange nvtx_range("dcu:infer_batch_host");
  if (!e || !frames_host || !counts_host || !offsets_host || !total_host || !kpts_host || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_infer_batch_host: bad argument");
  if (n > e->cfg.max_batch) return fail(DCU_ERR_INVALID, "dcu_infer_batch_host: n > max_batch");
  if (e->decode_only) return fail(DCU_ERR_INVALID, "decode-only engine: no detector");
  if (use_refinenet && !refined_host) return fail(DCU_ERR_INVALID, "dcu_infer_batch_host: refined_host is NULL");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t gbytes = (size_t)n * e->cfg.height * e->cfg.width;
  const size_t fbytes = gbytes * channels;
  *total_host = 0;
  if (n == 0) return DCU_OK;
  if (channels == 3 && e->bgr.p == nullptr) CK(e->bgr.alloc((size_t)e->cfg.max_batch * e->cfg.height * e->cfg.width * 3));
  if (e->use_graphs && n <= e->graph_max_n && n <= e->mb1 && !e->profiling && (!use_refinenet || e->has_ref)) {
    const int rc = infer_small_graph(e, frames_host, n, channels, dust_bin_ids, use_refinenet, counts_host, offsets_host, total_host, kpts_host,
                                     refined_host, s);
    if (rc != 1) return rc;
  }
  if (channels == 3 && e->bgr.p == nullptr) CK(e->bgr.alloc((size_t)e->cfg.max_batch * e->cfg.height * e->cfg.width * 3));
  // stage through the engine's pinned buffer unless the caller's memory is already pinned (BGR: pageable copies are used as is)
  const uint8_t* src = frames_host;
  cudaPointerAttributes pa;
  if (channels == 1 && (cudaPointerGetAttributes(&pa, frames_host) != cudaSuccess || pa.type != cudaMemoryTypeHost)) {
    cudaGetLastError();
    std::memcpy(e->h_frames, frames_host, fbytes);
    src = e->h_frames;
  }
  cudaGetLastError();
  if (channels == 3) {
    CK(cudaMemcpyAsync(e->bgr.p, src, fbytes, cudaMemcpyHostToDevice, s));
    launch_bgr_to_gray(e->bgr.as<uint8_t>(), e->frames.as<uint8_t>(), (long long)gbytes, s);
    e->launches++;
    CK(cudaGetLastError());
  } else if (e->chunked_h2d && n > e->mb1 && (e->mb2 % e->mb1) == 0 && !e->profiling) {
    const size_t frame_bytes = (size_t)e->cfg.height * e->cfg.width;
    const int n_chunks = (n + e->mb1 - 1) / e->mb1;
    while ((int)e->h2d_ev.size() < n_chunks) {
      cudaEvent_t ev;
      CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      e->h2d_ev.push_back(ev);
    }
    CK(cudaEventRecord(e->ev_copy_start, s));                  // earlier work on the caller's stream may still read e->frames
    CK(cudaStreamWaitEvent(e->copy, e->ev_copy_start, 0));
    for (int c = 0; c < n_chunks; ++c) {
      const int f0 = c * e->mb1, m = std::min(e->mb1, n - f0);
      CK(cudaMemcpyAsync(e->frames.as<uint8_t>() + (size_t)f0 * frame_bytes, src + (size_t)f0 * frame_bytes, (size_t)m * frame_bytes,
                         cudaMemcpyHostToDevice, e->copy));
      CK(cudaEventRecord(e->h2d_ev[c], e->copy));
    }
    e->h2d_active = true;
  } else {
    CK(cudaMemcpyAsync(e->frames.p, src, fbytes, cudaMemcpyHostToDevice, s));
  }
  int rc = dcu_infer_batch(e, e->frames.as<uint8_t>(), n, dust_bin_ids, use_refinenet, e->counts.as<int32_t>(),
                           e->offsets.as<int32_t>(), e->total.as<int32_t>(), e->kpts.as<int32_t>(),
                           e->refined.as<float>(), s);
  e->h2d_active = false;
  if (rc && rc != DCU_ERR_CAPACITY) return rc;
  int total;
  if (use_refinenet) {
    total = e->h_total[0];        // already fetched by dcu_infer_batch
  } else {
    CK(cudaMemcpyAsync(e->h_total, e->total.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    total = e->h_total[0];
  }
  const int kept = std::min(total, e->cfg.max_patches);
  CK(cudaMemcpyAsync(e->h_counts, e->counts.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(e->h_offsets, e->offsets.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  if (kept > 0) {
    CK(cudaMemcpyAsync(e->h_kpts, e->kpts.p, (size_t)kept * 16, cudaMemcpyDeviceToHost, s));
    if (use_refinenet) CK(cudaMemcpyAsync(e->h_refined, e->refined.p, (size_t)kept * 8, cudaMemcpyDeviceToHost, s));
  }
  CK(cudaStreamSynchronize(s));
  std::memcpy(counts_host, e->h_counts, (size_t)n * 4);
  std::memcpy(offsets_host, e->h_offsets, (size_t)n * 4);
  if (kept > 0) {
    std::memcpy(kpts_host, e->h_kpts, (size_t)kept * 16);
    if (use_refinenet) std::memcpy(refined_host, e->h_refined, (size_t)kept * 8);
  }
  *total_host = total;
  if (total > e->cfg.max_patches)
    return fail(DCU_ERR_CAPACITY, "corner count " + std::to_string(total) + " exceeds max_patches " +
                                      std::to_string(e->cfg.max_patches));
  return DCU_OK;
}

int dcu_dc_metrics(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev, int n,
                   const int64_t* loc_target_dev, const int64_t* ids_target_dev, int dust_bin_ids, float* l2_dev,
                   float* ratio_dev, int32_t* valid_dev, void* stream) {
  if (!e || !counts_dev || !offsets_dev || !kpts_dev || !loc_target_dev || !ids_target_dev || !l2_dev || !ratio_dev || !valid_dev || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_dc_metrics: bad argument");
  if (dust_bin_ids < 0 || dust_bin_ids > 63) return fail(DCU_ERR_INVALID, "dcu_dc_metrics: ids above 63 are not supported");
  CK(cudaSetDevice(e->cfg.device));
  MetricsParams p{};
  p.counts = counts_dev; p.offsets = offsets_dev; p.kpts = kpts_dev;
  p.loc_target = reinterpret_cast<const long long*>(loc_target_dev); p.ids_target = reinterpret_cast<const long long*>(ids_target_dev);
  p.n = n; p.h = e->cfg.height / 8; p.w = e->cfg.width / 8; p.dust_bin = dust_bin_ids; p.max_rows = e->cfg.max_patches;
  p.l2 = l2_dev; p.ratio = ratio_dev; p.valid = valid_dev;
  launch_dc_metrics(p, (cudaStream_t)stream);
  if (n > 0) e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

int dcu_refinenet_metrics(DcuEngine* e, const float* heat_pred_dev, const int32_t* corners_pred_dev, const float* heat_target_dev, int p,
                          float* dist_dev, void* stream) {
  if (!e || (!heat_pred_dev && !corners_pred_dev) || !heat_target_dev || !dist_dev || p < 0)
    return fail(DCU_ERR_INVALID, "dcu_refinenet_metrics: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  launch_heat_argmax_dist(heat_pred_dev, corners_pred_dev, heat_target_dev, p, 64, 64, dist_dev, (cudaStream_t)stream);
  if (p > 0) e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

int dcu_synth_frames(DcuEngine* e, const DcuSynthFrame* params_host, int n, uint64_t seed, int first_index, const uint8_t* board_dev,
                     int board_px, uint8_t* frames_dev, void* stream) {
  if (!e || !params_host || !board_dev || !frames_dev || n < 0 || board_px < 2 || first_index < 0)
    return fail(DCU_ERR_INVALID, "dcu_synth_frames: bad argument");
  if (n == 0) return DCU_OK;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int H = e->cfg.height, W = e->cfg.width;
  const int lat_cap = (H / 4 + 3) * (W / 4 + 3);
  for (int i = 0; i < n; ++i) {
    const DcuSynthFrame& P = params_host[i];
    if (P.lat_step < 4 || P.lat_h != H / P.lat_step + 3 || P.lat_w != W / P.lat_step + 3 || P.n_boards < 0 || P.n_boards > 4)
      return fail(DCU_ERR_INVALID, "dcu_synth_frames: bad frame parameters (lattice step >= 4, lattice size H/step+3 x W/step+3, <= 4 boards)");
  }
  if (e->synth_params.bytes < (size_t)n * sizeof(DcuSynthFrame)) { e->synth_params.release(); CK(e->synth_params.alloc((size_t)n * sizeof(DcuSynthFrame))); }
  if (e->synth_lat.bytes < (size_t)n * lat_cap) { e->synth_lat.release(); CK(e->synth_lat.alloc((size_t)n * lat_cap)); }
  CK(cudaMemcpyAsync(e->synth_params.p, params_host, (size_t)n * sizeof(DcuSynthFrame), cudaMemcpyHostToDevice, s));
  CK(launch_synth_frames(e->synth_params.as<DcuSynthFrame>(), board_dev, board_px, e->synth_lat.as<uint8_t>(), lat_cap, seed, first_index, n, H, W,
                         frames_dev, s));
  e->launches += 2;
  return DCU_OK;
}

int dcu_warp_perspective_u8(DcuEngine* e, const uint8_t* src_dev, int src_h, int src_w, const double* minv9_host, uint8_t* dst_dev, int dst_h,
                            int dst_w, void* stream) {
  if (!e || !src_dev || !minv9_host || !dst_dev || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1 || src_h > 32767 || src_w > 32767)
    return fail(DCU_ERR_INVALID, "dcu_warp_perspective_u8: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  if (e->synth_m.bytes < 72) CK(e->synth_m.alloc(72));
  CK(cudaMemcpyAsync(e->synth_m.p, minv9_host, 72, cudaMemcpyHostToDevice, s));
  CK(launch_warp_perspective_u8(src_dev, src_h, src_w, e->synth_m.as<double>(), dst_dev, dst_h, dst_w, s));
  e->launches++;
  return DCU_OK;
}

int dcu_pixel_error(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev, const float* refined_dev,
                    int n, const int32_t* tcounts_dev, const int32_t* toffsets_dev, const double* target_dev, int32_t* status_dev,
                    double* out_dev, void* stream) {
  if (!e || !counts_dev || !offsets_dev || !kpts_dev || !refined_dev || !tcounts_dev || !toffsets_dev || !target_dev || !status_dev ||
      !out_dev || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_pixel_error: bad argument");
  CK(cudaSetDevice(e->cfg.device));
  PixelErrorParams p{};
  p.counts = counts_dev; p.offsets = offsets_dev; p.kpts = kpts_dev; p.refined = refined_dev; p.n = n; p.max_rows = e->cfg.max_patches;
  p.tcounts = tcounts_dev; p.toffsets = toffsets_dev; p.target = target_dev; p.status = status_dev; p.out = out_dev;
  launch_pixel_error(p, (cudaStream_t)stream);
  if (n > 0) e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

// ---- batched solve_pnp (inference.py:15-29) ----
static int pnp_object_table(DcuEngine* e, int col_count, int row_count, double square_len) {
  // object_points[:, :2] = meshgrid(arange(1,row_count), arange(1,col_count)).reshape(2,-1).T * square_len  (float32 storage):
  // point p -> ((p % (row_count-1)) + 1, (p / (row_count-1)) + 1) * square_len
  if (col_count < 2 || row_count < 2) return fail(DCU_ERR_INVALID, "solve_pnp: board needs >= 2 x 2 squares");
  if (e->pnp_cols == col_count && e->pnp_rows == row_count && e->pnp_sq == square_len && e->pnp_obj.p) return DCU_OK;
  const int n_obj = (col_count - 1) * (row_count - 1);
  std::vector<float> t((size_t)n_obj * 2);
  for (int p = 0; p < n_obj; ++p) {
    t[2 * p] = (float)((double)((p % (row_count - 1)) + 1) * square_len);
    t[2 * p + 1] = (float)((double)((p / (row_count - 1)) + 1) * square_len);
  }
  e->pnp_obj.release();
  CK(upload(e->pnp_obj, t));
  e->pnp_cols = col_count; e->pnp_rows = row_count; e->pnp_sq = square_len;
  return DCU_OK;
}

static int solve_pnp_batch_rows(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev,
                                const float* refined_dev, int n, int max_rows, int col_count, int row_count, double square_len,
                                const double* camera_matrix9, const double* dist_coeffs, int n_dist, int32_t* ret_dev,
                                double* rvec_dev, double* tvec_dev, void* stream) {
  if (!e || !counts_dev || !offsets_dev || !kpts_dev || !camera_matrix9 || !ret_dev || !rvec_dev || !tvec_dev || n < 0 ||
      n_dist < 0 || (n_dist > 0 && !dist_coeffs))
    return fail(DCU_ERR_INVALID, "dcu_solve_pnp_batch: bad argument");
  if (n_dist > 8) return fail(DCU_ERR_UNSUPPORTED, "dcu_solve_pnp_batch: at most 8 distortion coefficients (k1 k2 p1 p2 k3 k4 k5 k6)");
  CK(cudaSetDevice(e->cfg.device));
  int rc = pnp_object_table(e, col_count, row_count, square_len);
  if (rc) return rc;
  PnpParams q{};
  q.counts = counts_dev; q.offsets = offsets_dev; q.kpts = kpts_dev; q.refined = refined_dev;
  q.obj = e->pnp_obj.as<float>(); q.n = n; q.n_obj = (col_count - 1) * (row_count - 1); q.max_rows = max_rows;
  q.ret = ret_dev; q.rvec = rvec_dev; q.tvec = tvec_dev;
  launch_pnp_batch(q, camera_matrix9, dist_coeffs, n_dist, (cudaStream_t)stream);
  if (n > 0) e->launches++;
  CK(cudaGetLastError());
  return DCU_OK;
}

int dcu_solve_pnp_batch(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev,
                        const float* refined_dev, int n, int col_count, int row_count, double square_len,
                        const double* camera_matrix9, const double* dist_coeffs, int n_dist, int32_t* ret_dev,
                        double* rvec_dev, double* tvec_dev, void* stream) {
  // kpts_dev / refined_dev are the [max_patches] row buffers of dcu_infer_batch: frames whose rows were dropped at capacity are cut there
  return solve_pnp_batch_rows(e, counts_dev, offsets_dev, kpts_dev, refined_dev, n, e ? e->cfg.max_patches : 0, col_count, row_count,
                              square_len, camera_matrix9, dist_coeffs, n_dist, ret_dev, rvec_dev, tvec_dev, stream);
}

int dcu_solve_pnp_batch_host(DcuEngine* e, const int32_t* counts_host, const int32_t* kpts_host, const float* refined_host, int n,
                             int col_count, int row_count, double square_len, const double* camera_matrix9,
                             const double* dist_coeffs, int n_dist, int32_t* ret_host, double* rvec_host, double* tvec_host,
                             void* stream) {
  if (!e || !counts_host || !kpts_host || !ret_host || !rvec_host || !tvec_host || n < 0)
    return fail(DCU_ERR_INVALID, "dcu_solve_pnp_batch_host: bad argument");
  if (n == 0) return DCU_OK;
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<int32_t> offs(n);
  long long total = 0;
  for (int i = 0; i < n; ++i) { if (counts_host[i] < 0) return fail(DCU_ERR_INVALID, "negative count"); offs[i] = (int32_t)total; total += counts_host[i]; }
  DevBuf c, o, k, r, ret, rv, tv;
  auto freeall = [&]() { c.release(); o.release(); k.release(); r.release(); ret.release(); rv.release(); tv.release(); };
#define PK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { freeall(); return fail(DCU_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(_e)); } } while (0)
  PK(c.alloc((size_t)n * 4)); PK(o.alloc((size_t)n * 4)); PK(k.alloc((size_t)total * 16)); PK(ret.alloc((size_t)n * 4));
  PK(rv.alloc((size_t)n * 24)); PK(tv.alloc((size_t)n * 24));
  PK(cudaMemcpyAsync(c.p, counts_host, (size_t)n * 4, cudaMemcpyHostToDevice, s));
  PK(cudaMemcpyAsync(o.p, offs.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
  if (total > 0) PK(cudaMemcpyAsync(k.p, kpts_host, (size_t)total * 16, cudaMemcpyHostToDevice, s));
  if (refined_host) {
    PK(r.alloc((size_t)total * 8));
    if (total > 0) PK(cudaMemcpyAsync(r.p, refined_host, (size_t)total * 8, cudaMemcpyHostToDevice, s));
  }
  int rc = solve_pnp_batch_rows(e, c.as<int32_t>(), o.as<int32_t>(), k.as<int32_t>(), refined_host ? r.as<float>() : nullptr, n,
                                (int)std::min<long long>(total, 0x7fffffff), col_count, row_count, square_len, camera_matrix9, dist_coeffs,
                                n_dist, ret.as<int32_t>(), rv.as<double>(), tv.as<double>(), stream);
  if (rc) { freeall(); return rc; }
  PK(cudaMemcpyAsync(ret_host, ret.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(rvec_host, rv.p, (size_t)n * 24, cudaMemcpyDeviceToHost, s));
  PK(cudaMemcpyAsync(tvec_host, tv.p, (size_t)n * 24, cudaMemcpyDeviceToHost, s));
  PK(cudaStreamSynchronize(s));
#undef PK
  freeall();
  return DCU_OK;
}

int dcu_debug_tc_stats(DcuEngine* e, int enable, uint64_t* out8) {
  if (!e) return fail(DCU_ERR_INVALID, "null engine");
  CK(cudaSetDevice(e->cfg.device));
  CK(cudaDeviceSynchronize());
  if (out8 && g_tc_stats) CK(cudaMemcpy(out8, g_tc_stats, 64, cudaMemcpyDeviceToHost));
  if (enable) {
    if (!g_tc_stats) CK(cudaMalloc(&g_tc_stats, 64));
    CK(cudaMemset(g_tc_stats, 0, 64));
  } else if (g_tc_stats) {
    cudaFree(g_tc_stats);
    g_tc_stats = nullptr;
  }
  return DCU_OK;
}

int dcu_debug_conv_layer(DcuEngine* e, int net, int layer, int conv_impl, const float* in_dev, int n, int h, int w,
                         float* out_dev, void* stream) {
  if (!e || !in_dev || !out_dev || n < 1) return fail(DCU_ERR_INVALID, "dcu_debug_conv_layer: bad argument");
  if (e->decode_only) return fail(DCU_ERR_INVALID, "decode-only engine: no networks");
  if (net == 1 && !e->has_ref) return fail(DCU_ERR_INVALID, "no RefineNet in this engine");
  CK(cudaSetDevice(e->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const FirstLayer* fl = nullptr;
  const Layer3x3* l = nullptr;
  if (net == 0) {
    if (layer == 0) fl = &e->det_first;
    else if (layer >= 1 && layer <= 8) l = &e->det[layer - 1];
  } else if (net == 1) {
    if (layer == 0) fl = &e->ref_first;
    else if (layer >= 1 && layer <= 10) l = &e->ref[layer - 1];
  }
  if (!fl && !l) return fail(DCU_ERR_INVALID, "dcu_debug_conv_layer: layer is not a 3x3 convolution");
  const int cin = fl ? 1 : l->cin, cout = fl ? 64 : l->cout, pad = fl ? fl->pad : l->pad;
  const int ho = h + 2 * pad - 2, wo = w + 2 * pad - 2;
  int hf = ho, wf = wo;
  if (l && l->pool) { hf = ho / 2; wf = wo / 2; }
  if (l && l->ups) { hf = ho * 2; wf = wo * 2; }
  DevBuf tin, tout;
  CK(tin.alloc((size_t)n * cin * h * w * 4));
  CK(tout.alloc((size_t)n * cout * hf * wf * 4));
  int rc = DCU_OK;
  const bool h2 = !fl && conv_impl == DCU_CONV_TCGEN05;     // the tcgen05 kernel reads and writes the H2 layout
  // upsample fusion (run_3x3): the consumer takes the low-resolution tensor, the producer stores one
  const bool fu = h2 && e->fuse_up && e->tc_pair && (l->ups_in || l->ups);
  // RefineNet layers 1..6 in the F2 layouts refine_run uses (same rule: pair kernel + upsample fusion + DCU_FLAT)
  const bool flat = h2 && net == 1 && layer >= 1 && layer <= 6 && e->flat && e->fuse_up && e->tc_pair && e->ref[5].ups_in;
  const bool flat_in = flat && layer != 3, flat_out = flat && layer != 2 && layer != 6;
  FlatIn fin{}; H2Layout lin{}, lout{};
  if (flat_in) {
    const int hi = (layer == 6) ? h / 2 : h, wi = (layer == 6) ? w / 2 : w, g = (layer >= 4) ? 1 : 0;   // g: zero gutter column / row
    fin.period = (hi + g) * (wi + g); fin.row = wi + g; fin.plane_px = flat_plane_px(n, fin.period);
    lin = h2_flat(cin, fin.period, fin.row, fin.plane_px);
    tin.release();
    CK(tin.alloc((size_t)fin.plane_px * 16 * 2 * (cin / 8)));
    CK(cudaMemsetAsync(tin.p, 0, tin.bytes, s));
  }
  if (flat_out) {
    const int hs = (layer == 5) ? ho : hf, ws = (layer == 5) ? wo : wf, g = (layer >= 3) ? 1 : 0;
    const int period = (hs + g) * (ws + g);
    const long long pl = flat_plane_px(n, period);
    lout = h2_flat(cout, period, ws + g, pl);
    tout.release();
    CK(tout.alloc((size_t)pl * 16 * 2 * (cout / 8)));
    CK(cudaMemsetAsync(tout.p, 0, tout.bytes, s));
  }
  if (fl) {
    rc = run_first(e, *fl, nullptr, in_dev, tout.as<float>(), n, h, w, 0, s);
  } else {
    if (flat_in) launch_nchw_to_h2(in_dev, tin.p, n, cin, layer == 6 ? h / 2 : h, layer == 6 ? w / 2 : w, s, layer == 6 ? 2 : 1, &lin);
    else if (h2 && fu && l->ups_in) launch_nchw_to_h2(in_dev, tin.p, n, cin, h / 2, w / 2, s, 2);
    else if (h2) launch_nchw_to_h2(in_dev, tin.p, n, cin, h, w, s);
    else launch_nchw_to_c4(in_dev, tin.as<float>(), n, cin, h, w, s);
    rc = run_3x3(e, *l, conv_impl, tin.as<float>(), tout.as<float>(), n, h, w, nullptr, s, fu, flat_in ? &fin : nullptr,
                 flat_out ? &lout : nullptr);
  }
  if (rc == DCU_OK) {
    if (flat_out && layer == 5) launch_h2_to_nchw(tout.p, out_dev, n, cout, ho, wo, s, 2, &lout);
    else if (flat_out) launch_h2_to_nchw(tout.p, out_dev, n, cout, hf, wf, s, 1, &lout);
    else if (h2 && fu && l->ups) launch_h2_to_nchw(tout.p, out_dev, n, cout, ho, wo, s, 2);
    else if (h2) launch_h2_to_nchw(tout.p, out_dev, n, cout, hf, wf, s);
    else launch_c4_to_nchw(tout.as<float>(), out_dev, n, cout, hf, wf, s);
    cudaError_t ce = cudaStreamSynchronize(s);
    if (ce != cudaSuccess) rc = fail(DCU_ERR_CUDA, std::string("dcu_debug_conv_layer: ") + cudaGetErrorString(ce));
  }
  tin.release(); tout.release();
  return rc;
}

}  // extern "C"
