// Shared declarations for the deepcharuco_b200 CUDA sources (sm_100a only).
//
// Activation layouts in HBM.  Both keep "planes of 16-byte pixels", so that a TMA box of a plane lands in shared memory
// as the no-swizzle K-major core-matrix layout tcgen05.mma reads (8 horizontally adjacent pixels x 16 B), at ANY pixel
// offset, and one halo tile serves all nine taps of a 3x3 convolution.
//   "C4" (fp32 CUDA-core path):  float  [n][C/4][H][W][4]      -- 4 fp32 channels per 16-byte pixel
//   "H2" (tcgen05 path):         __half [n][2][C/8][H][W][8]   -- index 0: x_hi = fp16(x), index 1: x_lo = fp16(x - x_hi);
//        8 fp16 channels per 16-byte pixel.  Same bytes per element as fp32 (2 + 2), 22 significant bits, and it is
//        exactly the operand pair the split-precision MMA consumes, so the consumer needs no conversion pass: the
//        producing kernel's epilogue does the split once per element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcu {

// Addressing of an H2 tensor in units of 16-byte pixels (8 fp16 channels): pixel (img, k-group kg, y, x) has its hi half at
//   img * img + kg * plane + y * row + x      and its lo half `lo` pixels further.
//   standard:  [n][hi|lo][C/8][H][W]          img = 2*(C/8)*H*W, plane = H*W, lo = (C/8)*H*W, row = W
//   flat "F2": [hi|lo][C/8][plane_px]         img = period, plane = plane_px, lo = (C/8)*plane_px, row = stored row stride.
// F2 keeps ALL images of a launch in one run of pixels per channel group (image k starts at pixel k*period), so that a
// convolution can treat the whole batch as a 1-D signal (conv_tc2.cu, FLAT mode): RefineNet's small maps (22x22 ... 8x8)
// then fill the 128-pixel MMA tiles instead of leaving most of a 16x8 tile empty.  Same-padded layers store their maps
// with one zero gutter column and row (8x8 data in a 9x9 cell) that serves as the padding of every neighbour.
struct H2Layout {
  long long img, plane, lo;
  int row;
};
__host__ __device__ inline H2Layout h2_standard(int c, int h, int w) {
  H2Layout l;
  l.plane = (long long)h * w; l.img = 2LL * (c >> 3) * l.plane; l.lo = (long long)(c >> 3) * l.plane; l.row = w;
  return l;
}
__host__ __device__ inline H2Layout h2_flat(int c, int period, int row, long long plane_px) {
  H2Layout l;
  l.img = period; l.plane = plane_px; l.lo = (long long)(c >> 3) * plane_px; l.row = row;
  return l;
}

// ---- one 3x3 convolution layer, device-side view ---------------------------------------------------
struct ConvParams {
  const float* in;        // C4 [n][cin/4][hin][win][4]
  float* out;             // C4 [n][cout_total/4][hout'][wout'][4]  (hout' after pool / upsample)
  const float* bias;      // [cout_total]
  const float* alpha;     // [cout_total]
  const float* beta;      // [cout_total]
  int n;                  // images (frames or patches)
  int cin, cout_total;    // cout_total: channel count of the output tensor
  int hin, win;           // input spatial size
  int hout, wout;         // conv output size before pool/upsample (hin + 2*pad - 2)
  int pad;                // 0 (valid) or 1 (same)
  int pool;               // 1: 2x2 max-pool in the epilogue
  int ups;                // 1: write each output to a 2x2 block (nearest x2 upsample)
  // fused RefineNet head (convPa -> convPb 1x1 -> 64x64 arg-max), enabled when head_w != nullptr
  const float* head_w;    // [64] convPb weight
  float head_b;           // convPb bias
  unsigned long long* head_key;  // [n] packed (orderable heat value << 32 | ~index)
  float* heat;            // optional [n][hout][wout] dump of the heat map (tests), may be null
  // tcgen05 path, 1x1 "logits" mode (detector heads convPb / convDb, net.py:74,77): ksize == 1, no BN / ReLU, fp32 NCHW out
  int ksize;              // 3 (default when 0) or 1
  int cin_offset;         // first input channel inside the input tensor (multiple of 8)
  float* logits;          // [n][n_valid][hout][wout] fp32, or null
  uint8_t* arg_out;       // 1x1 "logits" mode: [n][hout*wout] arg-max over the n_valid channels (first max wins), or null.  With logits ==
                          // null the logits are never written: the fused pipeline decodes from the two arg-max maps (82 planes -> 2 bytes / cell)
  int n_valid;            // real output channels (the weight block is zero-padded to NT rows)
  float wscale_inv;       // tcgen05 path: 2^-s, undoes the power-of-two weight scaling of the fp16 split (1.0 otherwise)
  unsigned long long* stats;   // optional [8] cycle counters for the tcgen05 kernel's roles (profiling), may be null
  // tcgen05 pair kernel only.  out_layout.plane != 0: explicit addressing of the output tensor (after pool) instead of the
  // standard one.  flat_in != 0: `in` is an F2 tensor with image period in_period and stored row stride in_row (hin x win
  // are the data extents inside it); conv_tc2.cu FLAT mode.
  H2Layout out_layout;
  int flat_in, in_period, in_row;
  // image count taken from DEVICE memory (sync-free detector -> RefineNet hand-off): when n_dev != nullptr the kernel processes
  // clamp(*n_dev - n_off, 0, n) images; n stays the host-side upper bound that sizes tensor maps, layouts and the grid
  const int* n_dev; int n_off;
  int mt1;                      // pair kernel: one m-tile (16 x 8 pixels) per CTA -- small launches; the caller's tensor map box is 10 x 18 pixels
  const struct TcBn* host_bn;   // host copy of bias / alpha / beta: the pair kernel takes them by value (constant bank)
  // FIRST mode of the pair kernel: `in` is unused; the 64 input channels are conv1a (+BN+ReLU) of the frames, computed in-kernel
  const struct FirstWeights* first_w;   // host copy of conv1a's weights / BN constants, or null
  const uint8_t* first_u8;              // [n][hin][win] u8 frames (normalised in-kernel), or
  const float* first_f32;               // [n][hin][win] already normalised images
};
struct TcBn { float v[3][512]; float head[64]; };   // [bias | alpha | beta][channel]; head: weights of a fused 1x1 head (64 -> 1), else unused
struct FirstWeights { float w[9 * 64]; float bias[64], alpha[64], beta[64]; };   // conv1a: w[tap][channel], BN as y = fma(acc + bias, alpha, beta)

// FFMA path: weights packed [cin/4][9 taps][4 cin][cout_total] with the cout axis permuted per 64-block
// so a thread's 8 channels are two float4 that are bank-conflict free (see conv_ffma.cu).
struct FfmaWeights {
  const float* w;
};

void launch_conv3x3_ffma(const ConvParams& p, const float* w_packed, cudaStream_t s);

// first layer (cin == 1): direct convolution from u8 frames through the (x-128)/255 LUT, or from fp32 patches
struct FirstConvParams {
  const uint8_t* in_u8;   // [n][hin][win] or null
  const float* in_f32;    // [n][hin][win] or null
  const float* lut;       // [256] fp32 (x-128)/255, host-computed (model_utils.py:46-50)
  float* out;             // C4 [n][16][hout][wout][4], or H2 [n][2][8][hout][wout][8] fp16 when out_h2 != 0
  int out_h2;
  H2Layout out_layout;    // out_h2 only; plane == 0: the standard H2 tensor
  const float* w;         // [9][64]
  const float* bias; const float* alpha; const float* beta;   // [64]
  int n, hin, win, hout, wout, pad;
  const int* n_dev; int n_off;     // as ConvParams: process clamp(*n_dev - n_off, 0, n) images
};
void launch_conv_first(const FirstConvParams& p, cudaStream_t s);

// 1x1 heads of the detector (convPb 256->65, convDb 256->n_ids+1), NCHW outputs (net.py:74,77)
struct HeadParams {
  const float* in;        // C4 [n][128][h][w][4] (or H2 [n][2][64][h][w][8] fp16 when in_h2 != 0): channels 0..255 = cPa, 256..511 = cDa
  int in_h2;
  const float* w_loc;     // [65][256]
  const float* b_loc;     // [65]
  const float* w_ids;     // [n_ids+1][256]
  const float* b_ids;     // [n_ids+1]
  float* loc;             // [n][65][h][w]
  float* ids;             // [n][n_ids+1][h][w]
  int n, h, w, n_ids1;
};
void launch_heads_1x1(const HeadParams& p, cudaStream_t s);

// decode + patch gather (model_utils.py:53-124, 19-36)
struct DecodeParams {
  const float* loc; const float* ids; const uint8_t* frames; const float* lut;
  int n, H, W, h, w, n_ids1, dust_bin;
  int append;
  const uint8_t* loc_arg; const uint8_t* ids_arg;   // when non-null: per-cell arg-max maps [n][h*w] from the head epilogues (loc / ids are not read)
  int32_t* counts; int32_t* offsets; int32_t* total; int32_t* kpts; float* patches;
  int max_patches;
  unsigned long long* scan_state;   // [max_batch] chained-scan cells
  unsigned int epoch;
};
void launch_decode_gather(const DecodeParams& p, cudaStream_t s);
size_t decode_smem_bytes(int cells);

void launch_bgr_to_gray(const uint8_t* bgr, uint8_t* gray, long long n_px, cudaStream_t s);   // n_px % 4 == 0
void launch_resize_linear_u8(const uint8_t* src, uint8_t* dst, const int* tab_dev, int n, int hs, int ws, int h, int w, int ch, cudaStream_t s);
void launch_extract_patches(const float* image, int H, int W, const int32_t* xy, int k, float* patches, cudaStream_t s);

// RefineNet tail: packed arg-max key -> (col,row) and refined (x,y)   (refinenet.py:111-114)
void launch_refine_finalize(const unsigned long long* keys, const int32_t* xy, int xy_stride, int p,
                            int32_t* corners, float* refined, cudaStream_t s, const int* n_dev = nullptr, int n_off = 0);

// layout converters used by the debug/test entry point
void launch_nchw_to_c4(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s);
void launch_c4_to_nchw(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s);
// sub: the NCHW input is [h*sub][w*sub] and pixel (y*sub, x*sub) is taken; rep: the NCHW output is [h*rep][w*rep] (nearest)
// lay: addressing of the H2 side (nullptr: standard)
void launch_nchw_to_h2(const float* in, void* out, int n, int c, int h, int w, cudaStream_t s, int sub = 1,
                       const H2Layout* lay = nullptr);
void launch_h2_to_nchw(const void* in, float* out, int n, int c, int h, int w, cudaStream_t s, int rep = 1,
                       const H2Layout* lay = nullptr);

// tcgen05 path (conv_tc.cu)
struct TcLayerPack {
  const float* w_blocks;  // per (chunk of 16 cin, tap): [2 k-groups][N hi rows | N lo rows][8] fp16
  int cin, cout;          // cout here = N handled per CTA pass (64 or 128)
};
int tc_supported_shape(int cin, int cout);
void tc_tile_arrangement(int nt, int hout, int wout, int* tr, int* tc);
cudaError_t launch_conv3x3_tc(const ConvParams& p, const float* w_blocks, int n_slices, int w_copies, const void* tmap_in,
                              int sm_count, cudaStream_t s);

// CTA-pair (cta_group::2) variant, conv_tc2.cu
int tc2_block_bytes(int nt);
bool tc2_segmented(int cin);   // two-level accumulation for a layer with cin input channels (DCU_SEG policy) -> it runs in 64-channel slices
int tc2_stage_blocks(int up);       // weight blocks per bulk-copy stage (tensor-map box)
// up != 0: p.in is the LOW-resolution tensor (hin x win) whose 2x nearest upsampling is the layer's input; hout = 2*hin
int tc2_flat_rows(int in_row, int pad_or_up, int tile128);   // FLAT mode: rows of 16 pixels per halo box (tensor-map box height); tile128: 128-pixel CTA tiles (UP / one m-tile)
// issued_flops (optional): 2 x the MACs the tensor pipes execute for this launch (3 products, padded tiles included)
cudaError_t launch_conv_tc2(const ConvParams& p, int n_slices, int up, const void* tmap_a, const void* tmap_w0, const void* tmap_w1,
                            int sm_count, cudaStream_t s, double* issued_flops = nullptr);

// batched board pose (pnp.cu): per frame f, corners kpts[offsets[f] .. +counts[f]) (x, y, id, cell), optional refined (x, y)
struct PnpParams {
  const int32_t* counts; const int32_t* offsets; const int32_t* kpts; const float* refined;
  const float* obj;       // [n_obj][2] board coordinates of the inner corners (device)
  int n, n_obj;
  int max_rows;           // rows the kpts / refined buffers hold: a frame whose rows were dropped at capacity (decode_gather) is cut there
  int32_t* ret; double* rvec; double* tvec;     // [n], [n][3], [n][3]
};
void launch_pnp_batch(const PnpParams& q, const double* camera9, const double* dist, int n_dist, cudaStream_t s);

// detector validation metric (metrics.cu): per-sample distance / match ratio of the decode output against label maps
struct MetricsParams {
  const int32_t* counts; const int32_t* offsets; const int32_t* kpts;
  const long long* loc_target; const long long* ids_target;     // [n][h][w] int64, as the reference's dataset yields them
  int n, h, w, dust_bin;
  int max_rows;                                                 // rows the kpts buffer holds (counts / offsets may point beyond: clamped)
  float* l2; float* ratio; int32_t* valid;                      // [n]
};
void launch_dc_metrics(const MetricsParams& p, cudaStream_t s);
// utils.pixel_error (utils.py:33-52) per frame on the engine's result rows and float64 labels (metrics.cu)
struct PixelErrorParams {
  const int32_t* counts; const int32_t* offsets; const int32_t* kpts; const float* refined;
  int n, max_rows;
  const int32_t* tcounts; const int32_t* toffsets; const double* target;     // labels: [sum tcounts][3] = x, y, id
  int32_t* status; double* out;                                              // [n], [n][6]
};
void launch_pixel_error(const PixelErrorParams& p, cudaStream_t s);
// per sample |argmax(pred) - argmax(target)|_2 over h x w maps; pred == null: the predicted arg-max comes from pred_corners (col, row)
void launch_heat_argmax_dist(const float* pred, const int32_t* pred_corners, const float* target, int p, int h, int w, float* dist,
                             cudaStream_t s);

// synthetic frames (synth.cu); DcuSynthFrame is declared in include/deepcharuco_b200.h
}  // namespace dcu
struct DcuSynthFrame;
namespace dcu {
cudaError_t launch_synth_frames(const ::DcuSynthFrame* params_dev, const uint8_t* board_dev, int board_px, uint8_t* lattice_dev, int lat_cap,
                                uint64_t seed, int first_index, int n, int H, int W, uint8_t* frames_dev, cudaStream_t s);
cudaError_t launch_warp_perspective_u8(const uint8_t* src, int sh, int sw, const double* minv_dev, uint8_t* dst, int H, int W, cudaStream_t s);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace dcu
