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
  int n_valid;            // real output channels (the weight block is zero-padded to NT rows)
  float wscale_inv;       // tcgen05 path: 2^-s, undoes the power-of-two weight scaling of the fp16 split (1.0 otherwise)
  unsigned long long* stats;   // optional [8] cycle counters for the tcgen05 kernel's roles (profiling), may be null
};

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
  const float* w;         // [9][64]
  const float* bias; const float* alpha; const float* beta;   // [64]
  int n, hin, win, hout, wout, pad;
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
  int32_t* counts; int32_t* offsets; int32_t* total; int32_t* kpts; float* patches;
  int max_patches;
  unsigned long long* scan_state;   // [max_batch] chained-scan cells
  unsigned int epoch;
};
void launch_decode_gather(const DecodeParams& p, cudaStream_t s);
size_t decode_smem_bytes(int cells);

void launch_bgr_to_gray(const uint8_t* bgr, uint8_t* gray, long long n_px, cudaStream_t s);   // n_px % 4 == 0
void launch_extract_patches(const float* image, int H, int W, const int32_t* xy, int k, float* patches, cudaStream_t s);

// RefineNet tail: packed arg-max key -> (col,row) and refined (x,y)   (refinenet.py:111-114)
void launch_refine_finalize(const unsigned long long* keys, const int32_t* xy, int xy_stride, int p,
                            int32_t* corners, float* refined, cudaStream_t s);

// layout converters used by the debug/test entry point
void launch_nchw_to_c4(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s);
void launch_c4_to_nchw(const float* in, float* out, int n, int c, int h, int w, cudaStream_t s);
// sub: the NCHW input is [h*sub][w*sub] and pixel (y*sub, x*sub) is taken; rep: the NCHW output is [h*rep][w*rep] (nearest)
void launch_nchw_to_h2(const float* in, void* out, int n, int c, int h, int w, cudaStream_t s, int sub = 1);
void launch_h2_to_nchw(const void* in, float* out, int n, int c, int h, int w, cudaStream_t s, int rep = 1);

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
int tc2_stage_blocks(int up);       // weight blocks per bulk-copy stage (tensor-map box)
// up != 0: p.in is the LOW-resolution tensor (hin x win) whose 2x nearest upsampling is the layer's input; hout = 2*hin
cudaError_t launch_conv_tc2(const ConvParams& p, int n_slices, int up, const void* tmap_a, const void* tmap_w0, const void* tmap_w1,
                            int sm_count, cudaStream_t s);

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace dcu
