/*
 * deepcharuco_b200 -- C ABI of the B200-native ChArUco keypoint inference engine.
 *
 * The reference (JunkyByte/deepcharuco) has no FFI: its boundary is the Python
 * function surface of src/inference.py (load_models :73, infer_image :32,
 * solve_pnp :15).  This header is what a ctypes binding of that surface calls
 * instead of torch.nn modules; every entry point cites the reference code it
 * replaces.  Plain pointers and sizes only -- no torch / C++ types.
 *
 * Conventions
 *   - all `*_dev` pointers are CUDA device pointers on the engine's device;
 *     `*_host` pointers are host memory (pinned or pageable);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous on that stream unless stated otherwise;
 *   - return value: 0 = DCU_OK, negative = error; dcu_last_error() returns a
 *     thread-local human-readable message for the last failing call;
 *   - one engine serves one (device, stream) at a time; calls on one engine are
 *     not re-entrant (the reference is single-threaded too, inference.py:32-70).
 *
 * Tensor layouts at the boundary are the reference's own:
 *   frames  : uint8  [N][H][W]            grayscale (after cv2.cvtColor, inference.py:40)
 *   loc     : float  [N][65][H/8][W/8]    NCHW, dcModel.forward output (net.py:74)
 *   ids     : float  [N][n_ids+1][H/8][W/8]                            (net.py:77)
 *   patches : float  [P][24][24]          extract_patches output (model_utils.py:19-36)
 */
#ifndef DEEPCHARUCO_B200_H
#define DEEPCHARUCO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCU_OK             0
#define DCU_ERR_INVALID   -1   /* bad argument (shape, NULL pointer, N > max_batch, ...) */
#define DCU_ERR_CUDA      -2   /* a CUDA runtime / driver call failed */
#define DCU_ERR_CAPACITY  -3   /* more corners than max_patches; results up to capacity are valid */
#define DCU_ERR_UNSUPPORTED -4 /* e.g. not an sm_100 device */

/* Which convolution implementation the engine uses for the 3x3 layers with Cin >= 64.
 * Both are this library's own sm_100a kernels; neither is a fallback to a library. */
#define DCU_CONV_FFMA   0      /* fp32 CUDA-core direct convolution (bring-up / strict-fp32 path) */
#define DCU_CONV_TCGEN05 1     /* tcgen05.mma kind::f16 on an fp16 hi/lo split of both operands (3 products, 22 bits), fp32 accumulate in TMEM */

typedef struct DcuEngine DcuEngine;

/* One convolution layer as the reference stores it (torch OIHW fp32), plus the eval-mode
 * BatchNorm folded to a per-channel affine that is applied AFTER the fp32 accumulation:
 *     y = relu(fma(acc + bias, alpha, beta))        (alpha == NULL: y = acc + bias, no ReLU)
 * alpha = gamma / sqrt(var + eps), beta = bn_bias - mean * alpha, computed by the caller in fp32
 * (net.py:60 `relu(bn(conv(x)))`, BatchNorm2d eps 1e-5).  All pointers are HOST pointers and are
 * copied during dcu_create. */
typedef struct DcuConvLayer {
  const float* weight;   /* [cout][cin][k][k] */
  const float* bias;     /* [cout] */
  const float* alpha;    /* [cout] or NULL */
  const float* beta;     /* [cout] or NULL */
  int32_t cin, cout, ksize;
} DcuConvLayer;

typedef struct DcuConfig {
  int32_t device;        /* CUDA device ordinal */
  int32_t height, width; /* frame size, multiples of 8 (three 2x2 pools, net.py:62,65,68) */
  int32_t n_ids;         /* board inner corners; ids head has n_ids+1 channels (net.py:48) */
  int32_t max_batch;     /* frames per dcu_infer_* call (workspace is sized for this) */
  int32_t max_patches;   /* corner capacity per call, summed over the batch */
  int32_t conv_impl;     /* DCU_CONV_FFMA or DCU_CONV_TCGEN05 */
  int32_t reserved;      /* flags: DCU_FLAG_* (0 = a full engine) */
} DcuConfig;

/* DcuConfig.reserved flag: an engine without networks or convolution workspace, for the stand-alone helpers
 * pred_to_keypoints / extract_patches (model_utils.py:81-88, :19-36 -> dcu_decode_gather / dcu_extract_patches on
 * caller-owned buffers).  det_layers / ref_layers may be NULL; every entry point that needs a network returns
 * DCU_ERR_INVALID. */
#define DCU_FLAG_DECODE_ONLY 1

/* Replaces inference.load_models (inference.py:73-84): builds packed device weights + workspace.
 * det_layers: 12 layers in net.py:22-48 order  (conv1a,1b,2a,2b,3a,3b,4a,4b,Pa,Pb,Da,Db);
 * ref_layers: 12 layers in refinenet.py:22-47 order (conv1a,1b,2a,2b,3a,3b,4a,4b,5a,5b,Pa,Pb),
 *             or NULL / n_ref == 0 for a detector-only engine (refinenet_ckpt=None). */
int dcu_create(const DcuConfig* cfg, const DcuConvLayer* det_layers, int n_det,
               const DcuConvLayer* ref_layers, int n_ref, DcuEngine** out);
int dcu_destroy(DcuEngine* e);

/* Replaces pre_bgr_image + dcModel.forward (model_utils.py:46-50, net.py:50-80) for N frames.
 * frames_dev uint8 [N][H][W] -> loc_dev [N][65][H/8][W/8], ids_dev [N][n_ids+1][H/8][W/8]. */
int dcu_detector_forward(DcuEngine* e, const uint8_t* frames_dev, int n,
                         float* loc_dev, float* ids_dev, void* stream);

/* Same from an already normalised fp32 image batch [N][H][W] -- the argument lModel.infer_image takes
 * (net.py:127-128: `deepc.infer_image(img_gray)` with img_gray = (x-128)/255). */
int dcu_detector_forward_f32(DcuEngine* e, const float* images_dev, int n,
                             float* loc_dev, float* ids_dev, void* stream);

/* Replaces extract_patches (model_utils.py:19-36) on its own: image_dev fp32 [H][W] (normalised),
 * xy_dev [K][2] int32 -> patches_dev [K][24][24]; zeros outside the image. */
int dcu_extract_patches(DcuEngine* e, const float* image_dev, const int32_t* xy_dev, int k,
                        float* patches_dev, void* stream);

/* Replaces pred_to_keypoints + extract_patches (model_utils.py:53-124, :19-36), batched, one kernel.
 * Per frame f: counts_dev[f] = K_f and offsets_dev[f] = sum of K over earlier frames (assigned by
 * frame index, deterministic).  Frame f owns rows [offsets_dev[f], +K_f) of
 *     kpts_dev    [max_patches][4] int32 {x, y, id, cell}
 *     patches_dev [max_patches][24][24] float   (may be NULL: raw decode only, no gather)
 * sorted by (id, cell) -- the order inference.py:68-69 returns (stable sort by id of the row-major
 * list; sort a frame's rows by `cell` to recover pred_to_keypoints order).  total_dev[0] = sum K_f;
 * rows beyond max_patches are dropped and the caller sees total > max_patches (DCU_ERR_CAPACITY in
 * the *_host entry point).  `append` != 0 continues numbering from the current *total_dev instead
 * of 0 (used to decode a large batch in micro-batches). */
int dcu_decode_gather(DcuEngine* e, const float* loc_dev, const float* ids_dev,
                      const uint8_t* frames_dev, int n, int dust_bin_ids, int append,
                      int32_t* counts_dev, int32_t* offsets_dev, int32_t* total_dev,
                      int32_t* kpts_dev, float* patches_dev, void* stream);

/* Replaces RefineNet.infer_patches (refinenet.py:85-115): patches [P][24][24] + integer keypoints
 * xy_dev [P][xy_stride] int32 (x at +0, y at +1; xy_stride = 2 for a packed (K,2) array, 4 to pass
 * kpts_dev rows directly)
 * -> corners_dev [P][2] int32 (col, row) of the 64x64 arg-max (speedy_bargmax2d, model_utils.py:39-43)
 *    refined_dev [P][2] float = (corners - 32) / 8 + (x, y).
 * heat_dev (may be NULL) receives the [P][64][64] heat map, for stage-level tests only. */
int dcu_refine_forward(DcuEngine* e, const float* patches_dev, const int32_t* xy_dev, int xy_stride, int p,
                       int32_t* corners_dev, float* refined_dev, float* heat_dev, void* stream);

/* Replaces the body of inference.infer_image (inference.py:41-60) for a batch resident in HBM:
 * detector -> decode+gather -> RefineNet.  RefineNet is enqueued right behind the decode for a chunk count predicted from
 * recent calls and takes the true patch count from device memory; the call waits (on an event, after the decode) only to
 * learn the count and to launch a chunk the prediction missed -- the reference synchronises at model_utils.py:114.
 * Outputs (device): counts_dev [N], offsets_dev [N], total_dev [1], kpts_dev [max_patches][4] int32
 * {x, y, id, cell} and refined_dev [max_patches][2] float, both indexed by offsets_dev[f] + j.
 * use_refinenet == 0 skips the RefineNet leg (refinenet=None, inference.py:54).
 * More corners than max_patches: rows up to capacity are valid (and refined), counts / offsets / total describe the full set;
 * with use_refinenet the call returns DCU_ERR_CAPACITY (it knows the count), without it the caller compares *total_dev with
 * max_patches.  dcu_dc_metrics and dcu_solve_pnp_batch never read rows beyond max_patches. */
int dcu_infer_batch(DcuEngine* e, const uint8_t* frames_dev, int n, int dust_bin_ids, int use_refinenet,
                    int32_t* counts_dev, int32_t* offsets_dev, int32_t* total_dev,
                    int32_t* kpts_dev, float* refined_dev, void* stream);

/* Same, end to end from HOST memory: H2D of the u8 frames, the pipeline, D2H of the packed result,
 * stream-synchronised on return.  counts_host [N], offsets_host [N], kpts_host [max_patches][4],
 * refined_host [max_patches][2]; *total_host = sum of counts.  Uses the engine's pinned staging. */
int dcu_infer_batch_host(DcuEngine* e, const uint8_t* frames_host, int n, int dust_bin_ids, int use_refinenet,
                         int32_t* counts_host, int32_t* offsets_host, int32_t* total_host,
                         int32_t* kpts_host, float* refined_host, void* stream);

/* Replaces cv2.cvtColor(img, COLOR_BGR2GRAY) (inference.py:40) on the device: bgr_dev uint8 [N][H][W][3] -> gray_dev uint8
 * [N][H][W], OpenCV's 8-bit fixed-point luma, bit-exact.  (SURVEY.md 8f "next" row 1.) */
int dcu_bgr_to_gray(DcuEngine* e, const uint8_t* bgr_dev, int n, uint8_t* gray_dev, void* stream);

/* Replaces cv2.resize(img, (W, H), cv2.INTER_LINEAR) (the reference's evaluation loop resizes the camera frame to the network input,
 * inference.py:131-132) on the device for uint8 images with 1 or 3 interleaved channels: src_dev [N][src_h][src_w][C] ->
 * dst_dev [N][H][W][C] with H, W the engine's frame size.  Bit-exact with OpenCV's fixed-point path when SHRINKING (src >= dst in both
 * dimensions, the camera-to-network direction); enlarging returns DCU_ERR_UNSUPPORTED (OpenCV's vector path rounds 0.1 % of the pixels
 * differently there).  Together with dcu_bgr_to_gray this is the whole input stage of SURVEY.md 8f row 1. */
int dcu_resize_u8(DcuEngine* e, const uint8_t* src_dev, int n, int src_h, int src_w, int channels, uint8_t* dst_dev, void* stream);

/* dcu_infer_batch_host for BGR frames: frames_host uint8 [N][H][W][3]; the colour conversion runs on the device too. */
int dcu_infer_batch_host_bgr(DcuEngine* e, const uint8_t* frames_host, int n, int dust_bin_ids, int use_refinenet,
                             int32_t* counts_host, int32_t* offsets_host, int32_t* total_host,
                             int32_t* kpts_host, float* refined_host, void* stream);

/* Test hook: run ONE 3x3 conv(+BN+ReLU[+pool][+2x nearest up]) layer of either network on fp32 NCHW
 * device tensors with the selected implementation (DCU_CONV_*), so each layer is checkable against
 * the oracle in isolation.  net: 0 detector, 1 RefineNet; layer: index into the dcu_create tables.
 * in_dev [n][cin][h][w] -> out_dev [n][cout][h'][w'] (pad/pool/upsample as the reference applies
 * to that layer). */
int dcu_debug_conv_layer(DcuEngine* e, int net, int layer, int conv_impl, const float* in_dev,
                         int n, int h, int w, float* out_dev, void* stream);

/* Profiling hook for the tcgen05 kernel: enable != 0 zeroes and arms eight device cycle counters that the kernel's
 * roles add to (MMA warp: total / wait halo / wait weights / wait epilogue; splitter: total / wait TMA; epilogue:
 * total / wait MMA), summed over CTAs; out8 (may be NULL) receives the counters accumulated so far. */
int dcu_debug_tc_stats(DcuEngine* e, int enable, uint64_t* out8);

/* Batched board pose: inference.solve_pnp (inference.py:15-29; called per frame by pose_estimation.py:61-63) for all frames
 * of a batch in one launch, directly on the engine's result buffers (dcu_infer_batch outputs): per frame f the corners
 * kpts[offsets[f] .. offsets[f]+counts[f]) with image points = refined (x, y) (or the integer pixels when refined_dev is
 * NULL) and object points = the board's inner-corner table the reference builds from (col_count, row_count, square_len).
 * Restates cv2.solvePnP(SOLVEPNP_ITERATIVE) for coplanar points in fp64 (homography start + Levenberg-Marquardt on the
 * pixel reprojection error, OpenCV's distortion model with up to 8 coefficients k1 k2 p1 p2 k3 k4 k5 k6).
 * ret[f] = 1 on success, 0 for frames with < 4 corners (the reference returns (False, None, None)) or a degenerate /
 * out-of-range configuration; rvec / tvec are [n][3] doubles (zeros where ret == 0).  camera_matrix9 (row-major 3x3) and
 * dist_coeffs are HOST pointers.  The _host variant takes and returns host arrays (kpts_host packed [sum counts][4]). */
int dcu_solve_pnp_batch(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev,
                        const float* refined_dev, int n, int col_count, int row_count, double square_len,
                        const double* camera_matrix9, const double* dist_coeffs, int n_dist, int32_t* ret_dev,
                        double* rvec_dev, double* tvec_dev, void* stream);
int dcu_solve_pnp_batch_host(DcuEngine* e, const int32_t* counts_host, const int32_t* kpts_host, const float* refined_host, int n,
                             int col_count, int row_count, double square_len, const double* camera_matrix9,
                             const double* dist_coeffs, int n_dist, int32_t* ret_host, double* rvec_host, double* tvec_host,
                             void* stream);

/* Detector validation metric on the device: the per-sample part of DC_Metrics.update (models/metrics.py:48-73, i.e.
 * compute_l2_distance :102-129 and compute_ratio :75-100) on the decode output of dcu_decode_gather / dcu_infer_batch
 * (counts / offsets / kpts) against the label maps loc_target / ids_target [n][H/8][W/8] int64 (the reference dataset's
 * label format, decoded like label_to_keypoints :25-35).  Per sample: l2[f] = mean over matched label ids of the worst
 * prediction-to-label distance in pixels, ratio[f] = share of labels matched within 3 px, valid[f] = 0 when the sample
 * has no labels (the reference skips it).  The caller accumulates like DC_Metrics: distance += sum(l2[valid]) / n. */
int dcu_dc_metrics(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev, int n,
                   const int64_t* loc_target_dev, const int64_t* ids_target_dev, int dust_bin_ids, float* l2_dev,
                   float* ratio_dev, int32_t* valid_dev, void* stream);

/* The reference's pixel-error evaluation (utils.pixel_error, /root/reference/src/utils.py:33-52, used by the up_scale = 8 loop of
 * inference.py:110-171) for every frame of a batch, on the device-resident rows of dcu_infer_batch (kpts_dev {x, y, id, cell},
 * refined_dev {x, y}) against float64 labels target_dev [sum tcounts][3] = {x, y, id}, frame f owning rows
 * [toffsets_dev[f], +tcounts_dev[f]).  out_dev [n][6] = {mean, mean_refined, mean_refined_vs_raw, max, max_refined, max_refined_vs_raw}
 * of compute_l2_distance (utils.py:6-30; float64, bit-identical with numpy incl. its summation order).  status_dev[f] = 1 evaluated;
 * 0 skipped as the reference skips it (no labels / no corners / a predicted id that has no label: returns (None, None)); -1 where
 * numpy's broadcasting would raise (several predictions AND several labels of one id in different numbers), ids outside [0, 64)
 * or more than 256 rows in a frame. */
int dcu_pixel_error(DcuEngine* e, const int32_t* counts_dev, const int32_t* offsets_dev, const int32_t* kpts_dev, const float* refined_dev,
                    int n, const int32_t* tcounts_dev, const int32_t* toffsets_dev, const double* target_dev, int32_t* status_dev,
                    double* out_dev, void* stream);

/* ---- synthetic frames on the device ----
 * What the reference's data path does per training / validation sample on the host (transformations.py:22-52,105-114,
 * custom_aug.PasteBoard: warp the rendered board, paste it on a background, blur, brightness, noise), as one launch per batch.
 * Per-frame parameters come from the host (deepcharuco_b200/synth.py: gpu_frame_params -- homographies and Gaussian taps need
 * float64 trigonometry / a linear solve); the per-pixel work is the kernel's: OpenCV's own fixed-point bilinear warp of the board
 * texture and its mask (bit-exact with cv2.warpPerspective on u8, INTER_LINEAR, BORDER_CONSTANT), a lattice-noise background,
 * a separable 13-tap blur with reflect-101 borders, gain, and Philox4x32-10 noise: frame i depends only on (seed, first_index + i). */
typedef struct DcuSynthFrame {
  int32_t lat_step, lat_h, lat_w;   /* background lattice: spacing in pixels, lattice size (H / step + 3, W / step + 3) */
  int32_t n_boards;                 /* 0..4 boards pasted in order */
  float bg_lo, bg_hi, gain, reserved;
  float blur_w[16];                 /* 13 normalised taps (-6..6), float32 */
  double hinv[4][9];                /* per board: frame pixel -> board pixel homography (row-major 3x3), what cv2.warpPerspective
                                       computes internally as invert(M) */
} DcuSynthFrame;

/* params_host [n] (host memory, copied inside), board_dev uint8 [board_px][board_px] (the rendered board, device),
 * frames_dev uint8 [n][H][W] (device) with H, W the engine's frame size. */
int dcu_synth_frames(DcuEngine* e, const DcuSynthFrame* params_host, int n, uint64_t seed, int first_index,
                     const uint8_t* board_dev, int board_px, uint8_t* frames_dev, void* stream);

/* cv2.warpPerspective(src, M, (dst_w, dst_h), flags=INTER_LINEAR, borderMode=BORDER_CONSTANT, 0) for one uint8 single-channel image
 * with minv9_host = invert(M) (float64, row-major; host memory): src_dev [src_h][src_w] -> dst_dev [dst_h][dst_w], bit-exact. */
int dcu_warp_perspective_u8(DcuEngine* e, const uint8_t* src_dev, int src_h, int src_w, const double* minv9_host,
                            uint8_t* dst_dev, int dst_h, int dst_w, void* stream);

/* RefineNet validation metric on the device: the per-sample part of Refinenet_Metrics.update (models/metrics.py:141-158):
 * dist[i] = L2 distance in heat-map pixels between the arg-max of the predicted 64x64 heat map and the arg-max of the target map
 * (first maximum of the flattened map).  The prediction is either a heat map tensor heat_pred_dev [p][64][64] or, when that is
 * NULL, the (col, row) arg-maxes corners_pred_dev [p][2] that dcu_refine_forward already produced.  The caller accumulates
 * distance += mean(dist) like the reference. */
int dcu_refinenet_metrics(DcuEngine* e, const float* heat_pred_dev, const int32_t* corners_pred_dev, const float* heat_target_dev, int p,
                          float* dist_dev, void* stream);

/* Select the 3x3 conv implementation after creation (DCU_CONV_*). */
int dcu_set_conv_impl(DcuEngine* e, int conv_impl);

/* Number of this library's kernels launched on behalf of `e` since creation (bench.py: gpu_launches). */
int64_t dcu_launch_count(const DcuEngine* e);

/* Per-kernel timing for bench.py's roofline figure.  While enabled, every launch made on behalf of `e`
 * is bracketed by CUDA events on the launching stream.  dcu_profile_read synchronises, sums the event
 * durations of kernel class `cls` since the last enable (0 = 3x3 conv [the dominant kernel], 1 = first-layer
 * conv, 2 = 1x1 heads, 3 = decode+gather, 4 = RefineNet finalize) and returns them with the class's
 * ALGORITHMIC work: flops (2*MAC) for classes 0-2, bytes for class 3. */
int dcu_profile_enable(DcuEngine* e, int on);
int dcu_profile_read(DcuEngine* e, int cls, double* total_ms, double* total_work, int64_t* n_launches);
/* Class 0 on the CTA-pair tcgen05 kernel: the flops the tensor pipes actually EXECUTE for the same launches (2 x MACs of all
 * issued MMAs: three fp16 products per MAC, 4 of 9 taps on upsample-fused layers, padded / wrap-around tile rows included);
 * 0 for launches of other kernels. */
int dcu_profile_read_issued(DcuEngine* e, int cls, double* issued_flops);
/* The individual records since the last enable, in launch order: rec8 [cap][8] doubles
 * {cls, ms, work, cin, cout, hout, wout, n}; *n_records = how many exist (may exceed cap). */
int dcu_profile_records(DcuEngine* e, int cap, double* rec8, int* n_records);

/* Algorithmic FLOPs (2*MAC) per frame of the detector and per patch of RefineNet for this engine's
 * shapes (12.879 GFLOP / 320x240 frame, 0.8711 GFLOP / patch). */
double dcu_detector_flops_per_frame(const DcuEngine* e);
double dcu_refine_flops_per_patch(const DcuEngine* e);

const char* dcu_last_error(void);
const char* dcu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DEEPCHARUCO_B200_H */
