/*
 * playaid_b200.h -- C ABI of the B200-native fighter action-recognition hot path.
 *
 * The reference (NathanBWaters/playaid_core) is pure Python and has no FFI; its boundary for
 * this path is a set of Python call sites. Each entry point below names the reference code it
 * replaces (paths relative to the reference root). INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions: every function returns an int status (PA_OK == 0, negative == error, never
 * throws / exits); pointers are caller-owned DEVICE pointers unless marked "host"; every launch
 * takes an explicit CUDA stream (cudaStream_t passed as void*). Device allocations: weights in pa_model_finalize();
 * pa_preprocess / pa_stage_windows keep their scratch (per-crop geometry, coefficient tables, the pool of vertical
 * coefficient tiles, counters: 0.46 MB per crop of capacity) PER STREAM,
 * sized for 1024 crops on a stream's first call and re-allocated (cudaMalloc + cudaFree: a device synchronisation) only
 * when a call brings more crops than any earlier call on that stream -- calls on different streams or from different
 * threads never share scratch. Nothing else allocates (TMA descriptors are cached in host memory).
 */
#ifndef PLAYAID_B200_H
#define PLAYAID_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_ABI_VERSION 2

/* status codes */
#define PA_OK 0
#define PA_ERR_INVALID_ARG (-1)
#define PA_ERR_CUDA (-2)
#define PA_ERR_UNSUPPORTED (-3)
#define PA_ERR_NOT_READY (-4)
#define PA_ERR_WORKSPACE (-5)
#define PA_ERR_MISSING_TENSOR (-6)

/* per-crop status written by pa_preprocess (mirrors YoloCrop.square_crop's outcomes) */
#define PA_CROP_OK 1        /* reference returns (True, crop)                                    */
#define PA_CROP_INVALID 0   /* reference returns (False, None): empty window / PIL ValueError    */
#define PA_CROP_ZERO_DIV (-2) /* reference lets ZeroDivisionError escape (square_dim == 0)       */
#define PA_CROP_TOO_LARGE (-7) /* window wider than PA_MAX_WINDOW pixels: not computed            */

/* element types / layouts */
#define PA_DTYPE_U8 0
#define PA_DTYPE_BF16 1
#define PA_DTYPE_F32 2
#define PA_DTYPE_BF16X2 3 /* bf16 hi plane followed by a bf16 lo (rounding residual) plane */
#define PA_DTYPE_F16 4    /* IEEE half */
#define PA_DTYPE_F16X2 5  /* half hi plane followed by a half lo (rounding residual) plane */
#define PA_DTYPE_BF16_U8 6 /* the resampled BYTE VALUE 0..255 itself as bfloat16 (exact): no /255, mean/std ignored */
#define PA_DTYPE_F16_U8 7  /* the same as IEEE half (exact); what pa_features_u8 reads: the classifier's stem applies 1/255
                              in its fp32 epilogue, so the 16-bit input carries no rounding error and needs no lo plane */
#define PA_LAYOUT_NHWC 0
#define PA_LAYOUT_NCHW 1
#define PA_LAYOUT_NHWC4 2 /* 4 channels per pixel, channel 3 == 0 */
#define PA_LAYOUT_NHWC4P 3 /* NHWC4 with 4 zero pixels left and right of every row ([n][out][out+8][4]):
                              the conv1-ready layout (TMA im2col needs the horizontal padding in memory) */

/* classifier arithmetic (pa_model_finalize) */
#define PA_PREC_BF16 0   /* bf16 operands, fp32 accumulate: 1 tcgen05 MMA per k-step            */
#define PA_PREC_BF16X2 1 /* activations split hi+lo bf16, weights bf16: 2 MMAs; exact to ~1e-5
                            when the weights are bf16-representable                              */
#define PA_PREC_BF16X3 2 /* weights split as well: 3 MMAs, for arbitrary fp32 checkpoints        */
#define PA_PREC_F16 3    /* IEEE-half operands (11-bit significand), fp32 accumulate: same tensor
                            rate as bf16, 8x smaller rounding error; activations must stay < 65504 */
#define PA_PREC_F16X2 4  /* activations split hi+lo half (~22 bits), weights half: the fp32-parity
                            mode (1e-4) when the weights are half-representable                   */
#define PA_PREC_F16X3 5  /* weights split as well: fp32-parity for arbitrary fp32 checkpoints       */

#define PA_BOX_STRIDE 8 /* int32 per crop record */
/* crop record layout: {frame_index, cx, cy, cw, ch, reserved, reserved, reserved}; (cx,cy,cw,ch)
 * are YoloCrop.yolo_pixels(W, H), i.e. int() truncations (playaid/fighter.py:305-314). */

#define PA_MAX_WINDOW 1920 /* widest raw window (pixels) the preprocess kernel stages */

#define PA_LOG_STRIDE 10 /* doubles per ult_logger record handed to pa_boxes_from_log */

typedef struct pa_ctx pa_ctx;
typedef struct pa_model pa_model;

int pa_abi_version(void);
const char* pa_status_string(int status);
/* last CUDA error text seen by this context (host string, valid until the next call) */
const char* pa_last_error(pa_ctx* ctx);

int pa_ctx_create(int device, pa_ctx** out);
int pa_ctx_destroy(pa_ctx* ctx);

/*
 * Fused crop -> Pillow-BICUBIC letterbox -> OpenCV INTER_AREA -> (127-row letterbox) ->
 * channel swap -> /255 -> (x-mean)/std -> cast. One launch for all crops.
 * Replaces: YoloCrop.square_crop (playaid/fighter.py:323-381) incl. PIL.ImageOps.pad and
 * imutils.resize / cv2.resize(INTER_AREA); cv2.cvtColor(BGR2RGB) + permute + .float()/255
 * (playaid/ult_action_dataset.py:302,349-359; playaid/ai_runner.py:448,461-463); the per-frame
 * loop of playaid/data_gen_scripts/gen_gt_action_detection.py:38-56.
 *
 * frames   u8 [n_frames][H][W][3], row pitch `pitch_bytes`, frame stride `frame_stride_bytes`
 * boxes    int32 [n_crops][PA_BOX_STRIDE]
 * out      [n_crops][out][out][3] (NHWC), [n_crops][3][out][out] (NCHW), [n_crops][out][out][4] (NHWC4) or
 *          [n_crops][out][out+8][4] (NHWC4P, 16-bit only: what pa_features / pa_resformer_forward read) of
 *          out_dtype; the *X2 dtypes write a second ("lo") plane of the same shape behind the first.
 *          PA_DTYPE_U8 ignores mean/std and stores the resampled bytes (square_crop's result).
 *          Crops whose status != PA_CROP_OK are written as zeros.
 * status   int32 [n_crops] or NULL
 * n_crops == 0 is a valid no-op (boxes / out / status may then be NULL); the same holds for pa_stage_windows and, with
 * n == 0, for pa_boxes_from_log.
 */
int pa_preprocess(pa_ctx* ctx, const uint8_t* frames, int n_frames, int H, int W, int64_t pitch_bytes,
                  int64_t frame_stride_bytes, const int32_t* boxes, int n_crops, int out_size, int padding,
                  int swap_rb, const float* mean3_host, const float* std3_host, void* out, int out_dtype,
                  int out_layout, int32_t* status, void* stream);

/*
 * Fighter boxes on the device. Replaces, for n records at once, the bbox part of
 * Fighter.set_from_json (playaid/fighter.py:487-539) with its camera math (calculate_intrinsic_matrix :66-84,
 * calculate_lookat_matrix :87-120, project_point_to_pixel :123-155, np.round half-to-even) and
 * YoloCrop.yolo_pixels' int() truncation (:305-314). fp64, bit-equal to the reference on its own records.
 * log_records  double [n][PA_LOG_STRIDE] = {pos_x, pos_y, camera_position x y z, camera_target_position x y z,
 *              focal length 1280 / (2 tan(fov / 2)) of the stage, frame index of the record}
 * boxes        double [n][4] normalised (cx, cy, w, h) = Fighter.crop.yolo_crop(), or NULL
 * crop_records int32 [n][PA_BOX_STRIDE] ready for pa_preprocess / pa_stage_windows, or NULL
 */
int pa_boxes_from_log(pa_ctx* ctx, const double* log_records, int n, int W, int H, double* boxes, int32_t* crop_records,
                      void* stream);

/*
 * Frames in PINNED HOST memory (what a decoder thread hands over): instead of copying whole frames to
 * the device, pull only the bytes pa_preprocess will read -- the clipped window rows of every crop,
 * widened to 16-byte chunks -- over PCIe into `dev_frames`, a device buffer with the same pitch and
 * frame stride (bytes outside the windows are left untouched and never read). Run it on a side stream
 * so the transfer of one batch overlaps the kernels of the previous one, then hand `dev_frames` to
 * pa_preprocess with the same boxes. record.frame - frame_base indexes the batch, so the match's
 * full record table can be passed without rewriting it per batch. Stands where the reference
 * hands each decoded numpy frame to square_crop (gen_gt_action_detection.py:38-56, ai_runner.py:440-452).
 */
int pa_stage_windows(pa_ctx* ctx, const uint8_t* host_frames, int n_frames, int H, int W, int64_t pitch_bytes,
                     int64_t frame_stride_bytes, const int32_t* boxes, int n_crops, int padding, int frame_base,
                     uint8_t* dev_frames, void* stream);

/*
 * Classifier. Replaces CNNActionDetector / SpatialStreamCNN forward
 * (playaid/models/cnn_action_detector.py:13-43,86-92) and the argmax / exp head of
 * AIRunner.action_recognition (playaid/ai_runner.py:472-477).
 */
int pa_model_create(pa_ctx* ctx, int n_actions, int seq_len, pa_model** out);
int pa_model_destroy(pa_model* m);
/* Hand over one state_dict entry (host fp32, contiguous, PyTorch layout) under the reference's
 * key, e.g. "model.cnn2d.layer1.0.conv1.weight", "model.cnn1d.0.bias", "model.classifier.2.weight".
 * Keys ending in "num_batches_tracked" are ignored. */
int pa_model_set_tensor(pa_model* m, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* Fold eval-mode BatchNorm into fp32 scale/shift, pack the weights for the kernels, upload. */
int pa_model_finalize(pa_model* m, int precision);
int pa_model_precision(const pa_model* m);
/* Workspace needed to push `n_crops` crops through the CNN (activations of every layer). */
int pa_model_workspace_bytes(const pa_model* m, int n_crops, size_t* bytes);

/*
 * ResNet-18 features, once per crop. crops: 16-bit NHWC4P [n_crops][128][136][4] (channel 3 == 0, 4 zero
 * pixels either side of a row) as written by pa_preprocess(out_layout=PA_LAYOUT_NHWC4P) with the
 * out_dtype matching the model's precision; in split precisions the lo plane follows the hi plane.
 * feat: fp32 [n_crops][1000].
 */
int pa_features(pa_model* m, const void* crops, int n_crops, float* feat, void* workspace,
                size_t workspace_bytes, void* stream);
/* Same for crops that hold the raw byte values (pa_preprocess out_dtype PA_DTYPE_F16_U8 / PA_DTYPE_BF16_U8, one plane in
 * every precision): x = v / 255 of ai_runner.py:463 is folded into the stem's fp32 scale, which equals the reference's
 * conv(fl(v / 255)) to fp32 rounding and removes the 16-bit rounding of the input (the largest single error source of
 * the one-product modes) and the stem's second product in the split modes. Only valid for mean 0 / std 1. */
int pa_features_u8(pa_model* m, const void* crops, int n_crops, float* feat, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Elements (of out_dtype) per crop in the internal conv-ready layout. */
size_t pa_crop_elems(int out_size);

/*
 * Temporal head over windows of cached features: Conv1d(1000->512,k=seq)+ReLU, Linear+ReLU,
 * Linear, log_softmax, argmax (first maximum of the log-probs), exp(log-prob of the label).
 * feat      fp32 [n_feat][1000]
 * win_idx   int32 [n_win][seq_len] rows of feat (action_sample_from_frame_middle_out,
 *           playaid/dataset_utils.py:109-138, already offset into feat)
 * feat_status int32 [n_feat] per-crop status as written by pa_preprocess, or NULL. A window with a slot whose
 *           crop is not PA_CROP_OK gets label -1 and conf 0 (its log-probs are still written): the reference never
 *           classifies such a window -- process_pairing skips the crop (gen_gt_action_detection.py:54-56) and
 *           AIRunner asserts on the missing file (ai_runner.py:447)
 * logp      fp32 [n_win][n_actions]; label int32 [n_win]; conf fp32 [n_win] (probability;
 *           AIRunner multiplies by 100.0 on the host)
 */
int pa_head(pa_model* m, const float* feat, int n_feat, const int32_t* feat_status, const int32_t* win_idx, int n_win,
            float* logp, int32_t* label, float* conf, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Second detector: ResnetTransformerDetector / ResFormer, reference
 * playaid/models/resnet_transformer_detector.py:25-143 -- ResNet-50 features (2048) -> Linear(2048, 247) ->
 * + 9 time-encoding features -> three post-norm TransformerEncoder layers (d 256, 8 heads, feed-forward 2048)
 * -> Linear(256, A) -> log_softmax, one output per frame of the window. The handle is a pa_model: tensors are
 * handed over with pa_model_set_tensor under the checkpoint's keys ("model.resnet.conv1.weight",
 * "model.transformer.layers.0.self_attn.in_proj_weight", "model.freq_encoding", ...) and released with
 * pa_model_destroy. The reference builds the encoder with batch_first=False and feeds it [B,S,256], so attention
 * runs across the B windows of a call, slot by slot; pa_resformer_forward keeps that: its result for a window
 * depends on the other windows of the same call, exactly as the reference's does on its batch.
 * crops   16-bit NHWC4P [n_windows*seq][128][136][4] (pa_preprocess output; lo plane behind the hi plane in split
 *         precisions), window-major: crop index = window * seq + slot
 * logp    fp32 [n_windows][seq][n_actions]
 */
int pa_resformer_create(pa_ctx* ctx, int n_actions, int seq_len, pa_model** out);
int pa_resformer_finalize(pa_model* m, int precision);
int pa_resformer_workspace_bytes(const pa_model* m, int n_windows, size_t* bytes);
int pa_resformer_forward(pa_model* m, const void* crops, int n_windows, float* logp, void* workspace, size_t workspace_bytes,
                         void* stream);

/*
 * Single layers (synchronous; scratch allocated and freed inside): one convolution as a tcgen05
 * implicit GEMM on NHWC bf16 activations [n][hin][hin][cin] (cin % 64 == 0, hin in {32,16,8,4,1} for
 * the output), and the fused stem (7x7/s2 conv + scale/shift + ReLU + 3x3/s2 max-pool) on NHWC4P crops
 * [n][128][136][4] -> [n][32][32][64]. Weights / scale / shift are host fp32 in PyTorch
 * layout; y = conv(x, w) * scale + shift (+ residual) (ReLU). These are the building blocks
 * pa_features sequences; exposed for layer-level parity tests against torch.nn.functional.conv2d.
 * in_lo / out_lo / res_lo may be NULL (no split). split_w is a flag word: bit 0 = split the weights
 * too (3 MMAs), bit 1 = the 16-bit planes are IEEE half instead of bfloat16.
 */
int pa_conv2d(pa_ctx* ctx, const void* in_hi, const void* in_lo, int n, int hin, int cin, const float* w_host, int cout,
              int k, int stride, int pad, const float* scale_host, const float* shift_host, const void* res_hi,
              const void* res_lo, int relu, void* out_hi, void* out_lo, float* out_f32, int split_w, void* stream);
int pa_stem(pa_ctx* ctx, const void* in_hi, const void* in_lo, int n, const float* w_host, const float* scale_host,
            const float* shift_host, void* out_hi, void* out_lo, int split_w, void* stream);

/* Per-kernel device timing: between begin and end every launch of this context is bracketed by a
 * CUDA event pair on its stream; end writes "name<TAB>launches<TAB>total_ms" lines into buf (host). */
int pa_profile_begin(pa_ctx* ctx);
int pa_profile_end(pa_ctx* ctx, char* buf, size_t buflen);

/* Kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t pa_launch_count(pa_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PLAYAID_B200_H */
