// Internal declarations shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/playaid_b200.h"

namespace pa {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- preprocess
struct PPParams {
    const uint8_t* frames;
    int64_t frames_bytes;
    int n_frames, H, W;
    int64_t pitch, fstride;
    const int32_t* boxes;
    int n_crops, out, padding, swap_rb;
    float mean[3], stdv[3];
    void* outp;
    int out_dtype, out_layout;
    int out_f16, out_split, out_raw;  // 16-bit outputs: IEEE half (else bfloat16) / lo plane present / byte value itself (no /255)
    int64_t plane_elems;  // distance between the hi and lo planes (PA_DTYPE_BF16X2)
    int32_t* status;
    int smem_bytes;
    int first_pass_smem;  // > 0 in the second pass: skip slabs that fit in this many bytes
    int defer_too_large;  // first pass: leave slabs that do not fit to the second pass
    int overlap_prev;     // first pass: the kernel in front of it is the tensor-core kernel (disjoint crops): launch without waiting for it
    int* deferred;        // device counter of slabs the first pass left for the second
    uint8_t* geoms;       // per-crop geometry written by preprocess_plan_kernel (or nullptr)
    int* tables;          // per-crop coefficient tables, table_stride int32 per crop (or nullptr)
    int table_stride;
    int tc_enable;        // route eligible crops to the tensor-core kernel (frames in 16-byte aligned device memory)
    int2* tc_items;       // work items of the tensor-core kernel: {crop, strip | part << 16}
    int* tc_counters;     // [0] items enqueued by the plan kernel, [1] items taken by the tensor-core kernel, [2] CTAs done, [3] tiles reserved
    uint8_t* tc_tiles;    // pool of vertical coefficient tiles (24 KB each) the plan kernel builds per crop
    int* tc_tile_rec;     // per tile: first raw row of its K window | 32-row K steps << 16
    int tc_pool_tiles;
    int threads;          // CTA size of the main kernel: 256 (3 CTAs/SM) or 384 (2 CTAs/SM)
    int num_sms;          // grid of the large-window pass (persistent CTAs)
    int use_xb;           // general area regime: keep fp32 x-pass rows in a ring instead of re-reading bytes
};
int launch_preprocess(const PPParams& p, cudaStream_t stream);
// numpy window image[cy-half-pad : cy+half+pad, cx-half-pad : cx+half+pad] of fighter.py:333-346, clipped as the
// reference clips it (incl. numpy's negative-stop semantics for boxes entirely above / left of the frame)
__host__ __device__ inline void crop_window(int cx, int cy, int sd, int H, int W, int padding, int& x0, int& y0, int& rw, int& rh) {
    int half = sd / 2;
    y0 = cy - half - padding; if (y0 < 0) y0 = 0;
    int y1 = cy + half + padding; if (y1 > H) y1 = H;
    x0 = cx - half - padding; if (x0 < 0) x0 = 0;
    int x1 = cx + half + padding; if (x1 > W) x1 = W;
    if (y1 < 0) { y1 += H; if (y1 < 0) y1 = 0; }  // numpy negative-stop semantics
    if (x1 < 0) { x1 += W; if (x1 < 0) x1 = 0; }
    if (y0 > H) y0 = H;
    if (x0 > W) x0 = W;
    rh = y1 - y0; rw = x1 - x0;
    if (rh < 0) rh = 0;
    if (rw < 0) rw = 0;
}

struct StageParams {
    const uint8_t* src;     // pinned host frames (device-visible)
    uint8_t* dst;           // device frame batch with the same pitch / frame stride
    int64_t frames_bytes;
    int n_frames, H, W;
    int64_t pitch, fstride;
    const int32_t* boxes;   // crop records; record.frame - frame_base indexes the batch
    int n_crops, padding, frame_base;
    int max_sm;             // workers only on SMs with %smid below this
    int* sched;             // [PA_STAGE_SCHED_INTS] device scratch: work-item counter + workers per SM
};
constexpr int PA_STAGE_SCHED_INTS = 1 + 256;
int launch_stage_windows(const StageParams& p, int num_sms, cudaStream_t stream);
int launch_preprocess_plan(const PPParams& p, cudaStream_t stream);
// tensor-core kernel over the work items the plan kernel enqueued; frames_map: u8 tensor map {W*3 bytes, H rows, frames},
// box {128 bytes, 128 rows, 1}, SWIZZLE_128B
int launch_preprocess_tc(const PPParams& p, const CUtensorMap& frames_map, int num_sms, cudaStream_t stream);
constexpr int PA_TC_ITEMS_PER_CROP = 512;   // 128 strips x 4 parts at most
constexpr int PA_TC_TILES_PER_CROP = 12;    // tile pool size per crop of scratch capacity (a 380-px fighter crop takes 6)
constexpr int PA_TC_TILE_BYTES = 192 * 128;
size_t preprocess_geom_bytes();

// ---------------------------------------------------------------- implicit-GEMM convolution (tcgen05 + TMA)
// One launch = one convolution / linear layer over all crops, as D[M, Cout] = A[M, K] * W[Cout, K]^T
// with M = output pixels (NHWC, all crops), K = taps * Cin.
struct ConvMaps {
    CUtensorMap a[2][4];  // [plane hi/lo][input parity for stride 2, index 0 for stride 1]
    CUtensorMap b[2];     // [plane hi/lo] weights [Cout][K] K-major
    CUtensorMap bh[2];    // the same with half-tile boxes (BLOCK_N / 2 rows): what one CTA of a pair stages
};

struct ConvArgs {
    int m_total;          // valid output rows (pixels)
    int m_tiles, n_tiles;
    int cout;             // true output channels (row pitch of the output, in elements)
    int taps_h, taps_w, stride, pad;
    int kb_per_tap;       // 64-channel blocks per tap
    int k_per_tap;        // Cin (K offset between taps in the weight matrix)
    int ho, wo;           // output spatial size
    int num_stages;
    const float* scale;   // [cout] or nullptr (== 1)
    const float* shift;   // [cout] or nullptr (== 0)
    const bf16* res_hi;   // residual [M][cout] or nullptr
    const bf16* res_lo;   // residual lo plane or nullptr
    int relu;
    bf16* out_hi;         // bf16 output [M][cout] or nullptr
    bf16* out_lo;         // lo plane (split precision) or nullptr
    float* out_f32;       // fp32 output [M][cout] or nullptr
    int f16;              // 16-bit operand format: 0 bf16, 1 IEEE half
    int n_b;              // weight planes (2: hi + lo, the f16x3 / bf16x3 modes)
    int multipass;        // conv_gemm: split precision as separate passes over K (residual products first) instead of interleaved
    int debug;            // PA_CONV_DEBUG experiments (results are wrong): 1 = epilogue only drains barriers, 2 = one TMA patch per tile
};

int launch_conv_gemm(const ConvMaps& maps, const ConvArgs& args, int block_n, int n_a, int n_b, int num_sms,
                     cudaStream_t stream);
size_t conv_gemm_smem_bytes(int block_n, int n_a, int n_b, int num_stages, bool multipass);
// 3x3 / stride-1 layers on full-width tiles: patch staging (conv_patch.cu)
int conv_patch_plan(int block_n, int n_a, int wo, int ht, int kb, bool* wres_out, size_t* smem_out);
int launch_conv_patch(const ConvMaps& maps, const ConvArgs& args, int block_n, int n_a, int ht, bool wres, size_t smem,
                      int num_sms, cudaStream_t stream);
int conv_patch2_plan(int block_n, int n_a, int n_b, int wo, int ht, int kb, bool* wres_out, size_t* smem_out);
int launch_conv_patch2(const ConvMaps& maps, const ConvArgs& args, int block_n, int n_a, int ht, bool wres, size_t smem,
                       int num_sms, cudaStream_t stream);
int conv_gemm_pick_stages(int block_n, int n_a, int n_b, bool multipass);
// CTA-pair (cta_group::2) variant for BLOCK_N = 256: maps.b[1] must be the weight map with a BLOCK_N/2-row box
int conv_gemm2_pick_stages(int block_n, int n_a);
int launch_conv_gemm2(const ConvMaps& maps, const ConvArgs& args, int block_n, int n_a, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- conv1 (7x7 stride 2, TMA im2col)
struct Conv1Maps {
    CUtensorMap a[2][2];  // [plane hi/lo][input-row parity]: [32 elems, 64 ox (16-B stride), 64 row pairs, n]
    CUtensorMap b[2];     // [plane hi/lo] packed weights [64][256]: k = ky*32 + (kx+1)*4 + c
};
struct Conv1Args {
    const float* scale;   // [64]
    const float* shift;   // [64]
    bf16* out_hi;         // [n][32][32][64]: conv + BN + ReLU + 3x3/s2 max-pool
    bf16* out_lo;         // or nullptr
    int n_crops;
    int f16;              // 0 bf16, 1 IEEE half
    int split_a, split_w; // activation / weight lo planes present
};
int launch_conv1(const Conv1Maps& maps, const Conv1Args& a, int num_sms, cudaStream_t stream);

// ---------------------------------------------------------------- small memory-bound kernels
int launch_avgpool(const bf16* in_hi, const bf16* in_lo, bf16* out_hi, bf16* out_lo, int n, int hw, int c, int f16,
                   cudaStream_t stream);
int launch_split_f32(const float* in, bf16* out_hi, bf16* out_lo, int64_t n, int f16, cudaStream_t stream);

struct HeadArgs {
    const float* proj;     // [n_feat][seq*512] per-frame temporal projections
    const int32_t* feat_status;  // [n_feat] per-crop status (PA_CROP_*) or nullptr
    const int32_t* win_idx;  // [n_win][seq]
    int n_win, n_feat, seq, n_actions;
    const float* b1d;      // [512]
    const float* w1t;      // [512][128] (Linear(512,128) transposed)
    const float* b1;       // [128]
    const float* w2t;      // [128][n_actions]
    const float* b2;       // [n_actions]
    float* logp;           // [n_win][n_actions]
    int32_t* label;        // [n_win]
    float* conf;           // [n_win]
};
int launch_head(const HeadArgs& a, cudaStream_t stream);
int launch_boxes(const double* rec, int n, int W, int H, double* boxes, int32_t* crops, cudaStream_t stream);

// ---------------------------------------------------------------- ResFormer encoder pieces (transformer_kernels.cu)
int launch_tokens(const float* ffn, const float* enc, float* x, int T, int S, int hidden, cudaStream_t stream);
int launch_attention(const float* qkv, bf16* out_hi, bf16* out_lo, int B, int S, int f16, cudaStream_t stream);
int launch_add_layernorm(float* x, const float* y, const float* gamma, const float* beta, float eps, bf16* out_hi, bf16* out_lo,
                         int T, int f16, cudaStream_t stream);
int launch_logsoftmax(float* logits, int T, int A, cudaStream_t stream);

}  // namespace pa
