// Small fp32 kernels around the tensor-core GEMMs of the ResFormer encoder (reference
// playaid/models/resnet_transformer_detector.py:25-96, torch.nn.TransformerEncoderLayer with batch_first=False):
// token assembly, attention across the windows of a batch, residual + LayerNorm, log-softmax.
// Tokens are stored row-major as t = b * S + s (window b, slot s), d_model = 256 = 8 heads x 32.
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

constexpr int TF_D = 256, TF_H = 8, TF_DH = 32;

// X[t][0:hidden] = ffn[t][:], X[t][hidden:256] = enc[s][:]   (:77-85 of the reference module)
__global__ void tokens_kernel(const float* __restrict__ ffn, const float* __restrict__ enc, float* __restrict__ x, int T, int S,
                              int hidden) {
    const int64_t total = (int64_t)T * TF_D;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i / TF_D), c = (int)(i % TF_D);
        x[i] = c < hidden ? ffn[(int64_t)t * hidden + c] : enc[(t % S) * (TF_D - hidden) + (c - hidden)];
    }
}

// Scaled dot-product attention with the reference's axis convention: for slot s and head h, the B windows of the
// batch attend to each other. qkv [T][768] fp32 (q | k | v, each 8 heads x 32); out 16-bit hi (+lo) [T][256].
// One thread per query; keys/values of the (slot, head) pair stream through shared memory in tiles of 64; online softmax.
template <bool F16>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, bf16* __restrict__ out_hi,
                                                        bf16* __restrict__ out_lo, int B, int S) {
    __shared__ float ks[64][TF_DH + 1];
    __shared__ float vs[64][TF_DH + 1];
    const int s = blockIdx.x / TF_H, h = blockIdx.x % TF_H;
    const int l = blockIdx.y * blockDim.x + threadIdx.x;   // query window
    const bool active = l < B;
    float q[TF_DH], acc[TF_DH];
    const float scale = rsqrtf((float)TF_DH);
    if (active) {
        const float* qp = qkv + ((int64_t)l * S + s) * (3 * TF_D) + h * TF_DH;
#pragma unroll
        for (int d = 0; d < TF_DH; d++) { q[d] = qp[d] * scale; acc[d] = 0.f; }
    }
    float mx = -INFINITY, sum = 0.f;
    for (int m0 = 0; m0 < B; m0 += 64) {
        const int nk = min(64, B - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < nk * TF_DH; i += blockDim.x) {
            const int m = i / TF_DH, d = i % TF_DH;
            const float* base = qkv + ((int64_t)(m0 + m) * S + s) * (3 * TF_D) + h * TF_DH + d;
            ks[m][d] = base[TF_D];
            vs[m][d] = base[2 * TF_D];
        }
        __syncthreads();
        if (active) {
            for (int m = 0; m < nk; m++) {
                float sc = 0.f;
#pragma unroll
                for (int d = 0; d < TF_DH; d++) sc = fmaf(q[d], ks[m][d], sc);
                const float nmx = fmaxf(mx, sc);
                const float corr = __expf(mx - nmx), p = __expf(sc - nmx);
                sum = sum * corr + p;
#pragma unroll
                for (int d = 0; d < TF_DH; d++) acc[d] = fmaf(acc[d], corr, p * vs[m][d]);
                mx = nmx;
            }
        }
    }
    if (active) {
        const float inv = 1.f / sum;
        uint16_t* oh = (uint16_t*)out_hi + ((int64_t)l * S + s) * TF_D + h * TF_DH;
        uint16_t* ol = out_lo ? (uint16_t*)out_lo + ((int64_t)l * S + s) * TF_D + h * TF_DH : nullptr;
#pragma unroll
        for (int d = 0; d < TF_DH; d++) {
            const float v = acc[d] * inv;
            const uint16_t hv = enc16<F16>(v);
            oh[d] = hv;
            if (ol) ol[d] = enc16<F16>(v - dec16<F16>(hv));
        }
    }
}

// x <- LayerNorm(x + y) * gamma + beta (post-norm residual, eps inside the sqrt like torch); also the 16-bit copy the
// next GEMM reads. One warp per token, 8 channels per lane; mean and variance in two passes over registers.
template <bool F16>
__global__ void __launch_bounds__(256) add_layernorm_kernel(float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                            bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int T) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= T) return;
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = lane + 32 * i;
        v[i] = x[(int64_t)warp * TF_D + c] + (y ? y[(int64_t)warp * TF_D + c] : 0.f);
        s += v[i];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float mean = s * (1.f / TF_D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { const float c = v[i] - mean; q = fmaf(c, c, q); }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
    const float rstd = rsqrtf(q * (1.f / TF_D) + eps);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = lane + 32 * i;
        const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
        x[(int64_t)warp * TF_D + c] = o;
        const uint16_t hv = enc16<F16>(o);
        ((uint16_t*)out_hi)[(int64_t)warp * TF_D + c] = hv;
        if (out_lo) ((uint16_t*)out_lo)[(int64_t)warp * TF_D + c] = enc16<F16>(o - dec16<F16>(hv));
    }
}

// in-place log-softmax over the A classes of every token (one warp per token)
__global__ void __launch_bounds__(256) logsoftmax_kernel(float* __restrict__ logits, int T, int A) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= T) return;
    float* p = logits + (int64_t)warp * A;
    float m = -INFINITY;
    for (int i = lane; i < A; i += 32) m = fmaxf(m, p[i]);
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    float se = 0.f;
    for (int i = lane; i < A; i += 32) se += expf(p[i] - m);
    for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
    const float lse = m + logf(se);
    for (int i = lane; i < A; i += 32) p[i] -= lse;
}

static inline int grid_for(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b > 148 * 32) b = 148 * 32;
    return b < 1 ? 1 : (int)b;
}

int launch_tokens(const float* ffn, const float* enc, float* x, int T, int S, int hidden, cudaStream_t stream) {
    if (hidden <= 0 || hidden >= TF_D) return PA_ERR_INVALID_ARG;
    tokens_kernel<<<grid_for((int64_t)T * TF_D, 256), 256, 0, stream>>>(ffn, enc, x, T, S, hidden);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_attention(const float* qkv, bf16* out_hi, bf16* out_lo, int B, int S, int f16, cudaStream_t stream) {
    dim3 grid(S * TF_H, (B + 127) / 128);
    if (f16) attention_kernel<true><<<grid, 128, 0, stream>>>(qkv, out_hi, out_lo, B, S);
    else attention_kernel<false><<<grid, 128, 0, stream>>>(qkv, out_hi, out_lo, B, S);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_add_layernorm(float* x, const float* y, const float* gamma, const float* beta, float eps, bf16* out_hi, bf16* out_lo,
                         int T, int f16, cudaStream_t stream) {
    const int blocks = (T + 7) / 8;
    if (f16) add_layernorm_kernel<true><<<blocks, 256, 0, stream>>>(x, y, gamma, beta, eps, out_hi, out_lo, T);
    else add_layernorm_kernel<false><<<blocks, 256, 0, stream>>>(x, y, gamma, beta, eps, out_hi, out_lo, T);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_logsoftmax(float* logits, int T, int A, cudaStream_t stream) {
    logsoftmax_kernel<<<(T + 7) / 8, 256, 0, stream>>>(logits, T, A);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

}  // namespace pa
