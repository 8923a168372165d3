// Memory-bound helpers of the classifier: global average pool, fp32 -> 16-bit hi/lo split, and the temporal head (window gather + MLP + log-softmax / argmax / confidence).
//
// (the 3x3/s2 max-pool is fused into the stem kernel, conv1.cu)
// References: torchvision resnet18.avgpool (via playaid/models/cnn_action_detector.py:16,32);
// SpatialStreamCNN.cnn1d + classifier (:22-27,37-41); CNNActionDetector.forward log_softmax (:92);
// AIRunner.action_recognition argmax / exp (playaid/ai_runner.py:474-477).
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

// ---------------------------------------------------------------- global average pool [n][hw][c] -> [n][c]
template <bool F16>
__global__ void avgpool_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, bf16* __restrict__ out_hi,
                               bf16* __restrict__ out_lo, int n, int hw, int c) {
    const int64_t total = (int64_t)n * c;
    const float inv = 1.f / (float)hw;
    if (threadIdx.x == 0) pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const int64_t b = i / c;
        float s = 0.f;
        for (int p = 0; p < hw; p++) {
            const int64_t off = (b * hw + p) * c + ch;
            float v = dec16<F16>(((const uint16_t*)in_hi)[off]);
            if (in_lo) v += dec16<F16>(((const uint16_t*)in_lo)[off]);
            s += v;
        }
        s *= inv;
        const uint16_t h = enc16<F16>(s);
        ((uint16_t*)out_hi)[i] = h;
        if (out_lo) ((uint16_t*)out_lo)[i] = enc16<F16>(s - dec16<F16>(h));
    }
}

int launch_avgpool(const bf16* in_hi, const bf16* in_lo, bf16* out_hi, bf16* out_lo, int n, int hw, int c, int f16,
                   cudaStream_t stream) {
    const int64_t total = (int64_t)n * c;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cudaError_t e = f16 ? launch_pdl(avgpool_kernel<true>, dim3(blocks), dim3(256), 0, stream, in_hi, in_lo, out_hi, out_lo, n, hw, c)
                        : launch_pdl(avgpool_kernel<false>, dim3(blocks), dim3(256), 0, stream, in_hi, in_lo, out_hi, out_lo, n, hw, c);
    return e == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

// ---------------------------------------------------------------- fp32 -> bf16 hi (+ lo)
template <bool F16>
__global__ void split_kernel(const float* __restrict__ in, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int64_t n) {
    if (threadIdx.x == 0) pdl_launch_dependents();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = in[i];
        const uint16_t h = enc16<F16>(v);
        ((uint16_t*)out_hi)[i] = h;
        if (out_lo) ((uint16_t*)out_lo)[i] = enc16<F16>(v - dec16<F16>(h));
    }
}

int launch_split_f32(const float* in, bf16* out_hi, bf16* out_lo, int64_t n, int f16, cudaStream_t stream) {
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    cudaError_t e = f16 ? launch_pdl(split_kernel<true>, dim3(blocks), dim3(256), 0, stream, in, out_hi, out_lo, n)
                        : launch_pdl(split_kernel<false>, dim3(blocks), dim3(256), 0, stream, in, out_hi, out_lo, n);
    return e == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

// ---------------------------------------------------------------- temporal head, one CTA (512 threads) per window
// proj[f][t*512 + o] = sum_i W1d[o][i][t] * feat[f][i]; the Conv1d over a window is the sum over its
// 7 slots of the matching projection rows (+ bias), so no window tensor is ever built.
// The MLP rows are split four ways over the 512 threads (partial sums combined through shared memory) and the
// loops are unrolled 16-fold: the weights come from L2, and a single 512-long dependent chain per thread left the
// kernel waiting on one load at a time.
__global__ void __launch_bounds__(512) head_kernel(const HeadArgs a) {
    __shared__ float h[512];
    __shared__ float part[4][128];
    __shared__ float y1[128];
    __shared__ float y2[128];
    __shared__ int win_ok;
    const int w = blockIdx.x, t = threadIdx.x;
    const int pitch = a.seq * 512;
    if (t == 0) pdl_launch_dependents();
    pdl_wait();
    if (t == 0) {   // a window is only classified when every crop it reads exists (see pa_head)
        int ok = 1;
        if (a.feat_status)
            for (int k = 0; k < a.seq; k++) {
                int f = a.win_idx[(int64_t)w * a.seq + k];
                f = f < 0 ? 0 : (f >= a.n_feat ? a.n_feat - 1 : f);
                if (a.feat_status[f] != PA_CROP_OK) ok = 0;
            }
        win_ok = ok;
    }
    {
        const int o = t;
        float s = a.b1d[o];
        for (int k = 0; k < a.seq; k++) {
            int f = a.win_idx[(int64_t)w * a.seq + k];
            f = f < 0 ? 0 : (f >= a.n_feat ? a.n_feat - 1 : f);
            s += a.proj[(int64_t)f * pitch + k * 512 + o];
        }
        h[o] = fmaxf(s, 0.f);
    }
    __syncthreads();
    {
        const int j = t & 127, p = t >> 7;
        const float* wp = a.w1t + (size_t)(p * 128) * 128 + j;
        const float* hp = h + p * 128;
        float s = 0.f;
#pragma unroll 16
        for (int i = 0; i < 128; i++) s = fmaf(__ldg(wp + i * 128), hp[i], s);
        part[p][j] = s;
    }
    __syncthreads();
    if (t < 128) y1[t] = fmaxf(a.b1[t] + ((part[0][t] + part[1][t]) + (part[2][t] + part[3][t])), 0.f);
    __syncthreads();
    if (t < a.n_actions) {
        float s = a.b2[t];
#pragma unroll 16
        for (int i = 0; i < 128; i++) s = fmaf(__ldg(a.w2t + i * a.n_actions + t), y1[i], s);
        y2[t] = s;
    }
    __syncthreads();
    if (t < 32) {
        // log-softmax, then argmax over the log-probabilities themselves (first maximum, like
        // torch.argmax on the reference's output) -- all with warp shuffles (n_actions <= 128)
        float m = -INFINITY;
        for (int i = t; i < a.n_actions; i += 32) m = fmaxf(m, y2[i]);
        for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
        float se = 0.f;
        for (int i = t; i < a.n_actions; i += 32) se += expf(y2[i] - m);
        for (int d = 16; d > 0; d >>= 1) se += __shfl_xor_sync(0xffffffffu, se, d);
        const float lse = m + logf(se);
        float best = -INFINITY;
        int bi = 0x7FFFFFFF;   // stays there only when every log-prob is NaN: reported as class 0 below
        for (int i = t; i < a.n_actions; i += 32) {
            const float lp = y2[i] - lse;
            a.logp[(int64_t)w * a.n_actions + i] = lp;
            if (lp > best) { best = lp; bi = i; }
        }
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (t == 0) {
            if (bi == 0x7FFFFFFF) bi = 0;   // all-NaN logits: torch.argmax returns the first NaN's index
            a.label[w] = win_ok ? bi : -1;
            a.conf[w] = win_ok ? expf(best) : 0.f;  // probability; the host multiplies by 100.0 in double like the reference
        }
    }
}

// ---------------------------------------------------------------- fighter boxes from ult_logger records
// Reference playaid/fighter.py:487-539 (Fighter.set_from_json, bbox part) with its camera math :66-155: look-at pose
// from camera / target, the four world-space corners around the fighter projected through the pose's inverse and the
// pin-hole intrinsics of a virtual 1280 x 720 image, np.round (half to even), centre / extent of the rounded pixels,
// then YoloCrop.yolo_pixels' int() truncation (:305-314). fp64 throughout, no FMA contraction (the file is compiled
// with -fmad=false), IEEE sqrt / division: the rounded pixels -- hence the boxes -- equal the host path's.
// rec: [n][PA_LOG_STRIDE] doubles {pos_x, pos_y, cam x y z, target x y z, focal length, frame index}
__global__ void boxes_kernel(const double* __restrict__ rec, int n, int W, int H, double* __restrict__ boxes, int32_t* __restrict__ crops) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* r = rec + (size_t)i * PA_LOG_STRIDE;
    const double px_ = r[0], py_ = r[1];
    const double c[3] = {r[2], r[3], r[4]}, t[3] = {r[5], r[6], r[7]};
    const double f = r[8];
    // calculate_lookat_matrix: forward = normalize(cam - target); right = normalize(cross(up, forward)); up = cross(forward, right)
    double fw[3] = {c[0] - t[0], c[1] - t[1], c[2] - t[2]};
    double nrm = sqrt((fw[0] * fw[0] + fw[1] * fw[1]) + fw[2] * fw[2]);
    fw[0] /= nrm; fw[1] /= nrm; fw[2] /= nrm;
    double rt[3] = {1.0 * fw[2] - 0.0 * fw[1], 0.0 * fw[0] - 0.0 * fw[2], 0.0 * fw[1] - 1.0 * fw[0]};
    nrm = sqrt((rt[0] * rt[0] + rt[1] * rt[1]) + rt[2] * rt[2]);
    rt[0] /= nrm; rt[1] /= nrm; rt[2] /= nrm;
    const double up[3] = {fw[1] * rt[2] - fw[2] * rt[1], fw[2] * rt[0] - fw[0] * rt[2], fw[0] * rt[1] - fw[1] * rt[0]};
    // pose M = [R | cam] with rows right / up / -forward; its inverse = [R^-1 | -R^-1 cam], R^-1 by cofactors
    const double R[3][3] = {{rt[0], rt[1], rt[2]}, {up[0], up[1], up[2]}, {-fw[0], -fw[1], -fw[2]}};
    const double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0]) +
                       R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    double Ri[3][3];
    Ri[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det; Ri[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det; Ri[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det;
    Ri[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det; Ri[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det; Ri[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
    Ri[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det; Ri[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det; Ri[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
    double ti[3];
    for (int a = 0; a < 3; a++) ti[a] = -((Ri[a][0] * c[0] + Ri[a][1] * c[1]) + Ri[a][2] * c[2]);
    const double ox[4] = {-10.0, 10.0, -10.0, 10.0}, oy[4] = {20.0, 20.0, -3.0, -3.0};
    long long xs[4], ys[4];
    for (int k = 0; k < 4; k++) {
        const double w[3] = {px_ + ox[k], py_ + oy[k], 0.0};
        double pc[3];
        for (int a = 0; a < 3; a++) pc[a] = ((Ri[a][0] * w[0] + Ri[a][1] * w[1]) + Ri[a][2] * w[2]) + ti[a];
        const double nx = pc[0] / pc[2], ny = pc[1] / pc[2];
        const double u = f * nx + 640.0, v = 720.0 - (f * ny + 360.0);      // K @ normalised point, y flipped
        xs[k] = (long long)rint(u); ys[k] = (long long)rint(v);           // np.round: half to even
    }
    const long long sx = xs[0] + xs[1] + xs[2] + xs[3], sy = ys[0] + ys[1] + ys[2] + ys[3];
    const long long wx = max(max(xs[0], xs[1]), max(xs[2], xs[3])) - min(min(xs[0], xs[1]), min(xs[2], xs[3]));
    const long long wy = max(max(ys[0], ys[1]), max(ys[2], ys[3])) - min(min(ys[0], ys[1]), min(ys[2], ys[3]));
    const double b[4] = {(double)sx / 4.0 / 1280.0, (double)sy / 4.0 / 720.0, (double)wx / 1280.0, (double)wy / 720.0};
    if (boxes) for (int a = 0; a < 4; a++) boxes[(size_t)i * 4 + a] = b[a];
    if (crops) {
        int32_t* o = crops + (size_t)i * PA_BOX_STRIDE;
        o[0] = (int32_t)r[9];
        o[1] = (int32_t)(b[0] * (double)W); o[2] = (int32_t)(b[1] * (double)H);     // int(): truncation toward zero
        o[3] = (int32_t)(b[2] * (double)W); o[4] = (int32_t)(b[3] * (double)H);
        o[5] = o[6] = o[7] = 0;
    }
}

int launch_boxes(const double* rec, int n, int W, int H, double* boxes, int32_t* crops, cudaStream_t stream) {
    if (n <= 0) return PA_OK;
    boxes_kernel<<<(n + 127) / 128, 128, 0, stream>>>(rec, n, W, H, boxes, crops);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_head(const HeadArgs& a, cudaStream_t stream) {
    if (a.n_actions > 128 || a.n_win <= 0) return PA_ERR_INVALID_ARG;
    return launch_pdl(head_kernel, dim3(a.n_win), dim3(512), 0, stream, a) == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

}  // namespace pa
