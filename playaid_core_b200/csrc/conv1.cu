// ResNet stem: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BN + ReLU as a tcgen05 implicit GEMM.
//
// Cin = 3 cannot feed a TMA im2col box, so the A tile is assembled by four producer warps:
// the input is the 4-channel bf16 crop written by the preprocess kernel (NHWC4, channel 3 == 0);
// for one output pixel and one filter row ky the 8 input pixels x 4 channels it touches are one
// contiguous, 16-byte aligned 64-byte run (2*ox-4 .. 2*ox+3), so K is laid out as
//   k = ky*32 + kx8*4 + c      (kx8 = kx + 1; weight is 0 for kx8 == 0 and for c == 3)
// i.e. 7 filter rows x 32 = 224 = 14 tcgen05 k-steps of 16. Producers copy four 16-byte chunks per
// (pixel, ky) into the K-major SWIZZLE_128B tile with zero fill outside the image, fence to the
// async proxy and signal an mbarrier; one thread issues the 14 (x2 / x3 in split precision) MMAs
// into a double-buffered TMEM accumulator; four epilogue warps apply scale/shift + ReLU and store
// NHWC bf16 (hi [+ lo]).
//
// Replaces resnet18.conv1/bn1/relu (torchvision, via playaid/models/cnn_action_detector.py:16,32).
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

constexpr int C1_THREADS = 288;
constexpr int C1_A_PLANE = 4 * 16384;  // 4 k-blocks of [128 rows x 128 B]
constexpr int C1_B_PLANE = 4 * 8192;   // 4 k-blocks of [64 rows x 128 B]
constexpr int C1_IN = 128, C1_OUT = 64, C1_COUT = 64;

template <int NA, int NB>
__global__ void __launch_bounds__(C1_THREADS, 1) conv1_kernel(const Conv1Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int NBUF = (NA == 1) ? 2 : 1;
    uint8_t* sA = smem;                                  // [NBUF][NA][C1_A_PLANE]
    uint8_t* sB = smem + NBUF * NA * C1_A_PLANE;         // [NB][C1_B_PLANE]
    uint64_t* bars = (uint64_t*)(sB + NB * C1_B_PLANE);
    uint64_t* afull = bars;        // [2]
    uint64_t* aempty = bars + 2;   // [2]
    uint64_t* tfull = bars + 4;    // [2]
    uint64_t* tempty = bars + 6;   // [2]
    uint32_t* tmem_slot = (uint32_t*)(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = a.n_crops * (C1_OUT * C1_OUT / 128);

    // weights -> smem (SWIZZLE_128B K-major), once per CTA
    for (int pl = 0; pl < NB; pl++) {
        const uint4* w = (const uint4*)(pl == 0 ? a.w_hi : a.w_lo);
        for (int i = threadIdx.x; i < 64 * 32; i += C1_THREADS) {
            const int n = i >> 5, ch = i & 31;  // 32 chunks of 16 B per row of 256 bf16
            const int kb = ch >> 3, c = ch & 7;
            *(uint4*)(sB + pl * C1_B_PLANE + kb * 8192 + (n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4)) = __ldg(w + i);
        }
    }
    fence_proxy_async_smem();
    if (warp == 8 && lane == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&afull[i], 128); mbar_init(&aempty[i], 1);
            mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 && warp < 8) {
        // ===================== A-tile producers (one output pixel per thread) =====================
        const int r = threadIdx.x - 128;
        const int sw = r & 7;
        const uint32_t row_off = (r >> 3) * 1024 + sw * 128;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int buf = it % NBUF;
            const uint32_t ph = (it / NBUF) & 1;
            mbar_wait(&aempty[buf], ph ^ 1);
            const int n = tile >> 5;
            const int oy = ((tile & 31) << 1) + (r >> 6), ox = r & 63;
            const int ix0 = 2 * ox - 4;
#pragma unroll
            for (int pl = 0; pl < NA; pl++) {
                const bf16* in = (pl == 0 ? a.in_hi : a.in_lo) + (size_t)n * C1_IN * C1_IN * 4;
                uint8_t* dst = sA + (buf * NA + pl) * C1_A_PLANE + row_off;
#pragma unroll
                for (int ky = 0; ky < 7; ky++) {
                    const int iy = 2 * oy + ky - 3;
                    const bool yok = (iy >= 0) && (iy < C1_IN);
                    uint4 v[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int ix = ix0 + 2 * j;
                        v[j] = make_uint4(0, 0, 0, 0);
                        if (yok && ix >= 0 && ix < C1_IN) v[j] = __ldg((const uint4*)(in + ((size_t)iy * C1_IN + ix) * 4));
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int c = (ky & 1) * 4 + j;
                        *(uint4*)(dst + (ky >> 1) * 16384 + ((c ^ sw) << 4)) = v[j];
                    }
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&afull[buf]);
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = a.f16 ? umma_idesc_f16(128, C1_COUT) : umma_idesc_bf16(128, C1_COUT);
            const uint32_t sb0 = smem_u32(sB);
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int buf = it % NBUF;
                const uint32_t ph = (it / NBUF) & 1;
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                mbar_wait(&afull[buf], ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C1_COUT;
                const uint32_t sa0 = smem_u32(sA + buf * NA * C1_A_PLANE);
#pragma unroll
                for (int ks = 0; ks < 14; ks++) {
                    const uint32_t koff_a = (ks >> 2) * 16384 + (ks & 3) * 32;
                    const uint32_t koff_b = (ks >> 2) * 8192 + (ks & 3) * 32;
                    const uint64_t da = umma_desc_sw128(sa0 + koff_a);
                    const uint64_t db = umma_desc_sw128(sb0 + koff_b);
                    umma_bf16(d_tmem, da, db, idesc, ks != 0);
                    if (NA == 2) umma_bf16(d_tmem, umma_desc_sw128(sa0 + C1_A_PLANE + koff_a), db, idesc, 1);
                    if (NB == 2) umma_bf16(d_tmem, da, umma_desc_sw128(sb0 + C1_B_PLANE + koff_b), idesc, 1);
                }
                umma_commit(&aempty[buf]);
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        // ===================== epilogue (warps 0..3) =====================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            const int64_t o = ((int64_t)tile * 128 + r) * C1_COUT;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C1_COUT;
#pragma unroll 1
            for (int c0 = 0; c0 < C1_COUT; c0 += 16) {
                float v[16];
                tmem_ld16(t_addr + c0, v);
                uint32_t h[8], l[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    float x0 = fmaxf(v[2 * i] * __ldg(a.scale + c0 + 2 * i) + __ldg(a.shift + c0 + 2 * i), 0.f);
                    float x1 = fmaxf(v[2 * i + 1] * __ldg(a.scale + c0 + 2 * i + 1) + __ldg(a.shift + c0 + 2 * i + 1), 0.f);
                    if (a.f16) split2<true>(x0, x1, h[i], l[i]);
                    else split2<false>(x0, x1, h[i], l[i]);
                }
                uint4* op = (uint4*)(a.out_hi + o + c0);
                op[0] = make_uint4(h[0], h[1], h[2], h[3]);
                op[1] = make_uint4(h[4], h[5], h[6], h[7]);
                if (a.out_lo) {
                    uint4* lp = (uint4*)(a.out_lo + o + c0);
                    lp[0] = make_uint4(l[0], l[1], l[2], l[3]);
                    lp[1] = make_uint4(l[4], l[5], l[6], l[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<128>(tmem_base);
}

template <int NA, int NB>
static int launch_c1(const Conv1Args& a, int num_sms, cudaStream_t stream) {
    auto kern = conv1_kernel<NA, NB>;
    constexpr int NBUF = (NA == 1) ? 2 : 1;
    const size_t smem = 1024 + (size_t)NBUF * NA * C1_A_PLANE + (size_t)NB * C1_B_PLANE + 128;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PA_ERR_CUDA;
        attr_set = true;
    }
    int grid = a.n_crops * 32;
    if (grid > num_sms) grid = num_sms;
    kern<<<grid, C1_THREADS, smem, stream>>>(a);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv1(const Conv1Args& a, int num_sms, cudaStream_t stream) {
    const int na = a.in_lo ? 2 : 1, nb = a.w_lo ? 2 : 1;
    if (na == 1 && nb == 1) return launch_c1<1, 1>(a, num_sms, stream);
    if (na == 2 && nb == 1) return launch_c1<2, 1>(a, num_sms, stream);
    if (na == 2 && nb == 2) return launch_c1<2, 2>(a, num_sms, stream);
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
