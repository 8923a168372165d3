// ResNet stem: conv 7x7 stride 2 pad 3 (3 -> 64) + folded BN + ReLU as a tcgen05 implicit GEMM
// whose im2col is done entirely by the TMA unit.
//
// Input: the 16-bit crop written by the preprocess kernel in NHWC4P layout -- 4 channels per pixel
// (channel 3 == 0) and 4 zero pixels left/right of every row: [n][128][136][4]. For one output pixel
// (oy, ox) and one filter row ky the 8 input pixels x 4 channels it touches (2ox-4 .. 2ox+3, the
// leftmost one multiplied by a zero weight) are ONE contiguous 64-byte run starting at padded pixel
// 2ox. So K is laid out as
//     k = ky*32 + kx8*4 + c        (kx8 = kx + 1; weight 0 for kx8 == 0 and c == 3)
// and the A tile of filter row ky for 128 output pixels (2 output rows x 64 columns) is a 4-D TMA box
//     [32 elements, 64 ox (stride 16 B -- overlapping windows), 2 row-pairs (stride 2 rows), 1 crop]
// of a tensor map over every other input row (one map per row parity); rows above/below the image
// are zero-filled by the TMA unit, columns by the padding in memory. Seven 8 KB boxes per tile land
// in K-major SWIZZLE_64B tiles, 14 tcgen05.mma (M=128, N=64, K=16) accumulate in TMEM
// (double-buffered), four epilogue warps apply scale/shift + ReLU and store NHWC.
//
// Replaces resnet18.conv1/bn1/relu (torchvision, via playaid/models/cnn_action_detector.py:16,32).
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

constexpr int C1_EPI_WARPS = 16;
constexpr int C1_THREADS = (2 + C1_EPI_WARPS) * 32;  // TMA warp, MMA warp, 16 epilogue warps
constexpr int C1_A_KY = 128 * 64;          // one filter row of the A tile: 128 rows x 64 B
constexpr int C1_A_PLANE = 7 * C1_A_KY;    // 56 KB
constexpr int C1_B_KY = 64 * 64;           // 64 cout rows x 64 B
constexpr int C1_B_PLANE = 7 * C1_B_KY;    // 28 KB
constexpr int C1_COUT = 64;

template <int NA, int NB>
__global__ void __launch_bounds__(C1_THREADS, 1) conv1_kernel(const __grid_constant__ Conv1Maps maps, const Conv1Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    constexpr int NSTAGE = (NA == 1) ? 3 : 1;
    constexpr int STAGE_BYTES = NA * C1_A_PLANE;
    uint8_t* sA = smem;                                    // [NSTAGE][NA][7][128 x 64 B]
    uint8_t* sB = smem + NSTAGE * STAGE_BYTES;             // [NB][7][64 x 64 B]
    uint64_t* bars = (uint64_t*)(sB + NB * C1_B_PLANE);
    uint64_t* afull = bars;          // [NSTAGE]
    uint64_t* aempty = bars + 3;     // [NSTAGE]
    uint64_t* tfull = bars + 6;      // [2]
    uint64_t* tempty = bars + 8;     // [2]
    uint64_t* bfull = bars + 10;     // weights landed
    uint32_t* tmem_slot = (uint32_t*)(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = a.n_crops * 32;

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < NA; pl++) { tma_prefetch_desc(&maps.a[pl][0]); tma_prefetch_desc(&maps.a[pl][1]); }
        for (int pl = 0; pl < NB; pl++) tma_prefetch_desc(&maps.b[pl]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NSTAGE; i++) { mbar_init(&afull[i], 1); mbar_init(&aempty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], C1_EPI_WARPS); }
        mbar_init(bfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop; one elected lane issues) =====
        {
            if (elect_one()) {   // weights once: 7 boxes [32 k, 64 cout] per plane
                mbar_arrive_expect_tx(bfull, NB * C1_B_PLANE);
                for (int pl = 0; pl < NB; pl++)
                    for (int ky = 0; ky < 7; ky++) tma_load_2d(sB + pl * C1_B_PLANE + ky * C1_B_KY, &maps.b[pl], bfull, ky * 32, 0);
            }
            __syncwarp();
            int st = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n = tile >> 5, oy0 = (tile & 31) << 1;
                mbar_wait(&aempty[st], ph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&afull[st], STAGE_BYTES);
#pragma unroll
                    for (int pl = 0; pl < NA; pl++) {
#pragma unroll
                        for (int ky = 0; ky < 7; ky++) {
                            const int dy = ky - 3;           // input row = 2*oy + dy
                            const int py = dy & 1;           // row parity -> which tensor map
                            const int j0 = oy0 + (dy - py) / 2;
                            tma_load_4d(sA + st * STAGE_BYTES + pl * C1_A_PLANE + ky * C1_A_KY, &maps.a[pl][py], &afull[st], 0, 0, j0, n);
                        }
                    }
                }
                __syncwarp();
                if (++st == NSTAGE) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loop; one elected lane issues) =====
        {
            const uint32_t idesc = a.f16 ? umma_idesc_f16(128, C1_COUT) : umma_idesc_bf16(128, C1_COUT);
            const uint32_t sb0 = smem_u32(sB);
            mbar_wait(bfull, 0);
            int st = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                mbar_wait(&afull[st], ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * C1_COUT;
                const uint32_t sa0 = smem_u32(sA + st * STAGE_BYTES);
                if (elect_one()) {
                    const uint64_t da0 = umma_desc_sw64(sa0), db0 = umma_desc_sw64(sb0);
                    const uint64_t dal0 = (NA == 2) ? umma_desc_sw64(sa0 + C1_A_PLANE) : 0;
                    const uint64_t dbl0 = (NB == 2) ? umma_desc_sw64(sb0 + C1_B_PLANE) : 0;
#pragma unroll
                    for (int ks = 0; ks < 14; ks++) {
                        const uint32_t ia = ((ks >> 1) * C1_A_KY + (ks & 1) * 32) >> 4;   // start-address field increments
                        const uint32_t ib = ((ks >> 1) * C1_B_KY + (ks & 1) * 32) >> 4;
                        umma_bf16(d_tmem, da0 + ia, db0 + ib, idesc, ks != 0);
                        if (NA == 2) umma_bf16(d_tmem, dal0 + ia, db0 + ib, idesc, 1);
                        if (NB == 2) umma_bf16(d_tmem, da0 + ia, dbl0 + ib, idesc, 1);
                    }
                    umma_commit(&aempty[st]);
                    umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++st == NSTAGE) { st = 0; ph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..17): four warps per TMEM lane quarter, 16 columns each ====
        const int q = warp & 3;
        const int c0 = ((warp - 2) >> 2) * 16;   // this warp's fixed column chunk
        const int r = q * 32 + lane;
        float sc[16], sh[16];
#pragma unroll
        for (int i = 0; i < 16; i++) { sc[i] = __ldg(a.scale + c0 + i); sh[i] = __ldg(a.shift + c0 + i); }
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
            const int acc = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            const int64_t o = ((int64_t)tile * 128 + r) * C1_COUT + c0;
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C1_COUT + c0;
            float v[16];
            tmem_ld16(t_addr, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);   // accumulator drained into registers: release it early
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
            uint32_t h[8], l[8];
            if (a.out_lo) {
                if (a.f16) {
#pragma unroll
                    for (int i = 0; i < 8; i++) split2<true>(v[2 * i], v[2 * i + 1], h[i], l[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) split2<false>(v[2 * i], v[2 * i + 1], h[i], l[i]);
                }
                uint4* lp = (uint4*)(a.out_lo + o);
                lp[0] = make_uint4(l[0], l[1], l[2], l[3]);
                lp[1] = make_uint4(l[4], l[5], l[6], l[7]);
            } else if (a.f16) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const __half2 t = __floats2half2_rn(fminf(v[2 * i], 65504.f), fminf(v[2 * i + 1], 65504.f));
                    h[i] = *reinterpret_cast<const uint32_t*>(&t);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++) h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            }
            uint4* op = (uint4*)(a.out_hi + o);
            op[0] = make_uint4(h[0], h[1], h[2], h[3]);
            op[1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<128>(tmem_base);
}

template <int NA, int NB>
static int launch_c1(const Conv1Maps& maps, const Conv1Args& a, int num_sms, cudaStream_t stream) {
    auto kern = conv1_kernel<NA, NB>;
    constexpr int NSTAGE = (NA == 1) ? 3 : 1;
    const size_t smem = 1024 + (size_t)NSTAGE * NA * C1_A_PLANE + (size_t)NB * C1_B_PLANE + 128;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PA_ERR_CUDA;
        attr_set = true;
    }
    int grid = a.n_crops * 32;
    if (grid > num_sms) grid = num_sms;
    kern<<<grid, C1_THREADS, smem, stream>>>(maps, a);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv1(const Conv1Maps& maps, const Conv1Args& a, int num_sms, cudaStream_t stream) {
    const int na = a.split_a ? 2 : 1, nb = a.split_w ? 2 : 1;
    if (na == 1 && nb == 1) return launch_c1<1, 1>(maps, a, num_sms, stream);
    if (na == 2 && nb == 1) return launch_c1<2, 1>(maps, a, num_sms, stream);
    if (na == 2 && nb == 2) return launch_c1<2, 2>(maps, a, num_sms, stream);
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
