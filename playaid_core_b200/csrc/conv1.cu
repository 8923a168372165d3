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
// (double-buffered). The epilogue (16 warps) applies scale/shift + ReLU and FUSES the 3x3/s2 max-pool: a CTA
// walks the 32 tiles of a crop in order, keeps the last three conv rows in shared memory (fp32) and emits one
// pooled row [32 x 64] per tile -- the 64x64x64 stem output never touches HBM.
//
// Replaces resnet18.conv1/bn1/relu/maxpool (torchvision, via playaid/models/cnn_action_detector.py:16,32).
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
constexpr int C1_ROW_PITCH = 68;                              // floats per pixel in the pooling buffer (64 + 4: bank spread)
constexpr int C1_ROWS_BYTES = 3 * 64 * C1_ROW_PITCH * 4;      // three conv rows, fp32

template <int NA, int NB>
__global__ void __maxnreg__(88) conv1_kernel(const __grid_constant__ Conv1Maps maps, const Conv1Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1k(smem_raw);
    // one stage = the seven (ky) boxes of ONE operand plane; a split-precision tile (NA = 2) takes two consecutive
    // stages, hi then lo, accumulated into the same TMEM tile -- so both modes run the same two-stage pipeline
    constexpr int NSTAGE = 2;
    constexpr int STAGE_BYTES = C1_A_PLANE;
    uint8_t* sA = smem;                                    // [NSTAGE][7][128 x 64 B]
    uint8_t* sB = smem + NSTAGE * STAGE_BYTES;             // [NB][7][64 x 64 B]
    float* rows = (float*)(sB + NB * C1_B_PLANE);          // [3][64 px][C1_ROW_PITCH] conv rows for the fused max-pool
    uint64_t* bars = (uint64_t*)((uint8_t*)rows + C1_ROWS_BYTES);
    uint64_t* afull = bars;          // [NSTAGE]
    uint64_t* aempty = bars + 3;     // [NSTAGE]
    uint64_t* tfull = bars + 6;      // [2]
    uint64_t* tempty = bars + 8;     // [2]
    uint64_t* bfull = bars + 10;     // weights landed
    uint32_t* tmem_slot = (uint32_t*)(bars + 11);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

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
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (warp >= 2) pdl_wait();     // the epilogue writes an activation buffer an earlier kernel may still read

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop; one elected lane issues) =====
        {
            if (elect_one()) {   // weights once: 7 boxes [32 k, 64 cout] per plane
                mbar_arrive_expect_tx(bfull, NB * C1_B_PLANE);
                for (int pl = 0; pl < NB; pl++)
                    for (int ky = 0; ky < 7; ky++) tma_load_2d(sB + pl * C1_B_PLANE + ky * C1_B_KY, &maps.b[pl], bfull, ky * 32, 0);
            }
            __syncwarp();
            pdl_wait();      // the weights above are constants; the crops come from the preprocess kernel
            int st = 0; uint32_t ph = 0;
            for (int n = blockIdx.x; n < a.n_crops; n += gridDim.x)
            for (int t = 0; t < 32; t++) {   // a CTA walks the tiles of a crop in order (the fused max-pool needs it)
                const int oy0 = t << 1;
#pragma unroll
                for (int pl = 0; pl < NA; pl++) {
                    mbar_wait(&aempty[st], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&afull[st], STAGE_BYTES);
#pragma unroll
                        for (int ky = 0; ky < 7; ky++) {
                            const int dy = ky - 3;           // input row = 2*oy + dy
                            const int py = dy & 1;           // row parity -> which tensor map
                            const int j0 = oy0 + (dy - py) / 2;
                            tma_load_4d(sA + st * STAGE_BYTES + ky * C1_A_KY, &maps.a[pl][py], &afull[st], 0, 0, j0, n);
                        }
                    }
                    __syncwarp();
                    if (++st == NSTAGE) { st = 0; ph ^= 1; }
                }
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
            for (int n = blockIdx.x; n < a.n_crops; n += gridDim.x)
            for (int t = 0; t < 32; t++, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                const uint32_t d_tmem = tmem_base + acc * C1_COUT;
#pragma unroll
                for (int pl = 0; pl < NA; pl++) {            // plane 0: A_hi x (B_hi [+ B_lo]); plane 1: A_lo x B_hi
                    mbar_wait(&afull[st], ph);
                    tc_fence_after();
                    const uint32_t sa0 = smem_u32(sA + st * STAGE_BYTES);
                    if (elect_one()) {
                        const uint64_t da0 = umma_desc_sw64(sa0), db0 = umma_desc_sw64(sb0);
                        const uint64_t dbl0 = (NB == 2) ? umma_desc_sw64(sb0 + C1_B_PLANE) : 0;
#pragma unroll
                        for (int ks = 0; ks < 14; ks++) {
                            const uint32_t ia = ((ks >> 1) * C1_A_KY + (ks & 1) * 32) >> 4;   // start-address field increments
                            const uint32_t ib = ((ks >> 1) * C1_B_KY + (ks & 1) * 32) >> 4;
                            umma_bf16(d_tmem, da0 + ia, db0 + ib, idesc, (pl | ks) != 0);
                            if (NB == 2 && pl == 0) umma_bf16(d_tmem, da0 + ia, dbl0 + ib, idesc, 1);
                        }
                        umma_commit(&aempty[st]);
                        if (pl == NA - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++st == NSTAGE) { st = 0; ph ^= 1; }
                }
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
        const int et = threadIdx.x - 64;           // 0..511 among the epilogue threads
        const int ppx = et >> 4, pcg = (et & 15) * 4; // pooling role: pooled column, first of 4 channels
        int it = 0;
        for (int n = blockIdx.x; n < a.n_crops; n += gridDim.x)
        for (int t = 0; t < 32; t++, it++) {
            const int acc = it & 1;
            const uint32_t acc_ph = (it >> 1) & 1;
            mbar_wait(&tfull[acc], acc_ph);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C1_COUT + c0;
            float v[16];
            tmem_ld16(t_addr, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);   // accumulator drained into registers: release it early
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = fmaxf(fmaf(v[i], sc[i], sh[i]), 0.f);
            // conv row (2t + r/64), column r%64 -> pooling buffer slot (row % 3)
            {
                const int crow = 2 * t + (r >> 6);
                float4* d = (float4*)(rows + ((size_t)(crow % 3) * 64 + (r & 63)) * C1_ROW_PITCH + c0);
#pragma unroll
                for (int i = 0; i < 4; i++) d[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            asm volatile("bar.sync 1, 512;" ::: "memory");   // both conv rows of the tile are in shared memory
            // pooled row t: max over conv rows 2t-1..2t+1 and columns 2px-1..2px+1 (inputs are >= 0 after ReLU)
            float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int dr = -1; dr <= 1; dr++) {
                const int crow = 2 * t + dr;
                if (crow < 0) continue;
#pragma unroll
                for (int dc = -1; dc <= 1; dc++) {
                    const int col = 2 * ppx + dc;
                    if (col < 0) continue;
                    const float4 x = *(const float4*)(rows + ((size_t)(crow % 3) * 64 + col) * C1_ROW_PITCH + pcg);
                    m.x = fmaxf(m.x, x.x); m.y = fmaxf(m.y, x.y); m.z = fmaxf(m.z, x.z); m.w = fmaxf(m.w, x.w);
                }
            }
            const int64_t o = (((int64_t)n * 32 + t) * 32 + ppx) * C1_COUT + pcg;
            uint32_t h0, h1, l0, l1;
            if (a.f16) { split2<true>(m.x, m.y, h0, l0); split2<true>(m.z, m.w, h1, l1); }
            else { split2<false>(m.x, m.y, h0, l0); split2<false>(m.z, m.w, h1, l1); }
            *(uint2*)(a.out_hi + o) = make_uint2(h0, h1);
            if (a.out_lo) *(uint2*)(a.out_lo + o) = make_uint2(l0, l1);
            asm volatile("bar.sync 1, 512;" ::: "memory");   // pooling done before the next tile reuses a slot
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<128>(tmem_base);
}

template <int NA, int NB>
static int launch_c1(const Conv1Maps& maps, const Conv1Args& a, int num_sms, cudaStream_t stream) {
    auto kern = conv1_kernel<NA, NB>;
    constexpr int NSTAGE = 2;
    const size_t smem = 1024 + (size_t)NSTAGE * C1_A_PLANE + (size_t)NB * C1_B_PLANE + C1_ROWS_BYTES + 128;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return PA_ERR_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    int grid = a.n_crops;   // whole crops per CTA
    if (grid > num_sms) grid = num_sms;
    if (launch_pdl(kern, dim3(grid), dim3(C1_THREADS), smem, stream, maps, a) != cudaSuccess) return PA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv1(const Conv1Maps& maps, const Conv1Args& a, int num_sms, cudaStream_t stream) {
    const int na = a.split_a ? 2 : 1, nb = a.split_w ? 2 : 1;
    if (na == 1 && nb == 1) return launch_c1<1, 1>(maps, a, num_sms, stream);
    if (na == 2 && nb == 1) return launch_c1<2, 1>(maps, a, num_sms, stream);
    if (na == 1 && nb == 2) return launch_c1<1, 2>(maps, a, num_sms, stream);   // byte-valued crops with split weights
    if (na == 2 && nb == 2) return launch_c1<2, 2>(maps, a, num_sms, stream);
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
