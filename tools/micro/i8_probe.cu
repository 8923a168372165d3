// Probe for the tensor-core preprocess kernel (csrc/preprocess_tc.cu): checks on a B200 that
//   1. tcgen05.mma kind::i8 with a u8 operand and an s8 operand accumulates exactly in s32 (all four
//      signedness combinations of the instruction descriptor, K-major SWIZZLE_128B tiles written by threads),
//   2. a TMA box of u8 elements starting at an UNALIGNED byte column lands in the SWIZZLE_128B layout the UMMA
//      descriptor expects, with zero fill beyond the row end,
//   3. how many clocks one 128 x N x 32 i8 MMA takes from shared-memory operands (N = 64, 128, 256).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_probe i8_probe.cu && ./i8_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../playaid_core_b200/csrc/ptx.cuh"
using namespace pa;

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// D = s32 (2 << 4), A fmt at [7,10), B fmt at [10,13): 0 = u8, 1 = s8
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t afmt, uint32_t bfmt, uint32_t M, uint32_t N) {
    return (2u << 4) | (afmt << 7) | (bfmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16i(uint32_t taddr, int32_t* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    for (int i = 0; i < 16; i++) v[i] = (int32_t)r[i];
}

// ---- 1. correctness: A [128][128 B], B [N][128 B] row-major in global -> swizzled smem -> 4 k-steps -> D [128][N] s32
__global__ void __launch_bounds__(128) mma_probe(const uint8_t* A, const uint8_t* B, int32_t* D, int N, uint32_t afmt, uint32_t bfmt) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;              // 128 x 128 B
    uint8_t* sB = smem + 16384;      // up to 256 x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *(uint4*)(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *(const uint4*)(A + r * 128 + c * 16);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *(uint4*)(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *(const uint4*)(B + r * 128 + c * 16);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(&tslot);
    fence_proxy_async_smem();     // generic-proxy writes -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (tid == 0) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA)), db = umma_desc_sw128(smem_u32(sB));
        const uint32_t id = idesc_i8(afmt, bfmt, 128, (uint32_t)N);
        for (int k = 0; k < 4; k++) umma_i8(tb, da + 2 * k, db + 2 * k, id, k != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        int32_t v[16];
        tmem_ld16i(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; i++) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tb);
}

// ---- 2. TMA: box {128 B, 64 rows} of a u8 image at byte column c0 (unaligned) -> raw smem dump
__global__ void __launch_bounds__(128) tma_probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int c0, int c1) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tile = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, 64 * 128);
        tma_load_2d(tile, &map, &bar, c0, c1);
    }
    // bounded wait that reports instead of trapping
    bool done = false;
    for (int spin = 0; spin < 2000000 && !done; spin++) done = mbar_try_wait(&bar, 0);
    if (threadIdx.x == 0) out[8192] = done ? 1 : 0;
    for (int i = threadIdx.x; i < 64 * 128; i += 128) out[i] = tile[i];
}

// ---- 3. timing: `iters` x 4 k-steps, no operand refill
__global__ void __launch_bounds__(128) mma_time(long long* clk, int N, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (16384 + 32768) / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(&tslot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tslot;
    if (tid == 0) {
        const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 16384));
        const uint32_t id = idesc_i8(0, 1, 128, (uint32_t)N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; it++)
            for (int k = 0; k < 4; k++) umma_i8(tb, da + 2 * k, db + 2 * k, id, 1);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        clk[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tb);
}

int main() {
    srand(7);
    std::vector<uint8_t> A(128 * 128), B(256 * 128);
    for (auto& v : A) v = (uint8_t)(rand() & 255);
    for (auto& v : B) v = (uint8_t)(rand() & 255);
    uint8_t *dA, *dB; int32_t* dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 256 * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const int smem = 1024 + 16384 + 32768;
    cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mma_time, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int N : {64, 128, 256})
        for (int af = 0; af < 2; af++)
            for (int bf = 0; bf < 2; bf++) {
                cudaMemset(dD, 0xff, 128 * 256 * 4);
                mma_probe<<<1, 128, smem>>>(dA, dB, dD, N, af, bf);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mma_probe N=%d a=%d b=%d: %s\n", N, af, bf, cudaGetErrorString(e)); return 1; }
                std::vector<int32_t> D(128 * N);
                cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
                // expected under each interpretation: fmt 0 = unsigned, 1 = signed
                int bad = 0, bad_swapped = 0;
                for (int m = 0; m < 128; m++)
                    for (int n = 0; n < N; n++) {
                        long long s = 0, s2 = 0;
                        for (int k = 0; k < 128; k++) {
                            const int a_u = A[m * 128 + k], a_s = (int8_t)A[m * 128 + k], b_u = B[n * 128 + k], b_s = (int8_t)B[n * 128 + k];
                            s += (long long)(af ? a_s : a_u) * (bf ? b_s : b_u);
                            s2 += (long long)(af ? a_u : a_s) * (bf ? b_u : b_s);   // the opposite encoding (0 = signed)
                        }
                        if (D[m * N + n] != (int32_t)s) bad++;
                        if (D[m * N + n] != (int32_t)s2) bad_swapped++;
                    }
                printf("i8 mma N=%3d afmt=%d bfmt=%d: mismatches %d (fmt 0=u8,1=s8) / %d (fmt 0=s8,1=u8) of %d   D[0][0..3] = %d %d %d %d\n",
                       N, af, bf, bad, bad_swapped, 128 * N, D[0], D[1], D[2], D[3]);
            }
    // ---- TMA
    {
        const int rows = 96, pitch = 5760;
        std::vector<uint8_t> img((size_t)rows * pitch);
        for (int r = 0; r < rows; r++) for (int c = 0; c < pitch; c++) img[(size_t)r * pitch + c] = (uint8_t)((r * 131 + c * 7 + (c >> 8)) & 255);
        uint8_t *dI, *dO;
        cudaMalloc(&dI, img.size()); cudaMalloc(&dO, 64 * 128 + 16);
        cudaMemcpy(dI, img.data(), img.size(), cudaMemcpyHostToDevice);
        void* fn = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        auto enc = (decltype(&cuTensorMapEncodeTiled))fn;
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)pitch};
        cuuint32_t box[2] = {128, 64}, es[2] = {1, 1};
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dI, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tma encode rc=%d\n", (int)rc);
        cudaFuncSetAttribute(tma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 8192);
        const int cases[6][2] = {{0, 0}, {16, 2}, {48, 3}, {5744, 0}, {5696, 40}, {-16, 10}};
        for (auto& cs : cases) {
            const int c0 = cs[0], c1 = cs[1];
            cudaMemset(dO, 0xEE, 8192);
            tma_probe<<<1, 128, 1024 + 8192>>>(map, dO, c0, c1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("tma_probe: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<uint8_t> o(8193);
            cudaMemcpy(o.data(), dO, 8193, cudaMemcpyDeviceToHost);
            if (!o[8192]) { printf("tma u8 box at (c0=%d, c1=%d): TIMED OUT\n", c0, c1); continue; }
            int bad = 0;
            for (int r = 0; r < 64; r++)
                for (int b = 0; b < 128; b++) {
                    const int gr = c1 + r, gc = c0 + b;
                    const uint8_t want = (gr >= 0 && gr < rows && gc >= 0 && gc < pitch) ? img[(size_t)gr * pitch + gc] : 0;
                    const int pos = r * 128 + ((((b >> 4) ^ (r & 7)) << 4) | (b & 15));
                    if (o[pos] != want && bad++ < 4) printf("   r %d b %d: got %u want %u\n", r, b, o[pos], want);
                }
            printf("tma u8 box at (c0=%d, c1=%d): %d mismatches of 8192 vs the SW128 model (OOB -> 0)\n", c0, c1, bad);
        }
    }
    // ---- timing
    {
        long long* dclk; cudaMalloc(&dclk, 8);
        for (int N : {64, 128, 256}) {
            for (int rep = 0; rep < 2; rep++) {
                mma_time<<<1, 128, smem>>>(dclk, N, 2000);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mma_time: %s\n", cudaGetErrorString(e)); return 1; }
            }
            long long c; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
            printf("i8 mma 128x%dx32 from smem: %.1f clk per MMA (%d MMAs)\n", N, (double)c / 8000.0, 8000);
        }
    }
    return 0;
}
