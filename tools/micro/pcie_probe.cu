// How should crop windows be pulled from PINNED HOST frames into HBM? Compares, on the window geometry of the bench
// (256 frames of 1080p, two 440 x 440 px windows per frame = 1 344-byte row segments at a 5 760-byte pitch):
//   A. SM-issued 16-byte loads (ld.global.nc.v4, 8 in flight per lane) + stores -- what stage_windows_kernel does
//   B. TMA bulk copies: cp.async.bulk host -> shared (one row segment per copy, mbarrier completion), then
//      cp.async.bulk shared -> HBM, a ring of stages per CTA driven by one thread
//   C. one cudaMemcpyAsync of the whole batch (copy-engine peak, 10x the bytes)
//   D. one cudaMemcpy2DAsync per window (pitched copy-engine transfers), 1 / 2 / 4 streams
//   E. D on half of the windows while B pulls the other half
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_probe pcie_probe.cu && ./pcie_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <vector>

#include "../../playaid_core_b200/csrc/ptx.cuh"
using namespace pa;

constexpr int H = 1080, W = 1920, NF = 256;
constexpr int64_t PITCH = W * 3, FSTRIDE = (int64_t)H * PITCH;
constexpr int RH = 440, SEG = 1344;       // rows per window, bytes per row segment (16-byte aligned cover of 440 px)

struct Win { int64_t off; };              // byte offset of the window's first segment inside the batch

__device__ __forceinline__ uint4 ldnc(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// A: one warp per CTA, rows of all windows dealt round-robin to CTAs
__global__ void __launch_bounds__(32) pull_ldst(const uint8_t* src, uint8_t* dst, const Win* wins, int n_win) {
    const int lane = threadIdx.x;
    const int total = n_win * RH;
    constexpr int CH = SEG / 16;          // 84 chunks per row
    for (int row = blockIdx.x; row < total; row += gridDim.x) {
        const int w = row / RH, r = row - w * RH;
        const int64_t o = wins[w].off + (int64_t)r * PITCH;
        uint4 v[3];
#pragma unroll
        for (int u = 0; u < 3; u++) if (lane + 32 * u < CH) v[u] = ldnc(src + o + (lane + 32 * u) * 16);
#pragma unroll
        for (int u = 0; u < 3; u++) if (lane + 32 * u < CH) *(uint4*)(dst + o + (lane + 32 * u) * 16) = v[u];
    }
}

// B: TMA bulk copies through a shared-memory ring
template <int STAGES, int LAG>
__global__ void __launch_bounds__(32) pull_bulk(const uint8_t* src, uint8_t* dst, const Win* wins, int n_win) {
    extern __shared__ __align__(128) uint8_t ring[];     // STAGES x SEG
    __shared__ uint64_t full[STAGES];
    const int total = n_win * RH;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) mbar_init(&full[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    if (threadIdx.x != 0) return;
    auto row_off = [&](int row) { const int w = row / RH, r = row - w * RH; return wins[w].off + (int64_t)r * PITCH; };
    auto load = [&](int j, int row) {
        const int st = j % STAGES;
        mbar_arrive_expect_tx(&full[st], SEG);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(ring + st * SEG)), "l"(src + row_off(row)), "r"(SEG), "r"(smem_u32(&full[st])) : "memory");
    };
    int n_mine = 0;
    for (int row = blockIdx.x; row < total; row += gridDim.x) n_mine++;
    int issued = 0;
    for (; issued < n_mine && issued < STAGES - LAG; issued++) load(issued, blockIdx.x + issued * gridDim.x);
    for (int j = 0; j < n_mine; j++) {
        const int st = j % STAGES;
        mbar_wait(&full[st], (j / STAGES) & 1);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(dst + row_off(blockIdx.x + j * gridDim.x)), "r"(smem_u32(ring + st * SEG)), "r"(SEG) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(LAG - 1) : "memory");     // the store LAG rows back has read its stage
        if (issued < n_mine) { load(issued, blockIdx.x + issued * gridDim.x); issued++; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)NF * FSTRIDE;
    uint8_t *h, *d;
    const bool wc = argc > 1 && argv[1][0] == 'w';      // ./pcie_probe w : write-combined pinned memory
    printf("pinned host memory: %s\n", wc ? "write-combined" : "default (cacheable)");
    if (cudaHostAlloc(&h, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { printf("host alloc failed\n"); return 1; }
    cudaMalloc(&d, bytes);
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (uint8_t)(i >> 12);
    std::vector<Win> wins;
    srand(3);
    for (int f = 0; f < NF; f++)
        for (int k = 0; k < 2; k++) {
            const int x0 = (200 + rand() % 1000) & ~15, y0 = 100 + rand() % 500;      // 16-byte aligned byte column after * 3? keep px multiple of 16
            wins.push_back({(int64_t)f * FSTRIDE + (int64_t)y0 * PITCH + (int64_t)x0 * 3});
        }
    Win* dw; cudaMalloc(&dw, wins.size() * sizeof(Win));
    cudaMemcpy(dw, wins.data(), wins.size() * sizeof(Win), cudaMemcpyHostToDevice);
    const double wbytes = (double)wins.size() * RH * SEG;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int grid : {32, 148}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            pull_ldst<<<grid, 32>>>(h, d, dw, (int)wins.size());
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("A ld/st    grid %4d: %.3f ms  %.1f GB/s  (%s)\n", grid, ms, wbytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
#define RUNB(ST, LG) \
        for (int rep = 0; rep < 2; rep++) { \
            cudaEventRecord(e0); \
            pull_bulk<ST, LG><<<grid, 32, ST * SEG>>>(h, d, dw, (int)wins.size()); \
            cudaEventRecord(e1); cudaEventSynchronize(e1); \
        } \
        cudaEventElapsedTime(&ms, e0, e1); \
        printf("B TMA bulk grid %4d stages %2d lag %d: %.3f ms  %.1f GB/s  (%s)\n", grid, ST, LG, ms, wbytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        RUNB(4, 2) RUNB(8, 2) RUNB(16, 4)
    }
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("C whole-batch cudaMemcpyAsync: %.3f ms  %.1f GB/s\n", ms, bytes / ms / 1e6);
    // D: one cudaMemcpy2DAsync per window (copy engine, pitched) -- GPU time and host issue time
    for (int nstreams : {1, 2, 4}) {
        cudaStream_t st[4];
        for (int i = 0; i < nstreams; i++) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        cudaEvent_t f0[4], f1[4];
        for (int i = 0; i < nstreams; i++) { cudaEventCreate(&f0[i]); cudaEventCreate(&f1[i]); }
        double host_ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaDeviceSynchronize();
            timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
            for (int i = 0; i < nstreams; i++) cudaEventRecord(f0[i], st[i]);
            for (size_t w = 0; w < wins.size(); w++)
                cudaMemcpy2DAsync(d + wins[w].off, PITCH, h + wins[w].off, PITCH, SEG, RH, cudaMemcpyHostToDevice, st[w % nstreams]);
            for (int i = 0; i < nstreams; i++) cudaEventRecord(f1[i], st[i]);
            clock_gettime(CLOCK_MONOTONIC, &t1);
            host_ms = (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
            cudaDeviceSynchronize();
        }
        float mx = 0;
        for (int i = 0; i < nstreams; i++) { cudaEventElapsedTime(&ms, f0[0], f1[i]); if (ms > mx) mx = ms; }
        printf("D memcpy2D per window, %d stream(s): %.3f ms  %.1f GB/s  (host issue %.3f ms for %zu calls; %s)\n", nstreams, mx, wbytes / mx / 1e6,
               host_ms, wins.size(), cudaGetErrorString(cudaGetLastError()));
    }
    // E: copy engine (half of the windows) and the TMA pull (other half) at the same time
    {
        cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
        const int half = (int)wins.size() / 2;
        for (int rep = 0; rep < 2; rep++) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            cudaStreamWaitEvent(s2, e0);
            pull_bulk<4, 2><<<148, 32, 4 * SEG, s2>>>(h, d, dw, half);
            for (int w = half; w < (int)wins.size(); w++)
                cudaMemcpy2DAsync(d + wins[w].off, PITCH, h + wins[w].off, PITCH, SEG, RH, cudaMemcpyHostToDevice, s1);
            cudaEventRecord(e1, s2);
            cudaStreamWaitEvent(s1, e1);
            cudaEventRecord(e1, s1);
            cudaDeviceSynchronize();
        }
        cudaEventElapsedTime(&ms, e0, e1);
        printf("E half copy engine + half TMA pull: %.3f ms  %.1f GB/s\n", ms, wbytes / ms / 1e6);
    }
    // correctness of B on a sample
    cudaMemset(d, 0, bytes);
    pull_bulk<8, 2><<<74, 32, 8 * SEG>>>(h, d, dw, (int)wins.size());
    cudaDeviceSynchronize();
    std::vector<uint8_t> back(SEG);
    int bad = 0;
    for (int w : {0, 17, 511}) for (int r : {0, 100, 439}) {
        const int64_t o = wins[w].off + (int64_t)r * PITCH;
        cudaMemcpy(back.data(), d + o, SEG, cudaMemcpyDeviceToHost);
        for (int i = 0; i < SEG; i++) if (back[i] != h[o + i]) bad++;
    }
    printf("B sample check: %d bad bytes\n", bad);
    return 0;
}
