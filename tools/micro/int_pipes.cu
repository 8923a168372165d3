// Microbenchmark: per-SM throughput of IMAD vs IDP.2A (dp2a) vs IDP.4A (dp4a) vs PRMT+IMAD pairs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int* out, int iters, int a0, int b0) {
    int acc[8];
    int a = a0 + threadIdx.x, b = b0 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = i + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) acc[i] = acc[i] * a + b;                       // IMAD (dependent chain per acc, 8 chains)
            if (MODE == 1) { int d; asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b + i), "r"(acc[i])); acc[i] = d; }  // IDP.2A s16 x u8
            if (MODE == 2) acc[i] = __dp4a(a, b + i, acc[i]);             // IDP.4A
            if (MODE == 3) acc[i] += (int)__byte_perm(b + acc[(i + 1) & 7], 0, 0x4441) * a;  // PRMT + IMAD
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name) {
    int* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(int));
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 1024>>>(out, 16, 3, 5);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 1024>>>(out, iters, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 1024 * iters * 8;
    printf("%-12s %.3f ms  %.2f Tops/s  (%.1f lane-ops/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main() { run<0>("IMAD"); run<1>("DP2A"); run<2>("DP4A"); run<3>("PRMT+IMAD"); return 0; }
