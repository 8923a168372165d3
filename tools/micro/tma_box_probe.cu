// What does a TMA box with a 64-byte inner dimension look like in shared memory under SWIZZLE_128B?
// Map: dims {32 elems, 8 row pairs (stride 2 rows), 64 ox (stride 16 B)}, box {32, 2, 64}. Prints, for a few (ox, rp, e),
// where the element landed compared with the dense-image + address-XOR model.   nvcc -arch=sm_100a tma_box_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, uint16_t* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    uint8_t* tile = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), dst = (uint32_t)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(8192));
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar_a), "r"(0), "r"(1), "r"(0) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar_a), "r"(0));
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = ((uint16_t*)tile)[i];
}

int main() {
    const int rows = 32, rowe = 136 * 4;   // elements per row
    std::vector<uint16_t> h(rows * rowe);
    for (int r = 0; r < rows; r++) for (int e = 0; e < rowe; e++) h[r * rowe + e] = (uint16_t)(r * 1024 + (e & 1023));
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 8192);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap map;
    cuuint64_t dims[3] = {32, 8, 64};
    cuuint64_t strides[2] = {(cuuint64_t)2 * rowe * 2, 16};
    cuuint32_t box[3] = {32, 2, 64}, estr[3] = {1, 1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)rc);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    probe<<<1, 128, 16384>>>(map, o);
    printf("launch: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<uint16_t> img(4096);
    cudaMemcpy(img.data(), o, 8192, cudaMemcpyDeviceToHost);
    // model: dense image [ox][rp][32 elems] (128 B per ox), 16-byte chunk index XOR (line index & 7)
    int bad = 0;
    for (int ox = 0; ox < 64; ox++) for (int rp = 0; rp < 2; rp++) for (int e = 0; e < 32; e++) {
        const int line = ox, byte_in_line = rp * 64 + e * 2, chunk = byte_in_line >> 4;
        const int pos = line * 128 + (((chunk ^ (line & 7)) << 4) | (byte_in_line & 15));
        const uint16_t want = h[(2 * (1 + rp)) * rowe + ox * 8 + e];
        if (img[pos / 2] != want && bad++ < 6) printf("ox %d rp %d e %d: got %u want %u\n", ox, rp, e, img[pos / 2], want);
    }
    printf("dense-image model mismatches: %d of 4096\n", bad);
    // alternative model: image [rp][ox][32] (64 B rows, 2 per line)
    int bad2 = 0;
    for (int rp = 0; rp < 2; rp++) for (int ox = 0; ox < 64; ox++) for (int e = 0; e < 32; e++) {
        const int lin = (rp * 64 + ox) * 64 + e * 2, line = lin >> 7, chunk = (lin >> 4) & 7;
        const int pos = line * 128 + (((chunk ^ (line & 7)) << 4) | (lin & 15));
        if (img[pos / 2] != h[(2 * (1 + rp)) * rowe + ox * 8 + e]) bad2++;
    }
    printf("[rp][ox] model mismatches: %d of 4096\n", bad2);
    printf("first 40 elements of the image:"); for (int i = 0; i < 40; i++) printf(" %u", img[i]); printf("\n");
    return 0;
}
