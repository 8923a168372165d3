// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (sm_100a).
//
//   D[M, Cout] = A[M, K] * W[Cout, K]^T,  M = output pixels of all crops (NHWC), K = taps * Cin
//
// * A is never materialised: for every filter tap the TMA engine loads a 4-D box
//   [64 channels, Wt, Ht, Nt] (Wt*Ht*Nt = 128 output pixels) straight from the NHWC activation;
//   out-of-bounds coordinates are zero-filled by the TMA unit, which implements the conv padding.
//   Stride-2 layers use four parity views of the input (one tensor map per (y&1, x&1)).
// * W tiles [BLOCK_N, 64] come from a K-major packed weight matrix through a 2-D tensor map.
// * Both operands land in shared memory in the canonical K-major SWIZZLE_128B layout and are
//   consumed by tcgen05.mma (M=128, N=BLOCK_N, K=16 per instruction) with the fp32 accumulator in
//   TMEM (double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1).
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..17 = epilogue
//   (TMEM -> registers -> folded-BN scale/shift (+ residual) (+ ReLU) -> 16-bit / fp32 global).
// * Split precision (PA_PREC_BF16X2 / X3): activations (and optionally weights) carry a bf16
//   "lo" plane with the rounding residual; the extra products are accumulated into the same
//   TMEM tile, reusing the staged weight tile.
//
// Replaces the cuDNN/oneDNN calls behind torchvision.resnet18 + nn.Conv1d/nn.Linear in
// playaid/models/cnn_action_detector.py:13-43.
#include "conv_common.cuh"

namespace pa {

// Split precisions can make NA + NB - 1 PASSES over K per tile into one accumulator, smallest products first
// (A_lo x B, A x B_lo, then A_hi x B_hi): the tensor core truncates the fp32 accumulator at every k-step, and a pass of
// residual products sums to ~2^-11 of the result, so only the last pass truncates at full scale (see conv_gemm2.cu and
// tests/test_gpu_layers.py::test_fp32_accumulation_floor_grows_with_k). A stage then holds one A plane and one B plane and
// the operands stream once per pass, which this 1-CTA kernel pays for (layer2.0.conv1: 52 -> 67 us): the truncation bias is
// linear in K, so plan_conv asks for passes only where it matters (K >= 1152, or split weights); otherwise the products of
// a k-step are interleaved in one pass and a stage holds every plane.
size_t conv_gemm_smem_bytes(int block_n, int n_a, int n_b, int num_stages, bool multipass) {
    size_t stage = multipass ? (size_t)CG_A_BYTES + (size_t)block_n * CG_BLOCK_K * 2
                             : (size_t)n_a * CG_A_BYTES + (size_t)n_b * block_n * CG_BLOCK_K * 2;
    return 1024 /*alignment slack*/ + stage * num_stages + 256 /*barriers*/;
}
int conv_gemm_pick_stages(int block_n, int n_a, int n_b, bool multipass) {
    size_t stage = multipass ? (size_t)CG_A_BYTES + (size_t)block_n * CG_BLOCK_K * 2
                             : (size_t)n_a * CG_A_BYTES + (size_t)n_b * block_n * CG_BLOCK_K * 2;
    int s = (int)((PA_CONV_SMEM_BUDGET - 1024 - 256) / stage);
    if (s > 8) s = 8;
    return s;
}

template <int BLOCK_N, int NA, int NB, bool MP>
__global__ void __maxnreg__(CG_MAX_REGS)
conv_gemm_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1k(smem_raw);
    constexpr int B_BYTES = BLOCK_N * CG_BLOCK_K * 2;
    constexpr int STAGE_BYTES = MP ? CG_A_BYTES + B_BYTES : NA * CG_A_BYTES + NB * B_BYTES;
    constexpr int NPASS = MP ? NA + NB - 1 : 1;      // pass p: p < NA - 1 -> (A_lo, B_hi); p < NPASS - 1 -> (A_hi, B_lo); last -> (A_hi, B_hi)
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64 ? 64 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512)));
    const int S = args.num_stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int taps = args.taps_h * args.taps_w;
    const int num_kb = taps * args.kb_per_tap;
    const int total_tiles = args.m_tiles * args.n_tiles;

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < NA; pl++)
            for (int q = 0; q < (args.stride == 2 ? 4 : 1); q++) tma_prefetch_desc(&maps.a[pl][q]);
        for (int pl = 0; pl < NB; pl++) tma_prefetch_desc(&maps.b[pl]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (warp != 1) pdl_wait();     // producer and epilogue warps touch the previous layer's tensors; the MMA warp only shared memory

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop; one elected lane issues) =====
        {
            int st = 0; uint32_t ph = 0;
            const int pix_per_img = args.ho * args.wo;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / args.n_tiles, nt = tile - mt * args.n_tiles;
                const int m0 = mt * CG_BLOCK_M;
                const int n0 = m0 / pix_per_img;
                const int oy0 = (m0 - n0 * pix_per_img) / args.wo;
                for (int pass = 0; pass < NPASS; pass++) {
                const int pla = (pass < NA - 1) ? 1 : 0, plb = (pass >= NA - 1 && pass < NPASS - 1) ? 1 : 0;
                for (int tap = 0; tap < taps; tap++) {
                    const int ky = tap / args.taps_w, kx = tap - ky * args.taps_w;
                    int cx, cy, q = 0;
                    if (args.stride == 1) {
                        cx = kx - args.pad; cy = oy0 + ky - args.pad;
                    } else {
                        const int dx = kx - args.pad, dy = ky - args.pad;
                        const int px = dx & 1, py = dy & 1;
                        q = py * 2 + px;
                        cx = (dx - px) / 2; cy = oy0 + (dy - py) / 2;
                    }
                    for (int kc = 0; kc < args.kb_per_tap; kc++) {
                        mbar_wait(&empty[st], ph ^ 1);
                        uint8_t* sa = smem + (size_t)st * STAGE_BYTES;
                        uint8_t* sb = sa + (MP ? 1 : NA) * CG_A_BYTES;
                        if (args.debug & 2) {   // experiment: no operand loads at all (pure MMA issue rate)
                            if (elect_one()) mbar_arrive(&full[st]);
                        } else if (elect_one()) {
                            mbar_arrive_expect_tx(&full[st], STAGE_BYTES);
                            if (MP) {
                                tma_load_4d(sa, &maps.a[pla][q], &full[st], kc * CG_BLOCK_K, cx, cy, n0);
                                tma_load_2d(sb, &maps.b[plb], &full[st], tap * args.k_per_tap + kc * CG_BLOCK_K, nt * BLOCK_N);
                            } else {
#pragma unroll
                                for (int pl = 0; pl < NA; pl++)
                                    tma_load_4d(sa + pl * CG_A_BYTES, &maps.a[pl][q], &full[st], kc * CG_BLOCK_K, cx, cy, n0);
#pragma unroll
                                for (int pl = 0; pl < NB; pl++)
                                    tma_load_2d(sb + pl * B_BYTES, &maps.b[pl], &full[st], tap * args.k_per_tap + kc * CG_BLOCK_K, nt * BLOCK_N);
                            }
                        }
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                    }
                }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loop; one elected lane issues) =====
        {
            const uint32_t idesc = args.f16 ? umma_idesc_f16(CG_BLOCK_M, BLOCK_N) : umma_idesc_bf16(CG_BLOCK_M, BLOCK_N);
            int st = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < NPASS * num_kb; kb++) {
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
                    const uint32_t sb = sa + (MP ? 1 : NA) * CG_A_BYTES;
                    if (elect_one()) {
                        // descriptors differ only in the 14-bit start-address field: add 32 B >> 4 per k-step
                        const uint64_t da0 = umma_desc_sw128(sa), db0 = umma_desc_sw128(sb);
                        const uint64_t dal0 = (!MP && NA == 2) ? umma_desc_sw128(sa + CG_A_BYTES) : 0;
                        const uint64_t dbl0 = (!MP && NB == 2) ? umma_desc_sw128(sb + B_BYTES) : 0;
#pragma unroll
                        for (int k = 0; k < CG_BLOCK_K / 16; k++) {
                            umma_bf16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                            if (!MP && NA == 2) umma_bf16(d_tmem, dal0 + 2 * k, db0 + 2 * k, idesc, 1);
                            if (!MP && NB == 2) umma_bf16(d_tmem, da0 + 2 * k, dbl0 + 2 * k, idesc, 1);
                        }
                        umma_commit(&empty[st]);
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1; }
                }
                if (elect_one()) umma_commit(&tfull[acc]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..17) =====================
        const bool split_out = args.out_lo != nullptr;
        if (BLOCK_N == 64 && epilogue_n64_ok(args)) {
            if (args.f16) {
                if (split_out) epilogue_n64<true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
                else epilogue_n64<true, false>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            } else {
                if (split_out) epilogue_n64<false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
                else epilogue_n64<false, false>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            }
        } else if (args.f16) {
            if (split_out) epilogue<BLOCK_N, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, true, false>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        } else {
            if (split_out) epilogue<BLOCK_N, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, false, false>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

template <int BLOCK_N, int NA, int NB, bool MP>
static int launch_t(const ConvMaps& maps, const ConvArgs& args, int num_sms, cudaStream_t stream) {
    auto kern = conv_gemm_kernel<BLOCK_N, NA, NB, MP>;
    const size_t smem = conv_gemm_smem_bytes(BLOCK_N, NA, NB, args.num_stages, MP);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PA_ERR_CUDA;
        attr_set = true;
    }
    int grid = args.m_tiles * args.n_tiles;
    if (grid > num_sms) grid = num_sms;
    if (launch_pdl(kern, dim3(grid), dim3(CG_THREADS), smem, stream, maps, args) != cudaSuccess) return PA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv_gemm(const ConvMaps& maps, const ConvArgs& args_in, int block_n, int n_a, int n_b, int num_sms, cudaStream_t stream) {
    static int dbg = -1;
#ifdef PA_EXPERIMENT
    if (dbg < 0) { const char* e = getenv("PA_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
#else
    dbg = 0;   // the debug modes (wrong results by design) exist only in -DPA_EXPERIMENT builds
#endif
    ConvArgs args = args_in;
    args.debug = dbg;
    const bool mp = args.multipass != 0 && (n_a + n_b) > 2;
#define PA_CG_CASE(BN, A, B, M) \
    if (block_n == BN && n_a == A && n_b == B && mp == M) return launch_t<BN, A, B, M>(maps, args, num_sms, stream);
    PA_CG_CASE(64, 1, 1, false) PA_CG_CASE(128, 1, 1, false) PA_CG_CASE(256, 1, 1, false)
    PA_CG_CASE(64, 2, 1, false) PA_CG_CASE(128, 2, 1, false) PA_CG_CASE(256, 2, 1, false)
    PA_CG_CASE(64, 2, 1, true) PA_CG_CASE(128, 2, 1, true) PA_CG_CASE(256, 2, 1, true)
    PA_CG_CASE(64, 2, 2, true) PA_CG_CASE(128, 2, 2, true) PA_CG_CASE(256, 2, 2, true)
#undef PA_CG_CASE
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
