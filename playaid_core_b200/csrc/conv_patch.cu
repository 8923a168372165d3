// 3x3 / stride-1 / pad-1 convolution as a tcgen05 implicit GEMM with "patch" operand staging.
//
// conv_gemm.cu fetches the tile's input once per filter tap (9 TMA boxes of 128 pixels per 64-channel
// block). For stride-1 layers whose tile is Ht full image rows (W = 32: Ht = 4, W = 16: Ht = 8) the three
// vertical taps of one horizontal shift are the SAME shared-memory data, displaced by whole image rows:
// a box of Ht+2 rows [64 ch, W, Ht+2, 1] is loaded once per horizontal shift dx (TMA zero-fills the
// halo), and the UMMA descriptor of vertical tap dy simply starts dy*W rows (a multiple of the 1024-byte
// swizzle atom) further down. A-operand traffic drops from 9 x 16 KB to 3 x (Ht+2)/Ht x 16 KB per 64
// channels. When the whole weight matrix is small (layer1: 64 x 576 halves = 72 KB) it is loaded once per
// CTA and stays resident, so the steady state streams only activation patches.
//
// Warp roles, TMEM double buffering and the epilogue are those of conv_gemm.cu.
#include "conv_common.cuh"

namespace pa {

struct PatchGeom {
    int patch_bytes;   // (Ht + 2) * W * 128
    int row_bytes;     // W * 128: displacement of one vertical tap
};

template <int BLOCK_N, int NA, bool WRES>
__global__ void __maxnreg__(CG_MAX_REGS)
conv_patch_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args, const PatchGeom pg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1k(smem_raw);
    constexpr int B_BYTES = BLOCK_N * CG_BLOCK_K * 2;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
    const int S = args.num_stages;
    const int kb = args.kb_per_tap;                       // 64-channel blocks
    const int stage_bytes = NA * pg.patch_bytes + (WRES ? 0 : 3 * B_BYTES);
    const int wres_bytes = WRES ? 9 * kb * B_BYTES : 0;
    uint8_t* wres = smem;                                  // resident weights: [tap][kc][BLOCK_N x 128 B]
    uint8_t* stages = smem + wres_bytes;
    uint64_t* bars = (uint64_t*)(stages + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint64_t* wfull = bars + 2 * S + 4;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = args.m_tiles * args.n_tiles;   // n_tiles == 1 for these layers

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < NA; pl++) tma_prefetch_desc(&maps.a[pl][0]);
        tma_prefetch_desc(&maps.b[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CG_EPI_WARPS); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (warp >= CG_FIRST_EPI_WARP) pdl_wait();   // epilogue warps read the residual / write the output of earlier layers' buffers

    if (warp == 0) {
        // ===================== TMA producer (whole warp walks the loop; one elected lane issues) =====
        {
            if (WRES && elect_one()) {
                mbar_arrive_expect_tx(wfull, wres_bytes);
                for (int tap = 0; tap < 9; tap++)
                    for (int kc = 0; kc < kb; kc++)
                        tma_load_2d(wres + (size_t)(tap * kb + kc) * B_BYTES, &maps.b[0], wfull, tap * args.k_per_tap + kc * CG_BLOCK_K, 0);
            }
            __syncwarp();
            pdl_wait();      // the resident weights above do not depend on the previous layer; the activation patches do
            int st = 0; uint32_t ph = 0;
            const int pix_per_img = args.ho * args.wo;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m0 = tile * CG_BLOCK_M;
                const int n0 = m0 / pix_per_img;
                const int oy0 = (m0 - n0 * pix_per_img) / args.wo;
                for (int kc = 0; kc < kb; kc++) {
                    for (int dxi = 0; dxi < 3; dxi++) {
                        mbar_wait(&empty[st], ph ^ 1);
                        uint8_t* sa = stages + (size_t)st * stage_bytes;
                        if ((args.debug & 2) && dxi > 0) {   // experiment: skip two of the three patch loads
                            if (elect_one()) mbar_arrive(&full[st]);
                        } else if (elect_one()) {
                            mbar_arrive_expect_tx(&full[st], stage_bytes);
#pragma unroll
                            for (int pl = 0; pl < NA; pl++)
                                tma_load_4d(sa + pl * pg.patch_bytes, &maps.a[pl][0], &full[st], kc * CG_BLOCK_K, dxi - 1, oy0 - 1, n0);
                            if (!WRES) {
                                uint8_t* sb = sa + NA * pg.patch_bytes;
#pragma unroll
                                for (int dy = 0; dy < 3; dy++)
                                    tma_load_2d(sb + dy * B_BYTES, &maps.b[0], &full[st], (dy * 3 + dxi) * args.k_per_tap + kc * CG_BLOCK_K, 0);
                            }
                        }
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp walks the loop; one elected lane issues) =====
        {
            const uint32_t idesc = args.f16 ? umma_idesc_f16(CG_BLOCK_M, BLOCK_N) : umma_idesc_bf16(CG_BLOCK_M, BLOCK_N);
            if (WRES) { mbar_wait(wfull, 0); tc_fence_after(); }
            const uint32_t wres_u = smem_u32(wres);
            int st = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                uint32_t first = 1;
                for (int kc = 0; kc < kb; kc++) {
                    for (int dxi = 0; dxi < 3; dxi++) {
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stages + (size_t)st * stage_bytes);
                        const uint32_t sb = sa + NA * pg.patch_bytes;
                        if (elect_one()) {
#pragma unroll
                            for (int dy = 0; dy < 3; dy++) {
                                const uint32_t a0 = sa + dy * pg.row_bytes;
                                const uint32_t b0 = WRES ? wres_u + (uint32_t)(((dy * 3 + dxi) * kb + kc) * B_BYTES) : sb + dy * B_BYTES;
                                // descriptors differ only in the 14-bit start-address field: add 32 B >> 4 per k-step
                                const uint64_t da0 = umma_desc_sw128(a0), db0 = umma_desc_sw128(b0);
                                const uint64_t dl0 = (NA == 2) ? umma_desc_sw128(a0 + pg.patch_bytes) : 0;
#pragma unroll
                                for (int k = 0; k < CG_BLOCK_K / 16; k++) {
                                    umma_bf16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, first ? 0u : 1u);
                                    first = 0;
                                    if (NA == 2) umma_bf16(d_tmem, dl0 + 2 * k, db0 + 2 * k, idesc, 1);
                                }
                            }
                            umma_commit(&empty[st]);
                        }
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                    }
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

// Shared-memory plan for a patch launch; returns the number of stages (0: does not fit).
int conv_patch_plan(int block_n, int n_a, int wo, int ht, int kb, bool* wres_out, size_t* smem_out) {
    const int patch = (ht + 2) * wo * 128;
    const int b_bytes = block_n * CG_BLOCK_K * 2;
    const size_t budget = PA_CONV_SMEM_BUDGET - 1024 - 256;
    const size_t wres_bytes = (size_t)9 * kb * b_bytes;
    bool wres = wres_bytes <= 80 * 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
        const size_t stage = (size_t)n_a * patch + (wres ? 0 : 3 * (size_t)b_bytes);
        const size_t avail = budget - (wres ? wres_bytes : 0);
        int s = (int)(avail / stage);
        if (s > 8) s = 8;
        if (s >= 2) {
            *wres_out = wres;
            *smem_out = 1024 + (wres ? wres_bytes : 0) + stage * s + 256;
            return s;
        }
        wres = false;
    }
    return 0;
}

template <int BLOCK_N, int NA, bool WRES>
static int launch_p(const ConvMaps& maps, const ConvArgs& args, const PatchGeom& pg, size_t smem, int num_sms, cudaStream_t stream) {
    auto kern = conv_patch_kernel<BLOCK_N, NA, WRES>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PA_ERR_CUDA;
        attr_set = true;
    }
    int grid = args.m_tiles * args.n_tiles;
    if (grid > num_sms) grid = num_sms;
    if (launch_pdl(kern, dim3(grid), dim3(CG_THREADS), smem, stream, maps, args, pg) != cudaSuccess) return PA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv_patch(const ConvMaps& maps, const ConvArgs& args_in, int block_n, int n_a, int ht, bool wres, size_t smem,
                      int num_sms, cudaStream_t stream) {
    static int dbg = -1;
#ifdef PA_EXPERIMENT
    if (dbg < 0) { const char* e = getenv("PA_CONV_DEBUG"); dbg = e ? atoi(e) : 0; }
#else
    dbg = 0;   // the debug modes (wrong results by design) exist only in -DPA_EXPERIMENT builds
#endif
    ConvArgs args = args_in;
    args.debug = dbg;
    PatchGeom pg;
    pg.patch_bytes = (ht + 2) * args.wo * 128;
    pg.row_bytes = args.wo * 128;
#define PA_CP_CASE(BN, A, W) \
    if (block_n == BN && n_a == A && wres == W) return launch_p<BN, A, W>(maps, args, pg, smem, num_sms, stream);
    PA_CP_CASE(64, 1, true) PA_CP_CASE(64, 2, true) PA_CP_CASE(64, 1, false) PA_CP_CASE(64, 2, false)
    PA_CP_CASE(128, 1, false) PA_CP_CASE(128, 2, false)
#undef PA_CP_CASE
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
