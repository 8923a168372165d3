// cta_group::2 variant of the patch-mode 3x3 / stride-1 convolution (conv_patch.cu) for layer1 (N = 64) and layer2
// (N = 128): with both operands in shared memory these narrow-N MMAs are bound by the per-SM operand reads
// (96 clk per 128x64x16 MMA against a 32-clk tensor floor). A CTA pair computes a 256-pixel tile;
// each CTA stages the patches of its own 128 pixels and HALF of every weight tile (BLOCK_N/2 rows), the leader issues
// `tcgen05.mma.cta_group::2`. Barrier protocol as in conv_gemm2.cu; patch / descriptor arithmetic as in conv_patch.cu.
#include "conv_common.cuh"

namespace pa {

struct PatchGeom2 {
    int patch_bytes;   // (Ht + 2) * W * 128
    int row_bytes;     // W * 128: displacement of one vertical tap
};

template <int BLOCK_N, int NA, bool WRES>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(CG_MAX_REGS)
conv_patch2_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args, const PatchGeom2 pg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1k(smem_raw);
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr int B_BYTES = HALF_N * CG_BLOCK_K * 2;       // this CTA's half of one [BLOCK_N x 64] weight tile
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
    const int S = args.num_stages;
    const int kb = args.kb_per_tap;
    // split precision: NA + n_b - 1 passes over K per tile (residual products first, hi x hi last) into one accumulator, one
    // activation plane per stage -- see conv_gemm2.cu
    const int stage_bytes = pg.patch_bytes + (WRES ? 0 : 3 * B_BYTES);
    const int wres_plane = 9 * kb * B_BYTES;
    const int wres_bytes = WRES ? args.n_b * wres_plane : 0;
    const int npass = NA + args.n_b - 1;                   // split weights too (x3 modes): A_lo x B_hi, A_hi x B_lo, A_hi x B_hi
    uint8_t* wres = smem;                                  // resident half weights: [plane][tap][kc][HALF_N x 128 B]
    uint8_t* stages = smem + wres_bytes;
    uint64_t* bars = (uint64_t*)(stages + (size_t)S * stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint64_t* wfull = bars + 2 * S + 4;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int total_tiles = (args.m_tiles + 1) >> 1;       // pair tiles (n_tiles == 1 for these layers)

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < NA; pl++) tma_prefetch_desc(&maps.a[pl][0]);
        for (int pl = 0; pl < args.n_b; pl++) tma_prefetch_desc(&maps.bh[pl]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * CG_EPI_WARPS); }
        mbar_init(wfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (warp >= CG_FIRST_EPI_WARP) pdl_wait();   // epilogue warps read the residual / write the output of earlier layers' buffers

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (WRES && elect_one()) {   // both halves of the resident weights are counted on the leader's barrier
            if (rank == 0) mbar_arrive_expect_tx(wfull, 2 * wres_bytes);
            const uint32_t lead_w = mapa_u32(wfull, 0);
            for (int pl = 0; pl < args.n_b; pl++)
                for (int tap = 0; tap < 9; tap++)
                    for (int kc = 0; kc < kb; kc++)
                        tma2_load_2d(wres + (size_t)pl * wres_plane + (size_t)(tap * kb + kc) * B_BYTES, &maps.bh[pl], lead_w,
                                     tap * args.k_per_tap + kc * CG_BLOCK_K, (int)rank * HALF_N);
        }
        __syncwarp();
        pdl_wait();      // the resident weights above do not depend on the previous layer; the activation patches do
        int st = 0; uint32_t ph = 0;
        const int pix_per_img = args.ho * args.wo;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int m0 = (2 * tile + (int)rank) * CG_BLOCK_M;
            const int n0 = m0 / pix_per_img;
            const int oy0 = (m0 - n0 * pix_per_img) / args.wo;
            for (int pass = 0; pass < npass; pass++)        // residual products first, hi x hi last
            for (int kc = 0; kc < kb; kc++) {
                const int pla = (pass < NA - 1) ? 1 : 0, plb = (pass >= NA - 1 && pass < npass - 1) ? 1 : 0;
                for (int dxi = 0; dxi < 3; dxi++) {
                    mbar_wait(&empty[st], ph ^ 1);
                    uint8_t* sa = stages + (size_t)st * stage_bytes;
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * stage_bytes);
                        const uint32_t lead_full = mapa_u32(&full[st], 0);
                        tma2_load_4d(sa, &maps.a[pla][0], lead_full, kc * CG_BLOCK_K, dxi - 1, oy0 - 1, n0);
                        if (!WRES) {
                            uint8_t* sb = sa + pg.patch_bytes;
#pragma unroll
                            for (int dy = 0; dy < 3; dy++)
                                tma2_load_2d(sb + dy * B_BYTES, &maps.bh[plb], lead_full, (dy * 3 + dxi) * args.k_per_tap + kc * CG_BLOCK_K,
                                             (int)rank * HALF_N);
                        }
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            const uint32_t idesc = args.f16 ? umma_idesc_f16(2 * CG_BLOCK_M, BLOCK_N) : umma_idesc_bf16(2 * CG_BLOCK_M, BLOCK_N);
            if (WRES) { mbar_wait(wfull, 0); tc_fence_after(); }
            const uint32_t wres_u = smem_u32(wres);
            int st = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                uint32_t first = 1;
                for (int pass = 0; pass < npass; pass++)
                for (int kc = 0; kc < kb; kc++) {
                    const uint32_t wres_p = wres_u + (uint32_t)((pass >= NA - 1 && pass < npass - 1) ? wres_plane : 0);
                    for (int dxi = 0; dxi < 3; dxi++) {
                        mbar_wait(&full[st], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stages + (size_t)st * stage_bytes);
                        const uint32_t sb = sa + pg.patch_bytes;
                        if (elect_one()) {
#pragma unroll
                            for (int dy = 0; dy < 3; dy++) {
                                const uint32_t a0 = sa + dy * pg.row_bytes;
                                const uint32_t b0 = WRES ? wres_p + (uint32_t)(((dy * 3 + dxi) * kb + kc) * B_BYTES) : sb + dy * B_BYTES;
                                const uint64_t da0 = umma_desc_sw128(a0), db0 = umma_desc_sw128(b0);
#pragma unroll
                                for (int k = 0; k < CG_BLOCK_K / 16; k++) {
                                    umma2_f16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, first ? 0u : 1u);
                                    first = 0;
                                }
                            }
                            umma2_commit_multicast(&empty[st]);
                        }
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                    }
                }
                if (elect_one()) umma2_commit_multicast(&tfull[acc]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..17 of both CTAs) =====================
        const bool split_out = args.out_lo != nullptr;
        if (BLOCK_N == 64 && epilogue_n64_ok(args)) {
            if (args.f16) {
                if (split_out) epilogue_n64<true, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
                else epilogue_n64<true, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            } else {
                if (split_out) epilogue_n64<false, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
                else epilogue_n64<false, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            }
        } else if (args.f16) {
            if (split_out) epilogue<BLOCK_N, true, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, true, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        } else {
            if (split_out) epilogue<BLOCK_N, false, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, false, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) tmem_dealloc2<TMEM_COLS>(tmem_base);
}

// Shared-memory plan of a pair launch; returns the number of stages (0: does not fit).
int conv_patch2_plan(int block_n, int n_a, int n_b, int wo, int ht, int kb, bool* wres_out, size_t* smem_out) {
    const int patch = (ht + 2) * wo * 128;
    const int b_bytes = (block_n / 2) * CG_BLOCK_K * 2;
    const size_t budget = PA_CONV_SMEM_BUDGET - 1024 - 256;
    const size_t wres_bytes = (size_t)n_b * 9 * kb * b_bytes;
    bool wres = wres_bytes <= 80 * 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
        (void)n_a;      // one activation plane per stage (split precision: more passes over K, not bigger stages)
        const size_t stage = (size_t)patch + (wres ? 0 : 3 * (size_t)b_bytes);
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
static int launch_p2(const ConvMaps& maps, const ConvArgs& args, const PatchGeom2& pg, size_t smem, int num_sms, cudaStream_t stream) {
    auto kern = conv_patch2_kernel<BLOCK_N, NA, WRES>;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PA_ERR_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    const int pair_tiles = (args.m_tiles + 1) / 2;
    int clusters = num_sms / 2;
    if (clusters > pair_tiles) clusters = pair_tiles;
    if (launch_pdl(kern, dim3(2 * clusters), dim3(CG_THREADS), smem, stream, maps, args, pg) != cudaSuccess) return PA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv_patch2(const ConvMaps& maps, const ConvArgs& args_in, int block_n, int n_a, int ht, bool wres, size_t smem,
                       int num_sms, cudaStream_t stream) {
    ConvArgs args = args_in;
    args.debug = 0;
    if (args.n_b < 1) args.n_b = 1;
    PatchGeom2 pg;
    pg.patch_bytes = (ht + 2) * args.wo * 128;
    pg.row_bytes = args.wo * 128;
#define PA_CP2_CASE(BN, A, W) \
    if (block_n == BN && n_a == A && wres == W) return launch_p2<BN, A, W>(maps, args, pg, smem, num_sms, stream);
    PA_CP2_CASE(64, 1, true) PA_CP2_CASE(64, 2, true) PA_CP2_CASE(64, 1, false) PA_CP2_CASE(64, 2, false)
    PA_CP2_CASE(128, 1, true) PA_CP2_CASE(128, 2, true) PA_CP2_CASE(128, 1, false) PA_CP2_CASE(128, 2, false)
#undef PA_CP2_CASE
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
