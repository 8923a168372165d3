// cta_group::2 variant of the implicit-GEMM convolution (conv_gemm.cu) for the wide layers (Cout >= 256).
//
// Measured on the 1-CTA kernel (PA_CONV_DEBUG): with both operands in shared memory a
// 128x256x16 MMA takes ~167 clk instead of the 128-clk tensor floor, because every MMA re-reads 4 KB of A and 8 KB of B
// from the SM's shared memory. A CTA pair (two SMs of one TPC, a cluster of 2) computes a 256 x BLOCK_N tile: each CTA
// stages its own 128 rows of A but only HALF of the B tile (BLOCK_N/2 rows), and one `tcgen05.mma.cta_group::2` issued
// by the leader CTA drives both SMs' tensor cores, each reading 4 KB + 4 KB per k-step.
//
// Protocol per pipeline stage (barriers live at the same shared-memory offset in both CTAs):
//   producers (warp 0 of each CTA)  wait own empty[st] -> TMA own A tile + own half of B; the completion bytes of BOTH
//                                   CTAs are counted on the LEADER's full[st] (leader: arrive.expect_tx(2 x stage))
//   MMA (warp 1 of the leader)      wait full[st] -> 4 x umma2 -> commit.multicast -> empty[st] of both CTAs
//   per tile                        leader waits tempty[acc] (32 arrivals: the 16 epilogue warps of each CTA),
//                                   commit.multicast -> tfull[acc] of both CTAs; every CTA drains its own TMEM half.
#include "conv_common.cuh"

namespace pa {

template <int BLOCK_N, int NA>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(CG_MAX_REGS)
conv_gemm2_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_align_1k(smem_raw);
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr int B_BYTES = HALF_N * CG_BLOCK_K * 2;
    // Split precision (NA == 2) runs the K loop TWICE per tile: first every A_lo x B product, then every A_hi x B product,
    // into the same accumulator. tcgen05 truncates the fp32 accumulator at every k-step (a bias of -0.5 ulp of the running
    // sum per step, tests/test_gpu_layers.py::test_fp32_accumulation_floor_grows_with_k); interleaving hi and lo doubled
    // the number of truncating steps on the full-size sum, whereas the lo pass alone sums to ~2^-11 of it (its ulp, and
    // bias, are negligible). A stage holds ONE activation plane + the weight half-tile, so the weights stream twice.
    // Split weights as well (args.n_b == 2, the x3 modes): three passes, A_lo x B_hi, A_hi x B_lo, then A_hi x B_hi.
    constexpr int STAGE_BYTES = CG_A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
    const int S = args.num_stages;
    uint64_t* bars = (uint64_t*)(smem + (size_t)S * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + S;
    uint64_t* tfull = bars + 2 * S;
    uint64_t* tempty = bars + 2 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int taps = args.taps_h * args.taps_w;
    const int num_kb = taps * args.kb_per_tap;
    const int pair_m_tiles = (args.m_tiles + 1) >> 1;
    const int total_tiles = pair_m_tiles * args.n_tiles;   // pair tiles
    const int npass = NA + args.n_b - 1;

    if (warp == 0 && lane == 0) {
        for (int pl = 0; pl < NA; pl++)
            for (int q = 0; q < (args.stride == 2 ? 4 : 1); q++) tma_prefetch_desc(&maps.a[pl][q]);
        for (int pl = 0; pl < args.n_b; pl++) tma_prefetch_desc(&maps.bh[pl]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; i++) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 2 * CG_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();            // both CTAs' barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) pdl_launch_dependents();
    if (warp != 1) pdl_wait();     // producer and epilogue warps touch the previous layer's tensors; the MMA warp only shared memory

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int st = 0; uint32_t ph = 0;
        const int pix_per_img = args.ho * args.wo;
        for (int tile = cluster_id; tile < total_tiles; tile += n_clusters) {
            const int pm = tile / args.n_tiles, nt = tile - pm * args.n_tiles;
            const int mt = 2 * pm + (int)rank;
            const int m0 = mt * CG_BLOCK_M;
            const int n0 = m0 / pix_per_img;
            const int oy0 = (m0 - n0 * pix_per_img) / args.wo;
            for (int pass = 0; pass < npass; pass++)        // residual products first, hi x hi last
            for (int tap = 0; tap < taps; tap++) {
                const int pla = (pass < NA - 1) ? 1 : 0, plb = (pass >= NA - 1 && pass < npass - 1) ? 1 : 0;
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
                    uint8_t* sb = sa + CG_A_BYTES;
                    if (elect_one()) {
                        if (rank == 0) mbar_arrive_expect_tx(&full[st], 2 * STAGE_BYTES);
                        const uint32_t lead_full = mapa_u32(&full[st], 0);
                        tma2_load_4d(sa, &maps.a[pla][q], lead_full, kc * CG_BLOCK_K, cx, cy, n0);
                        tma2_load_2d(sb, &maps.bh[plb], lead_full, tap * args.k_per_tap + kc * CG_BLOCK_K, nt * BLOCK_N + (int)rank * HALF_N);
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
            int st = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = cluster_id; tile < total_tiles; tile += n_clusters, it++) {
                const int acc = it & 1;
                const uint32_t acc_ph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < npass * num_kb; kb++) {   // every pass over K goes into the one accumulator
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)st * STAGE_BYTES);
                    const uint32_t sb = sa + CG_A_BYTES;
                    if (elect_one()) {
                        const uint64_t da0 = umma_desc_sw128(sa), db0 = umma_desc_sw128(sb);
#pragma unroll
                        for (int k = 0; k < CG_BLOCK_K / 16; k++) umma2_f16(d_tmem, da0 + 2 * k, db0 + 2 * k, idesc, (kb | k) != 0);
                        umma2_commit_multicast(&empty[st]);
                    }
                    __syncwarp();
                    if (++st == S) { st = 0; ph ^= 1; }
                }
                if (elect_one()) umma2_commit_multicast(&tfull[acc]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (warps 2..17 of both CTAs) =====================
        const bool split_out = args.out_lo != nullptr;
        if (args.f16) {
            if (split_out) epilogue<BLOCK_N, true, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, true, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        } else {
            if (split_out) epilogue<BLOCK_N, false, true, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
            else epilogue<BLOCK_N, false, false, true>(args, tfull, tempty, tmem_base, warp, lane, total_tiles);
        }
    }
    tc_fence_before();
    cluster_sync_all();            // neither CTA frees TMEM / exits while its peer may still signal or read
    if (warp == 2) tmem_dealloc2<TMEM_COLS>(tmem_base);
}

int conv_gemm2_pick_stages(int block_n, int n_a) {
    (void)n_a;      // a stage holds one activation plane: split precision makes more passes over K instead of bigger stages
    const size_t stage = (size_t)CG_A_BYTES + (size_t)(block_n / 2) * CG_BLOCK_K * 2;
    int s = (int)((PA_CONV_SMEM_BUDGET - 1024 - 256) / stage);
    return s > 8 ? 8 : s;
}

template <int BLOCK_N, int NA>
static int launch2_t(const ConvMaps& maps, const ConvArgs& args, int num_sms, cudaStream_t stream) {
    auto kern = conv_gemm2_kernel<BLOCK_N, NA>;
    const size_t stage = (size_t)CG_A_BYTES + (size_t)(BLOCK_N / 2) * CG_BLOCK_K * 2;
    const size_t smem = 1024 + stage * args.num_stages + 256;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return PA_ERR_CUDA;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    const int pair_tiles = ((args.m_tiles + 1) / 2) * args.n_tiles;
    int clusters = num_sms / 2;
    if (clusters > pair_tiles) clusters = pair_tiles;
    if (launch_pdl(kern, dim3(2 * clusters), dim3(CG_THREADS), smem, stream, maps, args) != cudaSuccess) return PA_ERR_CUDA;   // cluster shape comes from __cluster_dims__
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_conv_gemm2(const ConvMaps& maps, const ConvArgs& args_in, int block_n, int n_a, int num_sms, cudaStream_t stream) {
    ConvArgs args = args_in;
    args.debug = 0;
    if (args.n_b < 1) args.n_b = 1;
    if (block_n == 256 && n_a == 1) return launch2_t<256, 1>(maps, args, num_sms, stream);
    if (block_n == 256 && n_a == 2) return launch2_t<256, 2>(maps, args, num_sms, stream);
    return PA_ERR_UNSUPPORTED;
}

}  // namespace pa
