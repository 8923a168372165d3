// Shared by the tcgen05 convolution kernels (conv_gemm.cu, conv_patch.cu): tile constants and the
// 16-warp TMEM epilogue.
#pragma once
#include <cstdlib>
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

constexpr int CG_FIRST_EPI_WARP = 2;
constexpr int CG_EPI_WARPS = 16;
// Register cap of the 18-warp conv CTAs. An SM sub-partition (16,384 registers) hosts 5 of those warps: at 88
// registers they take 14,080 and leave room for one 64-register warp of the window-staging kernel, which must stay
// resident beside them (at the natural 96 the conv CTA would have to wait for the staging kernel to drain).
constexpr int CG_MAX_REGS = 88;
constexpr int CG_THREADS = (CG_FIRST_EPI_WARP + CG_EPI_WARPS) * 32;  // TMA warp, MMA warp, 16 epilogue warps
// Dynamic shared memory the conv kernels may plan with: 220 KB of the 228 KB per SM, so that one CTA of the
// window-staging kernel (a 5.5 KB TMA ring + 1 KB of reserved shared memory) can stay resident beside a conv CTA.
inline size_t conv_smem_budget() {
    static size_t v = 0;
    if (!v) {
        int kb = 220;
#ifdef PA_EXPERIMENT
        const char* e = getenv("PA_CONV_SMEM_KB");
        if (e && atoi(e) >= 64 && atoi(e) <= 227) kb = atoi(e);
#endif
        v = (size_t)kb * 1024;
    }
    return v;
}
#define PA_CONV_SMEM_BUDGET conv_smem_budget()
constexpr int CG_BLOCK_M = 128;
constexpr int CG_BLOCK_K = 64;
constexpr int CG_A_BYTES = CG_BLOCK_M * CG_BLOCK_K * 2;  // 16 KB

// Epilogue of one CTA: TMEM accumulator -> fp32 scale/shift (folded BN or bias) -> (+ residual) -> (ReLU)
// -> 16-bit hi (+ lo) planes or fp32. Sixteen warps: four per TMEM lane quarter, each owning every
// fourth 16-column chunk, so TMEM / global latencies of one warp hide behind the others. One thread per
// accumulator row (= output pixel); its channels are contiguous in NHWC, so every access is a 32-byte
// run. The residual of a warp's next chunk is prefetched while the current one is processed.
template <bool F16>
__device__ __forceinline__ uint32_t pack16x2(float a, float b) {
    if (F16) {
        const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
        return *reinterpret_cast<const uint32_t*>(&h);
    }
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// PAIR = true: the CTA is one half of a cta_group::2 pair (conv_gemm2.cu). Tiles are walked per cluster, this CTA
// owns m-tile 2*pm + rank of pair tile pm, and "accumulator drained" is signalled to the LEADER CTA's barrier.
template <int BLOCK_N, bool F16, bool SPLIT, bool PAIR = false>
__device__ __forceinline__ void epilogue(const ConvArgs& args, uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base,
                                         int warp, int lane, int total_tiles) {
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int wq = (warp - CG_FIRST_EPI_WARP) >> 2;  // which of the quarter's four warps
    const int r = q * 32 + lane;
    const bool has_res = args.res_hi != nullptr;
    const bool has_res_lo = args.res_lo != nullptr;
    int it = 0;
    const int t_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int pair_rank = PAIR ? (int)(blockIdx.x & 1) : 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step, it++) {
        const int mt0 = tile / args.n_tiles, nt = tile - mt0 * args.n_tiles;
        const int mt = PAIR ? 2 * mt0 + pair_rank : mt0;
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        const int64_t row = (int64_t)mt * CG_BLOCK_M + r;
        const bool row_ok = row < args.m_total;
        const int n_base = nt * BLOCK_N;
        const int64_t o_base = row * args.cout + n_base;
        uint4 rh0, rh1, rl0, rl1;  // residual of the current chunk
        rh0 = rh1 = rl0 = rl1 = make_uint4(0, 0, 0, 0);
        auto load_res = [&](int c0, uint4& a0, uint4& a1, uint4& b0, uint4& b1) {
            if (has_res && row_ok && c0 < BLOCK_N && n_base + c0 + 16 <= args.cout) {
                if (((o_base + c0) & 15) == 0) {
                    ldg_nc_256(args.res_hi + o_base + c0, a0, a1);
                    if (has_res_lo) ldg_nc_256(args.res_lo + o_base + c0, b0, b1);
                } else {
                    const uint4* rp = (const uint4*)(args.res_hi + o_base + c0);
                    a0 = __ldg(rp); a1 = __ldg(rp + 1);
                    if (has_res_lo) {
                        const uint4* lp = (const uint4*)(args.res_lo + o_base + c0);
                        b0 = __ldg(lp); b1 = __ldg(lp + 1);
                    }
                }
            }
        };
        if (args.debug & 1) {   // experiment: how fast is the kernel without an epilogue?
            mbar_wait(&tfull[acc], acc_ph);
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            continue;
        }
        load_res(wq * 16, rh0, rh1, rl0, rl1);
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
#pragma unroll 1
        for (int c0 = wq * 16; c0 < BLOCK_N; c0 += 64) {
            float v[16];
            __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after per-row predication
            tmem_ld16(t_addr + c0, v);
            uint4 nh0, nh1, nl0, nl1;
            nh0 = nh1 = nl0 = nl1 = make_uint4(0, 0, 0, 0);
            load_res(c0 + 64, nh0, nh1, nl0, nl1);
            const int n = n_base + c0;
            if (n < args.cout && row_ok) {
                const bool full16 = (n + 16 <= args.cout);
                if (args.scale) {
                    if (full16) {
                        const float4* sp4 = (const float4*)(args.scale + n);
#pragma unroll
                        for (int i = 0; i < 4; i++) { const float4 t = __ldg(sp4 + i); v[4 * i] *= t.x; v[4 * i + 1] *= t.y; v[4 * i + 2] *= t.z; v[4 * i + 3] *= t.w; }
                    } else {
                        for (int i = 0; i < 16; i++) if (n + i < args.cout) v[i] *= __ldg(args.scale + n + i);
                    }
                }
                if (args.shift) {
                    if (full16) {
                        const float4* sp4 = (const float4*)(args.shift + n);
#pragma unroll
                        for (int i = 0; i < 4; i++) { const float4 t = __ldg(sp4 + i); v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w; }
                    } else {
                        for (int i = 0; i < 16; i++) if (n + i < args.cout) v[i] += __ldg(args.shift + n + i);
                    }
                }
                if (has_res && full16) {
                    const uint32_t w[8] = {rh0.x, rh0.y, rh0.z, rh0.w, rh1.x, rh1.y, rh1.z, rh1.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        v[2 * i] += dec16<F16>((uint16_t)(w[i] & 0xFFFF));
                        v[2 * i + 1] += dec16<F16>((uint16_t)(w[i] >> 16));
                    }
                    if (has_res_lo) {
                        const uint32_t x[8] = {rl0.x, rl0.y, rl0.z, rl0.w, rl1.x, rl1.y, rl1.z, rl1.w};
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            v[2 * i] += dec16<F16>((uint16_t)(x[i] & 0xFFFF));
                            v[2 * i + 1] += dec16<F16>((uint16_t)(x[i] >> 16));
                        }
                    }
                }
                if (args.relu) {
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
                }
                const int64_t o = o_base + c0;
                if (args.out_f32) {
                    if (full16 && ((o & 3) == 0)) {
                        float4* op = (float4*)(args.out_f32 + o);
#pragma unroll
                        for (int i = 0; i < 4; i++) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
                        for (int i = 0; i < 16; i++) if (n + i < args.cout) args.out_f32[o + i] = v[i];
                    }
                }
                if (args.out_hi) {
                    if (full16 && ((o & 7) == 0)) {
                        uint32_t h[8], l[8];
                        if (SPLIT) {
#pragma unroll
                            for (int i = 0; i < 8; i++) split2<F16>(v[2 * i], v[2 * i + 1], h[i], l[i]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; i++) h[i] = pack16x2<F16>(v[2 * i], v[2 * i + 1]);
                        }
                        if ((o & 15) == 0) {
                            stg_256(args.out_hi + o, make_uint4(h[0], h[1], h[2], h[3]), make_uint4(h[4], h[5], h[6], h[7]));
                            if (SPLIT) stg_256(args.out_lo + o, make_uint4(l[0], l[1], l[2], l[3]), make_uint4(l[4], l[5], l[6], l[7]));
                        } else {
                            uint4* op = (uint4*)(args.out_hi + o);
                            op[0] = make_uint4(h[0], h[1], h[2], h[3]);
                            op[1] = make_uint4(h[4], h[5], h[6], h[7]);
                            if (SPLIT) {
                                uint4* lp = (uint4*)(args.out_lo + o);
                                lp[0] = make_uint4(l[0], l[1], l[2], l[3]);
                                lp[1] = make_uint4(l[4], l[5], l[6], l[7]);
                            }
                        }
                    } else {
                        uint16_t* oh = (uint16_t*)args.out_hi;
                        uint16_t* ol = (uint16_t*)args.out_lo;
                        for (int i = 0; i < 16; i++) if (n + i < args.cout) {
                            const uint16_t h0 = enc16<F16>(v[i]);
                            oh[o + i] = h0;
                            if (SPLIT) ol[o + i] = enc16<F16>(v[i] - dec16<F16>(h0));
                        }
                    }
                }
            }
            rh0 = nh0; rh1 = nh1; rl0 = nl0; rl1 = nl1;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(&tempty[acc], 0);   // the leader's MMA warp waits for both halves
            else mbar_arrive(&tempty[acc]);
        }
    }
}

// Fast path for the Cout = 64 layers (layer1): one 16-column chunk per warp for the whole kernel, so scale/shift
// live in registers, the residual is requested before the accumulator is waited for, and the TMEM buffer is
// handed back to the MMA warp as soon as it has been read (the stores overlap the next tile's MMAs).
template <bool F16, bool SPLIT, bool PAIR = false>
__device__ __forceinline__ void epilogue_n64(const ConvArgs& args, uint64_t* tfull, uint64_t* tempty, uint32_t tmem_base,
                                             int warp, int lane, int total_tiles) {
    const int q = warp & 3;
    const int c0 = ((warp - CG_FIRST_EPI_WARP) >> 2) * 16;
    const int r = q * 32 + lane;
    float sc[16], sh[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 a = __ldg((const float4*)(args.scale + c0) + i), b = __ldg((const float4*)(args.shift + c0) + i);
        sc[4 * i] = a.x; sc[4 * i + 1] = a.y; sc[4 * i + 2] = a.z; sc[4 * i + 3] = a.w;
        sh[4 * i] = b.x; sh[4 * i + 1] = b.y; sh[4 * i + 2] = b.z; sh[4 * i + 3] = b.w;
    }
    const bool has_res = args.res_hi != nullptr;
    const bool has_res_lo = args.res_lo != nullptr;
    const bool relu = args.relu != 0;
    int it = 0;
    const int t_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, t_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int pair_rank = PAIR ? (int)(blockIdx.x & 1) : 0;
    for (int tile = t_first; tile < total_tiles; tile += t_step, it++) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        const int mt = PAIR ? 2 * tile + pair_rank : tile;          // total_tiles counts pair tiles in PAIR mode
        const int64_t o = ((int64_t)mt * CG_BLOCK_M + r) * 64 + c0;
        uint4 rh0 = make_uint4(0, 0, 0, 0), rh1 = rh0, rl0 = rh0, rl1 = rh0;
        if (has_res) {
            ldg_nc_256(args.res_hi + o, rh0, rh1);     // o is a multiple of 16 elements: 32-byte aligned
            if (has_res_lo) ldg_nc_256(args.res_lo + o, rl0, rl1);
        }
        mbar_wait(&tfull[acc], acc_ph);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 64 + c0, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if (PAIR) mbar_arrive_cluster(&tempty[acc], 0);
            else mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fmaf(v[i], sc[i], sh[i]);
        if (has_res) {
            const uint32_t w[8] = {rh0.x, rh0.y, rh0.z, rh0.w, rh1.x, rh1.y, rh1.z, rh1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) { v[2 * i] += dec16<F16>((uint16_t)(w[i] & 0xFFFF)); v[2 * i + 1] += dec16<F16>((uint16_t)(w[i] >> 16)); }
            if (has_res_lo) {
                const uint32_t x[8] = {rl0.x, rl0.y, rl0.z, rl0.w, rl1.x, rl1.y, rl1.z, rl1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) { v[2 * i] += dec16<F16>((uint16_t)(x[i] & 0xFFFF)); v[2 * i + 1] += dec16<F16>((uint16_t)(x[i] >> 16)); }
            }
        }
        if (relu) {
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = fmaxf(v[i], 0.f);
        }
        uint32_t h[8], l[8];
        if (SPLIT) {
#pragma unroll
            for (int i = 0; i < 8; i++) split2<F16>(v[2 * i], v[2 * i + 1], h[i], l[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) h[i] = pack16x2<F16>(v[2 * i], v[2 * i + 1]);
        }
        stg_256(args.out_hi + o, make_uint4(h[0], h[1], h[2], h[3]), make_uint4(h[4], h[5], h[6], h[7]));
        if (SPLIT) stg_256(args.out_lo + o, make_uint4(l[0], l[1], l[2], l[3]), make_uint4(l[4], l[5], l[6], l[7]));
    }
}

// true when epilogue_n64 applies
__device__ __forceinline__ bool epilogue_n64_ok(const ConvArgs& a) {
    return a.n_tiles == 1 && a.cout == 64 && a.scale && a.shift && a.out_hi && !a.out_f32 && (a.m_total % CG_BLOCK_M) == 0 && !a.debug;
}

}  // namespace pa
