// Fused fighter-crop preprocessing for sm_100a.
//
// One launch turns every (frame, fighter) box into a normalised 128x128 crop, with exactly the
// arithmetic of the reference chain:
//   YoloCrop.square_crop (playaid/fighter.py:323-381)
//     numpy window -> PIL.ImageOps.pad (Pillow BICUBIC, 22-bit fixed point, u8 between passes,
//     centred black letterbox) -> imutils.resize == cv2.resize(INTER_AREA) (copy / integer-scale /
//     fp32 area tables / 11-bit linear upscale) -> 127-row letterbox
//   then BGR->RGB, HWC->CHW, /255, (x-mean)/std, cast
//     (playaid/ult_action_dataset.py:302,349-359; playaid/ai_runner.py:448,461-463).
//
// Work decomposition: grid = n_crops x PA_SPLIT; each CTA owns a slab of final output rows and
// walks it in bands. A band is stateless: it loads the raw source rows it needs with 128-bit
// reads into shared memory, runs the horizontal and vertical bicubic passes and the area pass
// entirely out of shared memory, and stores final values once. Coefficient tables are built
// per CTA in fp64 with contraction disabled (this file is compiled with -fmad=false).
//
// HBM traffic is the window bytes (+ a small vertical halo per band) and the output; the
// kernel's arithmetic is ~8 integer MACs per source byte, so it is issue-bound, not HBM-bound
// ("preprocess roofline").
#include "pa_internal.cuh"
#include "ptx.cuh"

namespace pa {

constexpr int PP_MAX_THREADS = 384;
constexpr int PP_SPLIT = 4;

// ---- tensor-core path (preprocess_tc.inc): tile constants shared with the plan kernel
constexpr int TC_MAX_STRIPS = 128;
constexpr int TC_NCOLS = 21;         // canvas columns per strip: 21 x 3 channels = 63 <= 64 accumulator columns per digit
constexpr int TC_TT_ROWS = 512;      // raw rows one work item may span (four 128-row blocks)
constexpr int TC_XB_FLOATS = 7936;   // fp32 x-pass buffer of one work item (31 KB)
constexpr int TC_KSPAN = 128;        // bytes of K one coefficient tile covers (one SWIZZLE_128B atom)
constexpr int TC_MAX_D = 24;         // final columns per strip (x-table slots in shared memory)
constexpr int TC_XT_CAP = 16;        // area-table entries per final column the strip tables hold (scale_x <= 14)
constexpr int TC_TILE_BYTES = 192 * 128;   // one vertical coefficient tile: 3 digit planes x [64 rows][128 B], the kernel's shared-memory image
// byte offset of (row, k) inside a K-major SWIZZLE_128B tile with 128-byte rows (8-row atoms of 1 KB)
__device__ __forceinline__ int sw128(int row, int k) { return row * 128 + ((((k >> 4) ^ (row & 7)) << 4) | (k & 15)); }
// balanced base-256 digits of a fixed-point coefficient: k = d0 + 256 d1 + 65536 d2, d0, d1 in [-128, 127]
__device__ __forceinline__ void coef_digits(int k, int& d0, int& d1, int& d2) {
    d0 = ((k + 128) & 255) - 128;
    const int k1 = (k - d0) >> 8;
    d1 = ((k1 + 128) & 255) - 128;
    d2 = (k1 - d1) >> 8;
}

enum { REG_COPY = 0, REG_FAST = 1, REG_GENERAL = 2, REG_LINEAR = 3 };

struct CropGeom {
    int status, frame;
    int x0, y0, rw, rh;  // raw window in the frame
    int sd;              // square_dim
    int pad1;            // first ImageOps.pad active
    int nw, nh, ox, oy;  // resized size and paste offset inside the sd x sd canvas
    int hact, vact;      // bicubic passes active
    int oh, oy2;         // area output rows, final paste row offset
    int regime, isx, isy;
    double scale_x, scale_y, inv_scale_x, inv_scale_y;
    double h_scale, h_fs, h_sup;
    double v_scale, v_fs, v_sup;
    int h_ks, v_ks;
    // per-crop coefficient tables precomputed by preprocess_plan_kernel (int32 offsets into the crop's block)
    int tab_ok, off_h, off_v, off_x, off_y;
    // routing (preprocess_plan_kernel): 1 = the tensor-core kernel (preprocess_tc.inc) computes this crop as
    // n_strips x n_parts work items, 0 = the streaming CUDA-core kernel below
    int route, n_strips, n_parts, rv;
    int off_tc;                            // strip / part records of the tensor-core kernel inside the table block
    int tile_first[4];                     // per part: index of its first vertical coefficient tile in the tile pool
};

__device__ __forceinline__ double cubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

__device__ __forceinline__ int rint_d(double v) { return __double2int_rn(v); }  // Python round()

// ImageOps.contain target size
__device__ void contain_size(int w, int h, int sw, int sh, int& nw, int& nh) {
    double im_ratio = (double)w / (double)h, dest_ratio = (double)sw / (double)sh;
    nw = sw; nh = sh;
    if (im_ratio != dest_ratio) {
        if (im_ratio > dest_ratio) {
            int new_h = rint_d((double)h / (double)w * (double)sw);
            if (new_h != sh) nh = new_h;
        } else {
            int new_w = rint_d((double)w / (double)h * (double)sh);
            if (new_w != sw) nw = new_w;
        }
    }
}

__device__ void bicubic_axis(int in_size, int out_size, double& scale, double& fs, double& sup, int& ks) {
    scale = (double)((float)in_size - 0.0f) / (double)out_size;
    fs = scale < 1.0 ? 1.0 : scale;
    sup = 2.0 * fs;
    ks = (int)ceil(sup) * 2 + 1;
}

// bounds + fixed-point coefficients of one output index (Pillow precompute_coeffs + normalize_coeffs_8bpc)
__device__ void bicubic_coeffs(int xx, int in_size, double scale, double fs, double sup, int ks, int& xmin_out,
                               int& n_out, int32_t* kk /*[ks]*/) {
    double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    double ss = 1.0 / fs;
    int xmin = (int)(center - sup + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + sup + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    // two passes so no fp64 scratch array is needed: the sum is taken in the same order
    for (int x = 0; x < xmax; x++) ww += cubic((x + xmin - center + 0.5) * ss);
    for (int x = 0; x < ks; x++) {
        int32_t k = 0;
        if (x < xmax) {
            double w = cubic((x + xmin - center + 0.5) * ss);
            if (ww != 0.0) w /= ww;
            k = (w < 0) ? (int32_t)(-0.5 + w * (double)(1 << 22)) : (int32_t)(0.5 + w * (double)(1 << 22));
        }
        kk[x] = k;
    }
    xmin_out = xmin;
    n_out = xmax;
}

__device__ __forceinline__ uint8_t clip8_fix(int32_t v) {
    v >>= 22;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
__device__ __forceinline__ int sat_u8f(float v) {
    int iv = __float2int_rn(v);
    return iv < 0 ? 0 : (iv > 255 ? 255 : iv);
}
__device__ __forceinline__ int sat_s16f(float v) {
    int iv = __float2int_rn(v);
    return iv < -32768 ? -32768 : (iv > 32767 ? 32767 : iv);
}

__device__ void compute_geom(CropGeom& g, const int32_t* box, int H, int W, int n_frames, int out, int padding) {
    g.frame = box[0];
    int cx = box[1], cy = box[2], cw = box[3], ch = box[4];
    g.status = PA_CROP_OK;
    int sd = cw > ch ? cw : ch;
    g.sd = sd;
    if (g.frame < 0 || g.frame >= n_frames || sd < 0) { g.status = PA_CROP_INVALID; return; }
    int x0, y0, rw, rh;
    crop_window(cx, cy, sd, H, W, padding, x0, y0, rw, rh);
    g.x0 = x0; g.y0 = y0; g.rw = rw; g.rh = rh;
    g.pad1 = (rh != sd || rw != sd);
    g.nw = sd; g.nh = sd; g.ox = 0; g.oy = 0; g.hact = 0; g.vact = 0;
    if (g.pad1) {
        // ImageOps.contain divides width / height and size[0] / size[1]: ZeroDivisionError escapes
        if (rh == 0 || sd == 0) { g.status = PA_CROP_ZERO_DIV; return; }
        int nw, nh;
        contain_size(rw, rh, sd, sd, nw, nh);
        // Image.resize copies when the size is unchanged (even an empty image -> black canvas);
        // otherwise an empty target raises ValueError -> (False, None)
        if (!(nw == rw && nh == rh) && (nw <= 0 || nh <= 0)) { g.status = PA_CROP_INVALID; return; }
        g.nw = nw; g.nh = nh;
        if (!(nw == sd && nh == sd)) {
            if (nw != sd) g.ox = rint_d((double)(sd - nw) * 0.5);
            else g.oy = rint_d((double)(sd - nh) * 0.5);
        }
        g.hact = (nw != rw);
        g.vact = (nh != rh);
        if (g.hact) bicubic_axis(rw, nw, g.h_scale, g.h_fs, g.h_sup, g.h_ks);
        if (g.vact) bicubic_axis(rh, nh, g.v_scale, g.v_fs, g.v_sup, g.v_ks);
    }
    if (sd == 0) { g.status = PA_CROP_INVALID; return; }
    if (rw > PA_MAX_WINDOW || sd > PA_MAX_WINDOW) { g.status = PA_CROP_TOO_LARGE; return; }
    int oh = (int)((double)sd * ((double)out / (double)sd));
    if (oh <= 0) { g.status = PA_CROP_INVALID; return; }
    g.oh = oh;
    g.oy2 = 0;
    if (oh != out) {
        // second ImageOps.pad: (out x oh) -> (out, out); contain keeps the size, paste is centred
        int nw2, nh2;
        contain_size(out, oh, out, out, nw2, nh2);
        if (nw2 != out || nh2 != oh) { g.status = PA_CROP_TOO_LARGE; return; }  // would need a third resample
        g.oy2 = rint_d((double)(out - oh) * 0.5);
    }
    // cv2.resize(INTER_AREA) dispatch, source sd x sd -> out x oh
    if (oh == sd && out == sd) { g.regime = REG_COPY; return; }
    g.inv_scale_x = (double)out / (double)sd;
    g.inv_scale_y = (double)oh / (double)sd;
    g.scale_x = 1. / g.inv_scale_x;
    g.scale_y = 1. / g.inv_scale_y;
    g.isx = rint_d(g.scale_x);
    g.isy = rint_d(g.scale_y);
    bool fast = fabs(g.scale_x - g.isx) < 2.220446049250313e-16 && fabs(g.scale_y - g.isy) < 2.220446049250313e-16;
    if (g.scale_x >= 1 && g.scale_y >= 1) g.regime = fast ? REG_FAST : REG_GENERAL;
    else g.regime = REG_LINEAR;
}

// OpenCV computeResizeAreaTab for one destination index: up to `cap` (si, alpha) entries.
__device__ int area_entries(int dx, int ssize, double scale, int* si, float* alpha, int cap) {
    double fsx1 = dx * scale;
    double fsx2 = fsx1 + scale;
    double cell = scale < (ssize - fsx1) ? scale : (ssize - fsx1);
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    int k = 0;
    if (sx1 - fsx1 > 1e-3) {
        if (k < cap) { si[k] = sx1 - 1; alpha[k] = (float)((sx1 - fsx1) / cell); }
        k++;
    }
    for (int sx = sx1; sx < sx2; sx++) {
        if (k < cap) { si[k] = sx; alpha[k] = (float)(1.0 / cell); }
        k++;
    }
    if (fsx2 - sx2 > 1e-3) {
        double t = fsx2 - sx2;
        if (t > 1.) t = 1.;
        if (t > cell) t = cell;
        if (k < cap) { si[k] = sx2; alpha[k] = (float)(t / cell); }
        k++;
    }
    return k;
}

// first / one-past-last source index touched by destination index d (area table), without the table
__device__ void area_span(int d, int ssize, double scale, int& lo, int& hi) {
    double fsx1 = d * scale, fsx2 = fsx1 + scale;
    int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
    if (sx2 > ssize - 1) sx2 = ssize - 1;
    if (sx1 > sx2) sx1 = sx2;
    lo = (sx1 - fsx1 > 1e-3) ? sx1 - 1 : sx1;
    hi = (fsx2 - sx2 > 1e-3) ? sx2 + 1 : sx2;
    if (hi <= lo) hi = lo + 1;
}

__device__ void linear_coef(int d, int ssize, double scale, double inv_scale, bool clamp_edges, int& s, int& a0, int& a1) {
    int sx = (int)floor(d * scale);
    float fx = (float)((d + 1) - (sx + 1) * inv_scale);
    fx = fx <= 0 ? 0.f : fx - floorf(fx);
    if (clamp_edges) {
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= ssize - 1) { fx = 0; sx = ssize - 1; }
    }
    s = sx;
    a0 = sat_s16f((1.f - fx) * 2048);
    a1 = sat_s16f(fx * 2048);
}


__device__ __forceinline__ int align16(int v) { return (v + 15) & ~15; }

__device__ __forceinline__ void store_pixel(const PPParams& p, const float* lut, int crop, int f, int dx, int v0, int v1, int v2) {
    // v0..v2 in source channel order; destination channel = swap ? 2-c : c
    const int out = p.out;
    if (p.out_layout == PA_LAYOUT_NHWC4P) {
        // 16-bit only: [crop][f][dx + 4][4] with zeroed 4-pixel borders
        float fv[3];
        fv[0] = lut[(p.swap_rb ? v2 : v0)]; fv[1] = lut[256 + v1]; fv[2] = lut[512 + (p.swap_rb ? v0 : v2)];
        const bool f16 = p.out_f16 != 0;
        const bool split = p.out_split != 0;
        uint16_t hi[3], lo[3];
        for (int c = 0; c < 3; c++) {
            if (f16) {
                const __half h = __float2half_rn(fv[c]);
                hi[c] = __half_as_ushort(h);
                lo[c] = __half_as_ushort(__float2half_rn(__fsub_rn(fv[c], __half2float(h))));
            } else {
                const __nv_bfloat16 h = __float2bfloat16_rn(fv[c]);
                hi[c] = __bfloat16_as_ushort(h);
                lo[c] = __bfloat16_as_ushort(__float2bfloat16_rn(__fsub_rn(fv[c], __bfloat162float(h))));
            }
        }
        uint16_t* o = (uint16_t*)p.outp;
        const int64_t rowbase = ((int64_t)crop * out + f) * (out + 8) * 4;
        const int64_t i = rowbase + (int64_t)(dx + 4) * 4;
        *(uint2*)(o + i) = make_uint2((uint32_t)hi[0] | ((uint32_t)hi[1] << 16), (uint32_t)hi[2]);
        if (split) *(uint2*)(o + p.plane_elems + i) = make_uint2((uint32_t)lo[0] | ((uint32_t)lo[1] << 16), (uint32_t)lo[2]);
        if (dx == 0 || dx == out - 1) {
            const int64_t b = rowbase + (dx == 0 ? 0 : (int64_t)(out + 4) * 4);
            for (int q = 0; q < 4; q++) {
                *(uint2*)(o + b + q * 4) = make_uint2(0, 0);
                if (split) *(uint2*)(o + p.plane_elems + b + q * 4) = make_uint2(0, 0);
            }
        }
        return;
    }
    int vv[3];
    if (p.swap_rb) { vv[0] = v2; vv[1] = v1; vv[2] = v0; } else { vv[0] = v0; vv[1] = v1; vv[2] = v2; }
    if (p.out_dtype == PA_DTYPE_U8) {
        uint8_t* o = (uint8_t*)p.outp;
        if (p.out_layout == PA_LAYOUT_NCHW) {
            for (int c = 0; c < 3; c++) o[(((int64_t)crop * 3 + c) * out + f) * out + dx] = (uint8_t)vv[c];
        } else if (p.out_layout == PA_LAYOUT_NHWC4) {
            uchar4 q = make_uchar4((uint8_t)vv[0], (uint8_t)vv[1], (uint8_t)vv[2], 0);
            *(uchar4*)(o + (((int64_t)crop * out + f) * out + dx) * 4) = q;
        } else {
            uint8_t* q = o + (((int64_t)crop * out + f) * out + dx) * 3;
            q[0] = (uint8_t)vv[0]; q[1] = (uint8_t)vv[1]; q[2] = (uint8_t)vv[2];
        }
        return;
    }
    float fv[3];
    for (int c = 0; c < 3; c++) fv[c] = lut[c * 256 + vv[c]];
    if (p.out_dtype == PA_DTYPE_F32) {
        float* o = (float*)p.outp;
        if (p.out_layout == PA_LAYOUT_NCHW) {
            for (int c = 0; c < 3; c++) o[(((int64_t)crop * 3 + c) * out + f) * out + dx] = fv[c];
        } else if (p.out_layout == PA_LAYOUT_NHWC4) {
            *(float4*)(o + (((int64_t)crop * out + f) * out + dx) * 4) = make_float4(fv[0], fv[1], fv[2], 0.f);
        } else {
            float* q = o + (((int64_t)crop * out + f) * out + dx) * 3;
            q[0] = fv[0]; q[1] = fv[1]; q[2] = fv[2];
        }
        return;
    }
    // 16-bit float (bf16 or IEEE half), optionally with a lo plane holding the rounding residual
    const bool f16 = p.out_f16 != 0;
    uint16_t hi[3], lo[3];
    for (int c = 0; c < 3; c++) {
        if (f16) {
            const __half h = __float2half_rn(fv[c]);
            hi[c] = __half_as_ushort(h);
            lo[c] = __half_as_ushort(__float2half_rn(__fsub_rn(fv[c], __half2float(h))));
        } else {
            const __nv_bfloat16 h = __float2bfloat16_rn(fv[c]);
            hi[c] = __bfloat16_as_ushort(h);
            lo[c] = __bfloat16_as_ushort(__float2bfloat16_rn(__fsub_rn(fv[c], __bfloat162float(h))));
        }
    }
    uint16_t* o = (uint16_t*)p.outp;
    const bool split = p.out_split != 0;
    if (p.out_layout == PA_LAYOUT_NCHW) {
        for (int c = 0; c < 3; c++) {
            int64_t i = (((int64_t)crop * 3 + c) * out + f) * out + dx;
            o[i] = hi[c];
            if (split) o[p.plane_elems + i] = lo[c];
        }
    } else if (p.out_layout == PA_LAYOUT_NHWC4) {
        int64_t i = (((int64_t)crop * out + f) * out + dx) * 4;
        uint2 q;
        q.x = (uint32_t)hi[0] | ((uint32_t)hi[1] << 16);
        q.y = (uint32_t)hi[2];
        *(uint2*)(o + i) = q;
        if (split) {
            q.x = (uint32_t)lo[0] | ((uint32_t)lo[1] << 16);
            q.y = (uint32_t)lo[2];
            *(uint2*)(o + p.plane_elems + i) = q;
        }
    } else {
        int64_t i = (((int64_t)crop * out + f) * out + dx) * 3;
        for (int c = 0; c < 3; c++) {
            o[i + c] = hi[c];
            if (split) o[p.plane_elems + i + c] = lo[c];
        }
    }
}

// zero-fill the slab of an invalid crop (true zeros, not normalised black)
__device__ void zero_rows(const PPParams& p, int crop, int F0, int F1) {
    const int out = p.out;
    int esz = (p.out_dtype == PA_DTYPE_U8) ? 1 : (p.out_dtype == PA_DTYPE_F32 ? 4 : 2);
    int ch = (p.out_layout == PA_LAYOUT_NHWC4 || p.out_layout == PA_LAYOUT_NHWC4P) ? 4 : 3;
    const int wpad = (p.out_layout == PA_LAYOUT_NHWC4P) ? 8 : 0;
    int nplanes = p.out_split ? 2 : 1;
    for (int pl = 0; pl < nplanes; pl++) {
        uint8_t* base = (uint8_t*)p.outp + (int64_t)pl * p.plane_elems * esz;
        if (p.out_layout == PA_LAYOUT_NCHW) {
            for (int c = 0; c < 3; c++) {
                uint8_t* q = base + ((((int64_t)crop * 3 + c) * out + F0) * out) * esz;
                int64_t n = (int64_t)(F1 - F0) * out * esz;
                for (int64_t i = threadIdx.x; i < n; i += blockDim.x) q[i] = 0;
            }
        } else {
            uint8_t* q = base + (((int64_t)crop * out + F0) * (out + wpad)) * ch * esz;
            int64_t n = (int64_t)(F1 - F0) * (out + wpad) * ch * esz;
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) q[i] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming kernel: one CTA per (crop, slab of final rows). Raw rows are pulled through three ring
// stages entirely in shared memory -- RAW batch -> T (after the horizontal bicubic pass) ->
// S (canvas rows after the vertical pass, with the black letterbox) -> area pass -> output --
// so every source row is fetched and filtered once per slab, whatever the crop size.
struct PartPlan {
    int a0, a1;            // area-output rows of this slab
    int s_begin, s_end;    // canvas rows they touch
    int v_begin, v_end;    // resized-image rows among those
    int t_begin, t_end;    // raw rows those need
    int RB, CT, CS;        // raw rows per batch, ring capacities (rows)
    int CX;                // capacity of the fp32 x-pass ring (general area regime), else 0
    int h_global;          // very wide windows: leave the horizontal coefficient table in global memory
    int ok;                // 0: does not fit in shared memory
};

__device__ __forceinline__ uint32_t clip8w(int32_t v) {
    v >>= 22;
    return (uint32_t)min(max(v, 0), 255);
}

// canvas-row span [lo, hi) that area-output row `a` reads
__device__ __forceinline__ void area_rows(const CropGeom& g, int a, int& lo, int& hi) {
    if (g.regime == REG_COPY) { lo = a; hi = a + 1; }
    else if (g.regime == REG_FAST) { lo = a * g.isy; hi = lo + g.isy; }
    else if (g.regime == REG_GENERAL) { area_span(a, g.sd, g.scale_y, lo, hi); }
    else {
        int s, b0, b1;
        linear_coef(a, g.sd, g.scale_y, g.inv_scale_y, false, s, b0, b1);
        lo = min(max(s, 0), g.sd - 1);
        hi = min(max(s + 1, 0), g.sd - 1) + 1;
    }
    lo = max(lo, 0); hi = min(hi, g.sd);
}

// Table sizes shared by the plan kernel (writer) and the main kernel (reader).
__device__ __forceinline__ int tab_ksh(const CropGeom& g) { return g.hact ? (g.h_ks <= 8 ? 8 : g.h_ks) : 0; }
__device__ __forceinline__ int tab_xcap(const CropGeom& g) { return (g.regime == REG_GENERAL) ? (int)ceil(g.scale_x) + 2 : 0; }
__device__ __forceinline__ int tab_ycap(const CropGeom& g) { return (g.regime == REG_GENERAL) ? (int)ceil(g.scale_y) + 2 : 2; }

// Rows of the chain one slab of final rows [F0, F1) needs: area rows [a0, a1), canvas rows [s_begin, s_end),
// resized rows [v_begin, v_end), raw rows [t_begin, t_end) (bounds from the crop's vertical table).
struct SlabRows { int a0, a1, s_begin, s_end, v_begin, v_end, t_begin, t_end; };
template <typename TY, typename TN>
__device__ void slab_rows(const CropGeom& g, const TY* v_ymin, const TN* v_n, int F0, int F1, SlabRows& q) {
    q.a0 = max(F0 - g.oy2, 0); q.a1 = min(F1 - g.oy2, g.oh);
    if (q.a1 < q.a0) q.a1 = q.a0;
    q.s_begin = q.s_end = q.v_begin = q.v_end = q.t_begin = q.t_end = 0;
    if (q.a1 > q.a0) {
        int lo, hi;
        area_rows(g, q.a0, lo, hi); q.s_begin = lo;
        area_rows(g, q.a1 - 1, lo, hi); q.s_end = hi;
        q.v_begin = min(max(q.s_begin - g.oy, 0), g.nh); q.v_end = min(max(q.s_end - g.oy, 0), g.nh);
        if (q.v_end > q.v_begin) {
            q.t_begin = (int)v_ymin[q.v_begin];
            q.t_end = (int)v_ymin[q.v_end - 1] + (int)v_n[q.v_end - 1];
        }
    }
}

// What one work item of the tensor-core kernel needs to know about its strip of final columns (8 ints, written by the
// plan kernel into the crop's table block at off_tc + 8 * strip; the part records, SlabRows, follow at off_tc + 8 * 128)
struct TcStrip { int dx0, dx1, cx_lo, xx_lo, ncols, kb0, nks_h, pad; };

// Shared-memory copies of the small index tables the routing decision walks (filled while the tables are generated)
struct PlanScratch {
    short h_xmin[PA_MAX_WINDOW]; uint8_t h_n[PA_MAX_WINDOW];
    short v_ymin[PA_MAX_WINDOW]; uint8_t v_n[PA_MAX_WINDOW];
    short x_first[128], x_last[128];      // first / last canvas column of every final column (area table)
};

// Shared scratch of the routing decision
struct RouteScratch {
    int ok, np, rv, dcap, ns, base;
    SlabRows parts[4];
    short end[128];       // strip that starts at final column i ends before end[i] (== i: no strip fits there)
    short rank[128];      // strip index of the strips the greedy chain picked, -1 elsewhere
};

// Can the tensor-core kernel take this crop? (both bicubic passes active, general INTER_AREA regime, every tile's taps
// inside one 128-byte K atom.) If so: cut the final columns into strips and the final rows into parts, write their
// records, reserve the vertical coefficient tiles and enqueue one work item per (strip, part). Called by the whole CTA
// (128 threads) once the crop's tables exist: the per-start-column strip search runs one column per thread, thread 0
// only walks the chain of strips.
__device__ void tc_route(CropGeom& g, const PPParams& p, int crop, const PlanScratch& ps, RouteScratch& rs) {
    const int out = p.out, tid = threadIdx.x;
    int* tab = p.tables + (int64_t)crop * p.table_stride;
    const int nw = g.nw;
    if (tid == 0) {
        rs.ok = 0; rs.ns = 0;
        bool ok = g.pad1 && g.hact && g.vact && g.regime == REG_GENERAL && out <= 128 && g.h_ks <= 8 && g.v_ks <= 16 &&
                  g.rw <= PA_MAX_WINDOW && g.rh <= PA_MAX_WINDOW && tab_xcap(g) <= TC_XT_CAP;
        // ---- parts: fewest row slabs whose raw-row span fits the T ring
        int np = 0, smax = 0;
        for (int cand = 1; cand <= 4 && !np && ok; cand++) {
            bool fit = true; int sm = 0;
            for (int part = 0; part < cand && fit; part++) {
                slab_rows(g, ps.v_ymin, ps.v_n, (int)((int64_t)part * out / cand), (int)((int64_t)(part + 1) * out / cand), rs.parts[part]);
                if (rs.parts[part].t_end - rs.parts[part].t_begin > TC_TT_ROWS) fit = false;
                sm = max(sm, rs.parts[part].s_end - rs.parts[part].s_begin);
            }
            if (fit) { np = cand; smax = sm; }
        }
        ok = ok && np > 0;
        // ---- vertical blocks: rv resized rows whose taps fit one K atom (window start aligned down to 32 raw rows)
        int rv = 0;
        for (int cand = 64; cand >= 16 && !rv && ok; cand -= 16) {
            bool fit = true;
            for (int part = 0; part < np && fit; part++) {
                const SlabRows& q = rs.parts[part];
                for (int v0 = q.v_begin; v0 < q.v_end && fit; v0 += cand) {
                    const int vl = min(v0 + cand, q.v_end) - 1;
                    const int kw0 = (ps.v_ymin[v0] - q.t_begin) & ~31;
                    if (ps.v_ymin[vl] + ps.v_n[vl] - q.t_begin - kw0 > TC_KSPAN) fit = false;
                }
            }
            if (fit) rv = cand;
        }
        ok = ok && rv > 0;
        const int dcap = min(smax > 0 ? TC_XB_FLOATS / (3 * smax) : out, TC_MAX_D);
        ok = ok && dcap >= 1;
        rs.np = np; rs.rv = rv; rs.dcap = dcap; rs.ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!rs.ok) return;
    // ---- strips of final columns: <= TC_NCOLS canvas columns, horizontal taps inside one K atom, x-pass buffer bound.
    // Thread i finds the longest strip that starts at final column i.
    TcStrip st;
    st.dx0 = tid; st.dx1 = tid; st.cx_lo = 0; st.xx_lo = 0; st.ncols = 0; st.kb0 = 0; st.nks_h = 0; st.pad = 0;
    if (tid < out) {
        const int dcap = rs.dcap, dx0 = tid;
        const int cx_lo = ps.x_first[dx0];
        st.cx_lo = cx_lo;
        int dx1 = dx0;
        while (dx1 < out && dx1 - dx0 < dcap) {
            const int cx_hi = ps.x_last[dx1] + 1;
            if (cx_hi - cx_lo > TC_NCOLS) break;
            const int xx_lo = min(max(cx_lo - g.ox, 0), nw), xx_hi = min(max(cx_hi - g.ox, 0), nw);
            int kb0 = 0, kend = 0;
            if (xx_hi > xx_lo) {
                kb0 = ((g.x0 + ps.h_xmin[xx_lo]) * 3) & ~15;
                kend = (g.x0 + ps.h_xmin[xx_hi - 1] + ps.h_n[xx_hi - 1]) * 3 - kb0;
                if (kend > TC_KSPAN) break;
            }
            st.xx_lo = xx_lo; st.ncols = xx_hi - xx_lo; st.kb0 = kb0; st.nks_h = (kend + 31) >> 5;
            dx1++;
        }
        st.dx1 = dx1;
        rs.end[tid] = (short)dx1;
    }
    if (tid < 128) rs.rank[tid] = -1;
    __syncthreads();
    if (tid == 0) {
        int ns = 0, dx0 = 0;
        bool ok = true;
        while (dx0 < out && ok) {
            if (ns >= TC_MAX_STRIPS || rs.end[dx0] == dx0) { ok = false; break; }   // a single column does not fit: leave the crop to the streaming kernel
            rs.rank[dx0] = (short)ns++;
            dx0 = rs.end[dx0];
        }
        // ---- vertical coefficient tiles: one per (part, block of rv resized rows), reserved from the stream's tile pool
        // (a bump counter the tensor-core kernel's last CTA re-zeroes); an exhausted pool leaves the crop to the streaming kernel
        if (ok) {
            const int np = rs.np, rv = rs.rv;
            int nt = 0;
            for (int i = 0; i < np; i++) { g.tile_first[i] = nt; nt += (rs.parts[i].v_end - rs.parts[i].v_begin + rv - 1) / rv; }
            const int tb = atomicAdd(p.tc_counters + 3, nt);
            if (tb + nt > p.tc_pool_tiles) ok = false;
            else {
                for (int i = 0; i < np; i++) g.tile_first[i] += tb;
                g.route = 1; g.n_strips = ns; g.n_parts = np; g.rv = rv;
                rs.base = atomicAdd(p.tc_counters, ns * np);
                rs.ns = ns;
            }
        }
        rs.ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!rs.ok) return;
    if (tid < out && rs.rank[tid] >= 0) ((TcStrip*)(tab + g.off_tc))[rs.rank[tid]] = st;
    if (tid < rs.np) ((SlabRows*)(tab + g.off_tc + 8 * TC_MAX_STRIPS))[tid] = rs.parts[tid];
    const int ns = rs.ns;
    for (int i = tid; i < ns * rs.np; i += 128) p.tc_items[rs.base + i] = make_int2(crop, (i % ns) | ((i / ns) << 16));
}

// One CTA per crop: geometry + every coefficient table of the crop, once, into global memory
// (L2-resident), so that the four slab CTAs of the main kernel only copy what they need.
__global__ void __launch_bounds__(128) preprocess_plan_kernel(const PPParams p) {
    __shared__ CropGeom g;
    __shared__ PlanScratch ps;
    __shared__ RouteScratch rs;
    __shared__ __align__(16) uint8_t tile_s[TC_TILE_BYTES];      // one vertical coefficient tile under construction
    const int crop = blockIdx.x, tid = threadIdx.x;
    const int out = p.out;
    if (tid == 0) pdl_launch_dependents();
    if (tid == 0) {
        if (p.status) p.status[crop] = 0x7f7f7f7f;     // "not computed yet": every kernel that finishes a crop atomicMin's its outcome in
        if (crop == 0) *p.deferred = 0;
        compute_geom(g, p.boxes + (int64_t)crop * PA_BOX_STRIDE, p.H, p.W, p.n_frames, out, p.padding);
        g.tab_ok = 0; g.off_h = g.off_v = g.off_x = g.off_y = 0;
        if (g.status == PA_CROP_OK) {
            int o = 0;
            g.off_h = o; if (g.hact) o += 2 * g.nw + 4 + g.nw * tab_ksh(g);
            o = (o + 3) & ~3;
            g.off_v = o; if (g.vact) o += 2 * g.nh + g.nh * g.v_ks;
            o = (o + 3) & ~3;
            g.off_x = o;
            if (g.regime == REG_GENERAL) o += (out + 1) + 2 * out * tab_xcap(g);
            else if (g.regime == REG_LINEAR) o += 2 * out;
            o = (o + 3) & ~3;
            g.off_y = o;
            if (g.regime == REG_GENERAL) o += g.oh + 2 * g.oh * tab_ycap(g);
            else if (g.regime == REG_LINEAR) o += 4 * g.oh;
            o = (o + 3) & ~3;
            g.off_tc = o; o += 8 * TC_MAX_STRIPS + 8 * 4;
            g.tab_ok = (o <= p.table_stride) ? 1 : 0;
        }
    }
    __syncthreads();
    if (g.status == PA_CROP_OK && g.tab_ok) {
        int* tab = p.tables + (int64_t)crop * p.table_stride;
        const int nw = g.nw, nh = g.nh, sd = g.sd;
        if (g.hact) {
            const int KSH = tab_ksh(g);
            int* h_xmin = tab + g.off_h; int* h_n = h_xmin + nw; int* h_kk = tab + g.off_h + ((2 * nw + 3) & ~3);
            for (int xx = tid; xx < nw; xx += 128) {
                int xm, n;
                bicubic_coeffs(xx, g.rw, g.h_scale, g.h_fs, g.h_sup, g.h_ks, xm, n, h_kk + (size_t)xx * KSH);
                for (int j = g.h_ks; j < KSH; j++) h_kk[(size_t)xx * KSH + j] = 0;
                h_xmin[xx] = xm; h_n[xx] = n;
                if (xx < PA_MAX_WINDOW) { ps.h_xmin[xx] = (short)xm; ps.h_n[xx] = (uint8_t)n; }
            }
        }
        if (g.vact) {
            int* v_ymin = tab + g.off_v; int* v_n = v_ymin + nh; int* v_kk = v_n + nh;
            for (int i = tid; i < nh; i += 128) {
                int ym, n;
                bicubic_coeffs(i, g.rh, g.v_scale, g.v_fs, g.v_sup, g.v_ks, ym, n, v_kk + (size_t)i * g.v_ks);
                v_ymin[i] = ym; v_n[i] = n;
                if (i < PA_MAX_WINDOW) { ps.v_ymin[i] = (short)ym; ps.v_n[i] = (uint8_t)n; }
            }
        }
        if (g.regime == REG_GENERAL) {
            const int xcap = tab_xcap(g), ycap = tab_ycap(g);
            int* xt_n = tab + g.off_x; int* xt_si = xt_n + (out + 1); float* xt_al = (float*)(xt_si + out * xcap);
            for (int dx = tid; dx < out; dx += 128) {
                int n = area_entries(dx, sd, g.scale_x, xt_si + dx * xcap, xt_al + dx * xcap, xcap);
                xt_n[dx] = n < xcap ? n : xcap;
                if (dx < 128 && n >= 1) { ps.x_first[dx] = (short)xt_si[dx * xcap]; ps.x_last[dx] = (short)xt_si[dx * xcap + (n < xcap ? n : xcap) - 1]; }
            }
            int* yt_n = tab + g.off_y; int* yt_s = yt_n + g.oh; float* yt_b = (float*)(yt_s + g.oh * ycap);
            for (int i = tid; i < g.oh; i += 128) {
                int n = area_entries(i, sd, g.scale_y, yt_s + i * ycap, yt_b + i * ycap, ycap);
                yt_n[i] = n < ycap ? n : ycap;
            }
        } else if (g.regime == REG_LINEAR) {
            int* lx_s = tab + g.off_x; int* lx_a = lx_s + out;
            for (int dx = tid; dx < out; dx += 128) {
                int s_, a0, a1;
                linear_coef(dx, sd, g.scale_x, g.inv_scale_x, true, s_, a0, a1);
                lx_s[dx] = s_; lx_a[dx] = (a0 & 0xFFFF) | (a1 << 16);
            }
            int* yt_s = tab + g.off_y; int* yt_b = yt_s + 2 * g.oh;
            for (int i = tid; i < g.oh; i += 128) {
                int s_, b0, b1;
                linear_coef(i, sd, g.scale_y, g.inv_scale_y, false, s_, b0, b1);
                yt_s[i * 2] = s_; yt_s[i * 2 + 1] = 0;
                yt_b[i * 2] = b0; yt_b[i * 2 + 1] = b1;
            }
        }
    }
    __syncthreads();   // the tables written above are read back below (same block: visible after the barrier)
    if (tid == 0) { g.route = 0; g.n_strips = 0; g.n_parts = 0; g.rv = 0; }
    if (p.tc_enable && g.status == PA_CROP_OK && g.tab_ok) tc_route(g, p, crop, ps, rs);      // CTA-uniform condition
    __syncthreads();
    // Vertical coefficient tiles of a crop the tensor-core kernel takes: the three digit planes of every block of rv
    // resized rows, laid out exactly as the kernel's shared-memory operand (K-major SWIZZLE_128B, K = raw row inside the
    // block's 32-aligned window). Built ONCE per crop here; each of the crop's ~20 strips then fetches a tile with one
    // bulk copy instead of scattering the same coefficients again.
    if (g.route == 1) {
        const int* tab = p.tables + (int64_t)crop * p.table_stride;
        const int* v_kk = tab + g.off_v + 2 * g.nh;
        const int rv = g.rv, vks = g.v_ks;
        for (int part = 0; part < g.n_parts; part++) {
            const SlabRows& q = rs.parts[part];
            const int nrows = q.v_end - q.v_begin, nblk = (nrows + rv - 1) / rv;
            for (int blk = 0; blk < nblk; blk++) {
                // the tile is assembled in shared memory (zero fill, then one byte per tap and digit) and leaves as whole
                // 128-byte rows: full-line writes, no partial sectors
                const int v0 = q.v_begin + blk * rv, nr = min(rv, q.v_end - v0);
                const int kw0 = (ps.v_ymin[v0] - q.t_begin) & ~31;
                // a thread's taps of this tile: slot = (row, tap) with 8 or 16 tap slots per row; the coefficients are fetched
                // BEFORE the zero fill so that their (L2) latency hides under it
                const int jsh = vks <= 8 ? 3 : 4;
                int kreg[8], koff[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int slot = tid + 128 * u, r = slot >> jsh, j = slot & ((1 << jsh) - 1);
                    koff[u] = -1; kreg[u] = 0;
                    if (r < nr && j < ps.v_n[v0 + r]) {
                        kreg[u] = v_kk[(size_t)(v0 + r) * vks + j];
                        koff[u] = sw128(r, ps.v_ymin[v0 + r] + j - q.t_begin - kw0);
                    }
                }
                for (int i = tid; i < TC_TILE_BYTES / 16; i += 128) ((uint4*)tile_s)[i] = make_uint4(0, 0, 0, 0);
                __syncthreads();
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (koff[u] >= 0) {
                        int d0, d1, d2;
                        coef_digits(kreg[u], d0, d1, d2);
                        uint8_t* t = tile_s + koff[u];
                        t[0] = (uint8_t)d0; t[64 * 128] = (uint8_t)d1; t[128 * 128] = (uint8_t)d2;
                    }
                }
                __syncthreads();
                // rows beyond nr feed accumulator lanes nobody reads: not written
                uint4* dst = (uint4*)(p.tc_tiles + (size_t)(g.tile_first[part] + blk) * TC_TILE_BYTES);
#pragma unroll
                for (int d = 0; d < 3; d++)
                    for (int o = tid; o < nr * 8; o += 128) dst[d * 512 + o] = ((const uint4*)tile_s)[d * 512 + o];
                if (tid == 0) {
                    const int vl = v0 + nr - 1;
                    const int kend = ps.v_ymin[vl] + ps.v_n[vl] - q.t_begin - kw0;      // <= TC_KSPAN: tc_route checked every block
                    p.tc_tile_rec[g.tile_first[part] + blk] = kw0 | (((kend + 31) >> 5) << 16);
                }
                __syncthreads();
            }
        }
    }
    // publish the geometry (plain words; the main kernel launches after this one on the same stream)
    const int* src = (const int*)&g;
    int* dst = (int*)(p.geoms + (size_t)crop * sizeof(CropGeom));
    for (int i = tid; i < (int)(sizeof(CropGeom) / 4); i += 128) dst[i] = src[i];
    // Launched with programmatic stream serialization and nothing above depends on the predecessor, so all of it ran
    // under the predecessor's tail. The wait at the very END keeps stream order transitive: this kernel does not complete
    // before its predecessor has, so whatever is launched behind it is still ordered after everything before it.
    if (tid == 0) pdl_wait();
}

__device__ void preprocess_slab(const PPParams& p, const int slab) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ CropGeom g;
    __shared__ PartPlan pl;

    const int crop = slab / PP_SPLIT, part = slab % PP_SPLIT;
    const int tid = threadIdx.x;
    const int NT = blockDim.x;   // 256 or 384 threads (runtime: see pa_preprocess)
    const int out = p.out;
    if (p.geoms) {
        const int* src = (const int*)(p.geoms + (size_t)crop * sizeof(CropGeom));
        int* dst = (int*)&g;
        for (int i = tid; i < (int)(sizeof(CropGeom) / 4); i += NT) dst[i] = src[i];
    } else if (tid == 0) {
        compute_geom(g, p.boxes + (int64_t)crop * PA_BOX_STRIDE, p.H, p.W, p.n_frames, out, p.padding);
        g.tab_ok = 0;
    }
    __syncthreads();
    if (p.geoms && g.route == 1) return;   // computed by the tensor-core kernel
    const int F0 = (int)((int64_t)part * out / PP_SPLIT), F1 = (int)((int64_t)(part + 1) * out / PP_SPLIT);
    if (g.status != PA_CROP_OK) {
        if (p.first_pass_smem > 0) return;  // reported by the first pass
        if (tid == 0 && p.status) atomicMin(p.status + crop, g.status);
        zero_rows(p, crop, F0, F1);
        return;
    }
    const int nw = g.nw, nh = g.nh, sd = g.sd, rw = g.rw, rh = g.rh;
    const int nw3 = nw * 3, sd3 = sd * 3;
    const int rawp = align16(rw * 3 + 15) + 32;          // 16-byte aligned copy incl. alignment shift + over-read pad
    const int tp = g.hact ? align16(nw3) + 16 : rawp;    // T row pitch (raw rows themselves when there is no H pass)
    const int sp = align16(sd3) + 16;
    const int KSH = tab_ksh(g), xcap = tab_xcap(g), ycap = tab_ycap(g);

    // ---- slab plan + shared-memory budget (thread 0)
    if (tid == 0) {
        PartPlan q;
        q.ok = 1;
        q.a0 = max(F0 - g.oy2, 0); q.a1 = min(F1 - g.oy2, g.oh);
        if (q.a1 < q.a0) q.a1 = q.a0;
        q.s_begin = q.s_end = q.v_begin = q.v_end = q.t_begin = q.t_end = 0;
        q.RB = q.CT = q.CS = q.CX = 0;
        q.h_global = 0;
        if (q.a1 > q.a0) {
            int lo, hi;
            area_rows(g, q.a0, lo, hi); q.s_begin = lo;
            area_rows(g, q.a1 - 1, lo, hi); q.s_end = hi;
            q.v_begin = min(max(q.s_begin - g.oy, 0), nh); q.v_end = min(max(q.s_end - g.oy, 0), nh);
            if (q.v_end > q.v_begin) {
                if (g.vact) {
                    double c0 = 0.0 + (q.v_begin + 0.5) * g.v_scale;
                    int ymin = (int)(c0 - g.v_sup + 0.5); if (ymin < 0) ymin = 0;
                    double c1 = 0.0 + ((q.v_end - 1) + 0.5) * g.v_scale;
                    int ymax = (int)(c1 + g.v_sup + 0.5); if (ymax > rh) ymax = rh;
                    q.t_begin = ymin; q.t_end = ymax;
                } else { q.t_begin = q.v_begin; q.t_end = q.v_end; }
            }
            // fixed tables
            const int na = q.a1 - q.a0, nv = q.v_end - q.v_begin;
            int fixed = 768 * 4;
            const int h_bytes = g.hact ? nw * 8 + nw * KSH * 4 : 0;
            fixed += h_bytes;
            if (g.regime == REG_GENERAL) fixed += (out + 1) * 4 + out * xcap * 8;
            else if (g.regime == REG_LINEAR) fixed += out * 4 + align16(out * 4);
            fixed += (na + 1) * 4 + na * ycap * 8 + na * 8;
            if (g.vact) fixed += nv * 8 + nv * g.v_ks * 4;
            fixed = align16(fixed) + 64;
            const int vwin = g.vact ? g.v_ks : 1;
            const int awin = (g.regime == REG_COPY) ? 1 : (g.regime == REG_FAST ? g.isy : (g.regime == REG_GENERAL ? (int)floor(g.scale_y) + 2 : 2));
            auto fit = [&](int limit_, int& RBo, int& CTo, int& CSo, int& CXo, bool hg = false) {
                const int limit = limit_ + (hg ? h_bytes : 0);  // table left in global memory: its bytes are free
                auto pow2 = [](int v) { int q = 1; while (q < v) q <<= 1; return q; };
                const bool gen = (g.regime == REG_GENERAL) && p.use_xb;
                // ring capacities are powers of two (slot = row & (cap - 1)); very wide windows that
                // only fit with exact capacities take the modulo path
                for (int exact = 0; exact < 2; exact++) {
                    for (int RB = exact ? 4 : 32; RB >= 1; RB >>= 1) {
                        int CT = RB + max(vwin, g.pad1 ? 0 : awin) + 1;
                        const int pr = g.vact ? (int)ceil(RB / g.v_scale) + 1 : RB;   // canvas rows one batch can complete
                        // general regime: S only holds the rows of the current batch (their x-pass goes to the
                        // fp32 ring XB, which is what the y-pass keeps); other regimes read S directly
                        int CS = gen ? pr : pr + awin;
                        int CX = gen ? (g.pad1 ? pr + awin : CT) : 0;
                        if (!exact) { CT = pow2(CT); CS = pow2(CS); if (gen) CX = g.pad1 ? pow2(CX) : CT; }
                        else if (gen && !g.pad1) CX = CT;
                        int need = fixed + CT * tp + 64;
                        if (g.hact) need += RB * rawp + 64;
                        if (g.pad1) need += CS * sp;
                        need += CX * out * 12 + 16;
                        if (need <= limit) { RBo = RB; CTo = CT; CSo = CS; CXo = CX; return true; }
                    }
                }
                return false;
            };
            int rb, ct, cs, cx;
            if (p.first_pass_smem > 0 && fit(p.first_pass_smem, rb, ct, cs, cx)) q.ok = -1;   // done by the first pass
            else {
                q.ok = fit(p.smem_bytes, q.RB, q.CT, q.CS, q.CX) ? 1 : 0;
                if (!q.ok && g.tab_ok && p.tables && g.hact && fit(p.smem_bytes, q.RB, q.CT, q.CS, q.CX, true)) { q.ok = 1; q.h_global = 1; }
            }
        } else if (p.first_pass_smem > 0) {
            q.ok = -1;
        }
        pl = q;
    }
    __syncthreads();
    if (pl.ok < 0) return;
    if (!pl.ok) {
        if (p.defer_too_large) {  // the second pass retries this slab with the full carve-out
            if (tid == 0) atomicAdd(p.deferred, 1);
            return;
        }
        if (tid == 0 && p.status) atomicMin(p.status + crop, PA_CROP_TOO_LARGE);
        zero_rows(p, crop, F0, F1);
        return;
    }
    if (tid == 0 && p.status) atomicMin(p.status + crop, PA_CROP_OK);  // status = worst outcome over the slabs
    const PartPlan P = pl;
    const int na = P.a1 - P.a0, nv = P.v_end - P.v_begin;

    // ---- carve shared memory
    int off = 0;
    float* lut = (float*)(smem + off); off += 768 * 4;
    int* h_xmin = nullptr; int* h_n = nullptr; int32_t* h_kk = nullptr;
    if (g.hact && P.h_global) {
        int* tabg = p.tables + (int64_t)crop * p.table_stride + g.off_h;
        h_xmin = tabg; h_n = tabg + nw; h_kk = tabg + ((2 * nw + 3) & ~3);
    } else if (g.hact) {
        h_xmin = (int*)(smem + off); off += nw * 4;
        h_n = (int*)(smem + off); off += nw * 4;
        off = align16(off);
        h_kk = (int32_t*)(smem + off); off += nw * KSH * 4;
    }
    int* xt_n = nullptr; int* xt_si = nullptr; float* xt_al = nullptr; int* lx_s = nullptr; short* lx_a = nullptr;
    if (g.regime == REG_GENERAL) {
        xt_n = (int*)(smem + off); off += (out + 1) * 4;
        xt_si = (int*)(smem + off); off += out * xcap * 4;
        xt_al = (float*)(smem + off); off += out * xcap * 4;
    } else if (g.regime == REG_LINEAR) {
        lx_s = (int*)(smem + off); off += out * 4;
        lx_a = (short*)(smem + off); off += align16(out * 4);
    }
    int* ar_lo = (int*)(smem + off); off += na * 4;   // canvas-row span [lo, hi) of every area row of the slab
    int* ar_hi = (int*)(smem + off); off += na * 4;
    int* yt_n = (int*)(smem + off); off += (na + 1) * 4;
    int* yt_s = (int*)(smem + off); off += na * ycap * 4;
    float* yt_b = (float*)(smem + off); off += na * ycap * 4;
    int* v_ymin = nullptr; int* v_n = nullptr; int32_t* v_kk = nullptr;
    if (g.vact) {
        v_ymin = (int*)(smem + off); off += nv * 4;
        v_n = (int*)(smem + off); off += nv * 4;
        v_kk = (int32_t*)(smem + off); off += nv * g.v_ks * 4;
    }
    off = align16(off);
    uint8_t* RAW = nullptr;
    if (g.hact) { RAW = smem + off; off += P.RB * rawp + 64; }
    uint8_t* T = smem + off; off += P.CT * tp + 64;
    uint8_t* S = nullptr;
    if (g.pad1) { S = smem + off; off += P.CS * sp; }
    // general area regime: horizontal area sums of every canvas row, computed once per row (fp32 [out*3])
    float* XB = nullptr;
    const int xbp = out * 3;
    if (P.CX > 0) { off = align16(off); XB = (float*)(smem + off); off += P.CX * xbp * 4; }
    if (off > p.smem_bytes) { __trap(); }  // budget computed above must hold: fail loudly

    // ---- tables
    for (int i = tid; i < 768; i += NT) {
        int c = i >> 8, v = i & 255;
        float f = __fdiv_rn((float)v, 255.0f);
        lut[i] = p.out_raw ? (float)v : __fdiv_rn(__fsub_rn(f, p.mean[c]), p.stdv[c]);
    }
    if (g.tab_ok && p.tables) {
        // copy the slices this slab needs from the crop's precomputed block (preprocess_plan_kernel)
        const int* tab = p.tables + (int64_t)crop * p.table_stride;
        if (g.hact && !P.h_global) {
            const int* gx = tab + g.off_h; const int* gk = tab + g.off_h + ((2 * nw + 3) & ~3);
            for (int i = tid; i < nw; i += NT) { h_xmin[i] = gx[i]; h_n[i] = gx[nw + i]; }
            const int4* gk4 = (const int4*)gk; int4* hk4 = (int4*)h_kk;
            if ((KSH & 3) == 0) { for (int i = tid; i < nw * KSH / 4; i += NT) hk4[i] = __ldg(gk4 + i); }
            else { for (int i = tid; i < nw * KSH; i += NT) h_kk[i] = gk[i]; }
        }
        if (g.vact) {
            const int* gv = tab + g.off_v;
            for (int i = tid; i < nv; i += NT) { v_ymin[i] = gv[P.v_begin + i]; v_n[i] = gv[nh + P.v_begin + i]; }
            const int* gk = gv + 2 * nh + (size_t)P.v_begin * g.v_ks;
            for (int i = tid; i < nv * g.v_ks; i += NT) v_kk[i] = gk[i];
        }
        if (g.regime == REG_GENERAL) {
            const int* gx = tab + g.off_x;
            for (int i = tid; i < out; i += NT) xt_n[i] = gx[i];
            const int* gsi = gx + (out + 1); const int* gal = gsi + out * xcap;
            for (int i = tid; i < out * xcap; i += NT) { xt_si[i] = gsi[i]; ((int*)xt_al)[i] = gal[i]; }
            const int* gy = tab + g.off_y;
            for (int i = tid; i < na; i += NT) yt_n[i] = gy[P.a0 + i];
            const int* gys = gy + g.oh + (size_t)P.a0 * ycap; const int* gyb = gy + g.oh + (size_t)g.oh * ycap + (size_t)P.a0 * ycap;
            for (int i = tid; i < na * ycap; i += NT) { yt_s[i] = gys[i]; ((int*)yt_b)[i] = gyb[i]; }
        } else if (g.regime == REG_LINEAR) {
            const int* gx = tab + g.off_x;
            for (int i = tid; i < out; i += NT) {
                lx_s[i] = gx[i];
                const int a = gx[out + i];
                lx_a[i * 2] = (short)(a & 0xFFFF); lx_a[i * 2 + 1] = (short)(a >> 16);
            }
            const int* gy = tab + g.off_y;
            for (int i = tid; i < na * 2; i += NT) { yt_s[i] = gy[P.a0 * 2 + i]; ((int*)yt_b)[i] = gy[2 * g.oh + P.a0 * 2 + i]; }
        }
    } else {
    if (g.hact) {
        for (int xx = tid; xx < nw; xx += NT) {
            int xm, n;
            bicubic_coeffs(xx, rw, g.h_scale, g.h_fs, g.h_sup, g.h_ks, xm, n, h_kk + (size_t)xx * KSH);
            for (int j = g.h_ks; j < KSH; j++) h_kk[(size_t)xx * KSH + j] = 0;
            h_xmin[xx] = xm; h_n[xx] = n;
        }
    }
    if (g.regime == REG_GENERAL) {
        for (int dx = tid; dx < out; dx += NT) {
            int n = area_entries(dx, sd, g.scale_x, xt_si + dx * xcap, xt_al + dx * xcap, xcap);
            xt_n[dx] = n < xcap ? n : xcap;
        }
        for (int i = tid; i < na; i += NT) {
            int n = area_entries(P.a0 + i, sd, g.scale_y, yt_s + i * ycap, yt_b + i * ycap, ycap);
            yt_n[i] = n < ycap ? n : ycap;
        }
    } else if (g.regime == REG_LINEAR) {
        for (int dx = tid; dx < out; dx += NT) {
            int s, a0, a1;
            linear_coef(dx, sd, g.scale_x, g.inv_scale_x, true, s, a0, a1);
            lx_s[dx] = s; lx_a[dx * 2] = (short)a0; lx_a[dx * 2 + 1] = (short)a1;
        }
        for (int i = tid; i < na; i += NT) {
            int s, b0, b1;
            linear_coef(P.a0 + i, sd, g.scale_y, g.inv_scale_y, false, s, b0, b1);
            yt_s[i * 2] = s;
            ((int*)yt_b)[i * 2] = b0; ((int*)yt_b)[i * 2 + 1] = b1;
        }
    }
    if (g.vact) {
        for (int i = tid; i < nv; i += NT) {
            int ym, n;
            bicubic_coeffs(P.v_begin + i, rh, g.v_scale, g.v_fs, g.v_sup, g.v_ks, ym, n, v_kk + (size_t)i * g.v_ks);
            v_ymin[i] = ym; v_n[i] = n;
        }
    }
    }
    for (int i = tid; i < na; i += NT) { int lo, hi; area_rows(g, P.a0 + i, lo, hi); ar_lo[i] = lo; ar_hi[i] = hi; }
    // final rows of this slab that lie in the output letterbox are black
    for (int i = tid; i < (F1 - F0) * out; i += NT) {
        const int f = F0 + i / out, dx = i % out;
        const int dy = f - g.oy2;
        if (dy < 0 || dy >= g.oh) store_pixel(p, lut, crop, f, dx, 0, 0, 0);
    }
    __syncthreads();
    if (na <= 0) return;

    const uint8_t* fbase = p.frames + (int64_t)g.frame * p.fstride;
    const int64_t row0 = (int64_t)g.y0 * p.pitch + (int64_t)g.x0 * 3;
    const bool vec_ok = ((p.pitch & 15) == 0) && ((((uintptr_t)p.frames + (uintptr_t)((int64_t)g.frame * p.fstride)) & 15) == 0);
    const int shift = vec_ok ? (int)(row0 & 15) : 0;   // byte offset of the window inside the aligned copy
    const int chunks = (shift + rw * 3 + 15) >> 4;
    const bool h_fast = g.hact && (g.h_ks <= 8);
    const uint32_t nw_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(nw, 1)) + 1u;
    const int nww = (nw3 + 3) >> 2;                      // words per T row (fast V path)
    const uint32_t nww_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(nww, 1)) + 1u;
    const uint32_t out_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(out, 2)) + 1u;
    const uint32_t ch_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(chunks, 2)) + 1u;
    const int spw = sp >> 2;
    const uint32_t spw_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(spw, 2)) + 1u;
    const int t_off = g.hact ? 0 : shift;                // byte offset of pixel 0 in a T row
    // canvas the area pass reads: S ring, or the raw rows themselves when there is no letterbox stage
    const uint8_t* CV = g.pad1 ? S : T + t_off;
    const int CVp = g.pad1 ? sp : tp;
    // ring slot of a row: mask when the capacity is a power of two (the usual case), modulo otherwise
    const bool ct_p2 = (P.CT & (P.CT - 1)) == 0, cs_p2 = (P.CS & (P.CS - 1)) == 0;
    auto slotT = [&](int row) { return ct_p2 ? (row & (P.CT - 1)) : (row % P.CT); };
    auto slotS = [&](int row) { return cs_p2 ? (row & (P.CS - 1)) : (row % P.CS); };
    auto slotCV = [&](int row) { return g.pad1 ? slotS(row) : slotT(row); };
    const bool cx_p2 = P.CX > 0 && (P.CX & (P.CX - 1)) == 0;
    auto slotX = [&](int row) { return cx_p2 ? (row & (P.CX - 1)) : (row % max(P.CX, 1)); };

    int t_done = P.t_begin;     // raw rows [t_begin, t_done) have been through the H pass (ring T)
    int s_done = P.s_begin;     // canvas rows [s_begin, s_done) produced (ring S)
    int a_done = P.a0;          // area rows emitted
    int xb_done = g.pad1 ? P.s_begin : P.t_begin;  // canvas rows whose x-pass is in XB (general regime)

    auto t_needed_from = [&](int s_next) -> int {   // lowest raw row still needed once canvas rows < s_next exist
        int v = min(max(s_next - g.oy, P.v_begin), P.v_end);
        if (v >= P.v_end) return P.t_end;
        return g.vact ? v_ymin[v - P.v_begin] : v;
    };

    while (a_done < P.a1) {
        // ================= 1. load a batch of raw rows, horizontal pass -> T ring
        {
            const int t_keep = g.pad1 ? t_needed_from(s_done) : ar_lo[a_done - P.a0];
            int nb = min(P.RB, P.t_end - t_done);
            nb = min(nb, P.CT - (t_done - t_keep));
            if (nb > 0) {
                uint8_t* dstbase = g.hact ? RAW : T;
                const int dpitch = g.hact ? rawp : tp;
                if (rw > 0) {
                    if (vec_ok) {
                        const int total = nb * chunks;
                        for (int i = tid; i < total; i += NT) {
                            const int r = chunks > 1 ? (int)__umulhi((uint32_t)i, ch_magic) : i;
                            const int c = i - r * chunks;
                            const int t = t_done + r;
                            const int64_t goff = (int64_t)g.frame * p.fstride + row0 - shift + (int64_t)t * p.pitch + (int64_t)c * 16;
                            uint4 v;
                            if (goff + 16 <= p.frames_bytes) v = __ldg((const uint4*)(p.frames + goff));
                            else {
                                uint8_t tmp[16];
                                for (int k = 0; k < 16; k++) tmp[k] = (goff + k < p.frames_bytes) ? p.frames[goff + k] : 0;
                                v = *(uint4*)tmp;
                            }
                            uint8_t* d = dstbase + (size_t)(g.hact ? r : slotT(t)) * dpitch + c * 16;
                            *(uint4*)d = v;
                        }
                    } else {
                        const int total = nb * rw * 3;
                        for (int i = tid; i < total; i += NT) {
                            const int r = i / (rw * 3), c = i - r * (rw * 3);
                            const int t = t_done + r;
                            dstbase[(size_t)(g.hact ? r : slotT(t)) * dpitch + c] = fbase[row0 + (int64_t)t * p.pitch + c];
                        }
                    }
                }
                if (g.hact) {
                    __syncthreads();
                    const int total = nb * nw;
                    if (h_fast) {
                        for (int i = tid; i < total; i += NT) {
                            const int r = nw > 1 ? (int)__umulhi((uint32_t)i, nw_magic) : i;
                            const int xx = i - r * nw;
                            const int4 ka = *(const int4*)(h_kk + (size_t)xx * 8);
                            const int4 kb = *(const int4*)(h_kk + (size_t)xx * 8 + 4);
                            const int k[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
                            const int boff = shift + h_xmin[xx] * 3;
                            const uint32_t* wp = (const uint32_t*)(RAW + (size_t)r * rawp + (boff & ~3));
                            const int sh = (boff & 3) * 8;
                            uint32_t w[7];
#pragma unroll
                            for (int q = 0; q < 7; q++) w[q] = wp[q];
                            uint32_t b[6];
#pragma unroll
                            for (int q = 0; q < 6; q++) b[q] = __funnelshift_r(w[q], w[q + 1], sh);
                            int32_t s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const int e = 3 * j;  // byte index of channel 0 of tap j
                                s0 += (int)__byte_perm(b[e >> 2], 0, 0x4440 | (e & 3)) * k[j];
                                s1 += (int)__byte_perm(b[(e + 1) >> 2], 0, 0x4440 | ((e + 1) & 3)) * k[j];
                                s2 += (int)__byte_perm(b[(e + 2) >> 2], 0, 0x4440 | ((e + 2) & 3)) * k[j];
                            }
                            uint8_t* d = T + (size_t)slotT(t_done + r) * tp + xx * 3;
                            d[0] = (uint8_t)clip8w(s0); d[1] = (uint8_t)clip8w(s1); d[2] = (uint8_t)clip8w(s2);
                        }
                    } else {
                        for (int i = tid; i < total; i += NT) {
                            const int r = i / nw, xx = i - r * nw;
                            const uint8_t* src = RAW + (size_t)r * rawp + shift + h_xmin[xx] * 3;
                            const int32_t* k = h_kk + (size_t)xx * KSH;
                            const int n = h_n[xx];
                            int32_t s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
                            for (int j = 0; j < n; j++) {
                                const int32_t kj = k[j];
                                s0 += src[j * 3] * kj; s1 += src[j * 3 + 1] * kj; s2 += src[j * 3 + 2] * kj;
                            }
                            uint8_t* d = T + (size_t)slotT(t_done + r) * tp + xx * 3;
                            d[0] = (uint8_t)clip8w(s0); d[1] = (uint8_t)clip8w(s1); d[2] = (uint8_t)clip8w(s2);
                        }
                    }
                }
                t_done += nb;
            }
        }
        __syncthreads();

        // ================= 2. vertical pass: canvas rows whose taps are all in T -> S ring
        if (g.pad1) {
            const int lo_keep = ar_lo[a_done - P.a0];       // lowest canvas row still needed
            const int keep_cap = (P.CX > 0) ? P.CX : P.CS;
            int s_new = s_done;
            while (s_new < P.s_end && (s_new - lo_keep) < keep_cap && (s_new - s_done) < P.CS) {
                const int v = s_new - g.oy;
                if (v >= 0 && v < nh) {
                    const int need = g.vact ? (v_ymin[v - P.v_begin] + v_n[v - P.v_begin]) : (v + 1);
                    if (need > t_done) break;
                }
                s_new++;
            }
            const int ns = s_new - s_done;
            if (ns > 0) {
                // (a) black rows and letterbox borders
                const int lb = g.ox * 3, rb = g.ox * 3 + nw3;   // data occupies [lb, rb) of an image row
                for (int i = tid; i < ns * spw; i += NT) {
                    const int r = (int)__umulhi((uint32_t)i, spw_magic), wc = i - r * spw;
                    const int s = s_done + r, v = s - g.oy;
                    uint32_t* d = (uint32_t*)(S + (size_t)slotS(s) * sp) + wc;
                    if (v < 0 || v >= nh) *d = 0;
                    else if (wc * 4 + 4 <= lb || wc * 4 >= rb) *d = 0;
                    else if (wc * 4 < lb || wc * 4 + 4 > rb) {    // word straddles a border: zero only the border bytes
                        uint8_t* db = (uint8_t*)d;
                        for (int q = 0; q < 4; q++) if (wc * 4 + q < lb || wc * 4 + q >= rb) db[q] = 0;
                    }
                }
                // (b) image rows
                if (g.hact) {       // T rows are 16-byte aligned: 16 bytes (4 words) per item
                    const bool aligned_dst = ((lb & 3) == 0);
                    const int nq = (nw3 + 15) >> 4;                      // 16-byte items per row
                    const uint32_t nq_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(nq, 2)) + 1u;
                    const int total = ns * nq;
                    for (int i = tid; i < total; i += NT) {
                        const int r = nq > 1 ? (int)__umulhi((uint32_t)i, nq_magic) : i;
                        const int qc = i - r * nq;
                        const int s = s_done + r, v = s - g.oy;
                        if (v < 0 || v >= nh) continue;
                        uint32_t o[4];
                        if (g.vact) {
                            const int vi = v - P.v_begin;
                            const int32_t* k = v_kk + (size_t)vi * g.v_ks;
                            const int n = v_n[vi];
                            int slot = slotT(v_ymin[vi]);
                            int32_t c[16];
#pragma unroll
                            for (int e = 0; e < 16; e++) c[e] = 1 << 21;
                            for (int j = 0; j < n; j++) {
                                const uint4 w4 = *((const uint4*)(T + (size_t)slot * tp) + qc);
                                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
                                const int32_t kj = k[j];
#pragma unroll
                                for (int e = 0; e < 4; e++) {
                                    c[4 * e] += (int)__byte_perm(w[e], 0, 0x4440) * kj;
                                    c[4 * e + 1] += (int)__byte_perm(w[e], 0, 0x4441) * kj;
                                    c[4 * e + 2] += (int)__byte_perm(w[e], 0, 0x4442) * kj;
                                    c[4 * e + 3] += (int)(w[e] >> 24) * kj;
                                }
                                if (++slot == P.CT) slot = 0;
                            }
#pragma unroll
                            for (int e = 0; e < 4; e++)
                                o[e] = clip8w(c[4 * e]) | (clip8w(c[4 * e + 1]) << 8) | (clip8w(c[4 * e + 2]) << 16) | (clip8w(c[4 * e + 3]) << 24);
                        } else {
                            const uint4 w4 = *((const uint4*)(T + (size_t)slotT(v) * tp) + qc);
                            o[0] = w4.x; o[1] = w4.y; o[2] = w4.z; o[3] = w4.w;
                        }
                        uint8_t* drow = S + (size_t)slotS(s) * sp + lb;
                        if (aligned_dst && qc * 16 + 16 <= nw3) {
                            uint32_t* dw = (uint32_t*)drow + qc * 4;   // lb is a multiple of 4, not necessarily of 16
                            dw[0] = o[0]; dw[1] = o[1]; dw[2] = o[2]; dw[3] = o[3];
                        } else {
                            for (int q = 0; q < 16; q++) if (qc * 16 + q < nw3) drow[qc * 16 + q] = (uint8_t)(o[q >> 2] >> (8 * (q & 3)));
                        }
                    }
                } else {            // no H pass: T holds raw rows at byte offset t_off
                    const int total = ns * nw3;
                    for (int i = tid; i < total; i += NT) {
                        const int r = i / nw3, x = i - r * nw3;
                        const int s = s_done + r, v = s - g.oy;
                        if (v < 0 || v >= nh) continue;
                        uint8_t val;
                        if (g.vact) {
                            const int vi = v - P.v_begin;
                            const int32_t* k = v_kk + (size_t)vi * g.v_ks;
                            const int n = v_n[vi];
                            int slot = slotT(v_ymin[vi]);
                            int32_t c = 1 << 21;
                            for (int j = 0; j < n; j++) {
                                c += T[(size_t)slot * tp + t_off + x] * k[j];
                                if (++slot == P.CT) slot = 0;
                            }
                            val = (uint8_t)clip8w(c);
                        } else {
                            val = T[(size_t)slotT(v) * tp + t_off + x];
                        }
                        S[(size_t)slotS(s) * sp + lb + x] = val;
                    }
                }
                s_done = s_new;
            }
            __syncthreads();
        }
        const int cv_done = g.pad1 ? s_done : t_done;

        // ================= 2b. general area regime: x-pass of the canvas rows that just became available
        if (P.CX > 0) {
            const int nrow = cv_done - xb_done;
            for (int i = tid; i < nrow * out; i += NT) {
                const int r = out > 1 ? (int)__umulhi((uint32_t)i, out_magic) : i;
                const int dx = i - r * out;
                const int cr = xb_done + r;
                const uint8_t* row = CV + (size_t)slotCV(cr) * CVp;
                const int nx = xt_n[dx];
                const int* xsi = xt_si + dx * xcap;
                const float* xal = xt_al + dx * xcap;
                float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                for (int k = 0; k < nx; k++) {
                    const uint8_t* q = row + xsi[k] * 3;
                    const float al = xal[k];
                    b0 = __fadd_rn(b0, __fmul_rn((float)q[0], al));
                    b1 = __fadd_rn(b1, __fmul_rn((float)q[1], al));
                    b2 = __fadd_rn(b2, __fmul_rn((float)q[2], al));
                }
                float* o = XB + (size_t)slotX(cr) * xbp + dx * 3;
                o[0] = b0; o[1] = b1; o[2] = b2;
            }
            xb_done = cv_done;
            __syncthreads();
        }

        // ================= 3. area pass for every output row whose canvas rows are ready
        int a_new = a_done;
        while (a_new < P.a1 && ar_hi[a_new - P.a0] <= cv_done) a_new++;
        const int npix = (a_new - a_done) * out;
        for (int i = tid; i < npix; i += NT) {
            const int ar = out > 1 ? (int)__umulhi((uint32_t)i, out_magic) : i;
            const int dx = i - ar * out;
            const int dy = a_done + ar;
            const int f = dy + g.oy2;
            const int ai = dy - P.a0;
            int v0, v1, v2;
            if (g.regime == REG_COPY) {
                const uint8_t* q = CV + (size_t)slotCV(dy) * CVp + dx * 3;
                v0 = q[0]; v1 = q[1]; v2 = q[2];
            } else if (g.regime == REG_FAST) {
                int a0 = 0, a1 = 0, a2 = 0;
                for (int sy = 0; sy < g.isy; sy++) {
                    const uint8_t* q = CV + (size_t)slotCV(dy * g.isy + sy) * CVp + (size_t)dx * g.isx * 3;
                    for (int sx = 0; sx < g.isx; sx++) { a0 += q[sx * 3]; a1 += q[sx * 3 + 1]; a2 += q[sx * 3 + 2]; }
                }
                if (g.isx == 2 && g.isy == 2) { v0 = (a0 + 2) >> 2; v1 = (a1 + 2) >> 2; v2 = (a2 + 2) >> 2; }
                else {
                    const float sc = __fdiv_rn(1.f, (float)(g.isx * g.isy));
                    v0 = sat_u8f(__fmul_rn((float)a0, sc)); v1 = sat_u8f(__fmul_rn((float)a1, sc)); v2 = sat_u8f(__fmul_rn((float)a2, sc));
                }
            } else if (g.regime == REG_GENERAL) {
                const int ny = yt_n[ai];
                float m0 = 0.f, m1 = 0.f, m2 = 0.f;
                const int nx = xt_n[dx];
                const int* xsi = xt_si + dx * xcap;
                const float* xal = xt_al + dx * xcap;
                for (int j = 0; j < ny; j++) {
                    const float beta = yt_b[ai * ycap + j];
                    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                    if (P.CX > 0) {      // x-pass already in the fp32 ring
                        const float* xr = XB + (size_t)slotX(yt_s[ai * ycap + j]) * xbp + dx * 3;
                        b0 = xr[0]; b1 = xr[1]; b2 = xr[2];
                    } else {
                        const uint8_t* row = CV + (size_t)slotCV(yt_s[ai * ycap + j]) * CVp;
                        for (int k = 0; k < nx; k++) {
                            const uint8_t* q = row + xsi[k] * 3;
                            const float al = xal[k];
                            b0 = __fadd_rn(b0, __fmul_rn((float)q[0], al));
                            b1 = __fadd_rn(b1, __fmul_rn((float)q[1], al));
                            b2 = __fadd_rn(b2, __fmul_rn((float)q[2], al));
                        }
                    }
                    if (j == 0) { m0 = __fmul_rn(beta, b0); m1 = __fmul_rn(beta, b1); m2 = __fmul_rn(beta, b2); }
                    else {
                        m0 = __fadd_rn(m0, __fmul_rn(beta, b0)); m1 = __fadd_rn(m1, __fmul_rn(beta, b1)); m2 = __fadd_rn(m2, __fmul_rn(beta, b2));
                    }
                }
                v0 = sat_u8f(m0); v1 = sat_u8f(m1); v2 = sat_u8f(m2);
            } else {  // REG_LINEAR
                const int sx = lx_s[dx];
                const int sx1 = sx + 1 < sd ? sx + 1 : sx;
                const int a0 = lx_a[dx * 2], a1 = lx_a[dx * 2 + 1];
                const int sy = yt_s[ai * 2];
                const int b0 = ((int*)yt_b)[ai * 2], b1 = ((int*)yt_b)[ai * 2 + 1];
                int r[2][3];
                for (int k = 0; k < 2; k++) {
                    int yy = sy + k;
                    yy = yy >= 0 ? (yy < sd ? yy : sd - 1) : 0;
                    const uint8_t* row = CV + (size_t)slotCV(yy) * CVp;
                    for (int c = 0; c < 3; c++) r[k][c] = row[sx * 3 + c] * a0 + row[sx1 * 3 + c] * a1;
                }
                int vv[3];
                for (int c = 0; c < 3; c++) {
                    int t = (((b0 * (r[0][c] >> 4)) >> 16) + ((b1 * (r[1][c] >> 4)) >> 16) + 2) >> 2;
                    vv[c] = t < 0 ? 0 : (t > 255 ? 255 : t);
                }
                v0 = vv[0]; v1 = vv[1]; v2 = vv[2];
            }
            store_pixel(p, lut, crop, f, dx, v0, v1, v2);
        }
        a_done = a_new;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(PP_MAX_THREADS, 2) preprocess_kernel(const PPParams p) {
    if (p.first_pass_smem > 0) {
        // large-window pass: one CTA per SM (it asks for the whole carve-out) walks the slabs; usually the first pass
        // deferred nothing and the kernel is gone after one load
        if (*((volatile int*)p.deferred) == 0) return;
        for (int slab = blockIdx.x; slab < p.n_crops * PP_SPLIT; slab += gridDim.x) {
            preprocess_slab(p, slab);
            __syncthreads();
        }
        return;
    }
    preprocess_slab(p, blockIdx.x);
    // first pass launched under the tensor-core kernel's tail (overlap_prev): complete only after that kernel has, so that
    // stream order stays transitive for the launches behind this one
    if (p.overlap_prev && threadIdx.x == 0) pdl_wait();
}

#include "preprocess_tc.inc"

int launch_preprocess(const PPParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return PA_ERR_CUDA;
        // every kernel of the path asks for the maximum shared-memory carve-out: an SM never has to drain to
        // re-partition L1/shared between kernels, so the staging kernel's CTAs can stay resident across them
        cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(preprocess_plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        attr_set = true;
    }
    // Programmatic stream serialization, and the kernel never waits: its inputs (geometry, tables) come from the plan
    // kernel, which is complete before the tensor-core kernel in front of it starts, and it touches only crops that kernel
    // does not. Its CTAs (usually with nothing to do) drain through the SMs the tensor-core kernel's tail leaves free.
    // The large-window pass needs the first pass's counter and is launched the ordinary way.
    if (p.first_pass_smem == 0 && p.overlap_prev)
        return launch_pdl(preprocess_kernel, dim3(p.n_crops * PP_SPLIT), dim3(p.threads), (size_t)p.smem_bytes, stream, p) == cudaSuccess ? PA_OK : PA_ERR_CUDA;
    int grid = p.n_crops * PP_SPLIT;
    if (p.first_pass_smem > 0 && p.num_sms > 0 && grid > p.num_sms) grid = p.num_sms;
    preprocess_kernel<<<grid, p.threads, p.smem_bytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

int launch_preprocess_plan(const PPParams& p, cudaStream_t stream) {
    return launch_pdl(preprocess_plan_kernel, dim3(p.n_crops), dim3(128), 0, stream, p) == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}
size_t preprocess_geom_bytes() { return sizeof(CropGeom); }


// ------------------------------------------------------------------------------------------------------------
// Window staging: frames that live in PINNED HOST memory are not copied whole. This kernel pulls, over PCIe,
// exactly the 16-byte chunks preprocess_kernel reads (the clipped window rows of every crop) and writes them at
// the same offsets of a device buffer with the frame batch's geometry, so a 1080p frame costs ~0.6 MB of bus
// traffic instead of 6.2 MB and the copy of chunk i+1 can overlap the kernels of chunk i on another stream.
// The kernel is a set of SMALL persistent workers (one warp, 8 x 16 B loads in flight per thread): a worker has
// to sit beside the 2 x 256-thread preprocess CTAs or the 576-thread conv CTA that own an SM's registers while the
// previous batch computes, and ~1 MB in flight GPU-wide covers PCIe latency. The block scheduler would pack many
// such CTAs onto whichever SMs happen to be free at launch and starve the 1-CTA-per-SM conv kernels there, so each
// CTA first registers on its SM (%smid) and exits if ST_PER_SM workers already live there; work items are handed
// out through a global counter, so it does not matter which CTAs survive.
constexpr int ST_SPLIT = 8;      // work items per crop
constexpr int ST_THREADS = 32;     // one warp per worker: measured best beside the compute kernels (PA_ST_THREADS overrides)
constexpr int ST_PER_SM = 1;
constexpr int ST_CTAS_PER_SM = 4;  // launched; all but ST_PER_SM per SM retire at once

// streaming accesses that leave the SM's small L1 to the kernels computing beside this one
__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream16(uint8_t* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int ST_UNROLL>   // 16-byte loads in flight per thread
__global__ void __maxnreg__(64) stage_windows_kernel(const StageParams p) {
  __shared__ int s_item;
  if (threadIdx.x == 0) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      s_item = ((int)smid < p.max_sm && atomicAdd(&p.sched[1 + (smid & 255)], 1) < ST_PER_SM) ? 0 : -1;
  }
  __syncthreads();
  if (s_item < 0) return;
  const int n_items = p.n_crops * ST_SPLIT;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(&p.sched[0], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int crop = item / ST_SPLIT, part = item - crop * ST_SPLIT;
    const int32_t* box = p.boxes + (int64_t)crop * PA_BOX_STRIDE;
    const int frame = box[0] - p.frame_base;
    const int cw = box[3], ch = box[4];
    const int sd = cw > ch ? cw : ch;
    if (frame < 0 || frame >= p.n_frames || sd < 0) continue;
    int x0, y0, rw, rh;
    crop_window(box[1], box[2], sd, p.H, p.W, p.padding, x0, y0, rw, rh);
    if (rw <= 0 || rh <= 0) continue;
    const int64_t fo = (int64_t)frame * p.fstride;
    const int64_t row0 = (int64_t)y0 * p.pitch + (int64_t)x0 * 3;
    const bool vec_ok = ((p.pitch & 15) == 0) && ((((uintptr_t)p.src + (uintptr_t)fo) & 15) == 0) &&
                        ((((uintptr_t)p.dst + (uintptr_t)fo) & 15) == 0);
    if (vec_ok) {
        const int shift = (int)(row0 & 15);
        const int chunks = (shift + rw * 3 + 15) >> 4;
        const uint32_t ch_magic = (uint32_t)(0xFFFFFFFFu / (uint32_t)max(chunks, 2)) + 1u;
        const int total = rh * chunks;
        const int begin = (int)((int64_t)part * total / ST_SPLIT), end = (int)((int64_t)(part + 1) * total / ST_SPLIT);
        const int64_t base = fo + row0 - shift;
        for (int i0 = begin + threadIdx.x; i0 < end; i0 += blockDim.x * ST_UNROLL) {
            uint4 v[ST_UNROLL];
            int64_t off[ST_UNROLL];
#pragma unroll
            for (int u = 0; u < ST_UNROLL; u++) {
                const int i = i0 + u * blockDim.x;
                off[u] = -1;
                if (i < end) {
                    const int r = chunks > 1 ? (int)__umulhi((uint32_t)i, ch_magic) : i;
                    const int c = i - r * chunks;
                    const int64_t o = base + (int64_t)r * p.pitch + (int64_t)c * 16;
                    if (o + 16 <= p.frames_bytes) { off[u] = o; v[u] = ld_stream16(p.src + o); }
                    else for (int k = 0; k < 16 && o + k < p.frames_bytes; k++) p.dst[o + k] = p.src[o + k];  // tail of the last row
                }
            }
#pragma unroll
            for (int u = 0; u < ST_UNROLL; u++)
                if (off[u] >= 0) st_stream16(p.dst + off[u], v[u]);
        }
    } else {
        const int rowb = rw * 3;
        const int total = rh * rowb;
        const int begin = (int)((int64_t)part * total / ST_SPLIT), end = (int)((int64_t)(part + 1) * total / ST_SPLIT);
        for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
            const int r = i / rowb, c = i - r * rowb;
            const int64_t o = fo + row0 + (int64_t)r * p.pitch + c;
            p.dst[o] = p.src[o];
        }
    }
  }
}

// Same job with the TMA unit: a worker is ONE thread that pulls 16-byte aligned row segments with cp.async.bulk
// (pinned host -> shared memory ring, mbarrier completion) and pushes them on with cp.async.bulk (shared -> HBM).
// Larger requests cross PCIe (48.9 GB/s against 46.4 for the 16-byte loads above, tools/micro/pcie_probe.cu), no data
// passes through registers or L1, and a worker needs one thread and a 5.5 KB ring beside the compute CTAs.
constexpr int SB_STAGES = 4, SB_LAG = 2, SB_PIECE = 1408;   // ring stages, stores in flight before a stage is reused, bytes per stage
__global__ void __launch_bounds__(32) stage_windows_bulk_kernel(const StageParams p) {
    __shared__ __align__(128) uint8_t ring[SB_STAGES * SB_PIECE];
    __shared__ uint64_t full[SB_STAGES];
    if (threadIdx.x != 0) return;
    {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if (!((int)smid < p.max_sm && atomicAdd(&p.sched[1 + (smid & 255)], 1) < ST_PER_SM)) return;
    }
    for (int i = 0; i < SB_STAGES; i++) mbar_init(&full[i], 1);
    fence_barrier_init();
    uint32_t uses = 0;      // pieces that have gone through the ring so far (stage = uses % SB_STAGES, parity from uses / SB_STAGES)
    const int n_items = p.n_crops * ST_SPLIT;
    for (;;) {
        const int item = atomicAdd(&p.sched[0], 1);
        if (item >= n_items) break;
        const int crop = item / ST_SPLIT, part = item - crop * ST_SPLIT;
        const int32_t* box = p.boxes + (int64_t)crop * PA_BOX_STRIDE;
        const int frame = box[0] - p.frame_base;
        const int cw = box[3], ch = box[4];
        const int sd = cw > ch ? cw : ch;
        if (frame < 0 || frame >= p.n_frames || sd < 0) continue;
        int x0, y0, rw, rh;
        crop_window(box[1], box[2], sd, p.H, p.W, p.padding, x0, y0, rw, rh);
        if (rw <= 0 || rh <= 0) continue;
        const int64_t fo = (int64_t)frame * p.fstride;
        const int64_t row0 = (int64_t)y0 * p.pitch + (int64_t)x0 * 3;
        const int shift = (int)(row0 & 15);
        const int rowb = ((shift + rw * 3 + 15) >> 4) << 4;                 // 16-byte aligned cover of a window row
        const int ppr = (rowb + SB_PIECE - 1) / SB_PIECE;                   // pieces per row
        const int plen = (((rowb + ppr - 1) / ppr) + 15) & ~15;             // bytes per piece (the last one may be shorter)
        const int r_begin = (int)((int64_t)part * rh / ST_SPLIT), r_end = (int)((int64_t)(part + 1) * rh / ST_SPLIT);
        const int n_pieces = (r_end - r_begin) * ppr;
        const int64_t base = fo + row0 - shift;
        // Bytes the PREVIOUS record's window already brings in (two fighters close to each other share up to a third of
        // their windows): if that record sits in the same frame, rows ya0 .. ya1 of its 16-byte aligned cover [ax0, ax1) are
        // staged by its own work items, so a piece of ours is cut back where the cover reaches over its left or right end
        // (a cover strictly inside a piece would leave two pieces: left as it is, the bytes just cross twice).
        int ya0 = 0, ya1 = 0; int64_t ax0 = 0, ax1 = 0;
        if (crop > 0) {
            const int32_t* pb = box - PA_BOX_STRIDE;
            const int pw = pb[3], ph = pb[4], psd = pw > ph ? pw : ph;
            if (pb[0] == box[0] && psd >= 0) {
                int px0, py0, prw, prh;
                crop_window(pb[1], pb[2], psd, p.H, p.W, p.padding, px0, py0, prw, prh);
                if (prw > 0 && prh > 0) {
                    ya0 = py0; ya1 = py0 + prh;
                    const int64_t prow0 = (int64_t)py0 * p.pitch + (int64_t)px0 * 3;      // first byte of its first row (frame-relative)
                    const int pshift = (int)(prow0 & 15);
                    ax0 = (int64_t)px0 * 3 - pshift;                                      // row-relative byte range of its cover
                    ax1 = ax0 + ((((int64_t)pshift + prw * 3 + 15) >> 4) << 4);
                }
            }
        }
        const int64_t sx0 = (int64_t)x0 * 3 - shift;      // row-relative first byte of OUR cover
        auto piece = [&](int j, int64_t& off, int& len, int& tail) {
            const int r = r_begin + j / ppr, q = j - (j / ppr) * ppr;
            off = base + (int64_t)r * p.pitch + (int64_t)q * plen;
            len = min(plen, rowb - q * plen);
            tail = 0;
            if (off + len > p.frames_bytes) {      // tail of the last frame: what a 16-byte multiple cannot fetch is copied by hand
                const int want = len;
                len = (int)max((p.frames_bytes - off) & ~(int64_t)15, (int64_t)0);
                tail = (int)min((int64_t)want, p.frames_bytes - off) - len;
                return;
            }
            const int y = y0 + r;
            if (y >= ya0 && y < ya1) {
                const int64_t p0 = sx0 + (int64_t)q * plen, p1 = p0 + len;
                if (ax0 <= p0 && ax1 >= p1) len = 0;
                else if (ax0 <= p0 && ax1 > p0) { off += ax1 - p0; len -= (int)(ax1 - p0); }
                else if (ax1 >= p1 && ax0 < p1) len = (int)(ax0 - p0);
            }
        };
        auto load = [&](int j) {
            int64_t off; int len, tail;
            piece(j, off, len, tail);
            const uint32_t st = (uses + (uint32_t)j) % SB_STAGES;
            if (len > 0) {
                mbar_arrive_expect_tx(&full[st], (uint32_t)len);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(ring + st * SB_PIECE)), "l"(p.src + off), "r"(len), "r"(smem_u32(&full[st])) : "memory");
            } else {
                mbar_arrive(&full[st]);      // nothing to fetch: complete the phase so the consumer's wait falls through
            }
        };
        int issued = 0;
        for (; issued < n_pieces && issued < SB_STAGES - SB_LAG; issued++) load(issued);
        for (int j = 0; j < n_pieces; j++) {
            const uint32_t u = uses + (uint32_t)j, st = u % SB_STAGES;
            mbar_wait(&full[st], (u / SB_STAGES) & 1);
            int64_t off; int len, tail;
            piece(j, off, len, tail);
            if (len > 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             :: "l"(p.dst + off), "r"(smem_u32(ring + st * SB_PIECE)), "r"(len) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(SB_LAG - 1) : "memory");   // the store SB_LAG pieces back has read its stage
            if (issued < n_pieces) { load(issued); issued++; }
            // bytes of the last frame's last rows that the aligned cover cannot fetch as a 16-byte multiple
            for (int k = len; k < len + tail; k++) p.dst[off + k] = p.src[off + k];
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        uses += (uint32_t)n_pieces;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int launch_stage_windows(const StageParams& p, int num_sms, cudaStream_t stream) {
    int grid = p.n_crops * ST_SPLIT;
    if (grid > ST_CTAS_PER_SM * num_sms) grid = ST_CTAS_PER_SM * num_sms;
    if (cudaMemsetAsync(p.sched, 0, PA_STAGE_SCHED_INTS * sizeof(int), stream) != cudaSuccess) return PA_ERR_CUDA;
    static int unroll = 0, threads = 0, max_sm = 0;
    static bool use_ldst = false;
    if (!unroll) {
        const char* e;
        unroll = 8; threads = ST_THREADS; max_sm = 1 << 20;
        cudaFuncSetAttribute(stage_windows_bulk_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#ifdef PA_EXPERIMENT
        use_ldst = getenv("PA_ST_LDST") != nullptr;
        unroll = (e = getenv("PA_ST_UNROLL")) ? atoi(e) : 8;
        threads = (e = getenv("PA_ST_THREADS")) ? atoi(e) : ST_THREADS;
        max_sm = (e = getenv("PA_ST_SMS")) ? atoi(e) : 1 << 20;
        if (unroll != 2 && unroll != 4) unroll = 8;
        if (threads < 32 || threads > 1024 || (threads & 31)) threads = ST_THREADS;
        if (max_sm < 1) max_sm = 1 << 20;
#else
        (void)e;
#endif
        cudaFuncSetAttribute(stage_windows_kernel<8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(stage_windows_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(stage_windows_kernel<2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    StageParams q = p;
    q.max_sm = max_sm;
    // TMA bulk copies need 16-byte aligned rows on both sides; anything else takes the load/store kernel
    const bool bulk_ok = ((p.pitch & 15) == 0) && ((p.fstride & 15) == 0) && (((uintptr_t)p.src & 15) == 0) && (((uintptr_t)p.dst & 15) == 0);
    if (bulk_ok && !use_ldst) {
        stage_windows_bulk_kernel<<<grid, 32, 0, stream>>>(q);
        return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
    }
    if (unroll == 2) stage_windows_kernel<2><<<grid, threads, 0, stream>>>(q);
    else if (unroll == 4) stage_windows_kernel<4><<<grid, threads, 0, stream>>>(q);
    else stage_windows_kernel<8><<<grid, threads, 0, stream>>>(q);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

}  // namespace pa
