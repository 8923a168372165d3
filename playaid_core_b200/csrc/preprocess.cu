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

namespace pa {

constexpr int PP_THREADS = 256;
constexpr int PP_SPLIT = 4;

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
    int half = sd / 2;
    int y0 = cy - half - padding; if (y0 < 0) y0 = 0;
    int y1 = cy + half + padding; if (y1 > H) y1 = H;
    int x0 = cx - half - padding; if (x0 < 0) x0 = 0;
    int x1 = cx + half + padding; if (x1 > W) x1 = W;
    if (y1 < 0) { y1 += H; if (y1 < 0) y1 = 0; }  // numpy negative-stop semantics
    if (x1 < 0) { x1 += W; if (x1 < 0) x1 = 0; }
    if (y0 > H) y0 = H;
    if (x0 > W) x0 = W;
    int rh = y1 - y0, rw = x1 - x0;
    if (rh < 0) rh = 0;
    if (rw < 0) rw = 0;
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
    const bool f16 = (p.out_dtype == PA_DTYPE_F16 || p.out_dtype == PA_DTYPE_F16X2);
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
    const bool split = (p.out_dtype == PA_DTYPE_BF16X2 || p.out_dtype == PA_DTYPE_F16X2);
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
    int ch = (p.out_layout == PA_LAYOUT_NHWC4) ? 4 : 3;
    int nplanes = (p.out_dtype == PA_DTYPE_BF16X2 || p.out_dtype == PA_DTYPE_F16X2) ? 2 : 1;
    for (int pl = 0; pl < nplanes; pl++) {
        uint8_t* base = (uint8_t*)p.outp + (int64_t)pl * p.plane_elems * esz;
        if (p.out_layout == PA_LAYOUT_NCHW) {
            for (int c = 0; c < 3; c++) {
                uint8_t* q = base + ((((int64_t)crop * 3 + c) * out + F0) * out) * esz;
                int64_t n = (int64_t)(F1 - F0) * out * esz;
                for (int64_t i = threadIdx.x; i < n; i += blockDim.x) q[i] = 0;
            }
        } else {
            uint8_t* q = base + (((int64_t)crop * out + F0) * out) * ch * esz;
            int64_t n = (int64_t)(F1 - F0) * out * ch * esz;
            for (int64_t i = threadIdx.x; i < n; i += blockDim.x) q[i] = 0;
        }
    }
}

struct BandPlan {
    int f0, f1;  // final rows
    int a0, a1;  // area-output rows (may be empty)
    int s0, s1;  // canvas rows
    int v0, v1;  // resized rows
    int t0, t1;  // raw rows
};

__global__ void __launch_bounds__(PP_THREADS) preprocess_kernel(const PPParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ CropGeom g;
    __shared__ BandPlan bp;
    __shared__ int s_band_rows;

    const int crop = blockIdx.x / PP_SPLIT, part = blockIdx.x % PP_SPLIT;
    const int tid = threadIdx.x;
    const int out = p.out;
    if (tid == 0) compute_geom(g, p.boxes + (int64_t)crop * PA_BOX_STRIDE, p.H, p.W, p.n_frames, out, p.padding);
    __syncthreads();
    const int F0 = (int)((int64_t)part * out / PP_SPLIT), F1 = (int)((int64_t)(part + 1) * out / PP_SPLIT);
    if (g.status != PA_CROP_OK) {
        if (p.first_pass_smem > 0) return;  // reported by the first pass
        if (part == 0 && tid == 0 && p.status) p.status[crop] = g.status;
        zero_rows(p, crop, F0, F1);
        return;
    }

    // ---- carve shared memory: [lut][tabH][xtab][per-band tables][RAW][T][S]
    int off = 0;
    float* lut = (float*)(smem + off); off += 768 * 4;
    // H-pass table: per output column xmin, n, ks coefficients
    const int nw = g.nw, nh = g.nh, sd = g.sd, rw = g.rw, rh = g.rh;
    int* h_xmin = nullptr; int* h_n = nullptr; int32_t* h_kk = nullptr;
    if (g.hact) {
        h_xmin = (int*)(smem + off); off += nw * 4;
        h_n = (int*)(smem + off); off += nw * 4;
        h_kk = (int32_t*)(smem + off); off += nw * g.h_ks * 4;
    }
    // area x tables
    int xcap = 0;
    int* xt_start = nullptr; int* xt_si = nullptr; float* xt_al = nullptr;
    int* lx_s = nullptr; short* lx_a = nullptr;
    if (g.regime == REG_GENERAL) {
        xcap = (int)ceil(g.scale_x) + 2;
        xt_start = (int*)(smem + off); off += (out + 1) * 4;
        xt_si = (int*)(smem + off); off += out * xcap * 4;
        xt_al = (float*)(smem + off); off += out * xcap * 4;
    } else if (g.regime == REG_LINEAR) {
        lx_s = (int*)(smem + off); off += out * 4;
        lx_a = (short*)(smem + off); off += align16(out * 2 * 2);
    }
    off = align16(off);

    // ---- band sizing
    const int rawp = align16(rw * 3 + 15) + 16;  // aligned copy incl. alignment shift
    const int nw3 = nw * 3, sd3 = sd * 3;
    const int tp = align16(nw3), sp = align16(sd3);
    const int ycap = (g.regime == REG_GENERAL) ? (int)ceil(g.scale_y) + 2 : 2;
    if (tid == 0) {
        // largest band (final rows per pass) whose staging fits in `limit` bytes of shared memory
        auto fit = [&](int limit) {
            int B = F1 - F0;
            for (; B >= 1; B = (B > 1 ? (B + 1) / 2 : 0)) {
                int ns, nt;  // upper bounds of canvas / raw rows needed for B final rows
                if (g.regime == REG_COPY) ns = B;
                else if (g.regime == REG_FAST) ns = B * g.isy;
                else ns = (int)ceil(B * g.scale_y) + 2;
                if (ns > sd) ns = sd;
                nt = ns;
                if (g.vact) nt = (int)ceil(ns * g.v_scale) + 2 * (int)ceil(g.v_sup) + 2;
                if (nt > rh) nt = rh;
                int need = off + B * (2 + ycap * 2) * 4 + 64;  // y tables
                if (g.vact) need += ns * (2 + g.v_ks) * 4;      // V tables
                need += nt * rawp;                              // RAW
                if (g.hact) need += nt * tp;                    // T
                if (g.pad1) need += ns * sp;                    // S
                if (need <= limit) return B;
                if (B == 1) break;
            }
            return 0;
        };
        // second pass (whole 227 KB carve-out): only crops the first pass could not stage
        if (p.first_pass_smem > 0 && fit(p.first_pass_smem) > 0) s_band_rows = -1;
        else s_band_rows = fit(p.smem_bytes);
    }
    __syncthreads();
    const int B = s_band_rows;
    if (B < 0) return;  // already produced by the first pass
    if (B == 0) {  // does not fit in shared memory even one output row at a time
        if (part == 0 && tid == 0 && p.status) p.status[crop] = PA_CROP_TOO_LARGE;
        zero_rows(p, crop, F0, F1);
        return;
    }
    if (part == 0 && tid == 0 && p.status) p.status[crop] = PA_CROP_OK;
    // ---- per-CTA tables
    for (int i = tid; i < 768; i += PP_THREADS) {
        int c = i >> 8, v = i & 255;
        float f = __fdiv_rn((float)v, 255.0f);
        lut[i] = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.stdv[c]);
    }
    if (g.hact) {
        for (int xx = tid; xx < nw; xx += PP_THREADS) {
            int xm, n;
            bicubic_coeffs(xx, rw, g.h_scale, g.h_fs, g.h_sup, g.h_ks, xm, n, h_kk + (size_t)xx * g.h_ks);
            h_xmin[xx] = xm; h_n[xx] = n;
        }
    }
    if (g.regime == REG_GENERAL) {
        for (int dx = tid; dx < out; dx += PP_THREADS) {
            int n = area_entries(dx, sd, g.scale_x, xt_si + dx * xcap, xt_al + dx * xcap, xcap);
            xt_start[dx] = n < xcap ? n : xcap;
        }
    } else if (g.regime == REG_LINEAR) {
        for (int dx = tid; dx < out; dx += PP_THREADS) {
            int s, a0, a1;
            linear_coef(dx, sd, g.scale_x, g.inv_scale_x, true, s, a0, a1);
            lx_s[dx] = s; lx_a[dx * 2] = (short)a0; lx_a[dx * 2 + 1] = (short)a1;
        }
    }
    __syncthreads();

    const uint8_t* fbase = p.frames + (int64_t)g.frame * p.fstride;
    const bool vec_ok = ((p.pitch & 15) == 0) && ((((uintptr_t)fbase) & 15) == 0);

    for (int f0 = F0; f0 < F1; f0 += B) {
        // ---------------- plan the band (thread 0)
        if (tid == 0) {
            BandPlan b;
            b.f0 = f0; b.f1 = min(f0 + B, F1);
            b.a0 = max(b.f0 - g.oy2, 0); b.a1 = min(b.f1 - g.oy2, g.oh);
            if (b.a1 < b.a0) b.a1 = b.a0;
            b.s0 = b.s1 = b.v0 = b.v1 = b.t0 = b.t1 = 0;
            if (b.a1 > b.a0) {
                if (g.regime == REG_COPY) { b.s0 = b.a0; b.s1 = b.a1; }
                else if (g.regime == REG_FAST) { b.s0 = b.a0 * g.isy; b.s1 = b.a1 * g.isy; }
                else if (g.regime == REG_GENERAL) {
                    int lo, hi, lo2, hi2;
                    area_span(b.a0, sd, g.scale_y, lo, hi);
                    area_span(b.a1 - 1, sd, g.scale_y, lo2, hi2);
                    b.s0 = lo; b.s1 = hi2;
                } else {
                    int s, a0, a1;
                    linear_coef(b.a0, sd, g.scale_y, g.inv_scale_y, false, s, a0, a1);
                    b.s0 = min(max(s, 0), sd - 1);
                    linear_coef(b.a1 - 1, sd, g.scale_y, g.inv_scale_y, false, s, a0, a1);
                    b.s1 = min(max(s + 1, 0), sd - 1) + 1;
                }
                b.s0 = max(b.s0, 0); b.s1 = min(b.s1, sd);
                b.v0 = max(b.s0 - g.oy, 0); b.v1 = min(b.s1 - g.oy, nh);
                if (b.v1 < b.v0) b.v1 = b.v0;
                if (b.v1 > b.v0) {
                    if (g.vact) {
                        double c0 = 0.0 + (b.v0 + 0.5) * g.v_scale;
                        int ymin = (int)(c0 - g.v_sup + 0.5); if (ymin < 0) ymin = 0;
                        double c1 = 0.0 + ((b.v1 - 1) + 0.5) * g.v_scale;
                        int ymax = (int)(c1 + g.v_sup + 0.5); if (ymax > rh) ymax = rh;
                        b.t0 = ymin; b.t1 = ymax;
                    } else { b.t0 = b.v0; b.t1 = b.v1; }
                }
            }
            bp = b;
        }
        __syncthreads();
        const BandPlan b = bp;
        const int ns = b.s1 - b.s0, nt = b.t1 - b.t0, nv = b.v1 - b.v0, na = b.a1 - b.a0;

        // ---------------- carve the band region
        int o2 = off;
        int* yt_n = (int*)(smem + o2); o2 += (B + 1) * 4;
        int* yt_s = (int*)(smem + o2); o2 += B * ycap * 4;      // general: si ; linear: sy
        float* yt_b = (float*)(smem + o2); o2 += B * ycap * 4;  // general: beta ; linear: b0,b1 as ints
        int* v_ymin = nullptr; int* v_n = nullptr; int32_t* v_kk = nullptr;
        if (g.vact) {
            v_ymin = (int*)(smem + o2); o2 += max(nv, 1) * 4;
            v_n = (int*)(smem + o2); o2 += max(nv, 1) * 4;
            v_kk = (int32_t*)(smem + o2); o2 += max(nv, 1) * g.v_ks * 4;
        }
        o2 = align16(o2);
        uint8_t* RAW = smem + o2; o2 += nt * rawp;
        uint8_t* T = RAW;
        if (g.hact) { T = smem + o2; o2 += nt * tp; }
        uint8_t* S = nullptr;
        if (g.pad1) { S = smem + o2; o2 += ns * sp; }
        if (o2 > p.smem_bytes) { __trap(); }  // sizing bound violated: fail loudly

        // ---------------- band tables
        if (g.regime == REG_GENERAL) {
            for (int i = tid; i < na; i += PP_THREADS) {
                int n = area_entries(b.a0 + i, sd, g.scale_y, yt_s + i * ycap, yt_b + i * ycap, ycap);
                yt_n[i] = n < ycap ? n : ycap;
            }
        } else if (g.regime == REG_LINEAR) {
            for (int i = tid; i < na; i += PP_THREADS) {
                int s, b0, b1;
                linear_coef(b.a0 + i, sd, g.scale_y, g.inv_scale_y, false, s, b0, b1);
                yt_s[i * 2] = s;
                ((int*)yt_b)[i * 2] = b0; ((int*)yt_b)[i * 2 + 1] = b1;
            }
        }
        if (g.vact) {
            for (int i = tid; i < nv; i += PP_THREADS) {
                int ym, n;
                bicubic_coeffs(b.v0 + i, rh, g.v_scale, g.v_fs, g.v_sup, g.v_ks, ym, n, v_kk + (size_t)i * g.v_ks);
                v_ymin[i] = ym; v_n[i] = n;
            }
        }

        // ---------------- load raw rows [t0, t1): 128-bit reads of the 16-byte-aligned cover
        int shift = 0;
        if (nt > 0 && rw > 0) {
            const int64_t row0 = (int64_t)g.y0 * p.pitch + (int64_t)g.x0 * 3;
            if (vec_ok) {
                shift = (int)(row0 & 15);
                const int chunks = (shift + rw * 3 + 15) >> 4;
                const int total = nt * chunks;
                for (int i = tid; i < total; i += PP_THREADS) {
                    int r = i / chunks, c = i - r * chunks;
                    int64_t goff = (int64_t)g.frame * p.fstride + row0 - shift + (int64_t)(b.t0 + r) * p.pitch + (int64_t)c * 16;
                    uint4 q;
                    if (goff + 16 <= p.frames_bytes) {
                        q = __ldg((const uint4*)(p.frames + goff));
                    } else {
                        uint8_t tmp[16];
                        for (int k = 0; k < 16; k++) tmp[k] = (goff + k < p.frames_bytes) ? p.frames[goff + k] : 0;
                        q = *(uint4*)tmp;
                    }
                    *(uint4*)(RAW + (size_t)r * rawp + c * 16) = q;
                }
            } else {
                const int total = nt * rw * 3;
                for (int i = tid; i < total; i += PP_THREADS) {
                    int r = i / (rw * 3), c = i - r * (rw * 3);
                    RAW[(size_t)r * rawp + c] = fbase[row0 + (int64_t)(b.t0 + r) * p.pitch + c];
                }
            }
        }
        if (g.pad1) {
            for (int i = tid * 16; i < ns * sp; i += PP_THREADS * 16) *(uint4*)(S + i) = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();

        // ---------------- horizontal bicubic: RAW -> T
        if (g.hact) {
            const int total = nt * nw;
            for (int i = tid; i < total; i += PP_THREADS) {
                int r = i / nw, xx = i - r * nw;
                const uint8_t* src = RAW + (size_t)r * rawp + shift + h_xmin[xx] * 3;
                const int32_t* k = h_kk + (size_t)xx * g.h_ks;
                const int n = h_n[xx];
                int32_t s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
                for (int j = 0; j < n; j++) {
                    int32_t kj = k[j];
                    s0 += src[j * 3] * kj; s1 += src[j * 3 + 1] * kj; s2 += src[j * 3 + 2] * kj;
                }
                uint8_t* d = T + (size_t)r * tp + xx * 3;
                d[0] = clip8_fix(s0); d[1] = clip8_fix(s1); d[2] = clip8_fix(s2);
            }
            __syncthreads();
        }
        const uint8_t* Tsrc = g.hact ? T : RAW + shift;
        const int Tp = g.hact ? tp : rawp;

        // ---------------- vertical bicubic (or copy): T -> S canvas rows
        if (g.pad1) {
            const int total = nv * nw3;
            for (int i = tid; i < total; i += PP_THREADS) {
                int r = i / nw3, x = i - r * nw3;
                uint8_t val;
                if (g.vact) {
                    const int32_t* k = v_kk + (size_t)r * g.v_ks;
                    const int n = v_n[r];
                    const uint8_t* src = Tsrc + (size_t)(v_ymin[r] - b.t0) * Tp + x;
                    int32_t s = 1 << 21;
                    for (int j = 0; j < n; j++) s += src[(size_t)j * Tp] * k[j];
                    val = clip8_fix(s);
                } else {
                    val = Tsrc[(size_t)(b.v0 + r - b.t0) * Tp + x];
                }
                S[(size_t)(b.v0 + r + g.oy - b.s0) * sp + g.ox * 3 + x] = val;
            }
            __syncthreads();
        }
        const uint8_t* CV = g.pad1 ? S : RAW + shift;  // canvas rows [s0, s1), row stride CVp
        const int CVp = g.pad1 ? sp : rawp;
        const int cv0 = g.pad1 ? b.s0 : b.t0;

        // ---------------- area pass + store, one thread per final pixel
        const int npix = (b.f1 - b.f0) * out;
        for (int i = tid; i < npix; i += PP_THREADS) {
            const int fr = i / out, dx = i - fr * out;
            const int f = b.f0 + fr;
            const int dy = f - g.oy2;
            int v0 = 0, v1 = 0, v2 = 0;  // letterbox rows are black
            if (dy >= 0 && dy < g.oh) {
                const int ai = dy - b.a0;
                if (g.regime == REG_COPY) {
                    const uint8_t* q = CV + (size_t)(dy - cv0) * CVp + dx * 3;
                    v0 = q[0]; v1 = q[1]; v2 = q[2];
                } else if (g.regime == REG_FAST) {
                    int a0 = 0, a1 = 0, a2 = 0;
                    for (int sy = 0; sy < g.isy; sy++) {
                        const uint8_t* q = CV + (size_t)(dy * g.isy + sy - cv0) * CVp + (size_t)dx * g.isx * 3;
                        for (int sx = 0; sx < g.isx; sx++) { a0 += q[sx * 3]; a1 += q[sx * 3 + 1]; a2 += q[sx * 3 + 2]; }
                    }
                    if (g.isx == 2 && g.isy == 2) { v0 = (a0 + 2) >> 2; v1 = (a1 + 2) >> 2; v2 = (a2 + 2) >> 2; }
                    else {
                        float sc = __fdiv_rn(1.f, (float)(g.isx * g.isy));
                        v0 = sat_u8f(__fmul_rn((float)a0, sc)); v1 = sat_u8f(__fmul_rn((float)a1, sc)); v2 = sat_u8f(__fmul_rn((float)a2, sc));
                    }
                } else if (g.regime == REG_GENERAL) {
                    const int nx = xt_start[dx];
                    const int* xsi = xt_si + dx * xcap;
                    const float* xal = xt_al + dx * xcap;
                    const int ny = yt_n[ai];
                    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
                    for (int j = 0; j < ny; j++) {
                        const float beta = yt_b[ai * ycap + j];
                        const uint8_t* row = CV + (size_t)(yt_s[ai * ycap + j] - cv0) * CVp;
                        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
                        for (int k = 0; k < nx; k++) {
                            const uint8_t* q = row + xsi[k] * 3;
                            const float al = xal[k];
                            b0 = __fadd_rn(b0, __fmul_rn((float)q[0], al));
                            b1 = __fadd_rn(b1, __fmul_rn((float)q[1], al));
                            b2 = __fadd_rn(b2, __fmul_rn((float)q[2], al));
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
                        const uint8_t* row = CV + (size_t)(yy - cv0) * CVp;
                        for (int c = 0; c < 3; c++) r[k][c] = row[sx * 3 + c] * a0 + row[sx1 * 3 + c] * a1;
                    }
                    int vv[3];
                    for (int c = 0; c < 3; c++) {
                        int t = (((b0 * (r[0][c] >> 4)) >> 16) + ((b1 * (r[1][c] >> 4)) >> 16) + 2) >> 2;
                        vv[c] = t < 0 ? 0 : (t > 255 ? 255 : t);
                    }
                    v0 = vv[0]; v1 = vv[1]; v2 = vv[2];
                }
            }
            store_pixel(p, lut, crop, f, dx, v0, v1, v2);
        }
        __syncthreads();
    }
}

int launch_preprocess(const PPParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return PA_ERR_CUDA;
        attr_set = true;
    }
    preprocess_kernel<<<p.n_crops * PP_SPLIT, PP_THREADS, p.smem_bytes, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? PA_OK : PA_ERR_CUDA;
}

}  // namespace pa
