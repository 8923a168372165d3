/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the fighter-crop resample chain.
 *
 * Plain-C restatement of the arithmetic the reference delegates to third-party
 * libraries on its hot path:
 *
 *   playaid/fighter.py:323-381  YoloCrop.square_crop
 *     -> PIL.ImageOps.pad  (Pillow, unpinned in requirements.txt; 12.2.0 here)
 *          contain() + Image.resize(BICUBIC) + paste on black
 *     -> imutils.resize(width=128) == cv2.resize(INTER_AREA)
 *          (imutils 0.5.4 requirements.txt:93; opencv 4.5.5/4.6.0 pinned, 4.13.0 here)
 *
 * Neither library is vendored under /root/reference, so the published algorithms
 * are restated here: Pillow's ImagingResample 8-bit path (two passes, 22-bit
 * fixed-point coefficients, u8 intermediate) and OpenCV's INTER_AREA dispatch
 * (copy / integer-scale fast path / fp32 area tables / upscale via fixed-point
 * linear with area-mode coefficients).
 *
 * Parity pin: tests/golden/ holds outputs of the *reference's own*
 * YoloCrop.square_crop run under oracle/ref_shims.py (see oracle/gen_golden.py);
 * tests/test_oracle_golden.py checks this file against them byte for byte, and
 * tests/test_oracle_vs_libs.py checks it against the installed cv2 / Pillow.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * link or call this file. The product path never does.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PA_OK 1
#define PA_INVALID 0
#define PA_ERR_ZERO_DIV (-2)
#define PA_ERR_ALLOC (-3)

/* ------------------------------------------------------------------ Pillow */

#define PRECISION_BITS (32 - 8 - 2)

static double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

static uint8_t clip8(int32_t v) {
    v >>= PRECISION_BITS;
    if (v < 0) return 0;
    if (v > 255) return 255;
    return (uint8_t)v;
}

/* Pillow precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter (support 2.0). */
static int pil_coeffs(int in_size, int out_size, int** bounds_out, int32_t** kk_out, int* ksize_out) {
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = 2.0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double* pre = (double*)malloc(sizeof(double) * (size_t)out_size * ksize);
    int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_size * ksize);
    int* bounds = (int*)malloc(sizeof(int) * 2 * (size_t)out_size);
    if (!pre || !kk || !bounds) { free(pre); free(kk); free(bounds); return PA_ERR_ALLOC; }
    for (int xx = 0; xx < out_size; xx++) {
        double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double* k = &pre[(size_t)xx * ksize];
        int x;
        for (x = 0; x < xmax; x++) {
            double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; x++)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; x++) k[x] = 0;
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    for (size_t i = 0; i < (size_t)out_size * ksize; i++) {
        if (pre[i] < 0)
            kk[i] = (int32_t)(-0.5 + pre[i] * (1 << PRECISION_BITS));
        else
            kk[i] = (int32_t)(0.5 + pre[i] * (1 << PRECISION_BITS));
    }
    free(pre);
    *bounds_out = bounds;
    *kk_out = kk;
    *ksize_out = ksize;
    return PA_OK;
}

/* Image.resize(size, BICUBIC) on a tightly packed HxWx3 u8 image. */
int pa_oracle_pil_bicubic(const uint8_t* src, int h, int w, int src_pitch, uint8_t* dst, int oh, int ow) {
    if (oh <= 0 || ow <= 0) return PA_INVALID; /* Pillow: ValueError("height and width must be > 0") */
    if (h <= 0 || w <= 0) return PA_INVALID;
    if (oh == h && ow == w) { /* Image.resize: same size -> copy() */
        for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * ow * 3, src + (size_t)y * src_pitch, (size_t)w * 3);
        return PA_OK;
    }
    const uint8_t* cur = src;
    int cur_pitch = src_pitch;
    uint8_t* tmp = NULL;
    if (ow != w) { /* horizontal pass first */
        int *bounds, ksize; int32_t* kk;
        int rc = pil_coeffs(w, ow, &bounds, &kk, &ksize);
        if (rc != PA_OK) return rc;
        uint8_t* out = (oh != h) ? (tmp = (uint8_t*)malloc((size_t)h * ow * 3)) : dst;
        if (!out) { free(bounds); free(kk); return PA_ERR_ALLOC; }
        for (int y = 0; y < h; y++) {
            const uint8_t* row = src + (size_t)y * src_pitch;
            for (int xx = 0; xx < ow; xx++) {
                int xmin = bounds[xx * 2], xmax = bounds[xx * 2 + 1];
                const int32_t* k = &kk[(size_t)xx * ksize];
                int32_t s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
                for (int x = 0; x < xmax; x++) {
                    const uint8_t* p = row + (size_t)(x + xmin) * 3;
                    s0 += p[0] * k[x]; s1 += p[1] * k[x]; s2 += p[2] * k[x];
                }
                uint8_t* o = out + ((size_t)y * ow + xx) * 3;
                o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
            }
        }
        free(bounds); free(kk);
        cur = out; cur_pitch = ow * 3;
    }
    if (oh != h) { /* vertical pass on the u8 intermediate */
        int *bounds, ksize; int32_t* kk;
        int rc = pil_coeffs(h, oh, &bounds, &kk, &ksize);
        if (rc != PA_OK) { free(tmp); return rc; }
        for (int yy = 0; yy < oh; yy++) {
            int ymin = bounds[yy * 2], ymax = bounds[yy * 2 + 1];
            const int32_t* k = &kk[(size_t)yy * ksize];
            for (int x = 0; x < ow * 3; x++) {
                int32_t s = 1 << (PRECISION_BITS - 1);
                for (int y = 0; y < ymax; y++) s += cur[(size_t)(y + ymin) * cur_pitch + x] * k[y];
                dst[(size_t)yy * ow * 3 + x] = clip8(s);
            }
        }
        free(bounds); free(kk);
    }
    free(tmp);
    return PA_OK;
}

/* ImageOps.contain target size (Python round() == rint in the default rounding mode). */
static void pil_contain_size(int w, int h, int sw, int sh, int* nw, int* nh) {
    double im_ratio = (double)w / h, dest_ratio = (double)sw / sh;
    *nw = sw; *nh = sh;
    if (im_ratio != dest_ratio) {
        if (im_ratio > dest_ratio) {
            int new_h = (int)rint((double)h / w * sw);
            if (new_h != sh) *nh = new_h;
        } else {
            int new_w = (int)rint((double)w / h * sh);
            if (new_w != sw) *nw = new_w;
        }
    }
}

/* ImageOps.pad(img, (sw, sh), color="black"): contain + BICUBIC + centred paste. */
int pa_oracle_pil_pad(const uint8_t* src, int h, int w, int src_pitch, uint8_t* dst, int sh, int sw) {
    if (h <= 0) return PA_ERR_ZERO_DIV;  /* contain(): im_ratio = image.width / image.height */
    if (sh == 0) return PA_ERR_ZERO_DIV; /* dest_ratio = size[0] / size[1] */
    if (w < 0 || sw <= 0) return PA_INVALID;
    int nw, nh;
    pil_contain_size(w, h, sw, sh, &nw, &nh);
    /* Image.resize: identical size -> copy() (even for an empty image); otherwise the C resampler
       raises ValueError("height and width must be > 0") for an empty target */
    const int same = (nw == w && nh == h);
    if (!same && (nw <= 0 || nh <= 0)) return PA_INVALID;
    if (nw == sw && nh == sh) return pa_oracle_pil_bicubic(src, h, w, src_pitch, dst, sh, sw);
    memset(dst, 0, (size_t)sw * sh * 3);
    if (nw <= 0 || nh <= 0) return PA_OK; /* empty paste: black canvas */
    uint8_t* res = (uint8_t*)malloc((size_t)nw * nh * 3);
    if (!res) return PA_ERR_ALLOC;
    int rc = pa_oracle_pil_bicubic(src, h, w, src_pitch, res, nh, nw);
    if (rc != PA_OK) { free(res); return rc; }
    int ox = 0, oy = 0;
    if (nw != sw) ox = (int)rint((sw - nw) * 0.5);
    else oy = (int)rint((sh - nh) * 0.5);
    for (int y = 0; y < nh; y++) memcpy(dst + ((size_t)(y + oy) * sw + ox) * 3, res + (size_t)y * nw * 3, (size_t)nw * 3);
    free(res);
    return PA_OK;
}

/* ------------------------------------------------------------------ OpenCV */

typedef struct { int si, di; float alpha; } DecimateAlpha;

static int cv_area_tab(int ssize, int dsize, int cn, double scale, DecimateAlpha* tab) {
    int k = 0;
    for (int dx = 0; dx < dsize; dx++) {
        double fsx1 = dx * scale;
        double fsx2 = fsx1 + scale;
        double cell = scale < (ssize - fsx1) ? scale : (ssize - fsx1);
        int sx1 = (int)ceil(fsx1), sx2 = (int)floor(fsx2);
        if (sx2 > ssize - 1) sx2 = ssize - 1;
        if (sx1 > sx2) sx1 = sx2;
        if (sx1 - fsx1 > 1e-3) {
            tab[k].di = dx * cn; tab[k].si = (sx1 - 1) * cn;
            tab[k++].alpha = (float)((sx1 - fsx1) / cell);
        }
        for (int sx = sx1; sx < sx2; sx++) {
            tab[k].di = dx * cn; tab[k].si = sx * cn;
            tab[k++].alpha = (float)(1.0 / cell);
        }
        if (fsx2 - sx2 > 1e-3) {
            double t = fsx2 - sx2; if (t > 1.) t = 1.; if (t > cell) t = cell;
            tab[k].di = dx * cn; tab[k].si = sx2 * cn;
            tab[k++].alpha = (float)(t / cell);
        }
    }
    return k;
}

static uint8_t sat_u8_from_float(float v) {
    long iv = lrintf(v); /* cvRound: round-half-even */
    return (uint8_t)(iv < 0 ? 0 : iv > 255 ? 255 : iv);
}

static short sat_s16_from_float(float v) {
    long iv = lrintf(v);
    return (short)(iv < -32768 ? -32768 : iv > 32767 ? 32767 : iv);
}

/* cv2.resize(src, (ow, oh), interpolation=INTER_AREA) on HxWx3 u8. */
int pa_oracle_cv_area(const uint8_t* src, int h, int w, int src_pitch, uint8_t* dst, int oh, int ow) {
    if (h <= 0 || w <= 0 || oh <= 0 || ow <= 0) return PA_INVALID;
    const int cn = 3;
    if (oh == h && ow == w) {
        for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * ow * 3, src + (size_t)y * src_pitch, (size_t)w * 3);
        return PA_OK;
    }
    double inv_scale_x = (double)ow / w, inv_scale_y = (double)oh / h;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y); /* saturate_cast<int>(double) == cvRound */
    int is_area_fast = fabs(scale_x - iscale_x) < DBL_EPSILON && fabs(scale_y - iscale_y) < DBL_EPSILON;

    if (scale_x >= 1 && scale_y >= 1) {
        if (is_area_fast) {
            int area = iscale_x * iscale_y;
            float scale = 1.f / (float)area;
            for (int dy = 0; dy < oh; dy++)
                for (int dx = 0; dx < ow; dx++)
                    for (int c = 0; c < cn; c++) {
                        int sum = 0;
                        for (int sy = 0; sy < iscale_y; sy++)
                            for (int sx = 0; sx < iscale_x; sx++)
                                sum += src[(size_t)(dy * iscale_y + sy) * src_pitch + (size_t)(dx * iscale_x + sx) * cn + c];
                        uint8_t v;
                        if (iscale_x == 2 && iscale_y == 2) v = (uint8_t)((sum + 2) >> 2);
                        else v = sat_u8_from_float((float)sum * scale);
                        dst[((size_t)dy * ow + dx) * cn + c] = v;
                    }
            return PA_OK;
        }
        DecimateAlpha* xtab = (DecimateAlpha*)malloc(sizeof(DecimateAlpha) * (size_t)(w * 2 + 2 * ow + 4));
        DecimateAlpha* ytab = (DecimateAlpha*)malloc(sizeof(DecimateAlpha) * (size_t)(h * 2 + 2 * oh + 4));
        float* buf = (float*)malloc(sizeof(float) * (size_t)ow * cn * 2);
        if (!xtab || !ytab || !buf) { free(xtab); free(ytab); free(buf); return PA_ERR_ALLOC; }
        float* sum = buf + (size_t)ow * cn;
        int xn = cv_area_tab(w, ow, cn, scale_x, xtab);
        int yn = cv_area_tab(h, oh, 1, scale_y, ytab);
        int wn = ow * cn;
        int prev_dy = ytab[0].di;
        for (int dx = 0; dx < wn; dx++) sum[dx] = 0.f;
        for (int j = 0; j < yn; j++) {
            float beta = ytab[j].alpha;
            int dy = ytab[j].di, sy = ytab[j].si;
            const uint8_t* S = src + (size_t)sy * src_pitch;
            for (int dx = 0; dx < wn; dx++) buf[dx] = 0.f;
            for (int k = 0; k < xn; k++) {
                int sxn = xtab[k].si, dxn = xtab[k].di;
                float alpha = xtab[k].alpha;
                /* separate multiply and add, no FMA contraction */
                volatile float m0 = S[sxn] * alpha, m1 = S[sxn + 1] * alpha, m2 = S[sxn + 2] * alpha;
                buf[dxn] = buf[dxn] + m0; buf[dxn + 1] = buf[dxn + 1] + m1; buf[dxn + 2] = buf[dxn + 2] + m2;
            }
            if (dy != prev_dy) {
                uint8_t* D = dst + (size_t)prev_dy * wn;
                for (int dx = 0; dx < wn; dx++) {
                    D[dx] = sat_u8_from_float(sum[dx]);
                    volatile float m = beta * buf[dx];
                    sum[dx] = m;
                }
                prev_dy = dy;
            } else {
                for (int dx = 0; dx < wn; dx++) {
                    volatile float m = beta * buf[dx];
                    sum[dx] = sum[dx] + m;
                }
            }
        }
        uint8_t* D = dst + (size_t)prev_dy * wn;
        for (int dx = 0; dx < wn; dx++) D[dx] = sat_u8_from_float(sum[dx]);
        free(xtab); free(ytab); free(buf);
        return PA_OK;
    }

    /* upscale on at least one axis: INTER_LINEAR machinery with area-mode coefficients,
       11-bit fixed-point coefficients (HResizeLinear / VResizeLinear for uchar). */
    int* xofs = (int*)malloc(sizeof(int) * (size_t)ow);
    short* ialpha = (short*)malloc(sizeof(short) * 2 * (size_t)ow);
    int* yofs = (int*)malloc(sizeof(int) * (size_t)oh);
    short* ibeta = (short*)malloc(sizeof(short) * 2 * (size_t)oh);
    int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)ow * cn);
    if (!xofs || !ialpha || !yofs || !ibeta || !rows) { free(xofs); free(ialpha); free(yofs); free(ibeta); free(rows); return PA_ERR_ALLOC; }
    for (int dx = 0; dx < ow; dx++) {
        int sx = (int)floor(dx * scale_x);
        float fx = (float)((dx + 1) - (sx + 1) * inv_scale_x);
        fx = fx <= 0 ? 0.f : fx - floorf(fx);
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= w - 1) { fx = 0; sx = w - 1; }
        xofs[dx] = sx;
        ialpha[dx * 2] = sat_s16_from_float((1.f - fx) * 2048);
        ialpha[dx * 2 + 1] = sat_s16_from_float(fx * 2048);
    }
    for (int dy = 0; dy < oh; dy++) {
        int sy = (int)floor(dy * scale_y);
        float fy = (float)((dy + 1) - (sy + 1) * inv_scale_y);
        fy = fy <= 0 ? 0.f : fy - floorf(fy);
        yofs[dy] = sy;
        ibeta[dy * 2] = sat_s16_from_float((1.f - fy) * 2048);
        ibeta[dy * 2 + 1] = sat_s16_from_float(fy * 2048);
    }
    for (int dy = 0; dy < oh; dy++) {
        for (int k = 0; k < 2; k++) {
            int sy = yofs[dy] + k;
            sy = sy >= 0 ? (sy < h ? sy : h - 1) : 0;
            const uint8_t* S = src + (size_t)sy * src_pitch;
            int32_t* R = rows + (size_t)k * ow * cn;
            for (int dx = 0; dx < ow; dx++) {
                int sx = xofs[dx];
                int sx1 = sx + 1 < w ? sx + 1 : sx; /* beyond xmax the library uses S[sx]*ONE (a1 == 0 there) */
                for (int c = 0; c < cn; c++)
                    R[dx * cn + c] = S[sx * cn + c] * ialpha[dx * 2] + S[sx1 * cn + c] * ialpha[dx * 2 + 1];
            }
        }
        int b0 = ibeta[dy * 2], b1 = ibeta[dy * 2 + 1];
        const int32_t *S0 = rows, *S1 = rows + (size_t)ow * cn;
        for (int x = 0; x < ow * cn; x++) {
            int v = (((b0 * (S0[x] >> 4)) >> 16) + ((b1 * (S1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)dy * ow * cn + x] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    free(xofs); free(ialpha); free(yofs); free(ibeta); free(rows);
    return PA_OK;
}

/* ------------------------------------------------- composite: square_crop */

/*
 * YoloCrop.square_crop (fighter.py:323-381) given the already-truncated pixel
 * box (cx, cy, cw, ch) = yolo_pixels(W, H) (fighter.py:305-314).
 * Returns PA_OK (crop written), PA_INVALID (reference returns (False, None)),
 * PA_ERR_ZERO_DIV (reference lets ZeroDivisionError escape: square_dim == 0 with a
 * non-empty window).
 */
int pa_oracle_square_crop(const uint8_t* image, int H, int W, int pitch, int cx, int cy, int cw, int ch,
                          int out_size, int padding, uint8_t* out) {
    int sd = cw > ch ? cw : ch;
    int half = sd / 2; /* int(sd / 2) for sd >= 0 */
    int y0 = cy - half - padding; if (y0 < 0) y0 = 0;
    int y1 = cy + half + padding; if (y1 > H) y1 = H;
    int x0 = cx - half - padding; if (x0 < 0) x0 = 0;
    int x1 = cx + half + padding; if (x1 > W) x1 = W;
    /* numpy slice semantics for a negative stop (box entirely above / left of the frame) */
    if (y1 < 0) { y1 += H; if (y1 < 0) y1 = 0; }
    if (x1 < 0) { x1 += W; if (x1 < 0) x1 = 0; }
    if (y0 > H) y0 = H;
    if (x0 > W) x0 = W;
    int rh = y1 - y0, rw = x1 - x0;
    if (rh < 0) rh = 0;
    if (rw < 0) rw = 0;

    const uint8_t* raw = image + (size_t)y0 * pitch + (size_t)x0 * 3;
    int raw_pitch = pitch;
    uint8_t* sq = NULL;
    if (rh != sd || rw != sd) {
        if (rh == 0 || sd == 0) return PA_ERR_ZERO_DIV; /* contain(): width / height, size[0] / size[1] */
        sq = (uint8_t*)malloc((size_t)sd * sd * 3);
        if (!sq) return PA_ERR_ALLOC;
        int rc = pa_oracle_pil_pad(raw, rh, rw, pitch, sq, sd, sd);
        if (rc != PA_OK) { free(sq); return rc; }
        raw = sq; raw_pitch = sd * 3;
    }
    if (sd == 0) { free(sq); return PA_INVALID; } /* "Bad crop" branch */

    int oh = (int)(sd * ((double)out_size / (double)sd)); /* imutils: int(h * (width / float(w))) */
    if (oh <= 0) { free(sq); return PA_INVALID; }
    if (oh == out_size) {
        int rc = pa_oracle_cv_area(raw, sd, sd, raw_pitch, out, oh, out_size);
        free(sq);
        return rc;
    }
    uint8_t* small = (uint8_t*)malloc((size_t)oh * out_size * 3);
    if (!small) { free(sq); return PA_ERR_ALLOC; }
    int rc = pa_oracle_cv_area(raw, sd, sd, raw_pitch, small, oh, out_size);
    free(sq);
    if (rc != PA_OK) { free(small); return rc; }
    rc = pa_oracle_pil_pad(small, oh, out_size, out_size * 3, out, out_size, out_size);
    free(small);
    return rc;
}
