/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the per-frame region pipeline.
 *
 * Plain-C restatement of the native parts of the reference hot path
 * (YeLyuUT/VOSDetectron).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * package (vosdetectron_b200/) never does.
 *
 * Parity pinning (see DESIGN.md "Oracle"): the reference ships no tests or
 * golden vectors for this path, so every function here
 * is pinned against the reference code itself executed in the build
 * container (tests/test_oracle_vs_reference.py, fixtures in tests/golden/).
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off [-mfma] oracle.c -lm
 *   -ffp-contract=off : no implicit fused multiply-add -- the Cython NMS of the
 *                       reference is built for baseline x86-64 (no FMA), and the
 *                       RoIAlign restatement places fmaf() exactly where nvcc
 *                       contracted the reference kernel (read from its sm_100a SASS).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* Greedy NMS -- lib/utils/cython_nms.pyx:37-87                               */
/*   areas = (x2-x1+1)*(y2-y1+1) in fp32 (:43), order = argsort(scores)[::-1] */
/*   (:45), suppress j when inter/(iarea+area_j-inter) >= thresh (:76-85),    */
/*   return np.where(suppressed==0)[0] (ascending index, :87).                */
/* Tie order of the reference's unstable argsort is unspecified; the oracle   */
/* (and the CUDA path) fix it as: higher score first, lower index first.      */
/* ------------------------------------------------------------------------- */
typedef struct { float s; int64_t i; } orc_key;
static int orc_key_cmp(const void* a, const void* b) {
    const orc_key* x = (const orc_key*)a; const orc_key* y = (const orc_key*)b;
    if (x->s > y->s) return -1;
    if (x->s < y->s) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

int64_t orc_nms(const float* dets, int64_t n, float thresh, int64_t* keep) {
    if (n <= 0) return 0;
    float* areas = (float*)malloc(sizeof(float) * n);
    orc_key* order = (orc_key*)malloc(sizeof(orc_key) * n);
    uint8_t* sup = (uint8_t*)calloc(n, 1);
    for (int64_t i = 0; i < n; i++) {
        const float* d = dets + 5 * i;
        areas[i] = ((d[2] - d[0]) + 1.0f) * ((d[3] - d[1]) + 1.0f);
        order[i].s = d[4]; order[i].i = i;
    }
    qsort(order, n, sizeof(orc_key), orc_key_cmp);
    for (int64_t _i = 0; _i < n; _i++) {
        int64_t i = order[_i].i;
        if (sup[i]) continue;
        const float ix1 = dets[5*i], iy1 = dets[5*i+1], ix2 = dets[5*i+2], iy2 = dets[5*i+3];
        const float iarea = areas[i];
        for (int64_t _j = _i + 1; _j < n; _j++) {
            int64_t j = order[_j].i;
            if (sup[j]) continue;
            float xx1 = ix1 >= dets[5*j]   ? ix1 : dets[5*j];
            float yy1 = iy1 >= dets[5*j+1] ? iy1 : dets[5*j+1];
            float xx2 = ix2 <= dets[5*j+2] ? ix2 : dets[5*j+2];
            float yy2 = iy2 <= dets[5*j+3] ? iy2 : dets[5*j+3];
            float w = (xx2 - xx1) + 1.0f; w = 0.0f >= w ? 0.0f : w;
            float h = (yy2 - yy1) + 1.0f; h = 0.0f >= h ? 0.0f : h;
            float inter = w * h;
            float ovr = inter / ((iarea + areas[j]) - inter);
            if (ovr >= thresh) sup[j] = 1;
        }
    }
    int64_t c = 0;
    for (int64_t i = 0; i < n; i++) if (!sup[i]) keep[c++] = i;
    free(areas); free(order); free(sup);
    return c;
}

/* ------------------------------------------------------------------------- */
/* RoIAlign -- lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu       */
/*   forward 65-121 (+bilinear_interpolate 16-63), backward 195-270           */
/*   (+bilinear_interpolate_gradient 150-193).                                */
/* fmaf() marks the contractions nvcc 12.9 makes when it compiles that file   */
/* for sm_100a with its default -fmad=true (verified in the SASS of           */
/* oracle/_ref/libref_roialign.so):                                           */
/*   roi_width  = fmaxf(fma(roi[3], scale, -(roi[1]*scale)), 1)               */
/*   y          = fma(ph, bin_h, start_h) + ((iy+.5f)*bin_h)/grid_h           */
/*   val        = fma(w4,v4, fma(w3,v3, fma(w1,v1, w2*v2)))                   */
/* ------------------------------------------------------------------------- */
typedef struct { int low, high; float l, h; int valid; } orc_coord;

static inline orc_coord orc_axis(float v, int size) {
    orc_coord c;
    c.valid = !(v < -1.0f || v > (float)size);
    if (v <= 0) v = 0;
    c.low = (int)v;
    if (c.low >= size - 1) { c.high = c.low = size - 1; v = (float)c.low; }
    else c.high = c.low + 1;
    c.l = v - (float)c.low;
    c.h = 1.0f - c.l;
    return c;
}

typedef struct {
    int b; float start_w, start_h, bin_w, bin_h; int grid_h, grid_w; float count;
} orc_roi;

static inline orc_roi orc_roi_geom(const float* r, float scale, int ph, int pw, int sr) {
    orc_roi g;
    g.b = (int)r[0];
    g.start_w = r[1] * scale;
    g.start_h = r[2] * scale;
    float roi_w = fmaxf(fmaf(r[3], scale, -g.start_w), 1.0f);
    float roi_h = fmaxf(fmaf(r[4], scale, -g.start_h), 1.0f);
    g.bin_h = roi_h / (float)ph;
    g.bin_w = roi_w / (float)pw;
    g.grid_h = sr > 0 ? sr : (int)ceilf(roi_h / (float)ph);
    g.grid_w = sr > 0 ? sr : (int)ceilf(roi_w / (float)pw);
    g.count = (float)(g.grid_h * g.grid_w);
    return g;
}

void orc_roialign_fwd(const float* feat, int N, int C, int H, int W,
                      const float* rois, int R, int PH, int PW,
                      float scale, int sr, float* out, int nthreads) {
    (void)N;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < R; n++) {
        orc_roi g = orc_roi_geom(rois + 5 * n, scale, PH, PW, sr);
        for (int c = 0; c < C; c++) {
            const float* d = feat + ((size_t)g.b * C + c) * H * W;
            float* o = out + ((size_t)n * C + c) * PH * PW;
            for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {
                float acc = 0.f;
                for (int iy = 0; iy < g.grid_h; iy++) {
                    float y = fmaf((float)ph, g.bin_h, g.start_h) + (((float)iy + .5f) * g.bin_h) / (float)g.grid_h;
                    orc_coord cy = orc_axis(y, H);
                    for (int ix = 0; ix < g.grid_w; ix++) {
                        float x = fmaf((float)pw, g.bin_w, g.start_w) + (((float)ix + .5f) * g.bin_w) / (float)g.grid_w;
                        orc_coord cx = orc_axis(x, W);
                        float val = 0.f;
                        if (cy.valid && cx.valid) {
                            float v1 = d[cy.low * W + cx.low], v2 = d[cy.low * W + cx.high];
                            float v3 = d[cy.high * W + cx.low], v4 = d[cy.high * W + cx.high];
                            float w1 = cy.h * cx.h, w2 = cy.h * cx.l, w3 = cy.l * cx.h, w4 = cy.l * cx.l;
                            val = fmaf(w4, v4, fmaf(w3, v3, fmaf(w1, v1, w2 * v2)));
                        }
                        acc += val;
                    }
                }
                o[ph * PW + pw] = acc / g.count;
            }
        }
    }
}

/* bottom_diff must be zero-filled by the caller (functions/roi_align.py:39-40).
 * Parallel over channels so every texel is summed in a fixed (roi, ph, pw, iy, ix)
 * order: deterministic, unlike the atomics of the reference. */
void orc_roialign_bwd(const float* top, int N, int C, int H, int W,
                      const float* rois, int R, int PH, int PW,
                      float scale, int sr, float* bottom, int nthreads) {
    (void)N;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        for (int n = 0; n < R; n++) {
            orc_roi g = orc_roi_geom(rois + 5 * n, scale, PH, PW, sr);
            float* d = bottom + ((size_t)g.b * C + c) * H * W;
            const float* t = top + ((size_t)n * C + c) * PH * PW;
            for (int ph = 0; ph < PH; ph++) for (int pw = 0; pw < PW; pw++) {
                float tv = t[ph * PW + pw];
                for (int iy = 0; iy < g.grid_h; iy++) {
                    float y = fmaf((float)ph, g.bin_h, g.start_h) + (((float)iy + .5f) * g.bin_h) / (float)g.grid_h;
                    orc_coord cy = orc_axis(y, H);
                    for (int ix = 0; ix < g.grid_w; ix++) {
                        float x = fmaf((float)pw, g.bin_w, g.start_w) + (((float)ix + .5f) * g.bin_w) / (float)g.grid_w;
                        orc_coord cx = orc_axis(x, W);
                        if (!(cy.valid && cx.valid)) continue;
                        float w1 = cy.h * cx.h, w2 = cy.h * cx.l, w3 = cy.l * cx.h, w4 = cy.l * cx.l;
                        d[cy.low * W + cx.low]   += (tv * w1) / g.count;
                        d[cy.low * W + cx.high]  += (tv * w2) / g.count;
                        d[cy.high * W + cx.low]  += (tv * w3) / g.count;
                        d[cy.high * W + cx.high] += (tv * w4) / g.count;
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* cv2.resize(src, (dw, dh)) for CV_32FC1, INTER_LINEAR -- the call at        */
/* lib/core/test.py:831.  OpenCV is an unpinned third-party dependency of the */
/* reference (4.13.0 in this image); this restates its published algorithm    */
/* (modules/imgproc/src/resize.cpp): half-pixel centres computed in double    */
/* and rounded to float, x taps zeroed+clamped at the border, y rows          */
/* replicate-clamped with weights kept, horizontal pass then vertical pass in */
/* fp32, and the exact-2x-shrink special case that switches to INTER_AREA.    */
/* Checked against cv2 itself in tests/test_oracle.py (<= 4e-6 abs).          */
/* ------------------------------------------------------------------------- */
static inline int orc_floor(float v) { int i = (int)v; return i - (v < (float)i); }

void orc_resize_linear(const float* src, int sh, int sw, float* dst, int dh, int dw) {
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    if (sw == 2 * dw && sh == 2 * dh) {            /* INTER_LINEAR -> INTER_AREA fast path */
        for (int y = 0; y < dh; y++) for (int x = 0; x < dw; x++) {
            const float* s = src + (2 * y) * sw + 2 * x;
            dst[y * dw + x] = (s[0] + s[1] + s[sw] + s[sw + 1]) * 0.25f;
        }
        return;
    }
    int* xofs = (int*)malloc(sizeof(int) * dw);
    float* xa = (float*)malloc(sizeof(float) * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = orc_floor(fx);
        fx -= (float)sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx; xa[dx] = fx;
    }
    float* r0 = (float*)malloc(sizeof(float) * dw);
    float* r1 = (float*)malloc(sizeof(float) * dw);
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = orc_floor(fy);
        fy -= (float)sy;
        int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
        const float* s0 = src + y0 * sw; const float* s1 = src + y1 * sw;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx]; int sx1 = sx + 1 < sw ? sx + 1 : sx;
            float a1 = xa[dx], a0 = 1.f - a1;
            r0[dx] = s0[sx] * a0 + s0[sx1] * a1;
            r1[dx] = s1[sx] * a0 + s1[sx1] * a1;
        }
        float b1 = fy, b0 = 1.f - fy;
        for (int dx = 0; dx < dw; dx++) dst[dy * dw + dx] = r0[dx] * b0 + r1[dx] * b1;
    }
    free(xofs); free(xa); free(r0); free(r1);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------- */
/* FlowAlign -- lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu    */
/*   forward :15-55, backward :57-117.  The reference mixes float and double  */
/*   operands (`1.` is a double literal, `h_ratio` a float), and nvcc         */
/*   contracts some of the double products into fma: the operation order and  */
/*   the fma() placements below are read from the reference's sm_100a SASS    */
/*   (DMUL/DFMA/DADD/F2F sequence), so on an IEEE machine this produces the   */
/*   reference kernel's bits.  The reference is CUDA-only (functions/         */
/*   flow_align.py:27-30 raises on CPU tensors): its pin is the reference     */
/*   kernel run on the GPU (tests/test_gpu_flowalign.py).                     */
/* ------------------------------------------------------------------------- */
typedef struct { int off; float h, w; } orc_flow_geom;

/* :24-45.  off < 0: sample outside [0,H-1) x [0,W-1) -> output 0, no gradient. */
static orc_flow_geom orc_flow_geometry(const float* flow_n, int H, int W, int h, int w) {
    orc_flow_geom g; g.off = -1; g.h = g.w = 0.f;
    float flo_x = flow_n[h * W + w];
    float flo_y = flow_n[H * W + h * W + w];
    float w_flo = (float)w + flo_x;
    float h_flo = (float)h + flo_y;
    if (h_flo < 0 || h_flo >= (float)(H - 1) || w_flo < 0 || w_flo >= (float)(W - 1)) return g;
    if (h_flo != h_flo || w_flo != w_flo) {          /* NaN passes the tests above; CUDA's float->int gives 0 */
        if (H < 2 || W < 2) return g;
        g.off = 0; g.h = h_flo; g.w = w_flo;          /* NaN - 0.f */
        return g;
    }
    int h_start = (int)floorf(h_flo), w_start = (int)floorf(w_flo);
    g.h = h_flo - (float)h_start;
    g.w = w_flo - (float)w_start;
    g.off = w_start + W * h_start;
    return g;
}

void orc_flow_align_fwd(const float* bottom, const float* flow, int N, int C, int H, int W,
                        float* top, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(static) collapse(2)
    for (int n = 0; n < N; n++) for (int c = 0; c < C; c++) {
        const float* b = bottom + ((size_t)n * C + c) * H * W;
        float* t = top + ((size_t)n * C + c) * H * W;
        for (int h = 0; h < H; h++) for (int w = 0; w < W; w++) {
            orc_flow_geom g = orc_flow_geometry(flow + (size_t)n * 2 * H * W, H, W, h, w);
            if (g.off < 0) { t[h * W + w] = 0.f; continue; }
            const float* p = b + g.off;
            double A = 1.0 - (double)g.h, B = 1.0 - (double)g.w;
            double acc = ((double)p[1] * A) * (double)g.w;                  /* :49 */
            acc = fma((double)p[0] * A, B, acc);                            /* :48 */
            acc = fma(B, (double)(p[W] * g.h), acc);                        /* :50, float product first */
            acc = acc + (double)((p[W + 1] * g.h) * g.w);                   /* :51, all float */
            t[h * W + w] = (float)acc;
        }
    }
}

/* bottomdiff (N,C,H,W) and flowdiff (N,2,H,W) must be zero-filled by the caller
 * (functions/flow_align.py:41-43).  Fixed (c, h, w) summation order: deterministic, unlike the
 * reference's atomics. */
void orc_flow_align_bwd(const float* topdiff, const float* bottom, const float* flow,
                        int N, int C, int H, int W, float* bottomdiff, float* flowdiff, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel for schedule(static)
    for (int n = 0; n < N; n++) {
        float* fdx = flowdiff + (size_t)n * 2 * H * W;
        float* fdy = fdx + H * W;
        for (int c = 0; c < C; c++) {
            const float* b = bottom + ((size_t)n * C + c) * H * W;
            const float* t = topdiff + ((size_t)n * C + c) * H * W;
            float* d = bottomdiff + ((size_t)n * C + c) * H * W;
            for (int h = 0; h < H; h++) for (int w = 0; w < W; w++) {
                orc_flow_geom g = orc_flow_geometry(flow + (size_t)n * 2 * H * W, H, W, h, w);
                if (g.off < 0) continue;
                float tv = t[h * W + w];
                double A = 1.0 - (double)g.h, B = 1.0 - (double)g.w;
                d[g.off]         += (float)(((double)tv * A) * B);           /* :89 */
                d[g.off + 1]     += (float)((A * (double)tv) * (double)g.w); /* :90 */
                d[g.off + W]     += (float)(B * (double)(g.h * tv));         /* :91 */
                d[g.off + W + 1] += g.w * (g.h * tv);                        /* :92 */
                float f1 = b[g.off], f2 = b[g.off + 1], f3 = b[g.off + W], f4 = b[g.off + W + 1];
                double dx = fma(A, (double)(-f1), A * (double)f2);           /* :104 */
                dx = dx - (double)(g.h * f3);
                dx = dx + (double)(g.h * f4);
                double dy = fma(B, (double)(-f1), -(double)(g.w * f2));      /* :106 */
                dy = fma(B, (double)f3, dy);
                dy = dy + (double)(g.w * f4);
                fdx[h * W + w] += tv * (float)dx;                            /* :111 */
                fdy[h * W + w] += tv * (float)dy;                            /* :112 */
            }
        }
    }
}
