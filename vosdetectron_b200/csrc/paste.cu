// Mask paste-back for sm_100a: one streaming pass that writes every byte of the
// (R, im_h, im_w) uint8 output exactly once with 128-bit stores.
// Reference: segm_results, lib/core/test.py:801-855 (copy lib_vos/tools/vos_test.py:867-921);
// expand_boxes, lib/utils/boxes.py:242-258; cv2.resize(INTER_LINEAR) on CV_32FC1
// (OpenCV imgproc/resize.cpp: half-pixel centres computed in double, x taps zeroed and clamped
// at the border, y rows replicate-clamped, horizontal then vertical pass in fp32, exact 2x
// shrink handled as INTER_AREA).
#include "common.cuh"

namespace vosd {

struct DetGeom {
    int x0, y0, x1, y1;      // int32-truncated expanded box (inclusive)
    int w, h;                // resized mask extent, >= 1
    double sx, sy;           // cv2 scale_x, scale_y
    int area2x;              // exact 2x shrink -> INTER_AREA
};

__device__ __forceinline__ DetGeom det_geometry(const float* __restrict__ b, int M) {
    // expand_boxes (boxes.py:242-258) in fp32, then .astype(np.int32) truncation (test.py:812-813)
    const float scale = (float)(((double)M + 2.0) / (double)M);
    float w_half = __fmul_rn(__fsub_rn(b[2], b[0]), .5f);
    float h_half = __fmul_rn(__fsub_rn(b[3], b[1]), .5f);
    const float xc = __fmul_rn(__fadd_rn(b[2], b[0]), .5f);
    const float yc = __fmul_rn(__fadd_rn(b[3], b[1]), .5f);
    w_half = __fmul_rn(w_half, scale);
    h_half = __fmul_rn(h_half, scale);
    DetGeom g;
    g.x0 = (int)__fsub_rn(xc, w_half);
    g.x1 = (int)__fadd_rn(xc, w_half);
    g.y0 = (int)__fsub_rn(yc, h_half);
    g.y1 = (int)__fadd_rn(yc, h_half);
    g.w = max(g.x1 - g.x0 + 1, 1);
    g.h = max(g.y1 - g.y0 + 1, 1);
    return g;
}

// cv2: inv_scale = dsize / ssize (double); scale = 1. / inv_scale.  Only needed for chunks that
// actually intersect the box, so it is kept out of det_geometry.
__device__ __forceinline__ void det_scales(DetGeom& g, int M) {
    const int S = M + 2;
    g.sx = __drcp_rn(__ddiv_rn((double)g.w, (double)S));
    g.sy = __drcp_rn(__ddiv_rn((double)g.h, (double)S));
    g.area2x = (2 * g.w == S) && (2 * g.h == S);
}

// zero-padded (M+2)x(M+2) source (test.py:820-823)
__device__ __forceinline__ float padded_at(const float* __restrict__ m, int M, int y, int x) {
    return (y >= 1 && y <= M && x >= 1 && x <= M) ? __ldg(m + (y - 1) * M + (x - 1)) : 0.f;
}

__device__ __forceinline__ float resized_at(const float* __restrict__ m, int M, const DetGeom& g, int dy, int dx) {
    const int S = M + 2;
    if (g.area2x) {
        const float a = padded_at(m, M, 2 * dy, 2 * dx), b = padded_at(m, M, 2 * dy, 2 * dx + 1);
        const float c = padded_at(m, M, 2 * dy + 1, 2 * dx), d = padded_at(m, M, 2 * dy + 1, 2 * dx + 1);
        return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
    }
    float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, g.sx), 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= S - 1) { fx = 0.f; sx = S - 1; }
    float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, g.sy), 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int y0 = min(max(sy, 0), S - 1), y1 = min(max(sy + 1, 0), S - 1);
    const int sx1 = min(sx + 1, S - 1);
    const float a1 = fx, a0 = __fsub_rn(1.f, fx);
    const float b1 = fy, b0 = __fsub_rn(1.f, fy);
    const float r0 = __fadd_rn(__fmul_rn(padded_at(m, M, y0, sx), a0), __fmul_rn(padded_at(m, M, y0, sx1), a1));
    const float r1 = __fadd_rn(__fmul_rn(padded_at(m, M, y1, sx), a0), __fmul_rn(padded_at(m, M, y1, sx1), a1));
    return __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
}

// 16 consecutive output bytes starting at pixel (y, x) of detection r; runs are split at row ends
// (and detection ends when the flat variant walks across frames).
template <bool kProb>
__device__ __forceinline__ void paste_chunk(const float* __restrict__ masks, const int* __restrict__ cls,
                                            const float* __restrict__ ref_boxes, int r, int y, int x, int nbytes,
                                            int K, int M, int im_h, int im_w, float thresh,
                                            uint32_t (&packed)[4], float (&pv)[16]) {
    int done = 0;
    while (done < nbytes) {
        const int run = min(nbytes - done, im_w - x);
        DetGeom g = det_geometry(ref_boxes + 4 * (size_t)r, M);
        const int xa = max(max(g.x0, 0), x), xb = min(min(g.x1 + 1, im_w), x + run);
        if (y >= max(g.y0, 0) && y < min(g.y1 + 1, im_h) && xa < xb) {
            det_scales(g, M);
            const int c = cls ? cls[r] : 0;
            const float* m = masks + ((size_t)r * K + c) * M * M;
            for (int xx = xa; xx < xb; xx++) {
                const float v = resized_at(m, M, g, y - g.y0, xx - g.x0);
                const int i = done + (xx - x);
                if (v > thresh) packed[i >> 2] |= 1u << (8 * (i & 3));
                if (kProb) pv[i] = v;
            }
        }
        done += run;
        x += run;
        if (x >= im_w) { x = 0; if (++y >= im_h) { y = 0; ++r; } }
    }
}

template <bool kProb>
__device__ __forceinline__ void store_chunk(uint8_t* __restrict__ out, float* __restrict__ out_prob, long long f0,
                                            int nbytes, const uint32_t (&packed)[4], const float (&pv)[16]) {
    if (nbytes == 16) {
        st_stream_u4(out + f0, make_uint4(packed[0], packed[1], packed[2], packed[3]));
        if (kProb) {
#pragma unroll
            for (int q = 0; q < 4; q++)
                st_stream_f4(out_prob + f0 + 4 * q, make_float4(pv[4 * q], pv[4 * q + 1], pv[4 * q + 2], pv[4 * q + 3]));
        }
    } else {
        for (int i = 0; i < nbytes; i++) {
            out[f0 + i] = (uint8_t)((packed[i >> 2] >> (8 * (i & 3))) & 0xffu);
            if (kProb) out_prob[f0 + i] = pv[i];
        }
    }
}

// Fast variant (im_h*im_w % 16 == 0): grid = (chunk ranges, detections); a CTA owns a contiguous range of
// 16-byte chunks of ONE detection's frame.  If the range misses the box it is a pure zero fill.  Otherwise the
// CTA first builds, in shared memory, the zero-padded (M+2)^2 source mask and cv2's per-column / per-row
// interpolation tables (source index pair + the two fp32 coefficients, with cv2's border rules) for the
// visible part of the box, so the per-pixel work is 2 table loads, 4 mask loads and 6 flops -- no double
// precision and no bounds checks in the pixel loop.
struct __align__(16) AxisCoef { int i0, i1; float c0, c1; };      // value = S[i0]*c0 + S[i1]*c1
// Column-table slot of entry j: one pad entry per 16, so that the lanes of a warp (16 entries apart: one
// 16-pixel chunk each) read 128-bit entries from different bank groups.
__device__ __forceinline__ int xpad(int j) { return j + (j >> 4); }
constexpr int kPasteHFloats = 6144;               // shared-memory budget of the per-CTA horizontal-pass table

// cv2 column table entry for destination x-offset dx (zeroes the far tap at the border)
__device__ __forceinline__ AxisCoef cv2_x_coef(int dx, double scale, int S) {
    float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale), 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= S - 1) { fx = 0.f; sx = S - 1; }
    return AxisCoef{sx, min(sx + 1, S - 1), __fsub_rn(1.f, fx), fx};
}
// cv2 row table entry (rows are replicate-clamped, weights kept)
__device__ __forceinline__ AxisCoef cv2_y_coef(int dy, double scale, int S) {
    float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale), 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    return AxisCoef{min(max(sy, 0), S - 1) * S, min(max(sy + 1, 0), S - 1) * S, __fsub_rn(1.f, fy), fy};
}

template <bool kProb>
__global__ void __launch_bounds__(256)
paste_det_kernel(const float* __restrict__ masks, const int* __restrict__ cls,
                 const float* __restrict__ ref_boxes, int K, int M, int im_h, int im_w, int chunks_per_cta, int rows_cap,
                 float thresh, uint8_t* __restrict__ out, float* __restrict__ out_prob, uint8_t* __restrict__ out_packed) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int S = M + 2;
    float* smask = reinterpret_cast<float*>(smem);                                   // [S*S]
    AxisCoef* xt = reinterpret_cast<AxisCoef*>(smem + (((size_t)S * S * 4 + 15) & ~(size_t)15));   // [<= im_w, padded: xpad]
    AxisCoef* yt = xt + im_w + im_w / 16 + 1;                                         // [rows of this CTA]

    const int r = blockIdx.y;
    const unsigned frame = (unsigned)im_h * (unsigned)im_w;
    const unsigned f_begin = blockIdx.x * (unsigned)chunks_per_cta * 16u;
    const unsigned f_end = min(frame, f_begin + (unsigned)chunks_per_cta * 16u);
    DetGeom g = det_geometry(ref_boxes + 4 * (size_t)r, M);
    const int ya = max(g.y0, 0), yb = min(g.y1 + 1, im_h), xa = max(g.x0, 0), xb = min(g.x1 + 1, im_w);
    const int row_first = (int)(f_begin / (unsigned)im_w), row_last = (int)((f_end - 1u) / (unsigned)im_w);
    const int ra = max(ya, row_first), rb = min(yb, row_last + 1);     // box rows this CTA touches
    const bool hit = ra < rb && xa < xb;                              // uniform across the CTA
    float* ht = reinterpret_cast<float*>(yt + rows_cap);              // [source rows of this CTA][hpitch]
    bool use_h = false;
    int hpitch = 0;
    uint8_t* __restrict__ o = out ? out + (size_t)r * frame : nullptr;
    float* __restrict__ op = kProb ? out_prob + (size_t)r * frame : nullptr;
    // optional 1-bit-per-pixel copy (pixel 8j+k -> bit k of byte j, the layout of vosd_pack_mask_bits): the 16
    // pixels of a chunk are two bytes, so a warp writes 64 contiguous bytes -- no second pass over the dense masks
    uint8_t* __restrict__ opk = out_packed ? out_packed + (size_t)r * (frame / 8u) : nullptr;

    if (hit) {
        det_scales(g, M);
        const int c = cls ? cls[r] : 0;
        const float* m = masks + ((size_t)r * K + c) * M * M;
        for (int i = threadIdx.x; i < S * S; i += 256) {
            const int y = i / S, x = i - y * S;
            smask[i] = (y >= 1 && y <= M && x >= 1 && x <= M) ? __ldg(m + (y - 1) * M + (x - 1)) : 0.f;   // test.py:820-823
        }
        for (int x = xa + threadIdx.x; x < xb; x += 256) {
            AxisCoef e = cv2_x_coef(x - g.x0, g.sx, S);
            if (g.area2x) e = AxisCoef{2 * (x - g.x0), 2 * (x - g.x0) + 1, 0.5f, 0.5f};       // exact 2x shrink: INTER_AREA
            xt[xpad(x - xa)] = e;
        }
        for (int y = ra + threadIdx.x; y < rb; y += 256) {
            AxisCoef e = cv2_y_coef(y - g.y0, g.sy, S);
            if (g.area2x) e = AxisCoef{2 * (y - g.y0) * S, (2 * (y - g.y0) + 1) * S, 0.5f, 0.5f};
            yt[y - ra] = e;
        }
        __syncthreads();
        // Separable form: cv2 resizes horizontally, then vertically, and the horizontal result of source row s at
        // column x, h[s][x] = S[s][i0(x)] * c0(x) + S[s][i1(x)] * c1(x), does not depend on the output row.  The
        // rows of this CTA touch only a few source rows, so h is built once per CTA (same two products and one
        // sum as the per-pixel path: identical bits) and a pixel then costs 2 loads and 3 flops instead of 2 table
        // entries, 4 mask loads and 6 flops.  Falls back to the per-pixel path when h does not fit (wide, flat boxes).
        const int s_lo = yt[0].i0 / S, s_hi = yt[rb - ra - 1].i1 / S;       // the y table is monotone
        const int nsr = s_hi - s_lo + 1;
        hpitch = xpad(xb - xa - 1) + 1;
        use_h = nsr * hpitch <= kPasteHFloats;
        if (use_h) {
            for (int e = threadIdx.x; e < nsr * (xb - xa); e += 256) {
                const int k = e / (xb - xa), j = e - k * (xb - xa);
                const AxisCoef ex = xt[xpad(j)];
                const float* srow = smask + (s_lo + k) * S;
                ht[k * hpitch + xpad(j)] = __fadd_rn(__fmul_rn(srow[ex.i0], ex.c0), __fmul_rn(srow[ex.i1], ex.c1));
            }
            __syncthreads();                            // every thread has read the source-row range of the y table
            for (int y = ra + threadIdx.x; y < rb; y += 256) {
                AxisCoef e = yt[y - ra];
                e.i0 = (e.i0 / S - s_lo) * hpitch;      // from now on: offsets of the two h rows
                e.i1 = (e.i1 / S - s_lo) * hpitch;
                yt[y - ra] = e;
            }
            __syncthreads();
        }
    }

    for (unsigned f0 = f_begin + threadIdx.x * 16u; f0 < f_end; f0 += 256u * 16u) {
        unsigned long long lo = 0ull, hi = 0ull;          // bytes 0-7 / 8-15 of the chunk (no local-memory array)
        float pv[16];
        if (kProb) {
#pragma unroll
            for (int i = 0; i < 16; i++) pv[i] = 0.f;
        }
        if (hit) {
            int y = (int)(f0 / (unsigned)im_w), x = (int)(f0 - (unsigned)y * (unsigned)im_w);
            if (x + 16 <= im_w) {
                // the chunk lies in one image row (all but one chunk per row): fully unrolled, static bit positions;
                // pixels outside [xa, xb) evaluate a clamped table entry and are masked out
                if (y >= ra && y < rb && x < xb && x + 16 > xa) {
                    const AxisCoef ey = yt[y - ra];
                    const float* s0 = (use_h ? ht : smask) + ey.i0;
                    const float* s1 = (use_h ? ht : smask) + ey.i1;
                    const int last = xb - xa - 1;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int j = x + i - xa;
                        const int jp = xpad(min(max(j, 0), last));
                        float r0, r1;
                        if (use_h) {
                            r0 = s0[jp]; r1 = s1[jp];
                        } else {
                            const AxisCoef ex = xt[jp];
                            r0 = __fadd_rn(__fmul_rn(s0[ex.i0], ex.c0), __fmul_rn(s0[ex.i1], ex.c1));
                            r1 = __fadd_rn(__fmul_rn(s1[ex.i0], ex.c0), __fmul_rn(s1[ex.i1], ex.c1));
                        }
                        const float v = __fadd_rn(__fmul_rn(r0, ey.c0), __fmul_rn(r1, ey.c1));
                        const bool in = j >= 0 && j <= last;
                        if (in && v > thresh) {
                            if (i < 8) lo |= 1ull << (8 * i); else hi |= 1ull << (8 * (i - 8));
                        }
                        if (kProb) pv[i] = in ? v : 0.f;
                    }
                }
            } else {
                int done = 0;
                while (done < 16) {                                   // runs never leave this detection's frame
                    const int run = min(16 - done, im_w - x);
                    const int x0r = max(xa, x), x1r = min(xb, x + run);
                    if (y >= ra && y < rb && x0r < x1r) {
                        const AxisCoef ey = yt[y - ra];
                        const float* s0 = (use_h ? ht : smask) + ey.i0;
                        const float* s1 = (use_h ? ht : smask) + ey.i1;
                        for (int xx = x0r; xx < x1r; xx++) {
                            float r0, r1;
                            if (use_h) {
                                r0 = s0[xpad(xx - xa)]; r1 = s1[xpad(xx - xa)];
                            } else {
                                const AxisCoef ex = xt[xpad(xx - xa)];
                                r0 = __fadd_rn(__fmul_rn(s0[ex.i0], ex.c0), __fmul_rn(s0[ex.i1], ex.c1));
                                r1 = __fadd_rn(__fmul_rn(s1[ex.i0], ex.c0), __fmul_rn(s1[ex.i1], ex.c1));
                            }
                            const float v = __fadd_rn(__fmul_rn(r0, ey.c0), __fmul_rn(r1, ey.c1));
                            const int i = done + (xx - x);
                            if (v > thresh) {
                                const unsigned long long bit = 1ull << (8 * (i & 7));
                                if (i < 8) lo |= bit; else hi |= bit;
                            }
                            if (kProb) pv[i] = v;
                        }
                    }
                    done += run;
                    x += run;
                    if (x >= im_w) { x = 0; ++y; }
                }
            }
        }
        const uint32_t packed[4] = {(uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)};
        if (o || kProb) {
            if (o) store_chunk<kProb>(o, op, (long long)f0, 16, packed, pv);
            else if (kProb) {
#pragma unroll
                for (int q = 0; q < 4; q++)
                    st_stream_f4(op + f0 + 4 * q, make_float4(pv[4 * q], pv[4 * q + 1], pv[4 * q + 2], pv[4 * q + 3]));
            }
        }
        if (opk) {
            // bytes {0,1} -> bits: gather bit 0 of each byte (multiply trick, as pack_bits_kernel)
            const unsigned b0 = (unsigned)((lo * 0x0102040810204080ull) >> 56);
            const unsigned b1 = (unsigned)((hi * 0x0102040810204080ull) >> 56);
            *reinterpret_cast<uint16_t*>(opk + (f0 >> 3)) = (uint16_t)(b0 | (b1 << 8));
        }
    }
}

// Generic flat variant: each thread produces 16 consecutive bytes of the (R*im_h*im_w) output.
template <bool kProb>
__global__ void __launch_bounds__(256)
paste_kernel(const float* __restrict__ masks, const int* __restrict__ cls,
             const float* __restrict__ ref_boxes, long long total, int K, int M, int im_h, int im_w,
             float thresh, uint8_t* __restrict__ out, float* __restrict__ out_prob) {
    const long long frame = (long long)im_h * im_w;
    const long long chunks = (total + 15) / 16;
    for (long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x; ch < chunks;
         ch += (long long)gridDim.x * blockDim.x) {
        const long long f0 = ch * 16;
        const int r = (int)(f0 / frame);
        const int rem = (int)(f0 - (long long)r * frame);
        const int y = rem / im_w, x = rem - y * im_w;
        const int nbytes = (int)min(16LL, total - f0);
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
        float pv[16];
        if (kProb) {
#pragma unroll
            for (int i = 0; i < 16; i++) pv[i] = 0.f;
        }
        paste_chunk<kProb>(masks, cls, ref_boxes, r, y, x, nbytes, K, M, im_h, im_w, thresh, packed, pv);
        store_chunk<kProb>(out, out_prob, f0, nbytes, packed, pv);
    }
}

// Dense {0,1} uint8 masks -> 1 bit per pixel, LSB first (pixel 8j+k -> bit k of byte j); the payload of the
// final all-gather.  Each thread turns 64 mask bytes into 8 packed bytes: 4 x 128-bit loads, one 64-bit store.
__global__ void __launch_bounds__(256)
pack_bits_kernel(const uint8_t* __restrict__ in, long long n_in_per_mask, long long n_out_per_mask,
                 uint8_t* __restrict__ out, int vec_ok) {
    const long long m = blockIdx.y;
    const uint8_t* src = in + m * n_in_per_mask;
    uint8_t* dst = out + m * n_out_per_mask;
    const long long groups = (n_out_per_mask + 7) / 8;              // 8 output bytes per thread
    for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
        const long long i0 = gi * 64;
        unsigned long long packed = 0;
        if (vec_ok && i0 + 64 <= n_in_per_mask) {
            const uint4* p4 = reinterpret_cast<const uint4*>(src + i0);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 v = __ldg(p4 + q);
                const unsigned long long lo = ((unsigned long long)v.y << 32) | v.x, hi = ((unsigned long long)v.w << 32) | v.z;
                packed |= ((lo & 0x0101010101010101ull) * 0x0102040810204080ull >> 56) << (16 * q);
                packed |= ((hi & 0x0101010101010101ull) * 0x0102040810204080ull >> 56) << (16 * q + 8);
            }
            *reinterpret_cast<unsigned long long*>(dst + gi * 8) = packed;
        } else {
            for (int b = 0; b < 8 && gi * 8 + b < n_out_per_mask; b++) {
                unsigned v = 0;
                for (int k = 0; k < 8; k++) {
                    const long long i = i0 + b * 8 + k;
                    if (i < n_in_per_mask && src[i]) v |= 1u << k;
                }
                dst[gi * 8 + b] = (uint8_t)v;
            }
        }
    }
}


// =======================================================================================================
// Fused paste -> COCO RLE: the last step of segm_results (lib/core/test.py:843-848,
// `rle = mask_util.encode(np.array(im_mask[:, :, np.newaxis], order='F'))[0]`) without ever materialising the
// dense (im_h, im_w) canvas.  pycocotools is a third-party dependency that is not vendored by the reference
// (unpinned, README.md:63-74); the algorithm restated here is its published common/maskApi.c: rleEncode (runs
// of the column-major pixel sequence, starting with a run of zeros) and rleToString (per run the difference to
// the run two back for i > 2, as 5-bit groups with a continuation bit, sign-extended, + 48).
//
// One CTA per detection.  Only the box can hold ones, so the CTA evaluates just the box's pixels (same tables
// and arithmetic as paste_det_kernel: identical bits), one warp per image column, lanes along y, and keeps the
// result as 32-row bit words in shared memory (column-major, like the RLE order).  A transition is a pixel that
// differs from its predecessor in column-major order; t = w ^ ((w << 1) | carry) marks the transitions of a word.
// Transitions are counted per thread segment, a block scan gives every thread its first run index and the last
// transition before its segment, the detection reserves T + 1 runs in the caller's arena with one atomicAdd, and
// the runs are written in order.  The string pass repeats the pattern over the runs.
// =======================================================================================================
__device__ __forceinline__ int block_scan_excl(int v, int* warp_buf, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();                               // warp_buf free (previous use consumed)
    if (lane == 31) warp_buf[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int x = warp_buf[w];
        if (w < warp) base += x;
        tot += x;
    }
    total = tot;
    return base + inc - v;
}
// exclusive running maximum (identity 0) over the threads of the block
__device__ __forceinline__ int block_scan_excl_max(int v, int* warp_buf, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, t);
    }
    int exc = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) exc = 0;
    __syncthreads();
    if (lane == 31) warp_buf[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int x = warp_buf[w];
        if (w < warp) base = max(base, x);
        tot = max(tot, x);
    }
    total = tot;
    return max(base, exc);
}

// chars rleToString emits for run value x (maskApi.c: `long x`, arithmetic shifts)
__device__ __forceinline__ int rle_chars(long long x, uint8_t* dst) {
    int n = 0;
    bool more = true;
    while (more) {
        int c = (int)(x & 0x1f);
        x >>= 5;
        more = (c & 0x10) ? x != -1 : x != 0;
        if (more) c |= 0x20;
        if (dst) dst[n] = (uint8_t)(c + 48);
        n++;
    }
    return n;
}

__global__ void __launch_bounds__(256)
paste_rle_kernel(const float* __restrict__ masks, const int* __restrict__ cls, const float* __restrict__ ref_boxes,
                 int K, int M, int im_h, int im_w, float thresh,
                 uint32_t* __restrict__ run_arena, long long run_cap, uint8_t* __restrict__ str_arena, long long str_cap,
                 unsigned long long* __restrict__ cursors, long long* __restrict__ run_offset, int* __restrict__ run_count,
                 long long* __restrict__ str_offset, int* __restrict__ str_len, int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int warp_buf[8];
    __shared__ long long s_base[2];
    const int S = M + 2;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = im_h * im_w;
    DetGeom g = det_geometry(ref_boxes + 4 * (size_t)r, M);
    const int ya = max(g.y0, 0), yb = min(g.y1 + 1, im_h), xa = max(g.x0, 0), xb = min(g.x1 + 1, im_w);
    const bool hit = ya < yb && xa < xb;
    // bit matrix: columns [xa, xb2), rows [ya2, yb2): one zero column / row beyond the box so that the 1 -> 0
    // transition behind it is seen; a box that touches the bottom edge is extended to the top edge, because the
    // pixel after (x, H-1) is (x+1, 0)
    const int ya2 = hit ? (yb >= im_h ? 0 : ya) : 0, yb2 = hit ? min(yb + 1, im_h) : 0;
    const int xb2 = hit ? min(xb + 1, im_w) : xa;
    const int cols = hit ? xb2 - xa : 0, rows = yb2 - ya2, words = (rows + 31) >> 5;
    const bool wraps = ya2 == 0 && yb2 == im_h;        // a column's predecessor is the last pixel of the column before

    float* smask = reinterpret_cast<float*>(smem);
    AxisCoef* xt = reinterpret_cast<AxisCoef*>(smem + (((size_t)S * S * 4 + 15) & ~(size_t)15));     // [im_w]
    AxisCoef* yt = xt + im_w;                                                                        // [im_h]
    uint32_t* bits = reinterpret_cast<uint32_t*>(yt + im_h);                                         // [cols * words]

    if (hit) {
        det_scales(g, M);
        const int c = cls ? cls[r] : 0;
        const float* m = masks + ((size_t)r * K + c) * M * M;
        for (int i = tid; i < S * S; i += 256) {
            const int y = i / S, x = i - y * S;
            smask[i] = (y >= 1 && y <= M && x >= 1 && x <= M) ? __ldg(m + (y - 1) * M + (x - 1)) : 0.f;
        }
        for (int x = xa + tid; x < xb; x += 256) {
            AxisCoef e = cv2_x_coef(x - g.x0, g.sx, S);
            if (g.area2x) e = AxisCoef{2 * (x - g.x0), 2 * (x - g.x0) + 1, 0.5f, 0.5f};
            xt[x - xa] = e;
        }
        for (int y = ya + tid; y < yb; y += 256) {
            AxisCoef e = cv2_y_coef(y - g.y0, g.sy, S);
            if (g.area2x) e = AxisCoef{2 * (y - g.y0) * S, (2 * (y - g.y0) + 1) * S, 0.5f, 0.5f};
            yt[y - ya] = e;
        }
        __syncthreads();
        // ---- bits: warp per column, lanes along y ----
        for (int j = warp; j < cols; j += 8) {
            const int x = xa + j;
            const bool colin = x < xb;
            AxisCoef ex = AxisCoef{0, 0, 0.f, 0.f};
            if (colin) ex = xt[j];
            for (int wd = 0; wd < words; wd++) {
                const int y = ya2 + 32 * wd + lane;
                bool bit = false;
                if (colin && y >= ya && y < yb) {
                    const AxisCoef ey = yt[y - ya];
                    const float* s0 = smask + ey.i0;
                    const float* s1 = smask + ey.i1;
                    const float r0 = __fadd_rn(__fmul_rn(s0[ex.i0], ex.c0), __fmul_rn(s0[ex.i1], ex.c1));
                    const float r1 = __fadd_rn(__fmul_rn(s1[ex.i0], ex.c0), __fmul_rn(s1[ex.i1], ex.c1));
                    bit = __fadd_rn(__fmul_rn(r0, ey.c0), __fmul_rn(r1, ey.c1)) > thresh;
                }
                const unsigned w = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) bits[j * words + wd] = w;
            }
        }
        __syncthreads();
    }

    // ---- transitions of this thread's segment of (column, word) entries ----
    const int n = cols * words;
    const int per = (n + 255) / 256;
    const int e0 = min(n, tid * per), e1 = min(n, e0 + per);
    const unsigned last_valid = (rows & 31) ? ((1u << (rows & 31)) - 1u) : 0xffffffffu;
    auto trans = [&](int e, int& pos0) -> unsigned {
        const int j = e / words, wd = e - j * words;
        const unsigned w = bits[e];
        unsigned prev;
        if (wd > 0) prev = bits[e - 1] >> 31;
        else if (wraps && j > 0) prev = (bits[e - 1] >> ((rows - 1) & 31)) & 1u;
        else prev = 0u;
        unsigned t = w ^ ((w << 1) | prev);
        if (wd == words - 1) t &= last_valid;
        pos0 = (xa + j) * im_h + ya2 + 32 * wd;
        return t;
    };
    int cnt = 0, last = 0;
    for (int e = e0; e < e1; e++) {
        int pos0;
        const unsigned t = trans(e, pos0);
        if (t) { cnt += __popc(t); last = pos0 + 31 - __clz(t); }
    }
    int T, last_all;
    const int k0 = block_scan_excl(cnt, warp_buf, T);
    int prev_pos = block_scan_excl_max(last, warp_buf, last_all);
    const int nruns = T + 1;
    if (tid == 0) {
        const long long base = (long long)atomicAdd(cursors, (unsigned long long)nruns);
        s_base[0] = base;
        run_offset[r] = base;
        run_count[r] = nruns;
    }
    __syncthreads();
    const long long rbase = s_base[0];
    const bool run_ok = rbase + nruns <= run_cap;
    uint32_t* __restrict__ runs = run_arena + rbase;
    if (run_ok) {
        int k = k0;
        for (int e = e0; e < e1; e++) {
            int pos0;
            unsigned t = trans(e, pos0);
            while (t) {
                const int b = __ffs(t) - 1;
                t &= t - 1;
                const int pos = pos0 + b;
                runs[k++] = (uint32_t)(pos - prev_pos);
                prev_pos = pos;
            }
        }
        if (tid == 0) runs[T] = (uint32_t)(frame - last_all);
    }
    __syncthreads();                                    // the runs of this detection are visible to the whole CTA

    // ---- rleToString over the runs ----
    const int per2 = (nruns + 255) / 256;
    const int i0 = min(nruns, tid * per2), i1 = min(nruns, i0 + per2);
    int nch = 0;
    if (run_ok)
        for (int i = i0; i < i1; i++) {
            long long x = (long long)runs[i];
            if (i > 2) x -= (long long)runs[i - 2];
            nch += rle_chars(x, nullptr);
        }
    int L;
    const int c0 = block_scan_excl(nch, warp_buf, L);
    if (tid == 0) {
        const long long base = run_ok ? (long long)atomicAdd(cursors + 1, (unsigned long long)L) : 0;
        s_base[1] = base;
        str_offset[r] = base;
        str_len[r] = run_ok ? L : 0;
    }
    __syncthreads();
    const long long sbase = s_base[1];
    const bool str_ok = run_ok && sbase + L <= str_cap;
    if (str_ok) {
        uint8_t* __restrict__ dst = str_arena + sbase + c0;
        for (int i = i0; i < i1; i++) {
            long long x = (long long)runs[i];
            if (i > 2) x -= (long long)runs[i - 2];
            dst += rle_chars(x, dst);
        }
    }
    if (tid == 0) status[r] = run_ok ? (str_ok ? 0 : 2) : 1;
}

}  // namespace vosd

using namespace vosd;

extern "C" int vosd_pack_mask_bits(const uint8_t* masks, int num_masks, long long pixels_per_mask,
                                   uint8_t* packed, cudaStream_t stream) {
    if (num_masks < 0 || pixels_per_mask < 1) return VOSD_ERR_BAD_SHAPE;
    if (num_masks == 0) return VOSD_OK;
    if (!masks || !packed) return VOSD_ERR_BAD_ARG;
    if (num_masks > 65535) return VOSD_ERR_UNSUPPORTED;
    const long long n_out = (pixels_per_mask + 7) / 8;
    const int vec_ok = aligned16(masks) && (reinterpret_cast<uintptr_t>(packed) & 7) == 0 &&
                       pixels_per_mask % 16 == 0 && n_out % 8 == 0;
    long long bx = ((n_out + 7) / 8 + 255) / 256;
    if (bx > 4096) bx = 4096;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)num_masks);
    pack_bits_kernel<<<grid, 256, 0, stream>>>(masks, pixels_per_mask, n_out, packed, vec_ok);
    count_launch();
    return check_launch();
}


static int paste_impl(const float* masks, const int* cls, const float* ref_boxes,
                      int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                      float thresh, uint8_t* out, float* out_prob, uint8_t* out_packed, bool want_packed,
                      cudaStream_t stream) {
    if (num_dets < 0 || num_classes < 1 || mask_size < 1 || im_h < 1 || im_w < 1) return VOSD_ERR_BAD_SHAPE;
    if (num_dets == 0) return VOSD_OK;
    if (!masks || !ref_boxes || (want_packed ? !out_packed : !out)) return VOSD_ERR_BAD_ARG;
    if ((out && !aligned16(out)) || (out_prob && !aligned16(out_prob))) return VOSD_ERR_BAD_ARG;
    if (out_packed && (reinterpret_cast<uintptr_t>(out_packed) & 1)) return VOSD_ERR_BAD_ARG;
    const long long total = (long long)num_dets * im_h * im_w;
    const long long chunks = (total + 15) / 16;
    long long blocks = (chunks + 255) / 256;
    if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
    const long long frame = (long long)im_h * im_w;
    if (frame % 16 == 0 && frame < (1LL << 31) && num_dets <= 65535) {
        // 2048 chunks (32 KB of output) per CTA: few, fat CTAs; tables cover at most the rows a CTA spans
#ifndef VOSD_PASTE_CHUNKS
#define VOSD_PASTE_CHUNKS 2048
#endif
        const int chunks_per_cta = VOSD_PASTE_CHUNKS;
        const unsigned bx = (unsigned)((frame / 16 + chunks_per_cta - 1) / chunks_per_cta);
        const int rows_per_cta = (chunks_per_cta * 16 + im_w - 1) / im_w + 2;
        const int S = mask_size + 2;
        const size_t smem = (((size_t)S * S * 4 + 15) & ~(size_t)15) + (size_t)(im_w + im_w / 16 + 1 + rows_per_cta) * 16 +
                            (size_t)kPasteHFloats * 4;
        if (smem > 200 * 1024) return VOSD_ERR_UNSUPPORTED;
        dim3 grid(bx, (unsigned)num_dets);
        if (out_prob) {
            if (cudaFuncSetAttribute(paste_det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            paste_det_kernel<true><<<grid, 256, smem, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                                chunks_per_cta, rows_per_cta, thresh, out, out_prob, out_packed);
        } else {
            if (cudaFuncSetAttribute(paste_det_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            paste_det_kernel<false><<<grid, 256, smem, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                                 chunks_per_cta, rows_per_cta, thresh, out, out_prob, out_packed);
        }
        count_launch();
        return check_launch();
    }
    // odd frame sizes: flat dense kernel, then (if asked) the stand-alone packer over the dense result
    if (!out) return VOSD_ERR_UNSUPPORTED;
    if (out_prob) {
        paste_kernel<true><<<(int)blocks, 256, 0, stream>>>(masks, cls, ref_boxes, total, num_classes, mask_size,
                                                            im_h, im_w, thresh, out, out_prob);
    } else {
        paste_kernel<false><<<(int)blocks, 256, 0, stream>>>(masks, cls, ref_boxes, total, num_classes, mask_size,
                                                             im_h, im_w, thresh, out, out_prob);
    }
    count_launch();
    const int st = check_launch();
    if (st != VOSD_OK || !out_packed) return st;
    return vosd_pack_mask_bits(out, num_dets, frame, out_packed, stream);
}

extern "C" int vosd_paste_masks(const float* masks, const int* cls, const float* ref_boxes,
                                int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                                float thresh, uint8_t* out, float* out_prob, cudaStream_t stream) {
    return paste_impl(masks, cls, ref_boxes, num_dets, num_classes, mask_size, im_h, im_w, thresh, out, out_prob,
                      nullptr, false, stream);
}

extern "C" int vosd_paste_masks_packed(const float* masks, const int* cls, const float* ref_boxes,
                                       int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                                       float thresh, uint8_t* out, uint8_t* out_packed, cudaStream_t stream) {
    return paste_impl(masks, cls, ref_boxes, num_dets, num_classes, mask_size, im_h, im_w, thresh, out, nullptr,
                      out_packed, true, stream);
}

extern "C" size_t vosd_paste_rle_smem_bytes(int mask_size, int im_h, int im_w) {
    const int S = mask_size + 2;
    return (((size_t)S * S * 4 + 15) & ~(size_t)15) + (size_t)(im_w + im_h) * 16 +
           (size_t)im_w * ((im_h + 31) / 32) * 4;
}

extern "C" int vosd_paste_rle(const float* masks, const int* cls, const float* ref_boxes,
                              int num_dets, int num_classes, int mask_size, int im_h, int im_w, float thresh,
                              uint32_t* run_arena, long long run_capacity, uint8_t* str_arena, long long str_capacity,
                              unsigned long long* cursors, long long* run_offset, int* run_count,
                              long long* str_offset, int* str_len, int* status, cudaStream_t stream) {
    if (num_dets < 0 || num_classes < 1 || mask_size < 1 || im_h < 1 || im_w < 1) return VOSD_ERR_BAD_SHAPE;
    if (run_capacity < 0 || str_capacity < 0) return VOSD_ERR_BAD_SHAPE;
    if ((long long)im_h * im_w >= (1LL << 31)) return VOSD_ERR_UNSUPPORTED;
    if (!cursors) return VOSD_ERR_BAD_ARG;
    if (cudaMemsetAsync(cursors, 0, 2 * sizeof(unsigned long long), stream) != cudaSuccess) return VOSD_ERR_LAUNCH;
    if (num_dets == 0) return VOSD_OK;
    if (!masks || !ref_boxes || !run_arena || !str_arena || !run_offset || !run_count || !str_offset || !str_len || !status)
        return VOSD_ERR_BAD_ARG;
    const size_t smem = vosd_paste_rle_smem_bytes(mask_size, im_h, im_w);
    if (smem > 220 * 1024) return VOSD_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(paste_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    paste_rle_kernel<<<num_dets, 256, smem, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w, thresh,
                                                     run_arena, run_capacity, str_arena, str_capacity, cursors,
                                                     run_offset, run_count, str_offset, str_len, status);
    count_launch();
    return check_launch();
}
