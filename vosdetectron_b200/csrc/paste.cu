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
                 const float* __restrict__ ref_boxes, int K, int M, int im_h, int im_w, int chunks_per_cta,
                 float thresh, uint8_t* __restrict__ out, float* __restrict__ out_prob) {
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
    uint8_t* __restrict__ o = out + (size_t)r * frame;
    float* __restrict__ op = kProb ? out_prob + (size_t)r * frame : nullptr;

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
                    const float* s0 = smask + ey.i0;
                    const float* s1 = smask + ey.i1;
                    const int last = xb - xa - 1;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int j = x + i - xa;
                        const AxisCoef ex = xt[xpad(min(max(j, 0), last))];
                        const float r0 = __fadd_rn(__fmul_rn(s0[ex.i0], ex.c0), __fmul_rn(s0[ex.i1], ex.c1));
                        const float r1 = __fadd_rn(__fmul_rn(s1[ex.i0], ex.c0), __fmul_rn(s1[ex.i1], ex.c1));
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
                        const float* s0 = smask + ey.i0;
                        const float* s1 = smask + ey.i1;
                        for (int xx = x0r; xx < x1r; xx++) {
                            const AxisCoef ex = xt[xpad(xx - xa)];
                            const float r0 = __fadd_rn(__fmul_rn(s0[ex.i0], ex.c0), __fmul_rn(s0[ex.i1], ex.c1));
                            const float r1 = __fadd_rn(__fmul_rn(s1[ex.i0], ex.c0), __fmul_rn(s1[ex.i1], ex.c1));
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
        store_chunk<kProb>(o, op, (long long)f0, 16, packed, pv);
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


extern "C" int vosd_paste_masks(const float* masks, const int* cls, const float* ref_boxes,
                                int num_dets, int num_classes, int mask_size, int im_h, int im_w,
                                float thresh, uint8_t* out, float* out_prob, cudaStream_t stream) {
    if (num_dets < 0 || num_classes < 1 || mask_size < 1 || im_h < 1 || im_w < 1) return VOSD_ERR_BAD_SHAPE;
    if (num_dets == 0) return VOSD_OK;
    if (!masks || !ref_boxes || !out) return VOSD_ERR_BAD_ARG;
    if (!aligned16(out) || (out_prob && !aligned16(out_prob))) return VOSD_ERR_BAD_ARG;
    const long long total = (long long)num_dets * im_h * im_w;
    const long long chunks = (total + 15) / 16;
    long long blocks = (chunks + 255) / 256;
    if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
    const long long frame = (long long)im_h * im_w;
    if (frame % 16 == 0 && frame < (1LL << 31) && num_dets <= 65535) {
        // 2048 chunks (32 KB of output) per CTA: few, fat CTAs; tables cover at most the rows a CTA spans
        const int chunks_per_cta = 2048;
        const unsigned bx = (unsigned)((frame / 16 + chunks_per_cta - 1) / chunks_per_cta);
        const int rows_per_cta = (chunks_per_cta * 16 + im_w - 1) / im_w + 2;
        const int S = mask_size + 2;
        const size_t smem = (((size_t)S * S * 4 + 15) & ~(size_t)15) + (size_t)(im_w + im_w / 16 + 1 + rows_per_cta) * 16;
        if (smem > 200 * 1024) return VOSD_ERR_UNSUPPORTED;
        dim3 grid(bx, (unsigned)num_dets);
        if (out_prob) {
            if (cudaFuncSetAttribute(paste_det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            paste_det_kernel<true><<<grid, 256, smem, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                                chunks_per_cta, thresh, out, out_prob);
        } else {
            if (cudaFuncSetAttribute(paste_det_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            paste_det_kernel<false><<<grid, 256, smem, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                                 chunks_per_cta, thresh, out, out_prob);
        }
    } else if (out_prob) {
        paste_kernel<true><<<(int)blocks, 256, 0, stream>>>(masks, cls, ref_boxes, total, num_classes, mask_size,
                                                            im_h, im_w, thresh, out, out_prob);
    } else {
        paste_kernel<false><<<(int)blocks, 256, 0, stream>>>(masks, cls, ref_boxes, total, num_classes, mask_size,
                                                             im_h, im_w, thresh, out, out_prob);
    }
    count_launch();
    return check_launch();
}
