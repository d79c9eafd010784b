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

// Fast variant (im_h*im_w % 16 == 0): grid = (chunks per frame / 256, detections).  The detection is
// fixed per CTA, so a chunk that misses the box rows/columns costs one 32-bit division, a few
// compares and one 128-bit store.
template <bool kProb>
__global__ void __launch_bounds__(256)
paste_det_kernel(const float* __restrict__ masks, const int* __restrict__ cls,
                 const float* __restrict__ ref_boxes, int K, int M, int im_h, int im_w,
                 float thresh, uint8_t* __restrict__ out, float* __restrict__ out_prob) {
    const int r = blockIdx.y;
    const unsigned frame = (unsigned)im_h * (unsigned)im_w;
    const DetGeom g = det_geometry(ref_boxes + 4 * (size_t)r, M);
    const int ya = max(g.y0, 0), yb = min(g.y1 + 1, im_h), xa = max(g.x0, 0), xb = min(g.x1 + 1, im_w);
    uint8_t* __restrict__ o = out + (size_t)r * frame;
    float* __restrict__ op = kProb ? out_prob + (size_t)r * frame : nullptr;
    for (unsigned f0 = (blockIdx.x * 256u + threadIdx.x) * 16u; f0 < frame; f0 += gridDim.x * 256u * 16u) {
        const int y = (int)(f0 / (unsigned)im_w), x = (int)(f0 - (unsigned)y * (unsigned)im_w);
        const int y_last = (int)((f0 + 15u) / (unsigned)im_w);
        uint32_t packed[4] = {0u, 0u, 0u, 0u};
        float pv[16];
        if (kProb) {
#pragma unroll
            for (int i = 0; i < 16; i++) pv[i] = 0.f;
        }
        const bool rows_hit = y_last >= ya && y < yb && xa < xb;
        const bool cols_hit = (y_last > y) || (x < xb && x + 16 > xa);
        if (rows_hit && cols_hit)
            paste_chunk<kProb>(masks, cls, ref_boxes, r, y, x, 16, K, M, im_h, im_w, thresh, packed, pv);
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

}  // namespace vosd

using namespace vosd;

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
        // ~8 chunks (128 B) per thread: few, fat CTAs instead of one 4 KB CTA per 256 chunks
        unsigned bx = (unsigned)((frame / 16 + 256 * 8 - 1) / (256 * 8));
        dim3 grid(bx < 1 ? 1 : bx, (unsigned)num_dets);
        if (out_prob)
            paste_det_kernel<true><<<grid, 256, 0, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                             thresh, out, out_prob);
        else
            paste_det_kernel<false><<<grid, 256, 0, stream>>>(masks, cls, ref_boxes, num_classes, mask_size, im_h, im_w,
                                                              thresh, out, out_prob);
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
