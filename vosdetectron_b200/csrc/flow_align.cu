// FlowAlign forward / backward for sm_100a: warp a feature map by an optical-flow field (bilinear),
// gradients to the features and to the flow.  SURVEY.md section 8f, rank 4.
//
// Reference: lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu
//   FlowAlignForward_kernel  :15-55   one thread per (n, c, h, w) element, geometry recomputed per channel
//   FlowAlignBackward_kernel :57-117  4 atomicAdd per element into bottomdiff + 2 into flowdiff
//   launchers                :120-156 (512-thread 1-D grid, exit(-1) on a launch error)
//
// Design (B200): the geometry (flow fetch, bounds test, floor, ratios) depends on the pixel only, so a
// thread owns a short column of pixels (2 rows forward, 1 backward), derives their geometry ONCE and then
// streams over a chunk of up to 64 channels with lanes along x: coalesced 128-byte stores, near-coalesced tap
// loads through L1 (the flow is a small displacement, so the four taps of neighbouring lanes share lines).
// Taps are fetched unconditionally (dead pixels read texel 0) into two register sets, channel c+1 in flight
// while channel c is computed, and the tap rows of channel c+4 are prefetched into L2: the kernels are
// bound by bytes in flight (measured: 0.34 -> 0.21 ms with the prefetch, -> 0.175 ms with 2 rows per thread
// and a 73-register cap; 62 % of the measured HBM copy bandwidth).  All FPN levels go in one launch
// (block ranges per level).
//
// Arithmetic: the reference's expressions mix float and double operands (the literals `1.` are doubles);
// the kernels spell the same expressions with the same operand types, so nvcc emits the same multiply / fma
// sequence and the forward is BIT-identical to the reference kernel (gated by the tests).  The backward's
// addends are the reference's; they are pre-summed in registers where adjacent lanes (horizontal neighbours)
// or consecutive rows of a thread hit the same texel, which the reference's atomics would add one by one in
// arbitrary order, and the flow gradient is summed over the channel chunk before one reduction per pixel.
// An alternative that widens float->double on the ALU pipe instead of the 16-lane XU pipe (also
// bit-identical) is kept behind vosd_debug_flow_align_fast(2): it measured slower (more issue slots and
// registers than the conversions cost).
#include <atomic>
#include "common.cuh"

namespace vosd {
namespace {

constexpr int kFlowWarps = 4;   // row bands per CTA (one warp each)
constexpr int kFlowThreads = kFlowWarps * 32;
// Tuning (A/B builds on the B200, profiles/README.md): rows per thread, channels per thread, register cap.
// The kernels are bound by bytes in flight, so occupancy beats per-thread reuse: 2 rows (forward) / 1 row
// (backward: the vertical pre-summing of shared texels is worth less than the extra warps).
#ifndef VOSD_FLOW_ROWS_FWD
#define VOSD_FLOW_ROWS_FWD 2
#endif
#ifndef VOSD_FLOW_ROWS_BWD
#define VOSD_FLOW_ROWS_BWD 1
#endif
#ifndef VOSD_FLOW_CHUNK
#define VOSD_FLOW_CHUNK 64
#endif
#ifndef VOSD_FLOW_MINB_FWD
#define VOSD_FLOW_MINB_FWD 7
#endif
#ifndef VOSD_FLOW_MINB_BWD
#define VOSD_FLOW_MINB_BWD 8
#endif
constexpr int kFlowRowsFwd = VOSD_FLOW_ROWS_FWD;
constexpr int kFlowRowsBwd = VOSD_FLOW_ROWS_BWD;

struct FlowLevelArgs {
    const float* bottom;    // (N,C,H,W) features
    const float* flow;      // (N,2,H,W) [x displacement plane, y displacement plane]
    const float* topdiff;   // backward only: (N,C,H,W)
    float* out0;            // forward: top (N,C,H,W); backward: bottomdiff (N,C,H,W)
    float* out1;            // backward: flowdiff (N,2,H,W)
    int H, W;
    int tiles_x, tiles_y;   // tiles of 32 columns x (kFlowWarps * rows per thread) rows
    int block_begin;        // first block of this level in the grid
    int pad;
};

struct FlowArgs {
    FlowLevelArgs lv[VOSD_MAX_LEVELS];
    int num_levels, N, C, chunk, chunks;
};

struct FlowTile {
    int level, tx, ty, chunk, n;
};

__device__ __forceinline__ FlowTile decode_tile(const FlowArgs& a) {
    const int b = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < VOSD_MAX_LEVELS; ++i)
        if (i < a.num_levels && b >= a.lv[i].block_begin) l = i;
    int t = b - a.lv[l].block_begin;
    FlowTile r;
    r.level = l;
    r.tx = t % a.lv[l].tiles_x; t /= a.lv[l].tiles_x;
    r.ty = t % a.lv[l].tiles_y; t /= a.lv[l].tiles_y;
    r.chunk = t % a.chunks;
    r.n = t / a.chunks;
    return r;
}

// Geometry of one pixel, flow_align_cuda_kernel.cu:24-45 (identical float operations).
// Returns the offset of the up-left tap inside a channel plane, or -1 when the reference writes 0 /
// propagates nothing (sample outside [0,H-1) x [0,W-1)).
__device__ __forceinline__ int flow_geometry(const float* __restrict__ flow_n, int plane, int H, int W,
                                             int h, int w, float& h_ratio, float& w_ratio) {
    const float flo_x = __ldg(flow_n + h * W + w);
    const float flo_y = __ldg(flow_n + plane + h * W + w);
    const float w_flo = w + flo_x;
    const float h_flo = h + flo_y;
    h_ratio = 0.f;
    w_ratio = 0.f;
    if (h_flo < 0 || h_flo >= H - 1 || w_flo < 0 || w_flo >= W - 1) return -1;
    // A NaN flow passes every comparison above; the reference then reads the taps of (0,0) and produces
    // NaN.  With H < 2 or W < 2 that read leaves the plane: refuse it (the output is 0 there).
    if ((h_flo != h_flo || w_flo != w_flo) && (H < 2 || W < 2)) return -1;
    const int h_start = floorf(h_flo);
    const int w_start = floorf(w_flo);
    h_ratio = h_flo - (float)h_start;
    w_ratio = w_flo - (float)w_start;
    return w_start + W * h_start;
}

// L2 prefetch distance in channels (0 = off): the tap rows of channel c + kFlowPrefetch are requested into L2 while
// channel c is computed, so the register-held loads of the next channel see L2 latency instead of DRAM latency
// (the kernels are bound by bytes in flight: 12-24 resident warps x 2 channels of taps per thread).
#ifndef VOSD_FLOW_PF
#define VOSD_FLOW_PF 4
#endif
constexpr int kFlowPrefetch = VOSD_FLOW_PF;

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// ------------------------------------------------------------------------------------- forward
// Arithmetic variants of the forward (same loads, same stores):
//   kFlowExactCvt   the reference's expression as written (:48-51): 4 F2F.F64.F32 + 1 F2F.F32.F64 per element, all on
//                   the 16-lane XU pipe, which then bounds the kernel (63 % XU utilisation in the first capture);
//   kFlowExactAlu   the same double-precision operation sequence with the float->double widening done as a bit
//                   shuffle on the ALU pipe: widen_scaled(x) = x * 2^-896 exactly (finite x, zero and denormals
//                   included); the 2^896 is folded into the per-pixel factors, so every rounding happens on the same
//                   real value as in the reference -> the same bits.  Elements with a non-finite tap take the
//                   F2F expression;
//   kFlowFp32       plain fp32 bilinear weights (test / tuning hook, not bit-identical).
enum { kFlowExactCvt = 0, kFlowFp32 = 1, kFlowExactAlu = 2 };

// x * 2^-896 as a double, exact for every finite float (a float's exponent field e8 becomes the double's
// exponent field; denormals map to double denormals with the same scale factor).
__device__ __forceinline__ double widen_scaled(float x) {
    const int u = __float_as_int(x);
    return __hiloint2double((u >> 3) & 0x8FFFFFFF, (int)((unsigned)u << 29));
}

// flow_align_cuda_kernel.cu:48-51 exactly as written there (operand types included).
__device__ __forceinline__ float flow_bilinear_ref(float b1, float b2, float b3, float b4, float h_ratio, float w_ratio) {
    return b1 * (1. - h_ratio) * (1. - w_ratio)
         + b2 * (1. - h_ratio) * (w_ratio)
         + b3 * (h_ratio) * (1. - w_ratio)
         + b4 * (h_ratio) * (w_ratio);
}
__device__ __noinline__ float flow_bilinear_ref_cold(float b1, float b2, float b3, float b4, float h_ratio, float w_ratio) {
    return flow_bilinear_ref(b1, b2, b3, b4, h_ratio, w_ratio);
}

// Per-pixel double factors kept in registers across the channel loop.
template <int MODE> struct FlowFactors;
template <> struct FlowFactors<kFlowExactCvt> { __device__ void set(float, float) {} };
template <> struct FlowFactors<kFlowFp32> { __device__ void set(float, float) {} };
template <> struct FlowFactors<kFlowExactAlu> {
    double oh_s, ow, ow_s, wd;      // (1-h)*2^896, (1-w), (1-w)*2^896, (double)w
    __device__ void set(float h_ratio, float w_ratio) {
        const double oh = 1. - h_ratio;
        ow = 1. - w_ratio;
        oh_s = oh * 0x1p896;
        ow_s = ow * 0x1p896;
        wd = w_ratio;
    }
};

template <int MODE>
__device__ __forceinline__ float flow_bilinear(float b1, float b2, float b3, float b4, float h_ratio, float w_ratio,
                                               const FlowFactors<MODE>& f) {
    if constexpr (MODE == kFlowFp32) {
        const float fh = 1.f - h_ratio, fw = 1.f - w_ratio;
        return fmaf(b4, h_ratio * w_ratio, fmaf(b3, h_ratio * fw, fmaf(b2, fh * w_ratio, b1 * (fh * fw))));
    } else if constexpr (MODE == kFlowExactAlu) {
        // operation order of the reference's SASS: T2 = (b2*A)*w; fma(b1*A, B, T2); fma(B, b3*h, .); + b4*h*w
        double acc = __dmul_rn(__dmul_rn(widen_scaled(b2), f.oh_s), f.wd);
        acc = __fma_rn(__dmul_rn(widen_scaled(b1), f.oh_s), f.ow, acc);
        acc = __fma_rn(f.ow_s, widen_scaled(__fmul_rn(b3, h_ratio)), acc);
        acc = __fma_rn(widen_scaled(__fmul_rn(__fmul_rn(b4, h_ratio), w_ratio)), 0x1p896, acc);
        float v = __double2float_rn(acc);
        const float s = (b1 + b2) + (b3 + b4);                      // non-finite iff some tap is (or the sum overflows)
        if ((__float_as_uint(s) & 0x7f800000u) == 0x7f800000u)      // rare: the widening shuffle is for finite taps
            v = flow_bilinear_ref_cold(b1, b2, b3, b4, h_ratio, w_ratio);
        return v;
    } else {
        return flow_bilinear_ref(b1, b2, b3, b4, h_ratio, w_ratio);
    }
}

template <int MODE>
__global__ void __launch_bounds__(kFlowThreads, VOSD_FLOW_MINB_FWD) flow_align_fwd_kernel(const __grid_constant__ FlowArgs a) {
    constexpr int kFlowRows = kFlowRowsFwd;
    const FlowTile t = decode_tile(a);
    const FlowLevelArgs& L = a.lv[t.level];
    const int H = L.H, W = L.W, plane = H * W;      // H, W >= 2 (smaller maps never reach the kernel)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = t.tx * 32 + lane;
    const int h0 = (t.ty * kFlowWarps + warp) * kFlowRows;
    if (h0 >= H) return;   // warp-uniform

    int off[kFlowRows];         // -2: no pixel here, -1: pixel written as 0, >= 0: up-left tap
    int ld[kFlowRows];          // tap offset actually loaded: taps are fetched unconditionally (offset 0 for dead pixels)
    float hr[kFlowRows], wr[kFlowRows];
    FlowFactors<MODE> fac[kFlowRows];
    const float* flow_n = L.flow + (size_t)t.n * 2 * plane;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        off[r] = -2;
        hr[r] = wr[r] = 0.f;
        if (h0 + r < H && w < W) off[r] = flow_geometry(flow_n, plane, H, W, h0 + r, w, hr[r], wr[r]);
        ld[r] = max(off[r], 0);
        fac[r].set(hr[r], wr[r]);
    }

    const int c0 = t.chunk * a.chunk;
    const int c1 = min(c0 + a.chunk, a.C);
    const float* base = L.bottom + ((size_t)t.n * a.C + c0) * plane;
    float* out = L.out0 + ((size_t)t.n * a.C + c0) * plane + (size_t)h0 * W + w;

    // Two register sets: the taps of channel c+1 are in flight while channel c is computed and stored.
    float cur[kFlowRows][4], nxt[kFlowRows][4];
    auto fetch = [&](float (&b)[kFlowRows][4], const float* p) {
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            b[r][0] = __ldg(p + ld[r]);
            b[r][1] = __ldg(p + ld[r] + 1);
            b[r][2] = __ldg(p + ld[r] + W);
            b[r][3] = __ldg(p + ld[r] + W + 1);
        }
    };
    auto emit = [&](const float (&b)[kFlowRows][4], float* o) {
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            float v = flow_bilinear<MODE>(b[r][0], b[r][1], b[r][2], b[r][3], hr[r], wr[r], fac[r]);
            if (off[r] < 0) v = 0.f;
            if (off[r] != -2) __stcs(o + r * W, v);
        }
    };
    auto prefetch = [&](const float* p) {          // one line per tap row of the thread's pixel column
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) prefetch_l2(p + ld[r]);
        prefetch_l2(p + ld[kFlowRows - 1] + W);
    };
    fetch(cur, base);
    int c = c0;
    for (; c + 2 <= c1; c += 2) {
        fetch(nxt, base + plane);
        if (kFlowPrefetch && c + kFlowPrefetch < c1) prefetch(base + kFlowPrefetch * (size_t)plane);
        emit(cur, out);
        if (c + 2 < c1) fetch(cur, base + 2 * (size_t)plane);
        if (kFlowPrefetch && c + 1 + kFlowPrefetch < c1) prefetch(base + (1 + kFlowPrefetch) * (size_t)plane);
        emit(nxt, out + plane);
        base += 2 * (size_t)plane;
        out += 2 * (size_t)plane;
    }
    if (c < c1) emit(cur, out);
}

// ------------------------------------------------------------------------------------- backward
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// The six values of one element, flow_align_cuda_kernel.cu:89-112 exactly as written there:
// a1..a4 = what the reference hands to atomicAdd(bottomdiff + tap, .), dx / dy = its flow-gradient factors.
struct FlowGrad {
    float a1, a2, a3, a4, dx, dy;
};
__device__ __forceinline__ FlowGrad flow_grad_ref(float g, float f1, float f2, float f3, float f4,
                                                  float h_ratio, float w_ratio) {
    FlowGrad o;
    o.a1 = g * (1. - h_ratio) * (1. - w_ratio);
    o.a2 = g * (1. - h_ratio) * (w_ratio);
    o.a3 = g * (h_ratio) * (1. - w_ratio);
    o.a4 = g * (h_ratio) * (w_ratio);
    o.dx = -f1 * (1. - h_ratio) + f2 * (1. - h_ratio) - f3 * (h_ratio) + f4 * (h_ratio);
    o.dy = -f1 * (1. - w_ratio) - f2 * (w_ratio) + f3 * (1. - w_ratio) + f4 * (w_ratio);
    return o;
}
__device__ __noinline__ void flow_grad_ref_cold(FlowGrad* o, float g, float f1, float f2, float f3, float f4,
                                                float h_ratio, float w_ratio) {
    *o = flow_grad_ref(g, f1, f2, f3, f4, h_ratio, w_ratio);
}

// Same double-precision operation sequence (read from the reference's SASS, see oracle/oracle.c) with most
// float->double widenings done on the ALU pipe (widen_scaled) and the 2^896 folded into the pixel's factors;
// three widenings stay on the XU pipe to balance the two.  Bit-identical for finite inputs; anything else
// takes the expression as written.
struct FlowBwdFactors {
    double oh, oh_s, ow, ow_s, wd;
    __device__ __forceinline__ void set(float h_ratio, float w_ratio) {
        oh = 1. - h_ratio;
        ow = 1. - w_ratio;
        oh_s = oh * 0x1p896;
        ow_s = ow * 0x1p896;
        wd = w_ratio;
    }
};
__device__ __forceinline__ FlowGrad flow_grad(float g, float f1, float f2, float f3, float f4, float h_ratio,
                                              float w_ratio, const FlowBwdFactors& k) {
    FlowGrad o;
    const float gh = __fmul_rn(g, h_ratio);
    const double ga = __dmul_rn(widen_scaled(g), k.oh_s);                         // d(g) * (1-h)
    o.a1 = __double2float_rn(__dmul_rn(ga, k.ow));
    o.a2 = __double2float_rn(__dmul_rn(ga, k.wd));
    o.a3 = __double2float_rn(__dmul_rn(k.ow, (double)gh));
    o.a4 = __fmul_rn(gh, w_ratio);
    const double nf1 = widen_scaled(-f1);
    double x = __dmul_rn(k.oh_s, widen_scaled(f2));                               // A*f2
    x = __fma_rn(k.oh_s, nf1, x);                                                 // A*(-f1) + .
    x = __fma_rn(widen_scaled(__fmul_rn(h_ratio, f3)), -0x1p896, x);              // . - d(h*f3)
    x = __fma_rn(widen_scaled(__fmul_rn(h_ratio, f4)), 0x1p896, x);               // . + d(h*f4)
    o.dx = __double2float_rn(x);
    double y = __dmul_rn(widen_scaled(__fmul_rn(w_ratio, f2)), -0x1p896);         // -d(w*f2), exact
    y = __fma_rn(k.ow_s, nf1, y);                                                 // B*(-f1) - d(w*f2)
    y = __fma_rn(k.ow, (double)f3, y);                                            // B*f3 + .
    y = y + (double)__fmul_rn(w_ratio, f4);                                       // . + d(w*f4)
    o.dy = __double2float_rn(y);
    const float s = ((f1 + f2) + (f3 + f4)) + g;
    if ((__float_as_uint(s) & 0x7f800000u) == 0x7f800000u)                        // rare: some input is not finite
        flow_grad_ref_cold(&o, g, f1, f2, f3, f4, h_ratio, w_ratio);
    return o;
}

template <bool ALU, bool WANT_FLOW>
__global__ void __launch_bounds__(kFlowThreads, VOSD_FLOW_MINB_BWD) flow_align_bwd_kernel(const __grid_constant__ FlowArgs a) {
    constexpr int kFlowRows = kFlowRowsBwd;
    const FlowTile t = decode_tile(a);
    const FlowLevelArgs& L = a.lv[t.level];
    const int H = L.H, W = L.W, plane = H * W;      // H, W >= 2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = t.tx * 32 + lane;
    const int h0 = (t.ty * kFlowWarps + warp) * kFlowRows;
    if (h0 >= H) return;   // warp-uniform: every shuffle below is executed by all 32 lanes

    int off[kFlowRows];         // < 0: nothing flows back from this pixel
    int ld[kFlowRows], ldg[kFlowRows];
    float hr[kFlowRows], wr[kFlowRows];
    FlowBwdFactors fac[kFlowRows];
    const float* flow_n = L.flow + (size_t)t.n * 2 * plane;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        off[r] = -1;
        hr[r] = wr[r] = 0.f;
        const bool inside = h0 + r < H && w < W;
        if (inside) off[r] = flow_geometry(flow_n, plane, H, W, h0 + r, w, hr[r], wr[r]);
        ld[r] = max(off[r], 0);                                  // loads are unconditional: dead pixels read texel 0
        ldg[r] = inside ? (h0 + r) * W + w : 0;
        if (ALU) fac[r].set(hr[r], wr[r]);
    }
    // Merge flags, one bit per row (pixel-only, hoisted out of the channel loop):
    //   vm: this pixel's top taps are the previous row's bottom taps (same thread) -> the pending pair is folded in
    //   hl: this pixel's left taps are the left lane's right taps -> that lane's right column is folded in here
    //   hrm: the right lane folds this lane's right column -> this lane does not emit it
    unsigned vm = 0, hl = 0, hrm = 0;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        if (r > 0 && off[r] >= 0 && off[r - 1] >= 0 && off[r] == off[r - 1] + W) vm |= 1u << r;
        const int left = __shfl_up_sync(0xffffffffu, off[r], 1);
        const bool l = lane > 0 && off[r] >= 0 && left >= 0 && off[r] == left + 1;
        const bool rr = __shfl_down_sync(0xffffffffu, (int)l, 1) && lane < 31;
        if (l) hl |= 1u << r;
        if (rr) hrm |= 1u << r;
    }

    const int c0 = t.chunk * a.chunk;
    const int c1 = min(c0 + a.chunk, a.C);
    const size_t cbase = ((size_t)t.n * a.C + c0) * plane;
    const float* base = L.bottom + cbase;
    const float* td = L.topdiff + cbase;
    float* bd = L.out0 + cbase;
    float gx[kFlowRows], gy[kFlowRows];
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) gx[r] = gy[r] = 0.f;

    // Two register sets: the five loads per pixel of channel c+1 are in flight while channel c is scattered.
    float cur[kFlowRows][5], nxt[kFlowRows][5];
    auto fetch = [&](float (&b)[kFlowRows][5], const float* p, const float* q) {
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            if (WANT_FLOW) {                              // the taps only feed the flow gradient
                b[r][0] = __ldg(p + ld[r]);
                b[r][1] = __ldg(p + ld[r] + 1);
                b[r][2] = __ldg(p + ld[r] + W);
                b[r][3] = __ldg(p + ld[r] + W + 1);
            } else {
                b[r][0] = b[r][1] = b[r][2] = b[r][3] = 0.f;
            }
            b[r][4] = __ldcs(q + ldg[r]);
        }
    };
    // emit: top taps of row r (+ the pending bottom taps of row r-1 when they coincide), bottom taps of the last
    // row; horizontally, a lane's right column goes to the right lane when that lane's left column is the same
    // texel column.
    auto scatter = [&](const float (&b)[kFlowRows][5], float* d) {
        float pl = 0.f, pr = 0.f;     // pending bottom pair of the previous row
#pragma unroll
        for (int r = 0; r <= kFlowRows; ++r) {
            float tl, tr;
            int o;
            unsigned bit;
            if (r < kFlowRows) {
                FlowGrad v = ALU ? flow_grad(b[r][4], b[r][0], b[r][1], b[r][2], b[r][3], hr[r], wr[r], fac[r])
                                 : flow_grad_ref(b[r][4], b[r][0], b[r][1], b[r][2], b[r][3], hr[r], wr[r]);
                if (off[r] < 0) v.a1 = v.a2 = v.a3 = v.a4 = v.dx = v.dy = 0.f;
                if (WANT_FLOW) {
                    gx[r] += __fmul_rn(b[r][4], v.dx);           // :111-112, summed over the chunk first
                    gy[r] += __fmul_rn(b[r][4], v.dy);
                }
                if (r > 0 && !((vm >> r) & 1u) && off[r - 1] >= 0) {   // previous row's bottom pair stands alone
                    red_add(d + off[r - 1] + W, pl);
                    red_add(d + off[r - 1] + W + 1, pr);
                    pl = pr = 0.f;
                }
                tl = v.a1 + (((vm >> r) & 1u) ? pl : 0.f);
                tr = v.a2 + (((vm >> r) & 1u) ? pr : 0.f);
                o = off[r];
                bit = r;
                pl = v.a3;
                pr = v.a4;
            } else {                                                   // bottom pair of the thread's last row
                tl = pl;
                tr = pr;
                o = off[kFlowRows - 1] >= 0 ? off[kFlowRows - 1] + W : -1;
                bit = kFlowRows - 1;
            }
            const float from_left = __shfl_up_sync(0xffffffffu, tr, 1);
            if ((hl >> bit) & 1u) tl += from_left;
            if (o >= 0) {
                red_add(d + o, tl);
                if (!((hrm >> bit) & 1u)) red_add(d + o + 1, tr);
            }
        }
    };
    auto prefetch = [&](const float* p, const float* q) {
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            if (WANT_FLOW) prefetch_l2(p + ld[r]);
            prefetch_l2(q + ldg[r]);
        }
        if (WANT_FLOW) prefetch_l2(p + ld[kFlowRows - 1] + W);
    };
    fetch(cur, base, td);
    int c = c0;
    for (; c + 2 <= c1; c += 2) {
        fetch(nxt, base + plane, td + plane);
        if (kFlowPrefetch && c + kFlowPrefetch < c1)
            prefetch(base + kFlowPrefetch * (size_t)plane, td + kFlowPrefetch * (size_t)plane);
        scatter(cur, bd);
        if (c + 2 < c1) fetch(cur, base + 2 * (size_t)plane, td + 2 * (size_t)plane);
        if (kFlowPrefetch && c + 1 + kFlowPrefetch < c1)
            prefetch(base + (1 + kFlowPrefetch) * (size_t)plane, td + (1 + kFlowPrefetch) * (size_t)plane);
        scatter(nxt, bd + plane);
        base += 2 * (size_t)plane;
        td += 2 * (size_t)plane;
        bd += 2 * (size_t)plane;
    }
    if (c < c1) scatter(cur, bd);
    // flow gradient: one reduction per (pixel, channel chunk) instead of one per channel
    if (!WANT_FLOW) return;
    float* fd = L.out1 + (size_t)t.n * 2 * plane;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r)
        if (off[r] >= 0) {
            red_add(fd + ldg[r], gx[r]);
            red_add(fd + plane + ldg[r], gy[r]);
        }
}

std::atomic<int> g_flow_fast{kFlowExactCvt};

// Fills the block ranges; returns the grid size (0: nothing to do) or a negative status.
long long plan(FlowArgs& a, int rows, int num_levels, int batches, int channels, const int* level_h, const int* level_w) {
    if (num_levels < 0 || num_levels > VOSD_MAX_LEVELS || batches < 0 || channels < 0) return VOSD_ERR_BAD_SHAPE;
    a.num_levels = num_levels;
    a.N = batches;
    a.C = channels;
    long long tiles = 0;
    for (int l = 0; l < num_levels; ++l) {
        if (level_h[l] < 0 || level_w[l] < 0) return VOSD_ERR_BAD_SHAPE;
        // the reference indexes elements with a 32-bit int (flow_align_cuda_kernel.cu:18-22)
        if ((long long)batches * channels * level_h[l] * level_w[l] >= (1ll << 31)) return VOSD_ERR_UNSUPPORTED;
        a.lv[l].H = level_h[l];
        a.lv[l].W = level_w[l];
        // a map with H < 2 or W < 2 has no sample inside [0,H-1) x [0,W-1): the output is all zero and nothing
        // flows back; it gets no tiles (the kernels fetch the 2x2 taps of offset 0 unconditionally)
        const bool live = level_h[l] >= 2 && level_w[l] >= 2;
        a.lv[l].tiles_x = live ? ceil_div(level_w[l], 32) : 0;
        a.lv[l].tiles_y = live ? ceil_div(level_h[l], kFlowWarps * rows) : 0;
        tiles += (long long)a.lv[l].tiles_x * a.lv[l].tiles_y;
    }
    if (tiles == 0 || batches == 0 || channels == 0) return 0;
    // channel chunk per thread: long enough to amortise the per-pixel geometry, short enough to fill 148 SMs
    int chunk = VOSD_FLOW_CHUNK;
    while (chunk > 8 && tiles * batches * ceil_div(channels, chunk) < (long long)kNumSMs * 16) chunk /= 2;
    a.chunk = chunk;
    a.chunks = ceil_div(channels, chunk);
    long long begin = 0;
    for (int l = 0; l < num_levels; ++l) {
        a.lv[l].block_begin = (int)begin;
        begin += (long long)a.lv[l].tiles_x * a.lv[l].tiles_y * a.chunks * batches;
        if (begin >= (1ll << 31)) return VOSD_ERR_UNSUPPORTED;
    }
    // (an empty level shares its block_begin with its successor; decode_tile() picks the last match)
    return begin;
}

int flow_fwd(int num_levels, int batches, int channels, const int* level_h, const int* level_w,
             const float* const* bottom, const float* const* flow, float* const* top, cudaStream_t stream) {
    if (num_levels > 0 && (!level_h || !level_w || !bottom || !flow || !top)) return VOSD_ERR_BAD_ARG;
    FlowArgs a = {};
    const long long grid = plan(a, kFlowRowsFwd, num_levels, batches, channels, level_h, level_w);
    if (grid < 0) return (int)grid;
    for (int l = 0; l < num_levels; ++l) {
        const size_t elems = (size_t)batches * channels * a.lv[l].H * a.lv[l].W;
        if (elems && (!bottom[l] || !flow[l] || !top[l])) return VOSD_ERR_BAD_ARG;
        a.lv[l].bottom = bottom[l];
        a.lv[l].flow = flow[l];
        a.lv[l].out0 = top[l];
        if (elems && a.lv[l].tiles_x == 0 &&
            cudaMemsetAsync(top[l], 0, elems * sizeof(float), stream) != cudaSuccess) return VOSD_ERR_LAUNCH;
    }
    if (grid == 0) return VOSD_OK;
    switch (g_flow_fast.load(std::memory_order_relaxed)) {
        case kFlowExactCvt: flow_align_fwd_kernel<kFlowExactCvt><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a); break;
        case kFlowFp32: flow_align_fwd_kernel<kFlowFp32><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a); break;
        case kFlowExactAlu: flow_align_fwd_kernel<kFlowExactAlu><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a); break;
        default: return VOSD_ERR_BAD_ARG;
    }
    count_launch();
    return check_launch();
}

int flow_bwd(int num_levels, int batches, int channels, const int* level_h, const int* level_w,
             const float* const* topdiff, const float* const* bottom, const float* const* flow,
             float* const* bottomdiff, float* const* flowdiff, int zero_init, cudaStream_t stream) {
    if (num_levels > 0 && (!level_h || !level_w || !topdiff || !bottom || !flow || !bottomdiff))
        return VOSD_ERR_BAD_ARG;
    // flowdiff == NULL (the table, or every entry): the flow gradient is not wanted -- the tap loads, their
    // conversions and the dx / dy arithmetic (:95-112) are compiled out
    bool want_flow = flowdiff != nullptr;
    if (want_flow) {
        int have = 0, live = 0;                       // empty levels carry no pointers at all
        for (int l = 0; l < num_levels; ++l) {
            if ((long long)level_h[l] * level_w[l] * batches == 0) continue;
            ++live;
            have += flowdiff[l] != nullptr;
        }
        if (have == 0) want_flow = false;
        else if (have != live) return VOSD_ERR_BAD_ARG;
    }
    FlowArgs a = {};
    const long long grid = plan(a, kFlowRowsBwd, num_levels, batches, channels, level_h, level_w);
    if (grid < 0) return (int)grid;
    for (int l = 0; l < num_levels; ++l) {
        const size_t plane = (size_t)level_h[l] * level_w[l];
        if (plane && batches && channels && (!topdiff[l] || !bottom[l] || !flow[l] || !bottomdiff[l]))
            return VOSD_ERR_BAD_ARG;
        a.lv[l].topdiff = topdiff[l];
        a.lv[l].bottom = bottom[l];
        a.lv[l].flow = flow[l];
        a.lv[l].out0 = bottomdiff[l];
        a.lv[l].out1 = want_flow ? flowdiff[l] : nullptr;
        if (zero_init && plane && batches) {
            // FlowAlignFunction.backward zero-fills both gradients itself (functions/flow_align.py:41-43)
            if (channels && cudaMemsetAsync(bottomdiff[l], 0, plane * batches * channels * sizeof(float), stream) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            if (want_flow && cudaMemsetAsync(flowdiff[l], 0, plane * batches * 2 * sizeof(float), stream) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
        }
    }
    if (grid == 0) return VOSD_OK;
    const bool alu = g_flow_fast.load(std::memory_order_relaxed) == kFlowExactAlu;
    if (alu && want_flow) flow_align_bwd_kernel<true, true><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    else if (alu) flow_align_bwd_kernel<true, false><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    else if (want_flow) flow_align_bwd_kernel<false, true><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    else flow_align_bwd_kernel<false, false><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    count_launch();
    return check_launch();
}
}  // namespace
}  // namespace vosd

extern "C" int vosd_flow_align_fwd(int batches, int height, int width, int channels, const float* bottom,
                                   const float* flow, float* top, cudaStream_t stream) {
    return vosd::flow_fwd(1, batches, channels, &height, &width, &bottom, &flow, &top, stream);
}

extern "C" int vosd_flow_align_bwd(int batches, int height, int width, int channels, const float* topdiff,
                                   const float* bottom, const float* flow, float* bottomdiff, float* flowdiff,
                                   int zero_init, cudaStream_t stream) {
    return vosd::flow_bwd(1, batches, channels, &height, &width, &topdiff, &bottom, &flow, &bottomdiff, &flowdiff,
                          zero_init, stream);
}

extern "C" int vosd_flow_align_ml_fwd(int num_levels, int batches, int channels, const int* level_h,
                                      const int* level_w, const float* const* bottom, const float* const* flow,
                                      float* const* top, cudaStream_t stream) {
    return vosd::flow_fwd(num_levels, batches, channels, level_h, level_w, bottom, flow, top, stream);
}

extern "C" int vosd_flow_align_ml_bwd(int num_levels, int batches, int channels, const int* level_h,
                                      const int* level_w, const float* const* topdiff, const float* const* bottom,
                                      const float* const* flow, float* const* bottomdiff, float* const* flowdiff,
                                      int zero_init, cudaStream_t stream) {
    return vosd::flow_bwd(num_levels, batches, channels, level_h, level_w, topdiff, bottom, flow, bottomdiff,
                          flowdiff, zero_init, stream);
}

extern "C" int vosd_debug_flow_align_fast(int on) {
    if (!vosd::test_hooks_enabled()) return vosd::g_flow_fast.load(std::memory_order_relaxed);
    return vosd::g_flow_fast.exchange(on, std::memory_order_relaxed);
}
