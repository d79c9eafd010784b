// FlowAlign forward / backward for sm_100a: warp a feature map by an optical-flow field (bilinear),
// gradients to the features and to the flow.  SURVEY.md section 8f, rank 4.
//
// Reference: lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu
//   FlowAlignForward_kernel  :15-55   one thread per (n, c, h, w) element, geometry recomputed per channel
//   FlowAlignBackward_kernel :57-117  4 atomicAdd per element into bottomdiff + 2 into flowdiff
//   launchers                :120-156 (512-thread 1-D grid, exit(-1) on a launch error)
//
// Design (B200): the geometry (flow fetch, bounds test, floor, ratios) depends on the pixel only, so a
// thread owns a column of kFlowRows pixels, derives their geometry ONCE and then streams over a chunk of
// channels with lanes along x: coalesced 128-byte stores, near-coalesced tap loads through L1 (the flow is a
// small displacement, so the four taps of neighbouring lanes share lines), kFlowRows x 4 independent loads
// in flight per thread.  All FPN levels go in one launch (block ranges per level).
//
// Arithmetic: the reference's expressions mix float and double operands (the literals `1.` are doubles);
// the EXACT kernels spell the same expressions with the same operand types, so nvcc emits the same
// multiply / fma sequence and the forward is BIT-identical to the reference kernel (gated by the tests).
// The backward's addends are the reference's; they are pre-summed in registers where two pixels of a
// thread (vertical neighbours) or of adjacent lanes (horizontal neighbours) hit the same texel, which the
// reference's atomics would add one by one in arbitrary order.
#include <atomic>
#include "common.cuh"

namespace vosd {
namespace {

constexpr int kFlowWarps = 4;   // row bands per CTA (one warp each)
constexpr int kFlowRows = 4;    // rows per thread
constexpr int kFlowThreads = kFlowWarps * 32;

struct FlowLevelArgs {
    const float* bottom;    // (N,C,H,W) features
    const float* flow;      // (N,2,H,W) [x displacement plane, y displacement plane]
    const float* topdiff;   // backward only: (N,C,H,W)
    float* out0;            // forward: top (N,C,H,W); backward: bottomdiff (N,C,H,W)
    float* out1;            // backward: flowdiff (N,2,H,W)
    int H, W;
    int tiles_x, tiles_y;   // tiles of 32 columns x (kFlowWarps * kFlowRows) rows
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

// ------------------------------------------------------------------------------------- forward
template <bool EXACT>
__global__ void __launch_bounds__(kFlowThreads) flow_align_fwd_kernel(const __grid_constant__ FlowArgs a) {
    const FlowTile t = decode_tile(a);
    const FlowLevelArgs& L = a.lv[t.level];
    const int H = L.H, W = L.W, plane = H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = t.tx * 32 + lane;
    const int h0 = (t.ty * kFlowWarps + warp) * kFlowRows;
    if (h0 >= H) return;   // warp-uniform

    int off[kFlowRows];         // -2: no pixel here, -1: pixel written as 0, >= 0: up-left tap
    float hr[kFlowRows], wr[kFlowRows];
    const float* flow_n = L.flow + (size_t)t.n * 2 * plane;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        off[r] = -2;
        hr[r] = wr[r] = 0.f;
        if (h0 + r < H && w < W) off[r] = flow_geometry(flow_n, plane, H, W, h0 + r, w, hr[r], wr[r]);
    }

    double oh[kFlowRows], ow[kFlowRows], wd[kFlowRows];
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        oh[r] = 1. - hr[r];
        ow[r] = 1. - wr[r];
        wd[r] = wr[r];
    }

    const int c0 = t.chunk * a.chunk;
    const int c1 = min(c0 + a.chunk, a.C);
    const float* base = L.bottom + ((size_t)t.n * a.C + c0) * plane;
    float* out = L.out0 + ((size_t)t.n * a.C + c0) * plane + (size_t)h0 * W + w;
#pragma unroll 2
    for (int c = c0; c < c1; ++c, base += plane, out += plane) {
        float v[kFlowRows];
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            v[r] = 0.f;
            if (off[r] >= 0) {
                const float* p = base + off[r];
                const float b1 = __ldg(p), b2 = __ldg(p + 1), b3 = __ldg(p + W), b4 = __ldg(p + W + 1);
                const float h_ratio = hr[r], w_ratio = wr[r];
                if (EXACT) {
                    // flow_align_cuda_kernel.cu:48-51, operand types as in the reference; the two double
                    // factors (1. - h_ratio), (1. - w_ratio) and the widened w_ratio are per pixel and kept in registers
                    v[r] = b1 * oh[r] * ow[r]
                         + b2 * oh[r] * wd[r]
                         + b3 * (h_ratio) * ow[r]
                         + b4 * (h_ratio) * (w_ratio);
                } else {
                    const float oh = 1.f - h_ratio, ow = 1.f - w_ratio;
                    v[r] = fmaf(b4, h_ratio * w_ratio, fmaf(b3, h_ratio * ow, fmaf(b2, oh * w_ratio, b1 * (oh * ow))));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r)
            if (off[r] != -2) __stcs(out + r * W, v[r]);
    }
}

// ------------------------------------------------------------------------------------- backward
__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kFlowThreads) flow_align_bwd_kernel(const __grid_constant__ FlowArgs a) {
    const FlowTile t = decode_tile(a);
    const FlowLevelArgs& L = a.lv[t.level];
    const int H = L.H, W = L.W, plane = H * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = t.tx * 32 + lane;
    const int h0 = (t.ty * kFlowWarps + warp) * kFlowRows;
    if (h0 >= H) return;   // warp-uniform: every shuffle below is executed by all 32 lanes

    int off[kFlowRows];
    float hr[kFlowRows], wr[kFlowRows];
    const float* flow_n = L.flow + (size_t)t.n * 2 * plane;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) {
        off[r] = -1;
        hr[r] = wr[r] = 0.f;
        if (h0 + r < H && w < W) off[r] = flow_geometry(flow_n, plane, H, W, h0 + r, w, hr[r], wr[r]);
    }
    // Merge flags, one bit per row (pixel-only, hoisted out of the channel loop):
    //   vm: this pixel's top taps are the previous row's bottom taps (same thread) -> the pending pair is folded in
    //   hl: this pixel's left taps are the left lane's right taps -> that lane's right column is folded in here
    //   hr: the right lane folds this lane's right column -> this lane does not emit it
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
    const float* td = L.topdiff + cbase + (size_t)h0 * W + w;
    float* bd = L.out0 + cbase;
    float gx[kFlowRows], gy[kFlowRows];
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r) gx[r] = gy[r] = 0.f;

    for (int c = c0; c < c1; ++c, base += plane, td += plane, bd += plane) {
        float a1[kFlowRows], a2[kFlowRows], a3[kFlowRows], a4[kFlowRows];
#pragma unroll
        for (int r = 0; r < kFlowRows; ++r) {
            a1[r] = a2[r] = a3[r] = a4[r] = 0.f;
            if (off[r] >= 0) {
                const float* p = base + off[r];
                const float f1 = __ldg(p), f2 = __ldg(p + 1), f3 = __ldg(p + W), f4 = __ldg(p + W + 1);
                const float g = __ldcs(td + r * W);
                const float h_ratio = hr[r], w_ratio = wr[r];
                // flow_align_cuda_kernel.cu:89-92: the values handed to atomicAdd(float*, float)
                a1[r] = g * (1. - h_ratio) * (1. - w_ratio);
                a2[r] = g * (1. - h_ratio) * (w_ratio);
                a3[r] = g * (h_ratio) * (1. - w_ratio);
                a4[r] = g * (h_ratio) * (w_ratio);
                // :95-110
                const float dx = -f1 * (1. - h_ratio) + f2 * (1. - h_ratio) - f3 * (h_ratio) + f4 * (h_ratio);
                const float dy = -f1 * (1. - w_ratio) - f2 * (w_ratio) + f3 * (1. - w_ratio) + f4 * (w_ratio);
                gx[r] += __fmul_rn(g, dx);
                gy[r] += __fmul_rn(g, dy);
            }
        }
        // emit: top taps of row r (+ the pending bottom taps of row r-1 when they coincide), bottom taps of the
        // last row; horizontally, a lane's right column goes to the right lane when that lane's left column is
        // the same texel column.
        float pl = 0.f, pr = 0.f;     // pending bottom pair of the previous row
#pragma unroll
        for (int r = 0; r <= kFlowRows; ++r) {
            float tl, tr;
            int o;
            unsigned bit;
            if (r < kFlowRows) {
                if (r > 0 && !((vm >> r) & 1u) && off[r - 1] >= 0) {   // previous row's bottom pair stands alone
                    red_add(bd + off[r - 1] + W, pl);
                    red_add(bd + off[r - 1] + W + 1, pr);
                    pl = pr = 0.f;
                }
                tl = a1[r] + (((vm >> r) & 1u) ? pl : 0.f);
                tr = a2[r] + (((vm >> r) & 1u) ? pr : 0.f);
                o = off[r];
                bit = r;
                pl = a3[r];
                pr = a4[r];
            } else {                                                   // bottom pair of the thread's last row
                tl = pl;
                tr = pr;
                o = off[kFlowRows - 1] >= 0 ? off[kFlowRows - 1] + W : -1;
                bit = kFlowRows - 1;
            }
            const float from_left = __shfl_up_sync(0xffffffffu, tr, 1);
            if ((hl >> bit) & 1u) tl += from_left;
            if (o >= 0) {
                red_add(bd + o, tl);
                if (!((hrm >> bit) & 1u)) red_add(bd + o + 1, tr);
            }
        }
    }
    // flow gradient: one reduction per (pixel, channel chunk) instead of one per channel (:111-112)
    float* fd = L.out1 + (size_t)t.n * 2 * plane + (size_t)h0 * W + w;
#pragma unroll
    for (int r = 0; r < kFlowRows; ++r)
        if (off[r] >= 0) {
            red_add(fd + r * W, gx[r]);
            red_add(fd + plane + r * W, gy[r]);
        }
}


std::atomic<int> g_flow_fast{0};

// Fills the block ranges; returns the grid size (0: nothing to do) or a negative status.
long long plan(FlowArgs& a, int num_levels, int batches, int channels, const int* level_h, const int* level_w) {
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
        a.lv[l].tiles_x = ceil_div(level_w[l], 32);
        a.lv[l].tiles_y = ceil_div(level_h[l], kFlowWarps * kFlowRows);
        tiles += (long long)a.lv[l].tiles_x * a.lv[l].tiles_y;
    }
    if (tiles == 0 || batches == 0 || channels == 0) return 0;
    // channel chunk per thread: long enough to amortise the per-pixel geometry, short enough to fill 148 SMs
    int chunk = 32;
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
    const long long grid = plan(a, num_levels, batches, channels, level_h, level_w);
    if (grid <= 0) return (int)grid;
    for (int l = 0; l < num_levels; ++l) {
        if (a.lv[l].H * a.lv[l].W && (!bottom[l] || !flow[l] || !top[l])) return VOSD_ERR_BAD_ARG;
        a.lv[l].bottom = bottom[l];
        a.lv[l].flow = flow[l];
        a.lv[l].out0 = top[l];
    }
    if (g_flow_fast.load(std::memory_order_relaxed))
        flow_align_fwd_kernel<false><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    else
        flow_align_fwd_kernel<true><<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
    count_launch();
    return check_launch();
}

int flow_bwd(int num_levels, int batches, int channels, const int* level_h, const int* level_w,
             const float* const* topdiff, const float* const* bottom, const float* const* flow,
             float* const* bottomdiff, float* const* flowdiff, int zero_init, cudaStream_t stream) {
    if (num_levels > 0 && (!level_h || !level_w || !topdiff || !bottom || !flow || !bottomdiff || !flowdiff))
        return VOSD_ERR_BAD_ARG;
    FlowArgs a = {};
    const long long grid = plan(a, num_levels, batches, channels, level_h, level_w);
    if (grid < 0) return (int)grid;
    for (int l = 0; l < num_levels; ++l) {
        const size_t plane = (size_t)level_h[l] * level_w[l];
        if (plane && batches && (!flowdiff[l] || (channels && (!topdiff[l] || !bottom[l] || !flow[l] || !bottomdiff[l]))))
            return VOSD_ERR_BAD_ARG;
        a.lv[l].topdiff = topdiff[l];
        a.lv[l].bottom = bottom[l];
        a.lv[l].flow = flow[l];
        a.lv[l].out0 = bottomdiff[l];
        a.lv[l].out1 = flowdiff[l];
        if (zero_init && plane && batches) {
            // FlowAlignFunction.backward zero-fills both gradients itself (functions/flow_align.py:41-43)
            if (channels && cudaMemsetAsync(bottomdiff[l], 0, plane * batches * channels * sizeof(float), stream) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
            if (cudaMemsetAsync(flowdiff[l], 0, plane * batches * 2 * sizeof(float), stream) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
        }
    }
    if (grid == 0) return VOSD_OK;
    flow_align_bwd_kernel<<<(unsigned)grid, kFlowThreads, 0, stream>>>(a);
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
    return vosd::g_flow_fast.exchange(on ? 1 : 0, std::memory_order_relaxed);
}
