// RoIAlign backward on channels-last gradient maps (memory order N, H, W, C): the twin of roialign_nhwc.cuh, so that a
// channels-last backbone trains through this library without a layout conversion.  Reference semantics:
// ROIAlignBackward, roi_align_kernel.cu:195-270 (every valid sample scatters its four bilinear taps of top_diff / 4).
//
// Formulation (as the separable NCHW backward, roialign_sep.cuh): the gradient of texel (y, x) is
//     sum_p sum_q  wy[y][p] * wx[x][q] * top_diff[p][q],     wy / wx = 0.5 * (sum of the axis weights with which the
//                                                             two samples of bin p / q touch row y / column x),
// so ONE reduction per (texel, channel) of the footprint instead of sixteen per bin.  In this layout the 32 channels
// of a slab are contiguous: lanes = channels, a warp's `red.global.add.f32` of one texel is one 128-byte transaction.
//   CTA = one 7 x 7 block of output bins of one RoI (14 x 14 and 28 x 28 heads: 4 / 16 blocks), 8 warps = 8 channel
//   slabs at a time.  The two weight tables are built once per CTA in shared memory (warp-uniform reads); a lane keeps
//   its channel's 49 top_diff values in registers, forms V[q] = sum_p wy[y][p] * top[p][q] per texel row (rows that do
//   not enter are skipped: warp-uniform test) and g = sum_q wx[x][q] * V[q] per texel (dense over the 7 columns: no
//   dynamic register indexing).
// Footprints beyond 64 rows or columns per block (RoIs far larger than their level's canonical size) take the direct
// path of the same kernel: sixteen reductions per bin, any size.
// Float behaviour: sums are re-associated (table products, atomics), parity is the backward gate of the tests; a
// zero-weight column still multiplies V, so a non-finite top_diff value spreads over its block's footprint rows.
#pragma once
#include "common.cuh"
#include "roialign_math.cuh"

namespace vosd {

constexpr int kNbWarps = 8;
constexpr int kNbMaxT = 64;           // table rows: texel rows / columns of one block's footprint

struct NbShared {
    float wy[kNbMaxT][8];             // [texel row - y_lo][bin row of the block], column 7 unused
    float wx[kNbMaxT][8];
    int y_lo, y_hi, x_lo, x_hi;       // inclusive texel ranges touched by valid samples (lo > hi: none)
};

__global__ void __launch_bounds__(32 * kNbWarps)
roialign_bwd_nhwc(const __grid_constant__ LevelTable t, int channels, int pooled_h, int pooled_w, int blocks_w,
                  const float* __restrict__ rois, const int* __restrict__ roi_level, const int* __restrict__ out_index,
                  const float* __restrict__ top_diff) {
    __shared__ NbShared sh;
    const int r = blockIdx.x;
    const int bp = blockIdx.y / blocks_w, bq = blockIdx.y - bp * blocks_w;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int level = roi_level ? __ldg(roi_level + r) : 0;
    const int H = t.h[level], W = t.w[level];
    const RoiGeom g = roi_geometry(rois + 5 * (size_t)r, t.scale[level], pooled_h, pooled_w, 2);
    const int row = out_index ? __ldg(out_index + r) : r;
    const int bins = pooled_h * pooled_w;

    // ---- footprint of the block and the two weight tables (warp 0: lanes 0-6 rows, lanes 8-14 columns)
    for (int i = threadIdx.x; i < 2 * kNbMaxT * 8; i += blockDim.x) (&sh.wy[0][0])[i] = 0.f;
    AxisTap tp[2];
    tp[0].valid = tp[1].valid = 0;
    const bool is_y = lane < 7, is_x = lane >= 8 && lane < 15;
    const int b = is_y ? lane : lane - 8;
    if (warp == 0) {
        int lo = 0x7fffffff, hi = -1;
        if (is_y || is_x) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                tp[i] = is_y ? axis_tap(sample_coord(g.start_h, g.bin_h, 7 * bp + b, i, 2), H)
                             : axis_tap(sample_coord(g.start_w, g.bin_w, 7 * bq + b, i, 2), W);
                if (tp[i].valid) { lo = min(lo, tp[i].low); hi = max(hi, tp[i].high); }
            }
        }
        // min / max over the y lanes (0-7) and the x lanes (8-15) separately
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { sh.y_lo = lo; sh.y_hi = hi; }
        if (lane == 8) { sh.x_lo = lo; sh.x_hi = hi; }
    }
    __syncthreads();
    const int y_lo = sh.y_lo, y_hi = sh.y_hi, x_lo = sh.x_lo, x_hi = sh.x_hi;
    if (y_lo > y_hi || x_lo > x_hi) return;                 // no valid sample: the block adds nothing
    const int th = y_hi - y_lo + 1, tw = x_hi - x_lo + 1;
    const bool tables = th <= kNbMaxT && tw <= kNbMaxT;
    if (tables && warp == 0 && (is_y || is_x)) {
        float (*tab)[8] = is_y ? sh.wy : sh.wx;
        const int base = is_y ? y_lo : x_lo;
#pragma unroll
        for (int i = 0; i < 2; i++)
            if (tp[i].valid) {                               // one lane per table column: no write conflicts
                tab[tp[i].low - base][b] += 0.5f * tp[i].h;
                tab[tp[i].high - base][b] += 0.5f * tp[i].l;
            }
    }
    __syncthreads();

    float* __restrict__ G = t.data[level] + (size_t)g.batch * H * W * channels;
    const int slabs = channels / 32;
    for (int s = warp; s < slabs; s += kNbWarps) {
        const int c = s * 32 + lane;
        const float* __restrict__ td = top_diff + ((size_t)row * channels + c) * bins + (7 * bp) * pooled_w + 7 * bq;
        float top[49];
#pragma unroll
        for (int p = 0; p < 7; p++)
#pragma unroll
            for (int q = 0; q < 7; q++) top[7 * p + q] = __ldg(td + p * pooled_w + q);
        if (tables) {
            for (int y = 0; y < th; y++) {
                const float4 a = *reinterpret_cast<const float4*>(&sh.wy[y][0]);
                const float4 bb = *reinterpret_cast<const float4*>(&sh.wy[y][4]);
                const float w[7] = {a.x, a.y, a.z, a.w, bb.x, bb.y, bb.z};
                float V[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                bool any = false;
#pragma unroll
                for (int p = 0; p < 7; p++)
                    if (w[p] != 0.f) {                       // warp-uniform
                        any = true;
#pragma unroll
                        for (int q = 0; q < 7; q++) V[q] = fmaf(w[p], top[7 * p + q], V[q]);
                    }
                if (!any) continue;
                float* __restrict__ grow = G + ((size_t)(y_lo + y) * W + x_lo) * channels + c;
                for (int x = 0; x < tw; x++) {
                    const float4 u = *reinterpret_cast<const float4*>(&sh.wx[x][0]);
                    const float4 v = *reinterpret_cast<const float4*>(&sh.wx[x][4]);
                    if ((u.x != 0.f) | (u.y != 0.f) | (u.z != 0.f) | (u.w != 0.f) | (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f)) {
                        const float gsum = fmaf(v.z, V[6], fmaf(v.y, V[5], fmaf(v.x, V[4], fmaf(u.w, V[3],
                                           fmaf(u.z, V[2], fmaf(u.y, V[1], u.x * V[0]))))));
                        red_add_f32(grow + (size_t)x * channels, gsum);
                    }
                }
            }
        } else {
            // direct path: every valid sample scatters its four taps (any footprint size)
            for (int p = 0; p < 7; p++)
                for (int iy = 0; iy < 2; iy++) {
                    const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, 7 * bp + p, iy, 2), H);
                    if (!ty.valid) continue;
                    for (int q = 0; q < 7; q++) {
                        float v = 0.f;                       // top[7 * p + q] without dynamic register indexing
#pragma unroll
                        for (int k = 0; k < 49; k++) v = (k == 7 * p + q) ? top[k] : v;
                        v *= 0.25f;
                        for (int ix = 0; ix < 2; ix++) {
                            const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, 7 * bq + q, ix, 2), W);
                            if (!tx.valid) continue;
                            red_add_f32(G + ((size_t)ty.low * W + tx.low) * channels + c, ty.h * tx.h * v);
                            red_add_f32(G + ((size_t)ty.low * W + tx.high) * channels + c, ty.h * tx.l * v);
                            red_add_f32(G + ((size_t)ty.high * W + tx.low) * channels + c, ty.l * tx.h * v);
                            red_add_f32(G + ((size_t)ty.high * W + tx.high) * channels + c, ty.l * tx.l * v);
                        }
                    }
                }
        }
    }
}

}  // namespace vosd
