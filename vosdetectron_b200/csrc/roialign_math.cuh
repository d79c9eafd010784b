// Per-element arithmetic of Caffe2-style RoIAlign, pinned operation by operation to what
// nvcc 12.9 emits for the reference kernel
// (lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:16-121,150-270) when that file
// is built for sm_100a with its default -fmad=true.  Read from the SASS of
// oracle/_ref/libref_roialign.so:
//   roi_start   = roi[k] * scale                                  FMUL
//   roi_width   = fmaxf(fma(roi[3], scale, -roi_start_w), 1)      FFMA + FMNMX
//   bin_size    = roi_width / pooled                              IEEE div.rn
//   y           = fma(ph, bin_h, start_h) + ((iy + .5f) * bin_h) / grid_h
//   hy          = 1 - ly  (the double `1.` of the source rounds to the same fp32 value)
//   val         = fma(w4, v4, fma(w3, v3, fma(w1, v1, w2 * v2)))
//   out         = (sum over iy, ix in that order) / (grid_h * grid_w)
//   grad addend = (top * w_k) / count
// Using the _rn intrinsics keeps the compiler from re-associating or contracting anything,
// so staging data through shared memory cannot change a single bit of the result.
#pragma once
#include <cuda_runtime.h>

namespace vosd {

struct RoiGeom {
    int batch;
    float start_w, start_h, bin_w, bin_h;
    int grid_h, grid_w;
    float count;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ r, float scale,
                                                int pooled_h, int pooled_w, int sampling_ratio) {
    RoiGeom g;
    g.batch = (int)r[0];
    g.start_w = __fmul_rn(r[1], scale);
    g.start_h = __fmul_rn(r[2], scale);
    const float roi_w = fmaxf(__fmaf_rn(r[3], scale, -g.start_w), 1.f);
    const float roi_h = fmaxf(__fmaf_rn(r[4], scale, -g.start_h), 1.f);
    g.bin_h = __fdiv_rn(roi_h, (float)pooled_h);
    g.bin_w = __fdiv_rn(roi_w, (float)pooled_w);
    g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bin_h);
    g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(g.bin_w);
    g.count = (float)(g.grid_h * g.grid_w);
    return g;
}

// Sample coordinate along one axis: bin index p, sub-sample i of `grid`.
__device__ __forceinline__ float sample_coord(float start, float bin, int p, int i, int grid) {
    const float base = __fmaf_rn((float)p, bin, start);
    const float off = __fdiv_rn(__fmul_rn((float)i + .5f, bin), (float)grid);
    return __fadd_rn(base, off);
}

struct AxisTap {
    int low, high;   // texel indices (clamped)
    float l, h;      // weights of high / low texel
    int valid;       // 0: sample lies outside [-1, size] -> contributes 0
};

__device__ __forceinline__ AxisTap axis_tap(float v, int size) {
    AxisTap t;
    t.valid = !(v < -1.0f || v > (float)size);
    if (v <= 0.f) v = 0.f;
    t.low = (int)v;
    if (t.low >= size - 1) {
        t.high = t.low = size - 1;
        v = (float)t.low;
    } else {
        t.high = t.low + 1;
    }
    t.l = __fsub_rn(v, (float)t.low);
    t.h = __fsub_rn(1.f, t.l);
    return t;
}

__device__ __forceinline__ float bilinear_value(float hy, float ly, float hx, float lx,
                                                float v1, float v2, float v3, float v4) {
    const float w1 = __fmul_rn(hy, hx), w2 = __fmul_rn(hy, lx);
    const float w3 = __fmul_rn(ly, hx), w4 = __fmul_rn(ly, lx);
    return __fmaf_rn(w4, v4, __fmaf_rn(w3, v3, __fmaf_rn(w1, v1, __fmul_rn(w2, v2))));
}

}  // namespace vosd
