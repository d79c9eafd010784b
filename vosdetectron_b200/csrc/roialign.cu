// Multi-level RoIAlign forward / backward for sm_100a.
// Reference: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:65-121 (fwd), :195-270 (bwd),
// driven per FPN level by lib/modeling/model_builder.py:262-303.
#include "common.cuh"
#include "roialign_math.cuh"

namespace vosd {

struct LevelTable {
    float* data[VOSD_MAX_LEVELS];       // fwd: feature maps (read); bwd: gradient maps (accumulated)
    int h[VOSD_MAX_LEVELS];
    int w[VOSD_MAX_LEVELS];
    float scale[VOSD_MAX_LEVELS];
};

// ---------------------------------------------------------------------------------------
// Generic gather path: one thread per output element, no staging.  Handles every shape the
// reference accepts (adaptive sampling grid, RoIs larger than any tile budget).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
roialign_fwd_generic(LevelTable lv, int channels, int pooled_h, int pooled_w, int sampling_ratio,
                     long long total, const float* __restrict__ rois,
                     const int* __restrict__ roi_level, const int* __restrict__ out_index,
                     float* __restrict__ top) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int pw = (int)(idx % pooled_w);
        const int ph = (int)((idx / pooled_w) % pooled_h);
        const int c = (int)((idx / pooled_w / pooled_h) % channels);
        const int n = (int)(idx / pooled_w / pooled_h / channels);
        const int l = roi_level ? roi_level[n] : 0;
        const int H = lv.h[l], W = lv.w[l];
        const RoiGeom g = roi_geometry(rois + 5 * (size_t)n, lv.scale[l], pooled_h, pooled_w, sampling_ratio);
        const float* __restrict__ d = lv.data[l] + ((size_t)g.batch * channels + c) * H * W;
        float acc = 0.f;
        for (int iy = 0; iy < g.grid_h; iy++) {
            const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h), H);
            for (int ix = 0; ix < g.grid_w; ix++) {
                const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w), W);
                float val = 0.f;
                if (ty.valid && tx.valid) {
                    val = bilinear_value(ty.h, ty.l, tx.h, tx.l,
                                         __ldg(d + ty.low * W + tx.low), __ldg(d + ty.low * W + tx.high),
                                         __ldg(d + ty.high * W + tx.low), __ldg(d + ty.high * W + tx.high));
                }
                acc = __fadd_rn(acc, val);
            }
        }
        const int row = out_index ? out_index[n] : n;
        top[(((size_t)row * channels + c) * pooled_h + ph) * pooled_w + pw] = __fdiv_rn(acc, g.count);
    }
}

__global__ void __launch_bounds__(256)
roialign_bwd_generic(LevelTable lv, int channels, int pooled_h, int pooled_w, int sampling_ratio,
                     long long total, const float* __restrict__ rois,
                     const int* __restrict__ roi_level, const int* __restrict__ out_index,
                     const float* __restrict__ top_diff) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int pw = (int)(idx % pooled_w);
        const int ph = (int)((idx / pooled_w) % pooled_h);
        const int c = (int)((idx / pooled_w / pooled_h) % channels);
        const int n = (int)(idx / pooled_w / pooled_h / channels);
        const int l = roi_level ? roi_level[n] : 0;
        const int H = lv.h[l], W = lv.w[l];
        const RoiGeom g = roi_geometry(rois + 5 * (size_t)n, lv.scale[l], pooled_h, pooled_w, sampling_ratio);
        float* d = lv.data[l] + ((size_t)g.batch * channels + c) * H * W;
        const int row = out_index ? out_index[n] : n;
        const float t = top_diff[(((size_t)row * channels + c) * pooled_h + ph) * pooled_w + pw];
        for (int iy = 0; iy < g.grid_h; iy++) {
            const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h), H);
            for (int ix = 0; ix < g.grid_w; ix++) {
                const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w), W);
                if (!(ty.valid && tx.valid)) continue;
                const float g1 = __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.h)), g.count);
                const float g2 = __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.l)), g.count);
                const float g3 = __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.h)), g.count);
                const float g4 = __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.l)), g.count);
                atomicAdd(d + ty.low * W + tx.low, g1);
                atomicAdd(d + ty.low * W + tx.high, g2);
                atomicAdd(d + ty.high * W + tx.low, g3);
                atomicAdd(d + ty.high * W + tx.high, g4);
            }
        }
    }
}

static int fill_table(LevelTable& t, const float* const* data, const int* h, const int* w,
                      const float* scale, int num_levels) {
    if (num_levels < 1 || num_levels > VOSD_MAX_LEVELS) return VOSD_ERR_UNSUPPORTED;
    if (!data || !h || !w || !scale) return VOSD_ERR_BAD_ARG;
    for (int l = 0; l < num_levels; l++) {
        if (!data[l]) return VOSD_ERR_BAD_ARG;
        if (h[l] <= 0 || w[l] <= 0) return VOSD_ERR_BAD_SHAPE;
        t.data[l] = const_cast<float*>(data[l]);
        t.h[l] = h[l]; t.w[l] = w[l]; t.scale[l] = scale[l];
    }
    return VOSD_OK;
}

static int grid_for(long long total, int block) {
    long long b = (total + block - 1) / block;
    const long long cap = (long long)kNumSMs * 64;
    return (int)(b < cap ? b : cap);
}

static int ml_fwd(const LevelTable& t, int channels, int ph, int pw, int sr, int num_rois,
                  const float* rois, const int* roi_level, const int* out_index, float* top,
                  cudaStream_t stream) {
    if (channels <= 0 || ph <= 0 || pw <= 0 || num_rois < 0) return VOSD_ERR_BAD_SHAPE;
    if (num_rois == 0) return VOSD_OK;
    if (!rois || !top) return VOSD_ERR_BAD_ARG;
    const long long total = (long long)num_rois * channels * ph * pw;
    roialign_fwd_generic<<<grid_for(total, 256), 256, 0, stream>>>(
        t, channels, ph, pw, sr, total, rois, roi_level, out_index, top);
    count_launch();
    return check_launch();
}

static int ml_bwd(const LevelTable& t, int num_levels, int batch, int channels, int ph, int pw, int sr,
                  int num_rois, const float* rois, const int* roi_level, const int* out_index,
                  const float* top_diff, int zero_init, cudaStream_t stream) {
    if (channels <= 0 || ph <= 0 || pw <= 0 || num_rois < 0 || batch <= 0) return VOSD_ERR_BAD_SHAPE;
    if (zero_init) {
        for (int l = 0; l < num_levels; l++)
            if (cudaMemsetAsync(t.data[l], 0, sizeof(float) * (size_t)batch * channels * t.h[l] * t.w[l],
                                stream) != cudaSuccess) return VOSD_ERR_LAUNCH;
    }
    if (num_rois == 0) return VOSD_OK;
    if (!rois || !top_diff) return VOSD_ERR_BAD_ARG;
    const long long total = (long long)num_rois * channels * ph * pw;
    roialign_bwd_generic<<<grid_for(total, 256), 256, 0, stream>>>(
        t, channels, ph, pw, sr, total, rois, roi_level, out_index, top_diff);
    count_launch();
    return check_launch();
}

}  // namespace vosd

using namespace vosd;

extern "C" int vosd_roialign_fwd(const float* bottom_data, float spatial_scale, int num_rois,
                                 int height, int width, int channels,
                                 int aligned_height, int aligned_width, int sampling_ratio,
                                 const float* bottom_rois, float* top_data, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, &bottom_data, &height, &width, &spatial_scale, 1);
    if (rc) return rc;
    return ml_fwd(t, channels, aligned_height, aligned_width, sampling_ratio, num_rois, bottom_rois,
                  nullptr, nullptr, top_data, stream);
}

extern "C" int vosd_roialign_bwd(const float* top_diff, float spatial_scale, int batch_size, int num_rois,
                                 int height, int width, int channels,
                                 int aligned_height, int aligned_width, int sampling_ratio,
                                 const float* bottom_rois, float* bottom_diff, int zero_init,
                                 cudaStream_t stream) {
    LevelTable t;
    const float* p = bottom_diff;
    int rc = fill_table(t, &p, &height, &width, &spatial_scale, 1);
    if (rc) return rc;
    return ml_bwd(t, 1, batch_size, channels, aligned_height, aligned_width, sampling_ratio, num_rois,
                  bottom_rois, nullptr, nullptr, top_diff, zero_init, stream);
}

extern "C" int vosd_roialign_ml_fwd(const float* const* level_data, const int* level_h, const int* level_w,
                                    const float* level_scale, int num_levels, int channels,
                                    int aligned_height, int aligned_width, int sampling_ratio,
                                    int num_rois, const float* rois, const int* roi_level,
                                    const int* out_index, float* top_data, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, level_data, level_h, level_w, level_scale, num_levels);
    if (rc) return rc;
    if (num_levels > 1 && !roi_level && num_rois > 0) return VOSD_ERR_BAD_ARG;
    return ml_fwd(t, channels, aligned_height, aligned_width, sampling_ratio, num_rois, rois,
                  roi_level, out_index, top_data, stream);
}

extern "C" int vosd_roialign_ml_bwd(const float* top_diff, float* const* level_diff, const int* level_h,
                                    const int* level_w, const float* level_scale, int num_levels,
                                    int batch_size, int channels,
                                    int aligned_height, int aligned_width, int sampling_ratio,
                                    int num_rois, const float* rois, const int* roi_level,
                                    const int* out_index, int zero_init, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, (const float* const*)level_diff, level_h, level_w, level_scale, num_levels);
    if (rc) return rc;
    if (num_levels > 1 && !roi_level && num_rois > 0) return VOSD_ERR_BAD_ARG;
    return ml_bwd(t, num_levels, batch_size, channels, aligned_height, aligned_width, sampling_ratio,
                  num_rois, rois, roi_level, out_index, top_diff, zero_init, stream);
}
