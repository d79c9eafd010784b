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


// ---------------------------------------------------------------------------------------
// Staged forward path.  One CTA = one RoI x one slab of 32 channels.
//   1. the first threads build per-axis tap tables (low/high texel, weights, validity) for the
//      PH*grid_h sample rows and PW*grid_w sample columns -- computed once per CTA instead of
//      once per output element;
//   2. the RoI's bounding tile of the feature map (rows y_lo..y_hi, columns x_lo..x_hi of the
//      32 channel planes) is staged in shared memory as tile[c][row*tw + col] with an ODD
//      per-channel stride, in bands of output rows when the whole tile does not fit;
//   3. lanes own channels, warps own output bins: every tap is one shared-memory wavefront
//      (bank = (c*stride + texel) mod 32, stride odd -> conflict-free), the weights are
//      warp-uniform, and the per-element arithmetic is the reference's (roialign_math.cuh);
//   4. results go through obuf[c][bin] (odd stride) so the slab leaves as ONE contiguous
//      32*PH*PW*4-byte streaming write (evict-first: keeps the feature maps in L2).
// RoIs whose tables or single-row tile exceed the budgets fall back to direct global gathers
// inside the same kernel (no CPU path).
// ---------------------------------------------------------------------------------------
constexpr int kSlab = 32;
constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kMaxTaps = 64;                 // per-axis table capacity (PH*grid_h, PW*grid_w)

struct __align__(16) Tap { int low, high; float l, h; };   // low < 0: sample outside the map

struct FwdShared {
    Tap ytab[kMaxTaps];
    Tap xtab[kMaxTaps];
    RoiGeom g;
    int level, H, W;
    int x_lo, tw;            // tile columns
    int ok;                  // 0 -> generic fallback for this RoI
};

__device__ __forceinline__ float ld_feat(const float* p) { return __ldg(p); }

__global__ void __launch_bounds__(kFwdThreads)
roialign_fwd_staged(const __grid_constant__ LevelTable lv, int channels, int pooled_h, int pooled_w,
                    int sampling_ratio, int tile_cap /* odd, words per channel */, int obuf_stride /* odd */,
                    const float* __restrict__ rois, const int* __restrict__ roi_level,
                    const int* __restrict__ out_index, float* __restrict__ top) {
    extern __shared__ __align__(16) float dyn[];
    float* tile = dyn;                                   // [kSlab][tile_cap]
    float* obuf = dyn + (size_t)kSlab * tile_cap;        // [kSlab][obuf_stride]
    __shared__ FwdShared sh;

    const int n = blockIdx.x;
    const int c0 = blockIdx.y * kSlab;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nch = min(kSlab, channels - c0);
    const int bins = pooled_h * pooled_w;

    if (tid == 0) {
        const int l = roi_level ? roi_level[n] : 0;
        sh.level = l; sh.H = lv.h[l]; sh.W = lv.w[l];
        sh.g = roi_geometry(rois + 5 * (size_t)n, lv.scale[l], pooled_h, pooled_w, sampling_ratio);
    }
    __syncthreads();
    const RoiGeom g = sh.g;
    const int H = sh.H, W = sh.W;
    const int ny = pooled_h * g.grid_h, nx = pooled_w * g.grid_w;
    const bool tables_fit = ny <= kMaxTaps && nx <= kMaxTaps;
    if (tables_fit) {
        if (tid < ny) {
            const AxisTap t = axis_tap(sample_coord(g.start_h, g.bin_h, tid / g.grid_h, tid % g.grid_h, g.grid_h), H);
            sh.ytab[tid] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        } else if (tid >= 64 && tid < 64 + nx) {
            const int k = tid - 64;
            const AxisTap t = axis_tap(sample_coord(g.start_w, g.bin_w, k / g.grid_w, k % g.grid_w, g.grid_w), W);
            sh.xtab[k] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        }
    }
    __syncthreads();
    if (tid == 0) {
        int x_lo = 1 << 30, x_hi = -1;
        if (tables_fit)
            for (int k = 0; k < nx; k++)
                if (sh.xtab[k].low >= 0) { x_lo = min(x_lo, sh.xtab[k].low); x_hi = max(x_hi, sh.xtab[k].high); }
        sh.x_lo = x_lo; sh.tw = x_hi - x_lo + 1;          // tw <= 0: no valid column at all
        // every single output row must fit the tile on its own
        int ok = tables_fit;
        if (ok && sh.tw > 0) {
            for (int ph = 0; ph < pooled_h && ok; ph++) {
                int lo = 1 << 30, hi = -1;
                for (int i = 0; i < g.grid_h; i++) {
                    const Tap t = sh.ytab[ph * g.grid_h + i];
                    if (t.low >= 0) { lo = min(lo, t.low); hi = max(hi, t.high); }
                }
                if (hi >= lo && (hi - lo + 1) * sh.tw > tile_cap) ok = 0;
            }
        }
        sh.ok = ok;
    }
    __syncthreads();
    const float* __restrict__ feat = lv.data[sh.level] + ((size_t)g.batch * channels + c0) * H * W;
    const int row = out_index ? out_index[n] : n;
    float* __restrict__ out = top + ((size_t)row * channels + c0) * bins;

    if (!sh.ok) {
        // direct-gather fallback: thread per (channel, bin) of this slab
        for (int e = tid; e < nch * bins; e += kFwdThreads) {
            const int c = e / bins, b = e - c * bins;
            const int ph = b / pooled_w, pw = b - ph * pooled_w;
            const float* d = feat + (size_t)c * H * W;
            float acc = 0.f;
            for (int iy = 0; iy < g.grid_h; iy++) {
                const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, g.grid_h), H);
                for (int ix = 0; ix < g.grid_w; ix++) {
                    const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, g.grid_w), W);
                    float val = 0.f;
                    if (ty.valid && tx.valid)
                        val = bilinear_value(ty.h, ty.l, tx.h, tx.l, __ldg(d + ty.low * W + tx.low),
                                             __ldg(d + ty.low * W + tx.high), __ldg(d + ty.high * W + tx.low),
                                             __ldg(d + ty.high * W + tx.high));
                    acc = __fadd_rn(acc, val);
                }
            }
            out[e] = __fdiv_rn(acc, g.count);
        }
        return;
    }

    const int x_lo = sh.x_lo, tw = sh.tw;
    const float inv_tw = tw > 0 ? 1.0f / (float)tw : 0.f;
    int p0 = 0;
    while (p0 < pooled_h) {
        // ---- band [p0, p1): as many output rows as fit the tile (uniform across the CTA) ----
        int p1 = p0, y_lo = 1 << 30, y_hi = -1;
        while (p1 < pooled_h) {
            int lo = y_lo, hi = y_hi;
            for (int i = 0; i < g.grid_h; i++) {
                const Tap t = sh.ytab[p1 * g.grid_h + i];
                if (t.low >= 0) { lo = min(lo, t.low); hi = max(hi, t.high); }
            }
            if (hi >= lo && tw > 0 && (hi - lo + 1) * tw > tile_cap) break;
            y_lo = lo; y_hi = hi; p1++;
        }
        const int rows = (y_hi >= y_lo && tw > 0) ? y_hi - y_lo + 1 : 0;
        const int elems = rows * tw;
        // ---- stage: warp w copies channels w, w+8, ... ; lanes run over the flattened tile ----
        for (int i0 = 0; i0 < elems; i0 += 32) {
            const int i = i0 + lane;
            if (i < elems) {
                const int ry = (int)(((float)i + 0.5f) * inv_tw);
                const int rx = i - ry * tw;
                const float* src = feat + (size_t)(y_lo + ry) * W + (x_lo + rx);
#pragma unroll
                for (int k = 0; k < kSlab / kFwdWarps; k++) {
                    const int c = warp + k * kFwdWarps;
                    if (c < nch) tile[c * tile_cap + i] = ld_feat(src + (size_t)c * H * W);
                }
            }
        }
        __syncthreads();
        // ---- gather: lanes = channels, warps = bins ----
        const int nb = (p1 - p0) * pooled_w;
        const float* tl = tile + lane * tile_cap;
        for (int q = warp; q < nb; q += kFwdWarps) {
            const int pr = q / pooled_w;
            const int ph = p0 + pr, pw = q - pr * pooled_w;
            float acc = 0.f;
            for (int iy = 0; iy < g.grid_h; iy++) {
                const Tap ty = sh.ytab[ph * g.grid_h + iy];
                const int r0 = (ty.low - y_lo) * tw - x_lo, r1 = (ty.high - y_lo) * tw - x_lo;
                for (int ix = 0; ix < g.grid_w; ix++) {
                    const Tap tx = sh.xtab[pw * g.grid_w + ix];
                    float val = 0.f;
                    if (ty.low >= 0 && tx.low >= 0)
                        val = bilinear_value(ty.h, ty.l, tx.h, tx.l, tl[r0 + tx.low], tl[r0 + tx.high],
                                             tl[r1 + tx.low], tl[r1 + tx.high]);
                    acc = __fadd_rn(acc, val);
                }
            }
            obuf[lane * obuf_stride + ph * pooled_w + pw] = __fdiv_rn(acc, g.count);
        }
        __syncthreads();
        p0 = p1;
    }
    // ---- one contiguous streaming write of the slab ----
    for (int e = tid; e < nch * bins; e += kFwdThreads) {
        const int c = e / bins, b = e - c * bins;
        __stcs(out + e, obuf[c * obuf_stride + b]);
    }
}

static int g_force_generic = 0;   // test hook (vosd_debug_force_generic)

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
    // staged path: shared-memory budget per CTA chosen so that 3 (small outputs) or 2 CTAs fit an SM
    const int bins = ph * pw;
    const int obuf_stride = bins | 1;
    const int budget = (bins <= 64 ? 74 : 112) * 1024 - (int)sizeof(FwdShared) - 64;
    int tile_cap = (budget - kSlab * obuf_stride * (int)sizeof(float)) / (kSlab * (int)sizeof(float));
    tile_cap = (tile_cap - 1) | 1;
    if (tile_cap >= 65 && !g_force_generic) {
        const size_t dyn = (size_t)kSlab * (tile_cap + obuf_stride) * sizeof(float);
        if (cudaFuncSetAttribute(roialign_fwd_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
            return VOSD_ERR_LAUNCH;
        dim3 grid(num_rois, ceil_div(channels, kSlab));
        roialign_fwd_staged<<<grid, kFwdThreads, dyn, stream>>>(t, channels, ph, pw, sr, tile_cap, obuf_stride,
                                                                rois, roi_level, out_index, top);
    } else {
        const long long total = (long long)num_rois * channels * ph * pw;
        roialign_fwd_generic<<<grid_for(total, 256), 256, 0, stream>>>(
            t, channels, ph, pw, sr, total, rois, roi_level, out_index, top);
    }
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

// Test hook: route RoIAlign through the generic (un-staged) kernels so both paths stay covered.
extern "C" VOSD_API int vosd_debug_force_generic(int on) {
    const int old = g_force_generic;
    g_force_generic = on;
    return old;
}
