// collect + FPN level assignment + distribute for sm_100a.
// Reference: lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:91-138,
// lib/utils/fpn.py:11-28, lib/utils/boxes.py:58-69.
//
// One CTA per group of images (the reference collects over its whole minibatch): exact
// top-`post` of the concatenated per-level proposals by score, then the FPN level of each
// survivor, then a stable per-level split giving `order` (concat of the per-level index
// lists) and `restore` (= rois_idx_restore_int32, the inverse permutation).
#include <math.h>
#include "common.cuh"
#include "select_sort.cuh"

namespace vosd {

// map_rois_to_fpn_levels (fpn.py:11-28): all fp32 as NumPy evaluates it; log2 is taken in
// fp64 and rounded to fp32 (NumPy's SIMD float32 log2 is not correctly rounded, so a level can
// only differ when 4 + log2(.) lies within ~2 ulp of an integer).
__device__ __forceinline__ int fpn_level(float x1, float y1, float x2, float y2, int k_min, int k_max,
                                         float s0, int lvl0) {
    const float w = __fadd_rn(__fsub_rn(x2, x1), 1.f);
    const float h = __fadd_rn(__fsub_rn(y2, y1), 1.f);
    float area = __fmul_rn(w, h);
    if (area < 0.f) area = 0.f;
    const float s = __fsqrt_rn(area);
    const float v = __fadd_rn(__fdiv_rn(s, s0), (float)1e-6);
    const float lg = (float)log2((double)v);
    float t = floorf(__fadd_rn((float)lvl0, lg));
    t = fminf(fmaxf(t, (float)k_min), (float)k_max);
    return (int)t;
}

// Stable split of items [0, n) by level: order[] = indices grouped by level (ascending inside a
// level), restore[i] = position of i in order[].  Block of kSelThreads.
__device__ void split_by_level(const int* lvl, int n, int k_min, int k_max, int* order, int* restore,
                               int* level_count, int* warp_sums) {
    if (n <= kSelThreads && k_max - k_min < 4) {
        // one tile, <= 4 levels: ONE block scan over four 16-bit counters packed in a 64-bit word instead of one scan per
        // level (same stable order: position within the level = number of earlier rows of that level)
        const int t = threadIdx.x;
        const int L = t < n ? lvl[t] - k_min : -1;
        const unsigned long long v = L >= 0 ? 1ull << (16 * L) : 0ull;
        const int lane = t & 31, warp = t >> 5;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        __shared__ unsigned long long wsum[32];
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            wsum[lane] = w;                              // inclusive sums of the warp totals
        }
        __syncthreads();
        const unsigned long long total = wsum[31];
        const unsigned long long excl = (warp == 0 ? 0ull : wsum[warp - 1]) + inc - v;
        if (L >= 0) {
            int base = 0;
            for (int k = 0; k < L; k++) base += (int)((total >> (16 * k)) & 0xffffull);
            const int pos = base + (int)((excl >> (16 * L)) & 0xffffull);
            order[pos] = t;
            restore[t] = pos;
        }
        if (t <= k_max - k_min) level_count[t] = (int)((total >> (16 * t)) & 0xffffull);
        __syncthreads();
        return;
    }
    int base = 0;
    for (int L = k_min; L <= k_max; L++) {
        const int start = base;
        for (int t0 = 0; t0 < n; t0 += kSelThreads) {
            const int t = t0 + threadIdx.x;
            const int f = (t < n && lvl[t] == L) ? 1 : 0;
            int total;
            const int off = block_exclusive_scan(f, warp_sums, total);
            if (f) { order[base + off] = t; restore[t] = base + off; }
            base += total;
        }
        if (threadIdx.x == 0) level_count[L - k_min] = base - start;
    }
}

struct CollectKeys {
    const float* probs;   // (L, N, cap)
    const int* count;     // (L, N)
    int N, cap, ipg, img0;
    // j enumerates (level, image-in-group, slot) -- the reference's concatenation order
    __device__ __forceinline__ uint64_t operator()(int j) const {
        // j / cap and s / ipg through fp32 (exact here: j < 2^22, and (j + 0.5) / cap stays >= 0.5 / cap away from every
        // integer, far more than the rounding error of the division); a handful of instructions instead of ~40
        int s, l;
        if (j < (1 << 22)) {
            s = __float2int_rd(__fdividef((float)j + 0.5f, (float)cap));
            l = __float2int_rd(__fdividef((float)s + 0.5f, (float)ipg));
        } else {
            s = j / cap;
            l = s / ipg;
        }
        const int slot = j - s * cap;
        const int img = img0 + (s - l * ipg);
        if (slot >= count[l * N + img]) return 0;
        const float p = __ldg(probs + ((size_t)l * N + img) * cap + slot);
        return ((uint64_t)float_to_ordered(p) << 32) | (uint64_t)(0xffffffffu - (uint32_t)j);
    }
};

// grid = groups, block = 1024, dyn smem = P*8 (keys) + P*4 (levels)
__global__ void __launch_bounds__(kSelThreads, 1)
collect_distribute_kernel(const float* __restrict__ rois, const float* __restrict__ probs,
                          const int* __restrict__ count, int num_levels, int N, int cap, int ipg,
                          int post, int P, int k_min, int k_max, float s0, int lvl0,
                          float* __restrict__ out_rois, int* __restrict__ out_count,
                          int* __restrict__ out_level, int* __restrict__ level_count,
                          int* __restrict__ order, int* __restrict__ restore) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(dyn);
    int* lvl = reinterpret_cast<int*>(dyn + (size_t)P * sizeof(uint64_t));
    __shared__ SelectShared sh;
    __shared__ int n_valid_s;

    const int g = blockIdx.x;
    const int img0 = g * ipg;
    if (threadIdx.x == 0) {
        int nv = 0;
        for (int l = 0; l < num_levels; l++)
            for (int i = 0; i < ipg; i++) nv += min(count[l * N + img0 + i], cap);
        n_valid_s = nv;
    }
    __syncthreads();
    const int n_valid = n_valid_s;
    CollectKeys kf{probs, count, N, cap, ipg, img0};
    const int take = select_and_sort(kf, num_levels * ipg * cap, n_valid, post, keys, P, sh);

    float* orow = out_rois + (size_t)g * post * 5;
    for (int t = threadIdx.x; t < take; t += kSelThreads) {
        const int j = (int)(0xffffffffu - (uint32_t)keys[t]);
        const int slot = j % cap, s = j / cap;
        const int l = s / ipg, img = img0 + (s - l * ipg);
        const float* r = rois + (((size_t)l * N + img) * cap + slot) * 5;
        const float b = r[0], x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
        float* o = orow + (size_t)t * 5;
        o[0] = b; o[1] = x1; o[2] = y1; o[3] = x2; o[4] = y2;
        const int L = fpn_level(x1, y1, x2, y2, k_min, k_max, s0, lvl0);
        lvl[t] = L;
        out_level[(size_t)g * post + t] = L;
    }
    // rows beyond the group's count: zero boxes, level 0, identity-free zeros (the caller does not have to clear the
    // outputs: six fill kernels less on the proposal chain of a step)
    for (int t = take + threadIdx.x; t < post; t += kSelThreads) {
        float* o = orow + (size_t)t * 5;
        o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f;
        out_level[(size_t)g * post + t] = 0;
        order[(size_t)g * post + t] = 0;
        restore[(size_t)g * post + t] = 0;
    }
    __syncthreads();
    split_by_level(lvl, take, k_min, k_max, order + (size_t)g * post, restore + (size_t)g * post,
                   level_count + (size_t)g * (k_max - k_min + 1), sh.warp_sums);
    if (threadIdx.x == 0) out_count[g] = take;
}

// distribute() on caller-supplied rois; 1 CTA, levels staged in global out_level.
__global__ void __launch_bounds__(kSelThreads, 1)
distribute_kernel(const float* __restrict__ rois, int n, int k_min, int k_max, float s0, int lvl0,
                  int* __restrict__ out_level, int* __restrict__ level_count,
                  int* __restrict__ order, int* __restrict__ restore) {
    __shared__ int warp_sums[32];
    for (int t = threadIdx.x; t < n; t += kSelThreads) {
        const float* r = rois + (size_t)t * 5;
        out_level[t] = fpn_level(r[1], r[2], r[3], r[4], k_min, k_max, s0, lvl0);
    }
    __syncthreads();
    split_by_level(out_level, n, k_min, k_max, order, restore, level_count, warp_sums);
}

}  // namespace vosd

using namespace vosd;

extern "C" size_t vosd_collect_distribute_workspace_bytes(int, int, int, int, int) {
    return 256;   // everything lives in shared memory; kept for ABI symmetry
}

extern "C" int vosd_collect_distribute(const float* rois, const float* probs, const int* count,
                                       int num_levels, int num_images, int cap, int images_per_group,
                                       int post_nms_topN, int k_min, int k_max,
                                       float canonical_scale, int canonical_level,
                                       float* out_rois, int* out_count, int* out_level, int* level_count,
                                       int* order, int* restore,
                                       void*, size_t, cudaStream_t stream) {
    if (num_levels < 1 || num_levels > VOSD_MAX_LEVELS || num_images < 1 || cap < 1 ||
        images_per_group < 1 || num_images % images_per_group != 0 || post_nms_topN < 1 || k_max < k_min)
        return VOSD_ERR_BAD_SHAPE;
    if (post_nms_topN > VOSD_MAX_TOPK) return VOSD_ERR_UNSUPPORTED;
    if ((long long)num_levels * images_per_group * cap > 0x7fffffffLL / 8) return VOSD_ERR_UNSUPPORTED;
    if (!rois || !probs || !count || !out_rois || !out_count || !out_level || !level_count || !order || !restore)
        return VOSD_ERR_BAD_ARG;
    const int P = next_pow2(post_nms_topN);
    const size_t dyn = (size_t)P * (sizeof(uint64_t) + sizeof(int));
    if (cudaFuncSetAttribute(collect_distribute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    collect_distribute_kernel<<<num_images / images_per_group, kSelThreads, dyn, stream>>>(
        rois, probs, count, num_levels, num_images, cap, images_per_group, post_nms_topN, P, k_min, k_max,
        canonical_scale, canonical_level, out_rois, out_count, out_level, level_count, order, restore);
    count_launch();
    return check_launch();
}

extern "C" int vosd_distribute(const float* rois, int num_rois, int k_min, int k_max,
                               float canonical_scale, int canonical_level,
                               int* out_level, int* level_count, int* order, int* restore,
                               cudaStream_t stream) {
    if (num_rois < 0 || k_max < k_min) return VOSD_ERR_BAD_SHAPE;
    if (!level_count) return VOSD_ERR_BAD_ARG;
    if (num_rois > 0 && (!rois || !out_level || !order || !restore)) return VOSD_ERR_BAD_ARG;
    distribute_kernel<<<1, kSelThreads, 0, stream>>>(rois, num_rois, k_min, k_max, canonical_scale,
                                                     canonical_level, out_level, level_count, order, restore);
    count_launch();
    return check_launch();
}
