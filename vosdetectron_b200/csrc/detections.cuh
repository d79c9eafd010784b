// Box-head post-processing on device.  Included by proposals.cu (reuses its
// decode / clip arithmetic and the NMS mask + reduce kernels).
// Reference: im_detect_bbox's decode, lib/core/test.py:166-181 (bbox_transform with BBOX_REG_WEIGHTS +
// clip_tiled_boxes), and box_results_with_nms_and_limit, lib/core/test.py:733-797 (twin:
// lib_vos/tools/vos_test.py:748-865): per class j >= 1 keep scores >= SCORE_THRESH, greedy NMS at TEST.NMS,
// results in ascending proposal index (np.where order); then, if more than DETECTIONS_PER_IM detections
// survive over all classes, keep those with score >= the DETECTIONS_PER_IM-th largest score (ties kept).
#pragma once

namespace vosd {

// ---- bbox_transform + clip for (n, 4k) deltas: thread per (row, class) ----
__global__ void __launch_bounds__(256)
bbox_transform_kernel(const float* __restrict__ boxes, const float* __restrict__ deltas, int n, int k,
                      float wx, float wy, float ww, float wh, double clip, float im_h, float im_w, int do_clip,
                      float* __restrict__ out) {
    const long long total = (long long)n * k;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / k);
        const float4 b = *reinterpret_cast<const float4*>(boxes + 4 * (size_t)r);
        const float4 d = *reinterpret_cast<const float4*>(deltas + 4 * (size_t)idx);
        float4 o = decode_box(b.x, b.y, b.z, b.w, d.x, d.y, d.z, d.w, wx, wy, ww, wh, clip);
        if (do_clip) o = clip_box(o, im_h, im_w);
        *reinterpret_cast<float4*>(out + 4 * (size_t)idx) = o;
    }
}

// ---- per (image, class) segment: threshold, sort by (score desc, index asc), gather the class's boxes ----
struct ClsKeys {
    const float* s;     // scores of one image, (R, K)
    int K, j;
    float thresh;
    __device__ __forceinline__ uint64_t operator()(int r) const {
        const float v = __ldg(s + (size_t)r * K + j);
        if (!(v >= thresh)) return 0;
        return ((uint64_t)float_to_ordered(v) << 32) | (uint64_t)(0xffffffffu - (uint32_t)r);
    }
};

// grid = N * (K - 1), block = 1024, dyn smem = P * 8.  Segment seg = img * (K-1) + (j-1).
__global__ void __launch_bounds__(kSelThreads, 1)
cls_sort_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const int* __restrict__ rows,
                int R, int K, float thresh, int P, float4* __restrict__ ws_boxes, int* __restrict__ ws_orig,
                int* __restrict__ ws_flag, int* __restrict__ ws_count) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(dyn);
    __shared__ SelectShared sh;
    __shared__ int warp_sums[32];
    const int seg = blockIdx.x;
    const int img = seg / (K - 1), j = 1 + seg % (K - 1);
    const int n = rows ? min(rows[img], R) : R;
    ClsKeys kf{scores + (size_t)img * R * K, K, j, thresh};
    // number of scores over the threshold
    int mine = 0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) mine += kf(r) != 0;
    int n_valid;
    block_exclusive_scan(mine, warp_sums, n_valid);
    for (int t = threadIdx.x; t < R; t += blockDim.x) ws_flag[(size_t)seg * R + t] = 0;
    const int take = select_and_sort(kf, n, n_valid, n_valid, keys, P, sh);
    const float* bimg = boxes + (size_t)img * R * 4 * K + 4 * j;
    for (int t = threadIdx.x; t < take; t += blockDim.x) {
        const int r = (int)(0xffffffffu - (uint32_t)keys[t]);
        ws_boxes[(size_t)seg * R + t] = *reinterpret_cast<const float4*>(bimg + (size_t)r * 4 * K);
        ws_orig[(size_t)seg * R + t] = r;
    }
    if (threadIdx.x == 0) ws_count[seg] = take;
}

// ---- per image: limit to max_per_image over all classes, emit in (class, proposal index) order ----
struct LimitKeys {
    const int* flag;    // (K-1, R) keep flags of this image
    const float* s;     // (R, K)
    int R, K;
    __device__ __forceinline__ uint64_t operator()(int e) const {
        if (!flag[e]) return 0;
        const int c = e / R, r = e - c * R;
        return ((uint64_t)float_to_ordered(__ldg(s + (size_t)r * K + c + 1)) << 32) | (uint64_t)(0xffffffffu - (uint32_t)e);
    }
};

// grid = N, block = 1024, dyn smem = K ints.
__global__ void __launch_bounds__(kSelThreads, 1)
det_limit_kernel(const float* __restrict__ scores, const float* __restrict__ boxes, const int* __restrict__ ws_flag,
                 int R, int K, int max_per_image, int cap, float* __restrict__ out_dets, int* __restrict__ out_count,
                 int* __restrict__ out_cls_count) {
    extern __shared__ int cls_cnt[];
    __shared__ SelectShared sh;
    __shared__ int warp_sums[32];
    __shared__ unsigned long long min_key;
    const int img = blockIdx.x;
    const int n = (K - 1) * R;
    LimitKeys kf{ws_flag + (size_t)img * n, scores + (size_t)img * R * K, R, K};
    for (int c = threadIdx.x; c < K; c += blockDim.x) cls_cnt[c] = 0;
    if (threadIdx.x == 0) min_key = ~0ull;
    int mine = 0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) mine += kf.flag[e] != 0;
    int total;
    block_exclusive_scan(mine, warp_sums, total);
    uint32_t thr_bits = 0;                               // ordered score bits of image_thresh (0: keep all)
    if (max_per_image > 0 && total > max_per_image) {
        uint64_t mask = 0, prefix = 0;
        radix_select(kf, n, max_per_image, sh, mask, prefix);
        // image_thresh = the smallest of the max_per_image largest scores (np.sort(...)[-max_per_image])
        unsigned long long lo = ~0ull;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const uint64_t k = kf(e);
            if (k != 0 && (k & mask) >= prefix) lo = min(lo, (unsigned long long)k);
        }
        atomicMin(&min_key, lo);
        __syncthreads();
        thr_bits = (uint32_t)(min_key >> 32);
    }
    __syncthreads();
    const float* s = kf.s;
    const float* bimg = boxes + (size_t)img * R * 4 * K;
    float* od = out_dets + (size_t)img * cap * 6;
    int base = 0;
    for (int e0 = 0; e0 < n; e0 += kSelThreads) {
        const int e = e0 + threadIdx.x;
        int f = 0, c = 0, r = 0;
        float sc = 0.f;
        if (e < n && kf.flag[e]) {
            c = e / R; r = e - c * R;
            sc = __ldg(s + (size_t)r * K + c + 1);
            f = float_to_ordered(sc) >= thr_bits;
        }
        int tot;
        const int off = block_exclusive_scan(f, warp_sums, tot);
        if (f) {
            const int pos = base + off;
            if (pos < cap) {
                const float4 b = *reinterpret_cast<const float4*>(bimg + (size_t)r * 4 * K + 4 * (c + 1));
                float* o = od + (size_t)pos * 6;
                o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w; o[4] = sc; o[5] = (float)(c + 1);
            }
            atomicAdd(&cls_cnt[c + 1], 1);
        }
        base += tot;
    }
    __syncthreads();
    if (threadIdx.x == 0) out_count[img] = base;
    if (out_cls_count)
        for (int c = threadIdx.x; c < K; c += blockDim.x) out_cls_count[(size_t)img * K + c] = cls_cnt[c];
}

struct DetLayout {
    size_t off_boxes, off_orig, off_flag, off_count, off_mask, total;
    int words, segs;
};
static DetLayout det_layout(int N, int R, int K) {
    DetLayout L;
    L.words = (R + 63) / 64;
    L.segs = N * (K - 1);
    size_t o = 0;
    L.off_boxes = o; o = align_up(o + (size_t)L.segs * R * sizeof(float4), 256);
    L.off_orig = o;  o = align_up(o + (size_t)L.segs * R * sizeof(int), 256);
    L.off_flag = o;  o = align_up(o + (size_t)L.segs * R * sizeof(int), 256);
    L.off_count = o; o = align_up(o + (size_t)L.segs * sizeof(int), 256);
    L.off_mask = o;  o = align_up(o + (size_t)L.segs * R * L.words * sizeof(unsigned long long), 256);
    L.total = o;
    return L;
}

}  // namespace vosd

using namespace vosd;

extern "C" int vosd_bbox_transform(const float* boxes, const float* deltas, int n, int k, const float* weights,
                                   float clip_h, float clip_w, float* out, cudaStream_t stream) {
    if (n < 0 || k <= 0) return VOSD_ERR_BAD_SHAPE;
    if (n == 0) return VOSD_OK;
    if (!boxes || !deltas || !weights || !out) return VOSD_ERR_BAD_ARG;
    if (!aligned16(boxes) || !aligned16(deltas) || !aligned16(out)) return VOSD_ERR_BAD_ARG;
    const long long total = (long long)n * k;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)kNumSMs * 32) blocks = (long long)kNumSMs * 32;
    bbox_transform_kernel<<<(int)blocks, 256, 0, stream>>>(boxes, deltas, n, k, weights[0], weights[1], weights[2],
                                                           weights[3], log(1000.0 / 16.0), clip_h, clip_w,
                                                           clip_h > 0.f && clip_w > 0.f, out);
    count_launch();
    return check_launch();
}

extern "C" size_t vosd_box_results_workspace_bytes(int num_images, int rois_per_image, int num_classes) {
    if (num_images <= 0 || rois_per_image <= 0 || num_classes < 2) return 256;
    return det_layout(num_images, rois_per_image, num_classes).total;
}

extern "C" int vosd_box_results(const float* scores, const float* boxes, const int* rows, int num_images,
                                int rois_per_image, int num_classes, float score_thresh, float nms_thresh,
                                int max_per_image, int cap, float* out_dets, int* out_count, int* out_cls_count,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int N = num_images, R = rois_per_image, K = num_classes;
    if (N <= 0 || R <= 0 || K < 2 || cap <= 0) return VOSD_ERR_BAD_SHAPE;
    if (R > VOSD_MAX_TOPK || K > 4096) return VOSD_ERR_UNSUPPORTED;
    if (!scores || !boxes || !out_dets || !out_count) return VOSD_ERR_BAD_ARG;
    if (!aligned16(boxes)) return VOSD_ERR_BAD_ARG;
    const DetLayout L = det_layout(N, R, K);
    if (!workspace || workspace_bytes < L.total || !aligned16(workspace)) return VOSD_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    float4* wb = reinterpret_cast<float4*>(ws + L.off_boxes);
    int* orig = reinterpret_cast<int*>(ws + L.off_orig);
    int* flag = reinterpret_cast<int*>(ws + L.off_flag);
    int* count = reinterpret_cast<int*>(ws + L.off_count);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + L.off_mask);
    const int P = next_pow2(R);
    const size_t dyn = (size_t)P * sizeof(uint64_t);
    if (cudaFuncSetAttribute(cls_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    cls_sort_kernel<<<L.segs, kSelThreads, dyn, stream>>>(scores, boxes, rows, R, K, score_thresh, P, wb, orig, flag, count);
    const int use_mask = nms_thresh > 0.f;
    if (use_mask) {
        nms_mask_kernel<<<nms_mask_grid(L.words, L.segs), 64, 0, stream>>>(wb, count, R, L.words, nms_thresh, mask);
    }
    if (launch_nms_reduce(L.segs, R, L.words, wb, nullptr, count, mask, use_mask, 0, 2, 1, R, nullptr, nullptr, nullptr, orig, flag,
                          stream) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    det_limit_kernel<<<N, kSelThreads, (size_t)K * sizeof(int), stream>>>(scores, boxes, flag, R, K, max_per_image, cap,
                                                                          out_dets, out_count, out_cls_count);
    count_launch(use_mask ? 4 : 3);
    return check_launch();
}
