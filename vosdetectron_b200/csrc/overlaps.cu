// bbox_overlaps (first piece of the training label assignment).
//
// Reference: lib/utils/cython_bbox.pyx:32-73 (bound as box_utils.bbox_overlaps, lib/utils/boxes.py:55); callers on
// the training path: datasets/json_dataset.py:450-456 (proposal -> gt overlaps, then .argmax(axis=1) / .max(axis=1)),
// roi_data/rpn.py:149-158 (anchor -> gt overlaps, row and column maxima), roi_data/mask_rcnn.py:58.
//
// Arithmetic as compiled from the .pyx: Cython types the literal in `x2 - x1 + 1` as the double 1.0, so widths,
// heights and the areas are formed in float64 (`box_area` and `iw`, `ih` are rounded to their float32 variables, the
// union `ua = float(area_n + box_area - iw * ih)` is summed in float64 and rounded once), `iw * ih` and the final
// division are float32; no fused multiply-add (baseline x86-64).  A pair overlaps only if iw > 0 and ih > 0.
// One thread per box row, query boxes in shared memory; the row maximum / first arg-maximum (NumPy argmax tie
// rule) come out of the same pass, so the (N,K) matrix need not be written at all when only the labels are wanted.
#include "common.cuh"
#include "select_sort.cuh"

namespace vosd {
namespace {

constexpr int kQueryTile = 1024;    // query boxes per shared-memory tile

__global__ void __launch_bounds__(256) bbox_overlaps_kernel(const float4* __restrict__ boxes, int N,
                                                            const float4* __restrict__ query, int K,
                                                            float* __restrict__ overlaps, float* __restrict__ row_max,
                                                            int* __restrict__ row_argmax) {
    __shared__ float4 q[kQueryTile];
    __shared__ float qarea[kQueryTile];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) b = boxes[n];
    // area of box n in float64 (exact: two 25-bit factors), cython_bbox.pyx:65-66
    const double barea = __dmul_rn((double)__fsub_rn(b.z, b.x) + 1.0, (double)__fsub_rn(b.w, b.y) + 1.0);
    float best = -1.f;          // overlaps are >= 0: the first column wins ties, like np.argmax
    int best_k = 0;
    for (int k0 = 0; k0 < K; k0 += kQueryTile) {
        const int kt = min(kQueryTile, K - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < kt; i += blockDim.x) {
            const float4 v = query[k0 + i];
            q[i] = v;
            qarea[i] = (float)__dmul_rn((double)__fsub_rn(v.z, v.x) + 1.0, (double)__fsub_rn(v.w, v.y) + 1.0);  // :49-52
        }
        __syncthreads();
        if (n >= N) continue;
        for (int i = 0; i < kt; ++i) {
            const float4 v = q[i];
            float o = 0.f;
            const float iw = (float)((double)__fsub_rn(fminf(b.z, v.z), fmaxf(b.x, v.x)) + 1.0);             // :54-57
            if (iw > 0.f) {
                const float ih = (float)((double)__fsub_rn(fminf(b.w, v.w), fmaxf(b.y, v.y)) + 1.0);         // :59-62
                if (ih > 0.f) {
                    const float inter = __fmul_rn(iw, ih);
                    const float ua = (float)__dsub_rn(__dadd_rn(barea, (double)qarea[i]), (double)inter);    // :64-68
                    o = __fdiv_rn(inter, ua);                                                                  // :69
                }
            }
            if (overlaps) overlaps[(size_t)n * K + k0 + i] = o;
            if (o > best) {
                best = o;
                best_k = k0 + i;
            }
        }
    }
    if (n < N) {
        if (row_max) row_max[n] = K > 0 ? best : 0.f;
        if (row_argmax) row_argmax[n] = best_k;
    }
}

}  // namespace
}  // namespace vosd

extern "C" int vosd_bbox_overlaps(const float* boxes, int num_boxes, const float* query_boxes, int num_query,
                                  float* overlaps, float* row_max, int* row_argmax, cudaStream_t stream) {
    using namespace vosd;
    if (num_boxes < 0 || num_query < 0) return VOSD_ERR_BAD_SHAPE;
    if (num_boxes == 0) return VOSD_OK;
    if (!boxes || (num_query && !query_boxes)) return VOSD_ERR_BAD_ARG;
    if (!aligned16(boxes) || (num_query && !aligned16(query_boxes))) return VOSD_ERR_BAD_ARG;
    bbox_overlaps_kernel<<<ceil_div(num_boxes, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(boxes), num_boxes, reinterpret_cast<const float4*>(query_boxes), num_query,
        overlaps, row_max, row_argmax);
    count_launch();
    return check_launch();
}

// ---------------------------------------------------------------------------------------------------------
// Box regression targets of the training label assignment:
//   _compute_targets      lib/roi_data/fast_rcnn.py:216-229  -> box_utils.bbox_transform_inv (lib/utils/boxes.py:208-239)
//   _expand_bbox_targets  lib/roi_data/fast_rcnn.py:232-260  (4-of-4K expansion, inside weights 1 on the label's class)
//   bbox_outside_weights = (bbox_inside_weights > 0)          lib/roi_data/fast_rcnn.py:206-208
// fp32 in NumPy's operation order (widths with + 1, centres, weights multiplied before the division, log of the
// ratio); one thread per RoI row writes the whole 4K-wide row (zeros + its four targets): every byte is written.
// ---------------------------------------------------------------------------------------------------------
namespace vosd {
namespace {

__global__ void __launch_bounds__(256) bbox_targets_kernel(const float4* __restrict__ ex, const float4* __restrict__ gt,
                                                           const int* __restrict__ labels, int n, int K, int agnostic,
                                                           float wx, float wy, float ww, float wh,
                                                           float* __restrict__ targets, float* __restrict__ inside,
                                                           float* __restrict__ outside) {
    // one warp per row: lanes stride over the 4K columns (coalesced zero fill), lane 0 derives the four targets
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    int cls = labels[row];
    if (agnostic) cls = min(cls, 1);                                     // clss.clip(max=1), :246-249
    const float4 b = ex[row], g = gt[row];
    const float ew = __fadd_rn(__fsub_rn(b.z, b.x), 1.f), eh = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);     // boxes.py:221-222
    const float ecx = __fadd_rn(b.x, __fmul_rn(0.5f, ew)), ecy = __fadd_rn(b.y, __fmul_rn(0.5f, eh));   // :223-224
    const float gw = __fadd_rn(__fsub_rn(g.z, g.x), 1.f), gh = __fadd_rn(__fsub_rn(g.w, g.y), 1.f);     // :226-227
    const float gcx = __fadd_rn(g.x, __fmul_rn(0.5f, gw)), gcy = __fadd_rn(g.y, __fmul_rn(0.5f, gh));   // :228-229
    float t[4];
    t[0] = __fdiv_rn(__fmul_rn(wx, __fsub_rn(gcx, ecx)), ew);                                           // :232
    t[1] = __fdiv_rn(__fmul_rn(wy, __fsub_rn(gcy, ecy)), eh);                                           // :233
    t[2] = __fmul_rn(ww, logf(__fdiv_rn(gw, ew)));                                                       // :234
    t[3] = __fmul_rn(wh, logf(__fdiv_rn(gh, eh)));                                                       // :235
    const size_t base = (size_t)row * 4 * K;
    for (int c = lane; c < 4 * K; c += 32) {
        const bool hit = cls > 0 && cls < K && (c >> 2) == cls;          // inds = where(clss > 0), :253-259
        const float v = hit ? t[c & 3] : 0.f;
        targets[base + c] = v;
        inside[base + c] = hit ? 1.f : 0.f;
        if (outside) outside[base + c] = hit ? 1.f : 0.f;
    }
}

}  // namespace
}  // namespace vosd

extern "C" int vosd_bbox_targets(const float* ex_rois, const float* gt_rois, const int* labels, int num_rois,
                                 int num_classes, int class_agnostic, const float* weights /*host*/,
                                 float* bbox_targets, float* inside_weights, float* outside_weights,
                                 cudaStream_t stream) {
    using namespace vosd;
    if (num_rois < 0 || num_classes < 1) return VOSD_ERR_BAD_SHAPE;
    if (num_rois == 0) return VOSD_OK;
    if (!ex_rois || !gt_rois || !labels || !weights || !bbox_targets || !inside_weights) return VOSD_ERR_BAD_ARG;
    if (!aligned16(ex_rois) || !aligned16(gt_rois)) return VOSD_ERR_BAD_ARG;
    const int K = class_agnostic ? 2 : num_classes;                      // :244-247
    bbox_targets_kernel<<<ceil_div(num_rois * 32, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(ex_rois), reinterpret_cast<const float4*>(gt_rois), labels, num_rois, K,
        class_agnostic ? 1 : 0, weights[0], weights[1], weights[2], weights[3], bbox_targets, inside_weights,
        outside_weights);
    count_launch();
    return check_launch();
}


// ---------------------------------------------------------------------------------------------------------
// Sampling of the training RoIs: _sample_rois (lib/roi_data/fast_rcnn.py:132-160).  The reference draws
// npr.choice(inds, size, replace=False) from NumPy's global generator; here the caller supplies one uniform key per
// box and "choice" is the `size` candidates with the SMALLEST keys in ascending key order (ties: lower index first),
// so the draw is reproducible, testable against the reference under the same contract, and the same distribution.
// One CTA per image: count the candidates, radix-select + sort them by key (select_sort.cuh), foreground first.
// ---------------------------------------------------------------------------------------------------------
namespace vosd {
namespace {

struct SampleKeys {
    const float* ov;
    const float* keys;
    float lo, hi;         // candidate: lo <= ov < hi  (foreground: lo = FG_THRESH, hi = +inf)
    __device__ __forceinline__ bool cand(int j) const { const float v = __ldg(ov + j); return v >= lo && v < hi; }
    __device__ __forceinline__ uint64_t operator()(int j) const {
        if (!cand(j)) return 0;
        return ((uint64_t)(~float_to_ordered(__ldg(keys + j))) << 32) | (uint64_t)(0xffffffffu - (uint32_t)j);
    }
};

__global__ void __launch_bounds__(kSelThreads, 1)
sample_rois_kernel(const float* __restrict__ max_overlaps, const float* __restrict__ keys, const int* __restrict__ num_boxes,
                   int stride, int rois_per_image, int fg_per_image, float fg_thresh, float bg_hi, float bg_lo, int P,
                   int* __restrict__ keep_inds, int* __restrict__ num_fg, int* __restrict__ num_keep) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(dyn);
    __shared__ SelectShared sh;
    const int b = blockIdx.x;
    const int n = num_boxes[b];
    int* keep = keep_inds + (size_t)b * rois_per_image;
    int written = 0, nfg = 0;
    for (int pass = 0; pass < 2; pass++) {
        SampleKeys kf{max_overlaps + (size_t)b * stride, keys + (size_t)b * stride, pass == 0 ? fg_thresh : bg_lo,
                      pass == 0 ? __int_as_float(0x7f800000) : bg_hi};
        int cnt = 0;
        for (int j0 = 0; j0 < n; j0 += kSelThreads) {
            const int j = j0 + threadIdx.x;
            cnt += __syncthreads_count(j < n && kf.cand(j));
        }
        const int want = min(pass == 0 ? fg_per_image : rois_per_image - nfg, cnt);
        const int take = select_and_sort(kf, n, cnt, want, sel, P, sh);
        for (int i = threadIdx.x; i < take; i += kSelThreads) keep[written + i] = (int)(0xffffffffu - (uint32_t)sel[i]);
        __syncthreads();
        if (pass == 0) nfg = take;
        written += take;
    }
    for (int i = written + threadIdx.x; i < rois_per_image; i += kSelThreads) keep[i] = -1;
    if (threadIdx.x == 0) { num_fg[b] = nfg; num_keep[b] = written; }
}

}  // namespace
}  // namespace vosd

extern "C" int vosd_sample_rois(const float* max_overlaps, const float* keys, const int* num_boxes, int num_images, int stride,
                                int rois_per_image, int fg_rois_per_image, float fg_thresh, float bg_thresh_hi,
                                float bg_thresh_lo, int* keep_inds, int* num_fg, int* num_keep, cudaStream_t stream) {
    using namespace vosd;
    if (num_images < 0 || stride < 0 || rois_per_image <= 0 || fg_rois_per_image < 0 || fg_rois_per_image > rois_per_image)
        return VOSD_ERR_BAD_SHAPE;
    if (num_images == 0) return VOSD_OK;
    if (!max_overlaps || !keys || !num_boxes || !keep_inds || !num_fg || !num_keep) return VOSD_ERR_BAD_ARG;
    if (rois_per_image > VOSD_MAX_TOPK) return VOSD_ERR_UNSUPPORTED;
    const int P = next_pow2(rois_per_image);
    const size_t dyn = (size_t)P * sizeof(uint64_t);
    if (cudaFuncSetAttribute(sample_rois_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    sample_rois_kernel<<<num_images, kSelThreads, dyn, stream>>>(max_overlaps, keys, num_boxes, stride, rois_per_image,
                                                                 fg_rois_per_image, fg_thresh, bg_thresh_hi, bg_thresh_lo, P,
                                                                 keep_inds, num_fg, num_keep);
    count_launch();
    return check_launch();
}
