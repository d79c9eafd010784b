// Mask-IoU suppression over bit-packed masks + RLE -> bit expansion (second half).
//
// Reference: lib_vos/tools/vos_test.py
//   nms_with_mask_iou :985-1029  decode every RLE to a dense (H,W) uint8 mask, sort by score, O(R^2) Python loop
//                                of iou_half_numpy over full frames, greedy discard
//   iou_half_numpy    :953-959   inter = sum(a & b); iou1 = inter / (sum(a) + 1e-6); iou2 = inter / (sum(b) + 1e-6)
//
// Here a mask is 1 bit per pixel (the layout vosd_paste_masks_packed writes, or vosd_rle_to_bits): a pair costs
// one AND + POPC per 32 pixels instead of a byte-wise NumPy pass over two dense frames.  One CTA owns mask i and
// a group of up to 8 later masks j (mask i's words are read once per group); the comparison is the reference's
// float64 expression, so the keep decisions are identical.  The greedy pass is one warp over the suppression
// bit matrix, in score order.
#include "common.cuh"

namespace vosd {
namespace {

constexpr int kPairGroup = 8;
constexpr int kMaskNmsMax = 2048;     // masks per call (greedy pass keeps `removed` as 32 x 64-bit words in a warp)

__global__ void __launch_bounds__(256) mask_area_kernel(const uint32_t* __restrict__ bits, long long words,
                                                        int* __restrict__ area) {
    const uint32_t* m = bits + (size_t)blockIdx.x * words;
    int acc = 0;
    for (long long i = threadIdx.x; i < words; i += blockDim.x) acc += __popc(__ldg(m + i));
    __shared__ int part[8];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < 8; ++k) s += part[k];
        area[blockIdx.x] = s;
    }
}

// grid.x = sorted position i, grid.y = group of kPairGroup later positions.  order[k] = mask index at sorted
// position k (NULL: identity).  Sets bit j of row i of `suppress` when the pair (i, j), i < j, exceeds the threshold.
__global__ void __launch_bounds__(256) mask_pair_kernel(const uint32_t* __restrict__ bits, long long words,
                                                        const int* __restrict__ order, const int* __restrict__ area,
                                                        int n, double iou_th, int row_words,
                                                        unsigned long long* __restrict__ suppress) {
    const int i = blockIdx.x;
    const int j0 = i + 1 + blockIdx.y * kPairGroup;
    if (j0 >= n) return;
    const int cnt = min(kPairGroup, n - j0);
    const int mi = order ? order[i] : i;
    const uint32_t* a = bits + (size_t)mi * words;
    const uint32_t* b[kPairGroup];
    int mj[kPairGroup];
#pragma unroll
    for (int k = 0; k < kPairGroup; ++k) {
        mj[k] = k < cnt ? (order ? order[j0 + k] : j0 + k) : mi;
        b[k] = bits + (size_t)mj[k] * words;
    }
    int acc[kPairGroup];
#pragma unroll
    for (int k = 0; k < kPairGroup; ++k) acc[k] = 0;
    for (long long w = threadIdx.x; w < words; w += blockDim.x) {
        const uint32_t x = __ldg(a + w);
        if (x == 0) continue;                                    // most of a frame is background
#pragma unroll
        for (int k = 0; k < kPairGroup; ++k) acc[k] += __popc(x & __ldg(b[k] + w));
    }
    __shared__ int part[8][kPairGroup];
#pragma unroll
    for (int k = 0; k < kPairGroup; ++k) {
        int v = acc[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
        const int k = threadIdx.x;
        int inter = 0;
        for (int q = 0; q < 8; ++q) inter += part[q][k];
        // vos_test.py:953-959, float64: intersection / (union + 1e-6)
        const double iou1 = (double)inter / ((double)area[mi] + 1e-6);
        const double iou2 = (double)inter / ((double)area[mj[k]] + 1e-6);
        if (iou1 > iou_th || iou2 > iou_th) {                    // :1009
            const int j = j0 + k;
            atomicOr(suppress + (size_t)i * row_words + (j >> 6), 1ull << (j & 63));
        }
    }
}

// One warp: lane l holds word l of `removed`.  Position i survives unless an earlier survivor suppressed it (:1001-1010).
__global__ void __launch_bounds__(32) mask_greedy_kernel(const unsigned long long* __restrict__ suppress, int n,
                                                         int row_words, int* __restrict__ removed_out,
                                                         int* __restrict__ num_keep) {
    const int lane = threadIdx.x;
    unsigned long long removed = 0;
    int kept = 0;
    for (int i = 0; i < n; ++i) {
        const unsigned long long word = __shfl_sync(0xffffffffu, removed, i >> 6);
        const bool dead = (word >> (i & 63)) & 1ull;
        if (!dead) {
            ++kept;
            if (lane < row_words) removed |= suppress[(size_t)i * row_words + lane];
        }
    }
    for (int base = 0; base < n; base += 32) {              // padded trip count: every lane takes part in the shuffle
        const int i = base + lane;
        const unsigned long long word = __shfl_sync(0xffffffffu, removed, min(i, n - 1) >> 6);
        if (i < n) removed_out[i] = (int)((word >> (i & 63)) & 1ull);
    }
    if (lane == 0) *num_keep = kept;
}

// ------------------------------------------------------------------------------------- RLE -> bits
// One CTA per mask.  runs[off .. off+cnt) are the uncompressed COCO counts (alternating 0-runs and 1-runs over the
// column-major pixel sequence, first run = zeros).  Each thread owns 32-bit output words and finds the run that
// covers its first pixel by binary search over the inclusive prefix sums kept in shared memory.
constexpr int kRleMaxRuns = 12000;

__global__ void __launch_bounds__(256) rle_to_bits_kernel(const uint32_t* __restrict__ runs,
                                                          const long long* __restrict__ run_offset,
                                                          const int* __restrict__ run_count, long long pixels,
                                                          long long words, uint32_t* __restrict__ out) {
    __shared__ uint32_t ends[kRleMaxRuns];      // ends[k] = number of pixels covered by runs 0..k
    __shared__ uint32_t carry;
    const int m = blockIdx.x;
    const int cnt = min(run_count[m], kRleMaxRuns);
    const uint32_t* r = runs + run_offset[m];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    // block-wide inclusive scan in tiles of 256
    for (int base = 0; base < cnt; base += 256) {
        const int k = base + threadIdx.x;
        uint32_t v = k < cnt ? r[k] : 0;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += t;
        }
        __shared__ uint32_t wsum[8];
        if (lane == 31) wsum[warp] = v;
        __syncthreads();
        uint32_t add = carry;
        for (int q = 0; q < warp; ++q) add += wsum[q];
        if (k < cnt) ends[k] = v + add;
        __syncthreads();
        if (threadIdx.x == 255) carry = v + add;
        __syncthreads();
    }
    uint32_t* o = out + (size_t)m * words;
    for (long long w = threadIdx.x; w < words; w += blockDim.x) {
        const long long p0 = w * 32;
        // first run whose end is > p0
        int lo = 0, hi = cnt;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((long long)ends[mid] > p0) hi = mid; else lo = mid + 1;
        }
        uint32_t word = 0;
        long long p = p0;
        const long long pend = min(p0 + 32, pixels);
        int k = lo;
        while (p < pend && k < cnt) {
            const long long e = min((long long)ends[k], pend);
            if ((k & 1) && e > p) {                                   // odd runs are ones
                const int a = (int)(p - p0), len = (int)(e - p);
                word |= (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << a;
            }
            p = max(p, e);
            ++k;
        }
        o[w] = word;
    }
}

}  // namespace
}  // namespace vosd

extern "C" size_t vosd_mask_iou_nms_workspace_bytes(int num_masks) {
    if (num_masks <= 0) return 256;
    const size_t row_words = (size_t)(num_masks + 63) / 64;
    return vosd::align_up((size_t)num_masks * sizeof(int), 256) + vosd::align_up((size_t)num_masks * row_words * 8, 256);
}

extern "C" int vosd_mask_iou_nms(const uint8_t* packed, int num_masks, long long bytes_per_mask, const int* order,
                                 double iou_th, int* removed, int* num_keep, void* workspace, size_t workspace_bytes,
                                 cudaStream_t stream) {
    using namespace vosd;
    if (num_masks < 0 || bytes_per_mask < 0) return VOSD_ERR_BAD_SHAPE;
    if (!num_keep) return VOSD_ERR_BAD_ARG;
    if (num_masks == 0) return cudaMemsetAsync(num_keep, 0, sizeof(int), stream) == cudaSuccess ? VOSD_OK : VOSD_ERR_LAUNCH;
    if (!packed || !removed) return VOSD_ERR_BAD_ARG;
    if (bytes_per_mask % 4 || (reinterpret_cast<uintptr_t>(packed) & 3)) return VOSD_ERR_BAD_ARG;   // 32-bit words
    if (num_masks > kMaskNmsMax) return VOSD_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < vosd_mask_iou_nms_workspace_bytes(num_masks)) return VOSD_ERR_WORKSPACE;
    const long long words = bytes_per_mask / 4;
    const int row_words = (num_masks + 63) / 64;
    int* area = static_cast<int*>(workspace);
    unsigned long long* suppress = reinterpret_cast<unsigned long long*>(
        static_cast<char*>(workspace) + align_up((size_t)num_masks * sizeof(int), 256));
    if (cudaMemsetAsync(suppress, 0, (size_t)num_masks * row_words * 8, stream) != cudaSuccess) return VOSD_ERR_LAUNCH;
    const uint32_t* bits = reinterpret_cast<const uint32_t*>(packed);
    mask_area_kernel<<<num_masks, 256, 0, stream>>>(bits, words, area);
    int launches = 2;
    if (num_masks > 1) {
        dim3 grid(num_masks - 1, ceil_div(num_masks - 1, kPairGroup));
        mask_pair_kernel<<<grid, 256, 0, stream>>>(bits, words, order, area, num_masks, iou_th, row_words, suppress);
        ++launches;
    }
    mask_greedy_kernel<<<1, 32, 0, stream>>>(suppress, num_masks, row_words, removed, num_keep);
    count_launch(launches);
    return check_launch();
}

extern "C" int vosd_rle_to_bits(const uint32_t* runs, const long long* run_offset, const int* run_count,
                                int num_masks, long long pixels, uint8_t* out_packed, int max_run_count,
                                cudaStream_t stream) {
    using namespace vosd;
    if (num_masks < 0 || pixels < 0) return VOSD_ERR_BAD_SHAPE;
    if (num_masks == 0 || pixels == 0) return VOSD_OK;
    if (!runs || !run_offset || !run_count || !out_packed) return VOSD_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(out_packed) & 3) return VOSD_ERR_BAD_ARG;
    if (max_run_count > kRleMaxRuns) return VOSD_ERR_UNSUPPORTED;
    const long long words = (pixels + 31) / 32;
    rle_to_bits_kernel<<<num_masks, 256, 0, stream>>>(runs, run_offset, run_count, pixels, words,
                                                      reinterpret_cast<uint32_t*>(out_packed));
    count_launch();
    return check_launch();
}
