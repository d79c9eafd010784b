// RPN proposal generation + greedy NMS for sm_100a.
// Reference: lib/modeling/generate_proposals.py:20-182, lib/utils/boxes.py:138-205,
// lib/utils/cython_nms.pyx:37-87.
//
// One call covers every (level, image) segment:
//   K1 topk_decode_kernel : 1 CTA / segment -- radix-select the pre_nms_topN best scores,
//                           bitonic-sort them, decode + clip + filter the survivors.
//   K2 nms_mask_kernel    : 64x64 IoU tiles -> suppression bitmask (upper triangle only).
//   K3 nms_reduce_cta_*   : 1 CTA / segment -- ordered greedy reduce over 64-box chunks,
//                           writes the first post_nms_topN kept boxes.
#include <math.h>
#include "common.cuh"
#include "select_sort.cuh"

namespace vosd {

// ------------------------------------------------------------------------------------
// Box arithmetic, operation for operation as NumPy evaluates it (no FMA contraction).
// ------------------------------------------------------------------------------------
// bbox_transform (boxes.py:156-205) for one anchor/delta pair.  BBOX_XFORM_CLIP is an
// np.float64 scalar (core/config.py:1009), so under NumPy >= 2 the dw/dh branch --
// np.minimum, np.exp, * widths, np.maximum(.,1), 0.5 * pred_w -- is carried in float64 and
// rounded once when stored into the float32 pred_boxes; the centre branch stays float32.
__device__ __forceinline__ float4 decode_box(float ax1, float ay1, float ax2, float ay2,
                                             float d0, float d1, float d2, float d3,
                                             float wx, float wy, float ww, float wh, double clip) {
    const float bw = __fadd_rn(__fsub_rn(ax2, ax1), 1.0f);
    const float bh = __fadd_rn(__fsub_rn(ay2, ay1), 1.0f);
    const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, bw));
    const float cy = __fadd_rn(ay1, __fmul_rn(0.5f, bh));
    // (x / 1 is x exactly: the RPN decode passes unit weights as literals and skips four IEEE divisions per anchor)
    const float dx = wx == 1.f ? d0 : __fdiv_rn(d0, wx), dy = wy == 1.f ? d1 : __fdiv_rn(d1, wy);
    const double dw = fmin((double)(ww == 1.f ? d2 : __fdiv_rn(d2, ww)), clip);
    const double dh = fmin((double)(wh == 1.f ? d3 : __fdiv_rn(d3, wh)), clip);
    const float pcx = __fadd_rn(__fmul_rn(dx, bw), cx);
    const float pcy = __fadd_rn(__fmul_rn(dy, bh), cy);
    const double pw = fmax(__dmul_rn(exp(dw), (double)bw), 1.0);
    const double ph = fmax(__dmul_rn(exp(dh), (double)bh), 1.0);
    const double hw = __dmul_rn(0.5, pw), hh = __dmul_rn(0.5, ph);
    float4 o;
    o.x = (float)__dsub_rn((double)pcx, hw);
    o.y = (float)__dsub_rn((double)pcy, hh);
    o.z = (float)__dsub_rn(__dadd_rn((double)pcx, hw), 1.0);
    o.w = (float)__dsub_rn(__dadd_rn((double)pcy, hh), 1.0);
    return o;
}

// clip_tiled_boxes (boxes.py:138-153): max(min(v, size - 1), 0) in fp32.
__device__ __forceinline__ float4 clip_box(float4 b, float im_h, float im_w) {
    const float mx = __fsub_rn(im_w, 1.f), my = __fsub_rn(im_h, 1.f);
    b.x = fmaxf(fminf(b.x, mx), 0.f);
    b.y = fmaxf(fminf(b.y, my), 0.f);
    b.z = fmaxf(fminf(b.z, mx), 0.f);
    b.w = fmaxf(fminf(b.w, my), 0.f);
    return b;
}

// _filter_boxes (generate_proposals.py:171-182).
__device__ __forceinline__ bool keep_box(float4 b, float min_size, float im_h, float im_w) {
    const float ws = __fadd_rn(__fsub_rn(b.z, b.x), 1.f);
    const float hs = __fadd_rn(__fsub_rn(b.w, b.y), 1.f);
    const float xc = __fadd_rn(b.x, __fdiv_rn(ws, 2.f));
    const float yc = __fadd_rn(b.y, __fdiv_rn(hs, 2.f));
    return ws >= min_size && hs >= min_size && xc < im_w && yc < im_h;
}

// IoU test of cython_nms.pyx:70-85 in IEEE fp32, +1 widths, `>=`.
__device__ __forceinline__ float box_area(float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
}
__device__ __forceinline__ bool suppresses(float4 a, float area_a, float4 b, float area_b, float thresh) {
    const float xx1 = a.x >= b.x ? a.x : b.x;
    const float yy1 = a.y >= b.y ? a.y : b.y;
    const float xx2 = a.z <= b.z ? a.z : b.z;
    const float yy2 = a.w <= b.w ? a.w : b.w;
    float w = __fadd_rn(__fsub_rn(xx2, xx1), 1.f);
    float h = __fadd_rn(__fsub_rn(yy2, yy1), 1.f);
    w = 0.f >= w ? 0.f : w;
    h = 0.f >= h ? 0.f : h;
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // The reference's decision is fl(inter / uni) >= thresh.  Two products decide all but the near-ties without the
    // IEEE division (about a dozen instructions): with r = inter / uni exactly,
    //   inter >= fl(fl(thresh * (1 + 2^-21)) * uni)  =>  r > thresh * (1 + 2^-22)  =>  fl(r) >= thresh   (rounding is monotone),
    //   inter <= fl(fl(thresh * (1 - 2^-21)) * uni)  =>  r < thresh * (1 - 2^-22)  =>  fl(r) <  thresh   (more than an ulp below),
    // (two roundings cost <= 2^-23 of the 2^-21 margin).  Disjoint boxes (inter = 0, uni > 0) leave through the second
    // test; near-ties, non-positive unions and NaNs fail both and take the division: the same decision bit for bit.
    const float hi = __fmul_rn(__fmul_rn(thresh, 1.00000047683715820312f), uni);      // 1 + 2^-21
    const float lo = __fmul_rn(__fmul_rn(thresh, 0.99999952316284179688f), uni);      // 1 - 2^-21
    if (inter >= hi && thresh > 0.f && uni > 0.f) return true;
    if (inter <= lo && thresh > 0.f && uni > 0.f) return false;
    return __fdiv_rn(inter, uni) >= thresh;
}

// ------------------------------------------------------------------------------------
struct RpnLevelDev {
    const float* scores;
    const float* deltas;
    int H, W, A;
    int n;            // A*H*W
    int take;         // boxes entering NMS for this level: min(pre_nms_topN or n, n)
    double stride;
    double anchors[4 * VOSD_MAX_ANCHORS];
};
struct RpnParams {
    RpnLevelDev lv[VOSD_MAX_LEVELS];
    int num_levels, num_images;
    int seg_stride;   // M: rows reserved per segment in the workspace
    int sort_cap;     // power of two >= M
    int cand_cap;     // keys of the candidate buffer behind the sort buffer (per CTA)
    float min_size;
    double xform_clip;
};

struct RpnKeys {
    const float* s;   // scores of one image at one level, (A, H*W)
    int A, HW;
    __device__ __forceinline__ uint64_t operator()(int j) const {
        // j / HW through fp32 (exact for j < 2^22: (j + 0.5) / HW stays >= 0.5 / HW away from every integer)
        const int a = j < (1 << 22) ? __float2int_rd(__fdividef((float)j + 0.5f, (float)HW)) : j / HW;
        const uint32_t flat = (uint32_t)(j - a * HW) * (uint32_t)A + (uint32_t)a;   // (h, w, a) order
        return ((uint64_t)float_to_ordered(__ldg(s + j)) << 32) | (uint64_t)(0xffffffffu - flat);
    }
};

// The proposal behind one ranked key: anchor of (h, w, a) shifted in fp64 (generate_proposals.py:69-89), decoded with
// unit weights and clipped to the image; `score` gets the key's score.
__device__ __forceinline__ float4 decode_ranked(const RpnLevelDev& L, const float* __restrict__ dl, int HW, uint64_t k,
                                                double xform_clip, float im_h, float im_w, float& score) {
    score = ordered_to_float((uint32_t)(k >> 32));
    const uint32_t flat = 0xffffffffu - (uint32_t)k;
    const int a = (int)(flat % (uint32_t)L.A);
    const int pos = (int)(flat / (uint32_t)L.A);
    const int h = pos / L.W, w = pos - h * L.W;
    const double sx = __dmul_rn((double)w, L.stride), sy = __dmul_rn((double)h, L.stride);
    const float ax1 = (float)__dadd_rn(L.anchors[4 * a + 0], sx), ay1 = (float)__dadd_rn(L.anchors[4 * a + 1], sy);
    const float ax2 = (float)__dadd_rn(L.anchors[4 * a + 2], sx), ay2 = (float)__dadd_rn(L.anchors[4 * a + 3], sy);
    const float* d = dl + (size_t)(4 * a) * HW + pos;
    const float4 box = decode_box(ax1, ay1, ax2, ay2, __ldg(d), __ldg(d + HW), __ldg(d + 2 * HW), __ldg(d + 3 * HW),
                                  1.f, 1.f, 1.f, 1.f, xform_clip);
    return clip_box(box, im_h, im_w);
}

// K1.  One thread-block CLUSTER of kTopkCluster CTAs per (level, image) segment: the score
// planes are streamed by 8 SMs at once, histograms meet in distributed shared memory, rank 0
// sorts the survivors and decodes them.  grid = segments * kTopkCluster, block = 1024,
// dyn smem = sort_cap * 8.
#ifndef VOSD_TOPK_CLUSTER
#define VOSD_TOPK_CLUSTER 4
#endif
constexpr int kTopkCluster = VOSD_TOPK_CLUSTER;
__global__ void __cluster_dims__(kTopkCluster, 1, 1) __launch_bounds__(kSelThreads, 1)
topk_decode_kernel(const __grid_constant__ RpnParams p, const float* __restrict__ im_info,
                   float4* __restrict__ ws_boxes, float* __restrict__ ws_scores,
                   int* __restrict__ ws_count) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(dyn);
    __shared__ ClusterSelectShared sh;
    cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();

    const int seg = blockIdx.x / kTopkCluster;
    const int l = seg / p.num_images, img = seg - l * p.num_images;
    const RpnLevelDev& L = p.lv[l];
    const int HW = L.H * L.W;
    RpnKeys kf{L.scores + (size_t)img * L.n, L.A, HW};
    const int P = next_pow2(L.take);
    uint64_t* cand = keys + p.sort_cap;                  // candidate keys of the threshold bin (select_sort.cuh)
    const int m = select_and_sort_cluster(cluster, kf, L.n, L.take, keys, P, sh, cand, p.cand_cap);
    if (cluster.block_rank() != 0) return;

    const float im_h = im_info[img * 3 + 0], im_w = im_info[img * 3 + 1], im_s = im_info[img * 3 + 2];
    const float min_size = __fmul_rn(p.min_size, im_s);
    const float* __restrict__ dl = L.deltas + (size_t)img * 4 * L.n;
    float4* ob = ws_boxes + (size_t)seg * p.seg_stride;
    float* os = ws_scores + (size_t)seg * p.seg_stride;

    // decode + clip + filter, then a stable compaction in rank order
    int base = 0;
    for (int t0 = 0; t0 < m; t0 += kSelThreads) {
        const int t = t0 + threadIdx.x;
        bool keep = false;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float score = 0.f;
        if (t < m) {
            box = decode_ranked(L, dl, HW, keys[t], p.xform_clip, im_h, im_w, score);
            keep = keep_box(box, min_size, im_h, im_w);
        }
        int total;
        const int off = block_exclusive_scan(keep ? 1 : 0, sh.warp_sums, total);
        if (keep) { ob[base + off] = box; os[base + off] = score; }
        base += total;
    }
    if (threadIdx.x == 0) ws_count[seg] = base;
}

// K2.  grid = (tiles of the upper triangle, 1, segments), block = 64.  Bit j of mask[i][c] is set
// when sorted box i suppresses sorted box 64c+j (only j-global > i is ever computed).  Only the W (W + 1) / 2 tiles
// with cb >= rb are launched (nms_mask_grid): a square grid spends half its blocks on a load of count[] and a return.
__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes, const int* __restrict__ count, int seg_stride,
                int words_per_row, float thresh, unsigned long long* __restrict__ mask) {
    const int seg = blockIdx.z;
    int t = blockIdx.x, rb = 0;                        // linear tile index -> (row block, column block >= row block)
    while (t >= words_per_row - rb) { t -= words_per_row - rb; rb++; }
    const int cb = rb + t;
    const int n = count[seg];
    if (cb * 64 >= n) return;                          // (rb <= cb)
    const float4* b = boxes + (size_t)seg * seg_stride;
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    const int cj = cb * 64 + threadIdx.x;
    if (cj < n) {
        const float4 v = b[cj];
        cbox[threadIdx.x] = v;
        carea[threadIdx.x] = box_area(v);
    }
    __syncthreads();
    const int i = rb * 64 + threadIdx.x;
    if (i >= n) return;
    const float4 a = b[i];
    const float area = box_area(a);
    const int ncol = min(64, n - cb * 64);
    unsigned long long bits = 0;
    const int start = (cb == rb) ? threadIdx.x + 1 : 0;
    for (int j = start; j < ncol; j++)
        if (suppresses(a, area, cbox[j], carea[j], thresh)) bits |= 1ull << j;
    mask[((size_t)seg * seg_stride + i) * words_per_row + cb] = bits;
}

static inline dim3 nms_mask_grid(int words, int segs) { return dim3((unsigned)(words * (words + 1) / 2), 1, (unsigned)segs); }

// K3: greedy reduce over the bitmask in sorted order, 64 boxes at a time: the chunk is resolved against itself, then
// the rows of its survivors are OR-ed into the running `removed` bitmap.
//   mode 0: write [img, box] / score of the first `post` kept boxes (proposal path);
//   mode 1: set keep_flag[orig_index[pos]] for every kept box (standalone nms);
//   mode 2: as 1 with orig_index / keep_flag laid out per segment (class-segmented nms, detections.cuh).
constexpr int kMaxWords = VOSD_MAX_TOPK / 64;

// K3, CTA-wide path for segments whose whole bitmask fits shared memory (rows * words * 8 bytes <= 128 KB, i.e. up to
// 1024 boxes: the TEST-mode RPN segments and the per-class segments of the box head).  A one-warp reduce (round 1) is
// bound by the latency of their global loads (~1.5 us per 64-box chunk); here 512 threads first copy the segment's upper
// triangle into shared memory in one coalesced sweep, warp 0 then resolves chunk after chunk from shared memory (64
// dependent steps per chunk on the diagonal word, then lanes = later words OR the survivors' rows), and all threads
// emit the kept boxes.  Same result as the warp kernels bit for bit (the greedy order is the same).
constexpr int kNmsCtaThreads = 512;
constexpr int kNmsCtaMaxWords = 16;
__global__ void __launch_bounds__(kNmsCtaThreads)
nms_reduce_cta_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                      const int* __restrict__ count, int seg_stride, int words_per_row,
                      const unsigned long long* __restrict__ mask, int use_mask, int post, int mode,
                      int num_images, int cap, float* __restrict__ out_rois, float* __restrict__ out_probs,
                      int* __restrict__ out_count, const int* __restrict__ orig_index,
                      int* __restrict__ keep_flag) {
    extern __shared__ __align__(16) unsigned long long sm_mask[];       // [n][nblk]
    __shared__ unsigned long long kept_w[kNmsCtaMaxWords];
    __shared__ int kept_prefix[kNmsCtaMaxWords + 1];
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = count[seg];
    const int nblk = (n + 63) / 64;
    const float4* b = boxes + (size_t)seg * seg_stride;
    const float* sc = scores ? scores + (size_t)seg * seg_stride : nullptr;
    const unsigned long long* mrow = mask + (size_t)seg * seg_stride * words_per_row;
    const int limit = post > 0 ? post : n;
    const float img = (float)(seg % num_images);
    if (mode == 2) { orig_index += (size_t)seg * seg_stride; keep_flag += (size_t)seg * seg_stride; }   // per-segment flags
    if (use_mask) {
        // only words w >= i / 64 of row i were written by nms_mask_kernel (upper triangle)
        // thread = (word w of the row, row i0 + k * rows-per-sweep): no division, four independent loads in flight
        const int w = tid % nblk, i0 = tid / nblk, rps = kNmsCtaThreads / nblk;
        if (i0 < rps) {
#pragma unroll 4
            for (int i = i0; i < n; i += rps)
                sm_mask[(size_t)i * nblk + w] = w >= (i >> 6) ? mrow[(size_t)i * words_per_row + w] : 0ull;
        }
    }
    __shared__ unsigned long long removed[kNmsCtaMaxWords];
    __shared__ unsigned long long kept_now;
    __shared__ int kept_sofar;
    if (tid < kNmsCtaMaxWords) removed[tid] = 0;
    if (tid == 0) kept_sofar = 0;
    __syncthreads();
    const int w_of = tid % kNmsCtaMaxWords, g_of = tid / kNmsCtaMaxWords;      // OR phase: thread = (later word, row group)
    for (int c = 0; c < nblk; c++) {
        const int i0 = c * 64;
        const int nin = min(64, n - i0);
        if (tid < 32) {
            // warp 0: the chunk against itself, 64 dependent steps on the diagonal word (shared-memory broadcasts)
            unsigned long long kept = nin < 64 ? (1ull << nin) - 1ull : ~0ull;
            const int before = kept_sofar;
            if (before >= limit) kept = 0;           // everything after the first `limit` survivors is dropped anyway
            else if (use_mask) {
                // Greedy result of the chunk = the unique fixpoint of K = alive & ~OR(rows of K): bit t of the right-hand
                // side only depends on bits < t of K (rows are upper-triangular), so iterating from K = alive fixes at least
                // one more leading bit per round and usually converges in a few rounds (the depth of the suppression
                // chains) instead of 64 dependent steps.  Lane L holds the diagonal words of boxes L and L + 32.
                const unsigned long long alive = ~removed[c] & kept;
                const unsigned long long* dcol = sm_mask + (size_t)i0 * nblk + c;
                const unsigned long long d0 = lane < nin ? dcol[(size_t)lane * nblk] : 0ull;
                const unsigned long long d1 = lane + 32 < nin ? dcol[(size_t)(lane + 32) * nblk] : 0ull;
                kept = alive;
                for (int round = 0; round < 65; round++) {
                    const unsigned long long sup = (((kept >> lane) & 1ull) ? d0 : 0ull) | (((kept >> (lane + 32)) & 1ull) ? d1 : 0ull);
                    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)sup);
                    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(sup >> 32));
                    const unsigned long long next = alive & ~(((unsigned long long)hi << 32) | lo);
                    if (next == kept) break;
                    kept = next;
                }
            }
            if (lane == 0) { kept_w[c] = kept; kept_prefix[c] = before; kept_now = kept; kept_sofar = before + __popcll(kept); }
        }
        __syncthreads();
        if (use_mask && c + 1 < nblk) {
            // all threads: OR the survivors' rows into the removed words of the later chunks
            const unsigned long long kept = kept_now;
            if (w_of > c && w_of < nblk) {
                unsigned long long acc = 0;
                for (int t = g_of; t < nin; t += kNmsCtaThreads / kNmsCtaMaxWords)
                    if ((kept >> t) & 1ull) acc |= sm_mask[(size_t)(i0 + t) * nblk + w_of];
                if (acc) atomicOr(&removed[w_of], acc);
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        kept_prefix[nblk] = kept_sofar;
        if (out_count) out_count[seg] = min(kept_sofar, limit);
    }
    __syncthreads();
    for (int i = tid; i < n; i += kNmsCtaThreads) {
        const int c = i >> 6, t = i & 63;
        const unsigned long long kept = kept_w[c];
        if ((kept >> t) & 1ull) {
            const int pos = kept_prefix[c] + __popcll(kept & ((1ull << t) - 1ull));
            if (pos < limit) {
                if (mode == 0) {
                    const float4 v = b[i];
                    float* o = out_rois + ((size_t)seg * cap + pos) * 5;
                    o[0] = img; o[1] = v.x; o[2] = v.y; o[3] = v.z; o[4] = v.w;
                    out_probs[(size_t)seg * cap + pos] = sc[i];
                } else {
                    keep_flag[orig_index[i]] = 1;
                }
            }
        }
    }
}

// K3, CTA-wide path for LONGER segments (1024 < boxes <= VOSD_MAX_TOPK: TRAIN-mode RPN segments of 2000, the
// standalone nms): the bitmask stays in global memory (L2), 512 threads share the work a one-warp reduce (round 1) did
// alone.  Per 64-box chunk: warp 0 resolves the diagonal word (fixpoint, as above) from two words per lane that were
// requested one chunk earlier; after one barrier every warp ORs the rows of ITS four boxes of the chunk, if they
// survived, into the shared `removed` bitmap (lanes = later words, coalesced), while warp 0 already requests the next
// chunk's diagonal words.  Two barriers and about one L2 round trip per chunk instead of ~4.5 us on one warp
// (2000-box TRAIN segments: 144 -> see DESIGN section 7).  Same greedy order: same result bit for bit.
__global__ void __launch_bounds__(kNmsCtaThreads)
nms_reduce_cta_global_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                             const int* __restrict__ count, int seg_stride, int words_per_row,
                             const unsigned long long* __restrict__ mask, int use_mask, int post, int mode,
                             int num_images, int cap, float* __restrict__ out_rois, float* __restrict__ out_probs,
                             int* __restrict__ out_count, const int* __restrict__ orig_index,
                             int* __restrict__ keep_flag) {
    __shared__ unsigned long long removed[kMaxWords], kept_w[kMaxWords];
    __shared__ int kept_prefix[kMaxWords + 1];
    __shared__ unsigned long long kept_now;
    __shared__ int kept_sofar;
    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = count[seg];
    const int nblk = (n + 63) / 64;
    const float4* b = boxes + (size_t)seg * seg_stride;
    const float* sc = scores ? scores + (size_t)seg * seg_stride : nullptr;
    const unsigned long long* mrow = mask + (size_t)seg * seg_stride * words_per_row;
    const int limit = post > 0 ? post : n;
    const float img = (float)(seg % num_images);
    if (mode == 2) { orig_index += (size_t)seg * seg_stride; keep_flag += (size_t)seg * seg_stride; }
    for (int w = tid; w < nblk; w += kNmsCtaThreads) { removed[w] = 0; kept_w[w] = 0; }
    if (tid == 0) kept_sofar = 0;
    __syncthreads();
    // warp 0: diagonal words of boxes `lane` and `lane + 32` of the current chunk (requested one chunk ahead)
    unsigned long long d0 = 0, d1 = 0;
    if (use_mask && warp == 0 && nblk > 0) {
        const int nin0 = min(64, n);
        d0 = lane < nin0 ? mrow[(size_t)lane * words_per_row] : 0ull;
        d1 = lane + 32 < nin0 ? mrow[(size_t)(lane + 32) * words_per_row] : 0ull;
    }
    int nchunks = 0;
    for (int c = 0; c < nblk; c++) {
        const int i0 = c * 64;
        const int nin = min(64, n - i0);
        if (kept_sofar >= limit) break;                  // uniform: read after the previous chunk's closing barrier
        nchunks = c + 1;
        if (warp == 0) {
            unsigned long long kept = nin < 64 ? (1ull << nin) - 1ull : ~0ull;
            if (use_mask) {
                const unsigned long long alive = ~removed[c] & kept;
                kept = alive;
                for (int round = 0; round < 65; round++) {
                    const unsigned long long sup = (((kept >> lane) & 1ull) ? d0 : 0ull) | (((kept >> (lane + 32)) & 1ull) ? d1 : 0ull);
                    const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)sup);
                    const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(sup >> 32));
                    const unsigned long long next = alive & ~(((unsigned long long)hi << 32) | lo);
                    if (next == kept) break;
                    kept = next;
                }
                if (c + 1 < nblk) {                      // next chunk's diagonal: in flight during the OR phase
                    const int j0 = i0 + 64, njn = min(64, n - j0);
                    d0 = lane < njn ? mrow[(size_t)(j0 + lane) * words_per_row + c + 1] : 0ull;
                    d1 = lane + 32 < njn ? mrow[(size_t)(j0 + lane + 32) * words_per_row + c + 1] : 0ull;
                }
            }
            if (lane == 0) { kept_w[c] = kept; kept_prefix[c] = kept_sofar; kept_now = kept; }
        }
        __syncthreads();
        if (use_mask && c + 1 < nblk) {
            const unsigned long long kept = kept_now;
            // warp j owns boxes 4j .. 4j+3 of the chunk; lanes = words c+1+lane, c+33+lane, ...
            for (int w = c + 1 + lane; w < nblk; w += 32) {
                unsigned long long acc = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int t = 4 * warp + k;
                    if ((kept >> t) & 1ull) acc |= mrow[(size_t)(i0 + t) * words_per_row + w];
                }
                if (acc) atomicOr(&removed[w], acc);
            }
        }
        if (tid == 0) kept_sofar += __popcll(kept_now);
        __syncthreads();
    }
    if (tid == 0) {
        kept_prefix[nchunks] = kept_sofar;
        if (out_count) out_count[seg] = min(kept_sofar, limit);
    }
    __syncthreads();
    for (int i = tid; i < min(n, nchunks * 64); i += kNmsCtaThreads) {
        const int c = i >> 6, t = i & 63;
        const unsigned long long kept = kept_w[c];
        if ((kept >> t) & 1ull) {
            const int pos = kept_prefix[c] + __popcll(kept & ((1ull << t) - 1ull));
            if (pos < limit) {
                if (mode == 0) {
                    const float4 v = b[i];
                    float* o = out_rois + ((size_t)seg * cap + pos) * 5;
                    o[0] = img; o[1] = v.x; o[2] = v.y; o[3] = v.z; o[4] = v.w;
                    out_probs[(size_t)seg * cap + pos] = sc[i];
                } else {
                    keep_flag[orig_index[i]] = 1;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------
// Segments with more than VOSD_MAX_TOPK boxes entering NMS (pre_nms_topN <= 0 or > 16384 on a large level: the full
// argsort branch of generate_proposals.py:131-132).  The n x n suppression bitmask of K2 would take gigabytes there, so
// these calls run ONE streamed kernel instead, one CTA per segment:
//   repeat: radix-select + sort the next 4096 best keys below the last one taken (a lazily produced full sort),
//           decode them 1024 at a time, drop the boxes an already kept box suppresses (kept boxes are read back from
//           the output), resolve the 1024 among themselves on a 1024 x 1024 bitmask in shared memory, append survivors;
//   until post_nms_topN boxes are kept or the `take` best keys are consumed.
// Same greedy order, same IoU arithmetic, same output layout as K1-K3.  Work is O(kept * n), memory O(1) per segment.
// ------------------------------------------------------------------------------------
constexpr int kStreamBatch = 4096;
constexpr int kStreamWords = kSelThreads / 64;            // 16 mask words per row of a 1024-box step
struct BoundedKeys {
    RpnKeys base;
    uint64_t bound;                                        // 0: no bound yet; else only keys < bound exist
    __device__ __forceinline__ uint64_t operator()(int j) const {
        const uint64_t k = base(j);
        return (bound == 0ull || k < bound) ? k : 0ull;
    }
};
constexpr size_t kStreamDyn = (size_t)kStreamBatch * 8 + (size_t)kSelThreads * kStreamWords * 8 + (size_t)kSelThreads * 20;

__global__ void __launch_bounds__(kSelThreads, 1)
proposals_stream_kernel(const __grid_constant__ RpnParams p, const float* __restrict__ im_info, float nms_thresh, int post,
                        int cap, float* __restrict__ out_rois, float* __restrict__ out_probs, int* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(dyn);                                            // [kStreamBatch]
    unsigned long long* smask = reinterpret_cast<unsigned long long*>(dyn + (size_t)kStreamBatch * 8);   // [1024][16]
    float4* cbox = reinterpret_cast<float4*>(dyn + (size_t)kStreamBatch * 8 + (size_t)kSelThreads * kStreamWords * 8);
    float* carea = reinterpret_cast<float*>(cbox + kSelThreads);
    float4* tile = reinterpret_cast<float4*>(smask);          // kept boxes of phase (a): the mask is built after it
    float* tarea = reinterpret_cast<float*>(tile + kSelThreads);
    __shared__ SelectShared sh;
    __shared__ unsigned long long removed[kStreamWords], kept_w[kStreamWords];
    __shared__ unsigned alive32[kSelThreads / 32];
    __shared__ int kept_prefix[kStreamWords + 1];

    const int seg = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int l = seg / p.num_images, img = seg - l * p.num_images;
    const RpnLevelDev& L = p.lv[l];
    const int HW = L.H * L.W;
    const float im_h = im_info[img * 3 + 0], im_w = im_info[img * 3 + 1], im_s = im_info[img * 3 + 2];
    const float min_size = __fmul_rn(p.min_size, im_s);
    const float* __restrict__ dl = L.deltas + (size_t)img * 4 * L.n;
    const bool use_nms = nms_thresh > 0.f;
    const int limit = (use_nms && post > 0) ? min(post, cap) : cap;
    float* rois = out_rois + (size_t)seg * cap * 5;
    float* probs = out_probs + (size_t)seg * cap;

    BoundedKeys kf{RpnKeys{L.scores + (size_t)img * L.n, L.A, HW}, 0ull};
    int consumed = 0, kept_total = 0;
    while (consumed < L.take && kept_total < limit) {
        const int batch = min(kStreamBatch, L.take - consumed);
        const int m = select_and_sort(kf, L.n, L.n - consumed, batch, keys, kStreamBatch, sh);
        __syncthreads();
        const uint64_t next_bound = keys[m - 1];
        for (int c0 = 0; c0 < m && kept_total < limit; c0 += kSelThreads) {
            const int t = c0 + tid;
            float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
            float score = 0.f;
            bool alive = false;
            if (t < m) {
                box = decode_ranked(L, dl, HW, keys[t], p.xform_clip, im_h, im_w, score);
                alive = keep_box(box, min_size, im_h, im_w);
            }
            const float area = box_area(box);
            cbox[tid] = box;
            carea[tid] = area;
            // (a) against the boxes kept so far
            if (use_nms) {
                for (int j0 = 0; j0 < kept_total; j0 += kSelThreads) {
                    const int nj = min(kSelThreads, kept_total - j0);
                    __syncthreads();
                    if (tid < nj) {
                        const float* r = rois + (size_t)(j0 + tid) * 5;
                        const float4 kb = make_float4(r[1], r[2], r[3], r[4]);
                        tile[tid] = kb;
                        tarea[tid] = box_area(kb);
                    }
                    __syncthreads();
                    if (alive)
                        for (int j = 0; j < nj; j++)
                            if (suppresses(tile[j], tarea[j], box, area, nms_thresh)) { alive = false; break; }
                }
            }
            __syncthreads();                                  // tile reads done: the region becomes the bitmask
            // (b) the step's boxes among themselves
            const unsigned bal = __ballot_sync(0xffffffffu, alive);
            if (lane == 0) alive32[tid >> 5] = bal;
            if (use_nms) {
                const int w0 = tid >> 6;
                for (int w = 0; w < kStreamWords; w++) {
                    unsigned long long bits = 0;
                    if (alive && w >= w0) {
                        const int jb = w * 64;
                        for (int j = (w == w0 ? (tid & 63) + 1 : 0); j < 64; j++)
                            if (suppresses(box, area, cbox[jb + j], carea[jb + j], nms_thresh)) bits |= 1ull << j;
                    }
                    smask[(size_t)tid * kStreamWords + w] = bits;
                }
            }
            __syncthreads();
            if (tid < kStreamWords) removed[tid] = ~((unsigned long long)alive32[2 * tid] | ((unsigned long long)alive32[2 * tid + 1] << 32));
            __syncthreads();
            for (int g = 0; g < kStreamWords; g++) {
                if (tid < 32) {
                    unsigned long long kept = ~removed[g];
                    if (use_nms) {
                        // greedy resolution of the 64 boxes of group g: the fixpoint of K = alive & ~OR(rows of K), as in K3
                        const unsigned long long al = kept;
                        const unsigned long long d0 = smask[(size_t)(g * 64 + lane) * kStreamWords + g];
                        const unsigned long long d1 = smask[(size_t)(g * 64 + lane + 32) * kStreamWords + g];
                        for (int round = 0; round < 65; round++) {
                            const unsigned long long sup = (((kept >> lane) & 1ull) ? d0 : 0ull) | (((kept >> (lane + 32)) & 1ull) ? d1 : 0ull);
                            const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)sup);
                            const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(sup >> 32));
                            const unsigned long long nx = al & ~(((unsigned long long)hi << 32) | lo);
                            if (nx == kept) break;
                            kept = nx;
                        }
                    }
                    if (lane == 0) kept_w[g] = kept;
                }
                __syncthreads();
                if (use_nms && g + 1 < kStreamWords) {
                    const int w_of = tid % kStreamWords, t_of = tid / kStreamWords;         // (later word, box of the group)
                    if (w_of > g && ((kept_w[g] >> t_of) & 1ull)) {
                        const unsigned long long row = smask[(size_t)(g * 64 + t_of) * kStreamWords + w_of];
                        if (row) atomicOr(&removed[w_of], row);
                    }
                    __syncthreads();
                }
            }
            if (tid == 0) {
                int acc = 0;
                for (int g = 0; g < kStreamWords; g++) { kept_prefix[g] = acc; acc += __popcll(kept_w[g]); }
                kept_prefix[kStreamWords] = acc;
            }
            __syncthreads();
            {
                const int g = tid >> 6, b = tid & 63;
                const unsigned long long kw = kept_w[g];
                if ((kw >> b) & 1ull) {
                    const int pos = kept_total + kept_prefix[g] + __popcll(kw & ((1ull << b) - 1ull));
                    if (pos < limit) {
                        float* r = rois + (size_t)pos * 5;
                        r[0] = (float)img; r[1] = box.x; r[2] = box.y; r[3] = box.z; r[4] = box.w;
                        probs[pos] = score;
                    }
                }
            }
            kept_total += kept_prefix[kStreamWords];
            __syncthreads();                                  // the appended rows are visible to phase (a) of the next step
        }
        consumed += m;
        kf.bound = next_bound;
    }
    if (tid == 0) out_count[seg] = min(kept_total, limit);
}

// dispatch of K3 (host): CTA-wide kernel when the segment's bitmask fits 128 KB of shared memory
static inline cudaError_t launch_nms_reduce(int segs, int seg_stride, int words, const float4* boxes, const float* scores,
                                            const int* count, const unsigned long long* mask, int use_mask, int post, int mode,
                                            int num_images, int cap, float* out_rois, float* out_probs, int* out_count,
                                            const int* orig_index, int* keep_flag, cudaStream_t stream) {
    const size_t dyn = (size_t)seg_stride * words * sizeof(unsigned long long);
    if (words <= kNmsCtaMaxWords && dyn <= 128 * 1024 && seg_stride >= 128) {
        cudaError_t e = cudaFuncSetAttribute(nms_reduce_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return e;
        nms_reduce_cta_kernel<<<segs, kNmsCtaThreads, dyn, stream>>>(boxes, scores, count, seg_stride, words, mask, use_mask, post, mode,
                                                                   num_images, cap, out_rois, out_probs, out_count, orig_index, keep_flag);
    } else {
        nms_reduce_cta_global_kernel<<<segs, kNmsCtaThreads, 0, stream>>>(boxes, scores, count, seg_stride, words, mask, use_mask, post,
                                                                          mode, num_images, cap, out_rois, out_probs, out_count,
                                                                          orig_index, keep_flag);
    }
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------
// Streaming decode of every anchor of a level (coalesced plane reads, 16-byte stores).
// grid-stride over (image, position); each thread handles all A anchors of one position.
// ------------------------------------------------------------------------------------
struct DecodeAllParams {
    const float* deltas;
    int H, W, A, N;
    double stride, clip;
    double anchors[4 * VOSD_MAX_ANCHORS];
    float fanchors[4 * VOSD_MAX_ANCHORS];     // kExact: the same anchors, exactly representable in fp32
    float fstride;
};
// kA > 0: the anchor count is a compile-time constant (A = 3, every FPN level): the anchor loop unrolls and the double-
// precision chains of the anchors of a position overlap; kA = 0: any A.
// kExact: every anchor coordinate is a multiple of 0.5, the stride an integer and all shifted coordinates stay below
// 2^22 (checked on the host), so "anchor + shift in fp64, rounded to fp32" (boxes.py:164) is an exact fp32 sum: the
// four fp64 additions and conversions per anchor are skipped with identical bits.
#ifndef VOSD_DECODE_MINB
#define VOSD_DECODE_MINB 3
#endif
template <int kA, bool kExact>
__global__ void __launch_bounds__(256, VOSD_DECODE_MINB)
decode_all_kernel(const __grid_constant__ DecodeAllParams p, const float* __restrict__ im_info, float4* __restrict__ out) {
    // grid = (position blocks, images): no 64-bit division per position
    const int HW = p.H * p.W;
    const int img = blockIdx.y;
    const float im_h = __ldg(im_info + img * 3), im_w = __ldg(im_info + img * 3 + 1);
    for (int pos = blockIdx.x * blockDim.x + threadIdx.x; pos < HW; pos += gridDim.x * blockDim.x) {
        const int h = pos / p.W, w = pos - h * p.W;
        const double sx = kExact ? 0.0 : __dmul_rn((double)w, p.stride), sy = kExact ? 0.0 : __dmul_rn((double)h, p.stride);
        const float fsx = __fmul_rn((float)w, p.fstride), fsy = __fmul_rn((float)h, p.fstride);
        const float* d = p.deltas + (size_t)img * 4 * p.A * HW + pos;
        const int A = kA > 0 ? kA : p.A;
        float4* o = out + ((size_t)img * HW + pos) * A;
#pragma unroll
        for (int a = 0; a < A; a++) {
            float ax1, ay1, ax2, ay2;
            if (kExact) {
                ax1 = __fadd_rn(p.fanchors[4 * a + 0], fsx); ay1 = __fadd_rn(p.fanchors[4 * a + 1], fsy);
                ax2 = __fadd_rn(p.fanchors[4 * a + 2], fsx); ay2 = __fadd_rn(p.fanchors[4 * a + 3], fsy);
            } else {
                ax1 = (float)__dadd_rn(p.anchors[4 * a + 0], sx); ay1 = (float)__dadd_rn(p.anchors[4 * a + 1], sy);
                ax2 = (float)__dadd_rn(p.anchors[4 * a + 2], sx); ay2 = (float)__dadd_rn(p.anchors[4 * a + 3], sy);
            }
            const float* da = d + (size_t)(4 * a) * HW;
            float4 box = decode_box(ax1, ay1, ax2, ay2, __ldg(da), __ldg(da + HW), __ldg(da + 2 * HW),
                                    __ldg(da + 3 * HW), 1.f, 1.f, 1.f, 1.f, p.clip);
            box = clip_box(box, im_h, im_w);
            st_stream_f4(o + a, box);
        }
    }
}

__global__ void __launch_bounds__(256)
any_nan_kernel(const float* __restrict__ x, size_t n, int* flag) {
    bool bad = false;
    const size_t n4 = n / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = ld_stream_f4(x4 + i);
        bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = x[n4 * 4 + threadIdx.x]; bad |= v != v; }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// ------------------------------------------------------------------------------------
// Standalone NMS helpers: sort (n,5) dets by score, then reuse K2/K3.
// ------------------------------------------------------------------------------------
struct DetKeys {
    const float* d;
    __device__ __forceinline__ uint64_t operator()(int j) const {
        return ((uint64_t)float_to_ordered(__ldg(d + 5 * (size_t)j + 4)) << 32) | (uint64_t)(0xffffffffu - (uint32_t)j);
    }
};
__global__ void __launch_bounds__(kSelThreads, 1)
nms_sort_kernel(const float* __restrict__ dets, int n, int P, float4* __restrict__ boxes,
                int* __restrict__ orig, int* __restrict__ count, int* __restrict__ keep_flag) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(dyn);
    __shared__ SelectShared sh;
    DetKeys kf{dets};
    select_and_sort(kf, n, n, n, keys, P, sh);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int j = (int)(0xffffffffu - (uint32_t)keys[t]);
        const float* r = dets + 5 * (size_t)j;
        boxes[t] = make_float4(r[0], r[1], r[2], r[3]);
        orig[t] = j;
        keep_flag[t] = 0;
    }
    if (threadIdx.x == 0) *count = n;
}

// keep = np.where(flag)[0] as int64 (ascending index), 1 CTA.
__global__ void __launch_bounds__(kSelThreads, 1)
compact_flags_kernel(const int* __restrict__ flag, int n, long long* __restrict__ keep, int* __restrict__ num_keep) {
    __shared__ int warp_sums[32];
    int base = 0;
    for (int t0 = 0; t0 < n; t0 += kSelThreads) {
        const int t = t0 + threadIdx.x;
        const int f = (t < n && flag[t]) ? 1 : 0;
        int total;
        const int off = block_exclusive_scan(f, warp_sums, total);
        if (f) keep[base + off] = t;
        base += total;
    }
    if (threadIdx.x == 0) *num_keep = base;
}

static int seg_take(const vosd_rpn_level& L, int pre) {
    const long long n = (long long)L.num_anchors * L.height * L.width;
    return (int)((pre <= 0 || pre >= n) ? n : pre);
}

struct PropLayout {
    int M, words, cap, S;
    bool stream;          // some level sends more than VOSD_MAX_TOPK boxes into NMS: proposals_stream_kernel
    size_t off_boxes, off_scores, off_count, off_mask, total;
};
static int prop_layout(const vosd_rpn_level* levels, int num_levels, int num_images, int pre, int post,
                       PropLayout& lay) {
    if (!levels) return VOSD_ERR_BAD_ARG;
    if (num_levels < 1 || num_levels > VOSD_MAX_LEVELS || num_images < 1) return VOSD_ERR_BAD_SHAPE;
    int M = 0;
    for (int l = 0; l < num_levels; l++) {
        const vosd_rpn_level& L = levels[l];
        if (L.height <= 0 || L.width <= 0 || L.num_anchors <= 0) return VOSD_ERR_BAD_SHAPE;
        if (L.num_anchors > VOSD_MAX_ANCHORS) return VOSD_ERR_UNSUPPORTED;
        if ((long long)L.num_anchors * L.height * L.width > 0x7fffffffLL / 8) return VOSD_ERR_UNSUPPORTED;
        const int t = seg_take(L, pre);
        M = t > M ? t : M;
    }
    lay.M = M;
    lay.stream = M > VOSD_MAX_TOPK;
    if (lay.stream) {
        lay.words = 0;
        lay.cap = post > 0 && post < M ? post : M;
        lay.S = num_levels * num_images;
        lay.off_boxes = lay.off_scores = lay.off_count = lay.off_mask = 0;
        lay.total = 256;                      // the streamed kernel keeps its state in shared memory and in the outputs
        return VOSD_OK;
    }
    lay.words = (M + 63) / 64;
    lay.cap = post > 0 && post < M ? post : M;
    lay.S = num_levels * num_images;
    size_t o = 0;
    lay.off_boxes = o;  o = align_up(o + (size_t)lay.S * M * sizeof(float4), 256);
    lay.off_scores = o; o = align_up(o + (size_t)lay.S * M * sizeof(float), 256);
    lay.off_count = o;  o = align_up(o + (size_t)lay.S * sizeof(int), 256);
    lay.off_mask = o;   o = align_up(o + (size_t)lay.S * M * lay.words * sizeof(unsigned long long), 256);
    lay.total = o;
    return VOSD_OK;
}

}  // namespace vosd

using namespace vosd;

extern "C" int vosd_proposals_capacity(const vosd_rpn_level* levels, int num_levels,
                                       int pre_nms_topN, int post_nms_topN) {
    PropLayout lay;
    const int rc = prop_layout(levels, num_levels, 1, pre_nms_topN, post_nms_topN, lay);
    return rc ? rc : lay.cap;
}

extern "C" size_t vosd_generate_proposals_workspace_bytes(const vosd_rpn_level* levels, int num_levels,
                                                          int num_images, int pre_nms_topN,
                                                          int post_nms_topN) {
    PropLayout lay;
    if (prop_layout(levels, num_levels, num_images, pre_nms_topN, post_nms_topN, lay)) return 0;
    return lay.total;
}

extern "C" int vosd_generate_proposals(const vosd_rpn_level* levels, int num_levels, int num_images,
                                       const float* im_info, int pre_nms_topN, int post_nms_topN,
                                       float nms_thresh, float min_size,
                                       float* out_rois, float* out_probs, int* out_count,
                                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    PropLayout lay;
    int rc = prop_layout(levels, num_levels, num_images, pre_nms_topN, post_nms_topN, lay);
    if (rc) return rc;
    if (!im_info || !out_rois || !out_probs || !out_count) return VOSD_ERR_BAD_ARG;
    if (!workspace || workspace_bytes < lay.total || !aligned16(workspace)) return VOSD_ERR_WORKSPACE;
    RpnParams p;
    for (int l = 0; l < num_levels; l++) {
        const vosd_rpn_level& L = levels[l];
        if (!L.scores || !L.deltas) return VOSD_ERR_BAD_ARG;
        RpnLevelDev& D = p.lv[l];
        D.scores = L.scores; D.deltas = L.deltas;
        D.H = L.height; D.W = L.width; D.A = L.num_anchors;
        D.n = L.num_anchors * L.height * L.width;
        D.take = seg_take(L, pre_nms_topN);
        D.stride = L.feat_stride;
        for (int k = 0; k < 4 * L.num_anchors; k++) D.anchors[k] = L.anchors[k];
    }
    p.num_levels = num_levels; p.num_images = num_images;
    p.seg_stride = lay.M; p.sort_cap = next_pow2(lay.M);
    p.min_size = min_size;
    p.xform_clip = log(1000.0 / 16.0);      // cfg.BBOX_XFORM_CLIP, core/config.py:1009
    char* ws = static_cast<char*>(workspace);
    float4* ws_boxes = reinterpret_cast<float4*>(ws + lay.off_boxes);
    float* ws_scores = reinterpret_cast<float*>(ws + lay.off_scores);
    int* ws_count = reinterpret_cast<int*>(ws + lay.off_count);
    unsigned long long* ws_mask = reinterpret_cast<unsigned long long*>(ws + lay.off_mask);

    // without NMS the reference keeps every filtered box (generate_proposals.py:159-166):
    // the caller must size the outputs with post_nms_topN = 0 in that case
    if (!(nms_thresh > 0.f) && lay.cap != lay.M) return VOSD_ERR_BAD_ARG;
    if (lay.stream) {
        if (cudaFuncSetAttribute(proposals_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStreamDyn) != cudaSuccess)
            return VOSD_ERR_LAUNCH;
        proposals_stream_kernel<<<lay.S, kSelThreads, kStreamDyn, stream>>>(p, im_info, nms_thresh, post_nms_topN, lay.cap,
                                                                          out_rois, out_probs, out_count);
        count_launch();
        return check_launch();
    }
    // candidate buffer: what is left of ~200 KB behind the sort buffer, at most 16384 keys per CTA
    long long cc = (200 * 1024 - (long long)p.sort_cap * 8) / 8;
    p.cand_cap = (int)(cc < 0 ? 0 : (cc > 16384 ? 16384 : cc));
    const size_t dyn = ((size_t)p.sort_cap + p.cand_cap) * sizeof(uint64_t);
    if (cudaFuncSetAttribute(topk_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    topk_decode_kernel<<<lay.S * kTopkCluster, kSelThreads, dyn, stream>>>(p, im_info, ws_boxes, ws_scores, ws_count);
    count_launch();
    const int use_mask = nms_thresh > 0.f;
    if (use_mask) {
        nms_mask_kernel<<<nms_mask_grid(lay.words, lay.S), 64, 0, stream>>>(ws_boxes, ws_count, lay.M, lay.words, nms_thresh, ws_mask);
        count_launch();
    }
    if (launch_nms_reduce(lay.S, lay.M, lay.words, ws_boxes, ws_scores, ws_count, ws_mask, use_mask, use_mask ? post_nms_topN : 0, 0,
                          num_images, lay.cap, out_rois, out_probs, out_count, nullptr, nullptr, stream) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    count_launch();
    return check_launch();
}

extern "C" int vosd_decode_anchors(const vosd_rpn_level* level, int num_images, const float* im_info,
                                   float* boxes, cudaStream_t stream) {
    if (!level || !level->deltas || !im_info || !boxes) return VOSD_ERR_BAD_ARG;
    if (level->height <= 0 || level->width <= 0 || level->num_anchors <= 0 || num_images <= 0) return VOSD_ERR_BAD_SHAPE;
    if (level->num_anchors > VOSD_MAX_ANCHORS) return VOSD_ERR_UNSUPPORTED;
    if (!aligned16(boxes)) return VOSD_ERR_BAD_ARG;
    DecodeAllParams p;
    p.deltas = level->deltas; p.H = level->height; p.W = level->width; p.A = level->num_anchors; p.N = num_images;
    p.stride = level->feat_stride; p.clip = log(1000.0 / 16.0);
    for (int k = 0; k < 4 * p.A; k++) p.anchors[k] = level->anchors[k];
    if (num_images > 65535) return VOSD_ERR_UNSUPPORTED;
    int blocks = ceil_div(p.H * p.W, 256);
    const int cap = ceil_div(kNumSMs * 32, num_images);
    if (blocks > cap) blocks = cap;
    // exact fp32 shifts?  (half-integer anchors, integer stride, everything below 2^22)
    bool exact = p.stride == floor(p.stride) && p.stride > 0 && p.stride * (double)(p.H > p.W ? p.H : p.W) < 4194304.0;
    for (int k = 0; k < 4 * p.A; k++) {
        exact = exact && 2.0 * p.anchors[k] == floor(2.0 * p.anchors[k]) && fabs(p.anchors[k]) < 4194304.0;
        p.fanchors[k] = (float)p.anchors[k];
    }
    p.fstride = (float)p.stride;
    const dim3 grid(blocks, num_images);
    float4* out = reinterpret_cast<float4*>(boxes);
    if (p.A == 3 && exact) decode_all_kernel<3, true><<<grid, 256, 0, stream>>>(p, im_info, out);
    else if (p.A == 3) decode_all_kernel<3, false><<<grid, 256, 0, stream>>>(p, im_info, out);
    else decode_all_kernel<0, false><<<grid, 256, 0, stream>>>(p, im_info, out);
    count_launch();
    return check_launch();
}

extern "C" int vosd_any_nan(const float* data, size_t n, int* flag, cudaStream_t stream) {
    if (!data || !flag) return VOSD_ERR_BAD_ARG;
    if (!aligned16(data)) return VOSD_ERR_BAD_ARG;
    if (n == 0) return VOSD_OK;
    size_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    any_nan_kernel<<<(int)blocks, 256, 0, stream>>>(data, n, flag);
    count_launch();
    return check_launch();
}

// workspace: boxes[n] float4 | orig[n] int | flag[n] int | count int | mask n*words u64
struct NmsLayout { size_t off_boxes, off_orig, off_flag, off_count, off_mask, total; int words; };
static NmsLayout nms_layout(int n) {
    NmsLayout L;
    L.words = (n + 63) / 64;
    size_t o = 0;
    L.off_boxes = o; o = align_up(o + (size_t)n * sizeof(float4), 256);
    L.off_orig = o;  o = align_up(o + (size_t)n * sizeof(int), 256);
    L.off_flag = o;  o = align_up(o + (size_t)n * sizeof(int), 256);
    L.off_count = o; o = align_up(o + sizeof(int), 256);
    L.off_mask = o;  o = align_up(o + (size_t)n * L.words * sizeof(unsigned long long), 256);
    L.total = o;
    return L;
}

extern "C" size_t vosd_nms_workspace_bytes(int n) {
    if (n <= 0) return 256;
    return nms_layout(n).total;
}

extern "C" int vosd_nms(const float* dets, int n, float thresh, int64_t* keep, int* num_keep,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n < 0) return VOSD_ERR_BAD_SHAPE;
    if (!num_keep) return VOSD_ERR_BAD_ARG;
    if (n == 0) return cudaMemsetAsync(num_keep, 0, sizeof(int), stream) == cudaSuccess ? VOSD_OK : VOSD_ERR_LAUNCH;
    if (n > VOSD_MAX_TOPK) return VOSD_ERR_UNSUPPORTED;
    if (!dets || !keep) return VOSD_ERR_BAD_ARG;
    const NmsLayout L = nms_layout(n);
    if (!workspace || workspace_bytes < L.total || !aligned16(workspace)) return VOSD_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    float4* boxes = reinterpret_cast<float4*>(ws + L.off_boxes);
    int* orig = reinterpret_cast<int*>(ws + L.off_orig);
    int* flag = reinterpret_cast<int*>(ws + L.off_flag);
    int* count = reinterpret_cast<int*>(ws + L.off_count);
    unsigned long long* mask = reinterpret_cast<unsigned long long*>(ws + L.off_mask);
    const int P = next_pow2(n);
    const size_t dyn = (size_t)P * sizeof(uint64_t);
    if (cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    nms_sort_kernel<<<1, kSelThreads, dyn, stream>>>(dets, n, P, boxes, orig, count, flag);
    nms_mask_kernel<<<nms_mask_grid(L.words, 1), 64, 0, stream>>>(boxes, count, n, L.words, thresh, mask);
    if (launch_nms_reduce(1, n, L.words, boxes, nullptr, count, mask, 1, 0, 1, 1, n, nullptr, nullptr, nullptr, orig, flag, stream) !=
        cudaSuccess)
        return VOSD_ERR_LAUNCH;
    compact_flags_kernel<<<1, kSelThreads, 0, stream>>>(flag, n, reinterpret_cast<long long*>(keep), num_keep);
    count_launch(4);
    return check_launch();
}

#include "detections.cuh"
