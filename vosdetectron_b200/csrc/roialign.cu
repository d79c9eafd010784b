// Multi-level RoIAlign forward / backward for sm_100a.
// Reference: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:65-121 (fwd), :195-270 (bwd),
// driven per FPN level by lib/modeling/model_builder.py:262-303.
#include <cstring>
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
// Staged forward path.  One CTA = one RoI x one group of output rows x several 32-channel slabs.
//   1. Per-axis tap tables (low/high texel, weights, validity) for the sample rows / columns are
//      built once per CTA, and from them one SAMPLE RECORD per bilinear sample: the four tile
//      offsets (bytes) and the four corner weights hy*hx, hy*lx, ly*hx, ly*lx.  They are shared by
//      every channel of the RoI, so the gather loop carries no coordinate arithmetic at all.
//   2. For each slab, the bounding tile of the feature map (rows y_lo..y_hi x columns x_lo..x_hi
//      of 32 channel planes) is staged in shared memory with 128-bit loads when rows are 16-byte
//      aligned (W % 4 == 0; scalar otherwise), in bands of output rows if it exceeds 512 texels.
//      Layout: word (t ^ c) of a 512-word row per channel c -- an XOR swizzle that keeps 16-byte
//      vectors intact (the 4 texels are permuted inside their vector) and makes the gather
//      bank-conflict-free: lanes = channels read the same texel t at banks (t ^ c) mod 32, all
//      distinct.  Address of a tap = K_lane ^ byte_offset: one LOP3 per shared-memory load.
//   3. Lanes own channels, warps own output bins; per-element arithmetic is the reference's
//      (roialign_math.cuh), so staging cannot change a bit of the result.
//   4. Results leave through obuf[c][bin] (odd stride) as contiguous streaming stores
//      (evict-first: the feature maps stay L2-resident for the RoIs that overlap them).
// RoIs whose tables / records exceed the budgets use direct global gathers inside the same kernel.
// ---------------------------------------------------------------------------------------
constexpr int kSlab = 32;
constexpr int kFwdThreads = 256;             // 8 warps x 128 registers, 2 CTAs / SM; each warp stages 4 channels
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kMaxTaps = 64;                 // per-axis table capacity (rows*grid_h, PW*grid_w)
constexpr int kMaxRecords = 1024;            // sample records per CTA
constexpr int kTileWords = 512;              // per channel; power of two (XOR addressing)

struct __align__(16) Tap { int low, high; float l, h; };   // low < 0: sample outside the map
struct __align__(16) SampleRec { int o1, o2, o3, o4; float w1, w2, w3, w4; };   // o1 < 0: contributes 0

struct FwdShared {
    Tap ytab[kMaxTaps];
    Tap xtab[kMaxTaps];
    RoiGeom g;
    int level, H, W;
    int x_lo, tw;            // tile columns (x_lo / tw aligned to 4 when vec)
    int vec;                 // 128-bit staging possible
    int ok;                  // 0 -> direct-gather fallback for this RoI
    int nbands;              // bands of output rows, each fitting the 512-texel tile
    short band_p1[kMaxTaps]; // end row (exclusive, relative to the row group) of band b
    short band_ylo[kMaxTaps];
    short band_rows[kMaxTaps];
    unsigned char band_tall[kMaxTaps];   // single output row taller than the tile: staged per sample row
};

// Texel j of an aligned 4-vector of channel c lives at position j ^ (c & 3) of that vector.
__device__ __forceinline__ float4 swizzle4(float4 v, int p) {
    switch (p) {
        case 0: return v;
        case 1: return make_float4(v.y, v.x, v.w, v.z);
        case 2: return make_float4(v.z, v.w, v.x, v.y);
        default: return make_float4(v.w, v.z, v.y, v.x);
    }
}

struct FwdCtx {
    const float* feat0;      // first channel of this CTA's first slab, image g.batch
    float* out0;             // first output of this CTA (row, first slab, first row of the group)
    float* tile; SampleRec* rec; float* obuf; int* binrec;
    int channels, pooled_w, bins, obuf_stride, slab0, nslab, H, W, gh, gw, nx;
    float count;
};

// Bands x slabs main loop.  kFast = 2x2 sampling grid, 128-bit staging, no tall rows: the common
// FPN case gets its own tightly register-allocated instance; everything else takes kFast = false.
template <bool kFast>
__device__ __noinline__ void run_bands(const FwdCtx cx, FwdShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* tile = cx.tile; SampleRec* rec = cx.rec; float* obuf = cx.obuf; int* binrec = cx.binrec;
    const int channels = cx.channels, pooled_w = cx.pooled_w, bins = cx.bins, obuf_stride = cx.obuf_stride;
    const int H = cx.H, W = cx.W, nx = cx.nx, slab0 = cx.slab0, nslab = cx.nslab;
    const int gh = kFast ? 2 : cx.gh, gw = kFast ? 2 : cx.gw;
    const int x_lo = sh.x_lo, tw = sh.tw;
    const bool vec = kFast ? true : (sh.vec != 0);
    const unsigned klane = (unsigned)(lane * kTileWords * 4) ^ (unsigned)(lane * 4);   // swizzled lane base (bytes)
    const char* tb = reinterpret_cast<const char*>(tile);
    const size_t plane = (size_t)H * W;
    const float* __restrict__ feat0 = cx.feat0;
    float* __restrict__ out0 = cx.out0;
    const int icount = gh * gw;
    const bool pow2 = (icount & (icount - 1)) == 0;       // x / 2^k == x * 2^-k exactly
    const float inv_count = 1.0f / cx.count;
    const int nbands = sh.nbands;

    int p0 = 0;                                            // rows are relative to ph_begin from here on
    for (int band = 0; band < nbands; band++) {
        const int p1 = sh.band_p1[band], y_lo = sh.band_ylo[band], rows_total = sh.band_rows[band];
        const bool tall = kFast ? false : (sh.band_tall[band] != 0);
        const int nsub = tall ? gh : 1;                    // tall row: one sample row (<= 2 texel rows) per pass
        const int band_bins = (p1 - p0) * pooled_w;
        // ---- once per band: sample records (tile byte offsets + corner weights), bin -> record base ----
        const int rec_n = (p1 - p0) * gh * nx;
        for (int s2 = tid; s2 < rec_n; s2 += kFwdThreads) {
            const int sy = s2 / nx, sx = s2 - sy * nx;
            const Tap ty = sh.ytab[p0 * gh + sy], tx = sh.xtab[sx];
            SampleRec r;
            if (ty.low >= 0 && tx.low >= 0) {
                const int base_y = tall ? ty.low : y_lo;            // tall: origin = the sample's own low row
                const int r0 = (ty.low - base_y) * tw - x_lo, r1 = (ty.high - base_y) * tw - x_lo;
                r.o1 = 4 * (r0 + tx.low); r.o2 = 4 * (r0 + tx.high);
                r.o3 = 4 * (r1 + tx.low); r.o4 = 4 * (r1 + tx.high);
                r.w1 = __fmul_rn(ty.h, tx.h); r.w2 = __fmul_rn(ty.h, tx.l);
                r.w3 = __fmul_rn(ty.l, tx.h); r.w4 = __fmul_rn(ty.l, tx.l);
            } else {
                r.o1 = -1; r.o2 = r.o3 = r.o4 = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
            }
            rec[s2] = r;
        }
        for (int q = tid; q < band_bins; q += kFwdThreads) {
            const int pr = q / pooled_w;
            binrec[q] = pr * gh * nx + (q - pr * pooled_w) * gw;
        }
        // ---- once per band: what this thread stages (same for every slab) ----
        int goff[4], toff[4];
        if (!tall) {
            const int tws = vec ? (tw >> 2) : tw;                   // row length in staging units
            const int nunit = rows_total * tws;
            const float inv = 1.0f / (float)max(tws, 1);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int i = lane + 32 * m;
                const int ry = (int)(((float)i + 0.5f) * inv);
                const int rx = vec ? ((i - ry * tws) << 2) : (i - ry * tws);
                goff[m] = i < nunit ? (y_lo + ry) * W + (x_lo + rx) : -1;
                toff[m] = ry * tw + rx;
            }
        }
        __syncthreads();

        for (int k = 0; k < nslab; k++) {
            const int nch = min(kSlab, channels - (slab0 + k) * kSlab);
            const float* __restrict__ feat = feat0 + (size_t)k * kSlab * plane;
            for (int sub = 0; sub < nsub; sub++) {
                int elems = rows_total * tw, sy_lo = y_lo;
                if (tall) {                                         // texel rows of sample row `sub` only
                    const Tap t = sh.ytab[p0 * gh + sub];
                    sy_lo = t.low >= 0 ? t.low : 0;
                    const int rows = (t.low >= 0 && tw > 0) ? t.high - t.low + 1 : 0;
                    elems = rows * tw;
                    const int tws = vec ? (tw >> 2) : tw;
                    const float inv = 1.0f / (float)max(tws, 1);
#pragma unroll
                    for (int m = 0; m < 4; m++) {
                        const int i = lane + 32 * m;
                        const int ry = (int)(((float)i + 0.5f) * inv);
                        const int rx = vec ? ((i - ry * tws) << 2) : (i - ry * tws);
                        goff[m] = i < rows * tws ? (sy_lo + ry) * W + (x_lo + rx) : -1;
                        toff[m] = ry * tw + rx;
                    }
                }
                // ---- stage: all of a thread's loads (2 channels x <= 4 units) are issued before any store ----
                if (vec) {
                    // one channel at a time: its <= 4 vectors are loaded before any is stored
                    for (int c = warp; c < nch; c += kFwdWarps) {
                        const float* pc = feat + c * plane;
                        float4 v[4];
#pragma unroll
                        for (int m = 0; m < 4; m++)
                            if (goff[m] >= 0) v[m] = __ldg(reinterpret_cast<const float4*>(pc + goff[m]));
                        float* tc = tile + c * kTileWords;
                        const int cx4 = c & ~3;
#pragma unroll
                        for (int m = 0; m < 4; m++)
                            if (goff[m] >= 0) *reinterpret_cast<float4*>(tc + (toff[m] ^ cx4)) = swizzle4(v[m], c & 3);
                    }
                } else {
                    // scalar rows (W % 4 != 0): 128 texels per pass, up to 4 passes
                    for (int i0 = 0; i0 < elems; i0 += 128) {
                        int go[4];
                        if (i0 == 0) {
#pragma unroll
                            for (int m = 0; m < 4; m++) go[m] = goff[m];
                        } else {
                            const float inv = 1.0f / (float)max(tw, 1);
#pragma unroll
                            for (int m = 0; m < 4; m++) {
                                const int i = i0 + lane + 32 * m;
                                const int ry = (int)(((float)i + 0.5f) * inv);
                                go[m] = i < elems ? (sy_lo + ry) * W + (x_lo + i - ry * tw) : -1;
                            }
                        }
                        for (int c = warp; c < nch; c += kFwdWarps) {
                            float va[4];
#pragma unroll
                            for (int m = 0; m < 4; m++)
                                if (go[m] >= 0) va[m] = __ldg(feat + c * plane + go[m]);
#pragma unroll
                            for (int m = 0; m < 4; m++)
                                if (go[m] >= 0) tile[c * kTileWords + ((i0 + lane + 32 * m) ^ c)] = va[m];
                        }
                    }
                }
                __syncthreads();
                // ---- gather: lanes = channels, warps = bins of the band ----
                if (kFast) {
                    for (int q = warp; q < band_bins; q += kFwdWarps) {
                        const SampleRec* rr = rec + binrec[q];
                        float acc = 0.f;
#pragma unroll
                        for (int iy = 0; iy < 2; iy++) {
#pragma unroll
                            for (int ix = 0; ix < 2; ix++) {
                                const SampleRec r = rr[iy * nx + ix];
                                float val = 0.f;
                                if (r.o1 >= 0) {
                                    const float v1 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o1));
                                    const float v2 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o2));
                                    const float v3 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o3));
                                    const float v4 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o4));
                                    val = __fmaf_rn(r.w4, v4, __fmaf_rn(r.w3, v3, __fmaf_rn(r.w1, v1, __fmul_rn(r.w2, v2))));
                                }
                                acc = __fadd_rn(acc, val);
                            }
                        }
                        obuf[lane * obuf_stride + q] = __fmul_rn(acc, 0.25f);          // count == 4: exact
                    }
                } else {
                    const int iy0 = tall ? sub : 0, iy1 = tall ? sub + 1 : gh;
                    for (int q = warp; q < band_bins; q += kFwdWarps) {
                        const SampleRec* rr0 = rec + binrec[q];
                        float* slot = obuf + lane * obuf_stride + q;
                        float acc = (tall && sub > 0) ? *slot : 0.f;     // tall rows carry the running sum in obuf
                        for (int iy = iy0; iy < iy1; iy++) {
                            const SampleRec* rr = rr0 + iy * nx;
                            for (int ix = 0; ix < gw; ix++) {
                                const SampleRec r = rr[ix];
                                float val = 0.f;
                                if (r.o1 >= 0) {
                                    const float v1 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o1));
                                    const float v2 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o2));
                                    const float v3 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o3));
                                    const float v4 = *reinterpret_cast<const float*>(tb + (klane ^ (unsigned)r.o4));
                                    val = __fmaf_rn(r.w4, v4, __fmaf_rn(r.w3, v3, __fmaf_rn(r.w1, v1, __fmul_rn(r.w2, v2))));
                                }
                                acc = __fadd_rn(acc, val);
                            }
                        }
                        *slot = (sub == nsub - 1) ? (pow2 ? __fmul_rn(acc, inv_count) : __fdiv_rn(acc, cx.count)) : acc;
                    }
                }
                __syncthreads();
            }
            // ---- contiguous streaming write of this band of the slab: warp -> its two channels ----
            float* __restrict__ out = out0 + (size_t)k * kSlab * bins + p0 * pooled_w;
            for (int c = warp; c < nch; c += kFwdWarps)
                for (int bq = lane; bq < band_bins; bq += 32)
                    __stcs(out + (size_t)c * bins + bq, obuf[c * obuf_stride + bq]);
        }
        __syncthreads();
        p0 = p1;
    }
}

// ---- shared-memory access by 32-bit shared-window address (keeps the XOR-swizzled address
//      arithmetic to a single LOP3 per tap and stops the compiler re-deriving generic pointers) ----
__device__ __forceinline__ float lds_f32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(unsigned a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void lds_rec(unsigned a, int& o1, int& o2, int& o3, int& o4, float& w1, float& w2, float& w3, float& w4) {
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o1), "=r"(o2), "=r"(o3), "=r"(o4) : "r"(a));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+16];" : "=f"(w1), "=f"(w2), "=f"(w3), "=f"(w4) : "r"(a));
}

// Ordered (volatile) read-only 128-bit global load: keeps channel a's loads/stores ahead of channel b's,
// i.e. 4 (not 8) vectors live per thread -- the kernel is capped at 64 registers for 2 CTAs / SM.
__device__ __forceinline__ float4 ldg_v4_ordered(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ void sts_v2(unsigned a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" :: "r"(a), "f"(x), "f"(y) : "memory");
}

// Stores vector v of a channel with c & 3 == p to its swizzled slot: texel j goes to position j ^ p.
// Bit 1 of p swaps the two 8-byte halves (done in the address: a_lo = slot ^ ((p & 2) * 4)), bit 0 swaps
// inside each half (2 x 2 selects on a warp-uniform predicate).  One code instance for all warps.
__device__ __forceinline__ void sts_swizzled(unsigned a_lo, bool swap, float4 v) {
    sts_v2(a_lo, swap ? v.y : v.x, swap ? v.x : v.y);
    sts_v2(a_lo ^ 8u, swap ? v.w : v.z, swap ? v.z : v.w);
}

constexpr int kChPerWarp = kSlab / kFwdWarps;      // 4

// Predicated (goff >= 0) ordered 128-bit read-only load: no branch around the volatile asm.
__device__ __forceinline__ void ldg_v4_pred(float4& v, const float* p, int goff) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %5, 0;\n\t@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(p + (goff >= 0 ? goff : 0)), "r"(goff));
}
// Predicated 128-bit shared store of channel 4*warp + H of the slab: its row is H * 2 KB further, its swizzle
// phase is H, so texel j goes to position j ^ H -- a compile-time permutation of the operands.
template <int H>
__device__ __forceinline__ void sts_v4_pred(unsigned a, const float4& v, int goff) {
    const float e0 = H == 0 ? v.x : H == 1 ? v.y : H == 2 ? v.z : v.w;
    const float e1 = H == 0 ? v.y : H == 1 ? v.x : H == 2 ? v.w : v.z;
    const float e2 = H == 0 ? v.z : H == 1 ? v.w : H == 2 ? v.x : v.y;
    const float e3 = H == 0 ? v.w : H == 1 ? v.z : H == 2 ? v.y : v.x;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0+%6], {%1, %2, %3, %4};\n\t}"
                 :: "r"(a), "f"(e0), "f"(e1), "f"(e2), "f"(e3), "r"(goff), "n"(H * kTileWords * 4) : "memory");
}

// A warp stages channels 4*warp .. 4*warp+3 of the slab: same XOR group (c & ~3 = 4*warp), phases 0..3.
__device__ __forceinline__ void prefetch_slab(float4 (&v)[kChPerWarp][4], const float* __restrict__ p0, size_t plane,
                                              int mcnt, const int (&goff)[4]) {
#pragma unroll
    for (int m = 0; m < 4; m++) {
        if (m < mcnt) {                      // warp-uniform
#pragma unroll
            for (int h = 0; h < kChPerWarp; h++) ldg_v4_pred(v[h][m], p0 + (size_t)h * plane, goff[m]);
        }
    }
}

__device__ __forceinline__ void commit_slab(const float4 (&v)[kChPerWarp][4], int mcnt, const int (&goff)[4],
                                            const unsigned (&ts0)[4]) {
#pragma unroll
    for (int m = 0; m < 4; m++) {
        if (m < mcnt) {
            sts_v4_pred<0>(ts0[m], v[0][m], goff[m]);
            sts_v4_pred<1>(ts0[m], v[1][m], goff[m]);
            sts_v4_pred<2>(ts0[m], v[2][m], goff[m]);
            sts_v4_pred<3>(ts0[m], v[3][m], goff[m]);
        }
    }
}

// Fast path of the bands x slabs loop: 2x2 sampling grid, 16-byte aligned rows, no tall rows, C % 32 == 0.
//
// Software pipeline over the slabs of the CTA: the 16 vectors a thread stages for slab k+1
// (channels 4*warp..4*warp+3 x <= 4 vectors, 64 registers) are requested from L2/HBM BEFORE the gather of
// slab k and written to the tile after it, so the global-load latency hides behind the gather instead of
// stalling all warps at a barrier.  8 warps x 128 registers, 2 CTAs per SM.  The four channels of a warp share
// the XOR group (c & ~3) and have swizzle phases 0..3, so the in-vector permutation is static register renaming.
__device__ __noinline__ void run_bands_fast(const FwdCtx cx, FwdShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pooled_w = cx.pooled_w, bins = cx.bins, obuf_stride = cx.obuf_stride;
    const int W = cx.W, nx = cx.nx, nslab = cx.nslab;
    const int x_lo = sh.x_lo, tw = sh.tw;
    const size_t plane = (size_t)cx.H * W;
    const unsigned tile_s = (unsigned)__cvta_generic_to_shared(cx.tile);      // 2 KB aligned
    const unsigned rec_s = (unsigned)__cvta_generic_to_shared(cx.rec);
    const unsigned obuf_s = (unsigned)__cvta_generic_to_shared(cx.obuf);
    const unsigned klane = (tile_s + (unsigned)lane * (kTileWords * 4)) ^ ((unsigned)lane * 4u);
    const int nbands = sh.nbands;
    const int tw4 = tw >> 2;
    const unsigned nxb = (unsigned)nx * 32u;                // bytes between the two sample rows of a bin

    int p0 = 0;
    for (int band = 0; band < nbands; band++) {
        const int p1 = sh.band_p1[band], y_lo = sh.band_ylo[band], rows_total = sh.band_rows[band];
        const int band_bins = (p1 - p0) * pooled_w;
        // ---- what this thread stages for every slab of the band ----
        int goff[4];
        unsigned ts0[4];
        {
            const int nvec = rows_total * tw4;
            const float inv = 1.0f / (float)max(tw4, 1);
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int i = lane + 32 * m;
                const int ry = (int)(((float)i + 0.5f) * inv);
                const int rx = (i - ry * tw4) << 2;
                goff[m] = i < nvec ? (y_lo + ry) * W + (x_lo + rx) : -1;
                ts0[m] = tile_s + 4u * (unsigned)(4 * warp * kTileWords + ((ry * tw + rx) ^ (4 * warp)));
            }
        }
        const int mcnt = (rows_total * tw4 + 31) >> 5;             // 32-vector groups in use (warp-uniform)
        const float* pw0 = cx.feat0 + (size_t)(4 * warp) * plane;  // channel 4*warp of the current slab
        float4 v[kChPerWarp][4];
        prefetch_slab(v, pw0, plane, mcnt, goff);                  // slab 0 in flight
        // ---- once per band: sample records, bin -> first record (sign bit: some sample is outside) ----
        const int rec_n = (p1 - p0) * 2 * nx;
        for (int s2 = tid; s2 < rec_n; s2 += kFwdThreads) {
            const int sy = s2 / nx, sx = s2 - sy * nx;
            const Tap ty = sh.ytab[p0 * 2 + sy], tx = sh.xtab[sx];
            SampleRec r;
            if (ty.low >= 0 && tx.low >= 0) {
                const int r0 = (ty.low - y_lo) * tw - x_lo, r1 = (ty.high - y_lo) * tw - x_lo;
                r.o1 = 4 * (r0 + tx.low); r.o2 = 4 * (r0 + tx.high);
                r.o3 = 4 * (r1 + tx.low); r.o4 = 4 * (r1 + tx.high);
                r.w1 = __fmul_rn(ty.h, tx.h); r.w2 = __fmul_rn(ty.h, tx.l);
                r.w3 = __fmul_rn(ty.l, tx.h); r.w4 = __fmul_rn(ty.l, tx.l);
            } else {
                r.o1 = -1; r.o2 = r.o3 = r.o4 = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
            }
            cx.rec[s2] = r;
        }
        for (int q = tid; q < band_bins; q += kFwdThreads) {
            const int pr = q / pooled_w, pw = q - pr * pooled_w;
            const bool all_in = sh.ytab[(p0 + pr) * 2].low >= 0 && sh.ytab[(p0 + pr) * 2 + 1].low >= 0 &&
                                sh.xtab[pw * 2].low >= 0 && sh.xtab[pw * 2 + 1].low >= 0;
            const int first = pr * 2 * nx + pw * 2;
            cx.binrec[q] = all_in ? first : (first | (int)0x80000000);
        }
        commit_slab(v, mcnt, goff, ts0);
        __syncthreads();                                            // tile(slab 0), records, binrec visible

        float* out = cx.out0 + p0 * pooled_w;
        for (int k = 0; k < nslab; k++) {
            const int mnext = k + 1 < nslab ? mcnt : 0;
            pw0 += (size_t)kSlab * plane;
            prefetch_slab(v, pw0, plane, mnext, goff);                         // slab k+1 -> registers
            // ---- gather slab k: lanes = channels, warps = bins ----
            for (int q = warp; q < band_bins; q += kFwdWarps) {
                const int br = cx.binrec[q];
                const unsigned ra = rec_s + (unsigned)(br & 0x7fffffff) * 32u;
                float acc = 0.f;
                if (br >= 0) {
                    // all four samples inside the map.  Two samples at a time: both records, then the 8 taps,
                    // then the arithmetic, so shared-memory latencies overlap instead of chaining per sample.
#pragma unroll
                    for (int iy = 0; iy < 2; iy++) {
                        int o[2][4]; float w[2][4], t[2][4];
#pragma unroll
                        for (int ix = 0; ix < 2; ix++)
                            lds_rec(ra + (unsigned)iy * nxb + (unsigned)ix * 32u, o[ix][0], o[ix][1], o[ix][2], o[ix][3],
                                    w[ix][0], w[ix][1], w[ix][2], w[ix][3]);
#pragma unroll
                        for (int ix = 0; ix < 2; ix++)
#pragma unroll
                            for (int c4 = 0; c4 < 4; c4++) t[ix][c4] = lds_f32(klane ^ (unsigned)o[ix][c4]);
#pragma unroll
                        for (int ix = 0; ix < 2; ix++)
                            acc = __fadd_rn(acc, __fmaf_rn(w[ix][3], t[ix][3], __fmaf_rn(w[ix][2], t[ix][2],
                                                           __fmaf_rn(w[ix][0], t[ix][0], __fmul_rn(w[ix][1], t[ix][1])))));
                    }
                } else {
#pragma unroll
                    for (int s4 = 0; s4 < 4; s4++) {
                        int o1, o2, o3, o4; float w1, w2, w3, w4;
                        lds_rec(ra + (unsigned)(s4 >> 1) * nxb + (unsigned)(s4 & 1) * 32u, o1, o2, o3, o4, w1, w2, w3, w4);
                        float val = 0.f;
                        if (o1 >= 0) {
                            const float v1 = lds_f32(klane ^ (unsigned)o1), v2 = lds_f32(klane ^ (unsigned)o2);
                            const float v3 = lds_f32(klane ^ (unsigned)o3), v4 = lds_f32(klane ^ (unsigned)o4);
                            val = __fmaf_rn(w4, v4, __fmaf_rn(w3, v3, __fmaf_rn(w1, v1, __fmul_rn(w2, v2))));
                        }
                        acc = __fadd_rn(acc, val);
                    }
                }
                sts_f32(obuf_s + 4u * (unsigned)(lane * obuf_stride + q), __fmul_rn(acc, 0.25f));   // count == 4: exact
            }
            __syncthreads();                                        // every warp is done reading the tile
            commit_slab(v, mnext, goff, ts0);                        // slab k+1 -> tile
            // ---- contiguous streaming write of slab k: warp -> its 4 channels (4*warp ..) ----
            if (band_bins == bins && obuf_stride == bins) {          // one band, odd bin count: 4*bins contiguous floats
                float* oc = out + (size_t)(4 * warp) * bins;
                const float* ob = cx.obuf + 4 * warp * obuf_stride;
                for (int i = lane; i < kChPerWarp * bins; i += 32) __stcs(oc + i, ob[i]);
            } else {
#pragma unroll
                for (int h = 0; h < kChPerWarp; h++) {
                    float* oc = out + (size_t)(4 * warp + h) * bins;
                    const float* ob = cx.obuf + (4 * warp + h) * obuf_stride;
                    for (int bq = lane; bq < band_bins; bq += 32) __stcs(oc + bq, ob[bq]);
                }
            }
            out += (size_t)kSlab * bins;
            __syncthreads();                                        // tile(slab k+1) visible, obuf free
        }
        p0 = p1;
    }
}

// Per-CTA setup shared by the staged forward and backward kernels: RoI geometry, per-axis tap tables for
// the sample rows of [ph_begin, ph_end) and all sample columns, the column extent of the tile (aligned to
// 4 texels when rows are 16-byte aligned), and the list of bands of output rows that fit the tile.
// Ends with a barrier; afterwards sh.* is read-only.
__device__ __forceinline__ void roi_setup(FwdShared& sh, const LevelTable& lv, const float* __restrict__ rois,
                                          const int* __restrict__ roi_level, int n, int pooled_h, int pooled_w,
                                          int sampling_ratio, int ph_begin, int ph_end, int rec_cap) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const int l = roi_level ? roi_level[n] : 0;
        sh.level = l; sh.H = lv.h[l]; sh.W = lv.w[l];
        sh.g = roi_geometry(rois + 5 * (size_t)n, lv.scale[l], pooled_h, pooled_w, sampling_ratio);
    }
    __syncthreads();
    const RoiGeom g = sh.g;
    const int H = sh.H, W = sh.W;
    const int gh = g.grid_h, gw = g.grid_w;
    const int ny = (ph_end - ph_begin) * gh, nx = pooled_w * gw;      // sample rows of this group / columns
    const bool fits = ny <= kMaxTaps && nx <= kMaxTaps && ny * nx <= rec_cap;
    if (fits) {
        if (tid < ny) {
            const int sy = ph_begin * gh + tid;
            const AxisTap t = axis_tap(sample_coord(g.start_h, g.bin_h, sy / gh, sy % gh, gh), H);
            sh.ytab[tid] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        } else if (tid >= 64 && tid < 64 + nx) {
            const int k = tid - 64;
            const AxisTap t = axis_tap(sample_coord(g.start_w, g.bin_w, k / gw, k % gw, gw), W);
            sh.xtab[k] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        }
    }
    __syncthreads();
    if (warp == 0) {
        // column extent of the tile (warp reduction over the x taps)
        int lo = 1 << 30, hi = -1;
        if (fits)
            for (int k = lane; k < nx; k += 32)
                if (sh.xtab[k].low >= 0) { lo = min(lo, sh.xtab[k].low); hi = max(hi, sh.xtab[k].high); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        const int l = sh.level;
        const bool vec = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(lv.data[l]) & 15) == 0;
        if (vec && hi >= lo) { lo &= ~3; hi |= 3; }
        const int tw = hi - lo + 1;                                  // <= 0: no valid column
        // every sample row must fit the tile on its own (2 texel rows)
        int ok = fits && (tw <= 0 || 2 * tw <= kTileWords);
        if (lane == 0) {
            sh.x_lo = lo; sh.tw = tw; sh.vec = vec; sh.ok = ok;
            int nb = 0;
            if (ok) {
                const int nrows = ph_end - ph_begin;
                int p0 = 0;
                while (p0 < nrows) {
                    int p1 = p0, y_lo = 1 << 30, y_hi = -1, tall = 0;
                    while (p1 < nrows) {
                        int l2 = y_lo, h2 = y_hi;
                        for (int i = 0; i < gh; i++) {
                            const Tap t = sh.ytab[p1 * gh + i];
                            if (t.low >= 0) { l2 = min(l2, t.low); h2 = max(h2, t.high); }
                        }
                        const bool over = h2 >= l2 && tw > 0 && (h2 - l2 + 1) * tw > kTileWords;
                        if (over && p1 > p0) break;
                        y_lo = l2; y_hi = h2; p1++;
                        if (over) { tall = 1; break; }
                    }
                    sh.band_p1[nb] = (short)p1;
                    sh.band_ylo[nb] = (short)(y_hi >= y_lo ? y_lo : 0);
                    sh.band_rows[nb] = (short)((y_hi >= y_lo && tw > 0) ? y_hi - y_lo + 1 : 0);
                    sh.band_tall[nb] = (unsigned char)tall;
                    nb++;
                    p0 = p1;
                }
            }
            sh.nbands = nb;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kFwdThreads, 2)
roialign_fwd_staged(const __grid_constant__ LevelTable lv, int channels, int pooled_h, int pooled_w,
                    int sampling_ratio, int rows_per_group, int slabs_per_cta, int rec_cap, int obuf_stride,
                    const float* __restrict__ rois, const int* __restrict__ roi_level,
                    const int* __restrict__ out_index, float* __restrict__ top) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    // the tile starts on a 2 KB boundary of the shared window: tap address = K_lane ^ offset (one LOP3)
    unsigned char* dyn = dyn_raw + ((2048u - ((unsigned)__cvta_generic_to_shared(dyn_raw) & 2047u)) & 2047u);
    float* tile = reinterpret_cast<float*>(dyn);                                        // [32][512] swizzled
    SampleRec* rec = reinterpret_cast<SampleRec*>(dyn + kSlab * kTileWords * 4);       // [rec_cap]
    float* obuf = reinterpret_cast<float*>(dyn + kSlab * kTileWords * 4 + (size_t)rec_cap * sizeof(SampleRec));
    int* binrec = reinterpret_cast<int*>(obuf + kSlab * obuf_stride);                   // [rows_per_group * pooled_w]
    __shared__ FwdShared sh;

    const int n = blockIdx.x;
    const int tid = threadIdx.x;
    const int ph_begin = blockIdx.z * rows_per_group;
    const int ph_end = min(pooled_h, ph_begin + rows_per_group);
    const int bins = pooled_h * pooled_w;

    roi_setup(sh, lv, rois, roi_level, n, pooled_h, pooled_w, sampling_ratio, ph_begin, ph_end, rec_cap);
    const RoiGeom g = sh.g;
    const int H = sh.H, W = sh.W;
    const int gh = g.grid_h, gw = g.grid_w;
    const int nx = pooled_w * gw;
    const int row = out_index ? out_index[n] : n;
    const int group_bins = (ph_end - ph_begin) * pooled_w;

    if (!sh.ok) {
        // direct-gather fallback: thread per (channel, bin) of this row group, all slabs of this CTA
        const int c_begin = blockIdx.y * slabs_per_cta * kSlab;
        const int c_end = min(channels, c_begin + slabs_per_cta * kSlab);
        const float* fbase = lv.data[sh.level] + (size_t)g.batch * channels * H * W;
        for (int e = tid; e < (c_end - c_begin) * group_bins; e += kFwdThreads) {
            const int c = c_begin + e / group_bins, b = e % group_bins;
            const int ph = ph_begin + b / pooled_w, pw = b % pooled_w;
            const float* d = fbase + (size_t)c * H * W;
            float acc = 0.f;
            for (int iy = 0; iy < gh; iy++) {
                const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, gh), H);
                for (int ix = 0; ix < gw; ix++) {
                    const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, gw), W);
                    float val = 0.f;
                    if (ty.valid && tx.valid)
                        val = bilinear_value(ty.h, ty.l, tx.h, tx.l, __ldg(d + ty.low * W + tx.low),
                                             __ldg(d + ty.low * W + tx.high), __ldg(d + ty.high * W + tx.low),
                                             __ldg(d + ty.high * W + tx.high));
                    acc = __fadd_rn(acc, val);
                }
            }
            top[((size_t)row * channels + c) * bins + ph * pooled_w + pw] = __fdiv_rn(acc, g.count);
        }
        return;
    }

    FwdCtx cx;
    cx.tile = tile; cx.rec = rec; cx.obuf = obuf; cx.binrec = binrec;
    cx.channels = channels; cx.pooled_w = pooled_w; cx.bins = bins; cx.obuf_stride = obuf_stride;
    cx.slab0 = blockIdx.y * slabs_per_cta;
    cx.nslab = min(slabs_per_cta, (channels + kSlab - 1) / kSlab - cx.slab0);
    cx.H = H; cx.W = W; cx.gh = gh; cx.gw = gw; cx.nx = nx; cx.count = g.count;
    cx.feat0 = lv.data[sh.level] + ((size_t)g.batch * channels + (size_t)cx.slab0 * kSlab) * H * W;
    cx.out0 = top + ((size_t)row * channels + (size_t)cx.slab0 * kSlab) * bins + ph_begin * pooled_w;
    // fast path: 2x2 sampling grid, 16-byte aligned rows, no output row taller than the tile
    bool fast = gh == 2 && gw == 2 && sh.vec != 0 && (channels % kSlab) == 0;
    for (int bnd = 0; bnd < sh.nbands; bnd++) fast = fast && sh.band_tall[bnd] == 0;
    if (fast) {
        run_bands_fast(cx, sh);
    } else {
        run_bands<false>(cx, sh);
    }
}

// ---------------------------------------------------------------------------------------
// Staged backward path: the mirror image of the forward kernel.  Float atomicAdd on shared memory is a
// CAS loop on sm_100a, so the scatter is organised to need no shared-memory atomics at all:
//   * same per-CTA setup (tap tables, tile extent, bands);
//   * per band a CSR list: for every tile row the (sample row, corner) pairs that touch it with their y
//     weight (built in ascending sample order, so the summation order is deterministic);
//   * per slab: top_diff of the band -> gbuf[c][bin]; then lanes = channels and every warp OWNS whole tile
//     rows: it zeroes its rows and adds (top * (wy * wx)) / count for each entry x sample column x corner
//     with plain read-modify-writes -- every addend is computed exactly like the reference's
//     (roi_align_kernel.cu:186-190,252-255), only the order of the sum differs;
//   * the finished tile is written to the gradient map with lanes along x: one 128-bit vector reduction
//     (red.global.add.v4.f32) per 4 texels instead of 16 scalar atomics per output element.
// RoIs that need a tall band, or exceed the table budgets, use direct atomics inside the same kernel.
// ---------------------------------------------------------------------------------------
struct __align__(8) ListEnt { int bin; float w; };     // bin part (row: pr*PW, column: pw) and axis weight

__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kFwdThreads, 2)
roialign_bwd_staged(const __grid_constant__ LevelTable lv, int channels, int pooled_h, int pooled_w,
                    int sampling_ratio, int rows_per_group, int slabs_per_cta, int gbuf_stride,
                    const float* __restrict__ rois, const int* __restrict__ roi_level,
                    const int* __restrict__ out_index, const float* __restrict__ top_diff) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    unsigned char* dyn = dyn_raw + ((2048u - ((unsigned)__cvta_generic_to_shared(dyn_raw) & 2047u)) & 2047u);
    float* tile = reinterpret_cast<float*>(dyn);                                        // [32][512] swizzled
    float* gbuf = reinterpret_cast<float*>(dyn + kSlab * kTileWords * 4);              // [32][gbuf_stride]
    ListEnt* rlist = reinterpret_cast<ListEnt*>(gbuf + kSlab * gbuf_stride);            // [2*kMaxTaps]
    ListEnt* clist = rlist + 2 * kMaxTaps;                                              // [2*kMaxTaps]
    int* rptr = reinterpret_cast<int*>(clist + 2 * kMaxTaps);                           // [kTileWords + 2]
    int* cptr = rptr + (kTileWords + 2);                                                // [kTileWords + 2]
    __shared__ FwdShared sh;

    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ph_begin = blockIdx.z * rows_per_group;
    const int ph_end = min(pooled_h, ph_begin + rows_per_group);
    const int bins = pooled_h * pooled_w;

    roi_setup(sh, lv, rois, roi_level, n, pooled_h, pooled_w, sampling_ratio, ph_begin, ph_end, kMaxRecords);
    const RoiGeom g = sh.g;
    const int H = sh.H, W = sh.W;
    const int gh = g.grid_h, gw = g.grid_w;
    const int nx = pooled_w * gw;
    const int row = out_index ? out_index[n] : n;
    const int icount = gh * gw;
    const bool pow2 = (icount & (icount - 1)) == 0;
    const float inv_count = 1.0f / g.count;
    const int c_begin = blockIdx.y * slabs_per_cta * kSlab;
    const int c_end = min(channels, c_begin + slabs_per_cta * kSlab);
    float* gbase = lv.data[sh.level] + (size_t)g.batch * channels * H * W;
    const size_t plane = (size_t)H * W;

    bool staged = sh.ok != 0 && sh.tw > 0;
    for (int b = 0; b < sh.nbands; b++) staged = staged && sh.band_tall[b] == 0;
    if (!staged) {
        if (sh.ok != 0 && sh.tw <= 0) return;                       // no valid column: nothing to add
        // direct atomics: thread per (channel, bin) of this row group
        const int group_bins = (ph_end - ph_begin) * pooled_w;
        for (int e = tid; e < (c_end - c_begin) * group_bins; e += kFwdThreads) {
            const int c = c_begin + e / group_bins, b = e % group_bins;
            const int ph = ph_begin + b / pooled_w, pw = b % pooled_w;
            float* d = gbase + (size_t)c * plane;
            const float t = top_diff[((size_t)row * channels + c) * bins + ph * pooled_w + pw];
            for (int iy = 0; iy < gh; iy++) {
                const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, gh), H);
                for (int ix = 0; ix < gw; ix++) {
                    const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, gw), W);
                    if (!(ty.valid && tx.valid)) continue;
                    atomicAdd(d + ty.low * W + tx.low, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.h)), g.count));
                    atomicAdd(d + ty.low * W + tx.high, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.l)), g.count));
                    atomicAdd(d + ty.high * W + tx.low, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.h)), g.count));
                    atomicAdd(d + ty.high * W + tx.high, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.l)), g.count));
                }
            }
        }
        return;
    }

    const int x_lo = sh.x_lo, tw = sh.tw;
    const bool vec = sh.vec != 0;
    // ---- per sample column: output column (bin) it belongs to ----
    for (int sx = tid; sx < nx; sx += kFwdThreads) cptr[sx] = sx / gw;
    int p0 = 0;
    for (int band = 0; band < sh.nbands; band++) {
        const int p1 = sh.band_p1[band], y_lo = sh.band_ylo[band], rows = sh.band_rows[band];
        const int band_bins = (p1 - p0) * pooled_w;
        const int ny_b = (p1 - p0) * gh;
        __syncthreads();                                            // previous band done with rlist / rptr
        // ---- row list of the band ----
        if (tid == 0) rptr[0] = 0;
        for (int r = tid; r < rows; r += kFwdThreads) {
            int cnt = 0;
            for (int sy = 0; sy < ny_b; sy++) {
                const Tap t = sh.ytab[p0 * gh + sy];
                if (t.low >= 0) cnt += (t.low - y_lo == r) + (t.high - y_lo == r);
            }
            rptr[r + 1] = cnt;
        }
        __syncthreads();
        if (tid == 0) for (int r = 0; r < rows; r++) rptr[r + 1] += rptr[r];
        __syncthreads();
        for (int r = tid; r < rows; r += kFwdThreads) {
            int pos = rptr[r];
            for (int sy = 0; sy < ny_b; sy++) {
                const Tap t = sh.ytab[p0 * gh + sy];
                if (t.low < 0) continue;
                if (t.low - y_lo == r) rlist[pos++] = ListEnt{(sy / gh) * pooled_w, t.h};
                if (t.high - y_lo == r) rlist[pos++] = ListEnt{(sy / gh) * pooled_w, t.l};
            }
        }
        // ---- what this thread flushes for every slab (mirror of the forward staging) ----
        const int tws = vec ? (tw >> 2) : tw;
        const int nunit = rows * tws;
        const float inv = 1.0f / (float)max(tws, 1);

        for (int c0 = c_begin; c0 < c_end; c0 += kSlab) {
            const int nch = min(kSlab, channels - c0);
            __syncthreads();                                        // lists ready / previous slab flushed
            // ---- top_diff of the band for this slab -> gbuf[c][bin] ----
            const float* tsrc = top_diff + ((size_t)row * channels + c0) * bins + (ph_begin + p0) * pooled_w;
            for (int c = warp; c < nch; c += kFwdWarps)
                for (int bq = lane; bq < band_bins; bq += 32)
                    gbuf[c * gbuf_stride + bq] = __ldg(tsrc + (size_t)c * bins + bq);
            __syncthreads();
            // ---- scatter in tile space: lanes = channels, each warp OWNS whole tile rows, so the
            //      read-modify-writes below never race and need no atomics ----
            const float* gl = gbuf + lane * gbuf_stride;
            float* trow = tile + lane * kTileWords;
            for (int r = warp; r < rows; r += kFwdWarps) {
                for (int x = 0; x < tw; x++) trow[(r * tw + x) ^ lane] = 0.f;
                const int r0 = rptr[r], r1 = rptr[r + 1];
                for (int a = r0; a < r1; a++) {
                    const ListEnt ey = rlist[a];               // (sample row, corner): bin row + y weight
                    for (int sx = 0; sx < nx; sx++) {
                        const Tap tx = sh.xtab[sx];
                        if (tx.low < 0) continue;              // uniform
                        const float tv = gl[ey.bin + cptr[sx]];
                        const float g_lo = __fmul_rn(tv, __fmul_rn(ey.w, tx.h));
                        const float g_hi = __fmul_rn(tv, __fmul_rn(ey.w, tx.l));
                        const int i_lo = (r * tw + tx.low - x_lo) ^ lane, i_hi = (r * tw + tx.high - x_lo) ^ lane;
                        trow[i_lo] = __fadd_rn(trow[i_lo], pow2 ? __fmul_rn(g_lo, inv_count) : __fdiv_rn(g_lo, g.count));
                        trow[i_hi] = __fadd_rn(trow[i_hi], pow2 ? __fmul_rn(g_hi, inv_count) : __fdiv_rn(g_hi, g.count));
                    }
                }
            }
            __syncthreads();
            // ---- flush: lanes along x, one vector reduction per 4 texels ----
            float* gdst = gbase + (size_t)c0 * plane;
            for (int i = lane; i < nunit; i += 32) {
                const int ry = (int)(((float)i + 0.5f) * inv);
                const int rx = vec ? ((i - ry * tws) << 2) : (i - ry * tws);
                const size_t goff = (size_t)(y_lo + ry) * W + (x_lo + rx);
                const int t = ry * tw + rx;
                for (int c = warp; c < nch; c += kFwdWarps) {
                    if (vec) {
                        const float4 s4 = *reinterpret_cast<const float4*>(tile + c * kTileWords + (t ^ (c & ~3)));
                        const float4 v = swizzle4(s4, c & 3);          // the permutation is an involution
                        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
                            red_add_v4(gdst + (size_t)c * plane + goff, v);
                    } else {
                        const float v = tile[c * kTileWords + (t ^ c)];
                        if (v != 0.f) atomicAdd(gdst + (size_t)c * plane + goff, v);
                    }
                }
            }
        }
        p0 = p1;
    }
}

// ---------------------------------------------------------------------------------------
// Record-based backward (default).  The generic kernel spends its time on arithmetic, not on the
// reductions (ncu: 66 % issue-active, 9 % lg_throttle; 4 IEEE divisions for the coordinates and 4 for the
// addends of every sample of every channel).  Here one CTA = one RoI x one group of output rows x a
// chunk of channels: the per-axis tap tables and one record per sample -- the 4 plane offsets y*W+x and
// the 4 corner weights, pre-divided by `count` when it is a power of two (x/2^k == x*2^-k exactly, and
// scaling by 2^-k commutes with the rounding of top*w) -- are built once in shared memory; then a thread
// per (channel, bin) element reads its top_diff (coalesced) and issues 16 reductions whose addends are
// bit-identical to the reference's (top*w)/count (roi_align_kernel.cu:252-265).
// ---------------------------------------------------------------------------------------
struct __align__(16) BwdRec { int o1, o2, o3, o4; float w1, w2, w3, w4; };     // o1 < 0: sample outside the map

__global__ void __launch_bounds__(256)
roialign_bwd_records(const __grid_constant__ LevelTable lv, int channels, int pooled_h, int pooled_w,
                     int sampling_ratio, int rows_per_group, int ch_per_cta,
                     const float* __restrict__ rois, const int* __restrict__ roi_level,
                     const int* __restrict__ out_index, const float* __restrict__ top_diff) {
    extern __shared__ __align__(16) unsigned char dynb[];
    BwdRec* rec = reinterpret_cast<BwdRec*>(dynb);                 // [<= kMaxRecords]
    __shared__ FwdShared sh;

    const int n = blockIdx.x;
    const int tid = threadIdx.x;
    const int ph_begin = blockIdx.z * rows_per_group;
    const int ph_end = min(pooled_h, ph_begin + rows_per_group);
    const int bins = pooled_h * pooled_w;
    roi_setup(sh, lv, rois, roi_level, n, pooled_h, pooled_w, sampling_ratio, ph_begin, ph_end, kMaxRecords);
    const RoiGeom g = sh.g;
    const int H = sh.H, W = sh.W;
    const int gh = g.grid_h, gw = g.grid_w;
    const int nx = pooled_w * gw, ny = (ph_end - ph_begin) * gh;
    const int row = out_index ? out_index[n] : n;
    const int c_begin = blockIdx.y * ch_per_cta, c_end = min(channels, c_begin + ch_per_cta);
    const int group_bins = (ph_end - ph_begin) * pooled_w;
    const size_t plane = (size_t)H * W;
    float* gbase = lv.data[sh.level] + ((size_t)g.batch * channels + c_begin) * plane;
    const float* tsrc = top_diff + ((size_t)row * channels + c_begin) * bins + ph_begin * pooled_w;
    const bool fits = ny <= kMaxTaps && nx <= kMaxTaps && ny * nx <= kMaxRecords;     // as in roi_setup

    if (!fits) {
        // adaptive grids beyond the table budget: per-element geometry (same arithmetic)
        for (int e = tid; e < (c_end - c_begin) * group_bins; e += 256) {
            const int c = e / group_bins, b = e - c * group_bins;
            const int ph = ph_begin + b / pooled_w, pw = b % pooled_w;
            float* d = gbase + (size_t)c * plane;
            const float t = tsrc[(size_t)c * bins + b];
            for (int iy = 0; iy < gh; iy++) {
                const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, gh), H);
                for (int ix = 0; ix < gw; ix++) {
                    const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, gw), W);
                    if (!(ty.valid && tx.valid)) continue;
                    atomicAdd(d + ty.low * W + tx.low, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.h)), g.count));
                    atomicAdd(d + ty.low * W + tx.high, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.h, tx.l)), g.count));
                    atomicAdd(d + ty.high * W + tx.low, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.h)), g.count));
                    atomicAdd(d + ty.high * W + tx.high, __fdiv_rn(__fmul_rn(t, __fmul_rn(ty.l, tx.l)), g.count));
                }
            }
        }
        return;
    }

    const int icount = gh * gw;
    const bool pow2 = (icount & (icount - 1)) == 0;
    const float scale = pow2 ? 1.0f / g.count : 1.0f;
    for (int s2 = tid; s2 < ny * nx; s2 += 256) {
        const int sy = s2 / nx, sx = s2 - sy * nx;
        const Tap ty = sh.ytab[sy], tx = sh.xtab[sx];
        BwdRec r;
        if (ty.low >= 0 && tx.low >= 0) {
            r.o1 = ty.low * W + tx.low;  r.o2 = ty.low * W + tx.high;
            r.o3 = ty.high * W + tx.low; r.o4 = ty.high * W + tx.high;
            r.w1 = __fmul_rn(__fmul_rn(ty.h, tx.h), scale); r.w2 = __fmul_rn(__fmul_rn(ty.h, tx.l), scale);
            r.w3 = __fmul_rn(__fmul_rn(ty.l, tx.h), scale); r.w4 = __fmul_rn(__fmul_rn(ty.l, tx.l), scale);
        } else {
            r.o1 = -1; r.o2 = r.o3 = r.o4 = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
        }
        rec[s2] = r;
    }
    __syncthreads();

    const int total = (c_end - c_begin) * group_bins;
    const float inv_gb = 1.0f / (float)group_bins;
    for (int e = tid; e < total; e += 256) {
        const int c = (int)(((float)e + 0.5f) * inv_gb);
        const int b = e - c * group_bins;
        const int pr = b / pooled_w, pw = b - pr * pooled_w;
        const float t = __ldg(tsrc + (size_t)c * bins + b);
        float* d = gbase + (size_t)c * plane;
        const BwdRec* rr = rec + (pr * gh) * nx + pw * gw;
        for (int iy = 0; iy < gh; iy++) {
            for (int ix = 0; ix < gw; ix++) {
                const BwdRec r = rr[iy * nx + ix];
                if (r.o1 < 0) continue;
                if (pow2) {
                    atomicAdd(d + r.o1, __fmul_rn(t, r.w1));
                    atomicAdd(d + r.o2, __fmul_rn(t, r.w2));
                    atomicAdd(d + r.o3, __fmul_rn(t, r.w3));
                    atomicAdd(d + r.o4, __fmul_rn(t, r.w4));
                } else {
                    atomicAdd(d + r.o1, __fdiv_rn(__fmul_rn(t, r.w1), g.count));
                    atomicAdd(d + r.o2, __fdiv_rn(__fmul_rn(t, r.w2), g.count));
                    atomicAdd(d + r.o3, __fdiv_rn(__fmul_rn(t, r.w3), g.count));
                    atomicAdd(d + r.o4, __fdiv_rn(__fmul_rn(t, r.w4), g.count));
                }
            }
        }
    }
}

}  // namespace vosd

#include "roialign_sep.cuh"
#include "roialign_nhwc.cuh"
#include "roialign_rw.cuh"
#include "roialign_nhwc_bwd.cuh"
namespace vosd {

// test hook (vosd_debug_force_generic): 0 = default (separable forward where it applies, record-based backward), 1 = generic kernels
// everywhere, 2 = staged kernels everywhere (the staged backward is parity-tested but not yet faster than
// the atomic scatter -- profiles/r01_roialign_bwd_staged_*_ncu.txt -- so it is not the default)
static int g_force_generic = 0;

// L2 promotion of the row-window kernel's tensor maps: a box row is 48 - 112 bytes per channel plane
#ifndef VOSD_RW_L2PROMO
#define VOSD_RW_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif

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
    if (g_force_generic == 0 && sr == 2 && (pw == 7 || pw == 14 || pw == 28) && channels % kSlab == 0) {
        // separable row-streaming kernel (roialign_sep.cuh): T warps per (RoI, slab) team, NPH output rows per CTA
        const int T = pw / 7, nph = 7, teams = kSepWarps / T;
        const int slabs_all = channels / kSlab, rgroups = ceil_div(ph, nph);
        if (rgroups <= 65535) {
            // slabs per CTA: a multiple of the teams, few enough that the grid has >= ~12 CTAs per SM
#ifndef VOSD_SEP_CTAS_PER_SM
#define VOSD_SEP_CTAS_PER_SM 12
#endif
            long long spc = (long long)slabs_all * num_rois * rgroups / ((long long)VOSD_SEP_CTAS_PER_SM * kNumSMs);
            spc = spc / teams * teams;
            if (spc < teams) spc = teams;
            if (spc > slabs_all) spc = slabs_all;
            dim3 grid(num_rois, ceil_div(slabs_all, (int)spc), rgroups);
            cudaError_t e;
            if (T == 1) {
                e = cudaFuncSetAttribute(roialign_fwd_sep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_dyn_bytes<1>());
                if (e == cudaSuccess)
                    roialign_fwd_sep<1><<<grid, kSepThreads, sep_dyn_bytes<1>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top);
            } else if (T == 2) {
                e = cudaFuncSetAttribute(roialign_fwd_sep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_dyn_bytes<2>());
                if (e == cudaSuccess)
                    roialign_fwd_sep<2><<<grid, kSepThreads, sep_dyn_bytes<2>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top);
            } else {
                e = cudaFuncSetAttribute(roialign_fwd_sep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_dyn_bytes<4>());
                if (e == cudaSuccess)
                    roialign_fwd_sep<4><<<grid, kSepThreads, sep_dyn_bytes<4>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top);
            }
            if (e != cudaSuccess) return VOSD_ERR_LAUNCH;
            count_launch();
            return check_launch();
        }
    }
    // staged path: row groups of <= ~112 bins, sample records <= kMaxRecords, >= ~8 waves of CTAs
    const int slabs = ceil_div(channels, kSlab);
    int groups = ceil_div(ph * pw, 112);
    if (groups > ph) groups = ph;
    int rpg = ceil_div(ph, groups);
    const int gs = sr > 0 ? sr : 1;
    while (rpg > 1 && ((long long)rpg * gs * pw * gs > kMaxRecords || rpg * gs > kMaxTaps)) rpg--;
    groups = ceil_div(ph, rpg);
    const bool staged_ok = g_force_generic != 1 && pw * gs <= kMaxTaps && rpg * gs <= kMaxTaps &&
                           (long long)rpg * gs * pw * gs <= kMaxRecords && groups <= 65535;
    if (staged_ok) {
        const int rec_cap = sr > 0 ? rpg * sr * pw * sr : kMaxRecords;
        const int obuf_stride = (rpg * pw) | 1;
        long long spc = (long long)slabs * num_rois * groups / (6LL * 2 * kNumSMs);
        if (spc < 1) spc = 1;
        if (spc > slabs) spc = slabs;
        const size_t dyn = (size_t)kSlab * kTileWords * 4 + (size_t)rec_cap * sizeof(SampleRec) +
                           (size_t)kSlab * obuf_stride * sizeof(float) + (size_t)rpg * pw * sizeof(int) + 2048;
        if (cudaFuncSetAttribute(roialign_fwd_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
            return VOSD_ERR_LAUNCH;
        dim3 grid(num_rois, ceil_div(slabs, (int)spc), groups);
        roialign_fwd_staged<<<grid, kFwdThreads, dyn, stream>>>(t, channels, ph, pw, sr, rpg, (int)spc, rec_cap,
                                                                obuf_stride, rois, roi_level, out_index, top);
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
    const int slabs = ceil_div(channels, kSlab);
    int groups = ceil_div(ph * pw, 112);
    if (groups > ph) groups = ph;
    int rpg = ceil_div(ph, groups);
    const int gs = sr > 0 ? sr : 1;
    while (rpg > 1 && rpg * gs > kMaxTaps) rpg--;
    groups = ceil_div(ph, rpg);
    if (g_force_generic == 0 && sr == 2 && (pw == 7 || pw == 14 || pw == 28) && channels % kSlab == 0) {
        // separable row-streaming scatter (roialign_sep.cuh): one reduction per (channel, texel) of the footprint
        const int T = pw / 7, teams = kSepWarps / T;
        const int slabs_all = channels / kSlab, rgroups = ceil_div(ph, 7);
        if (rgroups <= 65535) {
#ifndef VOSD_SEP_CTAS_PER_SM
#define VOSD_SEP_CTAS_PER_SM 12
#endif
            long long spc = (long long)slabs_all * num_rois * rgroups / ((long long)VOSD_SEP_CTAS_PER_SM * kNumSMs);
            spc = spc / teams * teams;
            if (spc < teams) spc = teams;
            if (spc > slabs_all) spc = slabs_all;
            dim3 grid(num_rois, ceil_div(slabs_all, (int)spc), rgroups);
            cudaError_t e;
            if (T == 1) {
                e = cudaFuncSetAttribute(roialign_bwd_sep<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_bwd_dyn_bytes<1>());
                if (e == cudaSuccess)
                    roialign_bwd_sep<1><<<grid, kSepThreads, sep_bwd_dyn_bytes<1>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top_diff);
            } else if (T == 2) {
                e = cudaFuncSetAttribute(roialign_bwd_sep<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_bwd_dyn_bytes<2>());
                if (e == cudaSuccess)
                    roialign_bwd_sep<2><<<grid, kSepThreads, sep_bwd_dyn_bytes<2>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top_diff);
            } else {
                e = cudaFuncSetAttribute(roialign_bwd_sep<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_bwd_dyn_bytes<4>());
                if (e == cudaSuccess)
                    roialign_bwd_sep<4><<<grid, kSepThreads, sep_bwd_dyn_bytes<4>(), stream>>>(t, channels, ph, (int)spc, rois, roi_level, out_index, top_diff);
            }
            if (e != cudaSuccess) return VOSD_ERR_LAUNCH;
            count_launch();
            return check_launch();
        }
    }
    if (g_force_generic == 0) {
        // record-based scatter; rows grouped so that one group's records fit (<= kMaxRecords)
        int rg = ph;
        while (rg > 1 && ((long long)rg * gs * pw * gs > kMaxRecords || rg * gs > kMaxTaps)) rg--;
        const int ngroups = ceil_div(ph, rg);
        long long cpc = (long long)channels * num_rois * ngroups / (3LL * 6 * kNumSMs);     // channels per CTA: ~3 waves of 6 CTAs/SM
        if (cpc < 8) cpc = 8;
        if (cpc > channels) cpc = channels;
        const size_t dyn = (size_t)kMaxRecords * sizeof(BwdRec);
        if (ngroups <= 65535) {
            dim3 grid(num_rois, ceil_div(channels, (int)cpc), ngroups);
            roialign_bwd_records<<<grid, 256, dyn, stream>>>(t, channels, ph, pw, sr, rg, (int)cpc, rois, roi_level,
                                                             out_index, top_diff);
            count_launch();
            return check_launch();
        }
    }
    const bool staged_ok = g_force_generic == 2 && pw * gs <= kMaxTaps && rpg * gs <= kMaxTaps && groups <= 65535;
    if (staged_ok) {
        const int gbuf_stride = (rpg * pw) | 1;
        long long spc = (long long)slabs * num_rois * groups / (6LL * 2 * kNumSMs);
        if (spc < 1) spc = 1;
        if (spc > slabs) spc = slabs;
        const size_t dyn = (size_t)kSlab * kTileWords * 4 + (size_t)kSlab * gbuf_stride * sizeof(float) +
                           4 * kMaxTaps * sizeof(ListEnt) + (size_t)(2 * kTileWords + 4) * sizeof(int) + 2048;
        if (cudaFuncSetAttribute(roialign_bwd_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
            return VOSD_ERR_LAUNCH;
        dim3 grid(num_rois, ceil_div(slabs, (int)spc), groups);
        roialign_bwd_staged<<<grid, kFwdThreads, dyn, stream>>>(t, channels, ph, pw, sr, rpg, (int)spc, gbuf_stride,
                                                                rois, roi_level, out_index, top_diff);
    } else {
        const long long total = (long long)num_rois * channels * ph * pw;
        roialign_bwd_generic<<<grid_for(total, 256), 256, 0, stream>>>(
            t, channels, ph, pw, sr, total, rois, roi_level, out_index, top_diff);
    }
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

// ---------------------------------------------------------------------------------------
// Channels-last forward (roialign_nhwc.cuh): tensor maps are encoded per call on the host stack.
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

template <int T>
static int launch_nhwc(const NhwcMaps& maps, const LevelTable& t, int channels, int ph, int num_rois,
                       const float* rois, const int* roi_level, const int* out_index, float* top, cudaStream_t stream) {
    const size_t dyn = (size_t)kNhwcWarps * kNhwcRingBytes;
    if (cudaFuncSetAttribute(roialign_fwd_nhwc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess)
        return VOSD_ERR_LAUNCH;
    dim3 grid(num_rois, ceil_div(channels / kSlab, kNhwcWarps / T), ceil_div(ph, 7));
    roialign_fwd_nhwc<T><<<grid, kNhwcThreads, dyn, stream>>>(maps, t, channels, ph, rois, roi_level, out_index, top);
    count_launch();
    return check_launch();
}

extern "C" int vosd_roialign_ml_fwd_nhwc(const float* const* level_data, const int* level_h, const int* level_w,
                                         const float* level_scale, int num_levels, int batch_size, int channels,
                                         int aligned_height, int aligned_width, int sampling_ratio,
                                         int num_rois, const float* rois, const int* roi_level,
                                         const int* out_index, float* top_data, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, level_data, level_h, level_w, level_scale, num_levels);
    if (rc) return rc;
    if (num_levels > 1 && !roi_level && num_rois > 0) return VOSD_ERR_BAD_ARG;
    if (channels <= 0 || aligned_height <= 0 || aligned_width <= 0 || num_rois < 0 || batch_size <= 0) return VOSD_ERR_BAD_SHAPE;
    if (sampling_ratio != 2 || (aligned_width != 7 && aligned_width != 14 && aligned_width != 28) || channels % kSlab ||
        ceil_div(aligned_height, 7) > 65535 || ceil_div(channels / kSlab, 2) > 65535)
        return VOSD_ERR_UNSUPPORTED;            // the caller falls back to the NCHW entry point
    if (num_rois == 0) return VOSD_OK;
    if (!rois || !top_data) return VOSD_ERR_BAD_ARG;
    const EncodeTiledFn enc = encode_tiled();
    if (!enc) return VOSD_ERR_UNSUPPORTED;
    if (num_levels > 4) return VOSD_ERR_UNSUPPORTED;    // NhwcMaps holds the 4 RoI levels of an FPN
    NhwcMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < num_levels; l++) {
        if (!aligned16(level_data[l])) return VOSD_ERR_BAD_ARG;
        const cuuint64_t C = (cuuint64_t)channels, W = (cuuint64_t)level_w[l], H = (cuuint64_t)level_h[l];
        const cuuint64_t dims[4] = {C, W, H, (cuuint64_t)batch_size};
        const cuuint64_t strides[3] = {C * 4, W * C * 4, H * W * C * 4};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        for (int b = 0; b < kNhwcBoxes; b++) {
            const cuuint32_t box[4] = {(cuuint32_t)kSlab, (cuuint32_t)(kNhwcBoxStep * (b + 1)), 1, 1};
            if (enc(&maps.m[l][b], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(level_data[l]), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return VOSD_ERR_BAD_SHAPE;
        }
    }
    switch (aligned_width / 7) {
        case 1: return launch_nhwc<1>(maps, t, channels, aligned_height, num_rois, rois, roi_level, out_index, top_data, stream);
        case 2: return launch_nhwc<2>(maps, t, channels, aligned_height, num_rois, rois, roi_level, out_index, top_data, stream);
        default: return launch_nhwc<4>(maps, t, channels, aligned_height, num_rois, rois, roi_level, out_index, top_data, stream);
    }
}

// ---------------------------------------------------------------------------------------
// Row-window forward (roialign_rw.cuh): plan kernel + persistent TMA-fed main kernel, caller-provided workspace.
// ---------------------------------------------------------------------------------------
static bool rw_applies(int channels, int ph, int pw, int sr) {
    return g_force_generic == 0 && sr == 2 && (pw == 7 || pw == 14 || pw == 28) && ph > 0 && channels % kSlab == 0;
}

static long long rw_base_items(int ph, int pw, int num_rois) {
    return (long long)num_rois * ((ph + 6) / 7) * (pw / 7);
}

static size_t rw_items_bytes(long long nbase) { return align_up((size_t)(2 * nbase + 4096) * sizeof(RwItem), 256); }

extern "C" size_t vosd_roialign_fwd_workspace_bytes(const int* level_h, const int* level_w, int num_levels, int batch_size,
                                                    int channels, int aligned_height, int aligned_width, int num_rois) {
    if (!level_h || !level_w || num_levels < 1 || num_levels > 4 || batch_size <= 0 || channels <= 0 || num_rois < 0 ||
        aligned_height <= 0 || aligned_width <= 0)
        return 0;
    size_t bytes = 256 + rw_items_bytes(rw_base_items(aligned_height, aligned_width, num_rois));
    for (int l = 0; l < num_levels; l++)
        if (level_w[l] % 4)
            bytes += align_up((size_t)batch_size * channels * level_h[l] * ((level_w[l] + 3) & ~3) * sizeof(float), 256);
    return bytes;
}

extern "C" int vosd_roialign_ml_fwd_ws(const float* const* level_data, const int* level_h, const int* level_w,
                                       const float* level_scale, int num_levels, int batch_size, int channels,
                                       int aligned_height, int aligned_width, int sampling_ratio,
                                       int num_rois, const float* rois, const int* roi_level,
                                       const int* out_index, float* top_data,
                                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, level_data, level_h, level_w, level_scale, num_levels);
    if (rc) return rc;
    if (num_levels > 1 && !roi_level && num_rois > 0) return VOSD_ERR_BAD_ARG;
    if (channels <= 0 || aligned_height <= 0 || aligned_width <= 0 || num_rois < 0 || batch_size <= 0) return VOSD_ERR_BAD_SHAPE;
    if (num_rois == 0) return VOSD_OK;
    if (!rois || !top_data) return VOSD_ERR_BAD_ARG;
    const EncodeTiledFn enc = encode_tiled();
    bool use_rw = rw_applies(channels, aligned_height, aligned_width, sampling_ratio) && enc && num_levels <= 4 &&
                  rw_base_items(aligned_height, aligned_width, num_rois) < (1ll << 27);
    for (int l = 0; l < num_levels && use_rw; l++) use_rw = aligned16(level_data[l]);
    if (!use_rw)
        return ml_fwd(t, channels, aligned_height, aligned_width, sampling_ratio, num_rois, rois, roi_level, out_index,
                      top_data, stream);
    const size_t need = vosd_roialign_fwd_workspace_bytes(level_h, level_w, num_levels, batch_size, channels,
                                                          aligned_height, aligned_width, num_rois);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255)) return VOSD_ERR_WORKSPACE;
    const int nbase = (int)rw_base_items(aligned_height, aligned_width, num_rois);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    RwCounters* ctr = reinterpret_cast<RwCounters*>(ws);
    RwItem* items = reinterpret_cast<RwItem*>(ws + 256);
    unsigned char* pad_ws = ws + 256 + rw_items_bytes(nbase);
    const int cap_extra = nbase + 4096;
    // tensor maps over (W, H, N * C), box (BX, 1, 32); levels whose rows are not 16-byte multiples read a padded copy
    RwMaps maps;
    memset(&maps, 0, sizeof(maps));
    RwPadJobs pj;
    memset(&pj, 0, sizeof(pj));
    long long pad_f4 = 0;
    for (int l = 0; l < num_levels; l++) {
        const int W = level_w[l], H = level_h[l], Wp = (W + 3) & ~3;
        const float* src = level_data[l];
        if (Wp != W) {
            float* dst = reinterpret_cast<float*>(pad_ws);
            const long long rows = (long long)batch_size * channels * H;
            pj.src[pj.n] = src; pj.dst[pj.n] = dst; pj.W[pj.n] = W; pj.Wp[pj.n] = Wp; pj.rows[pj.n] = rows;
            pj.n++;
            pad_f4 += rows * (Wp / 4);
            pad_ws += align_up((size_t)rows * Wp * sizeof(float), 256);
            src = dst;
        }
        const cuuint64_t dims[3] = {(cuuint64_t)Wp, (cuuint64_t)H, (cuuint64_t)batch_size * channels};
        const cuuint64_t strides[2] = {(cuuint64_t)Wp * 4, (cuuint64_t)H * Wp * 4};
        const cuuint32_t es[3] = {1, 1, 1};
        for (int v = 0; v < kRwVariants; v++) {
            const cuuint32_t box[3] = {(cuuint32_t)(12 + 8 * v), 1, (cuuint32_t)kSlab};
            if (enc(&maps.m[l][v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(src), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, VOSD_RW_L2PROMO,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return VOSD_ERR_BAD_SHAPE;
        }
    }
    if (cudaMemsetAsync(ctr, 0, sizeof(RwCounters), stream) != cudaSuccess) return VOSD_ERR_LAUNCH;
    pj.plan_blocks = ceil_div(nbase, 8);
    // padding copies ride in the same launch: ~8 float4 per thread, at most two blocks per SM beside the plan
    long long pad_blocks = (pad_f4 + 128 * 8 - 1) / (128 * 8);
    if (pad_blocks > 2 * kNumSMs) pad_blocks = 2 * kNumSMs;
    roialign_rw_plan<<<pj.plan_blocks + (int)pad_blocks, 128, 0, stream>>>(t, pj, channels, aligned_height, aligned_width, num_rois,
                                                                          rois, roi_level, out_index, items, nbase, cap_extra, ctr);
    count_launch();
    if (check_launch() != VOSD_OK) return VOSD_ERR_LAUNCH;
    RwArgs a;
    a.items = items; a.ctr = ctr; a.top = top_data; a.rois = rois;
    a.nbase = nbase; a.cap_extra = cap_extra;
    a.channels = channels; a.pooled_h = aligned_height; a.pooled_w = aligned_width;
    // work unit = (item, slab range): >= ~8 units per warp, >= 2 slabs per unit
    const int slabs_all = channels / kSlab;
    int split_log2 = 0;
#ifndef VOSD_RW_UNITS_PER_WARP
#define VOSD_RW_UNITS_PER_WARP 8
#endif
    while ((1 << (split_log2 + 1)) * 2 <= slabs_all && ((long long)nbase << split_log2) < (long long)VOSD_RW_UNITS_PER_WARP * kNumSMs * kRwWarps)
        split_log2++;
    a.split_log2 = split_log2;
    a.slabs_per_unit = ceil_div(slabs_all, 1 << split_log2);
    a.top_aligned = aligned16(top_data) ? 1 : 0;
    const size_t dyn = (size_t)kRwWarps * kRwWarpBytes;
    if (cudaFuncSetAttribute(roialign_fwd_rw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn) != cudaSuccess) {
        check_launch();
        return VOSD_ERR_LAUNCH;
    }
    const long long units = (long long)nbase << split_log2;
    const int grid = (int)(units < (long long)kNumSMs * kRwWarps ? ceil_div((int)units, kRwWarps) : kNumSMs);
    roialign_fwd_rw<<<grid, kRwThreads, dyn, stream>>>(maps, t, a);
    count_launch();
    return check_launch();
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

// Channels-last gradient maps (roialign_nhwc_bwd.cuh): level_diff[l] is (N, H_l, W_l, C) in memory.
extern "C" int vosd_roialign_ml_bwd_nhwc(const float* top_diff, float* const* level_diff, const int* level_h,
                                         const int* level_w, const float* level_scale, int num_levels,
                                         int batch_size, int channels,
                                         int aligned_height, int aligned_width, int sampling_ratio,
                                         int num_rois, const float* rois, const int* roi_level,
                                         const int* out_index, int zero_init, cudaStream_t stream) {
    LevelTable t;
    int rc = fill_table(t, (const float* const*)level_diff, level_h, level_w, level_scale, num_levels);
    if (rc) return rc;
    if (channels <= 0 || aligned_height <= 0 || aligned_width <= 0 || num_rois < 0 || batch_size <= 0) return VOSD_ERR_BAD_SHAPE;
    if (sampling_ratio != 2 || aligned_height % 7 || aligned_width % 7 || channels % 32) return VOSD_ERR_UNSUPPORTED;
    const int bh = aligned_height / 7, bw = aligned_width / 7;
    if (bh * bw > 65535) return VOSD_ERR_UNSUPPORTED;
    if (num_levels > 1 && !roi_level && num_rois > 0) return VOSD_ERR_BAD_ARG;
    if (zero_init) {
        for (int l = 0; l < num_levels; l++)
            if (cudaMemsetAsync(t.data[l], 0, sizeof(float) * (size_t)batch_size * channels * t.h[l] * t.w[l], stream) != cudaSuccess)
                return VOSD_ERR_LAUNCH;
    }
    if (num_rois == 0) return VOSD_OK;
    if (!rois || !top_diff) return VOSD_ERR_BAD_ARG;
    dim3 grid(num_rois, bh * bw);
    roialign_bwd_nhwc<<<grid, 32 * kNbWarps, 0, stream>>>(t, channels, aligned_height, aligned_width, bw, rois, roi_level,
                                                          out_index, top_diff);
    count_launch();
    return check_launch();
}

// Test hook: route RoIAlign through the generic (un-staged) kernels so both paths stay covered.
extern "C" VOSD_API int vosd_debug_force_generic(int on) {
    const int old = g_force_generic;
    if (!test_hooks_enabled()) return old;          // production processes cannot change the kernel family
    g_force_generic = on;
    return old;
}
