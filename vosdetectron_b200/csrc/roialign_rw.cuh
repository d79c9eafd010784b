// "Row-window" RoIAlign forward for NCHW maps, fed by tensor-mode TMA (sampling_ratio == 2; reference:
// lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:65-121).  Included by roialign.cu.  Default forward of
// vosd_roialign_ml_fwd_ws (the entry point with a caller-provided workspace); roialign_sep.cuh stays behind the
// workspace-free entry points.
//
// Why another kernel (profiles/r01_roialign_fwd_v3g_sep_ncu.txt): the separable kernel spends
// ~45 % of its 3900 warp-instructions per (RoI, 32-channel slab) in producer warps that transpose NCHW rows into
// slot[x][c], ~20 % in per-CTA geometry set-up, and its LSU pipe carries ~1800 wavefronts per (RoI, slab).  Here:
//   * ONE small PLAN kernel derives every RoI's sample taps once (thread per 7x7 output block = "item"), splits
//     blocks that are too wide / too dense into sub-blocks, and writes 464-byte item records to the workspace.
//   * The main kernel is persistent (one CTA per SM, kRwWarps independent warps, work units from an atomic queue).
//     A warp owns an item: lane 0 keeps the warp's ring of row slots full with one cp.async.bulk.tensor per texel
//     row.  The box (x = BX, y = 1, c = 32) of the NCHW map lands as slot[c][x], pitch BX in {12, 20, 28} floats:
//     BX / 4 is odd, so lanes = channels read their own row with 128-bit loads WITHOUT bank conflicts
//     (16-byte bank group of lane c, quad j: (c * BX / 4 + j) mod 8, distinct over 8 consecutive lanes).
//     No producer warps, no transposing stores, no CTA barrier.
//   * Vertical pass first, elementwise in x (static register index): V[s][x] += Wy[row][s] * F[row][x] for the
//     <= S output rows that are "open" at this texel row (slot s = output row mod S; a zero weight skips the FMAs
//     through a warp-uniform branch).  When an output row's last texel row has been added, its V[s][0..BX) goes
//     through a per-lane scratch (the ring slot just consumed: dynamic x index) for the horizontal taps:
//     R[pw] = sum of 4 taps wx * V[x]  ->  obuf[c][ph][pw]  ->  one bulk store per slab (7x7 head).
//   Per (RoI, slab) with the bench's 12x12-texel footprints: ~1100 instructions and ~480 LSU wavefronts.
// Non-finite features propagate exactly as in the reference: a texel enters a bin iff it is one of the 4 taps of a
// valid sample of that bin (zero weights included: a tap with weight 0 is stored as -0.0f, which still multiplies;
// tests/test_gpu_ext_shims.py::test_non_finite_texels_follow_the_reference).
#pragma once
#include <cuda.h>
#include <type_traits>

namespace vosd {

// Warps per CTA (one CTA per SM) and shared memory per warp.  Registers are allocated per SM sub-partition
// (16384 each, warps dealt round-robin), so the budget per thread is 16384 / (32 * ceil(warps / 4)): 12 warps -> 168.
#ifndef VOSD_RW_WARPS
#define VOSD_RW_WARPS 12
#endif
#ifndef VOSD_RW_WARP_BYTES
#define VOSD_RW_WARP_BYTES 18816
#endif
constexpr int kRwWarps = VOSD_RW_WARPS;
constexpr int kRwThreads = 32 * kRwWarps;
constexpr int kRwWarpBytes = VOSD_RW_WARP_BYTES;       // per warp: obuf + row-weight table + barriers + ring
constexpr int kRwMaxRows = 64;                         // texel rows of an item's footprint (row-weight table)
constexpr int kRwMaxSlots = 8;
constexpr int kRwObufBytes = kSlab * 49 * 4;           // 6272
constexpr int kRwWrowBytes = kRwMaxRows * 4 * 4;       // 1024: per row <= 3 weights + the slot mask
constexpr int kRwBarOff = kRwObufBytes + kRwWrowBytes; // 7296
constexpr int kRwRingOff = kRwBarOff + 128;            // 7424: 128-byte aligned
constexpr int kRwRingBytes = kRwWarpBytes - kRwRingOff;
constexpr int kRwVariants = 3;                         // box widths 12 / 20 / 28 texels with S = 4 / 3 / 2 open rows
enum { RW_V12 = 0, RW_V20 = 1, RW_V28 = 2, RW_ZERO = 3, RW_DIRECT = 4, RW_SKIP = 5 };

struct RwMaps {
    CUtensorMap m[4][kRwVariants];                     // [level][variant]; FPN RoI levels: <= 4
};

struct RwCounters {
    unsigned next;         // work-unit queue head (main kernel)
    int n_extra;           // items appended by splits (plan kernel)
    int pad[2];
};

struct __align__(16) RwItem {
    int out_row, plane0, level, variant;
    int x0, y_lo, th, roi;
    int ph0, nph, pw0, npw;
    unsigned long long lastp1;     // byte p: (last footprint row of local output row p) + 1; 0: no valid sample
    int pad[2];
    Tap ytab[14];                  // local sample row 2p+i: low / high RELATIVE to y_lo (low < 0: invalid), weights
    float xw[28];                  // local output column q: h0, l0, h1, l1
    int xoff[14];                  // sample column 2q+i: byte offset of its low tap in a lane's scratch row
    int pad2[2];
};
static_assert(sizeof(RwItem) == 464, "RwItem layout");

__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// mbarrier wait that traps instead of hanging the GPU if a load never lands (a mis-encoded tensor map would
// otherwise spin forever); the counter only runs on the retry path.
__device__ __forceinline__ unsigned mbar_try_wait(unsigned a, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait_guard(unsigned a, unsigned parity) {
    if (mbar_try_wait(a, parity)) return;
    for (unsigned spin = 0; !mbar_try_wait(a, parity); spin++)
        if (spin > (1u << 22)) __trap();
}
__device__ __forceinline__ float lds_f32_off4(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_zero2(unsigned a) {
    asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+4], %1;" :: "r"(a), "f"(0.f) : "memory");
}

// ------------------------------------------------------------------------------------------------------
// PLAN: one HALF-WARP per base block (RoI, group of 7 output rows, group of 7 output columns).  Lane k < 14 of
// the half-warp holds y sample k and x sample k of the block (sample 2p + i: bin p, sub-sample i); extents, slot
// checks and the record are formed with 16-wide shuffles, so a block costs ~300 warp-instructions spread over
// 16 lanes instead of ~3000 serial ones.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int rw_variant_of(int need) { return need <= 12 ? RW_V12 : (need <= 20 ? RW_V20 : (need <= 28 ? RW_V28 : -1)); }
__device__ __forceinline__ int rw_slots_of(int variant) { return variant == RW_V12 ? 4 : (variant == RW_V20 ? 3 : 2); }

// piece k of split kind (0: whole, 1: 4 + rest, 2: 3 + 2 + 2, 3: singles) of n bins -> [b0, b0 + nb)
__device__ __forceinline__ int rw_pieces(int kind, int n) {
    return kind == 0 ? 1 : (kind == 1 ? (n > 4 ? 2 : 1) : (kind == 2 ? (n > 5 ? 3 : (n > 3 ? 2 : 1)) : n));
}
__device__ __forceinline__ void rw_piece(int kind, int n, int k, int& b0, int& nb) {
    if (kind == 0) { b0 = 0; nb = n; }
    else if (kind == 1) { if (n > 4) { b0 = k ? 4 : 0; nb = k ? n - 4 : 4; } else { b0 = 0; nb = n; } }
    else if (kind == 2) { b0 = k == 0 ? 0 : (k == 1 ? 3 : 5); nb = min(n, k == 0 ? 3 : (k == 1 ? 5 : 7)) - b0; }
    else { b0 = k; nb = 1; }
}

// sample coordinate for the fixed 2x2 grid: x / 2 == x * 0.5 exactly, so this equals sample_coord(.., 2) bit for bit
__device__ __forceinline__ float rw_sample_coord(float start, float bin, int p, int i) {
    return __fadd_rn(__fmaf_rn((float)p, bin, start), __fmul_rn(__fmul_rn((float)i + .5f, bin), 0.5f));
}

__device__ __forceinline__ int hw_min(int v, unsigned m) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(m, v, o, 16));
    return v;
}
__device__ __forceinline__ int hw_max(int v, unsigned m) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(m, v, o, 16));
    return v;
}

struct RwLaneTaps {        // lane k of the half-warp: sample k along y and along x (low < 0: invalid / lane >= 14)
    int ylow, yhigh, xlow, xhigh;
    float yl, yh, xl, xh;
    int pfirst, plast;     // extent of the lane's output row (both lanes of a pair hold it); plast < 0: no valid sample
};

// extents of the piece (rows [p0, p0 + np), columns [q0, q0 + nq)); all lanes of the half-warp get the same values
__device__ __forceinline__ void rw_extents(const RwLaneTaps& t, int k, unsigned m, int p0, int np, int q0, int nq,
                                           int& ylo, int& yhi, int& xlo, int& xhi) {
    const bool iny = k >= 2 * p0 && k < 2 * (p0 + np) && t.ylow >= 0, inx = k >= 2 * q0 && k < 2 * (q0 + nq) && t.xlow >= 0;
    ylo = hw_min(iny ? t.ylow : 1 << 30, m); yhi = hw_max(iny ? t.yhigh : -1, m);
    xlo = hw_min(inx ? t.xlow : 1 << 30, m); xhi = hw_max(inx ? t.xhigh : -1, m);
}

// variant that can run the piece, RW_ZERO if it has no valid sample, -1 if none
__device__ __forceinline__ int rw_eval(const RwLaneTaps& t, int k, unsigned m, int p0, int np, int q0, int nq) {
    int ylo, yhi, xlo, xhi;
    rw_extents(t, k, m, p0, np, q0, nq, ylo, yhi, xlo, xhi);
    if (yhi < 0 || xhi < 0) return RW_ZERO;
    if (yhi - ylo + 1 > kRwMaxRows) return -1;
    const int variant = rw_variant_of(xhi - (xlo & ~3) + 1);
    if (variant < 0) return -1;
    const int S = rw_slots_of(variant);
    // output rows p and p + S share an accumulator slot: they must never be open at the same texel row
    const int other_first = __shfl_down_sync(m, t.pfirst, 2 * S, 16), other_last = __shfl_down_sync(m, t.plast, 2 * S, 16);
    const int p = k >> 1;
    const bool clash = k < 14 && p >= p0 && p + S < p0 + np && t.plast >= 0 && other_last >= 0 && other_first <= t.plast;
    return hw_max(clash ? 1 : 0, m) ? -1 : variant;
}

// writes the record of the piece; called by the whole half-warp
__device__ __forceinline__ void rw_write_item(RwItem* it, const RwLaneTaps& t, int k, unsigned m, int variant, int p0, int np, int q0,
                                              int nq, int roi, int out_row, int plane0, int level, int ph_base, int pw_base) {
    int ylo, yhi, xlo, xhi;
    rw_extents(t, k, m, p0, np, q0, nq, ylo, yhi, xlo, xhi);
    if (variant == RW_ZERO || variant == RW_DIRECT) { ylo = 0; yhi = -1; xlo = 0; }
    const int x0 = xlo & ~3;
    // lastp1: byte j = (last row of local output row j) + 1, relative to ylo; gathered from the even lanes
    const int pl = __shfl_sync(m, t.plast, min(2 * (p0 + (k & 7)), 15), 16);
    unsigned lo32 = 0, hi32 = 0;
    if (variant <= RW_V28 && k < 7 && k < np && pl >= 0) {
        const unsigned v = (unsigned)(pl - ylo + 1);
        if (k < 4) lo32 = v << (8 * k); else hi32 = v << (8 * (k - 4));
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { lo32 |= __shfl_xor_sync(m, lo32, o, 16); hi32 |= __shfl_xor_sync(m, hi32, o, 16); }
    int4* o = reinterpret_cast<int4*>(it);
    if (k == 0) {
        o[0] = make_int4(out_row, plane0, level, variant);
        o[1] = make_int4(x0, ylo, yhi - ylo + 1, roi);
        o[2] = make_int4(ph_base + p0, np, pw_base + q0, nq);
        o[3] = make_int4((int)lo32, (int)hi32, 0, 0);
    }
    // ytab: local sample j <- lane 2 * p0 + j
    {
        const int src = min(2 * p0 + k, 15);
        const int sl = __shfl_sync(m, t.ylow, src, 16), sh = __shfl_sync(m, t.yhigh, src, 16);
        const float fl = __shfl_sync(m, t.yl, src, 16), fh = __shfl_sync(m, t.yh, src, 16);
        const bool v = k < 2 * np && sl >= 0;
        if (k < 14 && variant <= RW_V28)
            o[4 + k] = make_int4(v ? sl - ylo : -1, v ? sh - ylo : -1, __float_as_int(fl), __float_as_int(fh));
    }
    // x taps: local sample j <- lane 2 * q0 + j
    const int BX = 12 + 8 * variant;
    const int srcx = min(2 * q0 + k, 15);
    const int sxl = __shfl_sync(m, t.xlow, srcx, 16);
    const float fxl = __shfl_sync(m, t.xl, srcx, 16), fxh = __shfl_sync(m, t.xh, srcx, 16);
    const bool vx = k < 2 * nq && sxl >= 0;
    const int my_off = vx ? 4 * (sxl - x0) : 4 * BX;          // invalid: the two zero cells behind a scratch row
    const float my_h = vx ? fxh : 0.f, my_l = vx ? fxl : 0.f;
    {
        // xw[q] = (h, l) of samples 2q, 2q + 1: lane q < 7 collects from lanes 2q, 2q + 1
        const int a = min(2 * k, 15), b = min(2 * k + 1, 15);
        const float h0 = __shfl_sync(m, my_h, a, 16), l0 = __shfl_sync(m, my_l, a, 16);
        const float h1 = __shfl_sync(m, my_h, b, 16), l1 = __shfl_sync(m, my_l, b, 16);
        if (k < 7 && variant <= RW_V28)
            o[18 + k] = make_int4(__float_as_int(h0), __float_as_int(l0), __float_as_int(h1), __float_as_int(l1));
        // xoff: lane j < 4 collects the offsets of samples 4j .. 4j + 3
        const int c0 = __shfl_sync(m, my_off, min(4 * k, 15), 16), c1 = __shfl_sync(m, my_off, min(4 * k + 1, 15), 16);
        const int c2 = __shfl_sync(m, my_off, min(4 * k + 2, 15), 16), c3 = __shfl_sync(m, my_off, min(4 * k + 3, 15), 16);
        if (k < 4 && variant <= RW_V28) o[25 + k] = make_int4(c0, c1, k == 3 ? 0 : c2, k == 3 ? 0 : c3);
    }
}

// Levels whose row pitch is not a multiple of 16 bytes are copied into zero-padded buffers of the workspace (tensor maps
// need 16-byte strides).  The copy rides in the plan launch: blocks beyond the plan's own do it (one aligned float4 of a
// row per thread and step), so it costs no launch of its own and runs beside the plan.
struct RwPadJobs {
    const float* src[4];
    float* dst[4];
    int W[4], Wp[4];
    long long rows[4];
    int n;              // jobs
    int plan_blocks;    // blocks [0, plan_blocks) plan, the rest pad
};

__device__ __forceinline__ void rw_pad_block(const RwPadJobs& pj, int blk, int nblk) {
    for (int j = 0; j < pj.n; j++) {
        const int q = pj.Wp[j] >> 2, W = pj.W[j], Wp = pj.Wp[j];
        const long long total = pj.rows[j] * q;
        const float* __restrict__ src = pj.src[j];
        float* __restrict__ dst = pj.dst[j];
        for (long long i = (long long)blk * blockDim.x + threadIdx.x; i < total; i += (long long)nblk * blockDim.x) {
            const long long r = i / q;
            const int x4 = (int)(i - r * q) * 4;
            const float* s = src + r * W;
            float4 v;
            v.x = x4 < W ? __ldg(s + x4) : 0.f;
            v.y = x4 + 1 < W ? __ldg(s + x4 + 1) : 0.f;
            v.z = x4 + 2 < W ? __ldg(s + x4 + 2) : 0.f;
            v.w = x4 + 3 < W ? __ldg(s + x4 + 3) : 0.f;
            *reinterpret_cast<float4*>(dst + r * Wp + x4) = v;
        }
    }
}

__global__ void __launch_bounds__(128)
roialign_rw_plan(const __grid_constant__ LevelTable lv, const __grid_constant__ RwPadJobs pj, int channels, int pooled_h,
                 int pooled_w, int num_rois, const float* __restrict__ rois, const int* __restrict__ roi_level,
                 const int* __restrict__ out_index, RwItem* __restrict__ items, int nbase, int cap_extra,
                 RwCounters* __restrict__ ctr) {
    if ((int)blockIdx.x >= pj.plan_blocks) {
        rw_pad_block(pj, (int)blockIdx.x - pj.plan_blocks, (int)gridDim.x - pj.plan_blocks);
        return;
    }
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int k = threadIdx.x & 15;
    const unsigned m = 0xffffu << (threadIdx.x & 16);          // this half of the warp
    if (b >= nbase) return;                                     // half-warp uniform
    const int T = pooled_w / 7, Z = (pooled_h + 6) / 7;
    const int n = b / (Z * T), z = (b / T) % Z, hq = b % T;
    const int level = roi_level ? __ldg(roi_level + n) : 0;
    const int H = lv.h[level], W = lv.w[level];
    const RoiGeom g = roi_geometry(rois + 5 * (size_t)n, lv.scale[level], pooled_h, pooled_w, 2);
    const int out_row = out_index ? __ldg(out_index + n) : n;
    const int nphz = min(7, pooled_h - 7 * z);
    RwLaneTaps t;
    t.ylow = t.yhigh = t.xlow = t.xhigh = -1; t.yl = t.yh = t.xl = t.xh = 0.f;
    if (k < 2 * nphz) {
        const int sy = 14 * z + k;
        const AxisTap a = axis_tap(rw_sample_coord(g.start_h, g.bin_h, sy >> 1, sy & 1), H);
        if (a.valid) { t.ylow = a.low; t.yhigh = a.high; t.yl = a.l; t.yh = a.h; }
    }
    if (k < 14) {
        const int sx = 14 * hq + k;
        const AxisTap a = axis_tap(rw_sample_coord(g.start_w, g.bin_w, sx >> 1, sx & 1), W);
        if (a.valid) {
            t.xlow = a.low; t.xhigh = a.high; t.xl = a.l; t.xh = a.h;
            // a sample clamped to the last column (low == high == W-1, weights 1 / 0) becomes (W-2, W-1) with
            // weights 0 / 1: same value, and the high tap is always "next column"
            if (a.low == a.high && W >= 2) { t.xlow = W - 2; t.xhigh = W - 1; t.xh = 0.f; t.xl = 1.f; }
        }
    }
    {
        const int ol = __shfl_xor_sync(m, t.ylow, 1, 16), oh = __shfl_xor_sync(m, t.yhigh, 1, 16);
        t.pfirst = min(t.ylow >= 0 ? t.ylow : 1 << 30, ol >= 0 ? ol : 1 << 30);
        t.plast = max(t.ylow >= 0 ? t.yhigh : -1, ol >= 0 ? oh : -1);
    }
    const int plane0 = g.batch * channels;
    // split kinds (rows, columns) in order of increasing piece count; the first one whose pieces all run is taken
    const unsigned char kinds[12][2] = {{0, 0}, {1, 0}, {0, 1}, {2, 0}, {1, 1}, {2, 1}, {3, 0}, {0, 3}, {3, 1}, {1, 3}, {2, 3}, {3, 3}};
    int pk = -1, qk = -1;
    if (W >= 2) {
        for (int c = 0; c < 12 && pk < 0; c++) {
            const int np_ = rw_pieces(kinds[c][0], nphz), nq_ = rw_pieces(kinds[c][1], 7);
            bool ok = true;
            for (int i = 0; i < np_ && ok; i++)
                for (int j = 0; j < nq_ && ok; j++) {
                    int p0, np, q0, nq;
                    rw_piece(kinds[c][0], nphz, i, p0, np);
                    rw_piece(kinds[c][1], 7, j, q0, nq);
                    ok = rw_eval(t, k, m, p0, np, q0, nq) >= 0;
                }
            if (ok) { pk = kinds[c][0]; qk = kinds[c][1]; }
        }
    }
    int cnt = pk < 0 ? 1 : rw_pieces(pk, nphz) * rw_pieces(qk, 7);
    int extra0 = 0;
    if (cnt > 1) {
        if (k == 0) extra0 = atomicAdd(&ctr->n_extra, cnt - 1);
        extra0 = __shfl_sync(m, extra0, 0, 16);
        if (extra0 + cnt - 1 > cap_extra) {
            // no room for the pieces: the slots this block owns are marked to be skipped, the block goes direct
            for (int e = extra0 + k; e < min(extra0 + cnt - 1, cap_extra); e += 16) items[nbase + e].variant = RW_SKIP;
            pk = -1; cnt = 1;
        }
    }
    if (pk < 0) {
        rw_write_item(items + b, t, k, m, RW_DIRECT, 0, nphz, 0, 7, n, out_row, plane0, level, 7 * z, 7 * hq);
        return;
    }
    const int nq_ = rw_pieces(qk, 7);
    for (int e = 0; e < cnt; e++) {
        int p0, np, q0, nq;
        rw_piece(pk, nphz, e / nq_, p0, np);
        rw_piece(qk, 7, e % nq_, q0, nq);
        RwItem* it = e == 0 ? items + b : items + nbase + extra0 + e - 1;
        rw_write_item(it, t, k, m, rw_eval(t, k, m, p0, np, q0, nq), p0, np, q0, nq, n, out_row, plane0, level, 7 * z, 7 * hq);
    }
}

// ------------------------------------------------------------------------------------------------------
// MAIN kernel
// ------------------------------------------------------------------------------------------------------
struct RwArgs {
    const RwItem* items;
    RwCounters* ctr;
    float* top;
    const float* rois;
    int nbase, cap_extra;
    int channels, pooled_h, pooled_w;
    int slabs_per_unit, split_log2;        // work unit = (item, slab range); 2^split_log2 ranges per item
    int top_aligned;                       // top is 16-byte aligned (bulk stores)
};

// general epilogue: obuf[c][p * 7 + q] (lane stride 49) -> rows of nq floats in the (PH, PW) output block
__device__ __forceinline__ void rw_store_block(unsigned obuf_s, float* __restrict__ out_c0, int bins, int PW, int np, int nq, int lane) {
    // lane -> up to two elements (p, q) of the np x nq block
    const int run = np * nq;
    const int e0 = lane, e1 = lane + 32;
    const int p0 = e0 / nq, q0 = e0 - p0 * nq, p1 = e1 / nq, q1 = e1 - p1 * nq;
    const bool v0 = e0 < run, v1 = e1 < run;
    const unsigned s0 = obuf_s + 4u * (unsigned)(p0 * 7 + q0), s1 = obuf_s + 4u * (unsigned)(p1 * 7 + q1);
    float* g0 = out_c0 + p0 * PW + q0;
    float* g1 = out_c0 + p1 * PW + q1;
#pragma unroll 4
    for (int c = 0; c < kSlab; c++) {
        if (v0) __stcs(g0 + (size_t)c * bins, lds_off(s0 + 196u * (unsigned)c));
        if (v1) __stcs(g1 + (size_t)c * bins, lds_off(s1 + 196u * (unsigned)c));
    }
}

// Warp-uniform value: REDUX writes a uniform register, so the compiler treats everything derived from it as
// warp-uniform (uniform branches without BSSY / BSYNC, uniform address arithmetic).  v must already be the same in
// every lane.
__device__ __forceinline__ unsigned uni(unsigned v) { return __reduce_or_sync(0xffffffffu, v); }
__device__ __forceinline__ int uni(int v) { return (int)__reduce_or_sync(0xffffffffu, (unsigned)v); }

__device__ __forceinline__ bool elect_one() {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok));
    return ok != 0;
}

template <int BX, int S>
__device__ __forceinline__ void rw_process(const RwArgs& a, const CUtensorMap* map, const RwItem* __restrict__ it, int slab0, int nslab,
                                           unsigned base_s, unsigned& phase, int lane) {
    static_assert(S <= 4 && BX % 4 == 0, "row-weight table: <= 4 weights per row");
    constexpr int SW = 4;
    constexpr int kSlotBytes = 128 * BX;                // one TMA box: 32 channels x BX texels
    constexpr int kScrBytes = 128 * (BX + 3);           // per-lane scratch rows of BX + 3 floats (odd pitch: conflict-free)
    constexpr int kRing = kRwRingBytes - kScrBytes;
    constexpr int NS = kRing / kSlotBytes < kRwMaxSlots ? kRing / kSlotBytes : kRwMaxSlots;
    static_assert(NS >= 2, "ring too small");
    const unsigned obuf_s = base_s, wrow_s = base_s + kRwObufBytes, bar_s = base_s + kRwBarOff, scr_s = base_s + kRwRingOff,
                   ring_s = scr_s + kScrBytes;

    // ---- header (warp-uniform)
    const int4 h0 = __ldg(reinterpret_cast<const int4*>(it)), h1 = __ldg(reinterpret_cast<const int4*>(it) + 1),
               h2 = __ldg(reinterpret_cast<const int4*>(it) + 2), h3 = __ldg(reinterpret_cast<const int4*>(it) + 3);
    const int out_row = uni(h0.x), plane = uni(h0.y) + slab0 * kSlab;
    const int x0 = uni(h1.x), y_lo = uni(h1.y), th = uni(h1.z);
    const int ph0 = uni(h2.x), nph = uni(h2.y), pw0 = uni(h2.z), npw = uni(h2.w);
    const unsigned long long lastp1 = (unsigned long long)uni((unsigned)h3.x) | ((unsigned long long)uni((unsigned)h3.y) << 32);
    const int total = nslab * th;                       // flat row sequence of the unit

    // ---- the elected lane owns the fill cursor: the first NS rows are requested before anything else
    const bool leader = elect_one();
    int fs = 0, fr = 0, left = total;                   // (slab, row) of the next row to request, rows still to request
    if (leader) {
        for (int k = 0; k < NS && left > 0; k++, left--) {
            mbar_expect_tx(bar_s + 8u * (unsigned)k, 128u * BX);
            tma_load_3d(ring_s + (unsigned)(k * kSlotBytes), map, x0, y_lo + fr, plane + fs * kSlab, bar_s + 8u * (unsigned)k);
            if (++fr == th) { fr = 0; fs++; }
        }
    }

    // ---- row-weight table: wrow[r][s] = 0.25 * (sum of the y weights with which row r enters the output row in slot s)
    for (int i = lane; i < th * SW; i += 32) sts_f32(wrow_s + 4u * (unsigned)i, 0.f);
    const unsigned lane_row = ring_s + (unsigned)lane * (BX * 4), lane_scr = scr_s + (unsigned)lane * ((BX + 3) * 4);
    sts_zero2(lane_scr + 4u * BX);                      // the two zero cells behind a scratch row (invalid samples read them)
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (lane < nph) {
            const Tap t = it->ytab[2 * lane + i];
            if (t.low >= 0) {
                const unsigned sl = (unsigned)(lane % S);
                const unsigned al = wrow_s + 4u * ((unsigned)t.low * SW + sl), ah = wrow_s + 4u * ((unsigned)t.high * SW + sl);
                // a tap whose weight is zero is stored as -0.0f: "entered, still multiplies" (the reference multiplies all
                // four taps of a valid sample, so 0 * NaN reaches the bin there too); +0.0f = the row does not enter the slot
                float v = lds_off(al) + 0.25f * t.h;
                sts_f32(al, v == 0.f ? -0.f : v);
                if (t.high != t.low) {
                    v = lds_off(ah) + 0.25f * t.l;
                    sts_f32(ah, v == 0.f ? -0.f : v);
                }
            }
        }
        __syncwarp();
    }
    // ---- x taps (registers): absolute scratch addresses of the low taps, weights
    unsigned xa[14];
    float xw[28];
    {
        const int4* xo4 = reinterpret_cast<const int4*>(it->xoff);
        const float4* xw4 = reinterpret_cast<const float4*>(it->xw);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int4 v = __ldg(xo4 + k);
            xa[4 * k] = lane_scr + v.x; xa[4 * k + 1] = lane_scr + v.y; xa[4 * k + 2] = lane_scr + v.z; xa[4 * k + 3] = lane_scr + v.w;
        }
        xa[12] = lane_scr + __ldg(it->xoff + 12); xa[13] = lane_scr + __ldg(it->xoff + 13);
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const float4 v = __ldg(xw4 + k);
            xw[4 * k] = v.x; xw[4 * k + 1] = v.y; xw[4 * k + 2] = v.z; xw[4 * k + 3] = v.w;
        }
    }
    const int bins = a.pooled_h * a.pooled_w;
    const bool full = npw == 7 && nph == 7 && bins == 49 && a.top_aligned;
    float* out_item = a.top + ((size_t)out_row * a.channels + (size_t)slab0 * kSlab) * bins + ph0 * a.pooled_w + pw0;
    const unsigned lane_obuf = obuf_s + (unsigned)lane * 196u;

    // V[s][x]: the <= S output rows that are open; packed pairs for FFMA2 (two fp32 FMAs per instruction, same rounding)
    float2 V[S][BX / 2];
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int x = 0; x < BX / 2; x++) V[s][x] = make_float2(0.f, 0.f);

    int slot = 0;
    unsigned ready = 0;                                 // the next row's barrier phase is already known to be complete
    for (int sl = 0; sl < nslab; sl++) {
        if (full) {
            if (leader) bulk_store_wait_read();         // the previous slab's bulk store has read obuf
            __syncwarp();
        }
        int p = 0;
        int nxt = (int)(lastp1 & 0xffu);
        unsigned wrow_r = wrow_s;
        for (int r = 0; r < th; r++) {
            const unsigned slot_bar = bar_s + 8u * (unsigned)slot;
            // the barrier test of THIS row was issued one row earlier (a try_wait costs ~90 cycles even when the phase
            // has long completed); only a row whose early test failed spins here
            if (!ready) mbar_wait_guard(slot_bar, (phase >> slot) & 1u);
            phase ^= 1u << slot;
            {
                const int nslot = slot + 1 == NS ? 0 : slot + 1;
                ready = (sl * th + r + 1 < total) ? mbar_try_wait(bar_s + 8u * (unsigned)nslot, (phase >> nslot) & 1u) : 0u;
            }
            const float4 wq = lds_v4(wrow_r);
            const float w[4] = {wq.x, wq.y, wq.z, wq.w};
            wrow_r += SW * 4;
            float2 f[BX / 2];
            const unsigned row_a = lane_row + (unsigned)(slot * kSlotBytes);
#pragma unroll
            for (int j = 0; j < BX / 4; j++) {
                const float4 q = lds_v4(row_a + 16u * j);
                f[2 * j] = make_float2(q.x, q.y); f[2 * j + 1] = make_float2(q.z, q.w);
            }
            __syncwarp();                               // every lane holds its row: the slot can be refilled at once
            if (leader && left > 0) {
                mbar_expect_tx(slot_bar, 128u * BX);
                tma_load_3d(ring_s + (unsigned)(slot * kSlotBytes), map, x0, y_lo + fr, plane + fs * kSlab, slot_bar);
                left--;
                if (++fr == th) { fr = 0; fs++; }
            }
            if (++slot == NS) slot = 0;
            // a row enters a slot iff its table entry is not +0.0f (warp-uniform: every lane reads the same word)
#pragma unroll
            for (int s = 0; s < S; s++) {
                if (__float_as_uint(w[s]) != 0u) {
                    const float2 w2 = make_float2(w[s], w[s]);
#pragma unroll
                    for (int x = 0; x < BX / 2; x++) V[s][x] = __ffma2_rn(w2, f[x], V[s][x]);
                }
            }
            // ---- output rows whose last texel row this was (or that have no valid sample at all)
            while (p < nph && nxt <= r + 1) {
                const unsigned orow = lane_obuf + (unsigned)p * 28u;
                if (nxt == 0) {
                    // no valid sample in this output row: zeros; its slot may already belong to row p + S
#pragma unroll
                    for (int q = 0; q < 7; q++) sts_f32(orow + 4u * q, 0.f);
                } else {
                    const int s = p % S;
                    switch (s) {
#define VOSD_RW_CASE(K) case K: if (K < S) { _Pragma("unroll") for (int x = 0; x < BX / 2; x++) { \
                            sts_f32(lane_scr + 8u * x, V[K < S ? K : 0][x].x); sts_f32(lane_scr + 8u * x + 4u, V[K < S ? K : 0][x].y); \
                            V[K < S ? K : 0][x] = make_float2(0.f, 0.f); } } break;
                        VOSD_RW_CASE(0) VOSD_RW_CASE(1) VOSD_RW_CASE(2) VOSD_RW_CASE(3)
#undef VOSD_RW_CASE
                        default: break;
                    }
                    // lanes only read back their own scratch row: no barrier needed
                    // all 28 tap loads first (the asm loads are issued in program order: interleaving them with the FMAs and
                    // stores of each column would serialise seven load -> FMA -> store round trips)
                    float tp[28];
#pragma unroll
                    for (int q = 0; q < 7; q++) {
                        tp[4 * q] = lds_off(xa[2 * q]); tp[4 * q + 1] = lds_f32_off4(xa[2 * q]);
                        tp[4 * q + 2] = lds_off(xa[2 * q + 1]); tp[4 * q + 3] = lds_f32_off4(xa[2 * q + 1]);
                    }
                    float R[7];
#pragma unroll
                    for (int q = 0; q < 7; q++)
                        R[q] = fmaf(xw[4 * q + 3], tp[4 * q + 3], fmaf(xw[4 * q + 2], tp[4 * q + 2], fmaf(xw[4 * q + 1], tp[4 * q + 1], xw[4 * q] * tp[4 * q])));
#pragma unroll
                    for (int q = 0; q < 7; q++) sts_f32(orow + 4u * q, R[q]);
                }
                p++;
                nxt = (int)((lastp1 >> (8 * p)) & 0xffu);
            }
        }
        // ---- epilogue of the slab
        float* out_s = out_item + (size_t)sl * kSlab * bins;
        if (full) {
            fence_proxy_async_smem();
            __syncwarp();
            if (leader) bulk_store_evict_first(out_s, obuf_s, kRwObufBytes);
        } else {
            __syncwarp();
            rw_store_block(obuf_s, out_s, bins, a.pooled_w, nph, npw, lane);
            __syncwarp();
        }
    }
    if (full) {
        if (leader) bulk_store_wait_read();
        __syncwarp();
    }
}

// blocks without a valid sample: zeros (the reference writes 0 for them)
__device__ __forceinline__ void rw_zero_block(const RwArgs& a, const RwItem* it, int slab0, int nslab, int lane) {
    const int bins = a.pooled_h * a.pooled_w, nq = it->npw, run = it->nph * nq;
    float* out = a.top + ((size_t)it->out_row * a.channels + (size_t)slab0 * kSlab) * bins + it->ph0 * a.pooled_w + it->pw0;
    for (int e = lane; e < run; e += 32) {
        const int p = e / nq, q = e - p * nq;
        for (int c = 0; c < nslab * kSlab; c++) __stcs(out + (size_t)c * bins + p * a.pooled_w + q, 0.f);
    }
}

// blocks no row-window variant can run (footprint wider than 7 x 23 texels per bin row, W < 2, ...): the reference's
// arithmetic element by element, lanes = output elements
__device__ void rw_direct_block(const RwArgs& a, const LevelTable& lv, const RwItem* it, int slab0, int nslab, int lane) {
    const int level = it->level, H = lv.h[level], W = lv.w[level];
    const RoiGeom g = roi_geometry(a.rois + 5 * (size_t)it->roi, lv.scale[level], a.pooled_h, a.pooled_w, 2);
    const int bins = a.pooled_h * a.pooled_w, nq = it->npw, run = it->nph * nq;
    const size_t plane = (size_t)H * W;
    const float* fbase = lv.data[level] + ((size_t)it->plane0 + (size_t)slab0 * kSlab) * plane;
    float* out = a.top + ((size_t)it->out_row * a.channels + (size_t)slab0 * kSlab) * bins;
    for (int e = lane; e < run; e += 32) {
        const int ph = it->ph0 + e / nq, pw = it->pw0 + e % nq;
        AxisTap ty[2], tx[2];
        for (int i = 0; i < 2; i++) {
            ty[i] = axis_tap(sample_coord(g.start_h, g.bin_h, ph, i, 2), H);
            tx[i] = axis_tap(sample_coord(g.start_w, g.bin_w, pw, i, 2), W);
        }
        for (int c = 0; c < nslab * kSlab; c++) {
            const float* d = fbase + (size_t)c * plane;
            float acc = 0.f;
            for (int iy = 0; iy < 2; iy++)
                for (int ix = 0; ix < 2; ix++) {
                    float val = 0.f;
                    if (ty[iy].valid && tx[ix].valid)
                        val = bilinear_value(ty[iy].h, ty[iy].l, tx[ix].h, tx[ix].l, __ldg(d + ty[iy].low * W + tx[ix].low),
                                             __ldg(d + ty[iy].low * W + tx[ix].high), __ldg(d + ty[iy].high * W + tx[ix].low),
                                             __ldg(d + ty[iy].high * W + tx[ix].high));
                    acc = __fadd_rn(acc, val);
                }
            __stcs(out + (size_t)c * bins + ph * a.pooled_w + pw, __fmul_rn(acc, 0.25f));
        }
    }
}

#ifndef VOSD_RW_MAXREG
#define VOSD_RW_MAXREG (512 / ((kRwWarps + 3) / 4) / 8 * 8 > 255 ? 255 : 512 / ((kRwWarps + 3) / 4) / 8 * 8)
#endif
__global__ void __maxnreg__(VOSD_RW_MAXREG)
roialign_fwd_rw(const __grid_constant__ RwMaps maps, const __grid_constant__ LevelTable lv, const __grid_constant__ RwArgs a) {
    extern __shared__ __align__(1024) unsigned char rw_dyn[];
    const int lane = threadIdx.x & 31, warp = uni((int)(threadIdx.x >> 5));
    const unsigned base_s = (unsigned)__cvta_generic_to_shared(rw_dyn) + (unsigned)warp * kRwWarpBytes;
    if (lane == 0) {
        for (int k = 0; k < kRwMaxSlots; k++) mbar_init(base_s + kRwBarOff + 8u * (unsigned)k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;                                 // bit k: parity the next wait on slot k expects
    const int n_extra = min(*reinterpret_cast<volatile const int*>(&a.ctr->n_extra), a.cap_extra);
    const unsigned units = (unsigned)(a.nbase + n_extra) << a.split_log2;
    const int slabs_all = a.channels / kSlab;
    for (;;) {
        unsigned u = 0;
        if (lane == 0) u = atomicAdd(&a.ctr->next, 1u);
        u = uni(u);                                     // lanes != 0 hold 0
        if (u >= units) break;
        const int idx = (int)(u >> a.split_log2), part = (int)(u & ((1u << a.split_log2) - 1u));
        // split pieces first (the blocks of wide RoIs: the longest units), then the base blocks
        const RwItem* it = a.items + (idx < n_extra ? a.nbase + idx : idx - n_extra);
        const int slab0 = part * a.slabs_per_unit;
        const int nslab = min(a.slabs_per_unit, slabs_all - slab0);
        if (nslab <= 0) continue;
        const int variant = uni(__ldg(&it->variant));
        const int level = uni(__ldg(&it->level));
        switch (variant) {
            case RW_V12: rw_process<12, 4>(a, &maps.m[level][0], it, slab0, nslab, base_s, phase, lane); break;
            case RW_V20: rw_process<20, 3>(a, &maps.m[level][1], it, slab0, nslab, base_s, phase, lane); break;
            case RW_V28: rw_process<28, 2>(a, &maps.m[level][2], it, slab0, nslab, base_s, phase, lane); break;
            case RW_ZERO: rw_zero_block(a, it, slab0, nslab, lane); break;
            case RW_DIRECT: rw_direct_block(a, lv, it, slab0, nslab, lane); break;
            default: break;
        }
    }
}

}  // namespace vosd
