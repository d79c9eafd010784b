// RoIAlign forward for CHANNELS-LAST feature maps (memory order N, H, W, C), fed by tensor-mode TMA.
// Included by roialign.cu; same arithmetic and summation order as the separable NCHW kernel (roialign_sep.cuh):
//     out[ph][pw] = sum_y Wy[y][ph] * R[y][pw],   R[y][pw] = sum over the <= 4 x taps of bin pw of wx * F[y][x]
// (reference: lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:65-121, sampling_ratio == 2).
//
// Why a second layout: with channels last, a texel's 32-channel slab is 128 contiguous bytes, so a TMA box
// (c = 32, x = BX, y = 1) of the map lands in shared memory as slot[x][c] -- exactly the layout the consumer reads
// with lanes = channels (consecutive words: conflict-free) -- and the whole producer half of the NCHW kernel (four
// warps, ~295 global-load and ~364 transposing-store wavefronts per (RoI, slab)) disappears.  The box
// origin is (32 * slab, x0, y, n): the innermost coordinate is always 128-byte aligned, x0 and y are free (the
// 16-byte rule of tensor-mode TMA only binds the innermost coordinate).
//
// CTA = one RoI x 7 output rows x 8 / T slabs; EVERY warp is a consumer and its own producer: lane 0 keeps the
// warp's ring of row slots full with one cp.async.bulk.tensor per texel row (FULL mbarrier per slot; a slot is
// refilled right after the warp has read it, so no EMPTY barrier is needed).  Warp `sub` of a team owns output
// columns 7*sub .. 7*sub+6 and keeps acc[7][7] in registers; the epilogue reuses the (then idle) rings of the team
// as obuf[c][bin] and leaves through one bulk store per slab (7x7) or coalesced streaming stores.
#pragma once
#include <cuda.h>

namespace vosd {

// Tuning knobs (A/B builds, tools/ab_build.sh): ring bytes per warp, CTAs per SM, box-width step in texels.
#ifndef VOSD_NHWC_RING
#define VOSD_NHWC_RING 12288
#endif
#ifndef VOSD_NHWC_MINB
#define VOSD_NHWC_MINB 2
#endif
#ifndef VOSD_NHWC_PFDIST
#define VOSD_NHWC_PFDIST 148        // > 0: L2-prefetch the footprint of the RoI this many CTAs ahead (A/B: -6 %)
#endif
#ifndef VOSD_NHWC_BOXSTEP
#define VOSD_NHWC_BOXSTEP 8
#endif
constexpr int kNhwcWarps = 8;
constexpr int kNhwcThreads = 32 * kNhwcWarps;
constexpr int kNhwcRingBytes = VOSD_NHWC_RING;         // per warp: 3 slots of 32 texels .. 12 slots of 8 texels at 12 KB
constexpr int kNhwcTexelBytes = kSlab * 4;             // 128: one texel of a slab
constexpr int kNhwcBoxStep = VOSD_NHWC_BOXSTEP;        // box widths kNhwcBoxStep, 2 * kNhwcBoxStep, ... 32 texels
constexpr int kNhwcBoxes = 32 / kNhwcBoxStep;
constexpr int kNhwcMaxSlots = kNhwcRingBytes / (kNhwcBoxStep * kNhwcTexelBytes);

struct NhwcMaps {
    CUtensorMap m[4][kNhwcBoxes];                      // [level][box width / kNhwcBoxStep - 1]; FPN RoI levels: <= 4
};

__device__ __forceinline__ float lds_off128(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+128];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 :: "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct NhwcShared {
    float wy[kSepMaxRows * kSepWyStride];
    int xoff[28][2];              // per output column: byte offset (texel * 128) of the low tap of its two samples
    float xw[28][4];              // h0, l0, h1, l1 (0 for an invalid sample)
    unsigned long long full[kNhwcWarps][kNhwcMaxSlots];
};

// grid = (RoIs, slab groups of 8 / T, groups of 7 output rows), block = 256, dynamic smem = 8 rings (1024-aligned).
template <int T>
__global__ void __launch_bounds__(kNhwcThreads, VOSD_NHWC_MINB)
roialign_fwd_nhwc(const __grid_constant__ NhwcMaps maps, const __grid_constant__ LevelTable lv, int channels,
                  int pooled_h, const float* __restrict__ rois, const int* __restrict__ roi_level,
                  const int* __restrict__ out_index, float* __restrict__ top) {
    constexpr int PW = 7 * T;
    constexpr int NPH = 7;
    constexpr int kTeams = kNhwcWarps / T;
    constexpr int kRun = NPH * PW;
    constexpr int kObufStride = kRun | 1;
    static_assert(kSlab * kObufStride * 4 <= T * kNhwcRingBytes, "obuf must fit in the team's rings");
    __shared__ NhwcShared sh;
    extern __shared__ __align__(1024) unsigned char nhwc_dyn[];

    const int n = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ph_begin = blockIdx.z * NPH;
    const int nph = min(NPH, pooled_h - ph_begin);
    const int bins = pooled_h * PW;

    const int level = roi_level ? __ldg(roi_level + n) : 0;
    const int H = lv.h[level], W = lv.w[level];
    const RoiGeom g = roi_geometry(rois + 5 * (size_t)n, lv.scale[level], pooled_h, PW, 2);
    const int row = out_index ? __ldg(out_index + n) : n;
    const int sub = warp % T, team = warp / T;

    // Every warp owns its barriers and derives the footprint in registers (lanes = samples, shuffles), so its first
    // TMA loads leave before any CTA-wide barrier: the per-CTA cold start is the load latency alone.
    if (lane == 0) {
        for (int k = 0; k < kNhwcMaxSlots; k++) mbar_init((unsigned)__cvta_generic_to_shared(&sh.full[warp][k]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    // the x tap of sample column k, with the "clamped to the last column" rewrite the consumer relies on
    auto x_tap = [&](int k) {
        const AxisTap t = axis_tap(sample_coord(g.start_w, g.bin_w, k >> 1, k & 1, 2), W);
        Tap e = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        // a sample clamped to the last column becomes (W-2, W-1) with weights (0, 1): the high tap is always "next texel"
        if (t.valid && t.low == t.high && W >= 2) e = Tap{W - 2, W - 1, 1.f, 0.f};
        return e;
    };
    auto y_tap = [&](int k) {                           // sample row k of this CTA's output rows
        const int sy = 2 * ph_begin + k;
        const AxisTap t = axis_tap(sample_coord(g.start_h, g.bin_h, sy >> 1, sy & 1, 2), H);
        return Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
    };
    // footprint: rows shared, one column extent per group of 7 output columns ("half")
    int xlo_h[T], tw_h[T];
    int y_lo, th;
    {
        int lo2 = 1 << 30, hi2 = -1;
        if (lane < 2 * nph) {
            const Tap t = y_tap(lane);
            if (t.low >= 0) { lo2 = t.low; hi2 = t.high; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo2 = min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
            hi2 = max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
        }
        y_lo = lo2; th = hi2 - lo2 + 1;
        // lanes 0..13 of quarter q hold the taps of half q (T <= 2: one pass; T == 4: 14 * 4 = 56 samples, two passes)
#pragma unroll
        for (int h = 0; h < T; h++) {
            int lo = 1 << 30, hi = -1;
            if (lane < 14) {
                const Tap t = x_tap(14 * h + lane);
                if (t.low >= 0) { lo = t.low; hi = t.high; }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
            }
            lo = __shfl_sync(0xffffffffu, lo, 0); hi = __shfl_sync(0xffffffffu, hi, 0);
            xlo_h[h] = lo; tw_h[h] = hi - lo + 1;
        }
    }
    int tw_max = 0, tw_any = 0;
#pragma unroll
    for (int h = 0; h < T; h++) { tw_max = max(tw_max, tw_h[h]); tw_any = max(tw_any, tw_h[h] > 0 ? 1 : 0); }
    const int slabs_all = channels / kSlab;
    const int slab0 = blockIdx.y * kTeams;
    const int nslab = min(kTeams, slabs_all - slab0);
    float* __restrict__ out_roi = top + ((size_t)row * channels + (size_t)slab0 * kSlab) * bins + ph_begin * PW;
    const int group_bins = nph * PW;

    if (!tw_any || th <= 0) {
        for (int e = tid; e < nslab * kSlab * group_bins; e += kNhwcThreads) {
            const int c = e / group_bins, b = e - c * group_bins;
            __stcs(out_roi + (size_t)c * bins + b, 0.f);
        }
        return;
    }
    if (tw_max > 32 || th > kSepMaxRows || W < 2) {
        // Footprint beyond the ring (~8 % of the synthetic RoIs): direct gather with the reference's arithmetic,
        // element by element.  Lanes = channels (every tap is one 128-byte line), the 28 + 14 sample taps come from
        // tables built once per CTA, and the results leave through the (idle) ring memory as coalesced stores.
        __shared__ Tap gy[2 * NPH], gx[2 * PW];
        if (tid < 2 * nph) {
            gy[tid] = y_tap(tid);
        } else if (tid >= 64 && tid < 64 + 2 * PW) {
            const int k = tid - 64;
            const AxisTap t = axis_tap(sample_coord(g.start_w, g.bin_w, k >> 1, k & 1, 2), W);
            gx[k] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        }
        __syncthreads();
        const int nch = nslab * kSlab;
        const int ostride = group_bins | 1;
        float* obuf = reinterpret_cast<float*>(nhwc_dyn);            // nch * ostride floats <= 8 rings
        const float* fbase = lv.data[level] + (size_t)g.batch * H * W * channels + (size_t)slab0 * kSlab;
        for (int b = 0; b < group_bins; b++) {
            const int pr = b / PW, pw = b - pr * PW;
            const Tap ty0 = gy[2 * pr], ty1 = gy[2 * pr + 1], tx0 = gx[2 * pw], tx1 = gx[2 * pw + 1];
            for (int c = tid; c < nch; c += kNhwcThreads) {
                const float* d = fbase + c;
                float acc = 0.f;
#pragma unroll
                for (int iy = 0; iy < 2; iy++) {
                    const Tap ty = iy ? ty1 : ty0;
#pragma unroll
                    for (int ix = 0; ix < 2; ix++) {
                        const Tap tx = ix ? tx1 : tx0;
                        float val = 0.f;
                        if (ty.low >= 0 && tx.low >= 0)
                            val = bilinear_value(ty.h, ty.l, tx.h, tx.l,
                                                 __ldg(d + ((size_t)ty.low * W + tx.low) * channels),
                                                 __ldg(d + ((size_t)ty.low * W + tx.high) * channels),
                                                 __ldg(d + ((size_t)ty.high * W + tx.low) * channels),
                                                 __ldg(d + ((size_t)ty.high * W + tx.high) * channels));
                        acc = __fadd_rn(acc, val);
                    }
                }
                obuf[c * ostride + b] = __fmul_rn(acc, 0.25f);
            }
        }
        __syncthreads();
        for (int e = tid; e < nch * group_bins; e += kNhwcThreads) {
            const int c = e / group_bins, b = e - c * group_bins;
            __stcs(out_roi + (size_t)c * bins + b, obuf[c * ostride + b]);
        }
        return;
    }

    // ---- this warp's ring: box width = footprint width of its half rounded up to 8 texels
    int hx = 0, htw = 0;
#pragma unroll
    for (int h = 0; h < T; h++) if (h == sub) { hx = xlo_h[h]; htw = tw_h[h]; }
    const bool half_empty = htw <= 0;                   // no valid sample column in this half: its outputs are 0
    const int th_my = half_empty ? 0 : th;
    if (half_empty) { hx = 0; htw = 2; }
    const int bsel = (htw + kNhwcBoxStep - 1) / kNhwcBoxStep - 1;
    const int slot_bytes = (bsel + 1) * kNhwcBoxStep * kNhwcTexelBytes;
    const int NS = kNhwcRingBytes / slot_bytes;         // 12, 6, 4, 3 at 12 KB / step 8
    const unsigned dyn_s = (unsigned)__cvta_generic_to_shared(nhwc_dyn);
    if (dyn_s & 127u) __trap();                         // TMA destinations must be 128-byte aligned
    const unsigned ring_s = dyn_s + (unsigned)warp * kNhwcRingBytes;
    const unsigned full_s = (unsigned)__cvta_generic_to_shared(&sh.full[warp][0]);
    const int s = slab0 + team;                         // this warp's slab
    const bool live = team < nslab;
    const CUtensorMap* map = &maps.m[level][bsel];

    // prologue: fill the ring (lane 0), then build the tap tables while the first rows are in flight
    if (live && lane == 0) {
        for (int y = 0; y < min(NS, th_my); y++) {
            mbar_expect_tx(full_s + 8u * (unsigned)y, (unsigned)slot_bytes);
            tma_load_4d(ring_s + (unsigned)(y * slot_bytes), map, s * kSlab, hx, y_lo + y, g.batch, full_s + 8u * (unsigned)y);
        }
    }
    if (VOSD_NHWC_PFDIST > 0 && live) {
        // L2 prefetch of this warp's slab of the footprint of RoI n + VOSD_NHWC_PFDIST (same row group): by the time
        // that CTA starts, its first TMA loads hit L2.  Geometry recomputed in registers, lanes = texel rows.
        const int n2 = n + VOSD_NHWC_PFDIST;
        if (n2 < (int)gridDim.x) {
            const int level2 = roi_level ? __ldg(roi_level + n2) : 0;
            const int H2 = lv.h[level2], W2 = lv.w[level2];
            const RoiGeom g2 = roi_geometry(rois + 5 * (size_t)n2, lv.scale[level2], pooled_h, PW, 2);
            int lo = 1 << 30, hi = -1, xl = 1 << 30, xh = -1;
            if (lane < 2 * nph) {
                const int sy = 2 * ph_begin + lane;
                const AxisTap t = axis_tap(sample_coord(g2.start_h, g2.bin_h, sy >> 1, sy & 1, 2), H2);
                if (t.valid) { lo = t.low; hi = t.high; }
            }
            if (lane < 14) {
                const int k = 14 * sub + lane;
                const AxisTap t = axis_tap(sample_coord(g2.start_w, g2.bin_w, k >> 1, k & 1, 2), W2);
                if (t.valid) { xl = t.low; xh = t.high; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                xl = min(xl, __shfl_xor_sync(0xffffffffu, xl, o));
                xh = max(xh, __shfl_xor_sync(0xffffffffu, xh, o));
            }
            const int tw2 = xh - xl + 1, th2 = hi - lo + 1;
            if (tw2 > 0 && tw2 <= 32 && th2 > 0 && th2 <= kSepMaxRows) {
                const CUtensorMap* map2 = &maps.m[level2][(tw2 + kNhwcBoxStep - 1) / kNhwcBoxStep - 1];
                for (int y = lane; y < th2; y += 32) tma_prefetch_4d(map2, s * kSlab, min(xl, W2 - 2 < 0 ? 0 : xl), lo + y, g2.batch);
            }
        }
    }
    if (tid < PW) {
        const Tap t0 = x_tap(2 * tid), t1 = x_tap(2 * tid + 1);
        int ox = 0;
#pragma unroll
        for (int h = 0; h < T; h++) if (h == tid / 7) ox = tw_h[h] > 0 ? xlo_h[h] : 0;
        sh.xoff[tid][0] = t0.low >= 0 ? (t0.low - ox) * kNhwcTexelBytes : 0;
        sh.xoff[tid][1] = t1.low >= 0 ? (t1.low - ox) * kNhwcTexelBytes : 0;
        sh.xw[tid][0] = t0.low >= 0 ? t0.h : 0.f; sh.xw[tid][1] = t0.low >= 0 ? t0.l : 0.f;
        sh.xw[tid][2] = t1.low >= 0 ? t1.h : 0.f; sh.xw[tid][3] = t1.low >= 0 ? t1.l : 0.f;
    } else if (tid >= 32 && tid < 32 + NPH) {
        // thread = output row pr: it owns column pr of Wy (zero fill + its two sample rows; 0.25 = 1 / count, exact)
        const int pr = tid - 32;
        for (int y = 0; y < th; y++) sh.wy[y * kSepWyStride + pr] = 0.f;
        if (pr < nph) {
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const Tap t = y_tap(2 * pr + k);
                if (t.low >= 0) {
                    sh.wy[(t.low - y_lo) * kSepWyStride + pr] += 0.25f * t.h;
                    if (t.high != t.low) sh.wy[(t.high - y_lo) * kSepWyStride + pr] += 0.25f * t.l;
                }
            }
        }
    }
    __syncthreads();
    if (!live) return;                                  // (no barrier below involves warps of other teams)

    unsigned a0[7], a1[7];
    float xw[7][4];
#pragma unroll
    for (int i = 0; i < 7; i++) {
        a0[i] = ring_s + (unsigned)lane * 4u + (unsigned)sh.xoff[7 * sub + i][0];
        a1[i] = ring_s + (unsigned)lane * 4u + (unsigned)sh.xoff[7 * sub + i][1];
#pragma unroll
        for (int k = 0; k < 4; k++) xw[i][k] = sh.xw[7 * sub + i][k];
    }
    unsigned wy_a = (unsigned)__cvta_generic_to_shared(sh.wy);
    float acc[NPH][7];
#pragma unroll
    for (int p = 0; p < NPH; p++)
#pragma unroll
        for (int i = 0; i < 7; i++) acc[p][i] = 0.f;
    int slot = 0;
    unsigned parity = 0, c_off = 0;
    for (int y = 0; y < th_my; y++) {
        mbar_wait(full_s + 8u * (unsigned)slot, parity);
        const float4 q0 = lds_v4(wy_a), q1 = lds_v4(wy_a + 16);
        const float wy[7] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};
        float f[7][4];
#pragma unroll
        for (int i = 0; i < 7; i++) {
            f[i][0] = lds_off(a0[i] + c_off);
            f[i][1] = lds_off128(a0[i] + c_off);
            f[i][2] = lds_off(a1[i] + c_off);
            f[i][3] = lds_off128(a1[i] + c_off);
        }
#pragma unroll
        for (int i = 0; i < 7; i++) {
            const float r = fmaf(xw[i][3], f[i][3], fmaf(xw[i][2], f[i][2], fmaf(xw[i][1], f[i][1], xw[i][0] * f[i][0])));
            // (updating only the 2-3 output rows with a non-zero Wy[y][p] -- a set-bit loop with a switch -- measured
            //  slower in the bench step: 0.486 -> 0.523 ms; the branch overhead outweighs the 34 saved FMAs)
#pragma unroll
            for (int p = 0; p < NPH; p++) acc[p][i] = fmaf(wy[p], r, acc[p][i]);
        }
        __syncwarp();                                   // every lane's taps of this slot are in registers
        if (lane == 0 && y + NS < th_my) {              // refill the slot with the row NS further down
            mbar_expect_tx(full_s + 8u * (unsigned)slot, (unsigned)slot_bytes);
            tma_load_4d(ring_s + c_off, map, s * kSlab, hx, y_lo + y + NS, g.batch, full_s + 8u * (unsigned)slot);
        }
        wy_a += kSepWyStride * 4;
        c_off += slot_bytes;
        if (++slot == NS) { slot = 0; c_off = 0; parity ^= 1u; }
    }

    // ---- epilogue: the team's rings are idle now and become obuf[c][bin] (odd stride)
    team_sync<T>(team);                                 // teammates have consumed their last slot
    float* obuf = reinterpret_cast<float*>(nhwc_dyn + (size_t)team * T * kNhwcRingBytes);
    float* __restrict__ out_s = out_roi + (size_t)team * kSlab * bins;
    const int run = nph * PW;
#pragma unroll
    for (int p = 0; p < NPH; p++)
#pragma unroll
        for (int i = 0; i < 7; i++)
            obuf[lane * kObufStride + p * PW + 7 * sub + i] = acc[p][i];
    const bool bulk = T == 1 && run == bins && (reinterpret_cast<uintptr_t>(top) & 15) == 0;
    if (bulk) {
        // obuf[c][bin] with stride 49 IS the output layout of the slab: one bulk shared -> global store
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            bulk_store_evict_first(out_s, (unsigned)__cvta_generic_to_shared(obuf), kSlab * kRun * 4);
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // shared memory must outlive the store
        }
        return;
    }
    team_sync<T>(team);
    for (int i = lane + 32 * sub; i < kSlab * kRun; i += 32 * T) {
        const int c = i / kRun, b = i - c * kRun;
        if (b < run) __stcs(out_s + (size_t)c * bins + b, obuf[c * kObufStride + b]);
    }
}

}  // namespace vosd
