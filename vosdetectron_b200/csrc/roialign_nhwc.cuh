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

constexpr int kNhwcWarps = 8;
constexpr int kNhwcThreads = 32 * kNhwcWarps;
constexpr int kNhwcRingBytes = 12288;                  // per warp: 3 slots of 32 texels .. 12 slots of 8 texels
constexpr int kNhwcTexelBytes = kSlab * 4;             // 128: one texel of a slab
constexpr int kNhwcBoxes = 4;                          // box widths 8, 16, 24, 32 texels
constexpr int kNhwcMaxSlots = 12;

struct NhwcMaps {
    CUtensorMap m[VOSD_MAX_LEVELS][kNhwcBoxes];        // [level][box width / 8 - 1]
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

struct NhwcShared {
    Tap ytab[16];                 // sample rows of this CTA's 7 output rows
    Tap xtab[64];                 // all sample columns (2 * PW <= 56)
    float wy[kSepMaxRows * kSepWyStride];
    int xoff[28][2];              // per output column: byte offset (texel * 128) of the low tap of its two samples
    float xw[28][4];              // h0, l0, h1, l1 (0 for an invalid sample)
    unsigned long long full[kNhwcWarps][kNhwcMaxSlots];
};

// grid = (RoIs, slab groups of 8 / T, groups of 7 output rows), block = 256, dynamic smem = 8 rings (1024-aligned).
template <int T>
__global__ void __launch_bounds__(kNhwcThreads, 2)
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
    if (tid < 2 * nph) {
        const int sy = 2 * ph_begin + tid;
        const AxisTap t = axis_tap(sample_coord(g.start_h, g.bin_h, sy >> 1, sy & 1, 2), H);
        sh.ytab[tid] = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
    } else if (tid >= 64 && tid < 64 + 2 * PW) {
        const int k = tid - 64;
        const AxisTap t = axis_tap(sample_coord(g.start_w, g.bin_w, k >> 1, k & 1, 2), W);
        Tap e = Tap{t.valid ? t.low : -1, t.high, t.l, t.h};
        // a sample clamped to the last column becomes (W-2, W-1) with weights (0, 1): the high tap is always "next texel"
        if (t.valid && t.low == t.high && W >= 2) e = Tap{W - 2, W - 1, 1.f, 0.f};
        sh.xtab[k] = e;
    } else if (tid >= 128 && tid < 128 + kNhwcWarps * kNhwcMaxSlots) {
        mbar_init((unsigned)__cvta_generic_to_shared(&sh.full[0][0]) + 8u * (unsigned)(tid - 128), 1);
    }
    for (int i = tid; i < kSepMaxRows * kSepWyStride; i += kNhwcThreads) sh.wy[i] = 0.f;
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // footprint: rows shared, one column extent per group of 7 output columns ("half")
    const int sub = warp % T, team = warp / T;
    int xlo_h[T], tw_h[T];
    int y_lo, th;
    {
        int lo2 = 1 << 30, hi2 = -1;
        if (lane < 2 * nph) {
            const Tap t = sh.ytab[lane];
            if (t.low >= 0) { lo2 = t.low; hi2 = t.high; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo2 = min(lo2, __shfl_xor_sync(0xffffffffu, lo2, o));
            hi2 = max(hi2, __shfl_xor_sync(0xffffffffu, hi2, o));
        }
        y_lo = lo2; th = hi2 - lo2 + 1;
#pragma unroll
        for (int h = 0; h < T; h++) {
            int lo = 1 << 30, hi = -1;
            if (lane < 14) {
                const Tap t = sh.xtab[14 * h + lane];
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
        // footprint beyond the ring: direct gather, the reference's arithmetic element by element (lanes = channels)
        const float* fbase = lv.data[level] + (size_t)g.batch * H * W * channels + (size_t)slab0 * kSlab;
        for (int e = tid; e < nslab * kSlab * group_bins; e += kNhwcThreads) {
            const int c = e % (nslab * kSlab), b = e / (nslab * kSlab);
            const int ph = ph_begin + b / PW, pw = b % PW;
            const float* d = fbase + c;
            float acc = 0.f;
            for (int iy = 0; iy < 2; iy++) {
                const AxisTap ty = axis_tap(sample_coord(g.start_h, g.bin_h, ph, iy, 2), H);
                for (int ix = 0; ix < 2; ix++) {
                    const AxisTap tx = axis_tap(sample_coord(g.start_w, g.bin_w, pw, ix, 2), W);
                    float val = 0.f;
                    if (ty.valid && tx.valid)
                        val = bilinear_value(ty.h, ty.l, tx.h, tx.l,
                                             __ldg(d + ((size_t)ty.low * W + tx.low) * channels),
                                             __ldg(d + ((size_t)ty.low * W + tx.high) * channels),
                                             __ldg(d + ((size_t)ty.high * W + tx.low) * channels),
                                             __ldg(d + ((size_t)ty.high * W + tx.high) * channels));
                    acc = __fadd_rn(acc, val);
                }
            }
            __stcs(out_roi + (size_t)c * bins + b, __fmul_rn(acc, 0.25f));
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
    const int bsel = (htw + 7) / 8 - 1;                 // 0..3
    const int slot_bytes = (bsel + 1) * 8 * kNhwcTexelBytes;
    const int NS = kNhwcRingBytes / slot_bytes;         // 12, 6, 4, 3
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
    if (tid < PW) {
        const Tap t0 = sh.xtab[2 * tid], t1 = sh.xtab[2 * tid + 1];
        int ox = 0;
#pragma unroll
        for (int h = 0; h < T; h++) if (h == tid / 7) ox = tw_h[h] > 0 ? xlo_h[h] : 0;
        sh.xoff[tid][0] = t0.low >= 0 ? (t0.low - ox) * kNhwcTexelBytes : 0;
        sh.xoff[tid][1] = t1.low >= 0 ? (t1.low - ox) * kNhwcTexelBytes : 0;
        sh.xw[tid][0] = t0.low >= 0 ? t0.h : 0.f; sh.xw[tid][1] = t0.low >= 0 ? t0.l : 0.f;
        sh.xw[tid][2] = t1.low >= 0 ? t1.h : 0.f; sh.xw[tid][3] = t1.low >= 0 ? t1.l : 0.f;
    } else if (tid >= 32 && tid < 32 + nph) {
        const int pr = tid - 32;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const Tap t = sh.ytab[2 * pr + k];
            if (t.low >= 0) {
                sh.wy[(t.low - y_lo) * kSepWyStride + pr] += 0.25f * t.h;
                if (t.high != t.low) sh.wy[(t.high - y_lo) * kSepWyStride + pr] += 0.25f * t.l;
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
