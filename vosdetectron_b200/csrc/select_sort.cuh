// Block-wide exact top-m selection (MSD radix select on 64-bit keys) + bitonic sort.
//
// Keys are unique 64-bit integers: (order-preserving score bits << 32) | ~tie_index, so
// "largest key first" == "highest score first, lower index first on equal scores".
// The reference uses np.argpartition + np.argsort (lib/modeling/generate_proposals.py:131-139)
// and np.argsort(-scores) (collect_and_distribute_fpn_rpn_proposals.py:104), both unstable:
// on tie-free inputs the orders coincide bit for bit.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace vosd {

constexpr int kSelThreads = 1024;
constexpr int kRadixBits = 11;
constexpr int kBins = 1 << kRadixBits;     // 2048 bins, two per thread
constexpr int kKeyBatch = 8;               // keys fetched per thread before they are consumed

struct SelectShared {
    int hist[kBins];
    int warp_sums[32];
    int found_bin;
    int found_above;
    int found_count;
    int counter;
};

// Exclusive prefix sum over the block (blockDim.x == kSelThreads); returns the exclusive
// value, `total` gets the block sum.  Two __syncthreads inside.
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = warp_sums[lane];
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sums[lane] = winc;           // inclusive sums of warp totals
    }
    __syncthreads();
    total = warp_sums[31];
    const int warp_off = warp == 0 ? 0 : warp_sums[warp - 1];
    const int r = warp_off + inc - v;
    __syncthreads();                        // warp_sums may be reused by the caller
    return r;
}

// Finds (mask, prefix) such that exactly `m` of the n keys satisfy (key & mask) >= prefix.
// KeyFn: uint64_t operator()(int j) for j in [0, n) (any enumeration order).  Non-zero keys are
// unique; a zero key marks an empty slot and is ignored.  Requires 0 < m < #non-zero keys.
template <class KeyFn>
__device__ void radix_select(const KeyFn& key_at, int n, int m, SelectShared& sh,
                             uint64_t& out_mask, uint64_t& out_prefix) {
    uint64_t mask = 0, prefix = 0;
    int need = m;
    int shift = 64;
    while (shift > 0) {
        const int bits = shift >= kRadixBits ? kRadixBits : shift;
        shift -= bits;
        const uint32_t dmask = (1u << bits) - 1u;
        for (int b = threadIdx.x; b < kBins; b += kSelThreads) sh.hist[b] = 0;
        __syncthreads();
        // kKeyBatch keys per thread are fetched before the first atomic: the loads of a batch are in flight together
        // (one key per iteration exposes the full L2 latency per key: the passes were bound by it)
        for (int j = threadIdx.x; j < n; j += kKeyBatch * kSelThreads) {
            uint64_t kb[kKeyBatch];
#pragma unroll
            for (int u = 0; u < kKeyBatch; u++) kb[u] = j + u * kSelThreads < n ? key_at(j + u * kSelThreads) : 0;
#pragma unroll
            for (int u = 0; u < kKeyBatch; u++)
                if (kb[u] != 0 && (kb[u] & mask) == prefix) atomicAdd(&sh.hist[(uint32_t)(kb[u] >> shift) & dmask], 1);
        }
        __syncthreads();
        // descending suffix scan: thread t owns bins (kBins-1-2t) and (kBins-2-2t)
        const int b0 = kBins - 1 - 2 * threadIdx.x, b1 = b0 - 1;
        const int h0 = sh.hist[b0], h1 = sh.hist[b1];
        int total;
        const int above0 = block_exclusive_scan(h0 + h1, sh.warp_sums, total);
        const int above1 = above0 + h0;
        if (above0 < need && need <= above0 + h0) { sh.found_bin = b0; sh.found_above = above0; sh.found_count = h0; }
        if (above1 < need && need <= above1 + h1) { sh.found_bin = b1; sh.found_above = above1; sh.found_count = h1; }
        __syncthreads();
        need -= sh.found_above;
        prefix |= (uint64_t)sh.found_bin << shift;
        mask |= (uint64_t)dmask << shift;
        const bool done = sh.found_count == need;
        __syncthreads();
        if (done) break;
    }
    out_mask = mask;
    out_prefix = prefix;
}

// P = K * 1024 keys, blockDim.x == 1024: K keys per thread in registers (element e = thread + 1024 * k).  Stages whose
// partner sits in the same warp (j < 32) are two shuffles per key, stages with j >= 1024 pair two registers of the same
// thread, only the 32 <= j < 1024 stages go through shared memory (for P = 1024: 15 of 55 stages).
template <int K>
__device__ __forceinline__ void bitonic_sort_desc_regs(uint64_t* keys) {
    const int i = threadIdx.x;
    uint64_t v[K];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = keys[i + 1024 * k];
    for (int kk = 2; kk <= 1024 * K; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            if (j >= 1024) {
                // partner register k ^ dk, dk = j / 1024 (unrolled over the possible dk: every index is static)
#pragma unroll
                for (int dk = K >> 1; dk >= 1; dk >>= 1) {
                    if (j != (dk << 10)) continue;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        if ((k & dk) == 0) {
                            const int e = i + 1024 * k;
                            const bool desc = (e & kk) == 0;  // lower element keeps the max when the run is descending
                            const uint64_t a = v[k], b = v[k | dk];
                            const bool sw = desc ? (a < b) : (a > b);
                            v[k] = sw ? b : a;
                            v[k | dk] = sw ? a : b;
                        }
                    }
                }
                continue;
            }
            uint64_t w[K];
            if (j >= 32) {
                __syncthreads();                        // the previous exchange has been read
#pragma unroll
                for (int k = 0; k < K; k++) keys[i + 1024 * k] = v[k];
                __syncthreads();
#pragma unroll
                for (int k = 0; k < K; k++) w[k] = keys[(i ^ j) + 1024 * k];
            } else {
#pragma unroll
                for (int k = 0; k < K; k++) w[k] = __shfl_xor_sync(0xffffffffu, v[k], j);
            }
#pragma unroll
            for (int k = 0; k < K; k++) {
                const int e = i + 1024 * k;
                const bool keep_max = ((e & kk) == 0) == ((e & j) == 0);
                v[k] = keep_max ? (v[k] > w[k] ? v[k] : w[k]) : (v[k] < w[k] ? v[k] : w[k]);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++) keys[i + 1024 * k] = v[k];
    __syncthreads();
}

// In-place bitonic sort of `P` (power of two) keys in shared memory, largest first.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P) {
    if (blockDim.x == 1024) {
        switch (P) {
            case 1024: bitonic_sort_desc_regs<1>(keys); return;
            case 2048: bitonic_sort_desc_regs<2>(keys); return;
            case 4096: bitonic_sort_desc_regs<4>(keys); return;
            case 8192: bitonic_sort_desc_regs<8>(keys); return;
            default: break;
        }
    }
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const uint64_t a = keys[i], b = keys[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// Selects the m largest of the n_valid non-zero keys among key_at(0..n) into keys_out[0..take)
// sorted descending; pads [take, P) with 0.  take = min(m, n_valid) is returned.
// P >= take, power of two.
template <class KeyFn>
__device__ int select_and_sort(const KeyFn& key_at, int n, int n_valid, int m, uint64_t* keys_out,
                               int P, SelectShared& sh) {
    const int take = m < n_valid ? m : n_valid;
    if (threadIdx.x == 0) sh.counter = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys_out[i] = 0;
    __syncthreads();
    if (take == 0) return 0;
    uint64_t mask = 0, prefix = 0;
    if (take < n_valid) radix_select(key_at, n, take, sh, mask, prefix);
    for (int j = threadIdx.x; j < n; j += kKeyBatch * blockDim.x) {
        uint64_t kb[kKeyBatch];
#pragma unroll
        for (int u = 0; u < kKeyBatch; u++) kb[u] = j + u * (int)blockDim.x < n ? key_at(j + u * (int)blockDim.x) : 0;
#pragma unroll
        for (int u = 0; u < kKeyBatch; u++)
            if (kb[u] != 0 && (kb[u] & mask) >= prefix) {
                const int pos = atomicAdd(&sh.counter, 1);
                if (pos < P) keys_out[pos] = kb[u];
            }
    }
    __syncthreads();
    bitonic_sort_desc(keys_out, P);
    return take;
}

// ------------------------------------------------------------------------------------
// Cluster variant: the n keys of ONE segment are split over the CTAs of a thread-block cluster.
// Every pass each CTA histograms its slice into its own shared memory; after one cluster
// barrier all CTAs read the peer histograms through distributed shared memory (DSMEM) and
// derive the same threshold bin redundantly, so nothing has to be broadcast.  Histograms are
// double-buffered: one cluster barrier per pass.
// ------------------------------------------------------------------------------------
struct ClusterSelectShared {
    int hist[2][kBins];
    int warp_sums[32];
    int found_bin;
    int found_above;
    int found_count;
    int counter;          // keys this CTA selected from its slice
    int ncand;            // keys of this CTA's slice inside the threshold bin of the leading digit
};

// Cluster-wide select + sort.  On return rank 0 holds the `take` largest keys sorted descending in
// its keys_out[0..take) (padded with 0 to P); the other ranks may exit.  All threads of all CTAs of
// the cluster must call it.
//
// Scans of the score plane are what the kernel costs (every key is rebuilt from its score and index each time), so
// only TWO of them touch global memory: the histogram of the leading digit, and a split scan that appends the keys
// above the threshold bin to the output (they are selected whatever the later digits say) and the keys INSIDE the
// threshold bin to `cand` (shared memory, `cap_cand` keys per CTA).  The remaining digits are resolved on `cand`
// alone.  If the threshold bin of some CTA does not fit (a degenerate score distribution), every CTA of the cluster
// falls back to rescanning its slice per digit, as before.  cap_cand == 0 disables the candidate buffer.
template <class KeyFn>
__device__ int select_and_sort_cluster(cooperative_groups::cluster_group& cluster, const KeyFn& key_at,
                                       int n, int m, uint64_t* keys_out, int P, ClusterSelectShared& sh,
                                       uint64_t* cand = nullptr, int cap_cand = 0) {
    const unsigned rank = cluster.block_rank(), nranks = cluster.num_blocks();
    const int take = m < n ? m : n;
    const int chunk = (n + (int)nranks - 1) / (int)nranks;
    const int j0 = min(n, (int)rank * chunk), j1 = min(n, j0 + chunk);
    if (threadIdx.x == 0) { sh.counter = 0; sh.ncand = 0; }
    __syncthreads();
    uint64_t mask = 0, prefix = 0;
    bool from_cand = false;                           // later digits are resolved on cand[] instead of the slice
    if (take < n) {
        int need = take, shift = 64, buf = 0;
        bool first = true;
        while (shift > 0) {
            const int bits = shift >= kRadixBits ? kRadixBits : shift;
            shift -= bits;
            const uint32_t dmask = (1u << bits) - 1u;
            int* h = sh.hist[buf];
            for (int b = threadIdx.x; b < kBins; b += kSelThreads) h[b] = 0;
            __syncthreads();
            if (!from_cand) {
                for (int j = j0 + threadIdx.x; j < j1; j += kKeyBatch * kSelThreads) {
                    uint64_t kb[kKeyBatch];
#pragma unroll
                    for (int u = 0; u < kKeyBatch; u++) kb[u] = j + u * kSelThreads < j1 ? key_at(j + u * kSelThreads) : 0;
#pragma unroll
                    for (int u = 0; u < kKeyBatch; u++)
                        if (kb[u] != 0 && (kb[u] & mask) == prefix) atomicAdd(&h[(uint32_t)(kb[u] >> shift) & dmask], 1);
                }
            } else {
                const int nc = sh.ncand;
                for (int i = threadIdx.x; i < nc; i += kSelThreads) {
                    const uint64_t k = cand[i];
                    if ((k & mask) == prefix) atomicAdd(&h[(uint32_t)(k >> shift) & dmask], 1);
                }
            }
            cluster.sync();
            const int b0 = kBins - 1 - 2 * threadIdx.x, b1 = b0 - 1;
            int h0 = 0, h1 = 0;
            for (unsigned r = 0; r < nranks; r++) {
                const int* rh = cluster.map_shared_rank(h, r);
                h0 += rh[b0];
                h1 += rh[b1];
            }
            int total;
            const int above0 = block_exclusive_scan(h0 + h1, sh.warp_sums, total);
            const int above1 = above0 + h0;
            if (above0 < need && need <= above0 + h0) { sh.found_bin = b0; sh.found_above = above0; sh.found_count = h0; }
            if (above1 < need && need <= above1 + h1) { sh.found_bin = b1; sh.found_above = above1; sh.found_count = h1; }
            __syncthreads();
            need -= sh.found_above;
            prefix |= (uint64_t)sh.found_bin << shift;
            mask |= (uint64_t)dmask << shift;
            const bool done = sh.found_count == need;
            __syncthreads();
            buf ^= 1;
            if (done) break;
            if (first && cap_cand > 0) {
                // split scan (mask covers the leading digit only): above the threshold bin -> selected, inside -> candidate
                for (int j = j0 + threadIdx.x; j < j1; j += kKeyBatch * kSelThreads) {
                    uint64_t kb[kKeyBatch];
#pragma unroll
                    for (int u = 0; u < kKeyBatch; u++) kb[u] = j + u * kSelThreads < j1 ? key_at(j + u * kSelThreads) : 0;
#pragma unroll
                    for (int u = 0; u < kKeyBatch; u++) {
                        if (kb[u] == 0) continue;
                        const uint64_t top = kb[u] & mask;
                        if (top > prefix) {
                            const int pos = atomicAdd(&sh.counter, 1);
                            if (pos < P) keys_out[pos] = kb[u];
                        } else if (top == prefix) {
                            const int c = atomicAdd(&sh.ncand, 1);
                            if (c < cap_cand) cand[c] = kb[u];
                        }
                    }
                }
                cluster.sync();                        // every CTA's candidate count is final
                bool fits = true;
                for (unsigned r = 0; r < nranks; r++) fits = fits && *cluster.map_shared_rank(&sh.ncand, r) <= cap_cand;
                if (fits) {
                    from_cand = true;
                } else {
                    __syncthreads();
                    if (threadIdx.x == 0) sh.counter = 0;   // nothing is collected yet on the rescan path
                    __syncthreads();
                }
            }
            first = false;
        }
    }
    if (from_cand) {
        const int nc = sh.ncand;
        for (int i = threadIdx.x; i < nc; i += kSelThreads) {
            const uint64_t k = cand[i];
            if ((k & mask) >= prefix) {
                const int pos = atomicAdd(&sh.counter, 1);
                if (pos < P) keys_out[pos] = k;
            }
        }
    } else {
        for (int j = j0 + threadIdx.x; j < j1; j += kKeyBatch * kSelThreads) {
            uint64_t kb[kKeyBatch];
#pragma unroll
            for (int u = 0; u < kKeyBatch; u++) kb[u] = j + u * kSelThreads < j1 ? key_at(j + u * kSelThreads) : 0;
#pragma unroll
            for (int u = 0; u < kKeyBatch; u++)
                if (kb[u] != 0 && (kb[u] & mask) >= prefix) {
                    const int pos = atomicAdd(&sh.counter, 1);
                    if (pos < P) keys_out[pos] = kb[u];
                }
        }
    }
    cluster.sync();                                   // every slice collected, counters final
    if (rank == 0) {
        int off = sh.counter;
        for (unsigned r = 1; r < nranks; r++) {
            const int cnt = *cluster.map_shared_rank(&sh.counter, r);
            const uint64_t* rk = cluster.map_shared_rank(keys_out, r);
            for (int i = threadIdx.x; i < cnt; i += kSelThreads)
                if (off + i < P) keys_out[off + i] = rk[i];
            off += cnt;
        }
        for (int i = off + threadIdx.x; i < P; i += kSelThreads) keys_out[i] = 0;
    }
    cluster.sync();                                   // peers may leave once rank 0 has copied
    if (rank == 0) {
        __syncthreads();
        bitonic_sort_desc(keys_out, P);
    }
    return take;
}

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace vosd
