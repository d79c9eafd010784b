// Block-wide exact top-m selection (MSD radix select on 64-bit keys) + bitonic sort.
//
// Keys are unique 64-bit integers: (order-preserving score bits << 32) | ~tie_index, so
// "largest key first" == "highest score first, lower index first on equal scores".
// The reference uses np.argpartition + np.argsort (lib/modeling/generate_proposals.py:131-139)
// and np.argsort(-scores) (collect_and_distribute_fpn_rpn_proposals.py:104), both unstable:
// on tie-free inputs the orders coincide bit for bit.
#pragma once
#include "common.cuh"

namespace vosd {

constexpr int kSelThreads = 1024;
constexpr int kRadixBits = 11;
constexpr int kBins = 1 << kRadixBits;     // 2048 bins, two per thread

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
        for (int j = threadIdx.x; j < n; j += kSelThreads) {
            const uint64_t k = key_at(j);
            if (k != 0 && (k & mask) == prefix) atomicAdd(&sh.hist[(uint32_t)(k >> shift) & dmask], 1);
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

// In-place bitonic sort of `P` (power of two) keys in shared memory, largest first.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* keys, int P) {
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
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const uint64_t k = key_at(j);
        if (k != 0 && (k & mask) >= prefix) {
            const int pos = atomicAdd(&sh.counter, 1);
            if (pos < P) keys_out[pos] = k;
        }
    }
    __syncthreads();
    bitonic_sort_desc(keys_out, P);
    return take;
}

__host__ __device__ inline int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace vosd
