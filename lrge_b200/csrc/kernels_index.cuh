// kernels_index.cuh -- K2/K3: target minimizer index in HBM.
//
// Replaces minimap2 mm_idx_add / worker_post / mm_idx_get / mm_idx_cal_max_occ (index.c) and
// mm_mapopt_update (options.c), SURVEY.md rows a4/a5.  minimap2 buckets minimizers by the low 14 hash
// bits, sorts each bucket and builds a hash table per bucket; what a lookup returns is the list of
// y = rid<<32|pos<<1|strand of one key in ascending order.  Here all target minimizers are radix-sorted
// by hash (stable, so y stays ascending inside a key) into one CSR-like array, and a direct-address
// table over the TOP b hash bits gives the sub-range to search:
//     tx[M], ty[M]   sorted by tx>>8;   boff[2^b + 1]:  bucket(key) = key >> (2k - b);   tk32[M] = low 32 key bits
// A lookup is one read of two adjacent boff entries plus the bucket's <= 8 compact keys fetched with independent loads.
#pragma once
#include "common.cuh"

namespace lrge {

constexpr int OCC_BINS = 4096;   // occurrence histogram bins; the last bin collects everything >= OCC_BINS-1

// boff[b] = first i with bucket(tx[i]) >= b, for b in [0, nb]; nb = 2^bits
__global__ void bucket_bounds_kernel(const uint64_t *__restrict__ tx, uint64_t M, int shift, uint32_t nb,
                                     uint32_t *__restrict__ boff) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > M) return;
    const int64_t b_prev = i == 0 ? -1 : (int64_t)(tx[i - 1] >> shift);
    const int64_t b_cur = i == M ? (int64_t)nb : (int64_t)(tx[i] >> shift);
    for (int64_t b = b_prev + 1; b <= b_cur; ++b) boff[b] = (uint32_t)i;
}

// K2b + K3 fused: one coalesced pass over the sorted keys writes the bucket table and the occurrence histogram.
// Run lengths come from the warp's run-start ballot; only the last run of each warp may continue into the next
// warp's elements and is finished by galloping.  hist[OCC_BINS] bins + hist[OCC_BINS] = number of keys.
__global__ void __launch_bounds__(256) index_post_kernel(const uint64_t *__restrict__ tx, uint64_t M, int shift, uint32_t nb,
                                                         uint32_t *__restrict__ boff, uint32_t *__restrict__ tk32,
                                                         unsigned long long *hist) {
    __shared__ unsigned int sh[OCC_BINS];
    for (int i = threadIdx.x; i < OCC_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    unsigned int my_keys = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (M + 1 + stride - 1) / stride;
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t i = r * stride + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // warp-contiguous
        const bool in = i < M;
        const uint64_t x = in ? tx[i] : 0;
        // the element after the warp's last one, fetched together with the main load (ends most boundary runs)
        const uint64_t xnext = (lane == 31 && i + 1 < M) ? tx[i + 1] : 0;
        uint64_t xp = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0 && i > 0 && i <= M) xp = tx[i - 1];
        if (in) tk32[i] = (uint32_t)(x >> 8);                     // compact lookup keys (low 32 bits of the hash)
        // bucket table: boff[b] = first i with bucket >= b
        if (i <= M) {
            const int32_t b_prev = i == 0 ? -1 : (int32_t)(xp >> shift);
            const int32_t b_cur = i == M ? (int32_t)nb : (int32_t)(x >> shift);
            for (int32_t b = b_prev + 1; b <= b_cur; ++b) boff[b] = (uint32_t)i;
        }
        // occurrence of every key that STARTS in this warp
        const bool start = in && (i == 0 || (xp >> 8) != (x >> 8));
        const uint32_t S = __ballot_sync(0xffffffffu, start);
        const uint32_t valid = __ballot_sync(0xffffffffu, in);
        const uint64_t key_next = __shfl_sync(0xffffffffu, xnext, 31) >> 8;
        const bool full_warp = valid == 0xffffffffu;
        uint64_t n = 0;
        if (start) {
            const uint32_t above = lane == 31 ? 0u : (S >> (lane + 1)) << (lane + 1);
            if (above) n = (uint32_t)(__ffs(above) - 1) - lane;
            else {
                const uint32_t nvalid = __popc(valid);
                const uint64_t last = i + (nvalid - 1 - lane);         // last element of the warp, same key
                const uint64_t key = x >> 8;
                if (!full_warp || last + 1 >= M || key_next != key) n = last - i + 1;
                else {
                    // the run continues past the prefetched element: finish it by galloping
                    uint64_t lo = last + 1, step = 1;
                    while (lo + step < M && (tx[lo + step] >> 8) == key) { lo += step; step <<= 1; }
                    uint64_t hi = lo + step < M ? lo + step : M;
                    while (hi - lo > 1) {
                        const uint64_t mid = (lo + hi) >> 1;
                        if ((tx[mid] >> 8) == key) lo = mid; else hi = mid;
                    }
                    n = hi - i;
                }
            }
            ++my_keys;
        }
        // most keys occur once or twice: add those bins once per warp instead of 32 same-address atomics
        const uint32_t n1 = __ballot_sync(0xffffffffu, n == 1), n2 = __ballot_sync(0xffffffffu, n == 2);
        if (lane == 0) { if (n1) atomicAdd(&sh[1], (unsigned)__popc(n1)); if (n2) atomicAdd(&sh[2], (unsigned)__popc(n2)); }
        if (n > 2) atomicAdd(&sh[n < OCC_BINS - 1 ? n : OCC_BINS - 1], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OCC_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
    for (int o = 16; o > 0; o >>= 1) my_keys += __shfl_down_sync(0xffffffffu, my_keys, o);
    if (lane == 0 && my_keys) atomicAdd(&hist[OCC_BINS], (unsigned long long)my_keys);
}

// sharded path: the compact lookup keys alone
__global__ void tk32_kernel(const uint64_t *__restrict__ tx, uint64_t M, uint32_t *__restrict__ tk32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) tk32[i] = (uint32_t)(tx[i] >> 8);
}

// Run lengths that fell into the overflow bin, collected for an exact order statistic on the host
// (only reached when the mid_occ rank lands among keys occurring >= OCC_BINS-1 times).
__global__ void occ_overflow_kernel(const uint64_t *__restrict__ tx, uint64_t M, uint32_t *out, unsigned long long *n_out,
                                    uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = tx[i] >> 8;
        if (i > 0 && (tx[i - 1] >> 8) == key) continue;
        if (i + OCC_BINS - 2 >= M || (tx[i + OCC_BINS - 2] >> 8) != key) continue;   // n < OCC_BINS-1
        uint64_t lo = i + OCC_BINS - 2, step = 1;
        while (lo + step < M && (tx[lo + step] >> 8) == key) { lo += step; step <<= 1; }
        uint64_t hi = lo + step < M ? lo + step : M;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if ((tx[mid] >> 8) == key) lo = mid; else hi = mid;
        }
        const unsigned long long slot = atomicAdd(n_out, 1ull);
        if (slot < cap) out[slot] = (uint32_t)(hi - i);
    }
}

struct IndexView {
    const uint64_t *tx;      // sorted by tx>>8
    const uint32_t *tk32;    // low 32 bits of tx>>8: inside a bucket (top hash bits equal, 2k - bits <= 32) they order and
                             // identify keys exactly, at half the bytes per probe
    const uint64_t *ty;
    const uint32_t *boff;
    uint64_t M;
    int shift;               // bucket = tx >> shift  (shift = 8 + 2k - bits)
    const uint32_t *bl;      // sharded index only: bit i set = the run starting at i occurs > mid_occ times GLOBALLY
};

// mm_idx_get, bounded: returns the start of the run of `key` and its length capped at cap+1
// (callers only distinguish n <= mid_occ from n > mid_occ).  n = 0 if absent.
__device__ __forceinline__ uint32_t index_lookup(const IndexView &I, uint64_t key, uint32_t cap, uint32_t *start) {
    const uint64_t b = (key << 8) >> I.shift;
    uint32_t lo = I.boff[b], hi = I.boff[b + 1];
    const uint32_t k32 = (uint32_t)key;
    if (hi - lo <= 8) {
        // small bucket (the usual case, ~6 entries): fetch it with independent loads -- one memory round trip instead
        // of a chain of dependent probes -- and match in registers.  tk32 is allocated with >= 64 entries of slack.
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = I.tk32[lo + j];
        uint32_t n = 0, first = 0;
#pragma unroll
        for (int j = 7; j >= 0; --j)
            if (lo + j < hi && e[j] == k32) { ++n; first = j; }
        *start = lo + first;
        return n > cap + 1 ? cap + 1 : n;
    }
    // lower bound of key in [lo, hi)
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (I.tk32[mid] < k32) lo = mid + 1; else hi = mid;
    }
    *start = lo;
    if (lo >= I.M || lo >= I.boff[b + 1] || I.tk32[lo] != k32) return 0;
    uint32_t n = 1;
    const uint64_t lim0 = (uint64_t)lo + cap + 1 < I.M ? (uint64_t)lo + cap + 1 : I.M;
    const uint64_t lim = lim0 < I.boff[b + 1] ? lim0 : I.boff[b + 1];
    while ((uint64_t)lo + n < lim && I.tk32[lo + n] == k32) ++n;
    return n;
}

// mark the local runs of blacklisted keys (entries equal to ~0 are padding)
__global__ void bl_mark_kernel(const uint64_t *__restrict__ keys, uint64_t n, IndexView I, uint32_t *__restrict__ bitmap) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = keys[i];
    if (key == ~0ull) return;
    uint32_t start;
    if (index_lookup(I, key, 0, &start)) atomicOr(&bitmap[start >> 5], 1u << (start & 31));
}

// ---- sharded index: exact global occurrence counts across target shards ------------------------------------
// Every rank holds the sorted minimizers of ITS targets.  Distinct keys with their local counts are sent to the
// rank owning the key range (owner = key * W >> 2k), summed there, histogrammed (all-reduced -> the same mid_occ
// everywhere); keys above mid_occ come back as a blacklist that each rank marks in a bitmap over its own runs.

__global__ void run_flag_kernel(const uint64_t *__restrict__ tx, uint64_t M, int rshift, uint8_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    flags[i] = (i == 0 || (tx[i] >> rshift) != (tx[i - 1] >> rshift)) ? 1 : 0;
}

// one thread per run: key and run length; starts[U] must hold M
__global__ void rle_emit_kernel(const uint64_t *__restrict__ tx, const uint32_t *__restrict__ starts, uint32_t U,
                                uint64_t *__restrict__ ukey, uint32_t *__restrict__ ucnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    ukey[i] = tx[starts[i]] >> 8;
    ucnt[i] = starts[i + 1] - starts[i];
}

// bounds[d] = first run whose key belongs to owner >= d, d in [0, W]
__global__ void owner_bounds_kernel(const uint64_t *__restrict__ ukey, uint32_t U, int W, int kbits, unsigned long long *bounds) {
    const int d = threadIdx.x;
    if (d > W) return;
    const uint64_t first = d == W ? ~0ull : (((uint64_t)d << kbits) + W - 1) / (uint64_t)W;
    uint32_t lo = 0, hi = U;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (ukey[mid] < first) lo = mid + 1; else hi = mid; }
    bounds[d] = d == W ? U : lo;
}

// owner side: received (key, count) sorted by key; one thread per distinct key sums the <= W partial counts
__global__ void owner_sum_kernel(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ scnt, const uint32_t *__restrict__ starts,
                                 uint32_t G, uint32_t *__restrict__ gcnt, unsigned long long *hist) {
    __shared__ unsigned int sh[OCC_BINS];
    for (int i = threadIdx.x; i < OCC_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < G; g += gridDim.x * blockDim.x) {
        uint64_t n = 0;
        for (uint32_t j = starts[g]; j < starts[g + 1]; ++j) n += scnt[j];
        gcnt[g] = n > 0xffffffffull ? 0xffffffffu : (uint32_t)n;
        atomicAdd(&sh[n < OCC_BINS - 1 ? n : OCC_BINS - 1], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OCC_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// owner side: keys occurring more than mid_occ times -> blacklist (unordered), and overflow-bin values if asked
__global__ void owner_select_kernel(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ gcnt,
                                    uint32_t G, uint32_t thr, uint64_t *__restrict__ out, unsigned long long *n_out, uint64_t cap) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G || gcnt[g] <= thr) return;
    const unsigned long long slot = atomicAdd(n_out, 1ull);
    if (slot < cap) out[slot] = skey[starts[g]];
}
__global__ void owner_overflow_kernel(const uint32_t *__restrict__ gcnt, uint32_t G, uint32_t *__restrict__ out, unsigned long long *n_out, uint64_t cap) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G || gcnt[g] < OCC_BINS - 1) return;
    const unsigned long long slot = atomicAdd(n_out, 1ull);
    if (slot < cap) out[slot] = gcnt[g];
}

}  // namespace lrge
