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
// XT = uint64_t: tx holds hash << 8 | span (KSH = 8), the compact keys tk32 are written here.
// XT = uint32_t: tx IS the compact key array (sorted 32-bit hashes, KSH = 0), nothing to write.
template <typename XT, int KSH>
__global__ void __launch_bounds__(256) index_post_kernel(const XT *__restrict__ tx, uint64_t M, int shift, uint32_t nb,
                                                         uint32_t *__restrict__ boff, uint32_t *__restrict__ tk32,
                                                         unsigned long long *hist, uint32_t *__restrict__ run_n = nullptr) {
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
        const uint64_t x = in ? (uint64_t)tx[i] : 0;
        // the element after the warp's last one, fetched together with the main load (ends most boundary runs)
        const uint64_t xnext = (lane == 31 && i + 1 < M) ? (uint64_t)tx[i + 1] : 0;
        uint64_t xp = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0 && i > 0 && i <= M) xp = (uint64_t)tx[i - 1];
        if (KSH && in) tk32[i] = (uint32_t)(x >> KSH);            // compact lookup keys (low 32 bits of the hash)
        // bucket table: boff[b] = first i with bucket >= b
        if (i <= M) {
            const int32_t b_prev = i == 0 ? -1 : (int32_t)(xp >> shift);
            const int32_t b_cur = i == M ? (int32_t)nb : (int32_t)(x >> shift);
            for (int32_t b = b_prev + 1; b <= b_cur; ++b) boff[b] = (uint32_t)i;
        }
        // occurrence of every key that STARTS in this warp
        const bool start = in && (i == 0 || (xp >> KSH) != (x >> KSH));
        const uint32_t S = __ballot_sync(0xffffffffu, start);
        const uint32_t valid = __ballot_sync(0xffffffffu, in);
        const uint64_t key_next = __shfl_sync(0xffffffffu, xnext, 31) >> KSH;
        const bool full_warp = valid == 0xffffffffu;
        uint64_t n = 0;
        if (start) {
            const uint32_t above = lane == 31 ? 0u : (S >> (lane + 1)) << (lane + 1);
            if (above) n = (uint32_t)(__ffs(above) - 1) - lane;
            else {
                const uint32_t nvalid = __popc(valid);
                const uint64_t last = i + (nvalid - 1 - lane);         // last element of the warp, same key
                const uint64_t key = x >> KSH;
                if (!full_warp || last + 1 >= M || key_next != key) n = last - i + 1;
                else {
                    // the run continues past the prefetched element: finish it by galloping
                    uint64_t lo = last + 1, step = 1;
                    while (lo + step < M && ((uint64_t)tx[lo + step] >> KSH) == key) { lo += step; step <<= 1; }
                    uint64_t hi = lo + step < M ? lo + step : M;
                    while (hi - lo > 1) {
                        const uint64_t mid = (lo + hi) >> 1;
                        if (((uint64_t)tx[mid] >> KSH) == key) lo = mid; else hi = mid;
                    }
                    n = hi - i;
                }
            }
            ++my_keys;
        }
        // sharded index: run length at the first element of every run (saturated like the packed entries), 0 elsewhere
        if (run_n && in) run_n[i] = (uint32_t)(n < ((1ull << 20) - 1) ? n : ((1ull << 20) - 1));
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

// 32-bit hashes of the (read, position)-ordered minimizers: the sort keys of the single-GPU index for 2k <= 32
__global__ void key32_kernel(const uint64_t *__restrict__ x, uint64_t M, uint32_t *__restrict__ k32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) k32[i] = (uint32_t)(x[i] >> 8);
}

// sharded path: the compact lookup keys alone
__global__ void tk32_kernel(const uint64_t *__restrict__ tx, uint64_t M, uint32_t *__restrict__ tk32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) tk32[i] = (uint32_t)(tx[i] >> 8);
}

// Run lengths that fell into the overflow bin, collected for an exact order statistic on the host
// (only reached when the mid_occ rank lands among keys occurring >= OCC_BINS-1 times).
template <typename XT, int KSH>
__global__ void occ_overflow_kernel(const XT *__restrict__ tx, uint64_t M, uint32_t *out, unsigned long long *n_out,
                                    uint64_t cap) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = (uint64_t)tx[i] >> KSH;
        if (i > 0 && ((uint64_t)tx[i - 1] >> KSH) == key) continue;
        if (i + OCC_BINS - 2 >= M || ((uint64_t)tx[i + OCC_BINS - 2] >> KSH) != key) continue;   // n < OCC_BINS-1
        uint64_t lo = i + OCC_BINS - 2, step = 1;
        while (lo + step < M && ((uint64_t)tx[lo + step] >> KSH) == key) { lo += step; step <<= 1; }
        uint64_t hi = lo + step < M ? lo + step : M;
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (((uint64_t)tx[mid] >> KSH) == key) lo = mid; else hi = mid;
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


// ---- sharded index, fast path: packed (key << 24 | count) entries, sizes kept on the device --------------------------
// A rank's count saturates at 2^20 - 1 (> max_mid_occ = 10^6, so "count > mid_occ" and the occurrence histogram, whose
// bins saturate at OCC_BINS - 1, are unaffected); the sum over <= 16 ranks stays inside the 24-bit field.
constexpr int PK_SHIFT = 24;
constexpr uint64_t PK_CMAX = (1ull << 20) - 1;
constexpr uint64_t PK_FIELD = (1ull << PK_SHIFT) - 1;
constexpr uint64_t PK_SENTINEL = ~0ull;

// element i of the sorted keys as a packed entry (key << 24 | run length): the input of the DeviceSelect that keeps the
// run starts (run_n != 0) -- distinct keys and their counts in one pass, sizes stay on the device
template <typename XT, int KSH>
struct PackRunOp {
    const XT *tx; const uint32_t *run_n;
    __host__ __device__ __forceinline__ uint64_t operator()(uint32_t i) const {
        return ((uint64_t)tx[i] >> KSH) << 24 | (uint64_t)run_n[i];
    }
};

// one thread per run: packed entry; U and the run starts come from DeviceSelect (device side), the last run ends at M
__global__ void rle_pack_kernel(const uint64_t *__restrict__ tx, const uint32_t *__restrict__ starts, const uint32_t *__restrict__ U_dev,
                                uint64_t M, uint64_t *__restrict__ out) {
    const uint32_t U = *U_dev;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    const uint64_t s0 = starts[i], s1 = i + 1 < U ? (uint64_t)starts[i + 1] : M;
    const uint64_t n = s1 - s0;
    out[i] = (tx[s0] >> 8) << PK_SHIFT | (n < PK_CMAX ? n : PK_CMAX);
}

// bounds[d] = first packed entry whose key belongs to owner >= d (d in [0, W]); sendc[d] = entries for owner d
__global__ void owner_bounds_dev_kernel(const uint64_t *__restrict__ pk, const uint32_t *__restrict__ U_dev, int W, int kbits,
                                        unsigned long long *bounds, unsigned long long *sendc) {
    __shared__ unsigned long long sb[64];
    const uint32_t U = *U_dev;
    const int d = threadIdx.x;
    if (d <= W) {
        const uint64_t first = d == W ? ~0ull : (((uint64_t)d << kbits) + W - 1) / (uint64_t)W;
        uint32_t lo = 0, hi = U;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((pk[mid] >> PK_SHIFT) < first) lo = mid + 1; else hi = mid; }
        sb[d] = d == W ? U : lo;
        bounds[d] = sb[d];
    }
    __syncthreads();
    if (d < W) sendc[d] = sb[d + 1] - sb[d];
}

// owner side: the received packed entries are W runs, each sorted by key, holding every rank's distinct keys of this
// owner's key range (<= W entries per key).  The range is cut into slices of equal key width; a CTA binary-searches its
// slice in every run, pulls the W sub-runs (contiguous, coalesced) into a SHARED-MEMORY open-addressing table of packed
// words (claim an empty slot with one 64-bit CAS; a slot already holding the key takes the count with one atomicAdd on
// the same word) and reads the summed counts back: streaming HBM reads, no sort, no global atomics per entry.
// A slice that does not fit the table is halved until it does (a single key never exceeds W entries).
// MODE 0: occurrence histogram of the global counts (hist[OCC_BINS] counts the keys).
// MODE 1: keys whose global count exceeds mid_occ (meta[0]) -> blacklist (unordered); meta[2] set if it does not fit cap.
constexpr int OM_THREADS = 256;
constexpr int OM_SLOTS = 4096;             // 32 KB of shared memory: several CTAs per SM hide the load latency
constexpr int OM_HBINS = 64;               // counts below this go through a shared-memory histogram, the rest straight to global
constexpr int OM_CAP = OM_SLOTS / 2;
constexpr int OM_MAXW = 16;

struct OwnerMergeArgs {
    const uint64_t *rk;                    // received runs, run w = rk[roff[w] .. roff[w + 1])
    uint64_t roff[OM_MAXW + 1];
    int W;
    uint64_t klo, khi;                     // this owner's key range
    uint64_t slice_w;                      // keys per CTA slice
    const uint32_t *cuts;                  // [run][slice + 1]: first entry of run w (relative to roff[w]) with key >= slice start
    uint32_t n_slices;
    unsigned long long *hist;              // MODE 0
    const unsigned long long *meta;        // MODE 1: meta[0] = mid_occ
    uint64_t *bl; unsigned long long *bl_n; uint64_t bl_cap; unsigned long long *bl_ovf;
    // MODE 0 with candidates: summed entries above cand_thr (a lower bound of mid_occ) are kept, so that the blacklist
    // is a filter over them instead of a second merge
    uint64_t *cand; unsigned long long *cand_n; uint64_t cand_thr;
};

// cuts[w][p] for every run and slice boundary: one thread per binary search
__global__ void owner_cuts_kernel(OwnerMergeArgs A, uint32_t *__restrict__ cuts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per = (uint64_t)A.n_slices + 1;
    if (i >= per * A.W) return;
    const int w = (int)(i / per);
    const uint64_t p = i % per;
    const uint64_t key = min(A.klo + p * A.slice_w, A.khi);
    uint64_t lo = A.roff[w], hi = A.roff[w + 1];
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((A.rk[mid] >> PK_SHIFT) < key) lo = mid + 1; else hi = mid; }
    cuts[i] = (uint32_t)(lo - A.roff[w]);
}

template <int MODE>
__global__ void __launch_bounds__(OM_THREADS) owner_merge_kernel(OwnerMergeArgs A) {
    __shared__ __align__(16) unsigned long long om_table[OM_SLOTS];
    __shared__ unsigned int sh[OM_HBINS];
    __shared__ uint64_t s_lo[OM_MAXW], s_hi[OM_MAXW];
    __shared__ uint32_t s_pre[OM_MAXW + 1];
    __shared__ uint32_t s_cw[OM_THREADS / 32];
    __shared__ unsigned long long s_gbase;
    const uint32_t tid = threadIdx.x;
    if (MODE == 0) for (int i = tid; i < OM_HBINS; i += OM_THREADS) sh[i] = 0;
    unsigned int my_keys = 0;
    const uint64_t slice_lo = A.klo + (uint64_t)blockIdx.x * A.slice_w;
    const uint64_t slice_hi = min(slice_lo + A.slice_w, A.khi);
    auto lower = [&](int w, uint64_t key) -> uint64_t {                // first entry of run w with key >= `key`
        uint64_t lo = A.roff[w], hi = A.roff[w + 1];
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if ((A.rk[mid] >> PK_SHIFT) < key) lo = mid + 1; else hi = mid; }
        return lo;
    };
    uint64_t a = slice_lo;
    while (a < slice_hi) {
        uint64_t b = slice_hi;
        uint32_t n = 0;
        for (;;) {
            __syncthreads();
            if (tid < (uint32_t)A.W) {
                const uint32_t *cw = A.cuts + (uint64_t)tid * (A.n_slices + 1) + blockIdx.x;
                s_lo[tid] = a == slice_lo ? A.roff[tid] + cw[0] : lower((int)tid, a);        // slice boundaries are precomputed
                s_hi[tid] = b == slice_hi ? A.roff[tid] + cw[1] : lower((int)tid, b);
            }
            __syncthreads();
            if (tid == 0) { uint32_t run = 0; for (int w = 0; w < A.W; ++w) { s_pre[w] = run; run += (uint32_t)min(s_hi[w] - s_lo[w], (uint64_t)0x7fffffff); } s_pre[A.W] = run; }
            __syncthreads();
            n = s_pre[A.W];
            if (n <= (uint32_t)OM_CAP || b - a <= 1) break;
            b = a + (b - a) / 2;
        }
        for (int i = tid; i < OM_SLOTS; i += OM_THREADS) om_table[i] = 0ull;
        __syncthreads();
        for (uint32_t j = tid; j < n; j += OM_THREADS) {
            int w = 0;
            while (w + 1 < A.W && j >= s_pre[w + 1]) ++w;
            const unsigned long long e = A.rk[s_lo[w] + (j - s_pre[w])];
            const uint64_t key = e >> PK_SHIFT;
            uint32_t s = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 40) & (OM_SLOTS - 1);
            for (;;) {
                unsigned long long cur = om_table[s];
                if (cur == 0ull) { cur = atomicCAS(&om_table[s], 0ull, e); if (cur == 0ull) break; }
                if ((cur >> PK_SHIFT) == key) { atomicAdd(&om_table[s], e & PK_FIELD); break; }
                s = (s + 1) & (OM_SLOTS - 1);
            }
        }
        __syncthreads();
        uint32_t my_cand = 0;
        for (int i = tid; i < OM_SLOTS; i += OM_THREADS) {
            const unsigned long long e = om_table[i];
            if (e == 0ull) continue;
            const uint64_t g = e & PK_FIELD;
            if (MODE == 0) {
                if (g < OM_HBINS) atomicAdd(&sh[g], 1u);
                else atomicAdd(&A.hist[g < OCC_BINS - 1 ? g : OCC_BINS - 1], 1ull);
                ++my_keys;
                if (g > A.cand_thr) ++my_cand;
            }
            else if (g > A.meta[0]) {
                const unsigned long long slot = atomicAdd(A.bl_n, 1ull);
                if (slot < A.bl_cap) A.bl[slot] = e >> PK_SHIFT; else *A.bl_ovf = 1;
            }
        }
        if (MODE == 0 && A.cand) {
            // candidates of this table: one global atomic per CTA step (<= one candidate per received entry: cand always fits)
            uint32_t incl = my_cand;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)(tid & 31) >= o) incl += u; }
            if ((tid & 31) == 31) s_cw[tid >> 5] = incl;
            __syncthreads();
            if (tid == 0) {
                uint32_t run = 0;
                for (int w = 0; w < OM_THREADS / 32; ++w) { const uint32_t t = s_cw[w]; s_cw[w] = run; run += t; }
                s_gbase = run ? atomicAdd(A.cand_n, (unsigned long long)run) : 0ull;
            }
            __syncthreads();
            if (my_cand) {
                unsigned long long pos = s_gbase + s_cw[tid >> 5] + incl - my_cand;
                for (int i = tid; i < OM_SLOTS; i += OM_THREADS) {
                    const unsigned long long e = om_table[i];
                    if (e != 0ull && (e & PK_FIELD) > A.cand_thr) A.cand[pos++] = e;
                }
            }
        }
        a = b;
    }
    if (MODE == 0) {
        __syncthreads();
        for (int i = tid; i < OM_HBINS; i += OM_THREADS)
            if (sh[i]) atomicAdd(&A.hist[i], (unsigned long long)sh[i]);
        for (int o = 16; o > 0; o >>= 1) my_keys += __shfl_down_sync(0xffffffffu, my_keys, o);
        if ((tid & 31) == 0 && my_keys) atomicAdd(&A.hist[OCC_BINS], (unsigned long long)my_keys);
    }
}

// ---- key exchange over peer memory (NVLink / NVSwitch) --------------------------------------------------------------
// Every rank exports a receive window and a small array of landing flags (CUDA IPC); a sender stores its slice of packed
// entries for rank d straight into d's window at the offset the send-count matrix gives it, and then raises
// flags_d[sender] = epoch.  Replaces ncclSend / ncclRecv (point-to-point NCCL moved ~100 GB/s per peer here).
struct PeerPushArgs {
    const uint64_t *src;                      // this rank's packed entries, in key order
    uint64_t *dst[OM_MAXW];                   // window of every rank (dst[me] = the local one)
    unsigned long long *flag[OM_MAXW];        // &flags_d[me] in rank d's flag array
    uint64_t sbeg[OM_MAXW + 1];               // slice of src for rank d: [sbeg[d], sbeg[d + 1])
    uint64_t doff[OM_MAXW];                   // where this rank's slice starts in d's window (entries)
    unsigned int *done;                       // [OM_MAXW] CTAs finished per destination (self-resetting)
    unsigned long long epoch;
    int W, me;
    uint32_t ctas_per_dest;
};

__global__ void __launch_bounds__(256) peer_push_kernel(PeerPushArgs A) {
    const int d = blockIdx.x / A.ctas_per_dest;
    const uint32_t j = blockIdx.x % A.ctas_per_dest;
    const uint64_t n = A.sbeg[d + 1] - A.sbeg[d];
    const uint64_t per = (n + A.ctas_per_dest - 1) / A.ctas_per_dest;
    const uint64_t lo = min(n, per * j), hi = min(n, lo + per);
    const uint64_t *s = A.src + A.sbeg[d];
    uint64_t *t = A.dst[d] + A.doff[d];
    for (uint64_t i = lo + threadIdx.x; i < hi; i += 4ull * blockDim.x) {      // four independent 8-byte copies per thread
        uint64_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint64_t k = i + (uint64_t)u * blockDim.x; v[u] = k < hi ? s[k] : 0; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint64_t k = i + (uint64_t)u * blockDim.x; if (k < hi) t[k] = v[u]; }
    }
    __threadfence_system();                                                     // this CTA's stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&A.done[d], 1u);                    // ... before it counts as finished
        if (prev == A.ctas_per_dest - 1) {                                      // last CTA of this destination: raise the flag
            A.done[d] = 0;
            __threadfence_system();
            if (d != A.me) *reinterpret_cast<volatile unsigned long long *>(A.flag[d]) = A.epoch;
        }
    }
}

// wait until every peer's slice of this epoch has landed in the local window; gives up after ~4 s (err = 1)
__global__ void peer_wait_kernel(const unsigned long long *flags, int W, int me, unsigned long long epoch, unsigned int *err) {
    const int s = threadIdx.x;
    if (s >= W || s == me) return;
    const volatile unsigned long long *f = flags + s;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < epoch) {
        __nanosleep(200);
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { *err = 1; break; }
    }
    __threadfence_system();
}

// blacklist = candidates whose global count exceeds mid_occ (meta[0])
__global__ void bl_from_cand_kernel(const uint64_t *__restrict__ cand, const unsigned long long *__restrict__ cand_n,
                                    const unsigned long long *__restrict__ meta, uint64_t *__restrict__ bl, unsigned long long *bl_n,
                                    uint64_t bl_cap, unsigned long long *bl_ovf) {
    const unsigned long long n = *cand_n, thr = meta[0];
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint64_t e = cand[i];
        if ((e & PK_FIELD) <= thr) continue;
        const unsigned long long slot = atomicAdd(bl_n, 1ull);
        if (slot < bl_cap) bl[slot] = e >> PK_SHIFT; else *bl_ovf = 1;
    }
}

// mm_idx_cal_max_occ + mm_mapopt_update on the (all-reduced) histogram, on the device: meta[0] = mid_occ,
// meta[1] = 1 if the rank lies in the overflow bin (the exact host path must run), meta[2] is the blacklist overflow flag.
__global__ void mid_occ_kernel(const unsigned long long *__restrict__ hist, float mid_occ_frac, int min_mid_occ, int max_mid_occ,
                               unsigned long long *meta) {
    if (threadIdx.x || blockIdx.x) return;
    const unsigned long long n_keys = hist[OCC_BINS];
    long long mid = 0x7fffffff;
    unsigned long long need_exact = 0;
    if (n_keys) {
        const unsigned long long rank = (unsigned int)((1. - (double)mid_occ_frac) * (double)n_keys);
        unsigned long long cum = 0;
        long long val = -1;
        for (int b = 0; b < OCC_BINS - 1; ++b) { cum += hist[b]; if (cum > rank) { val = b; break; } }
        if (val < 0) need_exact = 1;
        mid = val + 1;
        if (mid < min_mid_occ) mid = min_mid_occ;
        if (max_mid_occ > min_mid_occ && mid > max_mid_occ) mid = max_mid_occ;
    }
    meta[0] = (unsigned long long)mid;
    meta[1] = need_exact;
}

}  // namespace lrge
