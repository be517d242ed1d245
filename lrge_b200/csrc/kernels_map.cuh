// kernels_map.cuh -- K4a seed lookup and K5: per-query shared-memory sort of the anchors.
//
// Replaces minimap2's mm_collect_matches lookup (seed.c) and radix_sort_128x in collect_seed_hits (map.c), SURVEY.md rows
// a8-a9, for ALL queries of a batch:
//
//   seed_lookup_kernel   every query minimizer: bucket table -> compact keys -> (start, n) of its hit list
//   (anchor_fill_kernel, kernels_seed.cuh: hit lists -> anchors of each query, contiguous, in generation order)
//   sort_query_kernel    persistent CTAs pull queries (longest first) from a work counter; per query
//     1. stage   its anchors HBM -> shared memory (coalesced), as (sort key, payload)
//     2. sort    block-wide stable LSD radix sort in smem on the (target, strand) bits: groups become contiguous and keep
//                their generation order
//     3. groups  one pass writes the pair / group boundaries of the sorted array as bit maps (read-only from here on);
//                warps claim 32-anchor windows of the sorted array, take the (query, target) pairs that start in them and,
//                for every strand group that can hold a chain, bring it into target-position order (already ascending /
//                strictly descending for a colinear overlap; warp-level radix sort otherwise), write it back as compact
//                anchors and append a group descriptor for the chaining kernel (chain_group_kernel, kernels_chain.cuh)
//   A query whose anchors exceed the smem capacity is sorted in ROUNDS over disjoint target-id ranges (groups never span
//   targets, so rounds are independent); ranges come from a 2048-bin histogram of its anchors over target ids.
//   A query that cannot be planned (one bin alone above capacity, > 255 rounds) is appended to `deferred` and the host
//   sorts it with the device-wide radix sort of the global path.
//
// Stable sort + generation order = the tie order of the global path (ksort's radix sort is
// unstable; ties are equal (target, strand, position) anchors from repeated query k-mers).
#pragma once
#include "common.cuh"
#include "kernels_index.cuh"
#include "kernels_seed.cuh"
#include "kernels_chain.cuh"

namespace lrge {

constexpr int MAP_MAX_WARPS = 32;
constexpr int MAP_BINS = 2048;         // target-id histogram bins for round planning (fewer when the smem capacity is tiny)
constexpr int MAP_MAX_ROUNDS = 255;
constexpr int SL_ILP = 4;              // independent probes per thread in seed_lookup_kernel
constexpr int SL_THREADS = 256;

// ---- K4a: index lookup of every query minimizer ----------------------------------------------------------------
// mm_collect_matches (seed.c): n > mid_occ => the seed is dropped (hn = 0).  SL_ILP probes per thread are issued
// back to back (bucket bounds first, then the compact keys of all buckets) so that a thread keeps 4 x 9 independent
// loads in flight; the table (boff) is L2-resident, the compact keys are one or two 32-byte sectors per probe.
__global__ void __launch_bounds__(SL_THREADS) seed_lookup_kernel(const uint64_t *__restrict__ qx, const uint8_t *__restrict__ flt,
                                                                 uint64_t MQ, IndexView I, int mid_occ,
                                                                 uint32_t *__restrict__ hstart, uint32_t *__restrict__ hn,
                                                                 uint32_t *__restrict__ acnt,
                                                                 unsigned long long *n_seeds, unsigned long long *n_hits) {
    const uint64_t base = (uint64_t)blockIdx.x * (SL_THREADS * SL_ILP) + threadIdx.x;
    uint64_t key[SL_ILP];
    uint32_t lo[SL_ILP], hi[SL_ILP];
    bool act[SL_ILP];
#pragma unroll
    for (int u = 0; u < SL_ILP; ++u) {
        const uint64_t i = base + (uint64_t)u * SL_THREADS;
        act[u] = i < MQ && !flt[i];
        key[u] = act[u] ? qx[i] >> 8 : 0;
    }
#pragma unroll
    for (int u = 0; u < SL_ILP; ++u) {
        const uint64_t b = (key[u] << 8) >> I.shift;
        lo[u] = act[u] ? I.boff[b] : 0;
        hi[u] = act[u] ? I.boff[b + 1] : 0;
    }
    uint32_t e[SL_ILP][8];
#pragma unroll
    for (int u = 0; u < SL_ILP; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) e[u][j] = I.tk32[lo[u] + j];     // tk32 has >= 64 entries of slack
    unsigned int seeds = 0, hits = 0;
#pragma unroll
    for (int u = 0; u < SL_ILP; ++u) {
        const uint64_t i = base + (uint64_t)u * SL_THREADS;
        if (i >= MQ) continue;
        uint32_t n = 0, start = 0;
        if (act[u]) {
            const uint32_t k32 = (uint32_t)key[u];
            if (hi[u] - lo[u] <= 8) {
                uint32_t first = 0;
#pragma unroll
                for (int j = 7; j >= 0; --j)
                    if (lo[u] + j < hi[u] && e[u][j] == k32) { ++n; first = j; }
                start = lo[u] + first;
            } else {
                n = index_lookup(I, key[u], (uint32_t)mid_occ, &start);       // large bucket: bisect
            }
            bool rep = n > (uint32_t)mid_occ;                                   // m[i].flt = 1
            if (n && I.bl && ((I.bl[start >> 5] >> (start & 31)) & 1)) rep = true;   // > mid_occ over all target shards
            if (rep) { n = 0; start = SEED_REPETITIVE; }                        // present in the index, dropped: counts for rl
        }
        hstart[i] = start;
        hn[i] = n;
        acnt[i] = n;                                                            // no skip_seed condition can apply (host decides)
        seeds += n ? 1 : 0;
        hits += n;
    }
    for (int o = 16; o > 0; o >>= 1) { seeds += __shfl_down_sync(0xffffffffu, seeds, o); hits += __shfl_down_sync(0xffffffffu, hits, o); }
    if ((threadIdx.x & 31) == 0 && hits) { atomicAdd(n_seeds, (unsigned long long)seeds); atomicAdd(n_hits, (unsigned long long)hits); }
}

// ---- K5: per-query shared-memory sort ---------------------------------------------------------------------------
struct SortArgs {
    const void *akey;                // anchors of the batch, per query contiguous, in generation order (anchor_fill_kernel):
    const uint32_t *aval;            //   key (KT) = ((rid - rid_base) << 1 | rev) << PB | target pos, val = span << 24 | qpos'
    void *akey2;                     // scratch of the same size: a multi-round query's anchors partitioned by round
    uint32_t *aval2;
    const uint64_t *qaoff;           // per query (absolute id): anchors before it
    uint64_t a_base;                 // qaoff of the batch's first query: akey[0] is anchor a_base
    const uint32_t *qorder;          // queries of the batch, longest first
    uint32_t nq;
    uint32_t *work_ctr;
    uint32_t cap;                    // anchors per round (smem capacity), multiple of 32, < 65536
    int PB, key_bits;                // key_bits = bits of (rid - rid_base) + 1 + PB
    uint32_t rid_base;               // first target id held by this index (sharded index: the rank's block)
    int rel_bits;                    // bits of (rid - rid_base); histogram bin = (rid - rid_base) >> (rel_bits - log2(bins))
    int max_span, min_cnt, min_sc;   // a group of n anchors scores at most n * max_span
    uint2 *anc;                      // out: sorted compact anchors (target pos, payload), same offsets as akey
    GroupLists L;                    // out: descriptors of the groups that can hold a chain, 4 size classes
    uint32_t *gcount;                //      entries per class
    uint32_t *pair_ctr;              //      pair slots handed out
    uint32_t *deferred, *n_deferred; // queries left to the global path
    unsigned long long *stats;       // [0] anchors [1] (query, target, strand) groups [2] rounds [3] capacity violations
    unsigned long long *cyc;         // optional: SM cycles per phase summed over CTAs: [0] plan [1] stage [2] sort [3] groups [4] total
};

// exclusive scan of one value per thread over the CTA (NT threads); total = sum.  wsum: NT / 32 + 1 words of smem.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *wsum, uint32_t &total) {
    constexpr int MAP_WARPS = NT / 32;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const uint32_t x = lane < MAP_WARPS ? wsum[lane] : 0;
        uint32_t xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, xi, o);
            if ((int)lane >= o) xi += u;
        }
        if (lane < MAP_WARPS) wsum[lane] = xi - x;
        if (lane == MAP_WARPS - 1) wsum[MAP_WARPS] = xi;
    }
    __syncthreads();
    total = wsum[MAP_WARPS];
    const uint32_t r = wsum[warp] + incl - v;
    __syncthreads();
    return r;
}

// Stable LSD radix sort of N (key, payload) records in shared memory on key bits [lo_bit, lo_bit + n_bits), in
// ceil(n_bits / 8) passes of <= 8 bits.  Every warp owns a contiguous slice; per pass: (1) rank of each record among the
// records of its slice with the same digit (peers by __match_any_sync, no atomics) plus per-(warp, digit) counts,
// (2) column scan of the counts in (digit, warp) order, (3) scatter.  rank[] is N words of scratch.
// Returns 1 if the result is in (kB, vB).
template <typename KT, int NT>
__device__ int block_radix_sort(KT *kA, uint32_t *vA, KT *kB, uint32_t *vB, uint32_t N, int lo_bit, int n_bits,
                                uint16_t *hist, uint16_t *rank, uint32_t *wsum) {
    constexpr int MAP_THREADS = NT, MAP_WARPS = NT / 32;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int passes = n_bits > 8 ? (n_bits + 7) / 8 : 1;
    const int dbits = (n_bits + passes - 1) / passes;
    const uint32_t per_warp = ((N + MAP_THREADS - 1) / MAP_THREADS) * 32;
    const uint32_t lo = min(warp * per_warp, N), hi = min(lo + per_warp, N);
    const uint32_t lt = (1u << lane) - 1;
    uint16_t *myh = hist + warp * 256;
    KT *ks = kA, *kd = kB;
    uint32_t *vs = vA, *vd = vB;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = lo_bit + pass * dbits;
        const int nb = min(dbits, n_bits - pass * dbits);
        const uint32_t dmask = (1u << nb) - 1;
        for (uint32_t i = tid; i < MAP_WARPS * 128; i += MAP_THREADS) reinterpret_cast<uint32_t *>(hist)[i] = 0;
        __syncthreads();
        for (uint32_t b = lo; b < hi; b += 32) {
            const uint32_t e = b + lane;
            const bool valid = e < hi;
            const uint32_t d = valid ? (uint32_t)(ks[e] >> shift) & dmask : 0;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);   // lanes holding the same digit
            if (valid) rank[e] = (uint16_t)(myh[d] + __popc(peers & lt));
            __syncwarp();
            if (valid && (peers & lt) == 0) myh[d] = (uint16_t)(myh[d] + __popc(peers));     // lowest peer updates the count
            __syncwarp();
        }
        __syncthreads();
        uint32_t tot = 0;
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < MAP_WARPS; ++w) tot += hist[w * 256 + tid];
        }
        uint32_t total;
        uint32_t run = block_excl_scan<NT>(tot, wsum, total);
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < MAP_WARPS; ++w) {
                const uint32_t c = hist[w * 256 + tid];
                hist[w * 256 + tid] = (uint16_t)run;
                run += c;
            }
        }
        __syncthreads();
        for (uint32_t e = lo + lane; e < hi; e += 32) {
            const KT k = ks[e];
            const uint32_t pos = (uint32_t)myh[(uint32_t)(k >> shift) & dmask] + rank[e];
            kd[pos] = k;
            vd[pos] = vs[e];
        }
        __syncthreads();
        KT *tk = ks; ks = kd; kd = tk;
        uint32_t *tv = vs; vs = vd; vd = tv;
    }
    return passes & 1;
}

// One warp brings the n anchors of a (target, strand) group -- contiguous in (K, V), in generation order after the
// block sort on (target, strand) -- into ascending target-position order, stably.  Anchors of a colinear overlap are
// generated in ascending query position, i.e. already ascending (+ strand) or strictly descending (- strand) in target
// position: those cases cost one read.  Anything else (repeats) is sorted by a warp-level LSD radix sort on the PB
// position bits, with scr (8 bytes per anchor) as the second buffer, myh (256 counters) and rank (n words) as scratch.
template <typename KT>
__device__ void warp_sort_group(KT *K, uint32_t *V, int32_t n, int PB, uint2 *scr, uint16_t *myh, uint16_t *rank, uint32_t lane) {
    const KT posmask = (KT)(((KT)1 << PB) - 1);
    bool asc = true, desc = true;
    for (int32_t b = 0; b + 1 < n; b += 32) {
        const int32_t i = b + (int32_t)lane;
        if (i + 1 < n) {
            const uint32_t a = (uint32_t)(K[i] & posmask), c = (uint32_t)(K[i + 1] & posmask);
            if (a > c) asc = false;
            if (a <= c) desc = false;
        }
    }
    asc = __all_sync(0xffffffffu, asc);
    desc = __all_sync(0xffffffffu, desc);
    if (asc) return;
    if (desc) {                                                         // all positions distinct: the reversal is THE sorted order
        for (int32_t i = lane; i < n / 2; i += 32) {
            const KT ka = K[i], kb = K[n - 1 - i];
            const uint32_t va = V[i], vb = V[n - 1 - i];
            K[i] = kb; K[n - 1 - i] = ka; V[i] = vb; V[n - 1 - i] = va;
        }
        __syncwarp();
        return;
    }
    const KT high = K[0] & ~posmask;
    const uint32_t lt = (1u << lane) - 1;
    const int passes = (PB + 7) / 8;
    for (int pass = 0; pass < passes; ++pass) {
        const int shift = pass * 8;
        const bool from_k = (pass & 1) == 0;                             // even passes read (K, V), odd passes read scr
        for (uint32_t i = lane; i < 128; i += 32) reinterpret_cast<uint32_t *>(myh)[i] = 0;
        __syncwarp();
        for (int32_t b = 0; b < n; b += 32) {
            const int32_t e = b + (int32_t)lane;
            const bool valid = e < n;
            uint32_t pos = 0;
            if (valid) pos = from_k ? (uint32_t)(K[e] & posmask) : scr[e].x;
            const uint32_t d = (pos >> shift) & 255u;
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 0xffffffffu);
            if (valid) rank[e] = (uint16_t)(myh[d] + __popc(peers & lt));
            __syncwarp();
            if (valid && (peers & lt) == 0) myh[d] = (uint16_t)(myh[d] + __popc(peers));
            __syncwarp();
        }
        // exclusive scan of the 256 counters: 8 per lane
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { c[j] = myh[lane * 8 + j]; sum += c[j]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
        uint32_t run = incl - sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) { myh[lane * 8 + j] = (uint16_t)run; run += c[j]; }
        __syncwarp();
        for (int32_t e = lane; e < n; e += 32) {
            uint32_t pos, val;
            if (from_k) { pos = (uint32_t)(K[e] & posmask); val = V[e]; } else { pos = scr[e].x; val = scr[e].y; }
            const uint32_t at = (uint32_t)myh[(pos >> shift) & 255u] + rank[e];
            if (from_k) scr[at] = make_uint2(pos, val); else { K[at] = high | (KT)pos; V[at] = val; }
        }
        __syncwarp();
    }
    if (passes & 1) {                                                    // result sits in scr
        for (int32_t e = lane; e < n; e += 32) { const uint2 r = scr[e]; K[e] = high | (KT)r.x; V[e] = r.y; }
        __syncwarp();
    }
}

template <typename KT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) sort_query_kernel(SortArgs A) {
    constexpr int MAP_THREADS = NT, MAP_WARPS = NT / 32;
    extern __shared__ __align__(16) uint8_t map_smem[];
    __shared__ uint32_t s_wsum[MAP_WARPS + 1];
    __shared__ uint32_t s_rbound[MAP_MAX_ROUNDS + 2], s_rE[MAP_MAX_ROUNDS + 2], s_rN[MAP_MAX_ROUNDS + 2];
    __shared__ uint32_t s_q, s_nrounds, s_unit;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cap = A.cap;
    KT *kX = reinterpret_cast<KT *>(map_smem);
    uint32_t *vX = reinterpret_cast<uint32_t *>(map_smem + (size_t)cap * sizeof(KT));
    uint8_t *bufY = map_smem + (size_t)cap * (sizeof(KT) + 4);
    KT *kY = reinterpret_cast<KT *>(bufY);
    uint32_t *vY = reinterpret_cast<uint32_t *>(bufY + (size_t)cap * sizeof(KT));
    uint16_t *rank = reinterpret_cast<uint16_t *>(bufY + (size_t)cap * (sizeof(KT) + 4));
    uint16_t *hist = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(rank) + (size_t)cap * 2);
    uint32_t *pbm = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(hist) + (size_t)(NT / 32) * 512);   // pair starts, 1 bit per slot
    uint32_t *gbm = pbm + cap / 32 + 1;                                   // (target, strand) group starts
    uint32_t *bins = reinterpret_cast<uint32_t *>(bufY);                  // round planning only (before staging): <= 8 * cap bytes
    int nbins_log = 11;                                                   // MAP_BINS = 2^11, fewer if the buffer is smaller
    while ((4u << nbins_log) > 8u * cap) --nbins_log;
    const uint32_t nbins = 1u << nbins_log;
    const int hshift = A.rel_bits > nbins_log ? A.rel_bits - nbins_log : 0;
    constexpr int BPT = MAP_BINS / MAP_THREADS;                           // bins per thread in the prefix pass

    const int PB = A.PB;
    unsigned long long st_anchors = 0, st_groups = 0, st_rounds = 0;     // st_groups: per warp (lane 0), others: thread 0
    long long cy[5] = {0, 0, 0, 0, 0};                                    // thread 0: cycles per phase
    const long long t_begin = clock64();

    // exclusive scan of one total per warp: returns this warp's base, *total = sum
    auto scan_warp_totals = [&](uint32_t wt, uint32_t *total) -> uint32_t {
        if (lane == 0) s_wsum[warp] = wt;
        __syncthreads();
        if (warp == 0) {
            const uint32_t x = lane < MAP_WARPS ? s_wsum[lane] : 0;
            uint32_t xi = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, xi, o); if ((int)lane >= o) xi += u; }
            if (lane < MAP_WARPS) s_wsum[lane] = xi - x;
            if (lane == MAP_WARPS - 1) s_wsum[MAP_WARPS] = xi;
        }
        __syncthreads();
        *total = s_wsum[MAP_WARPS];
        const uint32_t base = s_wsum[warp];
        __syncthreads();
        return base;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(A.work_ctr, 1u);
        __syncthreads();
        if (s_q >= A.nq) break;
        long long tc = clock64();
        const uint32_t qid = A.qorder[s_q];
        const uint64_t a0 = A.qaoff[qid] - A.a_base;
        const uint32_t Aq = (uint32_t)(A.qaoff[qid + 1] - A.qaoff[qid]);
        if (Aq == 0) continue;
        const KT *gk = reinterpret_cast<const KT *>(A.akey) + a0;
        const uint32_t *gv = A.aval + a0;
        // every warp owns a contiguous slice of the query's anchors (order-preserving compaction in multi-round mode)
        const uint32_t slice = ((Aq + MAP_WARPS - 1) / MAP_WARPS + 31) / 32 * 32;
        const uint32_t w0 = min(warp * slice, Aq), w1 = min(w0 + slice, Aq);

        uint32_t nrounds = 1;
        if (Aq > cap) {
            // plan rounds: histogram of the anchors over target ids, prefix sums, then R ranges of about Aq / R anchors
            // each (R sized for 85 % fill); every range is checked against the capacity, and a plan that does not fit
            // falls back to the exact greedy partition
            for (uint32_t i = tid; i < nbins; i += MAP_THREADS) bins[i] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < Aq; i += MAP_THREADS) atomicAdd(&bins[(uint32_t)(gk[i] >> (PB + 1)) >> hshift], 1u);
            __syncthreads();
            uint32_t mine = 0;
#pragma unroll
            for (int j = 0; j < BPT; ++j) { const uint32_t b = tid * BPT + j; mine += b < nbins ? bins[b] : 0; }
            uint32_t tot;
            uint32_t E = block_excl_scan<NT>(mine, s_wsum, tot);             // anchors in the bins before this thread's first bin
            const uint32_t R = (uint32_t)(((uint64_t)Aq * 100 + (uint64_t)cap * 85 - 1) / ((uint64_t)cap * 85));
            const uint32_t tgt = (Aq + R - 1) / R;
            const bool quick = R <= MAP_MAX_ROUNDS;
            if (quick) {
                for (uint32_t k = tid; k <= R; k += MAP_THREADS) {
                    s_rbound[k] = k == 0 ? 0u : k == R ? 0xffffffffu : 0xfffffffeu;   // 0xfffffffe: range never started (empty tail)
                    s_rE[k] = k == 0 ? 0u : Aq;
                }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < BPT; ++j) {                            // range k starts at the first bin b with E[b] >= k * tgt
                    const uint32_t b = tid * BPT + j;
                    if (b > 0 && b < nbins) {
                        const uint32_t e_prev = E - bins[b - 1];
                        for (uint32_t k = e_prev / tgt + 1; k <= E / tgt && k < R; ++k) { s_rbound[k] = b << hshift; s_rE[k] = E; }
                    }
                    if (b < nbins) E += bins[b];
                }
                __syncthreads();
                const bool bad = tid < R && s_rE[tid + 1] - s_rE[tid] > cap;   // anchors of range tid
                const int any_bad = __syncthreads_or(bad);
                if (tid == 0) s_nrounds = any_bad ? 0 : R;
            } else if (tid == 0) s_nrounds = 0;
            __syncthreads();
            const bool greedy = s_nrounds == 0;                           // read by everyone BEFORE thread 0 rewrites it below
            __syncthreads();
            if (greedy) {                                                 // exact greedy partition (serial, rare)
                if (tid == 0) {
                    uint32_t r = 0, acc = 0;
                    bool ok = true;
                    s_rbound[0] = 0;
                    for (uint32_t b = 0; b < nbins; ++b) {
                        const uint32_t c = bins[b];
                        if (c > cap) { ok = false; break; }
                        if (acc + c > cap) {
                            if (++r >= MAP_MAX_ROUNDS) { ok = false; break; }
                            s_rbound[r] = b << hshift;
                            acc = 0;
                        }
                        acc += c;
                    }
                    s_rbound[r + 1] = 0xffffffffu;
                    s_nrounds = ok ? r + 1 : 0;
                }
                __syncthreads();
            }
            nrounds = s_nrounds;
            if (nrounds == 0) {
                if (tid == 0) A.deferred[atomicAdd(A.n_deferred, 1u)] = qid;
                continue;
            }
        }
        if (nrounds > 1) {
            // Partition the query's anchors by round ONCE, stably, into the scratch arrays (a counting sort by round id,
            // every warp over its contiguous slice): each round then stages a contiguous range instead of filtering the
            // whole query again.  Counters (round, warp) are u16 and alias the radix histogram.
            uint16_t *cnt = hist;                                          // [round * MAP_WARPS + warp], round < 256
            for (uint32_t i = tid; i < MAP_WARPS * 128; i += MAP_THREADS) reinterpret_cast<uint32_t *>(hist)[i] = 0;
            for (uint32_t b = tid; b < nbins; b += MAP_THREADS) {          // bin -> round (boundaries are multiples of a bin)
                const uint32_t first_rel = b << hshift;
                uint32_t lo = 0, hi = nrounds;                             // largest k with s_rbound[k] <= first_rel
                while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (s_rbound[mid] <= first_rel) lo = mid; else hi = mid; }
                bins[b] = lo;
            }
            __syncthreads();
            const uint32_t lt = (1u << lane) - 1;
            for (uint32_t b = w0; b < w1; b += 32) {
                const uint32_t i = b + lane;
                const uint32_t rd = i < w1 ? bins[(uint32_t)(gk[i] >> (PB + 1)) >> hshift] : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, rd);
                if (i < w1 && (peers & lt) == 0) cnt[rd * MAP_WARPS + warp] = (uint16_t)(cnt[rd * MAP_WARPS + warp] + __popc(peers));
                __syncwarp();
            }
            __syncthreads();
            if (tid < nrounds) {                                           // per round: exclusive offsets of the warps, total
                uint32_t run = 0;
                for (int w = 0; w < MAP_WARPS; ++w) { const uint32_t v = cnt[tid * MAP_WARPS + w]; cnt[tid * MAP_WARPS + w] = (uint16_t)run; run += v; }
                s_rN[tid] = run;
            }
            __syncthreads();
            if (tid == 0) { uint32_t run = 0; for (uint32_t k = 0; k < nrounds; ++k) { s_rE[k] = run; run += s_rN[k]; } }   // round starts
            __syncthreads();
            KT *gk2 = reinterpret_cast<KT *>(A.akey2) + a0;
            uint32_t *gv2 = A.aval2 + a0;
            for (uint32_t b = w0; b < w1; b += 32) {
                const uint32_t i = b + lane;
                const KT k = i < w1 ? gk[i] : (KT)0;
                const uint32_t rd = i < w1 ? bins[(uint32_t)(k >> (PB + 1)) >> hshift] : 0xffffffffu;
                const uint32_t peers = __match_any_sync(0xffffffffu, rd);
                if (i < w1) {
                    const uint32_t at = s_rE[rd] + cnt[rd * MAP_WARPS + warp] + __popc(peers & lt);
                    gk2[at] = k; gv2[at] = gv[i];
                }
                __syncwarp();
                if (i < w1 && (peers & lt) == 0) cnt[rd * MAP_WARPS + warp] = (uint16_t)(cnt[rd * MAP_WARPS + warp] + __popc(peers));
                __syncwarp();
            }
            __threadfence_block();
            __syncthreads();
        }
        { const long long now = clock64(); cy[0] += now - tc; tc = now; }
        uint32_t round_base = 0;                                         // anchors written back by earlier rounds
        for (uint32_t rd = 0; rd < nrounds; ++rd) {
            // ---- 1. stage: HBM -> smem, generation order kept: a straight copy, eight coalesced loads in flight per thread
            const uint32_t N = nrounds == 1 ? Aq : s_rN[rd];
            if (N > cap) { if (tid == 0) atomicAdd(A.stats + 3, 1ull); continue; }   // cannot happen (planned)
            const KT *sk = nrounds == 1 ? gk : reinterpret_cast<const KT *>(A.akey2) + a0 + s_rE[rd];
            const uint32_t *sv = nrounds == 1 ? gv : A.aval2 + a0 + s_rE[rd];
            for (uint32_t b = 0; b < N; b += 4 * MAP_THREADS) {
                KT k[4]; uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const uint32_t i = b + u * MAP_THREADS + tid; if (i < N) { k[u] = sk[i]; v[u] = sv[i]; } }
#pragma unroll
                for (int u = 0; u < 4; ++u) { const uint32_t i = b + u * MAP_THREADS + tid; if (i < N) { kX[i] = k[u]; vX[i] = v[u]; } }
            }
            __syncthreads();
            { const long long now = clock64(); cy[1] += now - tc; tc = now; }
            if (N == 0) continue;
            if (tid == 0) { st_anchors += N; ++st_rounds; }

            // ---- 2. sort by (target, strand): groups become contiguous, still in generation order ----------------
            const int fin = block_radix_sort<KT, NT>(kX, vX, kY, vY, N, PB, A.key_bits - PB, hist, rank, s_wsum);
            KT *K = fin ? kY : kX;
            uint32_t *V = fin ? vY : vX;
            uint2 *scr = reinterpret_cast<uint2 *>(fin ? map_smem : bufY);        // the other buffer: scratch of the group sorts
            if (tid == 0) s_unit = 0;
            __syncthreads();
            { const long long now = clock64(); cy[2] += now - tc; tc = now; }

            // ---- 3. groups -------------------------------------------------------------------------------------------
            // 3a. boundaries of the sorted array as bit maps, built BEFORE any group is reordered: bit i of pbm = slot i
            // starts a (query, target) pair, of gbm = slot i starts a (target, strand) group; bit N closes both.
            for (uint32_t b = warp * 32; b <= N; b += NT) {
                const uint32_t i = b + lane;
                bool ps = i == N || i == 0, gs = ps;
                if (i > 0 && i < N) {
                    const KT prev = K[i - 1] >> PB, cur = K[i] >> PB;
                    ps = (prev >> 1) != (cur >> 1);
                    gs = prev != cur;
                }
                if (i > N) ps = gs = false;
                const uint32_t pw = __ballot_sync(0xffffffffu, ps), gw = __ballot_sync(0xffffffffu, gs);
                if (lane == 0) { pbm[b >> 5] = pw; gbm[b >> 5] = gw; }
            }
            __syncthreads();
            // first set bit of a map strictly after slot `after` (bit N always ends the search)
            auto next_bit = [&](const uint32_t *bm, uint32_t after) -> uint32_t {
                uint32_t w = (after + 1) >> 5, first = 0xffffffffu << ((after + 1) & 31);
                for (;;) {
                    const uint32_t wi = w + lane;
                    uint32_t x = wi <= (N >> 5) ? bm[wi] : 0u;
                    if (lane == 0) x &= first;
                    const uint32_t nz = __ballot_sync(0xffffffffu, x != 0u);
                    if (nz) {
                        const int src = __ffs(nz) - 1;
                        return ((w + (uint32_t)src) << 5) + (uint32_t)(__ffs(__shfl_sync(0xffffffffu, x, src)) - 1);
                    }
                    w += 32; first = 0xffffffffu;
                }
            };
            // 3b. warps claim 32-slot windows; the pairs starting there are theirs.  A warp reads keys of its own pairs
            // only; everything about other pairs comes from the bit maps.
            uint2 *out = A.anc + a0 + round_base;
            for (;;) {
                uint32_t u = 0;
                if (lane == 0) u = atomicAdd(&s_unit, 1u);
                u = __shfl_sync(0xffffffffu, u, 0);
                const uint32_t wnd = u * 32;
                if (wnd >= N) break;
                const uint32_t pw = pbm[u], gw = gbm[u];                   // this window's boundary bits (same in every lane)
                uint32_t pm = pw;
                if ((N >> 5) == u) pm &= ~(1u << (N & 31));                // the closing bit is not a pair
                while (pm) {
                    const uint32_t sb = (uint32_t)(__ffs(pm) - 1), s = wnd + sb;
                    pm &= pm - 1;
                    // end of the pair / first strand change: usually inside this window (registers only), else searched
                    const uint32_t above = sb == 31 ? 0u : 0xffffffffu << (sb + 1);
                    const uint32_t pe = pw & above, ge = gw & above;
                    const uint32_t e = pe ? wnd + (uint32_t)(__ffs(pe) - 1) : next_bit(pbm, wnd + 31);
                    uint32_t split = ge ? wnd + (uint32_t)(__ffs(ge) - 1) : next_bit(gbm, wnd + 31);   // <= e: a pair start is a group start
                    if (e - s < (uint32_t)A.min_cnt) { if (lane == 0) st_groups += 1 + (split < e); continue; }   // too short for a chain
                    const KT ks = K[s];
                    const KT pk = ks >> (PB + 1);
                    if (split == e && ((ks >> PB) & 1)) split = s;         // one strand only, the reverse one
                    const uint32_t nf = split - s, nr = e - split;
                    if (lane == 0) st_groups += (nf > 0) + (nr > 0);
                    const bool vf = (int32_t)nf >= A.min_cnt && (int64_t)nf * A.max_span >= A.min_sc;
                    const bool vr = (int32_t)nr >= A.min_cnt && (int64_t)nr * A.max_span >= A.min_sc;
                    if (!vf && !vr) continue;
                    uint32_t pair = 0;
                    if (lane == 0) pair = atomicAdd(A.pair_ctr, 1u);
                    pair = __shfl_sync(0xffffffffu, pair, 0);
#pragma unroll 1
                    for (int strand = 0; strand < 2; ++strand) {
                        if (!(strand ? vr : vf)) continue;
                        const uint32_t g0 = strand ? split : s, n = strand ? nr : nf;
                        warp_sort_group<KT>(K + g0, V + g0, (int32_t)n, PB, scr + g0, hist + warp * 256, rank + g0, lane);
                        for (uint32_t j = lane; j < n; j += 32)
                            out[g0 + j] = make_uint2((uint32_t)(K[g0 + j] & (KT)(((KT)1 << PB) - 1)), V[g0 + j]);
                        if (lane == 0) {
                            GroupDesc d;
                            d.start = (uint32_t)(a0 + round_base + g0); d.n = n; d.qid = qid; d.rid = (uint32_t)pk + A.rid_base;
                            d.strand = (uint32_t)strand; d.pair = pair;
                            const int c = group_class(n);
                            const uint32_t slot = atomicAdd(&A.gcount[c], 1u);
                            if (slot < A.L.cap[c]) A.L.gl[c][slot] = d; else atomicAdd(A.stats + 3, 1ull);
                        }
                    }
                }
            }
            round_base += N;
            __syncthreads();
            { const long long now = clock64(); cy[3] += now - tc; tc = now; }
        }
    }
    if (tid == 0) {
        atomicAdd(A.stats + 0, st_anchors); atomicAdd(A.stats + 2, st_rounds);
        if (A.cyc) {
            cy[4] = clock64() - t_begin;
            for (int i = 0; i < 5; ++i) atomicAdd(A.cyc + i, (unsigned long long)cy[i]);
        }
    }
    if (lane == 0 && st_groups) atomicAdd(A.stats + 1, st_groups);
}

}  // namespace lrge
