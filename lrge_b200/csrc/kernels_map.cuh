// kernels_map.cuh -- K4-K6 fused: one CTA maps one query read end to end in shared memory.
//
// Replaces, per query, what minimap2's mm_map_frag does after collect_minimizers (map.c): mm_collect_matches
// (seed.c), collect_seed_hits + skip_seed + radix_sort_128x (map.c), mg_lchain_dp + mg_chain_backtrack (lchain.c),
// the coordinate part of mm_gen_regs (hit.c) and liblrge's per-read dedupe + count (twoset.rs:286-302,
// ava.rs:271-306, twoset.rs:477-525) -- SURVEY.md rows a8-a14 -- without the anchors ever touching HBM:
//
//   seed_lookup_kernel   every query minimizer of the batch: bucket table -> compact keys -> (start, n) of its hit list
//   map_query_kernel     persistent CTAs pull queries (longest first) from a work counter; per query
//     1. expand   hit lists (128-bit loads, two 8-byte entries each) -> anchors (sort key, payload) staged in smem,
//                 in minimizer order (the order collect_seed_hits generates them)
//     2. sort     block-wide stable LSD radix sort in smem by (target, strand, target position)
//     3. chain    warps claim 32-anchor windows of the sorted array, find the (query, target) pairs that start in
//                 them and run the chaining DP (+ backtrack where -F / records need it) on the smem-resident anchors
//     4. count    distinct targets with a kept chain -> counts[], had_mapping[]
//   A query whose anchors exceed the smem capacity is mapped in ROUNDS over disjoint target-id ranges (groups never
//   span targets, so rounds are independent); ranges come from a 1024-bin histogram of its hits over target ids.
//   A query that cannot be planned (one bin alone above capacity, > 255 rounds) is appended to `deferred` and the
//   host maps it through the global-memory path (kernels_seed.cuh / CUB sort / chain_group_kernel).
//
// Stable sort + minimizer-order generation = the tie order of the global path (ksort's radix
// sort is unstable; ties are equal (target, strand, position) anchors from repeated query k-mers).
#pragma once
#include "common.cuh"
#include "kernels_index.cuh"
#include "kernels_seed.cuh"
#include "kernels_chain.cuh"

namespace lrge {

constexpr int MAP_MAX_WARPS = 32;
constexpr int MAP_BINS = 1024;         // target-id histogram bins for round planning
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
            if (n > (uint32_t)mid_occ) n = 0;                                   // m[i].flt = 1
            if (n && I.bl && ((I.bl[start >> 5] >> (start & 31)) & 1)) n = 0;   // > mid_occ over all target shards
        }
        hstart[i] = start;
        hn[i] = n;
        seeds += n ? 1 : 0;
        hits += n;
    }
    for (int o = 16; o > 0; o >>= 1) { seeds += __shfl_down_sync(0xffffffffu, seeds, o); hits += __shfl_down_sync(0xffffffffu, hits, o); }
    if ((threadIdx.x & 31) == 0 && hits) { atomicAdd(n_seeds, (unsigned long long)seeds); atomicAdd(n_hits, (unsigned long long)hits); }
}

// ---- K4b-K6 fused ------------------------------------------------------------------------------------------------
struct MapArgs {
    const uint64_t *qx, *qy;         // query minimizers (read, position) order
    const uint32_t *moff;            // nq + 1
    const uint32_t *hstart, *hn;     // seed_lookup_kernel
    IndexView I;
    SkipCtx S;
    ChainParams P;
    CountOut O;
    const uint32_t *qorder;          // queries, longest first
    uint32_t nq;
    uint32_t *work_ctr;
    uint32_t cap;                    // anchors per round (smem capacity), multiple of 32, < 65536
    int PB, key_bits;                // key = ((rid - rid_base) << 1 | rev) << PB | target position
    uint32_t rid_base;               // first target id held by this index (sharded index: the rank's block)
    int hshift;                      // histogram bin = (rid - rid_base) >> hshift
    int max_span;                    // k, or 255 with HPC: a group of n anchors scores at most n * max_span
    uint32_t *deferred, *n_deferred; // queries left to the global path
    unsigned long long *stats;       // [0] anchors [1] (query, target, strand) groups [2] rounds [3] capacity violations
    unsigned long long *cyc;         // optional: SM cycles per phase summed over CTAs: [0] plan [1] expand [2] sort [3] chain [4] total
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
                                uint16_t *hist, uint32_t *rank, uint32_t *wsum) {
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
            if (valid) rank[e] = (uint32_t)myh[d] + __popc(peers & lt);
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
__device__ void warp_sort_group(KT *K, uint32_t *V, int32_t n, int PB, uint2 *scr, uint16_t *myh, uint32_t *rank, uint32_t lane) {
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
            if (valid) rank[e] = (uint32_t)myh[d] + __popc(peers & lt);
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
    for (int32_t e = lane; e < n; e += 32) rank[e] = 0;                   // rank aliases the DP's mark array
    __syncwarp();
}

// per-lane walk over one hit list with four independent 128-bit loads (two 8-byte entries each) in flight;
// fn(entry, j) is called in ascending j
template <typename F>
__device__ __forceinline__ void for_hits(const uint4 *__restrict__ ty4, uint32_t start, uint32_t n, F fn) {
    const uint32_t p1 = (start + n + 1) >> 1;
    for (uint32_t p = start >> 1; p < p1; p += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p + u < p1 ? __ldg(ty4 + p + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t e0 = 2 * (p + u), e1 = e0 + 1;
            if (e0 >= start && e0 < start + n) fn((uint64_t)v[u].y << 32 | v[u].x, e0 - start);
            if (e1 >= start && e1 < start + n) fn((uint64_t)v[u].w << 32 | v[u].z, e1 - start);
        }
    }
}

template <typename KT, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) map_query_kernel(MapArgs A) {
    constexpr int MAP_THREADS = NT, MAP_WARPS = NT / 32;
    extern __shared__ __align__(16) uint8_t map_smem[];
    __shared__ uint32_t s_wsum[MAP_WARPS + 1];
    __shared__ uint32_t s_rbound[MAP_MAX_ROUNDS + 2];
    __shared__ uint32_t s_q, s_qn, s_nrounds, s_unit, s_cnt, s_kept;
    __shared__ uint8_t s_pen[2048];

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cap = A.cap;
    KT *kX = reinterpret_cast<KT *>(map_smem);
    uint32_t *vX = reinterpret_cast<uint32_t *>(map_smem + (size_t)cap * sizeof(KT));
    uint8_t *bufY = map_smem + (size_t)cap * (sizeof(KT) + 4);
    KT *kY = reinterpret_cast<KT *>(bufY);
    uint32_t *vY = reinterpret_cast<uint32_t *>(bufY + (size_t)cap * sizeof(KT));
    int32_t *t = reinterpret_cast<int32_t *>(bufY + (size_t)cap * (sizeof(KT) + 4));
    uint16_t *hist = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(t) + (size_t)cap * 4);

    for (int i = tid; i <= A.P.bw && i < 2048; i += MAP_THREADS) s_pen[i] = (uint8_t)A.P.pen[i];
    const int PB = A.PB;
    const KT posmask = (KT)(((KT)1 << PB) - 1);
    const uint4 *ty4 = reinterpret_cast<const uint4 *>(A.I.ty);
    unsigned long long st_anchors = 0, st_groups = 0, st_rounds = 0;     // st_groups: per warp (lane 0), others: thread 0
    long long cy[5] = {0, 0, 0, 0, 0};                                    // thread 0: cycles per phase
    const long long t_begin = clock64();

    // exclusive scan of one total per warp: returns this warp's base, *total = sum (two barriers)
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

    // Work is pulled one query ahead: while the current query is sorted and chained, the hit lists of the next one are
    // brought into L2 (prefetch.global.L2), so that its expansion does not wait on first-touch DRAM latency.
    if (tid == 0) s_qn = atomicAdd(A.work_ctr, 1u);
    for (;;) {
        __syncthreads();
        if (tid == 0) { s_q = s_qn; s_qn = s_qn < A.nq ? atomicAdd(A.work_ctr, 1u) : s_qn; s_cnt = 0; s_kept = 0; }
        __syncthreads();
        if (s_q >= A.nq) break;
        long long tc = clock64();
        const uint32_t qid = A.qorder[s_q];
        if (s_qn < A.nq) {
            const uint32_t qn = A.qorder[s_qn];
            const uint32_t e1 = A.moff[qn + 1];
            for (uint32_t i = A.moff[qn] + tid; i < e1; i += MAP_THREADS) {
                const uint32_t n = A.hn[i];
                if (n) {
                    const uint64_t *p0 = A.I.ty + A.hstart[i];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p0));
                    if ((((uintptr_t)p0) & 127) + (size_t)n * 8 > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + n - 1));
                }
            }
        }
        const uint32_t m0 = A.moff[qid], m1 = A.moff[qid + 1];
        const uint32_t qlen = A.S.len_q[qid];
        // every warp owns a contiguous slice of the query's minimizers: expansion needs one block-wide scan per round
        const uint32_t slice = ((m1 - m0 + MAP_WARPS - 1) / MAP_WARPS + 31) / 32 * 32;
        const uint32_t w0 = min(m0 + warp * slice, m1), w1 = min(w0 + slice, m1);

        // total hit-list entries of this query (an upper bound of its anchors) and of every warp's slice
        uint32_t wt = 0;
        for (uint32_t b = w0; b < w1; b += 32) wt += b + lane < w1 ? A.hn[b + lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, o);
        uint32_t H;
        uint32_t wbase = scan_warp_totals(wt, &H);
        if (H == 0) continue;

        // does a hit survive skip_seed and belong to the target range [rlo, rhi) (ids relative to rid_base)?
        auto keep = [&](uint64_t r, uint32_t q_pos, uint32_t rlo, uint32_t rhi) -> bool {
            const uint32_t rel = (uint32_t)(r >> 32) - A.rid_base;
            if (rel < rlo || rel >= rhi) return false;
            return !skip_hit(r, qid, q_pos, A.S);
        };

        const bool filtered = A.S.check_names || A.S.sharded || H > cap;
        uint32_t nrounds = 1;
        if (H > cap) {
            // plan rounds: histogram of the kept hits over target ids, then greedy ranges of <= cap anchors
            uint32_t *bins = reinterpret_cast<uint32_t *>(t);
            for (uint32_t i = tid; i < MAP_BINS; i += MAP_THREADS) bins[i] = 0;
            __syncthreads();
            for (uint32_t b = w0; b < w1; b += 32) {
                const uint32_t i = b + lane;
                const uint32_t n = i < w1 ? A.hn[i] : 0;
                if (!n) continue;
                const uint32_t start = A.hstart[i], q_pos = (uint32_t)A.qy[i];
                for_hits(ty4, start, n, [&](uint64_t r, uint32_t) {
                    if (keep(r, q_pos, 0u, 0xffffffffu)) atomicAdd(&bins[((uint32_t)(r >> 32) - A.rid_base) >> A.hshift], 1u);
                });
            }
            __syncthreads();
            if (tid == 0) {
                uint32_t r = 0, acc = 0;
                bool ok = true;
                s_rbound[0] = 0;
                for (uint32_t b = 0; b < MAP_BINS; ++b) {
                    const uint32_t c = bins[b];
                    if (c > cap) { ok = false; break; }
                    if (acc + c > cap) {
                        if (++r >= MAP_MAX_ROUNDS) { ok = false; break; }
                        s_rbound[r] = b << A.hshift;
                        acc = 0;
                    }
                    acc += c;
                }
                s_rbound[r + 1] = 0xffffffffu;
                s_nrounds = ok ? r + 1 : 0;
            }
            __syncthreads();
            nrounds = s_nrounds;
            if (nrounds == 0) {
                if (tid == 0) A.deferred[atomicAdd(A.n_deferred, 1u)] = qid;
                continue;
            }
        } else {
            if (tid == 0) { s_rbound[0] = 0; s_rbound[1] = 0xffffffffu; }
            __syncthreads();
        }

        { const long long now = clock64(); cy[0] += now - tc; tc = now; }
        for (uint32_t rd = 0; rd < nrounds; ++rd) {
            const uint32_t rlo = s_rbound[rd], rhi = s_rbound[rd + 1];
            // ---- 1. expand: anchors in minimizer order into (kX, vX) ------------------------------------------
            uint32_t N = H;
            if (filtered) {                                               // kept anchors of every warp's slice in this round
                wt = 0;
                for (uint32_t b = w0; b < w1; b += 32) {
                    const uint32_t i = b + lane;
                    const uint32_t n = i < w1 ? A.hn[i] : 0;
                    if (!n) continue;
                    const uint32_t start = A.hstart[i], q_pos = (uint32_t)A.qy[i];
                    for_hits(ty4, start, n, [&](uint64_t r, uint32_t) { wt += keep(r, q_pos, rlo, rhi); });
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, o);
                wbase = scan_warp_totals(wt, &N);
            }
            if (N > cap) { if (tid == 0) atomicAdd(A.stats + 3, 1ull); N = 0; }   // cannot happen (planned)
            if (N) {
                uint32_t run = wbase;
                for (uint32_t b = w0; b < w1; b += 32) {
                    const uint32_t i = b + lane;
                    uint32_t start = 0, n = 0, q_pos = 0, q_span = 0;
                    if (i < w1) {
                        n = A.hn[i];
                        if (n) { start = A.hstart[i]; q_pos = (uint32_t)A.qy[i]; q_span = (uint32_t)(A.qx[i] & 0xff); }
                    }
                    uint32_t kept = n;
                    if (filtered && n) {
                        kept = 0;
                        for_hits(ty4, start, n, [&](uint64_t r, uint32_t) { kept += keep(r, q_pos, rlo, rhi); });
                    }
                    uint32_t incl = kept;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
                    uint32_t o = run + incl - kept;
                    run += __shfl_sync(0xffffffffu, incl, 31);
                    if (kept) {
                        const uint32_t qp_f = q_pos >> 1, qp_r = qlen - ((q_pos >> 1) + 1 - q_span) - 1;
                        auto emit = [&](uint64_t r, uint32_t at) {
                            if (at >= cap) return;
                            const uint32_t rev = ((uint32_t)r ^ q_pos) & 1;
                            const KT rel = (KT)((uint32_t)(r >> 32) - A.rid_base);
                            kX[at] = (KT)((((rel << 1) | (KT)rev) << PB) | (KT)((uint32_t)r >> 1));
                            vX[at] = q_span << 24 | (rev ? qp_r : qp_f);
                        };
                        if (!filtered) for_hits(ty4, start, n, [&](uint64_t r, uint32_t j) { emit(r, o + j); });
                        else for_hits(ty4, start, n, [&](uint64_t r, uint32_t) { if (keep(r, q_pos, rlo, rhi)) emit(r, o++); });
                    }
                }
            }
            __syncthreads();
            { const long long now = clock64(); cy[1] += now - tc; tc = now; }
            if (N == 0) continue;
            if (tid == 0) { st_anchors += N; ++st_rounds; }

            // ---- 2. sort by (target, strand): groups become contiguous, still in generation order ----------------
            const int fin = block_radix_sort<KT, NT>(kX, vX, kY, vY, N, PB, A.key_bits - PB, hist, reinterpret_cast<uint32_t *>(t), s_wsum);
            KT *K = fin ? kY : kX;
            uint32_t *V = fin ? vY : vX;
            int2 *fp = reinterpret_cast<int2 *>(fin ? map_smem : bufY);           // the other buffer: (f, p) scratch
            for (uint32_t i = tid; i < N; i += MAP_THREADS) t[i] = 0;
            if (tid == 0) s_unit = 0;
            __syncthreads();
            { const long long now = clock64(); cy[2] += now - tc; tc = now; }

            // ---- 3. chain: warps claim 32-anchor windows; the (query, target) pairs starting there are theirs ---
            uint32_t my_cnt = 0;
            for (;;) {
                uint32_t u = 0;
                if (lane == 0) u = atomicAdd(&s_unit, 1u);
                u = __shfl_sync(0xffffffffu, u, 0);
                const uint32_t wnd = u * 32;
                if (wnd >= N) break;
                const uint32_t i = wnd + lane;
                bool ps = false;
                if (i < N) ps = i == 0 || (K[i - 1] >> (PB + 1)) != (K[i] >> (PB + 1));
                uint32_t pm = __ballot_sync(0xffffffffu, ps);
                while (pm) {
                    const uint32_t s = wnd + (uint32_t)(__ffs(pm) - 1);
                    pm &= pm - 1;
                    const KT pk = K[s] >> (PB + 1);
                    // pair end and strand split, 32 anchors per step
                    uint32_t e = N, split = 0xffffffffu;
                    for (uint32_t b = s; b < N; b += 32) {
                        const uint32_t j = b + lane;
                        const bool in = j < N;
                        const KT kk = in ? K[j] : (KT)0;
                        const bool same = in && (kk >> (PB + 1)) == pk;
                        const uint32_t diff = __ballot_sync(0xffffffffu, in && !same);
                        const uint32_t rv = __ballot_sync(0xffffffffu, same && ((kk >> PB) & 1));
                        if (split == 0xffffffffu && rv) split = b + (uint32_t)(__ffs(rv) - 1);
                        if (diff) { e = b + (uint32_t)(__ffs(diff) - 1); break; }
                    }
                    if (split == 0xffffffffu || split > e) split = e;
                    const uint32_t rid = (uint32_t)pk + A.rid_base;
                    int kept_any = 0, pass_any = 0;
#pragma unroll 1
                    for (int strand = 0; strand < 2; ++strand) {
                        const uint32_t g0 = strand ? split : s;
                        const int32_t n = (int32_t)(strand ? e - split : split - s);
                        if (n == 0) continue;
                        if (lane == 0) ++st_groups;
                        if (n < A.P.min_cnt || (int64_t)n * A.max_span < A.P.min_sc) continue;
                        warp_sort_group<KT>(K + g0, V + g0, n, PB, reinterpret_cast<uint2 *>(fp + g0), hist + warp * 256,
                                            reinterpret_cast<uint32_t *>(t + g0), lane);
                        int kept = 0, pass = 0;
                        chain_one_group(AncSoA<KT>{K + g0, V + g0, posmask}, n, fp + g0, t + g0, A.P, s_pen, A.O, lane, qid, rid,
                                        strand, kept, pass);
                        kept_any |= kept; pass_any |= pass;
                    }
                    if (lane == 0) {
                        if (kept_any) s_kept = 1;
                        if (pass_any) {
                            if (A.O.mode == 0) ++my_cnt;
                            else if (A.O.mode == 1) {
                                if (A.O.rank_q[qid] != A.O.rank_t[rid]) { atomicAdd(&A.O.counts[qid], 1u); atomicAdd(&A.O.counts[rid], 1u); }
                            } else atomicAdd(&A.O.counts[rid], 1u);
                        }
                    }
                }
            }
            if (lane == 0 && my_cnt) atomicAdd(&s_cnt, my_cnt);
            __syncthreads();
            { const long long now = clock64(); cy[3] += now - tc; tc = now; }
        }
        // ---- 4. per-query results ------------------------------------------------------------------------------
        if (tid == 0 && A.O.mode == 0) {
            A.O.counts[qid] = s_cnt;
            if (s_kept) A.O.had_mapping[qid] = 1;
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
