// kernels_seed.cuh -- K4: query-side seed filter, index lookup and anchor emission.
//
// Replaces minimap2 mm_seed_mz_flt / mm_seed_collect_all / mm_collect_matches (seed.c) and
// skip_seed / collect_seed_hits (map.c), SURVEY.md rows a7-a9, for ALL queries of a batch at once.
// Anchors leave this stage as (64-bit sort key, 32-bit payload):
//     key = ((qid_local << RB | rid) << 1 | rev) << PB | rpos        payload = q_span << 24 | qpos'
// Sorting by key groups anchors per (query, target), forward strand before reverse, ascending target
// position -- the order mg_lchain_dp needs inside one (target, strand) group.  Groups never interact
// in minimap2's DP, so ordering targets before strands is free; it makes the two
// strand groups of one (query, target) pair adjacent for the distinct-target count.
#pragma once
#include "common.cuh"
#include "kernels_index.cuh"

namespace lrge {

constexpr int MZ_TABLE = 8192;
constexpr uint32_t SEED_REPETITIVE = 0xffffffffu;   // hstart of a seed that is in the index but occurs > mid_occ times

// mm_seed_mz_flt: flag query minimizers whose x occurs > mid_occ times AND > q_occ_frac * n times in
// the query.  A counting table over-estimates occurrences; candidates are then counted exactly.
__global__ void __launch_bounds__(256) mz_flt_kernel(const uint64_t *__restrict__ qx, const uint32_t *__restrict__ moff,
                                                     int mid_occ, float q_occ_frac, uint8_t *__restrict__ flt) {
    __shared__ uint32_t tab[MZ_TABLE];
    const uint32_t q = blockIdx.x;
    const uint32_t b = moff[q], n = moff[q + 1] - b;
    if ((int64_t)n <= (int64_t)mid_occ || q_occ_frac <= 0.0f || mid_occ <= 0) return;
    for (int i = threadIdx.x; i < MZ_TABLE; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t x = qx[b + i];
        atomicAdd(&tab[(uint32_t)((x * 0x9E3779B97F4A7C15ULL) >> 51)], 1u);
    }
    __syncthreads();
    const float thr = __fmul_rn((float)n, q_occ_frac);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t x = qx[b + i];
        if (tab[(uint32_t)((x * 0x9E3779B97F4A7C15ULL) >> 51)] <= (uint32_t)mid_occ) continue;
        int32_t cnt = 0;
        for (uint32_t j = 0; j < n; ++j) cnt += (qx[b + j] == x);
        if (cnt > mid_occ && (float)cnt > thr) flt[b + i] = 1;
    }
}

struct SkipCtx {
    const uint32_t *rank_q, *rank_t;   // lexicographic name rank over the union of both sets
    const uint32_t *len_q, *len_t;     // original read lengths
    int check_names;                   // any (query, target) name comparison can matter
    int no_dual;                       // MM_F_NO_DUAL: skip when qname > tname
    uint32_t shard_lo, shard_hi;       // keep hits of targets rid in [shard_lo, shard_hi) only
    int sharded;                       // the range above is a proper subset of the index
};

// skip_seed (map.c) + target sharding
__device__ __forceinline__ bool skip_hit(uint64_t r, uint32_t qid, uint32_t q_pos, const SkipCtx &S) {
    const uint32_t rid = (uint32_t)(r >> 32);
    if (S.sharded && (rid < S.shard_lo || rid >= S.shard_hi)) return true;
    if (S.check_names) {
        const uint32_t rq = S.rank_q[qid], rt = S.rank_t[rid];
        if (rq == rt && S.len_t[rid] == S.len_q[qid] && ((uint32_t)r >> 1) == (q_pos >> 1)) return true;   // diagonal
        if (S.no_dual && rq > rt) return true;
    }
    return false;
}

// A hit list of at least this many entries is expanded by the whole warp (coalesced reads of the list, coalesced writes
// of the anchors) instead of by its own thread: ava-pb lists hold up to mid_occ ~ 70 entries.
constexpr uint32_t COOP_MIN_HITS = 16;

// One thread per query minimizer: lookup, mid_occ filter, exact anchor count.
__global__ void seed_count_kernel(const uint64_t *__restrict__ qx, const uint64_t *__restrict__ qy,
                                  const uint8_t *__restrict__ flt, uint64_t MQ, IndexView I, int mid_occ, SkipCtx S,
                                  uint32_t *__restrict__ hstart, uint32_t *__restrict__ hn, uint32_t *__restrict__ acnt,
                                  unsigned long long *n_seeds, unsigned long long *n_hits) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool live = i < MQ && !flt[i];
    const bool scan = S.check_names || S.sharded;
    uint32_t cnt = 0, n = 0, start = 0;
    uint64_t y = 0;
    if (live) {
        const uint64_t x = qx[i];
        n = index_lookup(I, x >> 8, (uint32_t)mid_occ, &start);
        bool rep = n > (uint32_t)mid_occ;                     // m[i].flt = 1 (seed.c mm_collect_matches)
        if (n && I.bl && ((I.bl[start >> 5] >> (start & 31)) & 1)) rep = true;   // > mid_occ over all target shards
        if (rep) { n = 0; start = SEED_REPETITIVE; }
        if (n && scan) y = qy[i];
        if (!scan) cnt = n;
    }
    if (scan) {
        // long lists: the warp counts the kept hits of one list together; the first 32 entries of up to four lists are
        // fetched back to back (four loads in flight per lane)
        uint32_t big = __ballot_sync(0xffffffffu, n >= COOP_MIN_HITS);
        while (big) {
            int src[4];
            uint32_t s_start[4], s_n[4];
            uint64_t r0[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { src[k] = big ? __ffs(big) - 1 : -1; big &= big - 1; }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int sl = src[k] < 0 ? 0 : src[k];
                s_start[k] = __shfl_sync(0xffffffffu, start, sl);
                s_n[k] = src[k] < 0 ? 0u : __shfl_sync(0xffffffffu, n, sl);
                r0[k] = lane < s_n[k] ? __ldg(I.ty + s_start[k] + lane) : 0ull;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (src[k] < 0) continue;
                const uint64_t s_y = __shfl_sync(0xffffffffu, y, src[k]);
                uint32_t kept = 0;
                for (uint32_t b = 0; b < s_n[k]; b += 32) {
                    const uint32_t e = b + lane;
                    const uint64_t r = b == 0 ? r0[k] : (e < s_n[k] ? __ldg(I.ty + s_start[k] + e) : 0ull);
                    const bool keep = e < s_n[k] && !skip_hit(r, (uint32_t)(s_y >> 32), (uint32_t)s_y, S);
                    kept += __popc(__ballot_sync(0xffffffffu, keep));
                }
                if ((int)lane == src[k]) cnt = kept;
            }
        }
        if (n && n < COOP_MIN_HITS) {
            // 128-bit loads, two hit-list entries each, four loads in flight
            const uint4 *ty4 = reinterpret_cast<const uint4 *>(I.ty);
            const uint32_t p1 = (start + n + 1) >> 1;
            for (uint32_t p = start >> 1; p < p1; p += 4) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = p + u < p1 ? __ldg(ty4 + p + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t e0 = 2 * (p + u), e1 = e0 + 1;
                    if (e0 >= start && e0 < start + n) cnt += !skip_hit((uint64_t)v[u].y << 32 | v[u].x, (uint32_t)(y >> 32), (uint32_t)y, S);
                    if (e1 >= start && e1 < start + n) cnt += !skip_hit((uint64_t)v[u].w << 32 | v[u].z, (uint32_t)(y >> 32), (uint32_t)y, S);
                }
            }
        }
    }
    if (live) {
        hstart[i] = start;
        hn[i] = n;                                            // hits to visit
        acnt[i] = cnt;                                        // anchors kept
    } else if (i < MQ) { hstart[i] = 0; hn[i] = 0; acnt[i] = 0; }
    // stats
    unsigned int seeds = n ? 1 : 0, hits = n;
    for (int o = 16; o > 0; o >>= 1) { seeds += __shfl_down_sync(0xffffffffu, seeds, o); hits += __shfl_down_sync(0xffffffffu, hits, o); }
    if ((threadIdx.x & 31) == 0 && hits) { atomicAdd(n_seeds, (unsigned long long)seeds); atomicAdd(n_hits, (unsigned long long)hits); }
}

struct KeyLayout { int PB, RB; uint32_t rid_base; };   // bits of target position / target id in the sort key; first local target id

// One thread per query minimizer of the batch [m0, m1): expands its hit list to anchors, in hit-list order, so that a
// query's anchors are contiguous and in the order collect_seed_hits generates them.  The hit list (8-byte entries) is
// read with 128-bit loads, two entries each, four loads in flight per thread.
// aoff = exclusive scan of anchor counts over ALL query minimizers (64-bit); a0 = aoff[m0].
// WITH_QID: the 64-bit key of the device-wide sort (query bits on top).  Otherwise the key the per-query kernel sorts in
// shared memory: ((rid - K.rid_base) << 1 | rev) << PB | target pos, as KT.
template <typename KT, bool WITH_QID>
__global__ void __launch_bounds__(256) anchor_fill_kernel(const uint64_t *__restrict__ qx, const uint64_t *__restrict__ qy,
                                   const uint32_t *__restrict__ hstart, const uint32_t *__restrict__ hn,
                                   const uint32_t *__restrict__ acnt, const uint64_t *__restrict__ aoff, uint64_t m0, uint64_t m1, uint64_t a0,
                                   uint32_t q0, IndexView I, SkipCtx S, KeyLayout K,
                                   KT *__restrict__ akey, uint32_t *__restrict__ aval) {
    const uint64_t i = m0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, lt = (1u << lane) - 1;
    const bool active = i < m1 && acnt[i] != 0;
    const bool scan = S.check_names || S.sharded;
    uint32_t n = 0, start = 0, q_span = 0, qlen = 0;
    uint64_t y = 0, o = 0;
    if (active) {
        n = hn[i]; start = hstart[i];
        y = qy[i]; q_span = (uint32_t)(qx[i] & 0xff);
        qlen = S.len_q[(uint32_t)(y >> 32)];
        o = aoff[i] - a0;
    }
    auto emit = [&](uint64_t at, uint64_t r, uint32_t qid, uint32_t q_pos, uint32_t span, uint32_t ql) {
        const uint32_t rev = ((uint32_t)r ^ q_pos) & 1;
        const uint32_t qp_f = q_pos >> 1, qp_r = ql - ((q_pos >> 1) + 1 - span) - 1;
        if (WITH_QID) akey[at] = (KT)(((((uint64_t)(qid - q0) << K.RB | (r >> 32)) << 1 | rev) << K.PB) | ((uint32_t)r >> 1));
        else akey[at] = (KT)(((((KT)((uint32_t)(r >> 32) - K.rid_base)) << 1 | (KT)rev) << K.PB) | (KT)((uint32_t)r >> 1));
        aval[at] = span << 24 | (rev ? qp_r : qp_f);
    };
    // long hit lists: 32 consecutive entries per step across the warp; the first 32 entries of up to four lists are
    // fetched back to back (four loads in flight per lane)
    uint32_t big = __ballot_sync(0xffffffffu, active && n >= COOP_MIN_HITS);
    while (big) {
        int src[4];
        uint32_t s_start[4], s_n[4];
        uint64_t r0[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { src[k] = big ? __ffs(big) - 1 : -1; big &= big - 1; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sl = src[k] < 0 ? 0 : src[k];
            s_start[k] = __shfl_sync(0xffffffffu, start, sl);
            s_n[k] = src[k] < 0 ? 0u : __shfl_sync(0xffffffffu, n, sl);
            r0[k] = lane < s_n[k] ? __ldg(I.ty + s_start[k] + lane) : 0ull;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (src[k] < 0) continue;
            const uint32_t s_span = __shfl_sync(0xffffffffu, q_span, src[k]), s_qlen = __shfl_sync(0xffffffffu, qlen, src[k]);
            const uint64_t s_y = __shfl_sync(0xffffffffu, y, src[k]);
            uint64_t s_o = __shfl_sync(0xffffffffu, o, src[k]);
            const uint32_t s_qid = (uint32_t)(s_y >> 32), s_qpos = (uint32_t)s_y;
            for (uint32_t b = 0; b < s_n[k]; b += 32) {
                const uint32_t e = b + lane;
                bool keep = e < s_n[k];
                const uint64_t r = b == 0 ? r0[k] : (keep ? __ldg(I.ty + s_start[k] + e) : 0ull);
                if (keep && scan && skip_hit(r, s_qid, s_qpos, S)) keep = false;
                const uint32_t km = __ballot_sync(0xffffffffu, keep);
                if (keep) emit(s_o + __popc(km & lt), r, s_qid, s_qpos, s_span, s_qlen);
                s_o += __popc(km);
            }
        }
    }
    if (!active || n >= COOP_MIN_HITS) return;
    const uint32_t qid = (uint32_t)(y >> 32), q_pos = (uint32_t)y;
    const uint4 *ty4 = reinterpret_cast<const uint4 *>(I.ty);
    const uint32_t p1 = (start + n + 1) >> 1;
    for (uint32_t p = start >> 1; p < p1; p += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p + u < p1 ? __ldg(ty4 + p + u) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t e = 2 * (p + u) + h;
                if (e < start || e >= start + n) continue;
                const uint64_t r = h ? ((uint64_t)v[u].w << 32 | v[u].z) : ((uint64_t)v[u].y << 32 | v[u].x);
                if (scan && skip_hit(r, qid, q_pos, S)) continue;
                emit(o, r, qid, q_pos, q_span, qlen);
                ++o;
            }
        }
    }
}

}  // namespace lrge
