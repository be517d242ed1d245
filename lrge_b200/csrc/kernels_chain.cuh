// kernels_chain.cuh -- K6: chaining DP, backtrack and per-(query,target) overlap decision.
//
// Replaces minimap2 mg_lchain_dp / comput_sc / mg_chain_backtrack / mg_chain_bk_end (lchain.c),
// the coordinate part of mm_gen_regs / mm_reg_set_coor (hit.c), and liblrge's dedupe + count
// (twoset.rs:286-302, ava.rs:271-306, mapping.rs:59-77), SURVEY.md rows a10-a14.
//
// Unit of work: one (query, target) PAIR = its forward-strand group followed by its reverse-strand
// group in the sorted anchor array.  The DP of minimap2 never links anchors of different
// (target, strand) groups, and its backtrack never walks across groups, so each group is chained on
// its own with fresh `st` / `max_ii` state -- identical to what the sequential pass over the whole
// per-query array computes for that group.
//
// Order-dependent details that are kept as they are in lchain.c:
//   * predecessors are visited from i-1 downwards; ties in score keep the first (nearest) one;
//   * the max_skip counter (t[] marks of the predecessors' predecessors) and its early break;
//   * the max_iter clamp of the window and the max_ii rescue of a far best predecessor;
//   * chains are extracted from the highest f downwards; equal f: higher index first (what a stable
//     ascending sort read backwards gives; ksort's radix sort is unstable);
//   * mg_chain_bk_end's drop-trimming with max_drop = bw.
// The gap penalty (int)(chn_pen_gap*dd + .5f*mg_log2(dd+1)) depends only on dd <= bw; it is tabulated
// on the host in f32 with the reference's evaluation order (no FMA) and read from `pen`.
#pragma once
#include "common.cuh"

namespace lrge {

struct ChainParams {
    int max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc;
    const int32_t *pen;          // bw+1 entries
    int span_bound_ok;           // every anchor span is the constant k and k * (min_cnt-1) < min_sc (no HPC)
};

struct PairTask {
    const uint64_t *akey;        // sorted anchor keys
    const uint32_t *aval;        // payload: span<<24 | qpos'
    const uint32_t *pstart;      // n_pairs+1 starts into the anchor arrays (last = n_anchors)
    uint32_t n_pairs;
    int PB, RB;                  // key layout
    uint32_t q0;                 // first query of the batch
    int32_t *f, *p, *t;          // per-anchor scratch (batch sized); t zeroed before launch
};

// one kept chain = one mm_reg1_t as liblrge reads it (aligner.rs:244-290), plus the chain's end anchors: the host
// derives the PAF tags dv (mm_est_err) and rl, and the row order of mm_gen_regs, from them and the query's sketch
struct RegOut { uint32_t qid, rid; int32_t rev, qs, qe, rs, re, score, cnt, mlen, blen; uint32_t a0x, a0y, a1y; };

struct CountOut {
    RegOut *regs;                // optional: every kept chain (full DP + full backtrack, no counting shortcut)
    unsigned long long *n_regs;
    uint64_t regs_cap;
    uint32_t *counts;            // two-set: per query; ava: per read
    uint8_t  *had_mapping;       // two-set: any kept chain (before -F)
    uint8_t  *pair_res;          // per pair: bit0 kept chain, bit1 kept chain passing -F
    const uint32_t *len_q, *len_t;
    const uint32_t *rank_q, *rank_t;
    int mode;                    // 0 two-set, 1 ava, 2 two-set inverse (count indexed read)
    int remove_internal;
    float ratio;
    int names_unique;            // target names are unique: device-side counting is exact
};

struct ChainCoord { int32_t a_first, a_last, sc, cnt; };   // local anchor indices of the chain ends

// Sorted anchors of one group as the DP reads them: [i] -> (x = target position, y = q_span << 24 | query position').
// AncAoS: compact uint2 records in HBM (global path, after group_mark_kernel).
// AncSoA: the shared-memory sort buffers of the fused per-query kernel (key array + payload array), used in place.
struct AncAoS {
    const uint2 *p;
    __device__ __forceinline__ uint2 operator[](int32_t i) const { return p[i]; }
};
template <typename KT> struct AncSoA {
    const KT *k; const uint32_t *v; KT posmask;
    __device__ __forceinline__ uint2 operator[](int32_t i) const { return make_uint2((uint32_t)(k[i] & posmask), v[i]); }
};

__device__ __forceinline__ int32_t comput_sc(int32_t ri, int32_t qi, int32_t rj, int32_t qj, int32_t span_j,
                                             const ChainParams &P) {
    const int32_t dq = qi - qj;
    if (dq <= 0 || dq > P.max_dist_x) return INT32_MIN;
    const int32_t dr = ri - rj;
    if (dr == 0 || dq > P.max_dist_y) return INT32_MIN;
    const int32_t dd = dr > dq ? dr - dq : dq - dr;
    if (dd > P.bw) return INT32_MIN;
    const int32_t dg = dr < dq ? dr : dq;
    int32_t sc = span_j < dg ? span_j : dg;
    if (dd || dg > span_j) sc -= P.pen[dd];
    return sc;
}

// =================================================================================================
// v2: one WARP per (query, target, strand) group; anchors i are processed in order, the predecessors
// j = i-1, i-2, ... of one anchor are scored 32 at a time (lane l takes j = base - l, i.e. lane order =
// the serial visiting order of mg_lchain_dp).  The three order-dependent pieces of the inner loop are
// resolved with warp scans so that the result equals the sequential loop exactly:
//   * "sc > max_f"            -> exclusive prefix-max over lanes (plus the running max_f of earlier chunks);
//   * t[p[j]] = i marks       -> every scored lane stores its mark, __syncwarp, then lanes read t[j]:
//                                 a mark made by lane l targets p[j] < j, so only lanes > l can see it,
//                                 which is the serial order; marks made past the break are never read
//                                 again for this i and are stale for every later i;
//   * n_skip (+1 on a marked non-improving lane, -1 floored at 0 on an improving lane, break when > max_skip)
//                              -> a clamped walk: n_l = N_l - min(0, min_{k<=l} N_k) with N = unclamped
//                                 prefix sum; the first lane where a marked step reaches > max_skip is
//                                 the break lane, later lanes are discarded.
// =================================================================================================

constexpr int CG_THREADS = 256;

// One (query, target, strand) group of the sorted anchor array that can hold a chain.  Written by the per-query sort
// kernel (kernels_map.cuh) or, on the global path, by group_classify_kernel.
struct GroupDesc {
    uint32_t start, n;           // anchors [start, start + n) of the batch's sorted anchor array
    uint32_t qid, rid;           // streamed read, indexed read
    uint32_t strand;
    uint32_t pair;               // slot shared by the two strand groups of one (query, target) pair
};

inline __host__ __device__ int group_class(uint32_t n) { return n >= 2048 ? 0 : n >= 256 ? 1 : n >= 48 ? 2 : 3; }

struct GroupTask {
    const uint2 *anc;            // compact sorted anchors: x = target pos, y = span << 24 | qpos'
    const GroupDesc *gl[4];      // 4 size classes, largest first
    const uint32_t *gcount;      // entries per class
    uint32_t *work_ctr;
    int2 *fp;                    // per anchor (f, p)
    int32_t *t;                  // per anchor marks, zeroed before launch
    uint32_t *pflag;             // per pair slot: "already counted"
    uint2 *dd_out;               // duplicate target names (host-side HashSet<name> dedupe): passing (query, target) pairs
    unsigned long long *dd_n;
};

__device__ __forceinline__ int32_t warp_prefix_max_incl(int32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v = max(v, u);
    }
    return v;
}
__device__ __forceinline__ int32_t warp_prefix_min_incl(int32_t v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane >= o) v = min(v, u);
    }
    return v;
}

__device__ __forceinline__ int32_t comput_sc_smem(int32_t ri, int32_t qi, int32_t rj, int32_t qj, int32_t span_j,
                                                  const ChainParams &P, const uint8_t *s_pen) {
    const int32_t dq = qi - qj;
    if (dq <= 0 || dq > P.max_dist_x) return INT32_MIN;
    const int32_t dr = ri - rj;
    if (dr == 0 || dq > P.max_dist_y) return INT32_MIN;
    const int32_t dd = dr > dq ? dr - dq : dq - dr;
    if (dd > P.bw) return INT32_MIN;
    const int32_t dg = dr < dq ? dr : dq;
    int32_t sc = span_j < dg ? span_j : dg;
    if (dd || dg > span_j) sc -= (int32_t)s_pen[dd];
    return sc;
}

// mg_lchain_dp for one group, warp-cooperative.  Returns the first anchor whose score reaches stop_sc (the DP is
// causal, so f[0..i] are final at that point), or -1 after the whole group.
template <typename A>
__device__ int32_t warp_group_dp(const A anc, int32_t n, const ChainParams &P, const uint8_t *s_pen,
                                 int2 *fp, int32_t *t, uint32_t lane, int32_t stop_sc) {
    int32_t st = 0, r_st = n > 0 ? (int32_t)anc[0].x : 0, max_ii = -1, r_maxii = 0, f_maxii = 0;
    const uint32_t le_mask = lane == 31 ? 0xffffffffu : ((2u << lane) - 1);
    for (int32_t i = 0; i < n; ++i) {
        const uint2 ai = anc[i];
        const int32_t ri = (int32_t)ai.x, qi = (int32_t)(ai.y & 0xffffff);
        int32_t max_f = (int32_t)(ai.y >> 24), max_j = -1, n_skip = 0;
        while (st < i && ri > r_st + P.max_dist_x) r_st = (int32_t)anc[++st].x;
        if (i - st > P.max_iter) { st = i - P.max_iter; r_st = (int32_t)anc[st].x; }
        int32_t end_j = st - 1;
        for (int32_t base = i - 1; base >= st; base -= 32) {
            const int32_t j = base - (int32_t)lane;
            int32_t s = INT32_MIN, pj = -1;
            if (j >= st) {
                const uint2 aj = anc[j];
                s = comput_sc_smem(ri, qi, (int32_t)aj.x, (int32_t)(aj.y & 0xffffff), (int32_t)(aj.y >> 24), P, s_pen);
                if (s != INT32_MIN) {
                    const int2 v = fp[j];
                    s += v.x;
                    pj = v.y;
                }
            }
            const bool scored = s != INT32_MIN;
            if (pj >= 0) t[pj] = i;
            __syncwarp();
            const int32_t tj = scored ? t[j] : -1;
            // improved lanes: s exceeds max_f and every earlier lane's s.  Candidates are the lanes above the
            // incoming max_f; walk them in lane order (usually one).
            uint32_t I = 0;
            {
                uint32_t C = __ballot_sync(0xffffffffu, scored && s > max_f);
                while (C) {
                    const int c = __ffs(C) - 1;
                    I |= 1u << c;
                    const int32_t m = __shfl_sync(0xffffffffu, s, c);
                    C = __ballot_sync(0xffffffffu, scored && s > m) & ~((2u << c) - 1u) & (c == 31 ? 0u : 0xffffffffu);
                }
            }
            const bool improved = (I >> lane) & 1;
            const bool marked = scored && !improved && tj == i;
            const uint32_t M = __ballot_sync(0xffffffffu, marked);
            uint32_t keep = 0xffffffffu;
            int32_t brk_lane = -1;
            if (M) {
                int32_t nl;
                if (I == 0 || (31 - __clz(I)) < (__ffs(M) - 1)) {
                    // every improvement precedes every mark: decrements first (floored), then pure increments
                    const int32_t n1 = max(0, n_skip - __popc(I));
                    nl = n1 + __popc(M & le_mask);
                    n_skip = n1 + __popc(M);
                } else {
                    const int32_t N = n_skip + __popc(M & le_mask) - __popc(I & le_mask);
                    const int32_t mn = min(0, warp_prefix_min_incl(N, lane));
                    nl = N - mn;
                    n_skip = __shfl_sync(0xffffffffu, nl, 31);
                }
                const uint32_t B = __ballot_sync(0xffffffffu, marked && nl > P.max_skip);
                if (B) { brk_lane = __ffs(B) - 1; keep = (1u << brk_lane) - 1; }
            } else {
                n_skip = max(0, n_skip - __popc(I));
            }
            const uint32_t Ik = I & keep;
            if (Ik) {
                const int top = 31 - __clz(Ik);
                max_f = __shfl_sync(0xffffffffu, s, top);
                max_j = base - top;
            }
            if (brk_lane >= 0) { end_j = base - brk_lane; break; }
        }
        // rescue of a far best predecessor (max_ii)
        if (max_ii < 0 || ri - r_maxii > P.max_dist_x) {
            int32_t bf = INT32_MIN, bj = -1;
            for (int32_t base = i - 1; base >= st; base -= 32) {
                const int32_t j = base - (int32_t)lane;
                if (j >= st) { const int32_t fj = fp[j].x; if (fj > bf) bf = fj, bj = j; }   // descending j per lane: first max = largest j
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const int32_t of = __shfl_xor_sync(0xffffffffu, bf, o), oj = __shfl_xor_sync(0xffffffffu, bj, o);
                if (oj >= 0 && (bj < 0 || of > bf || (of == bf && oj > bj))) bf = of, bj = oj;
            }
            max_ii = bj;
            if (bj >= 0) { r_maxii = (int32_t)anc[bj].x; f_maxii = bf; }
        }
        if (max_ii >= 0 && max_ii < end_j) {
            const uint2 am = anc[max_ii];
            const int32_t tmp = comput_sc_smem(ri, qi, (int32_t)am.x, (int32_t)(am.y & 0xffffff), (int32_t)(am.y >> 24), P, s_pen);
            if (tmp != INT32_MIN && max_f < tmp + f_maxii) max_f = tmp + f_maxii, max_j = max_ii;
        }
        if (lane == 0) fp[i] = make_int2(max_f, max_j);
        __syncwarp();
        if (max_f >= stop_sc) return i;
        if (max_ii < 0 || (ri - r_maxii <= P.max_dist_x && f_maxii < max_f)) { max_ii = i; r_maxii = ri; f_maxii = max_f; }
    }
    return -1;
}

// Warp argmax of (f, index) over the group among entries with f >= min_sc, strictly below (last_f, last_i) in
// (f, index) order and, if use_t, not yet used (t == 0).
__device__ __forceinline__ void warp_next_start(int32_t n, const ChainParams &P, const int2 *fp, const int32_t *t, bool use_t,
                                                int32_t last_f, int32_t last_i, uint32_t lane, int32_t &bf, int32_t &bi) {
    bf = INT32_MIN, bi = -1;
    for (int32_t i = lane; i < n; i += 32) {
        const int32_t fi = fp[i].x;
        if (fi < P.min_sc) continue;
        if (fi > last_f || (fi == last_f && i >= last_i)) continue;
        if (use_t && t[i] != 0) continue;
        if (fi > bf || (fi == bf && i > bi)) bf = fi, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int32_t of = __shfl_xor_sync(0xffffffffu, bf, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || of > bf || (of == bf && oi > bi))) bf = of, bi = oi;
    }
}

// mg_chain_backtrack for one group: chain starts are visited from the highest (f, index) down; the search for
// the next start is warp-parallel, the walk itself is done by lane 0.  Calls on_chain (all lanes, uniform
// arguments) for every kept chain until it returns true.
// The first start is walked without touching t[] (nothing is used yet, so mg_chain_bk_end's "t[i] == 0" test is
// always true); if its chain settles the question -- the common case for counting -- no mark is ever written.
template <typename F>
__device__ void warp_group_backtrack(int32_t n, const ChainParams &P, const int2 *fp, int32_t *t, uint32_t lane, F on_chain) {
    int32_t bf, bi;
    warp_next_start(n, P, fp, t, false, INT32_MAX, n, lane, bf, bi);
    if (bi < 0) return;
    int32_t cnt = 0, first = bi, sc = 0, end_i = -1;
    if (lane == 0) {
        // mg_chain_bk_end on untouched marks: find max_i, counting the anchors above it on the way
        int32_t i = bi, max_i = bi, max_s = 0, steps = 0, prev = bi;
        do {
            const int32_t cur = i;
            i = fp[i].y;
            ++steps;
            const int32_t s = i < 0 ? bf : bf - fp[i].x;
            if (s > max_s) { max_s = s; max_i = i; cnt = steps; first = cur; }
            else if (max_s - s > P.bw) break;
            prev = cur;
        } while (i >= 0);
        (void)prev;
        end_i = max_i;
        sc = end_i < 0 ? bf : bf - fp[end_i].x;
        if (end_i == bi) { cnt = 0; sc = 0; }
    }
    cnt = __shfl_sync(0xffffffffu, cnt, 0); first = __shfl_sync(0xffffffffu, first, 0);
    sc = __shfl_sync(0xffffffffu, sc, 0); end_i = __shfl_sync(0xffffffffu, end_i, 0);
    if (sc >= P.min_sc && cnt > 0 && cnt >= P.min_cnt) {
        ChainCoord c; c.a_first = first; c.a_last = bi; c.sc = sc; c.cnt = cnt;
        if (on_chain(c)) return;
    }
    // general path: marks are needed from here on
    for (int32_t i = lane; i < n; i += 32) t[i] = 0;
    __syncwarp();
    if (lane == 0) for (int32_t i = bi; i != end_i; i = fp[i].y) t[i] = 1;
    __syncwarp();
    int32_t last_f = bf, last_i = bi;
    for (;;) {
        warp_next_start(n, P, fp, t, true, last_f, last_i, lane, bf, bi);
        if (bi < 0) break;
        last_f = bf, last_i = bi;
        cnt = 0, first = bi, sc = 0;
        if (lane == 0) {
            // mg_chain_bk_end
            int32_t i = bi, stop_i = -1, max_i = bi, max_s = 0;
            do {
                int32_t s;
                t[i] = 2;
                stop_i = i = fp[i].y;
                s = i < 0 ? bf : bf - fp[i].x;
                if (s > max_s) max_s = s, max_i = i;
                else if (max_s - s > P.bw) break;
            } while (i >= 0 && t[i] == 0);
            // one pass: anchors above max_i join the chain (t = 1), the rest of the walked path is released
            bool in_chain = max_i != bi;
            for (i = bi; i >= 0 && i != stop_i; i = fp[i].y) {
                if (i == max_i) in_chain = false;
                if (in_chain) { ++cnt; t[i] = 1; first = i; } else t[i] = 0;
            }
            sc = max_i == bi ? 0 : (max_i < 0 ? bf : bf - fp[max_i].x);
        }
        __syncwarp();
        cnt = __shfl_sync(0xffffffffu, cnt, 0); first = __shfl_sync(0xffffffffu, first, 0); sc = __shfl_sync(0xffffffffu, sc, 0);
        if (sc >= P.min_sc && cnt > 0 && cnt >= P.min_cnt) {
            ChainCoord c; c.a_first = first; c.a_last = bi; c.sc = sc; c.cnt = cnt;
            if (on_chain(c)) return;
        }
    }
}

// One (query, target, strand) group: DP (+ backtrack where needed) and the overlap decision.  All lanes of the warp
// call it with uniform arguments; kept = the group yields a kept chain, pass = one of them also survives -F.
template <typename A>
__device__ void chain_one_group(const A anc, int32_t n, int2 *fp, int32_t *t, const ChainParams &P, const uint8_t *s_pen,
                                const CountOut &O, uint32_t lane, uint32_t qid, uint32_t rid, int strand, int &kept, int &pass) {
    const int32_t qlen = (int32_t)O.len_q[qid], tlen = (int32_t)O.len_t[rid];
    kept = 0; pass = 0;
    bool full = true;
    if (!O.remove_internal && !O.regs) {
        // Counting only needs "does this group yield a kept chain".  The chain extracted first starts at the group's
        // highest f; nothing is used yet and no predecessor scores above the start, so mg_chain_bk_end keeps
        // sc >= min(f_max, max_drop) -- and with every span <= (min_sc-1)/(min_cnt-1) a score >= min_sc needs
        // >= min_cnt anchors.  Hence: kept chain exists <=> some f[i] >= min_sc, and the DP can stop there.
        const int32_t hit = warp_group_dp(anc, n, P, s_pen, fp, t, lane, P.min_sc);
        full = false;
        if (hit >= 0) {
            int32_t mspan = 0;
            if (!P.span_bound_ok) {
                for (int32_t i = lane; i < n; i += 32) mspan = max(mspan, (int32_t)(anc[i].y >> 24));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mspan = max(mspan, __shfl_xor_sync(0xffffffffu, mspan, o));
            }
            if (P.span_bound_ok || (mspan * (P.min_cnt - 1) < P.min_sc && P.bw >= P.min_sc)) kept = pass = 1;
            else {                                                  // HPC group with a very long span: literal path
                for (int32_t i = lane; i < n; i += 32) t[i] = 0;
                __syncwarp();
                full = true;
            }
        }
    }
    if (!full) return;
    warp_group_dp(anc, n, P, s_pen, fp, t, lane, INT32_MAX);
    __syncwarp();
    warp_group_backtrack(n, P, fp, t, lane, [&](const ChainCoord &c) -> bool {
        kept = 1;
        if (!O.remove_internal && !O.regs) { pass = 1; return true; }
        // mm_reg_set_coor (hit.c)
        const uint2 af = anc[c.a_first], al = anc[c.a_last];
        const int32_t span0 = (int32_t)(af.y >> 24);
        const int32_t x0 = (int32_t)af.x, x1 = (int32_t)al.x;
        const int32_t y0 = (int32_t)(af.y & 0xffffff), y1 = (int32_t)(al.y & 0xffffff);
        const int32_t rs = x0 + 1 > span0 ? x0 + 1 - span0 : 0, re = x1 + 1;
        int32_t qs, qe;
        if (!strand) { qs = y0 + 1 - span0; qe = y1 + 1; }
        else { qs = qlen - (y1 + 1); qe = qlen - (y0 + 1 - span0); }
        bool drop = false;
        if (O.remove_internal) {                                // PafRecord::is_internal / the inverse path's inline test
            int32_t overhang;
            if (!strand) overhang = min(qs, rs) + min(qlen - qe, tlen - re);
            else overhang = min(qs, tlen - re) + min(qlen - qe, rs);
            const int32_t maplen = max(qe - qs, re - rs);
            if (O.mode == 2) drop = overhang > (int32_t)__fmul_rn((float)maplen, O.ratio);
            else drop = __fdiv_rn((float)overhang, (float)maplen) < O.ratio;
        }
        if (!drop) pass = 1;
        if (O.regs) {
            if (lane == 0) {
                // mm_cal_fuzzy_len (hit.c): walk the chain once more, last anchor -> first
                int32_t mlen = span0, blen = span0;
                for (int32_t i = c.a_last; i != c.a_first;) {
                    const int32_t pi = fp[i].y;
                    const uint2 a1 = anc[i], a0 = anc[pi];
                    const int32_t span = (int32_t)(a1.y >> 24);
                    const int32_t tl = (int32_t)a1.x - (int32_t)a0.x, ql = (int32_t)(a1.y & 0xffffff) - (int32_t)(a0.y & 0xffffff);
                    blen += tl > ql ? tl : ql;
                    mlen += (tl > span && ql > span) ? span : (tl < ql ? tl : ql);
                    i = pi;
                }
                const unsigned long long slot = atomicAdd(O.n_regs, 1ull);
                if (slot < O.regs_cap) {
                    RegOut r; r.qid = qid; r.rid = rid; r.rev = strand; r.qs = qs; r.qe = qe; r.rs = rs; r.re = re;
                    r.score = c.sc; r.cnt = c.cnt; r.mlen = mlen; r.blen = blen;
                    r.a0x = af.x; r.a0y = af.y; r.a1y = al.y;
                    O.regs[slot] = r;
                }
            }
            return false;                                       // every chain of the group is reported
        }
        return !drop;
    });
}

// Global path (queries whose anchors do not fit the fused kernel's shared-memory rounds): one warp per group of the
// globally sorted anchor array, persistent CTAs pulling groups largest-first from an atomic work counter.
__global__ void __launch_bounds__(CG_THREADS) chain_group_kernel(GroupTask T, ChainParams P, CountOut O) {
    __shared__ uint8_t s_pen[2048];
    for (int i = threadIdx.x; i <= P.bw && i < 2048; i += blockDim.x) s_pen[i] = (uint8_t)P.pen[i];
    __syncthreads();
    const uint32_t lane = lane_id();
    const uint32_t c0 = T.gcount[0], c1 = c0 + T.gcount[1], c2 = c1 + T.gcount[2], total = c2 + T.gcount[3];
    for (;;) {
        uint32_t w = 0;
        if (lane == 0) w = atomicAdd(T.work_ctr, 1u);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= total) break;
        const GroupDesc *dp = w < c0 ? T.gl[0] + w : w < c1 ? T.gl[1] + (w - c0) : w < c2 ? T.gl[2] + (w - c1) : T.gl[3] + (w - c2);
        const GroupDesc d = *dp;
        int kept = 0, pass = 0;
        chain_one_group(AncAoS{T.anc + d.start}, (int32_t)d.n, T.fp + d.start, T.t + d.start, P, s_pen, O, lane, d.qid, d.rid,
                        (int)d.strand, kept, pass);
        if (lane == 0 && kept) {
            if (O.mode == 0) O.had_mapping[d.qid] = 1;
            if (pass && atomicExch(&T.pflag[d.pair], 1u) == 0u) {
                if (O.mode == 0) {
                    if (O.names_unique) atomicAdd(&O.counts[d.qid], 1u);
                    else T.dd_out[atomicAdd(T.dd_n, 1ull)] = make_uint2(d.qid, d.rid);
                } else if (O.mode == 1) {
                    if (O.rank_q[d.qid] != O.rank_t[d.rid]) { atomicAdd(&O.counts[d.qid], 1u); atomicAdd(&O.counts[d.rid], 1u); }
                } else atomicAdd(&O.counts[d.rid], 1u);
            }
        }
    }
}

// K5b: sorted (key, payload) -> compact anchors + group-start flags
__global__ void group_mark_kernel(const uint64_t *__restrict__ akey, const uint32_t *__restrict__ aval, uint64_t n, int PB,
                                  uint2 *__restrict__ anc, uint8_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = akey[i];
    anc[i] = make_uint2((uint32_t)(k & ((1ULL << PB) - 1)), aval[i]);
    flags[i] = (i == 0 || (k >> PB) != (akey[i - 1] >> PB)) ? 1 : 0;
}

// global path: groups that can hold a chain (n >= min_cnt and n * max_span >= min_sc: f <= sum of spans) -> descriptors
// in 4 size classes.  The pair slot is the index of the pair's forward group (or the group itself).
struct GroupLists { GroupDesc *gl[4]; uint32_t cap[4]; };
__global__ void group_classify_kernel(const uint64_t *__restrict__ akey, const uint32_t *__restrict__ gstart, uint32_t n_groups,
                                      int PB, int RB, uint32_t q0, const uint32_t *__restrict__ qmap, int min_cnt, int min_sc,
                                      int max_span, GroupLists L, uint32_t *__restrict__ gcount) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint32_t g0 = gstart[g], n = gstart[g + 1] - g0;
    if ((int)n < min_cnt || (int64_t)n * max_span < min_sc) return;
    const uint64_t k0 = akey[g0];
    GroupDesc d;
    d.start = g0; d.n = n;
    d.strand = (uint32_t)((k0 >> PB) & 1);
    d.rid = (uint32_t)((k0 >> (PB + 1)) & ((1ULL << RB) - 1));
    const uint32_t ql = (uint32_t)(k0 >> (PB + 1 + RB));                 // local query id in the batch
    d.qid = qmap ? qmap[ql] : q0 + ql;
    d.pair = g;
    if (d.strand && g > 0 && (akey[gstart[g - 1]] >> (PB + 1)) == (k0 >> (PB + 1))) d.pair = g - 1;
    const int c = group_class(n);
    const uint32_t slot = atomicAdd(&gcount[c], 1u);
    if (slot < L.cap[c]) L.gl[c][slot] = d;
}

}  // namespace lrge
