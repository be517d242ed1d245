// kernels_chain_v1.cuh -- TEST-ONLY literal chaining kernel (one thread per (query, target) pair, sequential
// mg_lchain_dp / mg_chain_backtrack exactly as lchain.c orders them).  Compiled only into
// liblrge_b200_v1check.so (-DLRGE_TEST_KERNELS), which tests/test_gpu_parity.py::test_chain_kernels_agree loads to
// cross-check the product's warp-cooperative kernels.  Not part of liblrge_b200.so.
#pragma once

namespace lrge {

// mg_lchain_dp over one (target, strand) group of n anchors starting at g0. Fills f/p (p local, -1 none).
__device__ void group_dp(const uint64_t *akey, const uint32_t *aval, uint32_t g0, int32_t n, uint64_t posmask,
                         const ChainParams &P, int32_t *f, int32_t *p, int32_t *t) {
    int32_t st = 0, max_ii = -1;
    for (int32_t i = 0; i < n; ++i) {
        const int32_t ri = (int32_t)(akey[g0 + i] & posmask);
        const uint32_t vi = aval[g0 + i];
        const int32_t qi = (int32_t)(vi & 0xffffff);
        int32_t max_j = -1, max_f = (int32_t)(vi >> 24), n_skip = 0, end_j, j;
        while (st < i && ri > (int32_t)(akey[g0 + st] & posmask) + P.max_dist_x) ++st;
        if (i - st > P.max_iter) st = i - P.max_iter;
        for (j = i - 1; j >= st; --j) {
            const uint32_t vj = aval[g0 + j];
            int32_t sc = comput_sc(ri, qi, (int32_t)(akey[g0 + j] & posmask), (int32_t)(vj & 0xffffff), (int32_t)(vj >> 24), P);
            if (sc == INT32_MIN) continue;
            sc += f[g0 + j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[g0 + j] == i) {
                if (++n_skip > P.max_skip) break;
            }
            if (p[g0 + j] >= 0) t[g0 + p[g0 + j]] = i;
        }
        end_j = j;
        if (max_ii < 0 || ri - (int32_t)(akey[g0 + max_ii] & posmask) > P.max_dist_x) {
            int32_t mx = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (mx < f[g0 + j]) mx = f[g0 + j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            const uint32_t vj = aval[g0 + max_ii];
            const int32_t tmp = comput_sc(ri, qi, (int32_t)(akey[g0 + max_ii] & posmask), (int32_t)(vj & 0xffffff), (int32_t)(vj >> 24), P);
            if (tmp != INT32_MIN && max_f < tmp + f[g0 + max_ii]) max_f = tmp + f[g0 + max_ii], max_j = max_ii;
        }
        f[g0 + i] = max_f, p[g0 + i] = max_j;
        if (max_ii < 0 || (ri - (int32_t)(akey[g0 + max_ii] & posmask) <= P.max_dist_x && f[g0 + max_ii] < f[g0 + i])) max_ii = i;
    }
}

// mg_chain_backtrack over one group, visiting chain starts in the reference's order and calling
// `on_chain` for every kept chain until it returns true (stop).  t[] is reused as the visit marks.
template <typename F>
__device__ void group_backtrack(uint32_t g0, int32_t n, const ChainParams &P, const int32_t *f, const int32_t *p,
                                int32_t *t, F on_chain) {
    for (int32_t i = 0; i < n; ++i) t[g0 + i] = 0;
    int32_t last_f = INT32_MAX, last_i = n;
    for (;;) {
        // next z entry: largest (f, i) below (last_f, last_i) with f >= min_sc
        int32_t best = -1, bf = INT32_MIN;
        for (int32_t i = 0; i < n; ++i) {
            const int32_t fi = f[g0 + i];
            if (fi < P.min_sc) continue;
            if (fi > last_f || (fi == last_f && i >= last_i)) continue;
            if (fi > bf || (fi == bf && i > best)) bf = fi, best = i;
        }
        if (best < 0) break;
        last_f = bf, last_i = best;
        if (t[g0 + best] != 0) continue;
        // mg_chain_bk_end
        int32_t i = best, end_i = -1, max_i = best, max_s = 0;
        do {
            int32_t s;
            t[g0 + i] = 2;
            end_i = i = p[g0 + i];
            s = i < 0 ? bf : bf - f[g0 + i];
            if (s > max_s) max_s = s, max_i = i;
            else if (max_s - s > P.bw) break;
        } while (i >= 0 && t[g0 + i] == 0);
        for (i = best; i >= 0 && i != end_i; i = p[g0 + i]) t[g0 + i] = 0;
        end_i = max_i;
        // mark the chain
        int32_t cnt = 0, first = best;
        for (i = best; i != end_i; i = p[g0 + i]) { ++cnt; t[g0 + i] = 1; first = i; }
        const int32_t sc = i < 0 ? bf : bf - f[g0 + i];
        if (sc >= P.min_sc && cnt > 0 && cnt >= P.min_cnt) {
            ChainCoord c; c.a_first = first; c.a_last = best; c.sc = sc; c.cnt = cnt;
            if (on_chain(c)) return;
        }
    }
}

// v1: one thread per (query, target) pair, literal sequential semantics.
__global__ void __launch_bounds__(128) chain_pair_kernel(PairTask T, ChainParams P, CountOut O) {
    const uint32_t pi = blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= T.n_pairs) return;
    const uint32_t a0 = T.pstart[pi], a1 = T.pstart[pi + 1];
    const uint64_t posmask = (1ULL << T.PB) - 1;
    const uint64_t k0 = T.akey[a0];
    const uint32_t rid = (uint32_t)((k0 >> (T.PB + 1)) & ((1ULL << T.RB) - 1));
    const uint32_t qid = T.q0 + (uint32_t)(k0 >> (T.PB + 1 + T.RB));
    // split point between strands: first anchor with the strand bit set
    uint32_t lo = a0, hi = a1;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((T.akey[mid] >> T.PB) & 1) hi = mid; else lo = mid + 1; }
    const uint32_t split = lo;
    const int32_t qlen = (int32_t)O.len_q[qid], tlen = (int32_t)O.len_t[rid];
    int kept = 0, pass = 0;
    for (int strand = 0; strand < 2 && !pass; ++strand) {
        const uint32_t g0 = strand ? split : a0;
        const int32_t n = (int32_t)(strand ? a1 - split : split - a0);
        if (n < P.min_cnt) continue;
        group_dp(T.akey, T.aval, g0, n, posmask, P, T.f, T.p, T.t);
        group_backtrack(g0, n, P, T.f, T.p, T.t, [&](const ChainCoord &c) -> bool {
            kept = 1;
            if (!O.remove_internal) { pass = 1; return true; }
            // mm_reg_set_coor (hit.c) then PafRecord::is_internal (mapping.rs:59-77)
            const uint32_t vf = T.aval[g0 + c.a_first], vl = T.aval[g0 + c.a_last];
            const int32_t span0 = (int32_t)(vf >> 24);
            const int32_t x0 = (int32_t)(T.akey[g0 + c.a_first] & posmask), x1 = (int32_t)(T.akey[g0 + c.a_last] & posmask);
            const int32_t y0 = (int32_t)(vf & 0xffffff), y1 = (int32_t)(vl & 0xffffff);
            const int32_t rs = x0 + 1 > span0 ? x0 + 1 - span0 : 0, re = x1 + 1;
            int32_t qs, qe;
            if (!strand) { qs = y0 + 1 - span0; qe = y1 + 1; }
            else { qs = qlen - (y1 + 1); qe = qlen - (y0 + 1 - span0); }
            int32_t overhang, maplen;
            if (!strand) overhang = min(qs, rs) + min(qlen - qe, tlen - re);
            else overhang = min(qs, tlen - re) + min(qlen - qe, rs);
            maplen = max(qe - qs, re - rs);
            bool drop;
            if (O.mode == 2) drop = overhang > (int32_t)__fmul_rn((float)maplen, O.ratio);          // twoset.rs:514-517
            else drop = __fdiv_rn((float)overhang, (float)maplen) < O.ratio;                          // mapping.rs:75-76
            if (!drop) { pass = 1; return true; }
            return false;
        });
    }
    O.pair_res[pi] = (uint8_t)(kept | pass << 1);
    if (O.mode == 0) {
        if (kept) O.had_mapping[qid] = 1;
        if (pass && O.names_unique) atomicAdd(&O.counts[qid], 1u);
    } else if (O.mode == 1) {
        if (pass && O.rank_q[qid] != O.rank_t[rid]) { atomicAdd(&O.counts[qid], 1u); atomicAdd(&O.counts[rid], 1u); }
    } else {
        if (pass) atomicAdd(&O.counts[rid], 1u);
    }
}

// flags[i] = 1 where a new (query, target) pair starts
__global__ void pair_flag_kernel(const uint64_t *__restrict__ akey, uint64_t n, int shift, uint8_t *__restrict__ flags) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || (akey[i] >> shift) != (akey[i - 1] >> shift)) ? 1 : 0;
}


}  // namespace lrge
