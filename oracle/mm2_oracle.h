/*
 * mm2_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the overlap hot path that liblrge reaches through
 * minimap2-sys 0.1.30+minimap2.2.30 (lh3/minimap2 v2.30), plus liblrge's own
 * counting / estimator logic.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * PARITY UNPINNED: minimap2's C source is a crates.io dependency that is not
 * vendored under /root/reference, and no reference test
 * pins an overlap count.  The algorithm below is restated from the published
 * minimap2 v2.30 behaviour (sketch.c, index.c, seed.c, map.c, lchain.c, hit.c,
 * options.c, ksort.h) and anchored on liblrge's call sites:
 *   liblrge/src/minimap2/aligner.rs:53-197,204-303   (options, index, mm_map)
 *   liblrge/src/twoset.rs:286-317,477-565            (two-set counting)
 *   liblrge/src/ava.rs:271-346                       (all-vs-all counting)
 *   liblrge/src/minimap2/mapping.rs:59-77            (is_internal)
 *   liblrge/src/estimate.rs:80-157                   (median, per_read_estimate)
 * The adjacent pieces that DO have reference golden vectors (estimate.rs:159-343,
 * mapping.rs:422-492) are pinned in tests/test_oracle_golden.py.
 */
#ifndef MM2_ORACLE_H
#define MM2_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } orc128_t;

/* minimap2 flag bits used on this path (minimap.h) */
#define ORC_F_NO_DIAG    0x001
#define ORC_F_NO_DUAL    0x002
#define ORC_F_NO_LJOIN   0x400
#define ORC_F_ALL_CHAINS 0x800000

/* tie handling of the two in-place radix sorts whose order among equal keys is
 * implementation defined */
#define ORC_TIE_KSORT  0   /* literal ksort.h American-flag radix sort (unstable) */
#define ORC_TIE_STABLE 1   /* stable sort by the same key (what the CUDA path implements) */

typedef struct {
    const uint8_t  *bases;     /* concatenated ASCII bases */
    const uint64_t *off;       /* n+1 offsets into bases */
    const char     *names;     /* concatenated read names, no terminators */
    const uint32_t *name_off;  /* n+1 offsets into names */
    uint32_t n;
} orc_readset_t;

/* effective option block: mm_idxopt_t + mm_mapopt_t fields that matter on this path */
typedef struct {
    int32_t k, w, is_hpc, bucket_bits;
    int64_t flag;
    int32_t bw, bw_long, max_gap, max_gap_ref;
    int32_t max_chain_skip, max_chain_iter, min_cnt, min_chain_score;
    float   chain_gap_scale, chain_skip_scale;
    float   mid_occ_frac, q_occ_frac;
    int32_t min_mid_occ, max_mid_occ, mid_occ;
    int32_t max_occ, occ_dist, max_max_occ;
    int32_t seed;
    int32_t tie_mode;
    /* target sharding as the CUDA path does it (not minimap2): keep only hits of the contiguous target block
     * [n*shard_index/shard_count, n*(shard_index+1)/shard_count); occurrence filtering still sees the whole index.
     * shard_count <= 1 disables it. */
    int32_t shard_index, shard_count;
} orc_opt_t;

/* one mm_reg1_t as liblrge reads it (aligner.rs:244-290) */
typedef struct {
    int32_t rid, rev;
    int32_t qs, qe, rs, re;
    int32_t score, cnt, mlen, blen;
    float   div;                 /* mm_est_err (PAF dv) */
    uint32_t hash;
    int32_t rl;                  /* rep_len of the query (PAF rl) */
    int32_t order;               /* position of the row in mm_map's output for its query (score desc, hash tie-break) */
} orc_reg_t;

typedef struct {
    uint64_t n_t_minimizers, n_q_minimizers, n_keys;
    uint64_t n_seeds, n_hits, n_anchors, n_chains, n_dp_iter;
    uint64_t n_overlaps;         /* sum of per-read distinct-target counts (two-set) or unique pairs (ava) */
    uint32_t no_mapping_count;
    int32_t  mid_occ;
    double   t_index_s, t_map_s;
    uint64_t n_tie_x_queries;    /* queries holding >=2 anchors with identical x */
    /* which order-dependent branches the inputs exercised (test reporting only) */
    uint64_t n_tie_x_anchors;    /* anchors whose x equals the previous anchor's */
    uint64_t n_mz_flt_deleted;   /* query minimizers deleted by mm_seed_mz_flt */
    uint64_t n_max_iter_clamp;   /* DP rows whose window was clamped to max_chain_iter */
    uint64_t n_skip_break;       /* DP rows left through the max_chain_skip break */
    uint64_t n_occ_filtered;     /* seeds dropped because n > mid_occ */
} orc_stats_t;

typedef struct orc_index_s orc_index_t;

/* mm_set_opt(NULL) + mm_set_opt(preset) + Aligner::dual (aligner.rs:53-103). preset: "ava-ont" | "ava-pb" */
int  orc_opt_init(orc_opt_t *opt, const char *preset, int dual);

/* mm_sketch; *out is malloc'd (caller frees with orc_free) */
int  orc_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k, uint32_t rid, int32_t is_hpc,
                orc128_t **out, int64_t *n_out);
/* all minimizers of a read set in rid order (index-time sketch); per-read counts in cnt[n] (may be NULL) */
int  orc_sketch_readset(const orc_readset_t *rs, const orc_opt_t *opt, orc128_t **out, int64_t *n_out, uint32_t *cnt);

/* mm_idx_reader_read + mm_mapopt_update: builds the index and sets opt->mid_occ */
orc_index_t *orc_index_build(const orc_readset_t *T, orc_opt_t *opt, int n_threads);
void orc_index_destroy(orc_index_t *idx);
uint64_t orc_index_n_keys(const orc_index_t *idx);
uint64_t orc_index_n_minimizers(const orc_index_t *idx);
/* mm_idx_get */
const uint64_t *orc_index_get(const orc_index_t *idx, uint64_t minier, int32_t *n);

/* mm_map for one query. Optional dumps (pass NULL to skip); all outputs malloc'd.
 *   anchors: the array handed to mg_lchain_dp (after collect_seed_hits' sort)
 *   regs   : mm_gen_regs output (score-desc with hash tie-break), parent unset -> all reported */
int  orc_map(const orc_index_t *idx, const orc_opt_t *opt, const uint8_t *seq, int32_t qlen, const char *qname,
             orc_reg_t **regs, int32_t *n_regs, int32_t *rep_len,
             orc128_t **anchors, int64_t *n_anchors, orc_stats_t *st);

/* PafRecord::is_internal (mapping.rs:59-77) */
int  orc_is_internal(int32_t qlen, int32_t qs, int32_t qe, int32_t rev, int32_t tlen, int32_t ts, int32_t te, float ratio);

/* TwoSetStrategy::align_reads (twoset.rs:204-367): counts[nQ], had_mapping[nQ] */
int  orc_twoset_counts(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt,
                       int remove_internal, float max_overhang_ratio, int n_threads,
                       uint32_t *counts, uint8_t *had_mapping, orc_stats_t *st);
/* TwoSetStrategy::align_reads_inverse (twoset.rs:370-584): index = Q, stream T; counts[nQ] */
int  orc_twoset_inverse_counts(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt,
                               int remove_internal, float max_overhang_ratio, int n_threads,
                               uint32_t *counts, orc_stats_t *st);
/* AvaStrategy::align_reads (ava.rs:165-366): counts[n] */
int  orc_ava_counts(const orc_readset_t *R, orc_opt_t *opt, int remove_internal, float max_overhang_ratio,
                    int n_threads, uint32_t *counts, orc_stats_t *st);

/* all regs of every query, flattened (PAF-level parity). out arrays malloc'd; qidx[i] = query of regs[i] */
int  orc_twoset_regs(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt, int n_threads,
                     orc_reg_t **regs, uint32_t **qidx, int64_t *n_regs);

/* estimate.rs */
float orc_per_read_estimate(uint64_t read_len, float avg_target_len, uint64_t n_target_reads, uint64_t n_ovlaps, uint32_t ovlap_thresh);
/* median(): returns bitmask of which outputs are Some: bit0 lower, bit1 median, bit2 upper. lower_q/upper_q < 0 => None */
int   orc_median(const float *vals, int64_t n, float lower_q, float upper_q, float *lower, float *median, float *upper);
/* Estimate::estimate on a vector of per-read estimates */
int   orc_estimate(const float *estimates, int64_t n, int finite, float lower_q, float upper_q,
                   float *lower, float *median, float *upper);

/* chain gap penalty table: pen[dd] = (int)(chn_pen_gap*dd + .5f*mg_log2(dd+1)) for dd in [0,n) (lchain.c comput_sc) */
void  orc_gap_penalty_table(int32_t k, float chain_gap_scale, int32_t n, int32_t *pen);

void  orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
