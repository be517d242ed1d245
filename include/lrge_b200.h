/*
 * lrge_b200.h -- C-ABI of the B200-native overlap hot path (liblrge_b200.so).
 *
 * Drop-in boundary for liblrge's hot path.  liblrge reaches minimap2 through
 * per-read FFI calls (mm_set_opt, mm_idx_reader_open/read/close, mm_mapopt_update, mm_map:
 * liblrge/src/minimap2/aligner.rs:53-303, thread_buf.rs:20-34).  A per-read call cannot feed a GPU, so
 * the boundary moves up one level to what the strategies consume -- everything between "target.fa /
 * query.fa written" and "Vec<f32> estimates returned":
 *
 *   lrge_gpu_twoset_counts   replaces AlignerWrapper::new + TwoSetStrategy::align_reads
 *                            (twoset.rs:602-603 -> aligner.rs:311-328, twoset.rs:204-367)
 *   lrge_gpu_twoset_counts (use_min_ref=1) ... align_reads_inverse (twoset.rs:596-599, 370-584)
 *   lrge_gpu_ava_counts      replaces AlignerWrapper::new + AvaStrategy::align_reads
 *                            (ava.rs:378-380 -> ava.rs:165-366)
 *   lrge_per_read_estimate   estimate.rs:142-157          lrge_estimate   estimate.rs:55-132
 *
 * Plain pointers and sizes only; the caller owns every input and output buffer, the library owns
 * device memory inside the context; nothing allocated on one side is freed on the other.  Every
 * function returns an lrge_status (0 = ok) and never throws or aborts across the boundary; the message
 * for the last failure on a context is available from lrge_gpu_last_error.  There is no CPU fallback:
 * without a CUDA device lrge_gpu_create fails with LRGE_ERR_CUDA.
 */
#ifndef LRGE_B200_H
#define LRGE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors LrgeError (liblrge/src/error.rs:6-33) + device-side failures */
typedef enum lrge_status {
    LRGE_OK = 0,
    LRGE_ERR_IO = 1,                /* IoError */
    LRGE_ERR_FASTQ_PARSE = 2,       /* FastqParseError */
    LRGE_ERR_TOO_MANY_READS = 3,    /* TooManyReadsError */
    LRGE_ERR_TOO_FEW_READS = 4,     /* TooFewReadsError */
    LRGE_ERR_INVALID_PLATFORM = 5,  /* InvalidPlatform */
    LRGE_ERR_THREAD = 6,            /* ThreadError */
    LRGE_ERR_PAF_WRITE = 7,         /* PafWriteError */
    LRGE_ERR_MAP = 8,               /* MapError, e.g. "Sequence is empty" (aligner.rs:214-216) */
    LRGE_ERR_DUPLICATE_READ_ID = 9, /* DuplicateReadIdentifier (ava.rs:195-199, twoset.rs:439-449) */
    LRGE_ERR_CUDA = 100,
    LRGE_ERR_NCCL = 101,
    LRGE_ERR_INVALID_ARG = 102,
    LRGE_ERR_OUT_OF_MEMORY = 103,
    LRGE_ERR_UNSUPPORTED = 104
} lrge_status;

/* Platform (liblrge/src/lib.rs:163-180): Nanopore -> ava-ont, PacBio -> ava-pb (twoset.rs:591-594) */
typedef enum lrge_platform { LRGE_PLATFORM_NANOPORE = 0, LRGE_PLATFORM_PACBIO = 1 } lrge_platform;

/* A read set = the temp FASTA liblrge hands to minimap2 (twoset.rs:157-186): read order is rid order;
 * names are the header tokens up to the first whitespace (io.rs:192-205). */
typedef struct lrge_readset {
    const uint8_t  *bases;     /* concatenated ASCII bases (ACGTUN, either case; anything else is ambiguous) */
    const uint64_t *off;       /* n+1 offsets into bases */
    const char     *names;     /* concatenated names, no terminators */
    const uint32_t *name_off;  /* n+1 offsets into names */
    uint32_t n;
} lrge_readset;

/* Builder fields that reach the hot path (twoset/builder.rs:22-39, ava/builder.rs:18-32) */
typedef struct lrge_opts {
    int32_t platform;            /* lrge_platform; default Nanopore */
    int32_t remove_internal;     /* -F: drop overlaps that PafRecord::is_internal (mapping.rs:59-77); default 0 */
    float   max_overhang_ratio;  /* default 0.2 */
    int32_t use_min_ref;         /* two-set: index the smaller set (twoset.rs:596-599); default 0 */
    /* target sharding (one process per GPU): this context keeps anchors only for its contiguous block of targets,
     * rid in [n*shard_index/shard_count, n*(shard_index+1)/shard_count); seed filtering still uses global
     * occurrence counts. */
    int32_t shard_index;         /* default 0 */
    int32_t shard_count;         /* default 1 */
    int32_t reserved[6];
} lrge_opts;

typedef struct lrge_stats {
    uint64_t n_target_bases, n_query_bases;
    uint64_t n_target_minimizers, n_query_minimizers, n_keys;
    uint64_t n_seeds;            /* query minimizers that survive mm_seed_mz_flt and the mid_occ filter */
    uint64_t n_hits;             /* hit-list entries read */
    uint64_t n_anchors;          /* anchors kept after skip_seed (and sharding) */
    uint64_t n_pairs;            /* (query, target) pairs holding anchors */
    uint64_t n_overlaps;         /* pairs with at least one kept chain that passes the -F filter */
    uint32_t no_mapping_count;   /* twoset.rs:303-309 / ava.rs:329-330 semantics */
    int32_t  mid_occ;
    uint32_t n_batches;          /* query batches (bounded by anchor memory) */
    uint32_t n_kernel_launches;
    /* device time per stage in ms (CUDA events on the context's stream) */
    float ms_h2d, ms_pack, ms_sketch, ms_index, ms_seed, ms_sort, ms_chain, ms_d2h, ms_total;
    float ms_chain_kernel;       /* the chaining kernel alone, summed over batches (0 unless lrge_gpu_set_profiling is on) */
    float ms_comm;               /* NCCL reduction of the per-read counts */
    uint32_t n_deferred;         /* queries the per-query shared-memory sort handed to the device-wide sort */
    uint32_t n_rounds;           /* shared-memory rounds of the per-query sort (>= queries holding anchors) */
    uint32_t n_index_runs;       /* sorted runs the index was merged from: 1, or 2 when the host-buffer path sorted per half */
    float reserved[2];
} lrge_stats;

typedef struct lrge_ctx lrge_ctx;
typedef struct lrge_staged lrge_staged;   /* a read set resident in HBM (ASCII + offsets), owned by the context */

void lrge_opts_default(lrge_opts *opts);

int  lrge_gpu_create(int device, lrge_ctx **out);
void lrge_gpu_destroy(lrge_ctx *ctx);
const char *lrge_gpu_last_error(const lrge_ctx *ctx);   /* ctx may be NULL: last error of lrge_gpu_create */

/* Host buffers in, per-read overlap counts out (host<->device copies inside).
 *   counts[Q->n]      number of distinct target reads overlapping each query (twoset.rs:286-302);
 *                     with use_min_ref and more target than query bases the index is built on Q and T is
 *                     streamed (twoset.rs:370-584) -- counts are still per query read
 *   had_mapping[Q->n] 1 if mm_map returned any region for the read (twoset.rs:288); may be NULL */
int  lrge_gpu_twoset_counts(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                            uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats);
/* The same for target sets that do not fit one host: every rank of the communicator (lrge_gpu_comm_init) passes ONLY ITS
 * block of the target reads (any partition; blocks need not be equal) and ALL query reads; opts->shard_index / shard_count
 * = rank / world; use_min_ref must be 0.  The minimizer occurrence counts and mid_occ are made global by the index
 * exchange, so every rank returns the complete per-query counts -- identical to one lrge_gpu_twoset_counts call on the
 * concatenated target set.  Name rules (twoset.rs:286-302 dedupe of equal target names, equal query/target names) see the
 * names of the local block only: target names must be unique across blocks.  This is how README.md:173-175's
 * 2 000 000-target human read set is meant to run (8 ranks x 250 000 targets). */
int  lrge_gpu_twoset_counts_shard(lrge_ctx *ctx, const lrge_readset *T_block, const lrge_readset *Q, const lrge_opts *opts,
                                  uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats);
/* counts[R->n]: unique overlapping partners per read (ava.rs:265-306) */
int  lrge_gpu_ava_counts(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts,
                         uint32_t *counts, lrge_stats *stats);

/* Every overlap minimap2 would report for the pair of sets (one record per kept chain = one PAF row / mm_reg1_t as
 * liblrge copies it out, aligner.rs:244-290; with MM_F_ALL_CHAINS every chain is reported, tp:A:S, mapq 0), sorted by
 * query and, within a query, in the order mm_map returns its regions (mm_gen_regs: score descending, hash tie-break).
 * query/target are read indices into Q / T.  dv and rl need the whole index on one GPU (no target sharding).
 * Up to cap records are written; *n_out receives the number found (call again with a larger buffer if > cap).
 * counts / had_mapping (optional) are filled as by lrge_gpu_twoset_counts. */
typedef struct lrge_overlap {
    uint32_t query, target;
    int32_t  rev;                 /* 0 '+', 1 '-' */
    int32_t  qs, qe, ts, te;      /* PafRecord query_start/query_end/target_start/target_end */
    int32_t  score, cnt;          /* s1:i, cm:i */
    int32_t  mlen, blen;          /* match_len, block_len */
    float    dv;                  /* dv:f, mm_est_err (aligner.rs:264-268); < 0: not estimated */
    int32_t  rl;                  /* rl:i, length of the query's repetitive seeds (aligner.rs:269) */
} lrge_overlap;
int  lrge_gpu_twoset_overlaps(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                              lrge_overlap *out, uint64_t cap, uint64_t *n_out,
                              uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats);
int  lrge_gpu_ava_overlaps(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts,
                           lrge_overlap *out, uint64_t cap, uint64_t *n_out, uint32_t *counts, lrge_stats *stats);
/* One PAF line (no newline) for an overlap, formatted as liblrge's PafRecord serialises it (mapping.rs:81-177):
 * 12 columns, then tp:A:S cm:i s1:i dv:f rl:i.  dv < 0 prints as 0. Returns the length written (truncated to cap-1). */
int  lrge_paf_format(const lrge_overlap *o, const char *qname, int qname_len, int32_t qlen,
                     const char *tname, int tname_len, int32_t tlen, float dv, int32_t rl, char *buf, size_t cap);

/* Same computation with the read sets already resident in HBM (what bench.py times as `value`). */
int  lrge_gpu_stage(lrge_ctx *ctx, const lrge_readset *rs, lrge_staged **out);
void lrge_gpu_unstage(lrge_ctx *ctx, lrge_staged *s);
int  lrge_gpu_twoset_counts_staged(lrge_ctx *ctx, lrge_staged *T, lrge_staged *Q, const lrge_opts *opts,
                                   uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats);
int  lrge_gpu_ava_counts_staged(lrge_ctx *ctx, lrge_staged *R, const lrge_opts *opts,
                                uint32_t *counts, lrge_stats *stats);

/* One process per GPU. Rank 0 makes an id (LRGE_COMM_ID_BYTES bytes) and hands it to every rank out of band;
 * each rank then joins.  With a communicator and opts->shard_count == world, the *_counts calls finish with
 * one NCCL all-reduce over NVLink: counts are summed (target shards are additive), had_mapping is OR-ed. */
#define LRGE_COMM_ID_BYTES 128
int  lrge_gpu_comm_unique_id(uint8_t *out, size_t size);
int  lrge_gpu_comm_init(lrge_ctx *ctx, int rank, int world, const uint8_t *unique_id, size_t size);

/* Per-kernel device times of the LAST *_counts call on this context (CUDA events on the context's stream, no
 * host syncs in between). Off by default; lrge_gpu_set_profiling(ctx, 1) turns the per-kernel events on.
 * lrge_gpu_kernel_times fills up to cap records and returns the number of distinct kernels (or < 0). */
typedef struct lrge_kernel_time {
    char     name[64];
    double   ms;                  /* summed over the launches of this kernel in the call */
    double   algorithmic_bytes;   /* summed algorithmic bytes of those launches */
    uint32_t launches;
    uint32_t passes;              /* a library radix sort makes several passes over its records inside one launch; else 1 */
} lrge_kernel_time;
int  lrge_gpu_set_profiling(lrge_ctx *ctx, int on);
int  lrge_gpu_kernel_times(lrge_ctx *ctx, lrge_kernel_time *out, int cap);

/* Pinned host memory for the caller's read buffers (optional; pageable memory works, slower). */
int  lrge_gpu_host_alloc(size_t bytes, void **out);
void lrge_gpu_host_free(void *p);

/* Stage dumps for parity tests (device -> host). mm128-style records, x/y as in minimap2:
 *   minimizers of a read set in (rid, position) order: x = hash<<8|span, y = rid<<32|pos<<1|strand;
 *   cnt[rs->n] = minimizers per read (may be NULL). *n_out receives the number written; cap = capacity. */
int  lrge_gpu_dump_minimizers(lrge_ctx *ctx, const lrge_readset *rs, int32_t platform,
                              uint64_t *x, uint64_t *y, uint64_t cap, uint64_t *n_out, uint32_t *cnt);
/* the same for the reads [first, first + count) only, staged and sketched the way a rank treats its target shard
 * (cnt[count]; y keeps the read ids of the whole set) */
int  lrge_gpu_dump_minimizers_block(lrge_ctx *ctx, const lrge_readset *rs, int32_t platform, uint32_t first, uint32_t count,
                                    uint64_t *x, uint64_t *y, uint64_t cap, uint64_t *n_out, uint32_t *cnt);
/* mid_occ of an index built over T (mm_mapopt_update, aligner.rs:189) */
int  lrge_gpu_dump_mid_occ(lrge_ctx *ctx, const lrge_readset *T, int32_t platform, int32_t *mid_occ, uint64_t *n_keys);

/* Plain or gzip FASTA / FASTQ ingestion on the host (io::count_records / io::iter_records, io.rs:121-184; the read id
 * is the header up to the first whitespace, io.rs:192-205).  bzip2 / zstd / xz / BAM inputs -> LRGE_ERR_UNSUPPORTED;
 * an unreadable or empty file -> LRGE_ERR_IO; a malformed record -> LRGE_ERR_FASTQ_PARSE.
 * lrge_fastx_load fills buffers sized from lrge_fastx_count in the lrge_readset layout (off[n+1], name_off[n+1]); the
 * capacities are checked while the file is parsed again (LRGE_ERR_INVALID_ARG if it no longer fits). */
int  lrge_fastx_count(const char *path, uint64_t *n_records, uint64_t *n_bases, uint64_t *n_name_bytes);
int  lrge_fastx_load(const char *path, uint8_t *bases, uint64_t cap_bases, uint64_t *off, uint64_t cap_records,
                     char *names, uint64_t cap_name_bytes, uint32_t *name_off);

/* Two-pass subsampled ingestion (TwoSetStrategy::split_fastq, twoset.rs:122-201; AvaStrategy::subsample_reads,
 * ava.rs:108-161): after lrge_fastx_count the caller draws its record indices (ascending, unique); this second pass keeps
 * only those records and writes them straight into pinned (pinned != 0) host buffers, up to two disjoint selections in
 * the same pass (target and query set; idx_b / out_b may be NULL).  A set keeps the file order of its records.
 * hint_bases_per_record sizes the first allocation (0: 10 000); buffers grow on demand. */
typedef struct lrge_hostset lrge_hostset;
int  lrge_fastx_load_subsets(const char *path, const uint64_t *idx_a, uint64_t n_a, const uint64_t *idx_b, uint64_t n_b,
                             uint64_t hint_bases_per_record, int pinned, lrge_hostset **out_a, lrge_hostset **out_b);
const lrge_readset *lrge_hostset_view(lrge_hostset *h);   /* valid until lrge_hostset_free */
void lrge_hostset_free(lrge_hostset *h);

/* host-side estimator arithmetic (f32, evaluation order of estimate.rs) */
float lrge_per_read_estimate(uint64_t read_len, float avg_target_len, uint64_t n_target_reads,
                             uint64_t n_ovlaps, uint32_t ovlap_thresh);
/* Estimate::estimate: filter finite (if finite != 0), median and optional quantiles (q < 0 => None).
 * Returns a bit mask of the outputs that are Some: 1 lower, 2 median, 4 upper; -1 on a bad quantile. */
int   lrge_estimate(const float *estimates, uint64_t n, int finite, float lower_q, float upper_q,
                    float *lower, float *median, float *upper);
/* PafRecord::is_internal (mapping.rs:59-77) */
int   lrge_is_internal(int32_t qlen, int32_t qs, int32_t qe, int32_t rev, int32_t tlen, int32_t ts, int32_t te, float ratio);

/* Whole strategies on host buffers: counts via the GPU, then per_read_estimate for every read.
 *   estimates[Q->n] / [R->n]; *no_mapping_count as the reference reports it.
 * TwoSetStrategy::generate_estimates (twoset.rs:587-606) / AvaStrategy::generate_estimates (ava.rs:369-381)
 * minus the random subsampling that precedes the hot path. */
int  lrge_twoset_generate_estimates(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                                    float *estimates, uint32_t *no_mapping_count, lrge_stats *stats);
int  lrge_ava_generate_estimates(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts,
                                 float *estimates, uint32_t *no_mapping_count, lrge_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
