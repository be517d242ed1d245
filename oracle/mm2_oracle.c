/*
 * mm2_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See mm2_oracle.h.
 *
 * Restates minimap2 v2.30 (NOT under /root/reference; crates.io dependency
 * minimap2-sys 0.1.30+minimap2.2.30, liblrge/Cargo.toml:21, Cargo.lock:710-713) for the
 * ava-ont / ava-pb overlap path, function by function:
 *   orc_sketch            <- sketch.c  mm_sketch, hash64, tiny queue
 *   orc_index_build       <- index.c   mm_idx_gen / mm_idx_add / worker_post / mm_idx_cal_max_occ,
 *                            options.c mm_mapopt_update           (liblrge aligner.rs:171-191)
 *   orc_index_get         <- index.c   mm_idx_get
 *   seed_mz_flt ..        <- seed.c    mm_seed_mz_flt, mm_seed_collect_all, mm_collect_matches
 *   collect_seed_hits     <- map.c     skip_seed, collect_seed_hits
 *   lchain_dp / backtrack <- lchain.c  comput_sc, mg_lchain_dp, mg_chain_bk_end, mg_chain_backtrack, compact_a
 *   gen_regs              <- hit.c     mm_gen_regs, mm_reg_set_coor, mm_cal_fuzzy_len
 *   radix_sort_*          <- ksort.h   KRADIX_SORT_INIT (RS_MIN_SIZE 64, RS_MAX_BITS 8)
 * and liblrge's own logic (twoset.rs:286-317,477-565; ava.rs:271-346; mapping.rs:59-77;
 * estimate.rs:80-157).  Compile with -ffp-contract=off: the chaining gap penalty is f32
 * arithmetic with separate multiply and add (x86-64 build of minimap2 has no FMA contraction).
 */
#define _GNU_SOURCE
#include "mm2_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <pthread.h>
#include <time.h>
#include <assert.h>

/* ------------------------------------------------------------------ utils */

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
void orc_free(void *p) { free(p); }

typedef struct { size_t n, m; orc128_t *a; } v128_t;
static inline void v128_push(v128_t *v, orc128_t e) {
    if (v->n == v->m) { v->m = v->m ? v->m << 1 : 256; v->a = (orc128_t*)realloc(v->a, v->m * sizeof(orc128_t)); }
    v->a[v->n++] = e;
}

/* ------------------------------------------------------------------ ksort.h radix sort (literal) */

#define RS_MIN_SIZE 64
#define RS_MAX_BITS 8

typedef struct { orc128_t *b, *e; } rsb128_t;

static void rs_insertsort_128x(orc128_t *beg, orc128_t *end) {
    orc128_t *i;
    for (i = beg + 1; i < end; ++i)
        if (i->x < (i - 1)->x) {
            orc128_t *j, tmp = *i;
            for (j = i; j > beg && tmp.x < (j - 1)->x; --j) *j = *(j - 1);
            *j = tmp;
        }
}
static void rs_sort_128x(orc128_t *beg, orc128_t *end, int n_bits, int s) {
    orc128_t *i;
    int size = 1 << n_bits, m = size - 1;
    rsb128_t *k, b[1 << RS_MAX_BITS], *be = b + size;
    for (k = b; k != be; ++k) k->b = k->e = beg;
    for (i = beg; i != end; ++i) ++b[i->x >> s & m].e;
    for (k = b + 1; k != be; ++k) k->e += (k - 1)->e - beg, k->b = (k - 1)->e;
    for (k = b; k != be;) {
        if (k->b != k->e) {
            rsb128_t *l;
            if ((l = b + (k->b->x >> s & m)) != k) {
                orc128_t tmp = *k->b, swap;
                do {
                    swap = tmp; tmp = *l->b; *l->b++ = swap;
                    l = b + (tmp.x >> s & m);
                } while (l != k);
                *k->b++ = tmp;
            } else ++k->b;
        } else ++k;
    }
    for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e;
    if (s) {
        s = s > n_bits ? s - n_bits : 0;
        for (k = b; k != be; ++k)
            if (k->e - k->b > RS_MIN_SIZE) rs_sort_128x(k->b, k->e, n_bits, s);
            else if (k->e - k->b > 1) rs_insertsort_128x(k->b, k->e);
    }
}
static void radix_sort_128x(orc128_t *beg, orc128_t *end) {
    if (end - beg <= RS_MIN_SIZE) rs_insertsort_128x(beg, end);
    else rs_sort_128x(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS);
}

typedef struct { uint64_t *b, *e; } rsb64_t;
static void rs_insertsort_64(uint64_t *beg, uint64_t *end) {
    uint64_t *i;
    for (i = beg + 1; i < end; ++i)
        if (*i < *(i - 1)) {
            uint64_t *j, tmp = *i;
            for (j = i; j > beg && tmp < *(j - 1); --j) *j = *(j - 1);
            *j = tmp;
        }
}
static void rs_sort_64(uint64_t *beg, uint64_t *end, int n_bits, int s) {
    uint64_t *i;
    int size = 1 << n_bits, m = size - 1;
    rsb64_t *k, b[1 << RS_MAX_BITS], *be = b + size;
    for (k = b; k != be; ++k) k->b = k->e = beg;
    for (i = beg; i != end; ++i) ++b[*i >> s & m].e;
    for (k = b + 1; k != be; ++k) k->e += (k - 1)->e - beg, k->b = (k - 1)->e;
    for (k = b; k != be;) {
        if (k->b != k->e) {
            rsb64_t *l;
            if ((l = b + (*k->b >> s & m)) != k) {
                uint64_t tmp = *k->b, swap;
                do {
                    swap = tmp; tmp = *l->b; *l->b++ = swap;
                    l = b + (tmp >> s & m);
                } while (l != k);
                *k->b++ = tmp;
            } else ++k->b;
        } else ++k;
    }
    for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e;
    if (s) {
        s = s > n_bits ? s - n_bits : 0;
        for (k = b; k != be; ++k)
            if (k->e - k->b > RS_MIN_SIZE) rs_sort_64(k->b, k->e, n_bits, s);
            else if (k->e - k->b > 1) rs_insertsort_64(k->b, k->e);
    }
}
static void radix_sort_64(uint64_t *beg, uint64_t *end) {
    if (end - beg <= RS_MIN_SIZE) rs_insertsort_64(beg, end);
    else rs_sort_64(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS);
}

/* stable alternative (ORC_TIE_STABLE): bottom-up merge sort by x */
static void stable_sort_128x(orc128_t *a, size_t n) {
    if (n < 2) return;
    orc128_t *tmp = (orc128_t*)malloc(n * sizeof(orc128_t)), *src = a, *dst = tmp;
    for (size_t wdt = 1; wdt < n; wdt <<= 1) {
        for (size_t lo = 0; lo < n; lo += 2 * wdt) {
            size_t mid = lo + wdt < n ? lo + wdt : n, hi = lo + 2 * wdt < n ? lo + 2 * wdt : n;
            size_t i = lo, j = mid, o = lo;
            while (i < mid && j < hi) dst[o++] = (src[j].x < src[i].x) ? src[j++] : src[i++];
            while (i < mid) dst[o++] = src[i++];
            while (j < hi) dst[o++] = src[j++];
        }
        orc128_t *t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(orc128_t));
    free(tmp);
}
static void sort_128x(orc128_t *a, size_t n, int tie_mode) {
    if (tie_mode == ORC_TIE_STABLE) stable_sort_128x(a, n);
    else radix_sort_128x(a, a + n);
}

/* ------------------------------------------------------------------ options (options.c) */

int orc_opt_init(orc_opt_t *o, const char *preset, int dual) {
    memset(o, 0, sizeof(*o));
    /* mm_idxopt_init */
    o->k = 15; o->w = 10; o->is_hpc = 0; o->bucket_bits = 14;
    /* mm_mapopt_init */
    o->seed = 11;
    o->mid_occ_frac = 2e-4f; o->min_mid_occ = 10; o->max_mid_occ = 1000000;
    o->max_max_occ = 4095; o->occ_dist = 500; o->max_occ = 0; o->mid_occ = 0;
    o->q_occ_frac = 0.01f;
    o->min_cnt = 3; o->min_chain_score = 40;
    o->bw = 500; o->bw_long = 20000;
    o->max_gap = 5000; o->max_gap_ref = -1;
    o->max_chain_skip = 25; o->max_chain_iter = 5000;
    o->chain_gap_scale = 0.8f; o->chain_skip_scale = 0.0f;
    o->flag = 0;
    o->tie_mode = ORC_TIE_KSORT;
    if (preset == 0) return 0;
    if (strcmp(preset, "ava-ont") == 0) {
        o->is_hpc = 0; o->k = 15; o->w = 5;
        o->flag |= ORC_F_ALL_CHAINS | ORC_F_NO_DIAG | ORC_F_NO_DUAL | ORC_F_NO_LJOIN;
        o->min_chain_score = 100; o->max_chain_skip = 25;
        o->bw = o->bw_long = 2000;
        o->occ_dist = 0;
    } else if (strcmp(preset, "ava-pb") == 0) {
        o->is_hpc = 1; o->k = 19; o->w = 5;
        o->flag |= ORC_F_ALL_CHAINS | ORC_F_NO_DIAG | ORC_F_NO_DUAL | ORC_F_NO_LJOIN;
        o->min_chain_score = 100; o->max_chain_skip = 25;
        o->bw_long = o->bw;
        o->occ_dist = 0;
    } else return -1;
    /* Aligner::dual (aligner.rs:93-103) */
    if (dual) o->flag &= ~(int64_t)ORC_F_NO_DUAL; else o->flag |= ORC_F_NO_DUAL;
    return 0;
}

/* ------------------------------------------------------------------ sketch.c */

static const unsigned char nt4_table[256] = {
    0, 1, 2, 3,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 0, 4, 1,  4, 4, 4, 2,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  3, 3, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,
    4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4,  4, 4, 4, 4
};

static inline uint64_t hash64_masked(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

typedef struct { int front, count; int a[32]; } tiny_queue_t;
static inline void tq_push(tiny_queue_t *q, int x) { q->a[((q->count++) + q->front) & 0x1f] = x; }
static inline int tq_shift(tiny_queue_t *q) {
    int x;
    if (q->count == 0) return -1;
    x = q->a[q->front++]; q->front &= 0x1f; --q->count;
    return x;
}

static void mm_sketch(const uint8_t *str, int len, int w, int k, uint32_t rid, int is_hpc, v128_t *p) {
    uint64_t shift1 = 2 * (k - 1), mask = (1ULL << 2 * k) - 1, kmer[2] = {0, 0};
    int i, j, l, buf_pos, min_pos, kmer_span = 0;
    orc128_t buf[256], min = { UINT64_MAX, UINT64_MAX };
    tiny_queue_t tq;

    assert(len > 0 && (w > 0 && w < 256) && (k > 0 && k <= 28));
    memset(buf, 0xff, w * 16);
    memset(&tq, 0, sizeof(tiny_queue_t));

    for (i = l = buf_pos = min_pos = 0; i < len; ++i) {
        int c = nt4_table[str[i]];
        orc128_t info = { UINT64_MAX, UINT64_MAX };
        if (c < 4) {
            int z;
            if (is_hpc) {
                int skip_len = 1;
                if (i + 1 < len && nt4_table[str[i + 1]] == c) {
                    for (skip_len = 2; i + skip_len < len; ++skip_len)
                        if (nt4_table[str[i + skip_len]] != c) break;
                    i += skip_len - 1;
                }
                tq_push(&tq, skip_len);
                kmer_span += skip_len;
                if (tq.count > k) kmer_span -= tq_shift(&tq);
            } else kmer_span = l + 1 < k ? l + 1 : k;
            kmer[0] = (kmer[0] << 2 | c) & mask;
            kmer[1] = (kmer[1] >> 2) | (3ULL ^ c) << shift1;
            if (kmer[0] == kmer[1]) continue;
            z = kmer[0] < kmer[1] ? 0 : 1;
            ++l;
            if (l >= k && kmer_span < 256) {
                info.x = hash64_masked(kmer[z], mask) << 8 | kmer_span;
                info.y = (uint64_t)rid << 32 | (uint32_t)i << 1 | z;
            }
        } else l = 0, tq.count = tq.front = 0, kmer_span = 0;
        buf[buf_pos] = info;
        if (l == w + k - 1 && min.x != UINT64_MAX) {
            for (j = buf_pos + 1; j < w; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) v128_push(p, buf[j]);
            for (j = 0; j < buf_pos; ++j)
                if (min.x == buf[j].x && buf[j].y != min.y) v128_push(p, buf[j]);
        }
        if (info.x <= min.x) {
            if (l >= w + k && min.x != UINT64_MAX) v128_push(p, min);
            min = info, min_pos = buf_pos;
        } else if (buf_pos == min_pos) {
            if (l >= w + k - 1 && min.x != UINT64_MAX) v128_push(p, min);
            for (j = buf_pos + 1, min.x = UINT64_MAX; j < w; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            for (j = 0; j <= buf_pos; ++j)
                if (min.x >= buf[j].x) min = buf[j], min_pos = j;
            if (l >= w + k - 1 && min.x != UINT64_MAX) {
                for (j = buf_pos + 1; j < w; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) v128_push(p, buf[j]);
                for (j = 0; j <= buf_pos; ++j)
                    if (min.x == buf[j].x && min.y != buf[j].y) v128_push(p, buf[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
    }
    if (min.x != UINT64_MAX) v128_push(p, min);
}

int orc_sketch(const uint8_t *seq, int32_t len, int32_t w, int32_t k, uint32_t rid, int32_t is_hpc,
               orc128_t **out, int64_t *n_out) {
    v128_t v = {0, 0, 0};
    if (len > 0) mm_sketch(seq, len, w, k, rid, is_hpc, &v);
    *out = v.a; *n_out = (int64_t)v.n;
    return 0;
}

int orc_sketch_readset(const orc_readset_t *rs, const orc_opt_t *opt, orc128_t **out, int64_t *n_out, uint32_t *cnt) {
    v128_t v = {0, 0, 0};
    for (uint32_t r = 0; r < rs->n; ++r) {
        size_t n0 = v.n;
        int64_t len = (int64_t)(rs->off[r + 1] - rs->off[r]);
        if (len > 0) mm_sketch(rs->bases + rs->off[r], (int)len, opt->w, opt->k, r, opt->is_hpc, &v);
        if (cnt) cnt[r] = (uint32_t)(v.n - n0);
    }
    *out = v.a; *n_out = (int64_t)v.n;
    return 0;
}

/* ------------------------------------------------------------------ index.c */

typedef struct {
    v128_t a;          /* minimizers of this bucket (freed after post) */
    int32_t n;         /* size of p */
    uint64_t *p;       /* positions of non-singleton minimizers */
    /* open-addressing table standing in for khash(idx): key = minier>>b<<1 | singleton */
    uint32_t h_cap;    /* power of two, 0 if empty */
    uint32_t h_n;
    uint64_t *h_key;   /* UINT64_MAX = empty */
    uint64_t *h_val;
} idx_bucket_t;

struct orc_index_s {
    int32_t b, w, k, is_hpc;
    uint32_t n_seq;
    char **name;
    uint32_t *len;
    idx_bucket_t *B;
    uint64_t n_minimizers;
};

static inline uint32_t h_slot(uint64_t key_no_flag, uint32_t cap) {
    uint64_t h = key_no_flag * 0x9E3779B97F4A7C15ULL;
    return (uint32_t)(h >> 32) & (cap - 1);
}
static void h_init(idx_bucket_t *b, uint32_t n_keys) {
    uint32_t cap = 4;
    while (cap * 0.7 < n_keys) cap <<= 1;
    b->h_cap = cap; b->h_n = 0;
    b->h_key = (uint64_t*)malloc((size_t)cap * 8);
    b->h_val = (uint64_t*)malloc((size_t)cap * 8);
    memset(b->h_key, 0xff, (size_t)cap * 8);
}
static void h_put(idx_bucket_t *b, uint64_t key, uint64_t val) { /* key includes the singleton flag bit */
    uint32_t s = h_slot(key >> 1, b->h_cap);
    while (b->h_key[s] != UINT64_MAX) { assert(b->h_key[s] >> 1 != key >> 1); s = (s + 1) & (b->h_cap - 1); }
    b->h_key[s] = key; b->h_val[s] = val; ++b->h_n;
}
static inline int64_t h_get(const idx_bucket_t *b, uint64_t key_no_flag) {
    if (b->h_cap == 0) return -1;
    uint32_t s = h_slot(key_no_flag, b->h_cap);
    while (b->h_key[s] != UINT64_MAX) {
        if (b->h_key[s] >> 1 == key_no_flag) return s;
        s = (s + 1) & (b->h_cap - 1);
    }
    return -1;
}

static void worker_post(orc_index_t *mi, long i) {
    int n, n_keys;
    size_t j, start_a, start_p;
    idx_bucket_t *b = &mi->B[i];
    if (b->a.n == 0) return;
    radix_sort_128x(b->a.a, b->a.a + b->a.n);
    for (j = 1, n = 1, n_keys = 0, b->n = 0; j <= b->a.n; ++j) {
        if (j == b->a.n || b->a.a[j].x >> 8 != b->a.a[j - 1].x >> 8) {
            ++n_keys;
            if (n > 1) b->n += n;
            n = 1;
        } else ++n;
    }
    h_init(b, n_keys);
    b->p = (uint64_t*)calloc(b->n > 0 ? b->n : 1, 8);
    for (j = 1, n = 1, start_a = start_p = 0; j <= b->a.n; ++j) {
        if (j == b->a.n || b->a.a[j].x >> 8 != b->a.a[j - 1].x >> 8) {
            orc128_t *p = &b->a.a[j - 1];
            uint64_t key = p->x >> 8 >> mi->b << 1;
            assert(j == start_a + n);
            if (n == 1) {
                h_put(b, key | 1, p->y);
            } else {
                int k;
                for (k = 0; k < n; ++k) b->p[start_p + k] = b->a.a[start_a + k].y;
                radix_sort_64(&b->p[start_p], &b->p[start_p + n]);
                h_put(b, key, (uint64_t)start_p << 32 | n);
                start_p += n;
            }
            start_a = j, n = 1;
        } else ++n;
    }
    assert(b->n == (int32_t)start_p);
    free(b->a.a); b->a.a = 0; b->a.n = b->a.m = 0;
}

const uint64_t *orc_index_get(const orc_index_t *mi, uint64_t minier, int32_t *n) {
    int mask = (1 << mi->b) - 1;
    const idx_bucket_t *b = &mi->B[minier & mask];
    int64_t s;
    *n = 0;
    s = h_get(b, minier >> mi->b);
    if (s < 0) return 0;
    if (b->h_key[s] & 1) { *n = 1; return &b->h_val[s]; }
    *n = (uint32_t)b->h_val[s];
    return &b->p[b->h_val[s] >> 32];
}

static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b; return x < y ? -1 : x > y; }

static int32_t idx_cal_max_occ(const orc_index_t *mi, float f) {
    int i;
    size_t n = 0;
    uint32_t thres, *a;
    if (f <= 0.) return INT32_MAX;
    for (i = 0; i < 1 << mi->b; ++i) n += mi->B[i].h_n;
    if (n == 0) return INT32_MAX;
    a = (uint32_t*)malloc(n * 4);
    for (i = 0, n = 0; i < 1 << mi->b; ++i) {
        const idx_bucket_t *b = &mi->B[i];
        for (uint32_t s = 0; s < b->h_cap; ++s) {
            if (b->h_key[s] == UINT64_MAX) continue;
            a[n++] = (b->h_key[s] & 1) ? 1 : (uint32_t)b->h_val[s];
        }
    }
    /* ks_ksmall(n, a, k): k-th smallest, 0-based; value is algorithm independent */
    qsort(a, n, 4, cmp_u32);
    thres = a[(uint32_t)((1. - f) * n)] + 1;
    free(a);
    return (int32_t)thres;
}

typedef struct { orc_index_t *mi; long next; long n; pthread_mutex_t mu; } post_job_t;
static void *post_thread(void *arg) {
    post_job_t *j = (post_job_t*)arg;
    for (;;) {
        long i;
        pthread_mutex_lock(&j->mu); i = j->next; j->next += 64; pthread_mutex_unlock(&j->mu);
        if (i >= j->n) break;
        for (long e = i + 64 < j->n ? i + 64 : j->n; i < e; ++i) worker_post(j->mi, i);
    }
    return 0;
}

typedef struct { const orc_readset_t *T; const orc_opt_t *opt; v128_t *per_read; uint32_t next; pthread_mutex_t mu; } sk_job_t;
static void *sketch_thread(void *arg) {
    sk_job_t *j = (sk_job_t*)arg;
    for (;;) {
        uint32_t r0, r1;
        pthread_mutex_lock(&j->mu); r0 = j->next; j->next += 32; pthread_mutex_unlock(&j->mu);
        if (r0 >= j->T->n) break;
        r1 = r0 + 32 < j->T->n ? r0 + 32 : j->T->n;
        for (uint32_t r = r0; r < r1; ++r) {
            int64_t len = (int64_t)(j->T->off[r + 1] - j->T->off[r]);
            if (len > 0) mm_sketch(j->T->bases + j->T->off[r], (int)len, j->opt->w, j->opt->k, r, j->opt->is_hpc, &j->per_read[r]);
        }
    }
    return 0;
}

orc_index_t *orc_index_build(const orc_readset_t *T, orc_opt_t *opt, int n_threads) {
    orc_index_t *mi = (orc_index_t*)calloc(1, sizeof(*mi));
    if (n_threads < 1) n_threads = 1;
    mi->b = opt->bucket_bits; mi->w = opt->w; mi->k = opt->k; mi->is_hpc = opt->is_hpc;
    mi->n_seq = T->n;
    mi->B = (idx_bucket_t*)calloc((size_t)1 << mi->b, sizeof(idx_bucket_t));
    mi->name = (char**)calloc(T->n ? T->n : 1, sizeof(char*));
    mi->len = (uint32_t*)calloc(T->n ? T->n : 1, 4);
    for (uint32_t r = 0; r < T->n; ++r) {
        uint32_t nl = T->name_off[r + 1] - T->name_off[r];
        mi->name[r] = (char*)malloc(nl + 1);
        memcpy(mi->name[r], T->names + T->name_off[r], nl); mi->name[r][nl] = 0;
        mi->len[r] = (uint32_t)(T->off[r + 1] - T->off[r]);
    }
    /* step 1 of mm_idx_gen pipeline: sketch every sequence (zero-length ones are skipped) */
    {
        sk_job_t job; pthread_t *th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        job.T = T; job.opt = opt; job.next = 0; pthread_mutex_init(&job.mu, 0);
        job.per_read = (v128_t*)calloc(T->n ? T->n : 1, sizeof(v128_t));
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, sketch_thread, &job);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
        /* step 2: mm_idx_add in rid order */
        int mask = (1 << mi->b) - 1;
        for (uint32_t r = 0; r < T->n; ++r) {
            v128_t *v = &job.per_read[r];
            for (size_t i = 0; i < v->n; ++i) v128_push(&mi->B[v->a[i].x >> 8 & mask].a, v->a[i]);
            mi->n_minimizers += v->n;
            free(v->a);
        }
        free(job.per_read); free(th); pthread_mutex_destroy(&job.mu);
    }
    /* mm_idx_post */
    {
        post_job_t job; pthread_t *th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
        job.mi = mi; job.next = 0; job.n = 1L << mi->b; pthread_mutex_init(&job.mu, 0);
        for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, post_thread, &job);
        for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
        free(th); pthread_mutex_destroy(&job.mu);
    }
    /* mm_mapopt_update (aligner.rs:189) */
    if (opt->mid_occ <= 0) {
        opt->mid_occ = idx_cal_max_occ(mi, opt->mid_occ_frac);
        if (opt->mid_occ < opt->min_mid_occ) opt->mid_occ = opt->min_mid_occ;
        if (opt->max_mid_occ > opt->min_mid_occ && opt->mid_occ > opt->max_mid_occ) opt->mid_occ = opt->max_mid_occ;
    }
    if (opt->bw_long < opt->bw) opt->bw_long = opt->bw;
    return mi;
}

void orc_index_destroy(orc_index_t *mi) {
    if (!mi) return;
    for (long i = 0; i < 1L << mi->b; ++i) { free(mi->B[i].p); free(mi->B[i].h_key); free(mi->B[i].h_val); free(mi->B[i].a.a); }
    for (uint32_t r = 0; r < mi->n_seq; ++r) free(mi->name[r]);
    free(mi->name); free(mi->len); free(mi->B); free(mi);
}
uint64_t orc_index_n_keys(const orc_index_t *mi) { uint64_t n = 0; for (long i = 0; i < 1L << mi->b; ++i) n += mi->B[i].h_n; return n; }
uint64_t orc_index_n_minimizers(const orc_index_t *mi) { return mi->n_minimizers; }

/* ------------------------------------------------------------------ seed.c */

typedef struct {
    uint32_t n;
    uint32_t q_pos;
    uint32_t q_span:8, flt:1, seg_id:8, is_tandem:1;
    const uint64_t *cr;
} seed_t;

static void seed_mz_flt(v128_t *mv, int32_t q_occ_max, float q_occ_frac, orc_stats_t *st_) {
    orc128_t *a;
    size_t i, j, st;
    if (mv->n <= (size_t)q_occ_max || q_occ_frac <= 0.0f || q_occ_max <= 0) return;
    a = (orc128_t*)malloc(mv->n * sizeof(orc128_t));
    for (i = 0; i < mv->n; ++i) a[i].x = mv->a[i].x, a[i].y = i;
    radix_sort_128x(a, a + mv->n);
    for (st = 0, i = 1; i <= mv->n; ++i) {
        if (i == mv->n || a[i].x != a[st].x) {
            int32_t cnt = (int32_t)(i - st);
            if (cnt > q_occ_max && cnt > mv->n * q_occ_frac)
            {
                for (j = st; j < i; ++j) mv->a[a[j].y].x = 0;
                if (st_) st_->n_mz_flt_deleted += cnt;
            }
            st = i;
        }
    }
    free(a);
    for (i = j = 0; i < mv->n; ++i)
        if (mv->a[i].x != 0) mv->a[j++] = mv->a[i];
    mv->n = j;
}

static seed_t *seed_collect_all(const orc_index_t *mi, const v128_t *mv, int32_t *n_m_) {
    seed_t *m = (seed_t*)malloc((mv->n ? mv->n : 1) * sizeof(seed_t));
    size_t i; int32_t k;
    for (i = k = 0; i < mv->n; ++i) {
        const uint64_t *cr; seed_t *q; orc128_t *p = &mv->a[i];
        uint32_t q_pos = (uint32_t)p->y, q_span = p->x & 0xff;
        int t;
        cr = orc_index_get(mi, p->x >> 8, &t);
        if (t == 0) continue;
        q = &m[k++];
        q->q_pos = q_pos, q->q_span = q_span, q->cr = cr, q->n = t, q->seg_id = p->y >> 32;
        q->is_tandem = q->flt = 0;
        if (i > 0 && p->x >> 8 == mv->a[i - 1].x >> 8) q->is_tandem = 1;
        if (i < mv->n - 1 && p->x >> 8 == mv->a[i + 1].x >> 8) q->is_tandem = 1;
    }
    *n_m_ = k;
    return m;
}

/* occ_dist == 0 for both presets -> the plain "n > max_occ => filtered" branch */
static seed_t *collect_matches(int *_n_m, int max_occ, const orc_index_t *mi, const v128_t *mv, int64_t *n_a, int *rep_len, uint64_t *n_flt) {
    int rep_st = 0, rep_en = 0, n_m, n_m0;
    int i;
    seed_t *m = seed_collect_all(mi, mv, &n_m0);
    for (i = 0; i < n_m0; ++i)
        if (m[i].n > (uint32_t)max_occ) { m[i].flt = 1; if (n_flt) ++*n_flt; }
    for (i = 0, n_m = 0, *rep_len = 0, *n_a = 0; i < n_m0; ++i) {
        seed_t *q = &m[i];
        if (q->flt) {
            int en = (q->q_pos >> 1) + 1, st = en - q->q_span;
            if (st > rep_en) { *rep_len += rep_en - rep_st; rep_st = st, rep_en = en; }
            else rep_en = en;
        } else {
            *n_a += q->n;
            m[n_m++] = *q;
        }
    }
    *rep_len += rep_en - rep_st;
    *_n_m = n_m;
    return m;
}

/* ------------------------------------------------------------------ map.c */

#define SEED_TANDEM (1ULL << 42)
#define SEED_SELF   (1ULL << 43)
#define SEED_SEG_SHIFT 48

static int skip_seed(int64_t flag, uint64_t r, const seed_t *q, const char *qname, int qlen, const orc_index_t *mi, int *is_self) {
    *is_self = 0;
    if (qname && (flag & (ORC_F_NO_DIAG | ORC_F_NO_DUAL))) {
        int cmp = strcmp(qname, mi->name[r >> 32]);
        if ((flag & ORC_F_NO_DIAG) && cmp == 0 && (int)mi->len[r >> 32] == qlen) {
            if ((uint32_t)r >> 1 == (q->q_pos >> 1)) return 1;
            if ((r & 1) == (q->q_pos & 1)) *is_self = 1;
        }
        if ((flag & ORC_F_NO_DUAL) && cmp > 0) return 1;
    }
    return 0;
}

static orc128_t *collect_seed_hits(const orc_opt_t *opt, int max_occ, const orc_index_t *mi, const char *qname,
                                   const v128_t *mv, int qlen, int64_t *n_a, int *rep_len, orc_stats_t *st) {
    int i, n_m;
    seed_t *m = collect_matches(&n_m, max_occ, mi, mv, n_a, rep_len, st ? &st->n_occ_filtered : 0);
    orc128_t *a = (orc128_t*)malloc((*n_a ? *n_a : 1) * sizeof(orc128_t));
    if (st) { st->n_seeds += n_m; st->n_hits += *n_a; }
    for (i = 0, *n_a = 0; i < n_m; ++i) {
        seed_t *q = &m[i];
        const uint64_t *r = q->cr;
        uint32_t k;
        for (k = 0; k < q->n; ++k) {
            int32_t is_self, rpos = (uint32_t)r[k] >> 1;
            orc128_t *p;
            if (skip_seed(opt->flag, r[k], q, qname, qlen, mi, &is_self)) continue;
            if (opt->shard_count > 1 && (int32_t)((r[k] >> 32) % (uint32_t)opt->shard_count) != opt->shard_index) continue;
            p = &a[(*n_a)++];
            if ((r[k] & 1) == (q->q_pos & 1)) {
                p->x = (r[k] & 0xffffffff00000000ULL) | rpos;
                p->y = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
            } else {
                p->x = 1ULL << 63 | (r[k] & 0xffffffff00000000ULL) | rpos;
                p->y = (uint64_t)q->q_span << 32 | (qlen - ((q->q_pos >> 1) + 1 - q->q_span) - 1);
            }
            p->y |= (uint64_t)q->seg_id << SEED_SEG_SHIFT;
            if (q->is_tandem) p->y |= SEED_TANDEM;
            if (is_self) p->y |= SEED_SELF;
        }
    }
    free(m);
    sort_128x(a, (size_t)*n_a, opt->tie_mode);
    return a;
}

/* ------------------------------------------------------------------ lchain.c */

static inline float mg_log2(float x) { /* NB: only valid for x >= 2 */
    union { float f; uint32_t i; } z = { x };
    float log_2 = ((z.i >> 23) & 255) - 128;
    z.i &= ~(255U << 23);
    z.i += 127U << 23;
    log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
    return log_2;
}

static inline int32_t comput_sc(const orc128_t *ai, const orc128_t *aj, int32_t max_dist_x, int32_t max_dist_y, int32_t bw,
                                float chn_pen_gap, float chn_pen_skip) {
    int32_t dq = (int32_t)ai->y - (int32_t)aj->y, dr, dd, dg, q_span, sc;
    if (dq <= 0 || dq > max_dist_x) return INT32_MIN;
    dr = (int32_t)(ai->x - aj->x);
    if (dr == 0 || dq > max_dist_y) return INT32_MIN;      /* sidi == sidj always (n_seg == 1) */
    dd = dr > dq ? dr - dq : dq - dr;
    if (dd > bw) return INT32_MIN;
    dg = dr < dq ? dr : dq;
    q_span = aj->y >> 32 & 0xff;
    sc = q_span < dg ? q_span : dg;
    if (dd || dg > q_span) {
        float lin_pen, log_pen;
        lin_pen = chn_pen_gap * (float)dd + chn_pen_skip * (float)dg;
        log_pen = dd >= 1 ? mg_log2(dd + 1) : 0.0f;
        sc -= (int)(lin_pen + .5f * log_pen);
    }
    return sc;
}

void orc_gap_penalty_table(int32_t k, float chain_gap_scale, int32_t n, int32_t *pen) {
    float chn_pen_gap = chain_gap_scale * 0.01 * k;
    for (int32_t dd = 0; dd < n; ++dd) {
        float lin_pen = chn_pen_gap * (float)dd + 0.0f * (float)1;
        float log_pen = dd >= 1 ? mg_log2(dd + 1) : 0.0f;
        pen[dd] = (int)(lin_pen + .5f * log_pen);
    }
}

static int64_t chain_bk_end(int32_t max_drop, const orc128_t *z, const int32_t *f, const int64_t *p, int32_t *t, int64_t k) {
    int64_t i = z[k].y, end_i = -1, max_i = i;
    int32_t max_s = 0;
    if (i < 0 || t[i] != 0) return i;
    do {
        int32_t s;
        t[i] = 2;
        end_i = i = p[i];
        s = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
        if (s > max_s) max_s = s, max_i = i;
        else if (max_s - s > max_drop) break;
    } while (i >= 0 && t[i] == 0);
    for (i = z[k].y; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
    return max_i;
}

static uint64_t *chain_backtrack(int64_t n, const int32_t *f, const int64_t *p, int32_t *v, int32_t *t,
                                 int32_t min_cnt, int32_t min_sc, int32_t max_drop, int32_t *n_u_, int32_t *n_v_, int tie_mode) {
    orc128_t *z;
    uint64_t *u;
    int64_t i, k, n_z, n_v;
    int32_t n_u;
    *n_u_ = *n_v_ = 0;
    for (i = 0, n_z = 0; i < n; ++i) if (f[i] >= min_sc) ++n_z;
    if (n_z == 0) return 0;
    z = (orc128_t*)malloc(n_z * sizeof(orc128_t));
    for (i = 0, k = 0; i < n; ++i) if (f[i] >= min_sc) z[k].x = f[i], z[k++].y = i;
    sort_128x(z, (size_t)n_z, tie_mode);
    memset(t, 0, n * 4);
    for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
        if (t[z[k].y] == 0) {
            int64_t n_v0 = n_v, end_i;
            int32_t sc;
            end_i = chain_bk_end(max_drop, z, f, p, t, k);
            for (i = z[k].y; i != end_i; i = p[i]) ++n_v, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) ++n_u;
            else n_v = n_v0;
        }
    }
    u = (uint64_t*)malloc((n_u ? n_u : 1) * 8);
    memset(t, 0, n * 4);
    for (k = n_z - 1, n_v = n_u = 0; k >= 0; --k) {
        if (t[z[k].y] == 0) {
            int64_t n_v0 = n_v, end_i;
            int32_t sc;
            end_i = chain_bk_end(max_drop, z, f, p, t, k);
            for (i = z[k].y; i != end_i; i = p[i]) v[n_v++] = (int32_t)i, t[i] = 1;
            sc = i < 0 ? (int32_t)z[k].x : (int32_t)z[k].x - f[i];
            if (sc >= min_sc && n_v > n_v0 && n_v - n_v0 >= min_cnt) u[n_u++] = (uint64_t)sc << 32 | (n_v - n_v0);
            else n_v = n_v0;
        }
    }
    free(z);
    *n_u_ = n_u, *n_v_ = (int32_t)n_v;
    return u;
}

static orc128_t *compact_a(int32_t n_u, uint64_t *u, int32_t n_v, int32_t *v, orc128_t *a, int tie_mode) {
    orc128_t *b, *w;
    uint64_t *u2;
    int64_t i, j, k;
    b = (orc128_t*)malloc((n_v ? n_v : 1) * sizeof(orc128_t));
    for (i = 0, k = 0; i < n_u; ++i) {
        int32_t k0 = (int32_t)k, ni = (int32_t)u[i];
        for (j = 0; j < ni; ++j) b[k++] = a[v[k0 + (ni - j - 1)]];
    }
    free(v);
    w = (orc128_t*)malloc((n_u ? n_u : 1) * sizeof(orc128_t));
    for (i = k = 0; i < n_u; ++i) {
        w[i].x = b[k].x, w[i].y = (uint64_t)k << 32 | i;
        k += (int32_t)u[i];
    }
    sort_128x(w, (size_t)n_u, tie_mode);
    u2 = (uint64_t*)malloc((n_u ? n_u : 1) * 8);
    for (i = k = 0; i < n_u; ++i) {
        int32_t j2 = (int32_t)w[i].y, n = (int32_t)u[j2];
        u2[i] = u[j2];
        memcpy(&a[k], &b[w[i].y >> 32], n * sizeof(orc128_t));
        k += n;
    }
    memcpy(u, u2, n_u * 8);
    memcpy(b, a, k * sizeof(orc128_t));
    free(a); free(w); free(u2);
    return b;
}

static orc128_t *lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                           float chn_pen_gap, float chn_pen_skip, int64_t n, orc128_t *a, int *n_u_, uint64_t **_u,
                           int tie_mode, orc_stats_t *st_) {
    int32_t *f, *t, *v, n_u, n_v, max_drop = bw;
    int64_t *p, i, j, max_ii, st = 0, n_iter = 0;
    uint64_t *u;
    if (_u) *_u = 0, *n_u_ = 0;
    if (n == 0 || a == 0) { free(a); return 0; }
    if (max_dist_x < bw) max_dist_x = bw;
    if (max_dist_y < bw) max_dist_y = bw;
    p = (int64_t*)malloc(n * 8);
    f = (int32_t*)malloc(n * 4);
    v = (int32_t*)malloc(n * 4);
    t = (int32_t*)calloc(n, 4);
    for (i = 0, max_ii = -1; i < n; ++i) {
        int64_t max_j = -1, end_j;
        int32_t max_f = a[i].y >> 32 & 0xff, n_skip = 0;
        while (st < i && (a[i].x >> 32 != a[st].x >> 32 || a[i].x > a[st].x + max_dist_x)) ++st;
        if (i - st > max_iter) { st = i - max_iter; if (st_) ++st_->n_max_iter_clamp; }
        for (j = i - 1; j >= st; --j) {
            int32_t sc;
            sc = comput_sc(&a[i], &a[j], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            ++n_iter;
            if (sc == INT32_MIN) continue;
            sc += f[j];
            if (sc > max_f) {
                max_f = sc, max_j = j;
                if (n_skip > 0) --n_skip;
            } else if (t[j] == (int32_t)i) {
                if (++n_skip > max_skip) { if (st_) ++st_->n_skip_break; break; }
            }
            if (p[j] >= 0) t[p[j]] = (int32_t)i;
        }
        end_j = j;
        if (max_ii < 0 || a[i].x - a[max_ii].x > (uint64_t)(int64_t)max_dist_x) {
            int32_t max = INT32_MIN;
            max_ii = -1;
            for (j = i - 1; j >= st; --j)
                if (max < f[j]) max = f[j], max_ii = j;
        }
        if (max_ii >= 0 && max_ii < end_j) {
            int32_t tmp = comput_sc(&a[i], &a[max_ii], max_dist_x, max_dist_y, bw, chn_pen_gap, chn_pen_skip);
            if (tmp != INT32_MIN && max_f < tmp + f[max_ii]) max_f = tmp + f[max_ii], max_j = max_ii;
        }
        f[i] = max_f, p[i] = max_j;
        v[i] = max_j >= 0 && v[max_j] > max_f ? v[max_j] : max_f;
        if (max_ii < 0 || (a[i].x - a[max_ii].x <= (uint64_t)(int64_t)max_dist_x && f[max_ii] < f[i])) max_ii = i;
    }
    if (st_) st_->n_dp_iter += n_iter;
    u = chain_backtrack(n, f, p, v, t, min_cnt, min_sc, max_drop, &n_u, &n_v, tie_mode);
    *n_u_ = n_u, *_u = u;
    free(p); free(f); free(t);
    if (n_u == 0) { free(a); free(v); return 0; }
    return compact_a(n_u, u, n_v, v, a, tie_mode);
}

/* ------------------------------------------------------------------ hit.c */

static inline uint64_t hash64(uint64_t key) {
    key = (~key + (key << 21));
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8));
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4));
    key = key ^ key >> 28;
    key = (key + (key << 31));
    return key;
}
static inline uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15);
    key ^=  (key >> 10);
    key +=  (key << 3);
    key ^=  (key >> 6);
    key += ~(key << 11);
    key ^=  (key >> 16);
    return key;
}
static inline uint32_t x31_hash_string(const char *s) {
    uint32_t h = (uint32_t)*s;
    if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
    return h;
}

static void cal_fuzzy_len(orc_reg_t *r, const orc128_t *a, int32_t as) {
    int i;
    r->mlen = r->blen = 0;
    if (r->cnt <= 0) return;
    r->mlen = r->blen = a[as].y >> 32 & 0xff;
    for (i = as + 1; i < as + r->cnt; ++i) {
        int span = a[i].y >> 32 & 0xff;
        int tl = (int32_t)a[i].x - (int32_t)a[i - 1].x;
        int ql = (int32_t)a[i].y - (int32_t)a[i - 1].y;
        r->blen += tl > ql ? tl : ql;
        r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
    }
}
static void reg_set_coor(orc_reg_t *r, int32_t qlen, const orc128_t *a, int32_t as) {
    int32_t k = as, q_span = (int32_t)(a[k].y >> 32 & 0xff);
    r->rev = a[k].x >> 63;
    r->rid = a[k].x << 1 >> 33;
    r->rs = (int32_t)a[k].x + 1 > q_span ? (int32_t)a[k].x + 1 - q_span : 0;
    r->re = (int32_t)a[k + r->cnt - 1].x + 1;
    if (!r->rev) {
        r->qs = (int32_t)a[k].y + 1 - q_span;
        r->qe = (int32_t)a[k + r->cnt - 1].y + 1;
    } else {
        r->qs = qlen - ((int32_t)a[k + r->cnt - 1].y + 1);
        r->qe = qlen - ((int32_t)a[k].y + 1 - q_span);
    }
    cal_fuzzy_len(r, a, as);
}
static orc_reg_t *gen_regs(uint32_t hash, int qlen, int n_u, uint64_t *u, orc128_t *a, int tie_mode) {
    orc128_t *z, tmp;
    orc_reg_t *r;
    int i, k;
    if (n_u == 0) return 0;
    z = (orc128_t*)malloc(n_u * 16);
    for (i = k = 0; i < n_u; ++i) {
        uint32_t h = (uint32_t)hash64((hash64(a[k].x) + hash64(a[k].y)) ^ hash);
        z[i].x = u[i] ^ h;
        z[i].y = (uint64_t)k << 32 | (int32_t)u[i];
        k += (int32_t)u[i];
    }
    sort_128x(z, (size_t)n_u, tie_mode);
    for (i = 0; i < n_u >> 1; ++i) tmp = z[i], z[i] = z[n_u - 1 - i], z[n_u - 1 - i] = tmp;
    r = (orc_reg_t*)calloc(n_u, sizeof(orc_reg_t));
    for (i = 0; i < n_u; ++i) {
        orc_reg_t *ri = &r[i];
        ri->score = z[i].x >> 32;
        ri->hash = (uint32_t)z[i].x;
        ri->cnt = (int32_t)z[i].y;
        ri->div = -1.0f;
        reg_set_coor(ri, qlen, a, (int32_t)(z[i].y >> 32));
    }
    free(z);
    return r;
}

/* ------------------------------------------------------------------ mm_map_frag (single segment, no CIGAR) */

int orc_map(const orc_index_t *mi, const orc_opt_t *opt, const uint8_t *seq, int32_t qlen, const char *qname,
            orc_reg_t **regs, int32_t *n_regs, int32_t *rep_len_,
            orc128_t **anchors, int64_t *n_anchors, orc_stats_t *st) {
    v128_t mv = {0, 0, 0};
    int64_t n_a = 0;
    int rep_len = 0, n_regs0 = 0;
    uint64_t *u = 0;
    orc128_t *a;
    uint32_t hash;
    float chn_pen_gap, chn_pen_skip;
    int max_chain_gap_qry, max_chain_gap_ref;

    if (regs) *regs = 0;
    if (n_regs) *n_regs = 0;
    if (rep_len_) *rep_len_ = 0;
    if (anchors) *anchors = 0;
    if (n_anchors) *n_anchors = 0;
    if (qlen <= 0) return -1;   /* Aligner::map: "Sequence is empty" (aligner.rs:214-216) */

    hash = qname ? x31_hash_string(qname) : 0;
    hash ^= wang_hash((uint32_t)qlen) + wang_hash((uint32_t)opt->seed);
    hash = wang_hash(hash);

    mm_sketch(seq, qlen, mi->w, mi->k, 0, mi->is_hpc, &mv);     /* collect_minimizers; sdust_thres == 0 */
    if (st) st->n_q_minimizers += mv.n;
    if (opt->q_occ_frac > 0.0f) seed_mz_flt(&mv, opt->mid_occ, opt->q_occ_frac, st);
    a = collect_seed_hits(opt, opt->mid_occ, mi, qname, &mv, qlen, &n_a, &rep_len, st);
    free(mv.a);
    if (st) {
        st->n_anchors += n_a;
        int64_t ties = 0;
        for (int64_t i = 1; i < n_a; ++i) if (a[i].x == a[i - 1].x) ++ties;
        st->n_tie_x_anchors += ties;
        if (ties) ++st->n_tie_x_queries;
    }
    if (anchors) {
        *anchors = (orc128_t*)malloc((n_a ? n_a : 1) * sizeof(orc128_t));
        memcpy(*anchors, a, n_a * sizeof(orc128_t));
        *n_anchors = n_a;
    }

    max_chain_gap_qry = opt->max_gap;
    max_chain_gap_ref = opt->max_gap_ref > 0 ? opt->max_gap_ref : opt->max_gap;
    chn_pen_gap  = opt->chain_gap_scale * 0.01 * mi->k;
    chn_pen_skip = opt->chain_skip_scale * 0.01 * mi->k;
    a = lchain_dp(max_chain_gap_ref, max_chain_gap_qry, opt->bw, opt->max_chain_skip, opt->max_chain_iter, opt->min_cnt,
                  opt->min_chain_score, chn_pen_gap, chn_pen_skip, n_a, a, &n_regs0, &u, opt->tie_mode, st);
    /* bw_long <= bw or NO_LJOIN -> no long-join re-chain; max_occ (0) <= mid_occ -> no rescue re-chain */
    if (rep_len_) *rep_len_ = rep_len;
    if (st) st->n_chains += n_regs0;
    if (n_regs0 > 0) {
        orc_reg_t *r = gen_regs(hash, qlen, n_regs0, u, a, opt->tie_mode);
        /* chain_post: no-op under ALL_CHAINS; mm_est_err -> div (PAF only, not restated: stays -1);
         * mm_filter_strand_retained keeps all; align_regs returns early without CIGAR; mapq stays 0 */
        if (regs) *regs = r; else free(r);
        if (n_regs) *n_regs = n_regs0;
    }
    free(a); free(u);
    return 0;
}

/* ------------------------------------------------------------------ liblrge */

int orc_is_internal(int32_t qlen, int32_t qs, int32_t qe, int32_t rev, int32_t tlen, int32_t ts, int32_t te, float ratio) {
    int32_t overhang, maplen;
    float overhang_ratio;
#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))
    if (!rev) overhang = IMIN(qs, ts) + IMIN(qlen - qe, tlen - te);
    else overhang = IMIN(qs, tlen - te) + IMIN(qlen - qe, ts);
    maplen = IMAX(qe - qs, te - ts);
    overhang_ratio = (float)overhang / (float)maplen;
    return overhang_ratio < ratio;
}

float orc_per_read_estimate(uint64_t read_len, float avg_target_len, uint64_t n_target_reads, uint64_t n_ovlaps, uint32_t ovlap_thresh) {
    if (n_ovlaps == 0) return INFINITY;
    float ovlap_ratio = (float)n_target_reads / (float)n_ovlaps;
    return (float)read_len + ovlap_ratio * ((float)read_len + avg_target_len - 2.0f * (float)ovlap_thresh + 1.0f);
}

static int cmp_f32(const void *a, const void *b) { float x = *(const float*)a, y = *(const float*)b; return x < y ? -1 : x > y; }
static int calc_quantile(const float *data, int64_t n, float q, float *out) {
    if (n == 0) return 0;
    float pos = q * (float)(n - 1);
    int64_t idx = (int64_t)floorf(pos);
    float frac = pos - (float)idx;
    if (idx + 1 < n) *out = data[idx] * (1.0f - frac) + data[idx + 1] * frac;
    else *out = data[idx];
    return 1;
}
int orc_median(const float *vals, int64_t n, float lower_q, float upper_q, float *lower, float *median, float *upper) {
    int mask = 0;
    if (n == 0) return 0;
    float *v = (float*)malloc(n * 4);
    memcpy(v, vals, n * 4);
    qsort(v, n, 4, cmp_f32);
    if (calc_quantile(v, n, 0.5f, median)) mask |= 2;
    if (lower_q >= 0.0f && calc_quantile(v, n, lower_q, lower)) mask |= 1;
    if (upper_q >= 0.0f && calc_quantile(v, n, upper_q, upper)) mask |= 4;
    free(v);
    return mask;
}
int orc_estimate(const float *estimates, int64_t n, int finite, float lower_q, float upper_q,
                 float *lower, float *median, float *upper) {
    float *v = (float*)malloc((n ? n : 1) * 4);
    int64_t m = 0;
    for (int64_t i = 0; i < n; ++i) if (!finite || isfinite(estimates[i])) v[m++] = estimates[i];
    int mask = orc_median(v, m, lower_q, upper_q, lower, median, upper);
    free(v);
    return mask;
}

/* ---- parallel driver over the streamed read set (rayon par_bridge in the reference) */

typedef struct map_job_s map_job_t;
typedef void (*consume_fn)(map_job_t *job, uint32_t q, const orc_reg_t *regs, int32_t n_regs);
struct map_job_s {
    const orc_index_t *mi; const orc_opt_t *opt; const orc_readset_t *S;
    uint32_t next; pthread_mutex_t mu;     /* work counter + the reference's global locks */
    consume_fn consume;
    /* outputs */
    uint32_t *counts; uint8_t *had_mapping;
    int remove_internal; float ratio;
    orc_stats_t st;
    /* name handling */
    const uint32_t *name_id_idx;   /* name id of each indexed read */
    const uint32_t *name_id_s;     /* name id of each streamed read */
    /* ava */
    uint64_t *seen; size_t seen_cap, seen_n;
    /* regs collection */
    orc_reg_t *all_regs; uint32_t *all_q; int64_t n_all, m_all;
    int error;
};

static void *map_thread(void *arg) {
    map_job_t *j = (map_job_t*)arg;
    orc_stats_t st; memset(&st, 0, sizeof(st));
    for (;;) {
        uint32_t q0, q1;
        pthread_mutex_lock(&j->mu); q0 = j->next; j->next += 4; pthread_mutex_unlock(&j->mu);
        if (q0 >= j->S->n) break;
        q1 = q0 + 4 < j->S->n ? q0 + 4 : j->S->n;
        for (uint32_t q = q0; q < q1; ++q) {
            int32_t qlen = (int32_t)(j->S->off[q + 1] - j->S->off[q]);
            uint32_t nl = j->S->name_off[q + 1] - j->S->name_off[q];
            char *qname = (char*)malloc(nl + 1);
            orc_reg_t *regs = 0; int32_t n_regs = 0;
            memcpy(qname, j->S->names + j->S->name_off[q], nl); qname[nl] = 0;
            if (orc_map(j->mi, j->opt, j->S->bases + j->S->off[q], qlen, qname, &regs, &n_regs, 0, 0, 0, &st) != 0) {
                j->error = 1;   /* LrgeError::MapError (twoset.rs:279-284) */
            } else {
                j->consume(j, q, regs, n_regs);
            }
            free(regs); free(qname);
        }
    }
    pthread_mutex_lock(&j->mu);
    j->st.n_q_minimizers += st.n_q_minimizers; j->st.n_seeds += st.n_seeds; j->st.n_hits += st.n_hits;
    j->st.n_anchors += st.n_anchors; j->st.n_chains += st.n_chains; j->st.n_dp_iter += st.n_dp_iter;
    j->st.n_tie_x_queries += st.n_tie_x_queries; j->st.n_tie_x_anchors += st.n_tie_x_anchors;
    j->st.n_mz_flt_deleted += st.n_mz_flt_deleted; j->st.n_max_iter_clamp += st.n_max_iter_clamp;
    j->st.n_skip_break += st.n_skip_break; j->st.n_occ_filtered += st.n_occ_filtered;
    pthread_mutex_unlock(&j->mu);
    return 0;
}

static void run_map_job(map_job_t *job, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    job->next = 0; pthread_mutex_init(&job->mu, 0);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], 0, map_thread, job);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], 0);
    pthread_mutex_destroy(&job->mu); free(th);
}

/* names -> dense ids with equal names sharing an id (HashSet<Vec<u8>> semantics of the reference) */
typedef struct { const char *s; uint32_t len; uint32_t set, idx; } name_ref_t;
static int cmp_name_ref(const void *a, const void *b) {
    const name_ref_t *x = (const name_ref_t*)a, *y = (const name_ref_t*)b;
    uint32_t m = x->len < y->len ? x->len : y->len;
    int c = memcmp(x->s, y->s, m);
    if (c) return c;
    return x->len < y->len ? -1 : x->len > y->len;
}
static void name_ids(const orc_readset_t *A, const orc_readset_t *B, uint32_t *idA, uint32_t *idB) {
    size_t n = (size_t)A->n + (B ? B->n : 0), i;
    name_ref_t *v = (name_ref_t*)malloc((n ? n : 1) * sizeof(name_ref_t));
    for (i = 0; i < A->n; ++i) { v[i].s = A->names + A->name_off[i]; v[i].len = A->name_off[i + 1] - A->name_off[i]; v[i].set = 0; v[i].idx = (uint32_t)i; }
    if (B) for (i = 0; i < B->n; ++i) { name_ref_t *e = &v[A->n + i]; e->s = B->names + B->name_off[i]; e->len = B->name_off[i + 1] - B->name_off[i]; e->set = 1; e->idx = (uint32_t)i; }
    qsort(v, n, sizeof(name_ref_t), cmp_name_ref);
    uint32_t id = 0;
    for (i = 0; i < n; ++i) {
        if (i > 0 && cmp_name_ref(&v[i - 1], &v[i]) != 0) ++id;
        if (v[i].set == 0) idA[v[i].idx] = id; else idB[v[i].idx] = id;
    }
    free(v);
}

/* twoset.rs:286-309 */
static void consume_twoset(map_job_t *j, uint32_t q, const orc_reg_t *regs, int32_t n_regs) {
    int32_t qlen = (int32_t)(j->S->off[q + 1] - j->S->off[q]);
    uint32_t n_unique = 0, *uniq = (uint32_t*)malloc((n_regs ? n_regs : 1) * 4);
    if (n_regs > 0) {
        for (int32_t i = 0; i < n_regs; ++i) {
            const orc_reg_t *r = &regs[i];
            if (j->remove_internal && orc_is_internal(qlen, r->qs, r->qe, r->rev, (int32_t)j->mi->len[r->rid], r->rs, r->re, j->ratio)) continue;
            uint32_t id = j->name_id_idx[r->rid], k;
            for (k = 0; k < n_unique; ++k) if (uniq[k] == id) break;
            if (k == n_unique) uniq[n_unique++] = id;
        }
        j->had_mapping[q] = 1;
    } else j->had_mapping[q] = 0;
    j->counts[q] = n_unique;
    free(uniq);
}

int orc_twoset_counts(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt,
                      int remove_internal, float max_overhang_ratio, int n_threads,
                      uint32_t *counts, uint8_t *had_mapping, orc_stats_t *st) {
    map_job_t job; memset(&job, 0, sizeof(job));
    double t0 = now_s();
    orc_index_t *mi = orc_index_build(T, opt, n_threads);
    double t1 = now_s();
    uint32_t *idT = (uint32_t*)malloc((T->n ? T->n : 1) * 4), *idQ = (uint32_t*)malloc((Q->n ? Q->n : 1) * 4);
    name_ids(T, Q, idT, idQ);
    job.mi = mi; job.opt = opt; job.S = Q; job.consume = consume_twoset;
    job.counts = counts; job.had_mapping = had_mapping; job.remove_internal = remove_internal; job.ratio = max_overhang_ratio;
    job.name_id_idx = idT; job.name_id_s = idQ;
    run_map_job(&job, n_threads);
    double t2 = now_s();
    if (st) {
        *st = job.st;
        st->n_t_minimizers = orc_index_n_minimizers(mi); st->n_keys = orc_index_n_keys(mi); st->mid_occ = opt->mid_occ;
        st->t_index_s = t1 - t0; st->t_map_s = t2 - t1;
        st->n_overlaps = 0; st->no_mapping_count = 0;
        for (uint32_t q = 0; q < Q->n; ++q) { st->n_overlaps += counts[q]; if (!had_mapping[q]) ++st->no_mapping_count; }
    }
    orc_index_destroy(mi); free(idT); free(idQ);
    return job.error ? -1 : 0;
}

/* twoset.rs:477-525: index = query set, stream targets; inline overhang test (sense opposite to is_internal) */
static void consume_inverse(map_job_t *j, uint32_t t, const orc_reg_t *regs, int32_t n_regs) {
    int32_t tlen_stream = (int32_t)(j->S->off[t + 1] - j->S->off[t]);
    uint32_t n_unique = 0, *uniq = (uint32_t*)malloc((n_regs ? n_regs : 1) * 4);
    if (n_regs > 0) {
        pthread_mutex_lock(&j->mu);
        for (int32_t i = 0; i < n_regs; ++i) {
            const orc_reg_t *r = &regs[i];
            uint32_t id = j->name_id_idx[r->rid], k;
            for (k = 0; k < n_unique; ++k) if (uniq[k] == id) break;
            if (k < n_unique) continue;
            if (j->remove_internal) {
                /* PafRecord: query = streamed read, target = indexed read */
                int32_t qs = r->qs, qe = r->qe, ql = tlen_stream, ts = r->rs, te = r->re, tl = (int32_t)j->mi->len[r->rid];
                int32_t overhang, maplen;
                if (!r->rev) overhang = IMIN(qs, ts) + IMIN(ql - qe, tl - te);
                else overhang = IMIN(qs, tl - te) + IMIN(ql - qe, ts);
                maplen = IMAX(qe - qs, te - ts);
                if (overhang > (int32_t)((float)maplen * j->ratio)) continue;
            }
            j->counts[r->rid] += 1;   /* names of indexed reads are unique here (DuplicateReadIdentifier otherwise) */
            uniq[n_unique++] = id;
        }
        pthread_mutex_unlock(&j->mu);
    }
    free(uniq);
}

static int has_dup_ids(const uint32_t *ids, uint32_t n) {
    uint32_t *s = (uint32_t*)malloc((n ? n : 1) * 4); int dup = 0;
    memcpy(s, ids, (size_t)n * 4); qsort(s, n, 4, cmp_u32);
    for (uint32_t i = 1; i < n; ++i) if (s[i] == s[i - 1]) { dup = 1; break; }
    free(s); return dup;
}

int orc_twoset_inverse_counts(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt,
                              int remove_internal, float max_overhang_ratio, int n_threads,
                              uint32_t *counts, orc_stats_t *st) {
    map_job_t job; memset(&job, 0, sizeof(job));
    double t0 = now_s();
    uint32_t *idT = (uint32_t*)malloc((T->n ? T->n : 1) * 4), *idQ = (uint32_t*)malloc((Q->n ? Q->n : 1) * 4);
    name_ids(T, Q, idT, idQ);
    if (has_dup_ids(idQ, Q->n)) { free(idT); free(idQ); return -9; }   /* LrgeError::DuplicateReadIdentifier (twoset.rs:439-449) */
    orc_index_t *mi = orc_index_build(Q, opt, n_threads);
    double t1 = now_s();
    memset(counts, 0, (size_t)Q->n * 4);
    job.mi = mi; job.opt = opt; job.S = T; job.consume = consume_inverse;
    job.counts = counts; job.remove_internal = remove_internal; job.ratio = max_overhang_ratio;
    job.name_id_idx = idQ; job.name_id_s = idT;
    run_map_job(&job, n_threads);
    double t2 = now_s();
    if (st) {
        *st = job.st;
        st->n_t_minimizers = orc_index_n_minimizers(mi); st->n_keys = orc_index_n_keys(mi); st->mid_occ = opt->mid_occ;
        st->t_index_s = t1 - t0; st->t_map_s = t2 - t1;
        st->n_overlaps = 0; st->no_mapping_count = 0;
        for (uint32_t q = 0; q < Q->n; ++q) { st->n_overlaps += counts[q]; if (counts[q] == 0) ++st->no_mapping_count; }
    }
    orc_index_destroy(mi); free(idT); free(idQ);
    return job.error ? -1 : 0;
}

/* ava.rs:265-306; pair set keyed by (min id, max id) */
static int seen_insert(map_job_t *j, uint64_t key) { /* returns 1 if newly inserted */
    if ((j->seen_n + 1) * 2 > j->seen_cap) {
        size_t ncap = j->seen_cap ? j->seen_cap * 2 : 1024;
        uint64_t *ns = (uint64_t*)malloc(ncap * 8);
        memset(ns, 0xff, ncap * 8);
        for (size_t i = 0; i < j->seen_cap; ++i) if (j->seen[i] != UINT64_MAX) {
            size_t s = (j->seen[i] * 0x9E3779B97F4A7C15ULL >> 20) & (ncap - 1);
            while (ns[s] != UINT64_MAX) s = (s + 1) & (ncap - 1);
            ns[s] = j->seen[i];
        }
        free(j->seen); j->seen = ns; j->seen_cap = ncap;
    }
    size_t s = (key * 0x9E3779B97F4A7C15ULL >> 20) & (j->seen_cap - 1);
    while (j->seen[s] != UINT64_MAX) { if (j->seen[s] == key) return 0; s = (s + 1) & (j->seen_cap - 1); }
    j->seen[s] = key; ++j->seen_n;
    return 1;
}
static void consume_ava(map_job_t *j, uint32_t q, const orc_reg_t *regs, int32_t n_regs) {
    int32_t qlen = (int32_t)(j->S->off[q + 1] - j->S->off[q]);
    if (n_regs <= 0) return;
    pthread_mutex_lock(&j->mu);
    for (int32_t i = 0; i < n_regs; ++i) {
        const orc_reg_t *r = &regs[i];
        uint32_t qid = j->name_id_s[q], tid = j->name_id_idx[r->rid];
        if (qid == tid) continue;   /* self overlap */
        if (j->remove_internal && orc_is_internal(qlen, r->qs, r->qe, r->rev, (int32_t)j->mi->len[r->rid], r->rs, r->re, j->ratio)) continue;
        uint64_t key = qid < tid ? (uint64_t)qid << 32 | tid : (uint64_t)tid << 32 | qid;
        if (!seen_insert(j, key)) continue;
        j->counts[r->rid] += 1;
        j->counts[q] += 1;
    }
    pthread_mutex_unlock(&j->mu);
}

int orc_ava_counts(const orc_readset_t *R, orc_opt_t *opt, int remove_internal, float max_overhang_ratio,
                   int n_threads, uint32_t *counts, orc_stats_t *st) {
    map_job_t job; memset(&job, 0, sizeof(job));
    double t0 = now_s();
    uint32_t *id = (uint32_t*)malloc((R->n ? R->n : 1) * 4);
    name_ids(R, 0, id, 0);
    if (has_dup_ids(id, R->n)) { free(id); return -9; }   /* LrgeError::DuplicateReadIdentifier (ava.rs:195-199) */
    orc_index_t *mi = orc_index_build(R, opt, n_threads);
    double t1 = now_s();
    memset(counts, 0, (size_t)R->n * 4);
    job.mi = mi; job.opt = opt; job.S = R; job.consume = consume_ava;
    job.counts = counts; job.remove_internal = remove_internal; job.ratio = max_overhang_ratio;
    job.name_id_idx = id; job.name_id_s = id;
    run_map_job(&job, n_threads);
    double t2 = now_s();
    if (st) {
        *st = job.st;
        st->n_t_minimizers = orc_index_n_minimizers(mi); st->n_keys = orc_index_n_keys(mi); st->mid_occ = opt->mid_occ;
        st->t_index_s = t1 - t0; st->t_map_s = t2 - t1;
        st->n_overlaps = job.seen_n; st->no_mapping_count = 0;
        for (uint32_t q = 0; q < R->n; ++q) if (counts[q] == 0) ++st->no_mapping_count;
    }
    free(job.seen); orc_index_destroy(mi); free(id);
    return job.error ? -1 : 0;
}

static void consume_regs(map_job_t *j, uint32_t q, const orc_reg_t *regs, int32_t n_regs) {
    if (n_regs <= 0) return;
    pthread_mutex_lock(&j->mu);
    if (j->n_all + n_regs > j->m_all) {
        j->m_all = (j->n_all + n_regs) * 2;
        j->all_regs = (orc_reg_t*)realloc(j->all_regs, j->m_all * sizeof(orc_reg_t));
        j->all_q = (uint32_t*)realloc(j->all_q, j->m_all * 4);
    }
    for (int32_t i = 0; i < n_regs; ++i) { j->all_regs[j->n_all] = regs[i]; j->all_q[j->n_all++] = q; }
    pthread_mutex_unlock(&j->mu);
}

int orc_twoset_regs(const orc_readset_t *T, const orc_readset_t *Q, orc_opt_t *opt, int n_threads,
                    orc_reg_t **regs, uint32_t **qidx, int64_t *n_regs) {
    map_job_t job; memset(&job, 0, sizeof(job));
    orc_index_t *mi = orc_index_build(T, opt, n_threads);
    job.mi = mi; job.opt = opt; job.S = Q; job.consume = consume_regs;
    run_map_job(&job, n_threads);
    *regs = job.all_regs; *qidx = job.all_q; *n_regs = job.n_all;
    orc_index_destroy(mi);
    return job.error ? -1 : 0;
}
