// host_api.cpp -- host-side arithmetic of liblrge that follows the hot path (f32, evaluation order of the
// reference): per_read_estimate (estimate.rs:142-157), median / calculate_quantile / Estimate::estimate
// (estimate.rs:55-132), PafRecord::is_internal (mapping.rs:59-77), and the generate_estimates wrappers
// (twoset.rs:587-606, ava.rs:369-381) over the GPU counts.
#include "../../include/lrge_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <cctype>
#include <vector>

#include <zlib.h>
#include <string>

namespace {

// Line reader over a plain or gzip file (zlib reads both); other compressions are recognised by their magic bytes
// (io.rs:35-61) and refused.
struct LineReader {
    gzFile f = nullptr;
    std::vector<char> buf;
    size_t pos = 0, len = 0;
    int status = LRGE_OK;
    explicit LineReader(const char *path) : buf(1 << 20) {
        FILE *fp = fopen(path, "rb");
        if (!fp) { status = LRGE_ERR_IO; return; }
        unsigned char m[5] = {0, 0, 0, 0, 0};
        const size_t n = fread(m, 1, 5, fp);
        fclose(fp);
        if (n >= 2 && m[0] == 0x42 && m[1] == 0x5a) { status = LRGE_ERR_UNSUPPORTED; return; }                         // bzip2
        if (n >= 4 && m[0] == 0x28 && m[1] == 0xb5 && m[2] == 0x2f && m[3] == 0xfd) { status = LRGE_ERR_UNSUPPORTED; return; }   // zstd
        if (n >= 5 && m[0] == 0xfd && m[1] == 0x37 && m[2] == 0x7a && m[3] == 0x58 && m[4] == 0x5a) { status = LRGE_ERR_UNSUPPORTED; return; }   // xz
        f = gzopen(path, "rb");
        if (!f) status = LRGE_ERR_IO;
        else gzbuffer(f, 1 << 20);
    }
    ~LineReader() { if (f) gzclose(f); }
    bool getline(std::string &out) {
        out.clear();
        for (;;) {
            if (pos == len) {
                const int n = gzread(f, buf.data(), (unsigned)buf.size());
                if (n < 0) { status = LRGE_ERR_IO; return false; }
                if (n == 0) return !out.empty();
                pos = 0; len = (size_t)n;
            }
            const char *p = buf.data() + pos;
            const char *nl = (const char *)memchr(p, '\n', len - pos);
            if (nl) { out.append(p, nl - p); pos = (size_t)(nl - buf.data()) + 1; break; }
            out.append(p, len - pos);
            pos = len;
        }
        if (!out.empty() && out.back() == '\r') out.pop_back();
        return true;
    }
};

// Calls rec(id, id_len, seq, seq_len) for every FASTA / FASTQ record; the id is the header up to the first
// whitespace (io.rs:192-205).
template <typename F> int for_each_record(const char *path, F rec) {
    LineReader in(path);
    if (in.status != LRGE_OK) return in.status;
    std::string line, name, seq, plus, qual;
    bool have = false, fastq = false, first = true;
    auto id_len = [](const std::string &h) { size_t e = 1; while (e < h.size() && !isspace((unsigned char)h[e])) ++e; return e - 1; };
    while (in.getline(line)) {
        if (line.empty()) continue;
        if (first) { first = false; if (line[0] == '@') fastq = true; else if (line[0] != '>') return LRGE_ERR_FASTQ_PARSE; }
        if (fastq) {
            if (line[0] != '@') return LRGE_ERR_FASTQ_PARSE;
            name = line;
            if (!in.getline(seq) || !in.getline(plus) || !in.getline(qual) || plus.empty() || plus[0] != '+' || qual.size() != seq.size())
                return LRGE_ERR_FASTQ_PARSE;
            rec(name.data() + 1, id_len(name), seq.data(), seq.size());
        } else if (line[0] == '>') {
            if (have) rec(name.data() + 1, id_len(name), seq.data(), seq.size());
            name = line; seq.clear(); have = true;
        } else {
            if (!have) return LRGE_ERR_FASTQ_PARSE;
            seq += line;
        }
    }
    if (in.status != LRGE_OK) return in.status;
    if (have && !fastq) rec(name.data() + 1, id_len(name), seq.data(), seq.size());
    return LRGE_OK;
}

}  // namespace

extern "C" {

// io::count_records (io.rs:121-147): number of records, plus the bases and name bytes a loader needs to allocate
int lrge_fastx_count(const char *path, uint64_t *n_records, uint64_t *n_bases, uint64_t *n_name_bytes) {
    if (!path) return LRGE_ERR_INVALID_ARG;
    uint64_t n = 0, b = 0, nb = 0;
    const int rc = for_each_record(path, [&](const char *, size_t il, const char *, size_t sl) { ++n; b += sl; nb += il; });
    if (rc != LRGE_OK) return rc;
    if (n == 0) return LRGE_ERR_IO;                       // "Is the file empty?" (io.rs:140-145)
    if (n_records) *n_records = n;
    if (n_bases) *n_bases = b;
    if (n_name_bytes) *n_name_bytes = nb;
    return LRGE_OK;
}

// Loads every record into caller buffers sized from lrge_fastx_count: bases[cap_bases], off[cap_records + 1],
// names[cap_name_bytes], name_off[cap_records + 1] -- the lrge_readset layout.  The file is parsed again, so every write is
// checked against the capacities (a pipe or a file still being written may have grown): LRGE_ERR_INVALID_ARG if a
// record does not fit, or if the names exceed the 32-bit offsets of lrge_readset.
int lrge_fastx_load(const char *path, uint8_t *bases, uint64_t cap_bases, uint64_t *off, uint64_t cap_records,
                    char *names, uint64_t cap_name_bytes, uint32_t *name_off) {
    if (!path || !bases || !off || !names || !name_off) return LRGE_ERR_INVALID_ARG;
    uint64_t n = 0, b = 0, nb = 0;
    bool fits = true;
    off[0] = 0; name_off[0] = 0;
    const int rc = for_each_record(path, [&](const char *id, size_t il, const char *seq, size_t sl) {
        if (!fits) return;
        if (n + 1 > cap_records || b + sl > cap_bases || nb + il > cap_name_bytes || nb + il > 0xffffffffull) { fits = false; return; }
        memcpy(bases + b, seq, sl); b += sl;
        memcpy(names + nb, id, il); nb += il;
        ++n; off[n] = b; name_off[n] = (uint32_t)nb;
    });
    if (rc != LRGE_OK) return rc;
    return fits ? LRGE_OK : LRGE_ERR_INVALID_ARG;
}

// ---- two-pass subsampled ingestion (TwoSetStrategy::split_fastq, twoset.rs:122-201; AvaStrategy::subsample_reads,
// ava.rs:108-161): pass 1 = lrge_fastx_count, the caller draws its record indices, pass 2 = this function, which keeps
// only the selected records and writes them straight into PINNED host buffers (what the H2D copies of the *_counts calls
// want), without ever holding the whole file in memory.  Up to two disjoint selections are filled in the same pass
// (target and query set).  Indices must be ascending and unique (liblrge sorts nothing, it tests HashSet membership;
// here the order of a set is the file order, as in the temp FASTA liblrge writes).
struct lrge_hostset {
    uint8_t *bases = nullptr; uint64_t *off = nullptr; char *names = nullptr; uint32_t *name_off = nullptr;
    uint64_t cap_bases = 0, cap_names = 0, n = 0, cap_n = 0;
    bool pinned = false;
    lrge_readset view{};
};

namespace {

void *hs_alloc(size_t bytes, bool pinned) {
    void *p = nullptr;
    if (pinned) { if (lrge_gpu_host_alloc(bytes ? bytes : 1, &p) != LRGE_OK) p = nullptr; }
    else p = malloc(bytes ? bytes : 1);
    return p;
}
void hs_free(void *p, bool pinned) { if (!p) return; if (pinned) lrge_gpu_host_free(p); else free(p); }

bool hs_grow(lrge_hostset *h, uint64_t need_bases, uint64_t need_names) {
    if (need_bases > h->cap_bases) {
        const uint64_t cap = std::max<uint64_t>(need_bases + need_bases / 4, 1 << 20);
        uint8_t *nb = (uint8_t *)hs_alloc(cap, h->pinned);
        if (!nb) return false;
        if (h->bases) { memcpy(nb, h->bases, h->off[h->n]); hs_free(h->bases, h->pinned); }
        h->bases = nb; h->cap_bases = cap;
    }
    if (need_names > h->cap_names) {
        const uint64_t cap = std::max<uint64_t>(need_names + need_names / 4, 1 << 16);
        char *nn = (char *)hs_alloc(cap, h->pinned);
        if (!nn) return false;
        if (h->names) { memcpy(nn, h->names, h->name_off[h->n]); hs_free(h->names, h->pinned); }
        h->names = nn; h->cap_names = cap;
    }
    return true;
}

lrge_hostset *hs_new(uint64_t n, uint64_t hint_bases, bool pinned) {
    lrge_hostset *h = new lrge_hostset();
    h->pinned = pinned; h->cap_n = n;
    h->off = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    h->name_off = (uint32_t *)malloc((n + 1) * sizeof(uint32_t));
    if (!h->off || !h->name_off) { free(h->off); free(h->name_off); delete h; return nullptr; }
    h->off[0] = 0; h->name_off[0] = 0;
    if (!hs_grow(h, std::max<uint64_t>(hint_bases, 1), std::max<uint64_t>(n * 40, 1))) { free(h->off); free(h->name_off); delete h; return nullptr; }
    return h;
}

}  // namespace

void lrge_hostset_free(lrge_hostset *h) {
    if (!h) return;
    hs_free(h->bases, h->pinned); hs_free(h->names, h->pinned);
    free(h->off); free(h->name_off);
    delete h;
}

const lrge_readset *lrge_hostset_view(lrge_hostset *h) {
    if (!h) return nullptr;
    h->view.bases = h->bases; h->view.off = h->off; h->view.names = h->names; h->view.name_off = h->name_off; h->view.n = (uint32_t)h->n;
    return &h->view;
}

int lrge_fastx_load_subsets(const char *path, const uint64_t *idx_a, uint64_t n_a, const uint64_t *idx_b, uint64_t n_b,
                            uint64_t hint_bases_per_record, int pinned, lrge_hostset **out_a, lrge_hostset **out_b) {
    if (!path || !out_a || (n_a && !idx_a) || (n_b && (!idx_b || !out_b))) return LRGE_ERR_INVALID_ARG;
    for (uint64_t i = 1; i < n_a; ++i) if (idx_a[i] <= idx_a[i - 1]) return LRGE_ERR_INVALID_ARG;
    for (uint64_t i = 1; i < n_b; ++i) if (idx_b[i] <= idx_b[i - 1]) return LRGE_ERR_INVALID_ARG;
    if (n_a >= (1ull << 32) || n_b >= (1ull << 32)) return LRGE_ERR_TOO_MANY_READS;
    const uint64_t hint = hint_bases_per_record ? hint_bases_per_record : 10000;
    bool pin = pinned != 0;
    if (pin) {                                   // no usable CUDA runtime: plain memory (the GPU call will fail loudly later)
        void *probe = nullptr;
        if (lrge_gpu_host_alloc(16, &probe) == LRGE_OK) lrge_gpu_host_free(probe); else pin = false;
    }
    lrge_hostset *A = hs_new(n_a, n_a * hint + n_a * hint / 8, pin);
    lrge_hostset *B = out_b ? hs_new(n_b, n_b * hint + n_b * hint / 8, pin) : nullptr;
    if (!A || (out_b && !B)) { lrge_hostset_free(A); lrge_hostset_free(B); return LRGE_ERR_OUT_OF_MEMORY; }
    uint64_t rec = 0, ia = 0, ib = 0;
    bool oom = false, big_names = false;
    auto push = [&](lrge_hostset *h, const char *id, size_t il, const char *seq, size_t sl) {
        const uint64_t b = h->off[h->n], nb = h->name_off[h->n];
        if (nb + il > 0xffffffffull) { big_names = true; return; }
        if (!hs_grow(h, b + sl, nb + il)) { oom = true; return; }
        memcpy(h->bases + b, seq, sl);
        memcpy(h->names + nb, id, il);
        ++h->n; h->off[h->n] = b + sl; h->name_off[h->n] = (uint32_t)(nb + il);
    };
    const int rc = for_each_record(path, [&](const char *id, size_t il, const char *seq, size_t sl) {
        if (!oom && !big_names) {
            if (ia < n_a && idx_a[ia] == rec) { push(A, id, il, seq, sl); ++ia; }
            else if (ib < n_b && idx_b[ib] == rec) { push(B, id, il, seq, sl); ++ib; }
        }
        ++rec;
    });
    int err = rc;
    if (err == LRGE_OK && oom) err = LRGE_ERR_OUT_OF_MEMORY;
    if (err == LRGE_OK && big_names) err = LRGE_ERR_INVALID_ARG;
    if (err == LRGE_OK && (ia != n_a || ib != n_b)) err = LRGE_ERR_INVALID_ARG;     // an index beyond the last record
    if (err != LRGE_OK) { lrge_hostset_free(A); lrge_hostset_free(B); return err; }
    *out_a = A;
    if (out_b) *out_b = B;
    return LRGE_OK;
}

float lrge_per_read_estimate(uint64_t read_len, float avg_target_len, uint64_t n_target_reads, uint64_t n_ovlaps,
                             uint32_t ovlap_thresh) {
    if (n_ovlaps == 0) return std::numeric_limits<float>::infinity();
    volatile float ovlap_ratio = (float)n_target_reads / (float)n_ovlaps;
    volatile float a = (float)read_len + avg_target_len;
    volatile float two_t = 2.0f * (float)ovlap_thresh;
    volatile float b = a - two_t;
    volatile float c = b + 1.0f;
    volatile float d = ovlap_ratio * c;
    volatile float e = (float)read_len + d;
    return e;
}

static bool quantile(const std::vector<float> &data, float q, float *out) {
    if (data.empty()) return false;
    const size_t n = data.size();
    volatile float pos = q * (float)(n - 1);
    const size_t idx = (size_t)std::floor((float)pos);
    volatile float frac = pos - (float)idx;
    if (idx + 1 < n) {
        volatile float w0 = 1.0f - frac;
        volatile float t0 = data[idx] * w0;
        volatile float t1 = data[idx + 1] * frac;
        volatile float r = t0 + t1;
        *out = r;
    } else *out = data[idx];
    return true;
}

int lrge_estimate(const float *estimates, uint64_t n, int finite, float lower_q, float upper_q, float *lower, float *median,
                  float *upper) {
    // calculate_quantile panics outside [0, 1] (estimate.rs:118-120)
    if ((lower_q >= 0.0f && lower_q > 1.0f) || (upper_q >= 0.0f && upper_q > 1.0f)) return -1;
    std::vector<float> v;
    v.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        if (!finite || std::isfinite(estimates[i])) v.push_back(estimates[i]);
    if (v.empty()) return 0;
    std::sort(v.begin(), v.end());   // partial_cmp: no NaN expected (estimate.rs:92-93)
    int mask = 0;
    float tmp;
    if (quantile(v, 0.5f, &tmp)) { if (median) *median = tmp; mask |= 2; }
    if (lower_q >= 0.0f && quantile(v, lower_q, &tmp)) { if (lower) *lower = tmp; mask |= 1; }
    if (upper_q >= 0.0f && quantile(v, upper_q, &tmp)) { if (upper) *upper = tmp; mask |= 4; }
    return mask;
}

int lrge_is_internal(int32_t qlen, int32_t qs, int32_t qe, int32_t rev, int32_t tlen, int32_t ts, int32_t te, float ratio) {
    int32_t overhang;
    if (!rev) overhang = std::min(qs, ts) + std::min(qlen - qe, tlen - te);
    else overhang = std::min(qs, tlen - te) + std::min(qlen - qe, ts);
    const int32_t maplen = std::max(qe - qs, te - ts);
    volatile float r = (float)overhang / (float)maplen;
    return r < ratio;
}

int lrge_paf_format(const lrge_overlap *o, const char *qname, int qname_len, int32_t qlen, const char *tname, int tname_len,
                    int32_t tlen, float dv, int32_t rl, char *buf, size_t cap) {
    if (!o || !buf || cap == 0) return -1;
    // serialize_dv (mapping.rs:135-147): "0" below f32::EPSILON, else 4 decimals
    char dvs[32];
    if (dv < 1.1920929e-07f) snprintf(dvs, sizeof dvs, "0");
    else snprintf(dvs, sizeof dvs, "%.4f", (double)dv);
    const int n = snprintf(buf, cap, "%.*s\t%d\t%d\t%d\t%c\t%.*s\t%d\t%d\t%d\t%d\t%d\t0\ttp:A:S\tcm:i:%d\ts1:i:%d\tdv:f:%s\trl:i:%d",
                           qname_len, qname, qlen, o->qs, o->qe, o->rev ? '-' : '+', tname_len, tname, tlen, o->ts, o->te, o->mlen,
                           o->blen, o->cnt, o->score, dvs, rl);
    return n < 0 ? -1 : (n >= (int)cap ? (int)cap - 1 : n);
}

int lrge_twoset_generate_estimates(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                                   float *estimates, uint32_t *no_mapping_count, lrge_stats *stats) {
    if (!T || !Q || !estimates || !opts) return LRGE_ERR_INVALID_ARG;
    std::vector<uint32_t> counts(Q->n);
    std::vector<uint8_t> had(Q->n);
    lrge_stats st;
    memset(&st, 0, sizeof st);
    const int rc = lrge_gpu_twoset_counts(ctx, T, Q, opts, counts.data(), had.data(), &st);
    if (rc != LRGE_OK) return rc;
    // twoset.rs:191: avg_target_len = sum_target_len as f32 / target_num_reads as f32
    const uint64_t sum_t = T->n ? T->off[T->n] - T->off[0] : 0;
    const float avg_target_len = (float)sum_t / (float)T->n;
    for (uint32_t q = 0; q < Q->n; ++q)   // twoset.rs:311-317; overlap threshold = min_chain_score (twoset.rs:213)
        estimates[q] = lrge_per_read_estimate(Q->off[q + 1] - Q->off[q], avg_target_len, T->n, counts[q], 100);
    if (no_mapping_count) *no_mapping_count = st.no_mapping_count;
    if (stats) *stats = st;
    return LRGE_OK;
}

int lrge_ava_generate_estimates(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts, float *estimates,
                                uint32_t *no_mapping_count, lrge_stats *stats) {
    if (!R || !estimates || !opts) return LRGE_ERR_INVALID_ARG;
    std::vector<uint32_t> counts(R->n);
    lrge_stats st;
    memset(&st, 0, sizeof st);
    const int rc = lrge_gpu_ava_counts(ctx, R, opts, counts.data(), &st);
    if (rc != LRGE_OK) return rc;
    // ava.rs:339-346: avg_read_len = sum_len / (n - 1), n_target_reads = n - 1 (sum_len includes the read itself)
    const uint64_t sum_len = R->n ? R->off[R->n] - R->off[0] : 0;
    const float avg_read_len = (float)sum_len / (float)(R->n - 1);
    uint32_t nomap = 0;
    for (uint32_t i = 0; i < R->n; ++i) {
        if (counts[i] == 0) { ++nomap; estimates[i] = std::numeric_limits<float>::infinity(); }
        else estimates[i] = lrge_per_read_estimate(R->off[i + 1] - R->off[i], avg_read_len, (uint64_t)R->n - 1, counts[i], 100);
    }
    if (no_mapping_count) *no_mapping_count = nomap;
    if (stats) *stats = st;
    return LRGE_OK;
}

}  // extern "C"
