// lrge_b200.hpp -- C++ host mirror of liblrge's public surface over the C ABI (lrge_b200.h).
//
// The reference's host language is Rust; this image has no Rust toolchain, so the host side above the
// C ABI is C++ with the same names, argument meaning and error behaviour:
//   lrge::Estimate            trait Estimate              (liblrge/src/estimate.rs:21-78)
//   lrge::EstimateResult      struct EstimateResult       (estimate.rs:8-17)
//   lrge::TwoSetStrategy      struct TwoSetStrategy       (twoset.rs:73-99) + twoset::Builder (twoset/builder.rs)
//   lrge::AvaStrategy         struct AvaStrategy          (ava.rs:66-84)    + ava::Builder    (ava/builder.rs)
//   lrge::Platform            enum Platform               (lib.rs:163-180)
//   lrge::LrgeError           enum LrgeError              (error.rs:6-33)
// The strategies take explicit read sets (what split_fastq / subsample_reads leave in target.fa / query.fa /
// reads.fa, twoset.rs:157-186, ava.rs:139-156); the random subsampling before the hot path stays with the caller.
#pragma once
#include "lrge_b200.h"

#include <cctype>
#include <cstdint>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace lrge {

constexpr float LOWER_QUANTILE = 0.15f;                 // estimate.rs:4
constexpr float UPPER_QUANTILE = 0.65f;                 // estimate.rs:6
constexpr size_t DEFAULT_TARGET_NUM_READS = 10000;      // twoset.rs:65
constexpr size_t DEFAULT_QUERY_NUM_READS = 5000;        // twoset.rs:66
constexpr size_t DEFAULT_AVA_NUM_READS = 25000;         // ava.rs:62

enum class Platform { PacBio, Nanopore };

struct LrgeError : std::runtime_error {
    lrge_status code;
    LrgeError(lrge_status c, const std::string &msg) : std::runtime_error(msg), code(c) {}
};

inline Platform platform_from_str(std::string s) {      // Platform::from_str (lib.rs:170-180)
    for (auto &c : s) c = (char)tolower(c);
    if (s == "pacbio" || s == "pb") return Platform::PacBio;
    if (s == "nanopore" || s == "ont") return Platform::Nanopore;
    throw LrgeError(LRGE_ERR_INVALID_PLATFORM, "Invalid platform: " + s);
}

// Reads in the layout the C ABI takes.  Either owns its storage (vectors below) or wraps a pinned lrge_hostset filled by
// the two-pass subsampled loader (from_fastx_subsets).
struct ReadSet {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> off{0};
    std::string names;
    std::vector<uint32_t> name_off{0};
    std::shared_ptr<lrge_hostset> host;      // set: the buffers live in (pinned) library memory, the vectors are unused

    uint32_t size() const { return host ? lrge_hostset_view(host.get())->n : (uint32_t)(off.size() - 1); }
    uint64_t total_bases() const { const lrge_readset v = view(); return v.n ? v.off[v.n] - v.off[0] : 0; }
    void push(const std::string &name, const std::string &seq) {
        bases.insert(bases.end(), seq.begin(), seq.end());
        off.push_back(bases.size());
        names += name;
        name_off.push_back((uint32_t)names.size());
    }
    lrge_readset view() const {
        if (host) return *lrge_hostset_view(host.get());
        lrge_readset v;
        v.bases = bases.data(); v.off = off.data(); v.names = names.data(); v.name_off = name_off.data(); v.n = size();
        return v;
    }
    // pass 2 of split_fastq / subsample_reads (twoset.rs:157-186, ava.rs:128-155): only the records idx_a / idx_b
    // (ascending record numbers) are kept, straight into pinned buffers; idx_b may be empty
    static std::pair<ReadSet, ReadSet> from_fastx_subsets(const std::string &path, const std::vector<uint64_t> &idx_a,
                                                          const std::vector<uint64_t> &idx_b, uint64_t hint_bases_per_record = 0) {
        lrge_hostset *a = nullptr, *b = nullptr;
        const int rc = lrge_fastx_load_subsets(path.c_str(), idx_a.data(), idx_a.size(), idx_b.empty() ? nullptr : idx_b.data(), idx_b.size(),
                                               hint_bases_per_record, 1, &a, idx_b.empty() ? nullptr : &b);
        if (rc == LRGE_ERR_IO) throw LrgeError(LRGE_ERR_IO, "IO error: cannot read " + path);
        if (rc != LRGE_OK) throw LrgeError((lrge_status)rc, "FASTQ parse error: " + path);
        ReadSet A, B;
        A.host.reset(a, lrge_hostset_free);
        if (b) B.host.reset(b, lrge_hostset_free);
        return {std::move(A), std::move(B)};
    }
    // plain or gzip FASTA / FASTQ; the read id is the header up to the first whitespace (io.rs:192-205)
    static ReadSet from_fastx(const std::string &path) {
        uint64_t n = 0, nb = 0, nn = 0;
        int rc = lrge_fastx_count(path.c_str(), &n, &nb, &nn);
        if (rc == LRGE_ERR_IO) throw LrgeError(LRGE_ERR_IO, "IO error: cannot read " + path + " (is the file empty?)");
        if (rc == LRGE_ERR_UNSUPPORTED) throw LrgeError(LRGE_ERR_UNSUPPORTED, "unsupported compression: " + path);
        if (rc != LRGE_OK) throw LrgeError(LRGE_ERR_FASTQ_PARSE, "FASTQ parse error: " + path);
        ReadSet rs;
        rs.bases.resize(nb); rs.off.resize(n + 1); rs.names.resize(nn); rs.name_off.resize(n + 1);
        rc = lrge_fastx_load(path.c_str(), rs.bases.data(), rs.bases.size(), rs.off.data(), rs.off.size() - 1, rs.names.data(), rs.names.size(), rs.name_off.data());
        if (rc != LRGE_OK) throw LrgeError((lrge_status)rc, "FASTQ parse error: " + path);
        return rs;
    }
};

struct EstimateResult {                                  // estimate.rs:8-17
    std::optional<float> lower, estimate, upper;
    uint32_t no_mapping_count = 0;
};

// One GPU context per process (lrge_ctx).
class Context {
  public:
    explicit Context(int device = 0) {
        const int rc = lrge_gpu_create(device, &h_);
        if (rc != LRGE_OK) throw LrgeError((lrge_status)rc, lrge_gpu_last_error(nullptr));
    }
    ~Context() { lrge_gpu_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    lrge_ctx *get() const { return h_; }
    void check(int rc) const { if (rc != LRGE_OK) throw LrgeError((lrge_status)rc, lrge_gpu_last_error(h_)); }

  private:
    lrge_ctx *h_ = nullptr;
};

// overlaps.paf as liblrge leaves it in the temp directory with -C (twoset.rs:244-250,290-293; ava.rs:219-226): one row per
// kept chain, serialised like PafRecord (mapping.rs:81-177).  Rows come sorted by query, in mm_map's order within a query
// (liblrge's own file interleaves queries in thread-completion order).
inline void write_paf(const std::string &path, const std::vector<lrge_overlap> &ov, const ReadSet &Q, const ReadSet &T) {
    FILE *f = fopen(path.c_str(), "w");
    if (!f) throw LrgeError(LRGE_ERR_PAF_WRITE, "PAF write error: cannot open " + path);
    char buf[2048];
    const lrge_readset vq = Q.view(), vt = T.view();
    for (const lrge_overlap &o : ov) {
        const uint32_t q = o.query, t = o.target;
        const int n = lrge_paf_format(&o, vq.names + vq.name_off[q], (int)(vq.name_off[q + 1] - vq.name_off[q]), (int32_t)(vq.off[q + 1] - vq.off[q]),
                                      vt.names + vt.name_off[t], (int)(vt.name_off[t + 1] - vt.name_off[t]), (int32_t)(vt.off[t + 1] - vt.off[t]),
                                      o.dv, o.rl, buf, sizeof buf);
        if (n < 0 || fwrite(buf, 1, (size_t)n, f) != (size_t)n || fputc('\n', f) == EOF) { fclose(f); throw LrgeError(LRGE_ERR_PAF_WRITE, "PAF write error: " + path); }
    }
    fclose(f);
}

// trait Estimate (estimate.rs:21-78)
class Estimate {
  public:
    virtual ~Estimate() = default;
    virtual std::pair<std::vector<float>, uint32_t> generate_estimates() = 0;
    EstimateResult estimate(bool finite, std::optional<float> lower_quant, std::optional<float> upper_quant) {
        auto [estimates, no_mapping_count] = generate_estimates();
        float lo = 0, md = 0, up = 0;
        const int m = lrge_estimate(estimates.data(), estimates.size(), finite ? 1 : 0, lower_quant.value_or(-1.0f),
                                    upper_quant.value_or(-1.0f), &lo, &md, &up);
        if (m < 0) throw std::invalid_argument("Quantile must be between 0.0 and 1.0");   // estimate.rs:118-120 panics
        EstimateResult r;
        if (m & 1) r.lower = lo;
        if (m & 2) r.estimate = md;
        if (m & 4) r.upper = up;
        r.no_mapping_count = no_mapping_count;
        return r;
    }
};

class TwoSetStrategy : public Estimate {
  public:
    class Builder {                                      // twoset/builder.rs:22-39 defaults
      public:
        Builder &remove_internal(bool b) { o_.remove_internal = b; return *this; }
        Builder &max_overhang_ratio(float r) { o_.max_overhang_ratio = r; return *this; }
        Builder &use_min_ref(bool b) { o_.use_min_ref = b; return *this; }
        Builder &platform(Platform p) { o_.platform = p == Platform::PacBio ? LRGE_PLATFORM_PACBIO : LRGE_PLATFORM_NANOPORE; return *this; }
        Builder &device(int d) { device_ = d; return *this; }
        Builder &paf(std::string path) { paf_ = std::move(path); return *this; }     // also write overlaps.paf (the -C artefact)
        TwoSetStrategy build(ReadSet target, ReadSet query) const {
            TwoSetStrategy s(std::move(target), std::move(query), o_, device_);
            s.paf_path_ = paf_;
            return s;
        }
        Builder() { lrge_opts_default(&o_); }

      private:
        lrge_opts o_;
        int device_ = 0;
        std::string paf_;
    };
    TwoSetStrategy(ReadSet target, ReadSet query) : TwoSetStrategy(Builder().build(std::move(target), std::move(query))) {}
    size_t target_num_reads() const { return target_.size(); }
    size_t query_num_reads() const { return query_.size(); }
    const lrge_stats &stats() const { return stats_; }

    std::pair<std::vector<float>, uint32_t> generate_estimates() override {    // twoset.rs:587-606
        Context ctx(device_);
        std::vector<float> est(query_.size());
        uint32_t no_mapping = 0;
        const lrge_readset t = target_.view(), q = query_.view();
        if (!paf_path_.empty()) {
            std::vector<lrge_overlap> ov(std::max<size_t>(1024, 64 * query_.size()));
            uint64_t n = 0;
            for (;;) {
                ctx.check(lrge_gpu_twoset_overlaps(ctx.get(), &t, &q, &opts_, ov.data(), ov.size(), &n, nullptr, nullptr, nullptr));
                if (n <= ov.size()) break;
                ov.resize(n);
            }
            ov.resize(n);
            // with use_min_ref and more target than query bases the records name streamed target reads as queries
            const bool inverse = opts_.use_min_ref && target_.total_bases() > query_.total_bases();
            write_paf(paf_path_, ov, inverse ? target_ : query_, inverse ? query_ : target_);
        }
        ctx.check(lrge_twoset_generate_estimates(ctx.get(), &t, &q, &opts_, est.data(), &no_mapping, &stats_));
        return {std::move(est), no_mapping};
    }

  private:
    TwoSetStrategy(ReadSet t, ReadSet q, lrge_opts o, int device) : target_(std::move(t)), query_(std::move(q)), opts_(o), device_(device) {}
    ReadSet target_, query_;
    lrge_opts opts_;
    int device_;
    std::string paf_path_;
    lrge_stats stats_{};
};

class AvaStrategy : public Estimate {
  public:
    class Builder {                                      // ava/builder.rs:18-32 defaults
      public:
        Builder &remove_internal(bool b) { o_.remove_internal = b; return *this; }
        Builder &max_overhang_ratio(float r) { o_.max_overhang_ratio = r; return *this; }
        Builder &platform(Platform p) { o_.platform = p == Platform::PacBio ? LRGE_PLATFORM_PACBIO : LRGE_PLATFORM_NANOPORE; return *this; }
        Builder &device(int d) { device_ = d; return *this; }
        Builder &paf(std::string path) { paf_ = std::move(path); return *this; }     // also write overlaps.paf (the -C artefact)
        AvaStrategy build(ReadSet reads) const {
            AvaStrategy s(std::move(reads), o_, device_);
            s.paf_path_ = paf_;
            return s;
        }
        Builder() { lrge_opts_default(&o_); }

      private:
        lrge_opts o_;
        int device_ = 0;
        std::string paf_;
    };
    explicit AvaStrategy(ReadSet reads) : AvaStrategy(Builder().build(std::move(reads))) {}
    size_t num_reads() const { return reads_.size(); }
    const lrge_stats &stats() const { return stats_; }

    std::pair<std::vector<float>, uint32_t> generate_estimates() override {    // ava.rs:369-381
        Context ctx(device_);
        std::vector<float> est(reads_.size());
        uint32_t no_mapping = 0;
        const lrge_readset r = reads_.view();
        if (!paf_path_.empty()) {
            std::vector<lrge_overlap> ov(std::max<size_t>(1024, 64 * reads_.size()));
            uint64_t n = 0;
            for (;;) {
                ctx.check(lrge_gpu_ava_overlaps(ctx.get(), &r, &opts_, ov.data(), ov.size(), &n, nullptr, nullptr));
                if (n <= ov.size()) break;
                ov.resize(n);
            }
            ov.resize(n);
            write_paf(paf_path_, ov, reads_, reads_);
        }
        ctx.check(lrge_ava_generate_estimates(ctx.get(), &r, &opts_, est.data(), &no_mapping, &stats_));
        return {std::move(est), no_mapping};
    }

  private:
    AvaStrategy(ReadSet r, lrge_opts o, int device) : reads_(std::move(r)), opts_(o), device_(device) {}
    ReadSet reads_;
    lrge_opts opts_;
    int device_;
    std::string paf_path_;
    lrge_stats stats_{};
};

}  // namespace lrge
