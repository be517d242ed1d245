// lrge-b200 -- command-line driver with the lrge CLI's flags (lrge/src/cli.rs:9-87, main.rs:32-121) over liblrge_b200.so.
//
// Differences, on purpose: plain or gzip FASTA/FASTQ only (no bzip2/zstd/xz/BAM: the rest of io.rs is out of scope); the
// random subsample uses this file's own generator, so `-s SEED` is repeatable here but does not pick the reads the Rust
// `rand` crate would (lib.rs:189-204); `-t` is accepted and ignored; `-C -D <dir>` writes <dir>/overlaps.paf (the one
// intermediate of liblrge that is worth keeping; there are no temp FASTA files on this path); and `--platform` IS
// forwarded (the reference CLI parses it but never passes it on, main.rs:56-85).
#include "../../include/lrge_b200.hpp"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <numeric>
#include <string>
#include <vector>

namespace {

struct Args {
    std::string input, output = "-", platform = "ont", temp_dir;
    bool keep_temp = false;
    size_t target = lrge::DEFAULT_TARGET_NUM_READS, query = lrge::DEFAULT_QUERY_NUM_READS, num = 0;
    bool have_num = false, have_target = false, have_query = false;
    bool filter_contained = false, with_infinity = false, precise = false, use_min_ref = false, seeded = false;
    uint64_t seed = 0;
    float q1 = lrge::LOWER_QUANTILE, q3 = lrge::UPPER_QUANTILE, max_overhang_ratio = 0.2f;
    int device = 0, quiet = 0;
};

[[noreturn]] void usage(int rc) {
    fprintf(rc ? stderr : stdout,
            "Usage: lrge-b200 [OPTIONS] <INPUT>\n"
            "  -o, --output <OUTPUT>        output file for the estimate [default: -]\n"
            "  -T, --target <INT>           target number of reads (two-set strategy) [default: 10000]\n"
            "  -Q, --query <INT>            query number of reads (two-set strategy) [default: 5000]\n"
            "  -n, --num <INT>              number of reads (all-vs-all strategy)\n"
            "  -P, --platform <ont|pb>      sequencing platform [default: ont]\n"
            "  -F, --filter-contained       exclude overlaps for internal matches\n"
            "  -s, --seed <INT>             random seed\n"
            "  -8, --inf                    median over all estimates, including infinite ones\n"
            "  -f, --float-my-boat          print the estimate as a floating point number\n"
            "      --q1 <FLOAT> --q3 <FLOAT>  quantiles of the reported range [default: 0.15 0.65]\n"
            "      --max-overhang-ratio <FLOAT> [default: 0.2]\n"
            "      --use-min-ref            index the smaller of the two sets\n"
            "      --device <INT>           CUDA device [default: 0]\n"
            "  -C, --keep-temp              keep overlaps.paf (written to the -D directory, default: current directory)\n"
            "  -D, --temp <DIR>             directory for overlaps.paf\n"
            "  -t, --threads <INT>          accepted for compatibility, ignored\n");
    exit(rc);
}

Args parse(int argc, char **argv) {
    Args a;
    auto need = [&](int &i) -> const char * { if (i + 1 >= argc) usage(2); return argv[++i]; };
    for (int i = 1; i < argc; ++i) {
        const std::string s = argv[i];
        if (s == "-h" || s == "--help") usage(0);
        else if (s == "-o" || s == "--output") a.output = need(i);
        else if (s == "-T" || s == "--target") { a.target = strtoull(need(i), nullptr, 10); a.have_target = true; }
        else if (s == "-Q" || s == "--query") { a.query = strtoull(need(i), nullptr, 10); a.have_query = true; }
        else if (s == "-n" || s == "--num") { a.num = strtoull(need(i), nullptr, 10); a.have_num = true; }
        else if (s == "-P" || s == "--platform") a.platform = need(i);
        else if (s == "-F" || s == "--filter-contained") a.filter_contained = true;
        else if (s == "-s" || s == "--seed") { a.seed = strtoull(need(i), nullptr, 10); a.seeded = true; }
        else if (s == "-8" || s == "--inf") a.with_infinity = true;
        else if (s == "-f" || s == "--float-my-boat") a.precise = true;
        else if (s == "--q1") a.q1 = strtof(need(i), nullptr);
        else if (s == "--q3") a.q3 = strtof(need(i), nullptr);
        else if (s == "--max-overhang-ratio") a.max_overhang_ratio = strtof(need(i), nullptr);
        else if (s == "--use-min-ref") a.use_min_ref = true;
        else if (s == "--device") a.device = atoi(need(i));
        else if (s == "-t" || s == "--threads") need(i);
        else if (s == "-D" || s == "--temp") a.temp_dir = need(i);
        else if (s == "-C" || s == "--keep-temp") a.keep_temp = true;
        else if (s == "--quiet") ++a.quiet;
        else if (s == "--verbose") {}
        else if (s.size() >= 2 && s[0] == '-' && s[1] == 'q' && s.find_first_not_of('q', 1) == std::string::npos) a.quiet += (int)s.size() - 1;
        else if (s.size() >= 2 && s[0] == '-' && s[1] == 'v' && s.find_first_not_of('v', 1) == std::string::npos) {}
        else if (!s.empty() && s[0] == '-' && s != "-") { fprintf(stderr, "error: unexpected argument '%s'\n", s.c_str()); usage(2); }
        else a.input = s;
    }
    if (a.input.empty()) usage(2);
    if (a.have_num && (a.have_target || a.have_query)) {      // cli.rs: num conflicts_with_all = ["target", "query"]
        fprintf(stderr, "error: the argument '--num <INT>' cannot be used with '--target <INT>' / '--query <INT>'\n");
        exit(2);
    }
    // cli.rs: --q1 in [0, 0.5], --q3 in [0.5, 1], ratio in [0, 1]
    if (!(a.q1 >= 0.0f && a.q1 <= 0.5f) || !(a.q3 >= 0.5f && a.q3 <= 1.0f) || !(a.max_overhang_ratio >= 0.0f && a.max_overhang_ratio <= 1.0f)) {
        fprintf(stderr, "error: quantile or ratio out of range\n");
        exit(2);
    }
    return a;
}

// k unique indices out of [0, n), in random order (stands in for lib.rs:189-204)
std::vector<uint32_t> unique_random_set(size_t k, uint32_t n, uint64_t seed) {
    std::vector<uint32_t> idx(n);
    std::iota(idx.begin(), idx.end(), 0u);
    uint64_t x = seed * 0x9E3779B97F4A7C15ULL + 0xD1B54A32D192ED03ULL;
    auto next = [&]() { uint64_t z = (x += 0x9E3779B97F4A7C15ULL); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); };
    for (size_t i = 0; i < k; ++i) { const size_t j = i + (size_t)(next() % (n - i)); std::swap(idx[i], idx[j]); }
    idx.resize(k);
    return idx;
}

std::string format_estimate(float est) {                 // utils.rs:18-47
    if (std::isinf(est)) return "\xE2\x88\x9E bp";
    const char *units[] = {"bp", "kbp", "Mbp", "Gbp", "Tbp", "Pbp"};
    float value = est; const char *suffix = "bp";
    for (int p = 0; p < 6; ++p) {
        const float thr = std::pow(10.0f, (float)(p * 3));
        if (est >= thr) { value = est / thr; suffix = units[p]; } else break;
    }
    char buf[64]; snprintf(buf, sizeof buf, "%.2f %s", value, suffix);
    return buf;
}

}  // namespace

int main(int argc, char **argv) {
    const Args a = parse(argc, argv);
    try {
        const lrge::Platform platform = lrge::platform_from_str(a.platform);
        // pass 1 of split_fastq / subsample_reads: count the records (io::count_records, twoset.rs:124, ava.rs:110)
        uint64_t n_rec = 0, n_bases = 0, n_name_bytes = 0;
        {
            const int rc = lrge_fastx_count(a.input.c_str(), &n_rec, &n_bases, &n_name_bytes);
            if (rc == LRGE_ERR_IO) throw lrge::LrgeError(LRGE_ERR_IO, "IO error: cannot read " + a.input + " (is the file empty?)");
            if (rc == LRGE_ERR_UNSUPPORTED) throw lrge::LrgeError(LRGE_ERR_UNSUPPORTED, "unsupported compression: " + a.input);
            if (rc != LRGE_OK) throw lrge::LrgeError(LRGE_ERR_FASTQ_PARSE, "FASTQ parse error: " + a.input);
            if (n_rec > 0xffffffffull) throw lrge::LrgeError(LRGE_ERR_TOO_MANY_READS, "Too many reads");   // twoset.rs:126-130
        }
        const uint32_t n_fq = (uint32_t)n_rec;
        const uint64_t hint = n_rec ? n_bases / n_rec + 1 : 0;
        const uint64_t seed = a.seeded ? a.seed : (uint64_t)time(nullptr);
        // pass 2: only the selected records are kept, straight into pinned buffers
        auto take = [&](const std::vector<uint32_t> &sa, const std::vector<uint32_t> &sb) {
            return lrge::ReadSet::from_fastx_subsets(a.input, std::vector<uint64_t>(sa.begin(), sa.end()),
                                                     std::vector<uint64_t>(sb.begin(), sb.end()), hint);
        };
        lrge::EstimateResult res;
        const std::string paf_path = a.keep_temp ? (a.temp_dir.empty() ? std::string("overlaps.paf") : a.temp_dir + "/overlaps.paf") : std::string();
        if (a.have_num) {                                         // AvaStrategy (ava.rs:108-161)
            size_t n = a.num;
            if (n_fq < n) { if (a.quiet < 1) fprintf(stderr, "[WARN] only %u reads in input, using all\n", n_fq); n = n_fq; }
            auto idx = unique_random_set(n, n_fq, seed);
            std::sort(idx.begin(), idx.end());
            auto strat = lrge::AvaStrategy::Builder().platform(platform).remove_internal(a.filter_contained)
                             .max_overhang_ratio(a.max_overhang_ratio).device(a.device).paf(paf_path).build(take(idx, {}).first);
            res = strat.estimate(!a.with_infinity, a.q1, a.q3);
        } else {                                                  // TwoSetStrategy (twoset.rs:122-201)
            size_t T = a.target, Q = a.query;
            if (n_fq <= Q) throw lrge::LrgeError(LRGE_ERR_TOO_FEW_READS, "Too few reads requested: input has <= query number of reads");
            if (n_fq < T + Q) { T = n_fq - Q; if (a.quiet < 1) fprintf(stderr, "[WARN] using %zu target reads\n", T); }
            auto idx = unique_random_set(T + Q, n_fq, seed);
            // split_into_hashsets pops from the END: the last T indices are the targets (twoset.rs:632-652)
            std::vector<uint32_t> t(idx.end() - T, idx.end()), q(idx.begin(), idx.end() - T);
            std::sort(t.begin(), t.end()); std::sort(q.begin(), q.end());
            auto sets = take(t, q);
            auto strat = lrge::TwoSetStrategy::Builder().platform(platform).remove_internal(a.filter_contained)
                             .max_overhang_ratio(a.max_overhang_ratio).use_min_ref(a.use_min_ref).device(a.device).paf(paf_path)
                             .build(std::move(sets.first), std::move(sets.second));
            res = strat.estimate(!a.with_infinity, a.q1, a.q3);
        }
        if (!res.estimate) {
            fprintf(stderr, "Error: %s\n", a.with_infinity ? "No estimates were generated" : "No finite estimates were generated");
            return 1;
        }
        if (a.quiet < 1) {
            std::string msg = "Estimated genome size: " + format_estimate(*res.estimate);
            if (res.lower && res.upper) msg += " (IQR: " + format_estimate(*res.lower) + " - " + format_estimate(*res.upper) + ")";
            fprintf(stderr, "[INFO] %s\n", msg.c_str());
        }
        FILE *out = a.output == "-" ? stdout : fopen(a.output.c_str(), "w");
        if (!out) throw lrge::LrgeError(LRGE_ERR_IO, "IO error: cannot open " + a.output);
        if (a.precise) {                                          // main.rs:103-107 `{est}`: shortest decimal that round-trips the f32
            char buf[64];
            auto r = std::to_chars(buf, buf + sizeof buf - 1, *res.estimate, std::chars_format::fixed);
            *r.ptr = 0;
            std::string t = buf;
            if (std::isinf(*res.estimate)) t = "inf";
            fprintf(out, "%s\n", t.c_str());
        } else fprintf(out, "%.0f\n", (double)*res.estimate);
        if (out != stdout) fclose(out);
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "Error: %s\n", e.what());
        return 1;
    }
}
