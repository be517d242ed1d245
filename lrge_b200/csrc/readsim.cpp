// readsim.cpp -- deterministic synthetic long-read simulator (host only, no CUDA).
//
// Produces the read sets named by BASELINE.json / SURVEY.md section 8(d): an i.i.d. uniform ACGT
// genome and reads with Gamma(shape 2) lengths, uniform circular start, random strand and
// independent per-base substitution / insertion / deletion errors.  Every read draws from
// its own counter-seeded generator, so the output does not depend on threading or call order.
// The explicit target / query sets stand in for liblrge's split_fastq / subsample_reads
// (twoset.rs:122-201, ava.rs:108-161), whose `rand` crate shuffle cannot be reproduced here.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <vector>
#include <string>

namespace {

struct Rng {  // splitmix64-seeded xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    Rng(uint64_t seed, uint64_t stream) {
        uint64_t x = seed * 0xD1342543DE82EF95ULL + stream * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
        for (int i = 0; i < 4; ++i) s[i] = splitmix(x);
    }
    static uint64_t rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    uint64_t below(uint64_t n) { return (uint64_t)(uniform() * (double)n); }
};

const char kBase[4] = {'A', 'C', 'G', 'T'};
inline char comp(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

}  // namespace

extern "C" {

// genome[G] <- i.i.d. uniform ACGT
void readsim_genome(uint64_t G, uint64_t seed, uint8_t *genome) {
    const uint64_t kBlock = 1 << 16;
    for (uint64_t b = 0; b * kBlock < G; ++b) {
        Rng rng(seed, b);
        uint64_t e = (b + 1) * kBlock < G ? (b + 1) * kBlock : G;
        for (uint64_t i = b * kBlock; i < e; i += 32) {
            uint64_t r = rng.next();
            for (uint64_t j = i; j < e && j < i + 32; ++j, r >>= 2) genome[j] = kBase[r & 3];
        }
    }
}

// Simulate n reads. Outputs are malloc'd (free with readsim_free):
//   *bases  concatenated ASCII, *off n+1 offsets, *names concatenated "<prefix>%08u", *name_off n+1.
// Optional per-read truth (may be NULL): start[n], strand[n], span[n] (genome bases covered).
int readsim_reads(const uint8_t *genome, uint64_t G, uint32_t n, double mean_len,
                  double sub_rate, double ins_rate, double del_rate, double n_rate,
                  uint64_t seed, const char *prefix, uint32_t first_id,
                  uint8_t **bases, uint64_t **off, char **names, uint32_t **name_off,
                  uint64_t *start_out, uint8_t *strand_out, uint32_t *span_out) {
    std::vector<uint8_t> out;
    out.reserve((size_t)(n * mean_len * 1.05) + 1024);
    uint64_t *o = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)n + 1));
    uint32_t *no = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + 1));
    std::string nm;
    std::vector<uint8_t> tmp;
    o[0] = 0; no[0] = 0;
    for (uint32_t r = 0; r < n; ++r) {
        Rng rng(seed, r);
        // Gamma(shape 2, scale mean/2) = sum of two exponentials
        double u1 = 1.0 - rng.uniform(), u2 = 1.0 - rng.uniform();
        double len_d = -(mean_len / 2.0) * (std::log(u1) + std::log(u2));
        if (len_d < 500.0) len_d = 500.0;
        if (len_d > 8.0 * mean_len) len_d = 8.0 * mean_len;
        uint64_t span = (uint64_t)len_d;
        if (span > G) span = G;
        uint64_t start = rng.below(G);
        int rev = (int)(rng.next() >> 63);
        tmp.clear();
        for (uint64_t i = 0; i < span; ++i) {
            char c = (char)genome[(start + i) % G];
            double u = rng.uniform();
            if (u < del_rate) continue;
            if (u < del_rate + ins_rate) { tmp.push_back(kBase[rng.next() & 3]); }
            else if (u < del_rate + ins_rate + sub_rate) {
                int k = (int)(rng.next() % 3);
                int ci = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                c = kBase[(ci + 1 + k) & 3];
            }
            if (n_rate > 0.0 && rng.uniform() < n_rate) c = 'N';
            tmp.push_back((uint8_t)c);
        }
        if (tmp.empty()) tmp.push_back('A');
        if (rev) {
            size_t m = tmp.size();
            for (size_t i = 0; i < m / 2; ++i) { uint8_t a = tmp[i]; tmp[i] = tmp[m - 1 - i]; tmp[m - 1 - i] = a; }
            for (size_t i = 0; i < m; ++i) tmp[i] = (uint8_t)comp((char)tmp[i]);
        }
        out.insert(out.end(), tmp.begin(), tmp.end());
        o[r + 1] = out.size();
        char buf[64];
        snprintf(buf, sizeof buf, "%s%08u", prefix, first_id + r);
        nm += buf;
        no[r + 1] = (uint32_t)nm.size();
        if (start_out) start_out[r] = start;
        if (strand_out) strand_out[r] = (uint8_t)rev;
        if (span_out) span_out[r] = (uint32_t)span;
    }
    *bases = (uint8_t *)malloc(out.size() ? out.size() : 1);
    memcpy(*bases, out.data(), out.size());
    *off = o;
    *names = (char *)malloc(nm.size() ? nm.size() : 1);
    memcpy(*names, nm.data(), nm.size());
    *name_off = no;
    return 0;
}

void readsim_free(void *p) { free(p); }

}  // extern "C"
