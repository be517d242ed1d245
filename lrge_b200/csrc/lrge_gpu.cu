// lrge_gpu.cu -- context, device pipeline and C-ABI of liblrge_b200.so (include/lrge_b200.h).
//
// Pipeline (main stream + a side stream for the streamed set's sketch + a copy stream; every kernel is hand-written
// except the index radix sort, the prefix sums / selects and the optional merge, which are CUB device primitives):
//   K1  ava-ont: sketch2_kernel (TMA-staged ASCII, lane-serial state machine) -> scan -> compact_kernel (+ 32-bit sort keys)
//       ava-pb / fallback: hpc_compress_kernel, pack_kernel, sketch_kernel<stage | count, emit>
//   K2  radix sort of the indexed set's minimizers by hash (12-byte records for ava-ont)       (index in HBM)
//   K3  index_post_kernel      bucket table (+ compact lookup keys) + occurrence histogram -> mid_occ (mm_mapopt_update)
//       build_index_sharded_*  the index sharded over ranks: packed key|count entries pushed over peer memory (CUDA IPC
//                              windows; ncclSend/Recv as the agreed fallback), owner_merge_kernel, mid_occ on the device
//   K4  mz_flt_kernel, seed_lookup_kernel (seed_count_kernel when name rules apply) -> scan -> anchor_fill_kernel
//   K5  sort_query_kernel      per query in shared memory: stage, radix sort on (target, strand), boundary bit maps,
//                              per-group order, compact anchors + group descriptors
//       (deferred queries / LRGE_FORCE_GLOBAL: device-wide radix sort, group_mark + select + group_classify_kernel)
//   K6  chain_group_kernel     chaining DP + backtrack + overlap decision, one warp per (query, target, strand) group
//                              (test build only: chain_pair_kernel, the literal one-thread-per-pair cross-check)
// There is no CPU fallback anywhere in this file: without a device every entry point fails.
#include "../../include/lrge_b200.h"
#include "common.cuh"
#include "kernels_sketch.cuh"
#include "kernels_index.cuh"
#include "kernels_seed.cuh"
#include "kernels_chain.cuh"
#include "kernels_map.cuh"
#ifdef LRGE_TEST_KERNELS
#include "../../tests/native/kernels_chain_v1.cuh"
#endif

#include <cub/cub.cuh>
#include <cub/device/device_merge.cuh>
#include <functional>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

using namespace lrge;

// ------------------------------------------------------------------------------------------------
// host-side option block: minimap2 defaults o preset o liblrge
// ------------------------------------------------------------------------------------------------
namespace {

struct MapOpt {
    int k, w, is_hpc;
    int bw, max_gap, max_chain_skip, max_chain_iter, min_cnt, min_chain_score;
    float chain_gap_scale;
    float mid_occ_frac, q_occ_frac;
    int min_mid_occ, max_mid_occ;
};

MapOpt preset_for(int platform) {
    MapOpt o{};
    o.w = 5;
    o.max_gap = 5000; o.max_chain_skip = 25; o.max_chain_iter = 5000; o.min_cnt = 3; o.min_chain_score = 100;
    o.chain_gap_scale = 0.8f;
    o.mid_occ_frac = 2e-4f; o.q_occ_frac = 0.01f; o.min_mid_occ = 10; o.max_mid_occ = 1000000;
    if (platform == LRGE_PLATFORM_PACBIO) { o.k = 19; o.is_hpc = 1; o.bw = 500; }   // ava-pb
    else { o.k = 15; o.is_hpc = 0; o.bw = 2000; }                                    // ava-ont
    return o;
}

// mg_log2 (mmpriv.h) and the comput_sc gap penalty (lchain.c) in f32, one rounding per operation.
inline float mg_log2_host(float x) {
    union { float f; uint32_t i; } z;
    z.f = x;
    volatile float log_2 = (float)(int)(((z.i >> 23) & 255) - 128);
    z.i &= ~(255U << 23);
    z.i += 127U << 23;
    volatile float a = -0.34484843f * z.f;
    volatile float b = a + 2.02466578f;
    volatile float c = b * z.f;
    volatile float d = c - 0.67487759f;
    volatile float e = log_2 + d;
    return e;
}
void gap_penalty_table(int k, float chain_gap_scale, int n, std::vector<int32_t> &pen) {
    const float chn_pen_gap = (float)(chain_gap_scale * 0.01 * k);   // float * double * int, then narrowed
    pen.resize(n);
    for (int dd = 0; dd < n; ++dd) {
        volatile float lin = chn_pen_gap * (float)dd;
        volatile float lg = dd >= 1 ? mg_log2_host((float)(dd + 1)) : 0.0f;
        volatile float half = 0.5f * lg;
        volatile float sum = lin + half;
        pen[dd] = (int32_t)sum;
    }
}

inline int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) ++b; return b; }   // bits to hold v

const uint8_t kNt4[256] = {
    0, 1, 2, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 0, 4, 1, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4,
    4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};

thread_local std::string g_create_error;

struct NcclError : std::runtime_error { using std::runtime_error::runtime_error; };

// Event-based profiler: spans are recorded on the stream without host syncs and resolved after the run.
enum Stage { ST_PACK, ST_SKETCH, ST_INDEX, ST_SEED, ST_SORT, ST_CHAIN, ST_D2H, ST_TOTAL, ST_COMM, ST_COUNT };
struct Prof {
    cudaStream_t s = nullptr;
    bool kernels = false;                              // per-kernel spans (lrge_gpu_set_profiling)
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct Span { int id; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<std::string> names;                    // kernel spans: id = ST_COUNT + index into names
    std::vector<double> ms, bytes;
    std::vector<uint32_t> count, passes;
    int open_kernel = -1;
    cudaEvent_t open_ev = nullptr;
    cudaEvent_t stage_ev[ST_COUNT] = {};
    ~Prof() { for (auto e : pool) cudaEventDestroy(e); }
    cudaEvent_t get() {
        if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        return pool[used++];
    }
    void reset() {
        used = 0; spans.clear(); open_kernel = -1;
        ms.assign(ST_COUNT + names.size(), 0.0); bytes.assign(ST_COUNT + names.size(), 0.0); count.assign(ST_COUNT + names.size(), 0);
        passes.resize(ST_COUNT + names.size(), 1);
    }
    int kid(const char *name) {
        for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return ST_COUNT + (int)i;
        names.emplace_back(name); ms.push_back(0); bytes.push_back(0); count.push_back(0); passes.push_back(1);
        return ST_COUNT + (int)names.size() - 1;
    }
    void sbegin(Stage st) { stage_ev[st] = get(); cudaEventRecord(stage_ev[st], s); }
    void send(Stage st) { cudaEvent_t b = get(); cudaEventRecord(b, s); spans.push_back({(int)st, stage_ev[st], b}); }
    void kbegin(const char *name, double nbytes) {
        if (!kernels) return;
        open_kernel = kid(name); bytes[open_kernel] += nbytes;
        open_ev = get(); cudaEventRecord(open_ev, s);
    }
    void kend() {
        if (!kernels || open_kernel < 0) return;
        cudaEvent_t b = get(); cudaEventRecord(b, s); spans.push_back({open_kernel, open_ev, b}); open_kernel = -1;
    }
    void set_bytes(const char *name, double nbytes, bool add) {   // algorithmic bytes known only after the launch
        if (!kernels) return;
        for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) bytes[ST_COUNT + i] = (add ? bytes[ST_COUNT + i] : 0.0) + nbytes;
    }
    void set_passes(const char *name, uint32_t n) {    // a library radix sort: passes over the data inside one "launch"
        if (!kernels) return;
        for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) passes[ST_COUNT + i] = n;
    }
    // LRGE_TIMELINE=1: marks (any stream) and spans as offsets from the start of the call, printed by resolve()
    struct Mark { const char *what; cudaEvent_t e; };
    std::vector<Mark> marks;
    bool timeline = false;
    void mark(const char *what, cudaStream_t st) {     // own events: marks are recorded before reset() opens the call
        if (!timeline) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); marks.push_back({what, e});
    }
    void resolve() {                                   // the streams must be idle
        if (timeline && stage_ev[ST_TOTAL]) {
            static const char *stn[] = {"pack", "sketch", "index", "seed", "sort", "chain", "d2h", "total", "comm"};
            for (auto &m : marks) { float t = 0; if (cudaEventElapsedTime(&t, stage_ev[ST_TOTAL], m.e) == cudaSuccess) fprintf(stderr, "[timeline] %8.3f ms  %s\n", t, m.what); else cudaGetLastError(); }
            for (auto &sp : spans) {
                float a = 0, b = 0;
                if (cudaEventElapsedTime(&a, stage_ev[ST_TOTAL], sp.a) != cudaSuccess || cudaEventElapsedTime(&b, stage_ev[ST_TOTAL], sp.b) != cudaSuccess) { cudaGetLastError(); continue; }
                fprintf(stderr, "[timeline] %8.3f .. %8.3f ms  %s\n", a, b, sp.id < ST_COUNT ? stn[sp.id] : names[sp.id - ST_COUNT].c_str());
            }
        }
        for (auto &m : marks) cudaEventDestroy(m.e);
        marks.clear();
        for (auto &sp : spans) { float t = 0; cudaEventElapsedTime(&t, sp.a, sp.b); ms[sp.id] += t; ++count[sp.id]; }
        spans.clear();
    }
};
#define KLAUNCH(c, name, nbytes, ...) do { (c)->prof.kbegin(name, (double)(nbytes)); __VA_ARGS__; (c)->prof.kend(); ++(c)->launches; } while (0)

}  // namespace

// ------------------------------------------------------------------------------------------------
// staged read set and context
// ------------------------------------------------------------------------------------------------
struct lrge_staged {
    uint32_t n = 0;
    uint64_t total_bases = 0;
    uint32_t max_len = 0;
    DevBuf asc;                            // device: ASCII bases of reads [blk_first, blk_first + blk_count) (+64 slack)
    uint32_t blk_first = 0, blk_count = 0; // which reads' bases are resident (a rank may stage only its target shard)
    uint64_t blk_base = 0;                 // h_off[blk_first]: host offset of device byte 0
    std::vector<uint64_t> h_off;           // host copies
    std::vector<uint32_t> h_len;
    std::string names;
    std::vector<uint32_t> name_off;
    std::vector<std::pair<uint64_t, uint32_t>> name_hash;   // (FNV-1a of the name, read), sorted: cheap equality joins
    bool names_unique = true;
    bool has_empty = false;                                // a zero-length read: the index build skips it, mapping it is an error
    const lrge_readset *names_src = nullptr;               // host-buffer calls: names are analysed later, while the GPU sketches
    std::vector<uint32_t> lex_rank;                        // lazily: lexicographic rank of each name (AVA's strcmp order)
    std::vector<uint32_t> len_order;                       // reads, longest first (work order of the fused map kernel)
    cudaEvent_t ready = nullptr;                           // async staging: recorded on the copy stream after the H2D copies
    bool pending = false;
    // async staging copies the block in chunks of whole reads, one event each, so that the sketch of chunk i overlaps the
    // copy of chunk i + 1: chunk j = reads [chunk_hi[j-1], chunk_hi[j]) (absolute read ids)
    std::vector<uint32_t> chunk_hi;
    std::vector<cudaEvent_t> chunk_ev;
    // async staging also uploads the tile tables of the lane-serial sketch for the staged block, AHEAD of the bases: queued
    // behind them (one H2D engine, FIFO) they would hold the first sketch launch back until every base has arrived
    DevBuf t_tile_off, t_tile_read, t_sub_off, t_len, t_rid_map;
    std::vector<uint32_t> tab_tile_off;                    // host copy of tile_off (launch offsets per chunk)
    uint32_t tab_first = 0, tab_count = 0, tab_tiles = 0;
    uint64_t tab_bases = 0;
    bool tab_valid = false;
    ~lrge_staged() { if (ready) cudaEventDestroy(ready); for (auto e : chunk_ev) cudaEventDestroy(e); }
};

struct Sketch {                            // minimizers of a read set, (rid, position) order
    DevBuf x, y, moff;                     // moff: n+1 u32 per-read offsets
    uint64_t M = 0;
};


struct lrge_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // H2D staging of host read sets, overlapped with compute on `stream`
    cudaStream_t side_stream = nullptr;    // the sketch of the streamed set runs here, next to the index build on `stream`
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool on_side = false;                  // `stream` and `side_stream` are swapped (SideStream)
    uint64_t min_chunk = 8ull << 20;       // LRGE_MIN_CHUNK=bytes: smallest async staging chunk (tests reach the chunked paths with small sets)
    bool fork_recorded = false;            // ev_fork already marks the point the side stream has to wait for
    bool no_halves = true;                 // LRGE_INDEX_HALVES=1 turns the per-half index build of the host-buffer path ON: measured
                                           // slower on B200 (7.84 vs 7.39 ms for C2: the copy takes 1.3 ms at ~60 GB/s, the
                                           // indexed set's GPU work 3 ms, and two half sorts + a merge cost more than one sort)
    DevBuf hk_a, hk_b, hv_b;               // per-half index build: composite keys (two buffers), values of the sorted halves
    bool no_side = false;                  // LRGE_NO_SIDE_STREAM=1: everything on one stream (A/B measurements)
    DevBuf cub_tmp2;                       // CUB temporary storage of the side stream
    std::string err;
    uint32_t launches = 0;
    Prof prof;
    // layout / sketch scratch
    DevBuf poff, len0, sub_off, rid_map, tile_off, w2, nm, wcnt, woff, casc, cend, clen;
    // sharded index exchange
    DevBuf ukey, ucnt, rkey, rcnt, skey, scnt, gcnt, xmeta, blist, blall, blmap;
    DevBuf cub_tmp;
    // index
    Sketch sk_idx, sk_str;
    DevBuf tx, ty, boff, hist, tk32;
    // seed / anchors
    DevBuf flt, hstart, hn, acnt, aoff, qaoff;
    DevBuf akey0, akey1, aval0, aval1, flags, pstart, npairs, f, p, t, pair_res;
    DevBuf anc, fp, glist, gmeta, pflag, regs;
    bool chain_v1 = false;                 // LRGE_CHAIN_V1=1 (test build only): literal thread-per-pair kernel
    bool force_global = false;             // LRGE_FORCE_GLOBAL=1: every query through the global-memory path
    bool local_block = false;              // lrge_gpu_twoset_counts_shard: the staged indexed set IS this rank's block
    int index_bits_delta = 0;              // LRGE_INDEX_BITS_DELTA=d: 2^d times more (d > 0) or fewer buckets in the index table
    uint32_t map_cap = 0;                  // LRGE_MAP_CAP=n: anchors per shared-memory round (tests force multi-round)
    int map_ctas_per_sm = 1;               // LRGE_MAP_CTAS=1|2
    int smem_optin = 0;                    // cudaDevAttrMaxSharedMemoryPerBlockOptin
    int n_sm = 148;
    DevBuf qorder, deferred, mapctr, ddbuf;
    uint64_t max_anchors = 0;              // LRGE_MAX_ANCHORS=n: cap anchors per query batch
    bool sketch_two_pass = false;          // LRGE_SKETCH_TWO_PASS=1: count + emit instead of stage + compact
    bool sketch_legacy = false;            // LRGE_SKETCH_LEGACY=1: pack + warp-shuffle sketch + compact instead of the lane-serial kernel
    DevBuf tile_read;                      // lane-serial sketch: tile -> entry
    // small device -> host results land in PINNED memory: a cudaMemcpyAsync into pageable memory blocks the host until
    // the stream gets there, which would turn every "queue the copy, sync later" below into a sync on the spot
    uint8_t *pin2 = nullptr;               // sharded index: bounds + send-count matrix, then mid_occ / flags / totals (32 x 33 + 8 u64)
    uint8_t *pin = nullptr;                // [0] sketch total [4] sketch overflow [64..] occurrence histogram, (OCC_BINS + 2) x u64
    uint32_t &sk_total() { return *reinterpret_cast<uint32_t *>(pin); }
    uint32_t &sk_ovf() { return *reinterpret_cast<uint32_t *>(pin + 4); }
    unsigned long long *hist_h() { return reinterpret_cast<unsigned long long *>(pin + 64); }
    size_t free_at_create = 0, held_at_create = 0;
    DevBuf stg_x, stg_y, xflag;
    lrge_staged tmp_a, tmp_b;              // device staging of the host-buffer entry points (reused across calls)
    DevBuf rank_idx, rank_str, len_idx, len_str, pen, counts, had, counters;
    std::vector<int32_t> pen_host; int pen_k = 0, pen_bw = 0;
    // one process per GPU: NCCL communicator for the per-read count reduction
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // key exchange over peer memory (kernels_index.cuh: peer_push_kernel): IPC-exported flags and receive window
    struct PeerLink {
        bool ok = false;                                   // every rank exported / mapped its flags (agreed collectively)
        unsigned long long *flags = nullptr;
        unsigned long long *flags_peer[16] = {};
        DevBuf win;                                        // receive window
        uint64_t *win_peer[16] = {};
        size_t win_bytes = 0;                              // agreed size: the same on every rank
        unsigned long long epoch = 0;
        DevBuf hbuf, done;
    } link;
};

namespace {

// device memory still available to this context: what was free at creation minus what the process holds since
size_t free_estimate(const lrge_ctx *c) {
    const size_t used = dev_bytes_in_use().load();
    const size_t base = c->free_at_create + c->held_at_create;
    return base > used + (1ull << 30) ? base - used - (1ull << 30) : 0;
}

void *cub_temp(lrge_ctx *c, size_t bytes) { DevBuf &b = c->on_side ? c->cub_tmp2 : c->cub_tmp; b.reserve(bytes ? bytes : 16); return b.p; }

// While alive, everything that launches on c->stream (kernels, copies, profiler events, CUB calls) goes to the side stream.
struct SideStream {
    lrge_ctx *c;
    explicit SideStream(lrge_ctx *ctx) : c(ctx) { std::swap(c->stream, c->side_stream); c->prof.s = c->stream; c->on_side = true; }
    ~SideStream() { std::swap(c->stream, c->side_stream); c->prof.s = c->stream; c->on_side = false; }
};

template <typename T> void h2d(lrge_ctx *c, DevBuf &d, const T *src, size_t n) {
    d.reserve(std::max<size_t>(n, 1) * sizeof(T));
    if (n) CUDA_CHECK(cudaMemcpyAsync(d.p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
}

// reads [lo, hi) owned by `rank` of `world`: contiguous blocks, so a rank's share of a host read set is one copy
inline void shard_range(uint32_t n, int rank, int world, uint32_t *lo, uint32_t *hi) {
    *lo = (uint32_t)((uint64_t)n * (uint64_t)rank / (uint64_t)world);
    *hi = (uint32_t)((uint64_t)n * (uint64_t)(rank + 1) / (uint64_t)world);
}

// Copies a host read set (or only the bases of reads [blk_first, blk_first + blk_count)) to HBM on the copy stream;
// names and lengths of ALL reads are kept on the host side either way.  async: return without waiting -- the first
// kernel that touches the set waits for `ready` on the compute stream, so the copy of the second set overlaps the
// sketch of the first.
// Host-side analysis of a staged set that no kernel launch depends on: work order by length, name copies, name hashes
// (duplicate detection, shared names).  The host-buffer entry points defer it until the first sketch is in flight.
void finish_staging(lrge_staged *s) {
    const lrge_readset *rs = s->names_src;
    if (!rs) return;
    s->names_src = nullptr;
    s->len_order.resize(rs->n);
    std::iota(s->len_order.begin(), s->len_order.end(), 0u);
    std::stable_sort(s->len_order.begin(), s->len_order.end(), [&](uint32_t a, uint32_t b) { return s->h_len[a] > s->h_len[b]; });
    s->name_off.assign(rs->name_off, rs->name_off + rs->n + 1);
    s->names.assign(rs->names + rs->name_off[0], rs->names + rs->name_off[rs->n]);
    if (rs->n && rs->name_off[0] != 0) { const uint32_t b = rs->name_off[0]; for (auto &o : s->name_off) o -= b; }
    // name hashes: equal names <=> equal hashes (verified); the usual case "all names distinct" needs no string sort
    s->name_hash.resize(rs->n);
    for (uint32_t i = 0; i < rs->n; ++i) {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t j = s->name_off[i]; j < s->name_off[i + 1]; ++j) h = (h ^ (uint8_t)s->names[j]) * 1099511628211ull;
        s->name_hash[i] = {h, i};
    }
    std::sort(s->name_hash.begin(), s->name_hash.end());
    s->names_unique = true;
    s->lex_rank.clear();
    for (uint32_t i = 1; i < rs->n && s->names_unique; ++i)
        if (s->name_hash[i].first == s->name_hash[i - 1].first) {
            const uint32_t a = s->name_hash[i - 1].second, b = s->name_hash[i].second;
            const uint32_t la = s->name_off[a + 1] - s->name_off[a], lb = s->name_off[b + 1] - s->name_off[b];
            if (la == lb && memcmp(s->names.data() + s->name_off[a], s->names.data() + s->name_off[b], la) == 0) s->names_unique = false;
        }
}

void stage_readset(lrge_ctx *c, const lrge_readset *rs, lrge_staged *s, bool async = false, uint32_t blk_first = 0,
                   uint32_t blk_count = 0xffffffffu, bool defer_names = false) {
    s->n = rs->n;
    s->total_bases = rs->n ? rs->off[rs->n] - rs->off[0] : 0;
    s->pending = false;
    if (blk_first > rs->n) blk_first = rs->n;
    if (blk_count > rs->n - blk_first) blk_count = rs->n - blk_first;
    s->blk_first = blk_first; s->blk_count = blk_count;
    const uint64_t b0 = rs->n ? rs->off[blk_first] : 0, b1 = rs->n ? rs->off[blk_first + blk_count] : 0;
    s->blk_base = rs->n ? b0 - rs->off[0] : 0;
    // start the copies first; the host-side bookkeeping below overlaps with them
    s->asc.reserve((b1 - b0) + 64);
    CUDA_CHECK(cudaMemsetAsync(s->asc.as<uint8_t>() + (b1 - b0), 0, 64, c->copy_stream));   // slack read by the word-wise packers
    s->chunk_hi.clear();
    s->tab_valid = false;
    if (async && blk_count) {
        const uint32_t n = blk_count;
        std::vector<uint64_t> soff(n);
        std::vector<uint32_t> slen(n);
        s->tab_tile_off.resize((size_t)n + 1);
        uint32_t tiles = 0;
        uint64_t nbases = 0;
        for (uint32_t i = 0; i < n; ++i) {
            const uint64_t l = rs->off[blk_first + i + 1] - rs->off[blk_first + i];
            s->tab_tile_off[i] = tiles;
            soff[i] = rs->off[blk_first + i] - b0;
            slen[i] = (uint32_t)std::min<uint64_t>(l, 0xffffffffu);
            nbases += l;
            tiles += (uint32_t)((l + TILE_POS - 1) / TILE_POS);
        }
        s->tab_tile_off[n] = tiles;
        std::vector<uint32_t> tile_read(tiles);
        for (uint32_t i = 0; i < n; ++i) for (uint32_t t = s->tab_tile_off[i]; t < s->tab_tile_off[i + 1]; ++t) tile_read[t] = i;
        auto up = [&](DevBuf &d, const void *src, size_t bytes) {
            d.reserve(std::max<size_t>(bytes, 16));
            if (bytes) CUDA_CHECK(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        };
        up(s->t_tile_off, s->tab_tile_off.data(), ((size_t)n + 1) * 4);
        up(s->t_tile_read, tile_read.data(), (size_t)tiles * 4);
        up(s->t_sub_off, soff.data(), (size_t)n * 8);
        up(s->t_len, slen.data(), (size_t)n * 4);
        if (blk_first != 0 || blk_count != rs->n) {             // a proper block: read ids of its entries
            std::vector<uint32_t> rid_map(n);
            std::iota(rid_map.begin(), rid_map.end(), blk_first);
            up(s->t_rid_map, rid_map.data(), (size_t)n * 4);
        }
        s->tab_first = blk_first; s->tab_count = n; s->tab_tiles = tiles; s->tab_bases = nbases;
        s->tab_valid = true;
    }
    if (b1 > b0) {
        const uint64_t min_chunk = c->min_chunk;
        const uint32_t n_chunks = async ? (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(8, (b1 - b0) / min_chunk)) : 1;
        uint32_t lo = blk_first;
        for (uint32_t j = 0; j < n_chunks && lo < blk_first + blk_count; ++j) {
            uint32_t hi = blk_first + blk_count;
            if (j + 1 < n_chunks) {
                const uint64_t want = b0 + (b1 - b0) * (j + 1) / n_chunks;
                hi = (uint32_t)(std::upper_bound(rs->off + lo + 1, rs->off + blk_first + blk_count, want) - rs->off);
                hi = std::min(hi, blk_first + blk_count);
            }
            if (hi <= lo) continue;
            CUDA_CHECK(cudaMemcpyAsync(s->asc.as<uint8_t>() + (rs->off[lo] - b0), rs->bases + rs->off[lo], rs->off[hi] - rs->off[lo],
                                       cudaMemcpyHostToDevice, c->copy_stream));
            if (async) {
                if (s->chunk_ev.size() <= s->chunk_hi.size()) { cudaEvent_t e; CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); s->chunk_ev.push_back(e); }
                CUDA_CHECK(cudaEventRecord(s->chunk_ev[s->chunk_hi.size()], c->copy_stream));
            }
            s->chunk_hi.push_back(hi);
            lo = hi;
        }
    }
    s->h_off.assign(rs->off, rs->off + rs->n + 1);
    if (rs->n && rs->off[0] != 0) {
        const uint64_t b = rs->off[0];
        for (auto &o : s->h_off) o -= b;
    }
    s->h_len.resize(rs->n);
    s->max_len = 0;
    s->has_empty = false;
    for (uint32_t i = 0; i < rs->n; ++i) {
        const uint64_t l = s->h_off[i + 1] - s->h_off[i];
        if (l == 0) s->has_empty = true;                  // fatal only for a read that is MAPPED (aligner.rs:214-216), see run_overlap
        if (l >= (1u << 24)) { cudaStreamSynchronize(c->copy_stream); throw std::length_error("reads of 16 Mbases or more are not supported"); }
        s->h_len[i] = (uint32_t)l;
        s->max_len = std::max(s->max_len, (uint32_t)l);
    }
    s->names_src = rs;
    if (!defer_names) finish_staging(s);
    if (async) {
        if (!s->ready) CUDA_CHECK(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(s->ready, c->copy_stream));
        c->prof.mark("bases of a staged set have arrived (copy stream)", c->copy_stream);
        s->pending = true;
    } else CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
}

__global__ void gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t mul, uint32_t n,
                                  uint32_t *__restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(size_t)idx[i] * mul];
}
__global__ void gather_u32_base_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t mul, uint32_t n,
                                       uint32_t base, uint32_t *__restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[(size_t)idx[i] * mul] + base;
}
// sort keys of one half of the indexed set: hash << 2 | half -- equal hashes of different halves never tie, so the
// (unstable) merge of the two sorted runs reproduces the stable sort of the whole set
__global__ void key32_half_kernel(const uint64_t *__restrict__ x, uint64_t M, uint32_t g, uint32_t *__restrict__ k32) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) k32[i] = (uint32_t)(x[i] >> 8) << 2 | g;
}
struct LessU32 { __host__ __device__ bool operator()(uint32_t a, uint32_t b) const { return a < b; } };
__global__ void gather_u64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n,
                                  uint64_t *__restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// K1, lane-serial form (kernels_sketch.cuh: sketch2_kernel): ava-ont only (no HPC, k = 15).  One launch per staged
// chunk straight from the ASCII bases (no pack kernel), staging slots + compact_kernel as in the single-pass path.
// sketch2_begin launches and returns without waiting (false: not applicable / no room, nothing launched); sketch2_end
// waits for the minimizer total, then packs.  It returns false if a sub-tile overflowed its slot (very dense reads):
// the caller falls back to the exact two-pass path.
struct Sketch2Job {
    bool active = false; uint32_t n = 0, nw = 0; Sketch *out = nullptr; const uint32_t *d_tile_off = nullptr;
    bool want_k32 = false;      // in: also write the 32-bit index sort keys into c->ukey while compacting
    bool have_k32 = false;      // out: c->ukey holds them
};

// half: called after the launches of the first and of the second half of the staged chunks with the sub-tile range
// [w0, w1) and the reads [r_lo, r_hi) they cover (sketch_index_halves); the scan over all sub-tiles is then left out.
using HalfFn = std::function<bool(uint32_t w0, uint32_t w1, uint32_t r_lo, uint32_t r_hi, bool last)>;
bool sketch2_begin(lrge_ctx *c, const lrge_staged *s, const MapOpt &o, Sketch *out, uint32_t first, uint32_t count, Sketch2Job *job,
                   const HalfFn *half = nullptr) {
    job->active = false;
    if (o.is_hpc || o.k != 15 || c->sketch_legacy) return false;
    if (first > s->n) first = s->n;
    if (count > s->n - first) count = s->n - first;
    if (count && (first < s->blk_first || first + count > s->blk_first + s->blk_count))
        throw std::runtime_error("internal: sketch of reads that were not staged");
    const uint32_t n = count;
    // tables uploaded ahead of the bases by the async staging (same block): nothing to build or copy here
    const bool pre = s->tab_valid && s->pending && first == s->tab_first && count == s->tab_count;
    std::vector<uint64_t> soff;
    std::vector<uint32_t> tile_off_own, slen, rid_map, tile_read;
    uint32_t tiles = 0;
    uint64_t nbases = 0;
    if (pre) { tiles = s->tab_tiles; nbases = s->tab_bases; }
    else {
        soff.resize(n); tile_off_own.resize((size_t)n + 1); slen.resize(n); rid_map.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t r = first + i;
            tile_off_own[i] = tiles;
            soff[i] = s->h_off[r] - s->blk_base;
            slen[i] = s->h_len[r];
            rid_map[i] = r;
            nbases += s->h_len[r];
            tiles += (s->h_len[r] + TILE_POS - 1) / TILE_POS;
        }
        tile_off_own[n] = tiles;
        tile_read.resize(tiles);
        for (uint32_t i = 0; i < n; ++i) for (uint32_t t = tile_off_own[i]; t < tile_off_own[i + 1]; ++t) tile_read[t] = i;
    }
    const std::vector<uint32_t> &tile_off = pre ? s->tab_tile_off : tile_off_own;
    const bool subset = first != 0 || count != s->n;
    const uint32_t nw = tiles * (SK_THREADS / 32);
    const uint64_t stage_bytes = (uint64_t)nw * SK_CAP * 16;
    if (stage_bytes > (16ull << 30) || stage_bytes > (free_estimate(c) + c->stg_x.cap + c->stg_y.cap) / 4) return false;
    c->prof.sbegin(ST_SKETCH);
    if (!pre) {
        h2d(c, c->tile_off, tile_off.data(), (size_t)n + 1);
        h2d(c, c->tile_read, tile_read.data(), tiles);
        h2d(c, c->sub_off, soff.data(), n);
        h2d(c, c->len0, slen.data(), n);
        if (subset) h2d(c, c->rid_map, rid_map.data(), n);
    }
    const uint32_t *d_rid_map = !subset ? nullptr : pre ? s->t_rid_map.as<uint32_t>() : c->rid_map.as<uint32_t>();
    const uint32_t *d_tile_off = pre ? s->t_tile_off.as<uint32_t>() : c->tile_off.as<uint32_t>();
    out->moff.reserve(((size_t)n + 1) * 4);
    out->M = 0;
    c->wcnt.reserve(((size_t)nw + 1) * 4);
    c->woff.reserve(((size_t)nw + 1) * 4);
    job->n = n; job->nw = nw; job->out = out; job->active = true; job->d_tile_off = d_tile_off;
    if (tiles == 0) { CUDA_CHECK(cudaMemsetAsync(out->moff.p, 0, ((size_t)n + 1) * 4, c->stream)); c->sk_total() = 0; c->sk_ovf() = 0; c->prof.send(ST_SKETCH); return true; }
    c->stg_x.reserve(stage_bytes / 2 + 64); c->stg_y.reserve(stage_bytes / 2 + 64);
    c->xflag.reserve(16);
    CUDA_CHECK(cudaMemsetAsync(c->xflag.p, 0, 16, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->wcnt.as<uint32_t>() + nw, 0, 4, c->stream));
    Sketch2Args A{};
    A.asc = s->asc.as<uint8_t>();
    A.off = pre ? s->t_sub_off.as<uint64_t>() : c->sub_off.as<uint64_t>();
    A.len = pre ? s->t_len.as<uint32_t>() : c->len0.as<uint32_t>();
    A.rid_map = d_rid_map; A.tile_off = d_tile_off;
    A.tile_read = pre ? s->t_tile_read.as<uint32_t>() : c->tile_read.as<uint32_t>();
    A.ox = c->stg_x.as<uint64_t>(); A.oy = c->stg_y.as<uint64_t>(); A.wcnt = c->wcnt.as<uint32_t>(); A.overflow = c->xflag.as<uint32_t>();
    if (half) {
        // two launches groups, each followed by its own compaction / sort (the callback), instead of one scan at the end
        const size_t nc = s->chunk_hi.size(), mid = nc / 2;
        uint32_t lo = first, w_prev = 0, r_prev = first;
        for (size_t j = 0; j < nc; ++j) {
            const uint32_t hi = std::min(s->chunk_hi[j], first + count);
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, s->chunk_ev[j], 0));
            if (hi > lo) {
                A.tile0 = tile_off[lo - first];
                const uint32_t t = tile_off[hi - first] - A.tile0;
                if (t) KLAUNCH(c, "sketch2_kernel", (double)(hi - lo) / n * ((double)nbases / 4 + 16.0 * nbases / 3), (sketch2_kernel<15><<<t, SK_THREADS, 0, c->stream>>>(A)));
                lo = hi;
            }
            if (j + 1 == mid || j + 1 == nc) {
                const bool last = j + 1 == nc;
                const uint32_t r_hi = last ? first + count : lo;
                const uint32_t w1 = tile_off[r_hi - first] * (SK_THREADS / 32);
                CUDA_CHECK(cudaGetLastError());
                if (!(*half)(w_prev, w1, r_prev, r_hi, last)) { c->prof.send(ST_SKETCH); job->active = false; return false; }
                w_prev = w1; r_prev = r_hi;
            }
        }
        c->prof.send(ST_SKETCH);
        return true;
    }
    KLAUNCH(c, "sketch2_kernel", (double)nbases / 4 + 16.0 * nbases / 3, {     // SURVEY.md 8(d) K1 (M ~ bases / 3 at launch time)
        if (s->pending && !s->chunk_hi.empty()) {
            // one launch per staged chunk, each waiting only for its own copy
            uint32_t lo = first;
            for (size_t j = 0; j < s->chunk_hi.size() && lo < first + count; ++j) {
                const uint32_t hi = std::min(s->chunk_hi[j], first + count);
                CUDA_CHECK(cudaStreamWaitEvent(c->stream, s->chunk_ev[j], 0));
                if (hi <= lo) continue;
                A.tile0 = tile_off[lo - first];
                const uint32_t t = tile_off[hi - first] - A.tile0;
                if (t) sketch2_kernel<15><<<t, SK_THREADS, 0, c->stream>>>(A);
                lo = hi;
            }
        } else {
            if (s->pending) CUDA_CHECK(cudaStreamWaitEvent(c->stream, s->ready, 0));
            A.tile0 = 0;
            sketch2_kernel<15><<<tiles, SK_THREADS, 0, c->stream>>>(A);
        }
    });
    CUDA_CHECK(cudaGetLastError());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, c->wcnt.as<uint32_t>(), c->woff.as<uint32_t>(), nw + 1, c->stream);
    KLAUNCH(c, "cub::DeviceScan(u32)", 8.0 * nw,
            cub::DeviceScan::ExclusiveSum(cub_temp(c, tb), tb, c->wcnt.as<uint32_t>(), c->woff.as<uint32_t>(), nw + 1, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&c->sk_total(), c->woff.as<uint32_t>() + nw, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&c->sk_ovf(), c->xflag.p, 4, cudaMemcpyDeviceToHost, c->stream));
    c->prof.send(ST_SKETCH);
    return true;
}

bool sketch2_end(lrge_ctx *c, Sketch2Job &job) {
    Sketch *out = job.out;
    const uint32_t n = job.n, nw = job.nw;
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (nw == 0) return true;
    if (c->sk_ovf()) return false;
    const uint32_t total = c->sk_total();
    c->prof.sbegin(ST_SKETCH);
    out->M = total;
    out->x.reserve(((size_t)total + 1) * 8);
    out->y.reserve(((size_t)total + 1) * 8);
    uint32_t *k32 = nullptr;
    if (job.want_k32) { c->ukey.reserve(((size_t)total + 1) * 8); k32 = c->ukey.as<uint32_t>(); }
    KLAUNCH(c, "compact_kernel", 32.0 * total + (k32 ? 4.0 * total : 0.0),
            compact_kernel<<<(nw + 7) / 8, 256, 0, c->stream>>>(c->stg_x.as<uint64_t>(), c->stg_y.as<uint64_t>(), c->wcnt.as<uint32_t>(),
                                                               c->woff.as<uint32_t>(), nw, SK_CAP, out->x.as<uint64_t>(), out->y.as<uint64_t>(), k32));
    job.have_k32 = k32 != nullptr;
    // per-read minimizer offsets: woff at the first warp sub-tile of every read
    KLAUNCH(c, "gather_kernel", 12.0 * n,
            gather_u32_kernel<<<(n + 1 + 255) / 256, 256, 0, c->stream>>>(c->woff.as<uint32_t>(), job.d_tile_off,
                                                                         SK_THREADS / 32, n + 1, out->moff.as<uint32_t>()));
    CUDA_CHECK(cudaGetLastError());
    c->prof.send(ST_SKETCH);
    return true;
}

// K0 + K1: sketch a staged read set, or only its reads [first, first + count), into `out`
void sketch_readset(lrge_ctx *c, const lrge_staged *s, const MapOpt &o, Sketch *out, uint32_t first = 0, uint32_t count = 0xffffffffu,
                    bool skip_fast = false) {
    if (first > s->n) first = s->n;
    if (count > s->n - first) count = s->n - first;
    if (count && (first < s->blk_first || first + count > s->blk_first + s->blk_count))
        throw std::runtime_error("internal: sketch of reads that were not staged");
    if (skip_fast) { /* the caller already tried the lane-serial kernel */ }
    else { Sketch2Job job; if (sketch2_begin(c, s, o, out, first, count, &job) && sketch2_end(c, job)) return; }
    const bool subset = first != 0 || count != s->n;
    const uint32_t n = count;
    std::vector<uint64_t> poff(n), soff(n);
    std::vector<uint32_t> tile_off(n + 1), slen(n), rid_map(n);
    uint64_t pos = PAD_BASES;
    uint32_t tiles = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t r = first + i;
        poff[i] = pos;
        tile_off[i] = tiles;
        soff[i] = s->h_off[r] - s->blk_base;
        slen[i] = s->h_len[r];
        rid_map[i] = r;
        pos += ((uint64_t)s->h_len[r] + PAD_BASES - 1) / PAD_BASES * PAD_BASES;
        tiles += (s->h_len[r] + TILE_POS - 1) / TILE_POS;
    }
    tile_off[n] = tiles;
    const uint64_t padded_words = (pos + TILE_POS + 2 * PAD_BASES) / 32 + 8;
    if (s->pending) CUDA_CHECK(cudaStreamWaitEvent(c->stream, s->ready, 0));
    c->prof.sbegin(ST_PACK);
    uint64_t nbases = 0;
    for (uint32_t i = 0; i < n; ++i) nbases += slen[i];
    h2d(c, c->poff, poff.data(), n);
    h2d(c, c->tile_off, tile_off.data(), n + 1);
    c->w2.reserve(padded_words * 8);
    c->nm.reserve(padded_words * 4);
    const uint32_t *len_dev;
    ReadLayout L{};
    h2d(c, c->sub_off, soff.data(), n);
    h2d(c, c->len0, slen.data(), n);
    if (subset) h2d(c, c->rid_map, rid_map.data(), n);
    L.off = c->sub_off.as<uint64_t>();
    L.rid_map = subset ? c->rid_map.as<uint32_t>() : nullptr;
    L.poff = c->poff.as<uint64_t>();
    L.tile_off = c->tile_off.as<uint32_t>();
    L.w2 = c->w2.as<uint64_t>();
    L.nm = c->nm.as<uint32_t>();
    L.n = n;
    if (o.is_hpc) {
        const uint64_t resident = s->h_off[s->blk_first + s->blk_count] - s->blk_base;
        c->casc.reserve(resident + 64);
        c->cend.reserve((resident + 1) * 4);
        c->clen.reserve((size_t)std::max(n, 1u) * 4);
        // compressed reads keep their original offsets: the gaps behind them are read (and masked) by pack_kernel's word loads
        CUDA_CHECK(cudaMemsetAsync(c->casc.p, 0, resident + 64, c->stream));
        if (n) {
            KLAUNCH(c, "hpc_compress_kernel", 6 * nbases,
                    hpc_compress_kernel<<<(n + 7) / 8, 256, 0, c->stream>>>(s->asc.as<uint8_t>(), c->sub_off.as<uint64_t>(), c->len0.as<uint32_t>(), n,
                                                                            c->casc.as<uint8_t>(), c->cend.as<uint32_t>(), c->clen.as<uint32_t>()));
        }
        L.asc = c->casc.as<uint8_t>();
        len_dev = c->clen.as<uint32_t>();
    } else {
        L.asc = s->asc.as<uint8_t>();
        len_dev = c->len0.as<uint32_t>();
    }
    L.len = len_dev;
    if (tiles) KLAUNCH(c, "pack_kernel", nbases + nbases * 3 / 8, pack_kernel<<<tiles, SK_THREADS, 0, c->stream>>>(L));
    CUDA_CHECK(cudaGetLastError());
    c->prof.send(ST_PACK);

    c->prof.sbegin(ST_SKETCH);
    const uint32_t nw = tiles * (SK_THREADS / 32);
    c->wcnt.reserve(((size_t)nw + 1) * 4);
    c->woff.reserve(((size_t)nw + 1) * 4);
    SketchArgs A{};
    A.L = L; A.cend = c->cend.as<uint32_t>(); A.wcnt = c->wcnt.as<uint32_t>(); A.woff = c->woff.as<uint32_t>(); A.k = o.k;
    out->M = 0;
    out->moff.reserve(((size_t)n + 1) * 4);
    if (tiles) {
        CUDA_CHECK(cudaMemsetAsync(c->wcnt.p, 0, ((size_t)nw + 1) * 4, c->stream));
        const bool x32 = !o.is_hpc && 2 * o.k <= 31;
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, c->wcnt.as<uint32_t>(), c->woff.as<uint32_t>(), nw + 1, c->stream);
        // single pass when the staging slots (SK_CAP minimizers per 1024-step sub-tile) fit comfortably
        const uint64_t stage_bytes = (uint64_t)nw * SK_CAP * 16;
        const size_t free_b = free_estimate(c);
        bool single = !c->sketch_two_pass && stage_bytes <= (16ull << 30) && stage_bytes <= (free_b + c->stg_x.cap + c->stg_y.cap) / 4;
        uint32_t total = 0;
        if (single) {
            c->stg_x.reserve(stage_bytes / 2 + 64); c->stg_y.reserve(stage_bytes / 2 + 64);
            c->xflag.reserve(16);
            CUDA_CHECK(cudaMemsetAsync(c->xflag.p, 0, 16, c->stream));
            A.ox = c->stg_x.as<uint64_t>(); A.oy = c->stg_y.as<uint64_t>(); A.cap = SK_CAP; A.overflow = c->xflag.as<uint32_t>();
            KLAUNCH(c, "sketch_kernel<stage>", nbases * 3 / 8 + 16.0 * nbases / 3, {
                if (o.is_hpc) sketch_kernel<true, 2, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else if (x32) sketch_kernel<false, 2, true><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else sketch_kernel<false, 2, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
            });
            KLAUNCH(c, "cub::DeviceScan(u32)", 8.0 * nw,
                    cub::DeviceScan::ExclusiveSum(cub_temp(c, tb), tb, c->wcnt.as<uint32_t>(), c->woff.as<uint32_t>(), nw + 1, c->stream));
            uint32_t ovf = 0;
            CUDA_CHECK(cudaMemcpyAsync(&total, c->woff.as<uint32_t>() + nw, 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(&ovf, c->xflag.p, 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            if (ovf) single = false;                          // a very dense sub-tile: redo with the two-pass path
            else {
                out->M = total;
                out->x.reserve(((size_t)total + 1) * 8);
                out->y.reserve(((size_t)total + 1) * 8);
                KLAUNCH(c, "compact_kernel", 32.0 * total,
                        compact_kernel<<<(nw + 7) / 8, 256, 0, c->stream>>>(c->stg_x.as<uint64_t>(), c->stg_y.as<uint64_t>(), c->wcnt.as<uint32_t>(),
                                                                           c->woff.as<uint32_t>(), nw, SK_CAP, out->x.as<uint64_t>(), out->y.as<uint64_t>()));
            }
        }
        if (!single) {
            KLAUNCH(c, "sketch_kernel<count>", nbases * 3 / 8, {
                if (o.is_hpc) sketch_kernel<true, 0, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else if (x32) sketch_kernel<false, 0, true><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else sketch_kernel<false, 0, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
            });
            KLAUNCH(c, "cub::DeviceScan(u32)", 8.0 * nw,
                    cub::DeviceScan::ExclusiveSum(cub_temp(c, tb), tb, c->wcnt.as<uint32_t>(), c->woff.as<uint32_t>(), nw + 1, c->stream));
            CUDA_CHECK(cudaMemcpyAsync(&total, c->woff.as<uint32_t>() + nw, 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            out->M = total;
            out->x.reserve(((size_t)total + 1) * 8);
            out->y.reserve(((size_t)total + 1) * 8);
            A.ox = out->x.as<uint64_t>(); A.oy = out->y.as<uint64_t>();
            KLAUNCH(c, "sketch_kernel<emit>", nbases * 3 / 8 + 16.0 * total, {
                if (o.is_hpc) sketch_kernel<true, 1, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else if (x32) sketch_kernel<false, 1, true><<<tiles, SK_THREADS, 0, c->stream>>>(A);
                else sketch_kernel<false, 1, false><<<tiles, SK_THREADS, 0, c->stream>>>(A);
            });
        }
        // per-read minimizer offsets: woff at the first warp sub-tile of every read
        KLAUNCH(c, "gather_kernel", 12.0 * n,
                gather_u32_kernel<<<(n + 1 + 255) / 256, 256, 0, c->stream>>>(c->woff.as<uint32_t>(), c->tile_off.as<uint32_t>(),
                                                                             SK_THREADS / 32, n + 1, out->moff.as<uint32_t>()));
    } else {
        uint32_t z = 0;
        h2d(c, out->moff, &z, 1);
    }
    CUDA_CHECK(cudaGetLastError());
    c->prof.send(ST_SKETCH);
}

struct BuiltIndex { IndexView view; int mid_occ; uint64_t n_keys; bool x32 = false; };

// K2 + K3.  build_index_launch queues the sort, the bucket table and the occurrence histogram (+ its copy to pinned
// host memory) and returns; build_index_finish -- after any later synchronisation of the stream -- turns the histogram
// into mid_occ on the host.  Nothing in between may depend on mid_occ (the sketch of the streamed set does not).
BuiltIndex build_index_launch(lrge_ctx *c, const Sketch &sk, const MapOpt &o, bool have_k32 = false) {   // have_k32: c->ukey holds the keys
    BuiltIndex B{};
    const uint64_t M = sk.M;
    if (M >= (1ull << 32) - 2) throw std::length_error("more than 2^32 target minimizers per GPU are not supported");
    // ~8 entries per bucket: the table (4 B x 2^bits) stays L2-resident, a lookup bisects <= 3 steps
    int bits = 8;
    while (bits < 2 * o.k && bits < 28 && (8ull << bits) < M) ++bits;
    bits = std::max(8, std::min({bits + c->index_bits_delta, 2 * o.k, 28}));   // LRGE_INDEX_BITS_DELTA: experiments
    if (bits < 2 * o.k - 32) bits = 2 * o.k - 32;              // tk32: the in-bucket part of a key must fit 32 bits
    const uint32_t nb = 1u << bits;
    c->boff.reserve(((size_t)nb + 2) * 4);
    c->hist.reserve((OCC_BINS + 2) * 8);
    CUDA_CHECK(cudaMemsetAsync(c->hist.p, 0, (OCC_BINS + 2) * 8, c->stream));
    unsigned long long *hist = c->hist.as<unsigned long long>();
    c->ty.reserve((M + 1) * 8);
    c->tk32.reserve((M + 64) * 4);
    const bool k32 = !o.is_hpc && 2 * o.k <= 32;               // ava-ont: the key is the 30-bit hash alone (span = k)
    B.x32 = k32;
    size_t tb = 0;
    if (k32) {
        // sort (32-bit hash, y): 12-byte records instead of 16, and the sorted keys ARE the compact lookup keys
        c->ukey.reserve((M + 1) * 4);
        if (M) {
            if (!have_k32) KLAUNCH(c, "key32_kernel", 12.0 * M,
                                   key32_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(sk.x.as<uint64_t>(), M, c->ukey.as<uint32_t>()));
            cub::DeviceRadixSort::SortPairs(nullptr, tb, c->ukey.as<uint32_t>(), c->tk32.as<uint32_t>(), sk.y.as<uint64_t>(), c->ty.as<uint64_t>(),
                                            (int64_t)M, 0, 2 * o.k, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(index: u32 key, u64 value)", 32.0 * M,       // 2 x 16 x M_T
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, c->ukey.as<uint32_t>(), c->tk32.as<uint32_t>(), sk.y.as<uint64_t>(),
                                                    c->ty.as<uint64_t>(), (int64_t)M, 0, 2 * o.k, c->stream));
            c->launches += (2 * o.k + 7) / 8;
            c->prof.set_passes("cub::DeviceRadixSort(index: u32 key, u64 value)", (2 * o.k + 7) / 8);
        }
        CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));          // slack read by the 8-wide bucket fetch
        const int shift = 2 * o.k - bits;
        KLAUNCH(c, "index_post_kernel", 4.0 * M + 4.0 * nb,
                (index_post_kernel<uint32_t, 0><<<148 * 8, 256, 0, c->stream>>>(c->tk32.as<uint32_t>(), M, shift, nb, c->boff.as<uint32_t>(), nullptr, hist)));
        B.view.tx = nullptr; B.view.shift = 8 + shift;          // index_lookup shifts hash << 8
    } else {
        c->tx.reserve((M + 1) * 8);
        if (M) {
            cub::DeviceRadixSort::SortPairs(nullptr, tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                            c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(index: u64 key, u64 value)", 32.0 * M,
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                                    c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream));
            c->launches += (2 * o.k + 7) / 8;
            c->prof.set_passes("cub::DeviceRadixSort(index: u64 key, u64 value)", (2 * o.k + 7) / 8);
        }
        CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));          // slack read by the 8-wide bucket fetch
        const int shift = 8 + 2 * o.k - bits;
        KLAUNCH(c, "index_post_kernel", 12.0 * M + 4.0 * nb,
                (index_post_kernel<uint64_t, 8><<<148 * 8, 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, shift, nb, c->boff.as<uint32_t>(), c->tk32.as<uint32_t>(), hist)));
        B.view.tx = c->tx.as<uint64_t>(); B.view.shift = shift;
    }
    CUDA_CHECK(cudaMemcpyAsync(c->hist_h(), hist, (OCC_BINS + 2) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaGetLastError());
    B.view.tk32 = c->tk32.as<uint32_t>(); B.view.ty = c->ty.as<uint64_t>(); B.view.boff = c->boff.as<uint32_t>();
    B.view.M = M; B.view.bl = nullptr;
    B.mid_occ = -1; B.n_keys = 0;
    return B;
}

void build_index_finish(lrge_ctx *c, const MapOpt &o, BuiltIndex *Bp) {       // the stream has been synchronised since the launch
    BuiltIndex &B = *Bp;
    const uint64_t M = B.view.M;
    const unsigned long long *h = c->hist_h();
    unsigned long long *hist = c->hist.as<unsigned long long>();
    const uint64_t n_keys = h[OCC_BINS];
    int mid_occ;
    if (n_keys == 0) mid_occ = INT32_MAX;
    else {
        const uint64_t rank = (uint32_t)((1. - (double)o.mid_occ_frac) * (double)n_keys);   // (uint32_t)((1. - f) * n)
        uint64_t cum = 0; int64_t val = -1;
        for (int b = 0; b < OCC_BINS - 1; ++b) { cum += h[b]; if (cum > rank) { val = b; break; } }
        if (val < 0) {   // rank lies among keys occurring >= OCC_BINS-1 times: exact order statistic of those
            const uint64_t n_over = h[OCC_BINS - 1];
            DevBuf over; over.reserve((n_over + 1) * 4);
            CUDA_CHECK(cudaMemsetAsync(hist + OCC_BINS + 1, 0, 8, c->stream));
            if (B.x32) occ_overflow_kernel<uint32_t, 0><<<148 * 4, 256, 0, c->stream>>>(c->tk32.as<uint32_t>(), M, over.as<uint32_t>(), hist + OCC_BINS + 1, n_over);
            else occ_overflow_kernel<uint64_t, 8><<<148 * 4, 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, over.as<uint32_t>(), hist + OCC_BINS + 1, n_over);
            ++c->launches;
            std::vector<uint32_t> ov(n_over);
            CUDA_CHECK(cudaMemcpyAsync(ov.data(), over.p, n_over * 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            std::sort(ov.begin(), ov.end());
            val = ov[rank - cum];
        }
        mid_occ = (int)val + 1;
    }
    if (mid_occ < o.min_mid_occ) mid_occ = o.min_mid_occ;
    if (o.max_mid_occ > o.min_mid_occ && mid_occ > o.max_mid_occ) mid_occ = o.max_mid_occ;
    B.mid_occ = mid_occ; B.n_keys = n_keys;
}

BuiltIndex build_index(lrge_ctx *c, const Sketch &sk, const MapOpt &o) {
    BuiltIndex B = build_index_launch(c, sk, o);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    build_index_finish(c, o, &B);
    return B;
}

// Host-buffer path, ava-ont, one GPU: the indexed set arrives in chunks (stage_readset, async).  Its sketch, compaction
// and index sort are done per HALF of the chunks, so the first half's index work runs while the second half is still on
// the wire; the two sorted runs are merged (keys carry the half in their low bits, see key32_half_kernel) and
// index_post_kernel strips the tag.  Leaves `out` complete (x, y, moff) and the index launched like build_index_launch.
// host_work runs once while the first half is in flight.  false: not applicable or a capacity was exceeded -- the
// caller takes the ordinary path (nothing it depends on has been touched).
bool sketch_index_halves(lrge_ctx *c, const lrge_staged *s, const MapOpt &o, Sketch *out, const std::function<void()> &host_work,
                         BuiltIndex *Bout) {
    if (c->no_halves || c->sketch_legacy || o.is_hpc || o.k != 15 || !s->pending || s->chunk_hi.size() < 2) return false;
    if (s->blk_first != 0 || s->blk_count != s->n || s->n == 0) return false;
    const uint32_t n = s->n;
    const uint64_t cap = s->total_bases / 2 + 4096;               // minimizers (density ~ 1/3 at w = 5)
    if (cap >= (1ull << 31) - 64) return false;
    if (cap * 64 > free_estimate(c) + out->x.cap + out->y.cap + c->ty.cap + c->tk32.cap + c->hk_a.cap + c->hk_b.cap + c->hv_b.cap + c->ukey.cap) return false;
    out->x.reserve((cap + 1) * 8); out->y.reserve((cap + 1) * 8);
    c->ukey.reserve((cap + 1) * 4);
    c->hk_a.reserve((cap + 1) * 4); c->hk_b.reserve((cap + 1) * 4); c->hv_b.reserve((cap + 1) * 8);
    c->ty.reserve((cap + 1) * 8); c->tk32.reserve((cap + 64) * 4);
    uint64_t base = 0;
    uint32_t run[2] = {0, 0}, g = 0;
    bool did_host = false;
    Sketch2Job job;
    const HalfFn half = [&](uint32_t w0, uint32_t w1, uint32_t r_lo, uint32_t r_hi, bool last) -> bool {
        const uint32_t nwg = w1 - w0;
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, c->wcnt.as<uint32_t>() + w0, c->woff.as<uint32_t>() + w0, nwg + 1, c->stream);
        KLAUNCH(c, "cub::DeviceScan(u32)", 8.0 * nwg,
                cub::DeviceScan::ExclusiveSum(cub_temp(c, tb), tb, c->wcnt.as<uint32_t>() + w0, c->woff.as<uint32_t>() + w0, nwg + 1, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(&c->sk_total(), c->woff.as<uint32_t>() + w1, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(&c->sk_ovf(), c->xflag.p, 4, cudaMemcpyDeviceToHost, c->stream));
        if (!did_host) { host_work(); did_host = true; }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (c->sk_ovf()) return false;
        const uint32_t total = c->sk_total();
        if (base + total > cap) return false;
        uint64_t *x = out->x.as<uint64_t>() + base, *y = out->y.as<uint64_t>() + base;
        if (nwg) KLAUNCH(c, "compact_kernel", 32.0 * total,
                         compact_kernel<<<(nwg + 7) / 8, 256, 0, c->stream>>>(c->stg_x.as<uint64_t>() + (size_t)w0 * SK_CAP, c->stg_y.as<uint64_t>() + (size_t)w0 * SK_CAP,
                                                                             c->wcnt.as<uint32_t>() + w0, c->woff.as<uint32_t>() + w0, nwg, SK_CAP, x, y));
        const uint32_t ng = r_hi - r_lo + (last ? 1u : 0u);
        if (ng) KLAUNCH(c, "gather_kernel", 12.0 * ng,
                        gather_u32_base_kernel<<<(ng + 255) / 256, 256, 0, c->stream>>>(c->woff.as<uint32_t>(), job.d_tile_off + r_lo, SK_THREADS / 32, ng,
                                                                                         (uint32_t)base, out->moff.as<uint32_t>() + r_lo));
        if (last) { CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream)); c->fork_recorded = true; }   // the staging slots are free again
        if (total) {
            uint32_t *kin = c->ukey.as<uint32_t>() + base;
            KLAUNCH(c, "key32_kernel", 12.0 * total, key32_half_kernel<<<(total + 255) / 256, 256, 0, c->stream>>>(x, total, g, kin));
            cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, c->hk_b.as<uint32_t>() + base, y, c->hv_b.as<uint64_t>() + base, (int64_t)total, 0, 32, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(index: u32 key, u64 value)", 32.0 * total,
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, kin, c->hk_b.as<uint32_t>() + base, y, c->hv_b.as<uint64_t>() + base, (int64_t)total, 0, 32, c->stream));
            c->launches += 4;
            c->prof.set_passes("cub::DeviceRadixSort(index: u32 key, u64 value)", 4);
        }
        run[g] = total; base += total; ++g;
        return true;
    };
    if (!sketch2_begin(c, s, o, out, 0, n, &job, &half) || g != 2) { if (!did_host) host_work(); return false; }
    const uint64_t M = base;
    out->M = M;
    // merge the two runs; strip the tag, bucket table, occurrence histogram
    BuiltIndex B{};
    int bits = 8;
    while (bits < 2 * o.k && bits < 28 && (8ull << bits) < M) ++bits;
    const uint32_t nb = 1u << bits;
    c->boff.reserve(((size_t)nb + 2) * 4);
    c->hist.reserve((OCC_BINS + 2) * 8);
    c->prof.sbegin(ST_INDEX);
    CUDA_CHECK(cudaMemsetAsync(c->hist.p, 0, (OCC_BINS + 2) * 8, c->stream));
    if (M) {
        size_t tb = 0;
        cub::DeviceMerge::MergePairs(nullptr, tb, c->hk_b.as<uint32_t>(), c->hv_b.as<uint64_t>(), (int)run[0], c->hk_b.as<uint32_t>() + run[0],
                                     c->hv_b.as<uint64_t>() + run[0], (int)run[1], c->hk_a.as<uint32_t>(), c->ty.as<uint64_t>(), LessU32{}, c->stream);
        KLAUNCH(c, "cub::DeviceMerge(index halves)", 24.0 * M,
                cub::DeviceMerge::MergePairs(cub_temp(c, tb), tb, c->hk_b.as<uint32_t>(), c->hv_b.as<uint64_t>(), (int)run[0], c->hk_b.as<uint32_t>() + run[0],
                                             c->hv_b.as<uint64_t>() + run[0], (int)run[1], c->hk_a.as<uint32_t>(), c->ty.as<uint64_t>(), LessU32{}, c->stream));
    }
    CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));
    const int shift = 2 * o.k - bits;
    KLAUNCH(c, "index_post_kernel", 8.0 * M + 4.0 * nb,
            (index_post_kernel<uint32_t, 2><<<148 * 8, 256, 0, c->stream>>>(c->hk_a.as<uint32_t>(), M, shift + 2, nb, c->boff.as<uint32_t>(), c->tk32.as<uint32_t>(),
                                                                            c->hist.as<unsigned long long>())));
    CUDA_CHECK(cudaMemcpyAsync(c->hist_h(), c->hist.p, (OCC_BINS + 2) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaGetLastError());
    c->prof.send(ST_INDEX);
    B.x32 = true;
    B.view.tx = nullptr; B.view.shift = 8 + shift;
    B.view.tk32 = c->tk32.as<uint32_t>(); B.view.ty = c->ty.as<uint64_t>(); B.view.boff = c->boff.as<uint32_t>();
    B.view.M = M; B.view.bl = nullptr;
    B.mid_occ = -1; B.n_keys = 0;
    *Bout = B;
    return true;
}

#define NCCL_CHECK(x) do { ncclResult_t _r = (x); if (_r != ncclSuccess) throw NcclError(std::string(#x) + ": " + ncclGetErrorString(_r)); } while (0)

// K2 + K3 with the index sharded over ranks: `sk` holds the minimizers of THIS rank's targets only.
// Exact global occurrence counts: distinct local keys (key, count) go to the rank owning their key range,
// are summed there, the occurrence histogram is all-reduced (=> the same mid_occ on every rank), and the keys
// above mid_occ come back as a blacklist marked in a bitmap over the local runs.
BuiltIndex build_index_sharded_exact(lrge_ctx *c, const Sketch &sk, const MapOpt &o, lrge_stats *S) {
    BuiltIndex B{};
    const int W = c->world, me = c->rank;
    const uint64_t M = sk.M;
    if (M >= (1ull << 32) - 2) throw std::length_error("more than 2^32 target minimizers per GPU are not supported");
    c->tx.reserve((M + 1) * 8);
    c->ty.reserve((M + 1) * 8);
    size_t tb = 0;
    if (M) {
        cub::DeviceRadixSort::SortPairs(nullptr, tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                        c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream);
        KLAUNCH(c, "cub::DeviceRadixSort(index: u64 key, u64 value)", 32.0 * M,
                cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                                c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream));
        c->launches += (2 * o.k + 7) / 8;
    }
    int bits = 8;
    while (bits < 2 * o.k && bits < 28 && (8ull << bits) < M) ++bits;
    bits = std::max(8, std::min({bits + c->index_bits_delta, 2 * o.k, 28}));   // LRGE_INDEX_BITS_DELTA: experiments
    if (bits < 2 * o.k - 32) bits = 2 * o.k - 32;              // tk32: the in-bucket part of a key must fit 32 bits
    const uint32_t nb = 1u << bits;
    const int shift = 8 + 2 * o.k - bits;
    c->boff.reserve(((size_t)nb + 2) * 4);
    KLAUNCH(c, "bucket_bounds_kernel", 8.0 * M + 4.0 * nb,
            bucket_bounds_kernel<<<(unsigned)((M + 1 + 255) / 256), 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, shift, nb, c->boff.as<uint32_t>()));

    c->tk32.reserve((M + 64) * 4);
    CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));
    if (M) KLAUNCH(c, "tk32_kernel", 12.0 * M, tk32_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, c->tk32.as<uint32_t>()));
    // local run-length encoding: (key, count) per distinct local key
    c->flags.reserve(M + 1);
    c->pstart.reserve((M + 2) * 4);
    c->npairs.reserve(16);
    uint32_t U = 0;
    if (M) {
        KLAUNCH(c, "run_flag_kernel", 9.0 * M, run_flag_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, 8, c->flags.as<uint8_t>()));
        cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(), c->pstart.as<uint32_t>(),
                                   c->npairs.as<uint32_t>(), (int64_t)M, c->stream);
        cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(),
                                   c->pstart.as<uint32_t>(), c->npairs.as<uint32_t>(), (int64_t)M, c->stream);
        c->launches += 3;
        CUDA_CHECK(cudaMemcpyAsync(&U, c->npairs.p, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const uint32_t sentinel = (uint32_t)M;
        CUDA_CHECK(cudaMemcpyAsync(c->pstart.as<uint32_t>() + U, &sentinel, 4, cudaMemcpyHostToDevice, c->stream));
    }
    c->ukey.reserve(((size_t)U + 1) * 8);
    c->ucnt.reserve(((size_t)U + 1) * 4);
    if (U) { rle_emit_kernel<<<(U + 255) / 256, 256, 0, c->stream>>>(c->tx.as<uint64_t>(), c->pstart.as<uint32_t>(), U, c->ukey.as<uint64_t>(), c->ucnt.as<uint32_t>()); ++c->launches; }

    // who owns what: bounds[W+1] over the local runs, then the W x W matrix of send counts
    c->xmeta.reserve((size_t)(W + 1) * 8 + (size_t)W * W * 8 + (OCC_BINS + 8) * 8 + 64);
    unsigned long long *bounds_d = c->xmeta.as<unsigned long long>();
    unsigned long long *mat_d = bounds_d + (W + 1);
    unsigned long long *hist_d = mat_d + (size_t)W * W;            // OCC_BINS bins + [OCC_BINS] n_keys + [OCC_BINS+1] minimizers + scratch
    owner_bounds_kernel<<<1, 32, 0, c->stream>>>(c->ukey.as<uint64_t>(), U, W, 2 * o.k, bounds_d);
    ++c->launches;
    std::vector<unsigned long long> bounds(W + 1), mat((size_t)W * W), sendc(W);
    CUDA_CHECK(cudaMemcpyAsync(bounds.data(), bounds_d, (W + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int d = 0; d < W; ++d) sendc[d] = bounds[d + 1] - bounds[d];
    CUDA_CHECK(cudaMemcpyAsync(hist_d + OCC_BINS + 8, sendc.data(), W * 8, cudaMemcpyHostToDevice, c->stream));   // staging slot
    NCCL_CHECK(ncclAllGather(hist_d + OCC_BINS + 8, mat_d, W, ncclUint64, c->comm, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(mat.data(), mat_d, (size_t)W * W * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> roff(W + 1, 0);
    for (int s2 = 0; s2 < W; ++s2) roff[s2 + 1] = roff[s2] + mat[(size_t)s2 * W + me];
    const uint64_t R = roff[W];
    if (R >= (1ull << 32) - 2) throw std::length_error("too many distinct keys per owner rank");
    c->rkey.reserve((R + 1) * 8); c->rcnt.reserve((R + 1) * 4);
    c->skey.reserve((R + 1) * 8); c->scnt.reserve((R + 1) * 4);
    NCCL_CHECK(ncclGroupStart());
    for (int d = 0; d < W; ++d) if (sendc[d]) {
        NCCL_CHECK(ncclSend(c->ukey.as<uint64_t>() + bounds[d], sendc[d], ncclUint64, d, c->comm, c->stream));
        NCCL_CHECK(ncclSend(c->ucnt.as<uint32_t>() + bounds[d], sendc[d], ncclUint32, d, c->comm, c->stream));
    }
    for (int s2 = 0; s2 < W; ++s2) if (roff[s2 + 1] > roff[s2]) {
        NCCL_CHECK(ncclRecv(c->rkey.as<uint64_t>() + roff[s2], roff[s2 + 1] - roff[s2], ncclUint64, s2, c->comm, c->stream));
        NCCL_CHECK(ncclRecv(c->rcnt.as<uint32_t>() + roff[s2], roff[s2 + 1] - roff[s2], ncclUint32, s2, c->comm, c->stream));
    }
    NCCL_CHECK(ncclGroupEnd());

    // owner: sort the received partial counts by key, sum per key, histogram
    uint32_t G = 0;
    CUDA_CHECK(cudaMemsetAsync(hist_d, 0, (OCC_BINS + 8) * 8, c->stream));
    if (R) {
        cub::DeviceRadixSort::SortPairs(nullptr, tb, c->rkey.as<uint64_t>(), c->skey.as<uint64_t>(), c->rcnt.as<uint32_t>(),
                                        c->scnt.as<uint32_t>(), (int64_t)R, 0, 2 * o.k, c->stream);
        cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, c->rkey.as<uint64_t>(), c->skey.as<uint64_t>(), c->rcnt.as<uint32_t>(),
                                        c->scnt.as<uint32_t>(), (int64_t)R, 0, 2 * o.k, c->stream);
        c->flags.reserve(R + 1);
        c->pstart.reserve((R + 2) * 4);      // NB: the local run starts are no longer needed
        run_flag_kernel<<<(unsigned)((R + 255) / 256), 256, 0, c->stream>>>(c->skey.as<uint64_t>(), R, 0, c->flags.as<uint8_t>());
        cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(), c->pstart.as<uint32_t>(),
                                   c->npairs.as<uint32_t>(), (int64_t)R, c->stream);
        cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(),
                                   c->pstart.as<uint32_t>(), c->npairs.as<uint32_t>(), (int64_t)R, c->stream);
        c->launches += 8;
        CUDA_CHECK(cudaMemcpyAsync(&G, c->npairs.p, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        const uint32_t sentinel = (uint32_t)R;
        CUDA_CHECK(cudaMemcpyAsync(c->pstart.as<uint32_t>() + G, &sentinel, 4, cudaMemcpyHostToDevice, c->stream));
        c->gcnt.reserve(((size_t)G + 1) * 4);
        owner_sum_kernel<<<148 * 4, 256, 0, c->stream>>>(c->skey.as<uint64_t>(), c->scnt.as<uint32_t>(), c->pstart.as<uint32_t>(), G,
                                                       c->gcnt.as<uint32_t>(), hist_d);
        ++c->launches;
    }
    const unsigned long long extra[2] = {G, M};
    CUDA_CHECK(cudaMemcpyAsync(hist_d + OCC_BINS, extra, 16, cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(ncclAllReduce(hist_d, hist_d, OCC_BINS + 2, ncclUint64, ncclSum, c->comm, c->stream));
    std::vector<unsigned long long> h(OCC_BINS + 2);
    CUDA_CHECK(cudaMemcpyAsync(h.data(), hist_d, (OCC_BINS + 2) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const uint64_t n_keys = h[OCC_BINS];
    if (S) S->n_target_minimizers = h[OCC_BINS + 1];
    int mid_occ;
    if (n_keys == 0) mid_occ = INT32_MAX;
    else {
        const uint64_t rank = (uint32_t)((1. - (double)o.mid_occ_frac) * (double)n_keys);
        uint64_t cum = 0; int64_t val = -1;
        for (int b = 0; b < OCC_BINS - 1; ++b) { cum += h[b]; if (cum > rank) { val = b; break; } }
        if (val < 0) {
            // the rank lies among keys occurring >= OCC_BINS-1 times: gather those counts from all owners (padded)
            const uint64_t cap = h[OCC_BINS - 1];
            DevBuf mine, all;
            mine.reserve((cap + 1) * 4); all.reserve((cap + 1) * 4 * W);
            CUDA_CHECK(cudaMemsetAsync(mine.p, 0xff, (cap + 1) * 4, c->stream));
            CUDA_CHECK(cudaMemsetAsync(hist_d + OCC_BINS + 4, 0, 8, c->stream));
            if (G) owner_overflow_kernel<<<(G + 255) / 256, 256, 0, c->stream>>>(c->gcnt.as<uint32_t>(), G, mine.as<uint32_t>(), hist_d + OCC_BINS + 4, cap);
            NCCL_CHECK(ncclAllGather(mine.p, all.p, cap, ncclUint32, c->comm, c->stream));
            std::vector<uint32_t> ov(cap * W);
            CUDA_CHECK(cudaMemcpyAsync(ov.data(), all.p, cap * W * 4, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            std::sort(ov.begin(), ov.end());                       // padding (0xffffffff) sorts last
            val = ov[rank - cum];
        }
        mid_occ = (int)val + 1;
    }
    if (mid_occ < o.min_mid_occ) mid_occ = o.min_mid_occ;
    if (o.max_mid_occ > o.min_mid_occ && mid_occ > o.max_mid_occ) mid_occ = o.max_mid_occ;

    // blacklist: owned keys above mid_occ -> everyone
    uint64_t n_above = 0;
    for (int b = std::min(mid_occ + 1, OCC_BINS - 1); b < OCC_BINS; ++b) n_above += h[b];   // global upper bound on the blacklist length
    const uint64_t cap = n_above + 1;
    c->blist.reserve(cap * 8); c->blall.reserve(cap * 8 * W);
    CUDA_CHECK(cudaMemsetAsync(c->blist.p, 0xff, cap * 8, c->stream));
    CUDA_CHECK(cudaMemsetAsync(hist_d + OCC_BINS + 4, 0, 8, c->stream));
    if (G) { owner_select_kernel<<<(G + 255) / 256, 256, 0, c->stream>>>(c->skey.as<uint64_t>(), c->pstart.as<uint32_t>(), c->gcnt.as<uint32_t>(), G,
                                                                          (uint32_t)mid_occ, c->blist.as<uint64_t>(), hist_d + OCC_BINS + 4, cap); ++c->launches; }
    NCCL_CHECK(ncclAllGather(c->blist.p, c->blall.p, cap, ncclUint64, c->comm, c->stream));
    c->blmap.reserve((M / 32 + 2) * 4);
    CUDA_CHECK(cudaMemsetAsync(c->blmap.p, 0, (M / 32 + 2) * 4, c->stream));
    B.view.tx = c->tx.as<uint64_t>(); B.view.tk32 = c->tk32.as<uint32_t>(); B.view.ty = c->ty.as<uint64_t>(); B.view.boff = c->boff.as<uint32_t>();
    B.view.M = M; B.view.shift = shift; B.view.bl = nullptr;
    bl_mark_kernel<<<(unsigned)((cap * W + 255) / 256), 256, 0, c->stream>>>(c->blall.as<uint64_t>(), cap * W, B.view, c->blmap.as<uint32_t>());
    ++c->launches;
    CUDA_CHECK(cudaGetLastError());
    B.view.bl = c->blmap.as<uint32_t>();
    B.mid_occ = mid_occ; B.n_keys = n_keys;
    return B;
}

// ---- peer-memory link: collective helpers (every rank calls them at the same point) ----------------------------------
bool all_agree(lrge_ctx *c, bool mine) {                  // true iff every rank passes true (also a barrier)
    c->link.hbuf.reserve(4096);
    uint32_t v = mine ? 1u : 0u;
    CUDA_CHECK(cudaMemcpyAsync(c->link.hbuf.p, &v, 4, cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(ncclAllReduce(c->link.hbuf.p, c->link.hbuf.p, 1, ncclUint32, ncclMin, c->comm, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(&v, c->link.hbuf.p, 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return v != 0;
}

// all-gather the IPC handle of `ptr` and map every peer's allocation; out[d] = mapped pointer (out[me] = ptr).
// Returns the collective verdict; on false everything that was opened here has been closed again.
bool exchange_and_open(lrge_ctx *c, void *ptr, bool have, void **out) {
    const int W = c->world, me = c->rank;
    struct Rec { cudaIpcMemHandle_t h; unsigned long long ok; unsigned long long pad; };
    static_assert(sizeof(Rec) == 80, "record layout");
    Rec mine{};
    if (have && cudaIpcGetMemHandle(&mine.h, ptr) == cudaSuccess) mine.ok = 1; else cudaGetLastError();
    c->link.hbuf.reserve(4096 + sizeof(Rec) * (size_t)(W + 1));
    uint8_t *base = c->link.hbuf.as<uint8_t>() + 4096;
    CUDA_CHECK(cudaMemcpyAsync(base + sizeof(Rec) * W, &mine, sizeof(Rec), cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(ncclAllGather(base + sizeof(Rec) * W, base, sizeof(Rec), ncclUint8, c->comm, c->stream));
    std::vector<Rec> all(W);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), base, sizeof(Rec) * W, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    bool ok = true;
    for (int d = 0; d < W; ++d) ok = ok && all[d].ok;
    for (int d = 0; d < W; ++d) out[d] = nullptr;
    if (ok) {
        out[me] = ptr;
        for (int d = 0; d < W && ok; ++d) {
            if (d == me) continue;
            if (cudaIpcOpenMemHandle(&out[d], all[d].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); out[d] = nullptr; ok = false; }
        }
    }
    const bool agreed = all_agree(c, ok);
    if (!agreed) for (int d = 0; d < W; ++d) { if (d != me && out[d]) cudaIpcCloseMemHandle(out[d]); out[d] = nullptr; }
    return agreed;
}

void peer_link_close(lrge_ctx *c) {
    auto &L = c->link;
    for (int d = 0; d < 16; ++d) {
        if (d != c->rank && L.flags_peer[d]) cudaIpcCloseMemHandle(L.flags_peer[d]);
        if (d != c->rank && L.win_peer[d]) cudaIpcCloseMemHandle(L.win_peer[d]);
        L.flags_peer[d] = nullptr; L.win_peer[d] = nullptr;
    }
    if (L.flags) { cudaFree(L.flags); L.flags = nullptr; }
    L.win.release(); L.win_bytes = 0; L.ok = false;
}

// The same, fast path: ONE host synchronisation (the W x W matrix of send counts, which ncclSend/ncclRecv need on the
// host), hidden behind `while_waiting` (the caller launches the sketch of the streamed set there).  Packed
// (key << 24 | count) entries, one array per exchange; the owner sorts keys only; mid_occ is computed on the device from
// the all-reduced histogram; results (mid_occ, key / minimizer totals, fallback flags, all-reduced so that every rank
// decides alike) land in pinned memory and are read by build_index_sharded_finish after the caller's next
// synchronisation.  Flags set = rank in the overflow bin or blacklist above its buffer: the exact path above runs.
struct ShardedJob { BuiltIndex B; bool active = false; };

// bl_cap: blacklist entries per rank in the all-gather -- MUST be the same number on every rank (derive it from a
// quantity all ranks share, not from the local minimizer count).
template <typename F>
ShardedJob build_index_sharded_launch(lrge_ctx *c, const Sketch &sk, const MapOpt &o, uint64_t bl_cap, F while_waiting) {
    ShardedJob J;
    BuiltIndex &B = J.B;
    const int W = c->world, me = c->rank;
    const uint64_t M = sk.M;
    if (M >= (1ull << 32) - 2) throw std::length_error("more than 2^32 target minimizers per GPU are not supported");
    if (W > 16) throw std::length_error("more than 16 ranks are not supported");
    c->ty.reserve((M + 1) * 8);
    c->tk32.reserve((M + 64) * 4);
    int bits = 8;
    while (bits < 2 * o.k && bits < 28 && (8ull << bits) < M) ++bits;
    if (bits < 2 * o.k - 32) bits = 2 * o.k - 32;
    const uint32_t nb = 1u << bits;
    c->boff.reserve(((size_t)nb + 2) * 4);
    c->hist.reserve((OCC_BINS + 2) * 8);
    CUDA_CHECK(cudaMemsetAsync(c->hist.p, 0, (OCC_BINS + 2) * 8, c->stream));      // local histogram: unused here
    c->flags.reserve((M + 1) * 4);                                                  // run length at every run start, 0 elsewhere
    uint32_t *run_n = c->flags.as<uint32_t>();
    c->npairs.reserve(16);
    c->ukey.reserve((M + 1) * 8);
    CUDA_CHECK(cudaMemsetAsync(c->npairs.p, 0, 16, c->stream));
    const bool k32 = !o.is_hpc && 2 * o.k <= 32;               // as in build_index_launch: sort (32-bit hash, y)
    B.x32 = k32;
    size_t tb = 0;
    int shift;
    if (k32) {
        c->tx.reserve((M + 1) * 4);                                                 // unsorted 32-bit keys
        if (M) {
            KLAUNCH(c, "key32_kernel", 12.0 * M,
                    key32_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c->stream>>>(sk.x.as<uint64_t>(), M, c->tx.as<uint32_t>()));
            cub::DeviceRadixSort::SortPairs(nullptr, tb, c->tx.as<uint32_t>(), c->tk32.as<uint32_t>(), sk.y.as<uint64_t>(), c->ty.as<uint64_t>(),
                                            (int64_t)M, 0, 2 * o.k, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(index: u32 key, u64 value)", 32.0 * M,
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, c->tx.as<uint32_t>(), c->tk32.as<uint32_t>(), sk.y.as<uint64_t>(),
                                                    c->ty.as<uint64_t>(), (int64_t)M, 0, 2 * o.k, c->stream));
            c->launches += (2 * o.k + 7) / 8;
            c->prof.set_passes("cub::DeviceRadixSort(index: u32 key, u64 value)", (2 * o.k + 7) / 8);
        }
        CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));
        shift = 8 + 2 * o.k - bits;
        KLAUNCH(c, "index_post_kernel", 8.0 * M + 4.0 * nb,
                (index_post_kernel<uint32_t, 0><<<148 * 8, 256, 0, c->stream>>>(c->tk32.as<uint32_t>(), M, shift - 8, nb, c->boff.as<uint32_t>(), nullptr,
                                                                                c->hist.as<unsigned long long>(), run_n)));
        if (M) {
            PackRunOp<uint32_t, 0> op{c->tk32.as<uint32_t>(), run_n};
            cub::TransformInputIterator<uint64_t, PackRunOp<uint32_t, 0>, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), op);
            cub::DeviceSelect::Flagged(nullptr, tb, it, run_n, c->ukey.as<uint64_t>(), c->npairs.as<uint32_t>(), (int64_t)M, c->stream);
            KLAUNCH(c, "cub::DeviceSelect(packed runs)", 8.0 * M,
                    cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, it, run_n, c->ukey.as<uint64_t>(), c->npairs.as<uint32_t>(), (int64_t)M, c->stream));
        }
    } else {
        c->tx.reserve((M + 1) * 8);
        if (M) {
            cub::DeviceRadixSort::SortPairs(nullptr, tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                            c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(index: u64 key, u64 value)", 32.0 * M,
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, sk.x.as<uint64_t>(), c->tx.as<uint64_t>(), sk.y.as<uint64_t>(),
                                                    c->ty.as<uint64_t>(), (int64_t)M, 8, 8 + 2 * o.k, c->stream));
            c->launches += (2 * o.k + 7) / 8;
            c->prof.set_passes("cub::DeviceRadixSort(index: u64 key, u64 value)", (2 * o.k + 7) / 8);
        }
        CUDA_CHECK(cudaMemsetAsync(c->tk32.as<uint32_t>() + M, 0, 64 * 4, c->stream));
        shift = 8 + 2 * o.k - bits;
        KLAUNCH(c, "index_post_kernel", 16.0 * M + 4.0 * nb,
                (index_post_kernel<uint64_t, 8><<<148 * 8, 256, 0, c->stream>>>(c->tx.as<uint64_t>(), M, shift, nb, c->boff.as<uint32_t>(), c->tk32.as<uint32_t>(),
                                                                                c->hist.as<unsigned long long>(), run_n)));
        if (M) {
            PackRunOp<uint64_t, 8> op{c->tx.as<uint64_t>(), run_n};
            cub::TransformInputIterator<uint64_t, PackRunOp<uint64_t, 8>, cub::CountingInputIterator<uint32_t>> it(cub::CountingInputIterator<uint32_t>(0), op);
            cub::DeviceSelect::Flagged(nullptr, tb, it, run_n, c->ukey.as<uint64_t>(), c->npairs.as<uint32_t>(), (int64_t)M, c->stream);
            KLAUNCH(c, "cub::DeviceSelect(packed runs)", 12.0 * M,
                    cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, it, run_n, c->ukey.as<uint64_t>(), c->npairs.as<uint32_t>(), (int64_t)M, c->stream));
        }
    }
    // xmeta: bounds[W+1] | sendc[W] | mat[W*W] | hist[OCC_BINS+2] | meta[8]
    c->xmeta.reserve(((size_t)(2 * W + 1) + (size_t)W * W + OCC_BINS + 2 + 8) * 8 + 64);
    unsigned long long *bounds_d = c->xmeta.as<unsigned long long>();
    unsigned long long *sendc_d = bounds_d + (W + 1);
    unsigned long long *mat_d = sendc_d + W;
    unsigned long long *hist_d = mat_d + (size_t)W * W;
    unsigned long long *meta_d = hist_d + OCC_BINS + 2;          // [0] mid_occ [1] need exact [2] blacklist overflow [3] blacklist length
    owner_bounds_dev_kernel<<<1, 64, 0, c->stream>>>(c->ukey.as<uint64_t>(), c->npairs.as<uint32_t>(), W, 2 * o.k, bounds_d, sendc_d);
    ++c->launches;
    NCCL_CHECK(ncclAllGather(sendc_d, mat_d, W, ncclUint64, c->comm, c->stream));
    unsigned long long *pin64 = reinterpret_cast<unsigned long long *>(c->pin2);   // bounds[W+1] then mat[W*W]
    CUDA_CHECK(cudaMemcpyAsync(pin64, bounds_d, (size_t)(W + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(pin64 + W + 1, mat_d, (size_t)W * W * 8, cudaMemcpyDeviceToHost, c->stream));
    while_waiting();                                               // e.g. launch the sketch of the streamed set
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<uint64_t> bounds(pin64, pin64 + W + 1), roff(W + 1, 0);
    const unsigned long long *mat = pin64 + W + 1;
    for (int s2 = 0; s2 < W; ++s2) roff[s2 + 1] = roff[s2] + mat[(size_t)s2 * W + me];
    const uint64_t R = roff[W];
    if (R >= (1ull << 32) - 2) throw std::length_error("too many distinct keys per owner rank");
    uint64_t Rmax = 0;                                            // the largest receive total of any rank (every rank knows the matrix)
    for (int d = 0; d < W; ++d) { uint64_t t = 0; for (int s2 = 0; s2 < W; ++s2) t += mat[(size_t)s2 * W + d]; Rmax = std::max(Rmax, t); }
    const uint64_t *recv_buf = nullptr;
    if (c->link.ok && (Rmax + 1) * 8 > c->link.win_bytes) {
        // (re)establish the receive windows -- the same decision on every rank; warm-up calls only.  Nobody frees a
        // window before everybody has unmapped it.
        auto &L = c->link;
        for (int d = 0; d < W; ++d) { if (d != me && L.win_peer[d]) cudaIpcCloseMemHandle(L.win_peer[d]); L.win_peer[d] = nullptr; }
        all_agree(c, true);
        const size_t nb = (Rmax + 1) * 8 + (Rmax + 1) * 2 + 4096;
        bool have = true;
        L.win.release();
        try { L.win.reserve(nb); } catch (const std::exception &) { have = false; }
        void *mapped[16] = {};
        if (exchange_and_open(c, L.win.p, have, mapped)) { for (int d = 0; d < W; ++d) L.win_peer[d] = (uint64_t *)mapped[d]; L.win_bytes = nb; }
        else { L.ok = false; L.win_bytes = 0; }                   // every rank falls back to NCCL send / recv alike
    }
    if (c->link.ok) {
        auto &L = c->link;
        ++L.epoch;
        L.done.reserve(256);
        PeerPushArgs PA{};
        PA.src = c->ukey.as<uint64_t>(); PA.done = L.done.as<unsigned int>(); PA.epoch = L.epoch; PA.W = W; PA.me = me;
        PA.ctas_per_dest = (uint32_t)std::max(1, 148 * 4 / W);
        for (int d = 0; d < W; ++d) {
            PA.dst[d] = d == me ? L.win.as<uint64_t>() : L.win_peer[d];
            PA.flag[d] = d == me ? nullptr : L.flags_peer[d] + me;
            PA.sbeg[d] = bounds[d];
            uint64_t off = 0;
            for (int s2 = 0; s2 < me; ++s2) off += mat[(size_t)s2 * W + d];
            PA.doff[d] = off;
        }
        PA.sbeg[W] = bounds[W];
        KLAUNCH(c, "peer_push_kernel", 8.0 * R, peer_push_kernel<<<PA.ctas_per_dest * W, 256, 0, c->stream>>>(PA));
        KLAUNCH(c, "peer_wait_kernel", 8.0 * W, peer_wait_kernel<<<1, 32, 0, c->stream>>>(L.flags, W, me, L.epoch, L.done.as<unsigned int>() + 32));
        recv_buf = L.win.as<uint64_t>();
    } else {
        c->rkey.reserve((R + 1) * 8);
        c->prof.kbegin("ncclSend/Recv(key ranges)", 8.0 * R);
        NCCL_CHECK(ncclGroupStart());
        for (int d = 0; d < W; ++d) {
            const uint64_t sc = bounds[d + 1] - bounds[d];
            if (sc) NCCL_CHECK(ncclSend(c->ukey.as<uint64_t>() + bounds[d], sc, ncclUint64, d, c->comm, c->stream));
        }
        for (int s2 = 0; s2 < W; ++s2)
            if (roff[s2 + 1] > roff[s2]) NCCL_CHECK(ncclRecv(c->rkey.as<uint64_t>() + roff[s2], roff[s2 + 1] - roff[s2], ncclUint64, s2, c->comm, c->stream));
        NCCL_CHECK(ncclGroupEnd());
        c->prof.kend();
        recv_buf = c->rkey.as<uint64_t>();
    }
    // owner: sum the <= W partial counts of every key (owner_merge_kernel: slices of the key range through a
    // shared-memory table), histogram of the sums
    CUDA_CHECK(cudaMemsetAsync(hist_d, 0, (OCC_BINS + 2 + 8) * 8, c->stream));
    OwnerMergeArgs OA{};
    OA.rk = recv_buf; OA.W = W;
    for (int s2 = 0; s2 <= W; ++s2) OA.roff[s2] = roff[s2];
    const int kbits = 2 * o.k;
    OA.klo = (((uint64_t)me << kbits) + W - 1) / (uint64_t)W;
    OA.khi = me + 1 == W ? (1ull << kbits) : (((uint64_t)(me + 1) << kbits) + W - 1) / (uint64_t)W;
    const uint64_t n_slices = std::max<uint64_t>(1, std::min<uint64_t>(R / (OM_CAP * 3 / 4) + 1, OA.khi - OA.klo));   // ~6k entries per slice (table: OM_CAP)
    OA.slice_w = (OA.khi - OA.klo + n_slices - 1) / n_slices;
    OA.hist = hist_d; OA.meta = meta_d;
    const size_t om_smem = 0;
    const unsigned om_grid = (unsigned)((OA.khi - OA.klo + OA.slice_w - 1) / OA.slice_w);
    OA.n_slices = om_grid;
    c->ucnt.reserve(((size_t)om_grid + 1) * W * 4);
    OA.cuts = c->ucnt.as<uint32_t>();
    if (R) KLAUNCH(c, "owner_cuts_kernel", 4.0 * (om_grid + 1) * W,
                   owner_cuts_kernel<<<(unsigned)(((uint64_t)(om_grid + 1) * W + 255) / 256), 256, 0, c->stream>>>(OA, c->ucnt.as<uint32_t>()));
    c->skey.reserve((R + 1) * 8);                                   // candidates: summed entries above min_mid_occ (<= mid_occ)
    OA.cand = c->skey.as<uint64_t>(); OA.cand_n = meta_d + 4; OA.cand_thr = (uint64_t)std::max(o.min_mid_occ, 0);
    if (R) KLAUNCH(c, "owner_merge_kernel<hist>", 8.0 * R, owner_merge_kernel<0><<<om_grid, OM_THREADS, om_smem, c->stream>>>(OA));
    const unsigned long long extra = M;
    CUDA_CHECK(cudaMemcpyAsync(hist_d + OCC_BINS + 1, &extra, 8, cudaMemcpyHostToDevice, c->stream));
    NCCL_CHECK(ncclAllReduce(hist_d, hist_d, OCC_BINS + 2, ncclUint64, ncclSum, c->comm, c->stream));
    mid_occ_kernel<<<1, 32, 0, c->stream>>>(hist_d, o.mid_occ_frac, o.min_mid_occ, o.max_mid_occ, meta_d);
    ++c->launches;
    // blacklist: owned keys above mid_occ -> everyone -> bitmap over the local runs
    const uint64_t cap_bl = bl_cap;
    c->blist.reserve(cap_bl * 8); c->blall.reserve(cap_bl * 8 * W);
    CUDA_CHECK(cudaMemsetAsync(c->blist.p, 0xff, cap_bl * 8, c->stream));
    OA.bl = c->blist.as<uint64_t>(); OA.bl_n = meta_d + 3; OA.bl_cap = cap_bl; OA.bl_ovf = meta_d + 2;
    if (R) KLAUNCH(c, "bl_from_cand_kernel", 8.0 * bl_cap,
                   bl_from_cand_kernel<<<148 * 4, 256, 0, c->stream>>>(OA.cand, OA.cand_n, meta_d, OA.bl, OA.bl_n, OA.bl_cap, OA.bl_ovf));
    NCCL_CHECK(ncclAllGather(c->blist.p, c->blall.p, cap_bl, ncclUint64, c->comm, c->stream));
    NCCL_CHECK(ncclAllReduce(meta_d + 1, meta_d + 1, 2, ncclUint64, ncclMax, c->comm, c->stream));     // fallback flags: every rank alike
    c->blmap.reserve((M / 32 + 2) * 4);
    CUDA_CHECK(cudaMemsetAsync(c->blmap.p, 0, (M / 32 + 2) * 4, c->stream));
    B.view.tx = k32 ? nullptr : c->tx.as<uint64_t>(); B.view.tk32 = c->tk32.as<uint32_t>(); B.view.ty = c->ty.as<uint64_t>(); B.view.boff = c->boff.as<uint32_t>();
    B.view.M = M; B.view.shift = shift; B.view.bl = nullptr;
    KLAUNCH(c, "bl_mark_kernel", 8.0 * cap_bl * W,
            bl_mark_kernel<<<(unsigned)((cap_bl * W + 255) / 256), 256, 0, c->stream>>>(c->blall.as<uint64_t>(), cap_bl * W, B.view, c->blmap.as<uint32_t>()));
    CUDA_CHECK(cudaGetLastError());
    B.view.bl = c->blmap.as<uint32_t>();
    // results for build_index_sharded_finish: meta[0..2], then n_keys and the minimizer total of all shards
    CUDA_CHECK(cudaMemcpyAsync(pin64, meta_d, 24, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(pin64 + 3, hist_d + OCC_BINS, 16, cudaMemcpyDeviceToHost, c->stream));
    pin64[5] = 0;
    if (c->link.ok) CUDA_CHECK(cudaMemcpyAsync(pin64 + 5, c->link.done.as<unsigned int>() + 32, 4, cudaMemcpyDeviceToHost, c->stream));
    B.mid_occ = -1; B.n_keys = 0;
    J.active = true;
    return J;
}

// after a synchronisation of the stream: false = one of the fallback conditions held on some rank (run the exact path)
bool build_index_sharded_finish(lrge_ctx *c, ShardedJob *J, lrge_stats *S) {
    const unsigned long long *r = reinterpret_cast<const unsigned long long *>(c->pin2);
    if (r[5]) throw std::runtime_error("key exchange over peer memory timed out (a peer rank did not deliver its slice)");
    if (r[1] || r[2]) return false;
    J->B.mid_occ = (int)r[0];
    J->B.n_keys = r[3];
    if (S) S->n_target_minimizers = r[4];
    return true;
}

// lexicographic name ranks over the union of two sets (equal names share a rank)
struct NameRanks { std::vector<uint32_t> a, b; bool a_unique = true, b_unique = true, shared = false; };
NameRanks rank_names(const lrge_staged *A, const lrge_staged *B) {
    struct Ref { const char *s; uint32_t len, set, idx; };
    std::vector<Ref> v;
    v.reserve((size_t)A->n + (B ? B->n : 0));
    for (uint32_t i = 0; i < A->n; ++i) v.push_back({A->names.data() + A->name_off[i], A->name_off[i + 1] - A->name_off[i], 0, i});
    if (B) for (uint32_t i = 0; i < B->n; ++i) v.push_back({B->names.data() + B->name_off[i], B->name_off[i + 1] - B->name_off[i], 1, i});
    auto cmp = [](const Ref &x, const Ref &y) {
        const uint32_t m = std::min(x.len, y.len);
        const int c = m ? memcmp(x.s, y.s, m) : 0;
        return c ? c : (x.len < y.len ? -1 : x.len > y.len ? 1 : 0);
    };
    std::sort(v.begin(), v.end(), [&](const Ref &x, const Ref &y) { return cmp(x, y) < 0; });
    NameRanks R;
    R.a.resize(A->n);
    if (B) R.b.resize(B->n);
    uint32_t rank = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (i > 0 && cmp(v[i - 1], v[i]) != 0) ++rank;
        else if (i > 0) {
            if (v[i - 1].set == v[i].set) (v[i].set ? R.b_unique : R.a_unique) = false;
            else R.shared = true;
        }
        (v[i].set ? R.b : R.a)[v[i].idx] = rank;
    }
    return R;
}

// ---- PAF annotations of the overlap records (host side; records mode only) -------------------------------------------
// mm_est_err (esterr.c) -> dv, tbuf->rep_len (seed.c mm_collect_matches) -> rl, and the row order of mm_gen_regs (hit.c:
// regions sorted by (score << 32 | cnt) ^ hash, descending).  All three are functions of the query's sketch and of the
// end anchors of each chain, which the chaining kernel returns with the record.
namespace paf {

inline uint64_t hash64(uint64_t key) {
    key = (~key + (key << 21)); key = key ^ key >> 24; key = ((key + (key << 3)) + (key << 8)); key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)); key = key ^ key >> 28; key = (key + (key << 31));
    return key;
}
inline uint32_t wang_hash(uint32_t key) {
    key += ~(key << 15); key ^= (key >> 10); key += (key << 3); key ^= (key >> 6); key += ~(key << 11); key ^= (key >> 16);
    return key;
}
inline uint32_t x31_hash(const char *s, uint32_t len) {
    if (len == 0) return 0;
    uint32_t h = (uint32_t)(unsigned char)s[0];
    for (uint32_t i = 1; i < len; ++i) h = (h << 5) - h + (uint32_t)(unsigned char)s[i];
    return h;
}

struct QuerySeeds {                       // the seeds of one query as mm_collect_matches leaves them
    std::vector<int32_t> pos;             // mini_pos: q_pos >> 1 of every kept seed (has hits, not repetitive), ascending
    std::vector<uint8_t> tandem;          // is_tandem of those seeds
    int32_t rep_len = 0;
    float avg_k = 0.0f;
};

QuerySeeds query_seeds(const uint64_t *x, const uint64_t *y, const uint8_t *flt, const uint32_t *hstart, const uint32_t *hn,
                       uint32_t m0, uint32_t m1) {
    QuerySeeds S;
    std::vector<uint32_t> mv;             // minimizers that survive mm_seed_mz_flt
    for (uint32_t i = m0; i < m1; ++i) if (!flt[i]) mv.push_back(i);
    int rep_st = 0, rep_en = 0;
    uint64_t sum_k = 0;
    for (size_t j = 0; j < mv.size(); ++j) {
        const uint32_t i = mv[j];
        const bool repetitive = hstart[i] == SEED_REPETITIVE;
        if (hn[i] == 0 && !repetitive) continue;                          // not in the index
        const uint32_t q_pos = (uint32_t)y[i], q_span = (uint32_t)(x[i] & 0xff);
        if (repetitive) {
            const int en = (int)(q_pos >> 1) + 1, st = en - (int)q_span;
            if (st > rep_en) { S.rep_len += rep_en - rep_st; rep_st = st; rep_en = en; }
            else rep_en = en;
        } else {
            bool t = false;
            if (j > 0 && (x[i] >> 8) == (x[mv[j - 1]] >> 8)) t = true;
            if (j + 1 < mv.size() && (x[i] >> 8) == (x[mv[j + 1]] >> 8)) t = true;
            S.pos.push_back((int32_t)(q_pos >> 1));
            S.tandem.push_back(t);
            sum_k += q_span;
        }
    }
    S.rep_len += rep_en - rep_st;
    if (!S.pos.empty()) S.avg_k = (float)sum_k / (int)S.pos.size();
    return S;
}

}  // namespace paf

enum Mode { MODE_TWOSET = 0, MODE_AVA = 1, MODE_INVERSE = 2 };

struct WidenU32 { __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; } };

// The whole hot path: `idx` is the indexed read set, `str` the streamed one (the same object for AVA).
// counts_out: MODE_TWOSET -> per streamed read; MODE_AVA -> per read; MODE_INVERSE -> per indexed read.
struct RegsRequest { lrge_overlap *out = nullptr; uint64_t cap = 0; uint64_t *n_out = nullptr; };

void run_overlap(lrge_ctx *c, lrge_staged *idx, lrge_staged *str, Mode mode, const lrge_opts &opts,
                 uint32_t *counts_out, uint8_t *had_out, lrge_stats *st, const RegsRequest *rq = nullptr) {
    const MapOpt o = preset_for(opts.platform);
    const uint32_t n_counts = mode == MODE_TWOSET ? str->n : idx->n;
    // Aligner::map fails on an empty sequence (aligner.rs:214-216 -> LrgeError::MapError); minimap2's index build just
    // skips empty target sequences, so only the streamed set is checked
    if (str->has_empty) throw std::invalid_argument("MapError: Sequence is empty");
    c->prof.reset();
    c->prof.sbegin(ST_TOTAL);
    lrge_stats S{};
    S.n_target_bases = idx->total_bases; S.n_query_bases = str->total_bases;
    const uint32_t launches0 = c->launches;

    // K0/K1 on the indexed set first: its kernels run while the host analyses the read names below.  With a
    // communicator the index itself is sharded: this rank sketches only its block of targets (shard_range) and the
    // occurrence filter is made exact by build_index_sharded.
    const int world = std::max(opts.shard_count, 1);
    const bool sharded_index = c->comm != nullptr && world > 1;
    if (sharded_index && (world != c->world || opts.shard_index != c->rank))
        throw std::runtime_error("shard_index/shard_count do not match the communicator");
    uint32_t i0 = 0, i1 = idx->n;
    if (sharded_index && !c->local_block) shard_range(idx->n, c->rank, world, &i0, &i1);
    Sketch2Job job_idx;
    BuiltIndex Bh{};
    c->fork_recorded = false;
    // host buffers, one GPU, ava-ont: sketch + index sort per half of the staged chunks (the name analysis runs inside)
    const bool halves = !sharded_index && sketch_index_halves(c, idx, o, &c->sk_idx, [&]() { finish_staging(idx); finish_staging(str); }, &Bh);
    job_idx.want_k32 = !sharded_index && !o.is_hpc && 2 * o.k <= 32;      // the compaction also writes the index's sort keys
    const bool fast_idx = !halves && sketch2_begin(c, idx, o, &c->sk_idx, i0, i1 - i0, &job_idx);

    // names -> ranks.  Only equality (two-set: NO_DIAG, HashSet<name>) or strcmp order (AVA: NO_DUAL) is ever used.
    finish_staging(idx);
    finish_staging(str);
    struct Drain { lrge_ctx *c; bool armed; ~Drain() { if (armed) cudaStreamSynchronize(c->stream); } } drain{c, true};   // D2H targets live in ctx
    if ((mode == MODE_AVA || mode == MODE_INVERSE) && !idx->names_unique) throw std::domain_error("DuplicateReadIdentifier");
    NameRanks R;
    if (mode == MODE_AVA) {
        if (idx->lex_rank.empty() && idx->n) idx->lex_rank = rank_names(idx, nullptr).a;
        R.a = idx->lex_rank;
    } else {
        // do the two sets share a name?  (sorted hash lists, equal hashes verified on the bytes)
        bool shared = false;
        size_t i = 0, j = 0;
        while (i < idx->name_hash.size() && j < str->name_hash.size() && !shared) {
            const uint64_t ha = idx->name_hash[i].first, hb = str->name_hash[j].first;
            if (ha < hb) ++i;
            else if (hb < ha) ++j;
            else {
                size_t j2 = j;
                for (; j2 < str->name_hash.size() && str->name_hash[j2].first == ha && !shared; ++j2) {
                    const uint32_t a = idx->name_hash[i].second, b = str->name_hash[j2].second;
                    const uint32_t la = idx->name_off[a + 1] - idx->name_off[a], lb = str->name_off[b + 1] - str->name_off[b];
                    if (la == lb && memcmp(idx->names.data() + idx->name_off[a], str->names.data() + str->name_off[b], la) == 0) shared = true;
                }
                ++i;
            }
        }
        if (!shared && idx->names_unique && str->names_unique) {   // all names distinct: any injective numbering will do
            R.a.resize(idx->n); R.b.resize(str->n);
            std::iota(R.a.begin(), R.a.end(), 0u);
            std::iota(R.b.begin(), R.b.end(), idx->n);
        } else R = rank_names(idx, str);
    }
    const std::vector<uint32_t> &rank_idx = R.a, &rank_str = mode == MODE_AVA ? R.a : R.b;
    drain.armed = false;

    bool idx_k32 = false;
    if (!halves) {
        if (fast_idx && sketch2_end(c, job_idx)) idx_k32 = job_idx.have_k32;
        else sketch_readset(c, idx, o, &c->sk_idx, i0, i1 - i0, true);
    }
    // K2/K3, then K1 on the streamed set.  Single GPU: the sketch's synchronisation also delivers the occurrence histogram
    // of the index (build_index_finish).  Sharded index: the sketch is launched while the host waits for the send counts
    // of the key exchange, and its synchronisation delivers mid_occ and the fallback flags.
    BuiltIndex B;
    Sketch *sq = &c->sk_idx;
    const bool need_str = mode != MODE_AVA || sharded_index;
    // The sketch of the streamed set goes to the side stream: it only has to wait for the sketch of the indexed set
    // (shared staging buffers), and then runs next to the index build (sort passes are HBM-bound, the sketch issue-bound).
    Sketch2Job job_str;
    bool fast_str = false;
    if (!c->fork_recorded) CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));   // the indexed set's sketch is complete here
    auto fork_str = [&]() {
        // per-kernel profiling (lrge_gpu_set_profiling) keeps everything on one stream: spans of kernels that share the
        // SMs with another stream's work would not be comparable with the serialised launch list of a profiler
        if (!need_str || c->no_side || c->prof.kernels) return;
        SideStream side(c);
        CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_fork, 0));
        fast_str = sketch2_begin(c, str, o, &c->sk_str, 0, str->n, &job_str);
    };
    // small tables of the seed stage (name ranks, read lengths, gap penalties): uploaded on the side stream behind the
    // streamed set's sketch when there is one, so that they are not on the path between the index and the first lookup
    bool seed_tables_up = false;
    auto upload_seed_tables = [&]() {
        h2d(c, c->rank_idx, rank_idx.data(), idx->n);
        h2d(c, c->rank_str, rank_str.data(), str->n);
        h2d(c, c->len_idx, idx->h_len.data(), idx->n);
        h2d(c, c->len_str, str->h_len.data(), str->n);
        if (c->pen_k != o.k || c->pen_bw != o.bw) { gap_penalty_table(o.k, o.chain_gap_scale, o.bw + 1, c->pen_host); c->pen_k = o.k; c->pen_bw = o.bw; }
        h2d(c, c->pen, c->pen_host.data(), c->pen_host.size());
        seed_tables_up = true;
    };
    auto finish_str = [&]() {                      // minimizer total of the streamed set -> compaction; join the main stream
        bool ok = false;
        if (fast_str) {
            SideStream side(c);
            ok = sketch2_end(c, job_str);          // synchronises the side stream only
            if (ok) { upload_seed_tables(); CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream)); }
        }
        if (ok) CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
        else sketch_readset(c, str, o, &c->sk_str, 0, 0xffffffffu, fast_str);   // not applicable, or a sub-tile overflowed
        sq = &c->sk_str;
    };
    if (sharded_index) {
        c->prof.sbegin(ST_INDEX);
        // the blacklist holds ~2e-4 of the distinct keys; room for 1/512 of a rank's nominal share of the bases
        // (a number every rank derives alike: with local blocks only the streamed set is common to all ranks)
        const uint64_t bl_cap = c->local_block ? std::max<uint64_t>(4096, str->total_bases / 256)
                                               : std::max<uint64_t>(4096, idx->total_bases / (uint64_t)world / 512);
        fork_str();
        ShardedJob J = build_index_sharded_launch(c, c->sk_idx, o, bl_cap, [&]() {});
        c->prof.send(ST_INDEX);
        finish_str();
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (build_index_sharded_finish(c, &J, &S)) B = J.B;
        else { c->prof.sbegin(ST_INDEX); B = build_index_sharded_exact(c, c->sk_idx, o, &S); c->prof.send(ST_INDEX); }
    } else {
        if (halves) B = Bh;                        // already launched, behind the second half's sort
        else {
            c->prof.sbegin(ST_INDEX);
            B = build_index_launch(c, c->sk_idx, o, idx_k32);
            c->prof.send(ST_INDEX);
        }
        fork_str();                                // after the index launches: those are on the critical path
        S.n_target_minimizers = c->sk_idx.M;
        S.n_index_runs = halves ? 2 : 1;
        if (need_str) finish_str();
        CUDA_CHECK(cudaStreamSynchronize(c->stream));   // the occurrence histogram of the index
        build_index_finish(c, o, &B);
    }
    S.n_keys = B.n_keys; S.mid_occ = B.mid_occ;
    const uint64_t MQ = sq->M;
    S.n_query_minimizers = MQ;

    c->prof.sbegin(ST_SEED);
    if (!seed_tables_up) upload_seed_tables();

    SkipCtx K{};
    K.rank_q = c->rank_str.as<uint32_t>(); K.rank_t = c->rank_idx.as<uint32_t>();
    K.len_q = c->len_str.as<uint32_t>(); K.len_t = c->len_idx.as<uint32_t>();
    K.no_dual = mode == MODE_AVA;                                      // aligner.rs:93-103, ava.rs:378
    K.check_names = mode == MODE_AVA || R.shared;                      // NO_DIAG only matters for equal names
    shard_range(idx->n, opts.shard_index, world, &K.shard_lo, &K.shard_hi);
    K.sharded = world > 1 && !sharded_index;                          // sharded index: every local hit already belongs to this shard

    c->counts.reserve(((size_t)n_counts + 1) * 4);
    c->had.reserve((size_t)str->n + 1);
    c->counters.reserve(128);
    CUDA_CHECK(cudaMemsetAsync(c->counts.p, 0, ((size_t)n_counts + 1) * 4, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->had.p, 0, (size_t)str->n + 1, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->counters.p, 0, 128, c->stream));
    // [0] seeds [1] hits [2] overlap records [3..6] fused kernel: anchors, groups, rounds, capacity violations [8,9] scratch
    unsigned long long *ctr = c->counters.as<unsigned long long>();
    const bool want_regs = rq && rq->n_out;
    if (want_regs) c->regs.reserve(std::max<uint64_t>(rq->cap, 1) * sizeof(RegOut));

    const int PB = bits_for(idx->max_len ? idx->max_len - 1 : 0), RB = bits_for(idx->n ? idx->n - 1 : 0);

    ChainParams P{};
    P.max_dist_x = std::max(o.max_gap, o.bw); P.max_dist_y = std::max(o.max_gap, o.bw);
    P.bw = o.bw; P.max_skip = o.max_chain_skip; P.max_iter = o.max_chain_iter; P.min_cnt = o.min_cnt; P.min_sc = o.min_chain_score;
    P.pen = c->pen.as<int32_t>();
    P.span_bound_ok = !o.is_hpc && o.k * (o.min_cnt - 1) < o.min_chain_score && o.bw >= o.min_chain_score;

    const bool host_dedupe = mode == MODE_TWOSET && !R.a_unique;      // duplicate target names: HashSet<name> semantics
    std::vector<std::vector<uint32_t>> dedupe_sets;
    if (host_dedupe) dedupe_sets.resize(str->n);
    c->prof.send(ST_SEED);

    // ---- K4-K6 --------------------------------------------------------------------------------------------------------
    // seed stage for all queries (mm_seed_mz_flt, index lookup, anchors per minimizer and their prefix sums), then per
    // batch of queries bounded by anchor memory: anchor_fill -> sort + group -> chain.
    //   sort + group = sort_query_kernel, the per-query shared-memory sort (kernels_map.cuh), by default;
    //                  the device-wide radix sort + group_mark / select / classify with LRGE_FORCE_GLOBAL=1 and for the
    //                  queries the per-query kernel defers.
    const bool use_smem_sort = !c->force_global;
    std::vector<uint64_t> qa((size_t)str->n + 1, 0);
    std::vector<uint32_t> moff_h((size_t)str->n + 1, 0);
    c->prof.sbegin(ST_SEED);
    c->flt.reserve(MQ + 1);
    c->hstart.reserve((MQ + 8) * 4); c->hn.reserve((MQ + 8) * 4); c->acnt.reserve((MQ + 8) * 4);
    c->aoff.reserve((MQ + 2) * 8);
    c->qaoff.reserve(((size_t)str->n + 2) * 8);
    CUDA_CHECK(cudaMemsetAsync(c->flt.p, 0, MQ + 1, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->acnt.p, 0, (MQ + 8) * 4, c->stream));
    CUDA_CHECK(cudaMemsetAsync(c->qaoff.p, 0, ((size_t)str->n + 2) * 8, c->stream));
    if (MQ) {
        KLAUNCH(c, "mz_flt_kernel", 16.0 * MQ,
                mz_flt_kernel<<<str->n, 256, 0, c->stream>>>(sq->x.as<uint64_t>(), sq->moff.as<uint32_t>(), B.mid_occ, o.q_occ_frac, c->flt.as<uint8_t>()));
        if (K.check_names || K.sharded) {            // skip_seed / shard membership is decided per hit: walk the lists
            KLAUNCH(c, "seed_count_kernel", 28.0 * MQ,
                    seed_count_kernel<<<(unsigned)((MQ + 255) / 256), 256, 0, c->stream>>>(
                        sq->x.as<uint64_t>(), sq->y.as<uint64_t>(), c->flt.as<uint8_t>(), MQ, B.view, B.mid_occ, K,
                        c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(), c->acnt.as<uint32_t>(), ctr, ctr + 1));
        } else {
            KLAUNCH(c, "seed_lookup_kernel", 28.0 * MQ,
                    seed_lookup_kernel<<<(unsigned)((MQ + SL_THREADS * SL_ILP - 1) / (SL_THREADS * SL_ILP)), SL_THREADS, 0, c->stream>>>(
                        sq->x.as<uint64_t>(), c->flt.as<uint8_t>(), MQ, B.view, B.mid_occ, c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(),
                        c->acnt.as<uint32_t>(), ctr, ctr + 1));
        }
        size_t tb = 0;
        cub::TransformInputIterator<uint64_t, WidenU32, const uint32_t *> in(c->acnt.as<uint32_t>(), WidenU32());
        cub::DeviceScan::ExclusiveSum(nullptr, tb, in, c->aoff.as<uint64_t>(), MQ + 1, c->stream);
        KLAUNCH(c, "cub::DeviceScan(u64)", 12.0 * MQ,
                cub::DeviceScan::ExclusiveSum(cub_temp(c, tb), tb, in, c->aoff.as<uint64_t>(), MQ + 1, c->stream));
        // anchors before each query: aoff[moff[q]]
        CUDA_CHECK(cudaMemcpyAsync(moff_h.data(), sq->moff.p, ((size_t)str->n + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        KLAUNCH(c, "gather_kernel", 20.0 * str->n,
                gather_u64_kernel<<<(str->n + 1 + 255) / 256, 256, 0, c->stream>>>(c->aoff.as<uint64_t>(), sq->moff.as<uint32_t>(), str->n + 1,
                                                                                  c->qaoff.as<uint64_t>()));
        CUDA_CHECK(cudaMemcpyAsync(qa.data(), c->qaoff.p, ((size_t)str->n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    CUDA_CHECK(cudaGetLastError());
    c->prof.send(ST_SEED);
    S.n_anchors = qa[str->n];

    // batches of queries bounded by anchor memory (34 bytes per anchor: unsorted 12, sorted 8, DP scratch 12, lists 2)
    const size_t free_b = free_estimate(c);
    size_t reusable = c->akey0.cap + c->akey1.cap + c->aval0.cap + c->aval1.cap + c->flags.cap + c->pstart.cap +
                      c->f.cap + c->p.cap + c->t.cap + c->pair_res.cap + c->anc.cap + c->fp.cap + c->glist.cap + c->pflag.cap;
    uint64_t a_max = (uint64_t)((free_b + reusable) * 0.7 / 56.0);
    a_max = std::min<uint64_t>(a_max, (1ull << 31) - 1024);
    if (c->max_anchors) a_max = std::min<uint64_t>(a_max, c->max_anchors);   // LRGE_MAX_ANCHORS: force small batches (tests)

    // per-query sort kernel configuration
    uint32_t r0 = 0, r1 = idx->n;
    if (sharded_index && !c->local_block) shard_range(idx->n, c->rank, world, &r0, &r1);
    const uint32_t n_local = r1 - r0;
    const int key_bits = bits_for(n_local ? n_local - 1 : 0) + 1 + PB;
    const bool k32 = key_bits <= 32, two = c->map_ctas_per_sm == 2;          // 2 CTAs of 512 threads, else 1 CTA of 1024
    const int NT = two ? 512 : 1024, NW = NT / 32;
    const void *sort_fn = two ? (k32 ? (const void *)sort_query_kernel<uint32_t, 512, 2> : (const void *)sort_query_kernel<uint64_t, 512, 2>)
                              : (k32 ? (const void *)sort_query_kernel<uint32_t, 1024, 1> : (const void *)sort_query_kernel<uint64_t, 1024, 1>);
    uint32_t map_cap = 0;
    size_t map_dyn = 0;
    if (use_smem_sort) {
        cudaFuncAttributes fa{};
        CUDA_CHECK(cudaFuncGetAttributes(&fa, sort_fn));
        const size_t rec = (k32 ? 4 : 8) + 4;
        size_t per_cta = (size_t)c->smem_optin;                                // static + dynamic, one CTA per SM
        if (two) per_cta = ((size_t)c->smem_optin + 1024) / 2 - 1024;
        const size_t fixed = fa.sharedSizeBytes + (size_t)NW * 512;
        // per slot: two (key, value) buffers, a u16 rank and 2 bits of the pair / group boundary maps
        map_cap = (uint32_t)std::min<size_t>((per_cta - fixed - 16) * 4 / (4 * (2 * rec + 2) + 1), 65504) / 32 * 32;
        if (c->map_cap) map_cap = std::max(64u, std::min(map_cap, c->map_cap) / 32 * 32);   // LRGE_MAP_CAP: force many rounds (tests)
        map_dyn = (size_t)map_cap * (2 * rec + 2) + (size_t)NW * 512 + 2 * ((size_t)map_cap / 32 + 1) * 4;
        CUDA_CHECK(cudaFuncSetAttribute(sort_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)map_dyn));
        c->deferred.reserve(((size_t)str->n + 1) * 4);
        c->mapctr.reserve(64);
        CUDA_CHECK(cudaMemsetAsync(c->mapctr.p, 0, 64, c->stream));     // [0] work counter (per batch) [1] deferred queries
    }
    const int min_viable = std::max({o.min_cnt, (o.min_chain_score + (o.is_hpc ? 255 : o.k) - 1) / (o.is_hpc ? 255 : o.k), 1});

    CountOut O{};
    O.counts = c->counts.as<uint32_t>(); O.had_mapping = c->had.as<uint8_t>();
    O.len_q = K.len_q; O.len_t = K.len_t; O.rank_q = K.rank_q; O.rank_t = K.rank_t;
    O.mode = mode; O.remove_internal = opts.remove_internal; O.ratio = opts.max_overhang_ratio;
    O.names_unique = !host_dedupe;
    if (want_regs) { O.regs = c->regs.as<RegOut>(); O.n_regs = ctr + 2; O.regs_cap = rq->cap; }

    // A batch is the contiguous query range [q0, q1), or -- for the queries the per-query kernel deferred -- the
    // scattered queries list[0 .. nlist), which then take local ids 0 .. nlist-1 in the 64-bit sort key.
    auto run_batch = [&](uint32_t q0, uint32_t q1, bool smem_sort, const uint32_t *list = nullptr, uint32_t nlist = 0) {
        uint64_t nA = 0;
        if (list) { for (uint32_t j = 0; j < nlist; ++j) nA += qa[list[j] + 1] - qa[list[j]]; }
        else nA = qa[q1] - qa[q0];
        const uint32_t nq_batch = list ? nlist : q1 - q0;
        if (nA > a_max) throw std::length_error("a single read yields more anchors than fit in device memory");
        const int QB = bits_for(nq_batch ? nq_batch - 1 : 0);
        if (QB + RB + 1 + PB > 64) throw std::length_error("sort key does not fit 64 bits (too many reads x read length)");
        if (!nA) return;
        c->prof.sbegin(ST_SEED);
        c->akey0.reserve(nA * 8); c->aval0.reserve(nA * 4);
        const uint64_t m0 = list ? 0 : moff_h[q0], m1 = list ? 0 : moff_h[q1];
        KeyLayout KL{PB, RB, r0};
        // SURVEY.md 8(d) K4 without the probes (seed_lookup_kernel): hit lists 8H + anchors 16A
        KLAUNCH(c, "anchor_fill_kernel", 8.0 * nA + 16.0 * nA, {
            const unsigned fg = (unsigned)((m1 - m0 + 255) / 256);
            if (list) {
                // one launch per listed query: its anchors follow those of the queries before it in the list
                uint64_t base = 0;
                for (uint32_t j = 0; j < nlist; ++j) {
                    const uint32_t q = list[j];
                    const uint64_t qm0 = moff_h[q], qm1 = moff_h[q + 1], nq_a = qa[q + 1] - qa[q];
                    if (qm1 > qm0 && nq_a)
                        anchor_fill_kernel<uint64_t, true><<<(unsigned)((qm1 - qm0 + 255) / 256), 256, 0, c->stream>>>(
                            sq->x.as<uint64_t>(), sq->y.as<uint64_t>(), c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(), c->acnt.as<uint32_t>(),
                            c->aoff.as<uint64_t>(), qm0, qm1, qa[q] - base, q - j, B.view, K, KL, c->akey0.as<uint64_t>(), c->aval0.as<uint32_t>());
                    base += nq_a;
                }
            } else if (!smem_sort)
                anchor_fill_kernel<uint64_t, true><<<fg, 256, 0, c->stream>>>(
                    sq->x.as<uint64_t>(), sq->y.as<uint64_t>(), c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(), c->acnt.as<uint32_t>(),
                    c->aoff.as<uint64_t>(), m0, m1, qa[q0], q0, B.view, K, KL, c->akey0.as<uint64_t>(), c->aval0.as<uint32_t>());
            else if (k32)
                anchor_fill_kernel<uint32_t, false><<<fg, 256, 0, c->stream>>>(
                    sq->x.as<uint64_t>(), sq->y.as<uint64_t>(), c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(), c->acnt.as<uint32_t>(),
                    c->aoff.as<uint64_t>(), m0, m1, qa[q0], q0, B.view, K, KL, c->akey0.as<uint32_t>(), c->aval0.as<uint32_t>());
            else
                anchor_fill_kernel<uint64_t, false><<<fg, 256, 0, c->stream>>>(
                    sq->x.as<uint64_t>(), sq->y.as<uint64_t>(), c->hstart.as<uint32_t>(), c->hn.as<uint32_t>(), c->acnt.as<uint32_t>(),
                    c->aoff.as<uint64_t>(), m0, m1, qa[q0], q0, B.view, K, KL, c->akey0.as<uint64_t>(), c->aval0.as<uint32_t>());
        });
        c->prof.send(ST_SEED);

        c->prof.sbegin(ST_SORT);
        // group descriptor lists, 4 size classes; a class holding groups of >= lb anchors has at most nA / lb entries
        const uint32_t lb[4] = {2048, 256, 48, (uint32_t)min_viable};
        GroupLists GL{};
        size_t n_desc = 0;
        for (int k = 0; k < 4; ++k) { GL.cap[k] = (uint32_t)(nA / lb[k] + 1); n_desc += GL.cap[k]; }
        c->glist.reserve(n_desc * sizeof(GroupDesc));
        { size_t off = 0; for (int k = 0; k < 4; ++k) { GL.gl[k] = c->glist.as<GroupDesc>() + off; off += GL.cap[k]; } }
        c->gmeta.reserve(64);
        c->pflag.reserve((n_desc + 1) * 4);
        c->anc.reserve(nA * 8);
        CUDA_CHECK(cudaMemsetAsync(c->gmeta.p, 0, 64, c->stream));          // [0..3] class counts [8] chain work counter [9] pair slots
        CUDA_CHECK(cudaMemsetAsync(c->pflag.p, 0, (n_desc + 1) * 4, c->stream));
        size_t tb = 0;
        bool v1_done = false;
        if (smem_sort) {
            std::vector<uint32_t> order;
            const uint32_t *order_p = str->len_order.data();
            if (q0 != 0 || q1 != str->n) {
                order.reserve(q1 - q0);
                for (uint32_t q : str->len_order) if (q >= q0 && q < q1) order.push_back(q);
                order_p = order.data();
                CUDA_CHECK(cudaStreamSynchronize(c->stream));               // `order` is a temporary: previous batch's copy must be done
            }
            h2d(c, c->qorder, order_p, q1 - q0);
            if (order_p != str->len_order.data()) CUDA_CHECK(cudaStreamSynchronize(c->stream));
            CUDA_CHECK(cudaMemsetAsync(c->mapctr.p, 0, 4, c->stream));
            SortArgs A{};
            c->akey1.reserve(nA * 8); c->aval1.reserve(nA * 4);
            A.akey = c->akey0.p; A.aval = c->aval0.as<uint32_t>(); A.qaoff = c->qaoff.as<uint64_t>(); A.a_base = qa[q0];
            A.akey2 = c->akey1.p; A.aval2 = c->aval1.as<uint32_t>();
            A.qorder = c->qorder.as<uint32_t>(); A.nq = q1 - q0;
            A.work_ctr = c->mapctr.as<uint32_t>(); A.n_deferred = c->mapctr.as<uint32_t>() + 1; A.deferred = c->deferred.as<uint32_t>();
            A.cap = map_cap; A.PB = PB; A.key_bits = key_bits; A.rid_base = r0;
            A.rel_bits = bits_for(n_local ? n_local - 1 : 0);
            A.max_span = o.is_hpc ? 255 : o.k; A.min_cnt = o.min_cnt; A.min_sc = o.min_chain_score;
            A.anc = c->anc.as<uint2>(); A.L = GL; A.gcount = c->gmeta.as<uint32_t>(); A.pair_ctr = c->gmeta.as<uint32_t>() + 9;
            A.stats = ctr + 3;
            A.cyc = c->prof.kernels ? ctr + 10 : nullptr;                   // [10..14]: cycles per phase (profiling only)
            const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)c->n_sm * (two ? 2 : 1), q1 - q0);
            // one read + one write of the 16-byte anchor records
            KLAUNCH(c, "sort_query_kernel", 32.0 * nA, {
                if (two) {
                    if (k32) sort_query_kernel<uint32_t, 512, 2><<<grid, 512, map_dyn, c->stream>>>(A);
                    else sort_query_kernel<uint64_t, 512, 2><<<grid, 512, map_dyn, c->stream>>>(A);
                } else {
                    if (k32) sort_query_kernel<uint32_t, 1024, 1><<<grid, 1024, map_dyn, c->stream>>>(A);
                    else sort_query_kernel<uint64_t, 1024, 1><<<grid, 1024, map_dyn, c->stream>>>(A);
                }
            });
            CUDA_CHECK(cudaGetLastError());
        } else {
            c->akey1.reserve(nA * 8); c->aval1.reserve(nA * 4);
            const int sort_bits = QB + RB + 1 + PB;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, c->akey0.as<uint64_t>(), c->akey1.as<uint64_t>(), c->aval0.as<uint32_t>(),
                                            c->aval1.as<uint32_t>(), (int64_t)nA, 0, sort_bits, c->stream);
            KLAUNCH(c, "cub::DeviceRadixSort(anchors: u64 key, u32 value)", 32.0 * nA,
                    cub::DeviceRadixSort::SortPairs(cub_temp(c, tb), tb, c->akey0.as<uint64_t>(), c->akey1.as<uint64_t>(), c->aval0.as<uint32_t>(),
                                                    c->aval1.as<uint32_t>(), (int64_t)nA, 0, sort_bits, c->stream));
            c->prof.set_passes("cub::DeviceRadixSort(anchors: u64 key, u32 value)", (sort_bits + 7) / 8);
            c->launches += (sort_bits + 7) / 8;
            c->flags.reserve(nA + 1);
            c->npairs.reserve(16);
#ifdef LRGE_TEST_KERNELS
            if (c->chain_v1) {
                v1_done = true;
                // ---- v1 (debug cross-check, LRGE_CHAIN_V1=1): one thread per (query, target) pair ----
                c->pstart.reserve((nA + 2) * 4);
                KLAUNCH(c, "pair_flag_kernel", 9.0 * nA,
                        pair_flag_kernel<<<(unsigned)((nA + 255) / 256), 256, 0, c->stream>>>(c->akey1.as<uint64_t>(), nA, PB + 1, c->flags.as<uint8_t>()));
                cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(), c->pstart.as<uint32_t>(),
                                           c->npairs.as<uint32_t>(), (int64_t)nA, c->stream);
                cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(),
                                           c->pstart.as<uint32_t>(), c->npairs.as<uint32_t>(), (int64_t)nA, c->stream);
                c->launches += 2;
                uint32_t n_pairs = 0;
                CUDA_CHECK(cudaMemcpyAsync(&n_pairs, c->npairs.p, 4, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaStreamSynchronize(c->stream));
                const uint32_t sentinel = (uint32_t)nA;
                CUDA_CHECK(cudaMemcpyAsync(c->pstart.as<uint32_t>() + n_pairs, &sentinel, 4, cudaMemcpyHostToDevice, c->stream));
                S.n_pairs += n_pairs;
                c->f.reserve(nA * 4); c->p.reserve(nA * 4); c->t.reserve(nA * 4);
                c->pair_res.reserve((size_t)n_pairs + 1);
                CUDA_CHECK(cudaMemsetAsync(c->t.p, 0, nA * 4, c->stream));
                PairTask T{};
                T.akey = c->akey1.as<uint64_t>(); T.aval = c->aval1.as<uint32_t>(); T.pstart = c->pstart.as<uint32_t>();
                T.n_pairs = n_pairs; T.PB = PB; T.RB = RB; T.q0 = q0;
                T.f = c->f.as<int32_t>(); T.p = c->p.as<int32_t>(); T.t = c->t.as<int32_t>();
                O.pair_res = c->pair_res.as<uint8_t>();
                KLAUNCH(c, "chain_pair_kernel", 16.0 * nA + 4.0 * str->n, chain_pair_kernel<<<(n_pairs + 127) / 128, 128, 0, c->stream>>>(T, P, O));
                CUDA_CHECK(cudaGetLastError());
                if (host_dedupe) {
                    std::vector<uint8_t> res(n_pairs);
                    std::vector<uint32_t> ps((size_t)n_pairs + 1);
                    std::vector<uint64_t> keys(nA);
                    CUDA_CHECK(cudaMemcpyAsync(res.data(), c->pair_res.p, n_pairs, cudaMemcpyDeviceToHost, c->stream));
                    CUDA_CHECK(cudaMemcpyAsync(ps.data(), c->pstart.p, ((size_t)n_pairs + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
                    CUDA_CHECK(cudaMemcpyAsync(keys.data(), c->akey1.p, nA * 8, cudaMemcpyDeviceToHost, c->stream));
                    CUDA_CHECK(cudaStreamSynchronize(c->stream));
                    for (uint32_t i = 0; i < n_pairs; ++i) if (res[i] & 2) {
                        const uint64_t kk = keys[ps[i]];
                        const uint32_t rid = (uint32_t)((kk >> (PB + 1)) & ((1ull << RB) - 1)), qid = q0 + (uint32_t)(kk >> (PB + 1 + RB));
                        dedupe_sets[qid].push_back(rank_idx[rid]);
                    }
                }
            }
#endif
            if (!v1_done) {
                c->pstart.reserve((nA + 2) * 4);
                KLAUNCH(c, "group_mark_kernel", 21.0 * nA,
                        group_mark_kernel<<<(unsigned)((nA + 255) / 256), 256, 0, c->stream>>>(c->akey1.as<uint64_t>(), c->aval1.as<uint32_t>(), nA, PB,
                                                                                               c->anc.as<uint2>(), c->flags.as<uint8_t>()));
                cub::DeviceSelect::Flagged(nullptr, tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(), c->pstart.as<uint32_t>(),
                                           c->npairs.as<uint32_t>(), (int64_t)nA, c->stream);
                KLAUNCH(c, "cub::DeviceSelect(group starts)", 1.0 * nA,
                        cub::DeviceSelect::Flagged(cub_temp(c, tb), tb, cub::CountingInputIterator<uint32_t>(0), c->flags.as<uint8_t>(),
                                                   c->pstart.as<uint32_t>(), c->npairs.as<uint32_t>(), (int64_t)nA, c->stream));
                uint32_t n_groups = 0;
                CUDA_CHECK(cudaMemcpyAsync(&n_groups, c->npairs.p, 4, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaStreamSynchronize(c->stream));
                const uint32_t sentinel = (uint32_t)nA;
                CUDA_CHECK(cudaMemcpyAsync(c->pstart.as<uint32_t>() + n_groups, &sentinel, 4, cudaMemcpyHostToDevice, c->stream));
                if ((size_t)n_groups + 1 > n_desc) { c->pflag.reserve(((size_t)n_groups + 2) * 4); CUDA_CHECK(cudaMemsetAsync(c->pflag.p, 0, ((size_t)n_groups + 2) * 4, c->stream)); }
                const uint32_t *qmap = nullptr;
                if (list) { h2d(c, c->qorder, list, nlist); qmap = c->qorder.as<uint32_t>(); }
                KLAUNCH(c, "group_classify_kernel", 8.0 * n_groups,
                        group_classify_kernel<<<(n_groups + 255) / 256, 256, 0, c->stream>>>(c->akey1.as<uint64_t>(), c->pstart.as<uint32_t>(), n_groups, PB, RB,
                                                                                            q0, qmap, P.min_cnt, P.min_sc, o.is_hpc ? 255 : o.k, GL,
                                                                                            c->gmeta.as<uint32_t>()));
                S.n_pairs += n_groups;
            }
        }
        c->prof.send(ST_SORT);
        if (v1_done) return;

        // ---- K6: one warp per (query, target, strand) group, persistent CTAs, largest groups first -------------------
        c->prof.sbegin(ST_CHAIN);
        c->fp.reserve(nA * 8); c->t.reserve(nA * 4);
        CUDA_CHECK(cudaMemsetAsync(c->t.p, 0, nA * 4, c->stream));
        GroupTask T{};
        T.anc = c->anc.as<uint2>();
        for (int k = 0; k < 4; ++k) T.gl[k] = GL.gl[k];
        T.gcount = c->gmeta.as<uint32_t>(); T.work_ctr = c->gmeta.as<uint32_t>() + 8;
        T.fp = c->fp.as<int2>(); T.t = c->t.as<int32_t>(); T.pflag = c->pflag.as<uint32_t>();
        if (host_dedupe) {
            c->ddbuf.reserve((n_desc + 1) * 8);
            CUDA_CHECK(cudaMemsetAsync(ctr + 9, 0, 8, c->stream));
            T.dd_out = c->ddbuf.as<uint2>(); T.dd_n = ctr + 9;
        }
        KLAUNCH(c, "chain_group_kernel", 16.0 * nA + 4.0 * nq_batch,
                chain_group_kernel<<<c->n_sm * (2048 / CG_THREADS), CG_THREADS, 0, c->stream>>>(T, P, O));
        CUDA_CHECK(cudaGetLastError());
        if (host_dedupe) {                                                   // passing (query, target) pairs -> HashSet<name> per query
            unsigned long long n_dd = 0;
            CUDA_CHECK(cudaMemcpyAsync(&n_dd, ctr + 9, 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            std::vector<uint2> dd(n_dd);
            if (n_dd) CUDA_CHECK(cudaMemcpyAsync(dd.data(), c->ddbuf.p, n_dd * 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            for (const uint2 &e : dd) dedupe_sets[e.x].push_back(rank_idx[e.y]);
        }
        c->prof.send(ST_CHAIN);
    };

    {
        uint32_t q0 = 0;
        while (q0 < str->n) {
            uint32_t q1 = q0 + 1;
            while (q1 < str->n && qa[q1 + 1] - qa[q0] <= a_max) ++q1;
            if (qa[q1] > qa[q0]) ++S.n_batches;
            run_batch(q0, q1, use_smem_sort);
            q0 = q1;
        }
    }
    // queries the per-query kernel could not plan (one target bin above the smem capacity): device-wide sort, in as few
    // batches as the anchor memory allows.  Single GPU: the check rides on the final copy of the results (no extra
    // synchronisation in the usual case); sharded: before the count reduction.
    std::vector<uint32_t> deferred;
    uint32_t mapctr_h[2] = {0, 0};
    auto run_deferred = [&]() {
        deferred.resize(mapctr_h[1]);
        if (!mapctr_h[1]) return;
        CUDA_CHECK(cudaMemcpyAsync(deferred.data(), c->deferred.p, (size_t)mapctr_h[1] * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::sort(deferred.begin(), deferred.end());
        size_t lo = 0;
        while (lo < deferred.size()) {
            size_t hi = lo;
            uint64_t sum = 0;
            while (hi < deferred.size() && (hi == lo || sum + (qa[deferred[hi] + 1] - qa[deferred[hi]]) <= a_max)) { sum += qa[deferred[hi] + 1] - qa[deferred[hi]]; ++hi; }
            ++S.n_batches;
            run_batch(0, 0, false, deferred.data() + lo, (uint32_t)(hi - lo));
            lo = hi;
        }
    };
    const bool have_map = use_smem_sort && str->n && MQ;
    if (have_map && sharded_index) {
        CUDA_CHECK(cudaMemcpyAsync(mapctr_h, c->mapctr.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        run_deferred();
    }

    // one exchange on the query path: sum the per-read counts of all target shards over NVLink
    if (sharded_index) {
        c->prof.sbegin(ST_COMM);
        ncclResult_t r1 = ncclAllReduce(c->counts.p, c->counts.p, n_counts, ncclUint32, ncclSum, c->comm, c->stream);
        ncclResult_t r2 = ncclAllReduce(c->had.p, c->had.p, str->n, ncclUint8, ncclMax, c->comm, c->stream);
        if (r1 != ncclSuccess || r2 != ncclSuccess) throw NcclError(std::string("ncclAllReduce: ") + ncclGetErrorString(r1 != ncclSuccess ? r1 : r2));
        c->prof.send(ST_COMM);
    }

    // results
    std::vector<uint32_t> counts(n_counts);
    std::vector<uint8_t> had(str->n);
    unsigned long long ctr_h[16] = {};
    auto fetch_results = [&]() {
        c->prof.sbegin(ST_D2H);
        if (n_counts) CUDA_CHECK(cudaMemcpyAsync(counts.data(), c->counts.p, (size_t)n_counts * 4, cudaMemcpyDeviceToHost, c->stream));
        if (str->n) CUDA_CHECK(cudaMemcpyAsync(had.data(), c->had.p, str->n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(ctr_h, ctr, 128, cudaMemcpyDeviceToHost, c->stream));
        c->prof.send(ST_D2H);
    };
    fetch_results();
    if (have_map && !sharded_index) CUDA_CHECK(cudaMemcpyAsync(mapctr_h, c->mapctr.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (have_map && !sharded_index && mapctr_h[1]) {
        run_deferred();
        fetch_results();
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    if (want_regs) {
#ifdef LRGE_TEST_KERNELS
        if (c->chain_v1) throw std::runtime_error("overlap records need the warp-per-group kernel (unset LRGE_CHAIN_V1)");
#endif
        *rq->n_out = ctr_h[2];
        const uint64_t m = std::min<uint64_t>(ctr_h[2], rq->cap);
        if (m && rq->out) {
            // records + the query sketch to the host, then the PAF annotations (namespace paf above)
            std::vector<RegOut> raw(m);
            std::vector<uint64_t> hx(MQ), hy(MQ);
            std::vector<uint8_t> hflt(MQ);
            std::vector<uint32_t> hhs(MQ), hhn(MQ);
            CUDA_CHECK(cudaMemcpyAsync(raw.data(), c->regs.p, m * sizeof(RegOut), cudaMemcpyDeviceToHost, c->stream));
            if (MQ) {
                CUDA_CHECK(cudaMemcpyAsync(hx.data(), sq->x.p, MQ * 8, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaMemcpyAsync(hy.data(), sq->y.p, MQ * 8, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaMemcpyAsync(hflt.data(), c->flt.p, MQ, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaMemcpyAsync(hhs.data(), c->hstart.p, MQ * 4, cudaMemcpyDeviceToHost, c->stream));
                CUDA_CHECK(cudaMemcpyAsync(hhn.data(), c->hn.p, MQ * 4, cudaMemcpyDeviceToHost, c->stream));
            }
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            std::vector<uint32_t> order(m);
            std::iota(order.begin(), order.end(), 0u);
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return raw[a].qid < raw[b].qid; });
            std::vector<uint64_t> zx(m);
            std::vector<float> dv(m, -1.0f);
            std::vector<int32_t> rl(m, 0);
            size_t lo = 0;
            while (lo < m) {
                size_t hi = lo;
                const uint32_t q = raw[order[lo]].qid;
                while (hi < m && raw[order[hi]].qid == q) ++hi;
                const paf::QuerySeeds QS = paf::query_seeds(hx.data(), hy.data(), hflt.data(), hhs.data(), hhn.data(), moff_h[q], moff_h[q + 1]);
                const int32_t qlen = (int32_t)str->h_len[q];
                uint32_t qhash = paf::x31_hash(str->names.data() + str->name_off[q], str->name_off[q + 1] - str->name_off[q]);
                qhash ^= paf::wang_hash((uint32_t)qlen) + paf::wang_hash(11u);      // mm_mapopt_t.seed = 11
                qhash = paf::wang_hash(qhash);
                const int32_t n_mp = (int32_t)QS.pos.size();
                auto mini_idx = [&](int32_t p) -> int32_t {
                    auto it = std::lower_bound(QS.pos.begin(), QS.pos.end(), p);
                    return it != QS.pos.end() && *it == p ? (int32_t)(it - QS.pos.begin()) : -1;
                };
                for (size_t e = lo; e < hi; ++e) {
                    const RegOut &r = raw[order[e]];
                    const int32_t span0 = (int32_t)(r.a0y >> 24), y0 = (int32_t)(r.a0y & 0xffffff);
                    const int32_t span1 = (int32_t)(r.a1y >> 24), y1 = (int32_t)(r.a1y & 0xffffff);
                    auto for_qpos = [&](int32_t y, int32_t span) { return r.rev ? qlen - 1 - (y + 1 - span) : y; };   // get_for_qpos
                    const int32_t i0 = mini_idx(for_qpos(y0, span0)), i1 = mini_idx(for_qpos(y1, span1));
                    rl[order[e]] = QS.rep_len;
                    // mm_est_err: every anchor of a chain is a kept seed of the query, visited in query order, so
                    // n_match = cnt and [st, en] are the seeds between the chain's two ends
                    const int32_t st = r.rev ? i1 : i0, en = r.rev ? i0 : i1;
                    if (n_mp && r.cnt > 0 && st >= 0 && en >= 0) {
                        const int32_t l_ref = (int32_t)idx->h_len[r.rid];
                        const int32_t n_match = r.cnt;
                        int32_t n_tot = en - st + 1;
                        if (r.qs > QS.avg_k && r.rs > QS.avg_k) ++n_tot;
                        if (qlen - r.qs > QS.avg_k && l_ref - r.re > QS.avg_k) ++n_tot;
                        dv[order[e]] = n_match >= n_tot ? 0.0f : (float)(1.0 - pow((double)n_match / n_tot, 1.0 / QS.avg_k));
                    }
                    // mm_gen_regs: h = hash64((hash64(a[k].x) + hash64(a[k].y)) ^ hash) of the chain's first anchor
                    const bool tandem = i0 >= 0 && QS.tandem[i0];
                    const bool self = rank_str[q] == rank_idx[r.rid] && idx->h_len[r.rid] == str->h_len[q] && !r.rev;
                    const uint64_t ax = (uint64_t)(r.rev ? 1 : 0) << 63 | (uint64_t)r.rid << 32 | r.a0x;
                    const uint64_t ay = (uint64_t)span0 << 32 | (uint64_t)y0 | (tandem ? 1ULL << 42 : 0) | (self ? 1ULL << 43 : 0);
                    const uint32_t h = (uint32_t)paf::hash64((paf::hash64(ax) + paf::hash64(ay)) ^ qhash);
                    zx[order[e]] = ((uint64_t)(uint32_t)r.score << 32 | (uint32_t)r.cnt) ^ h;
                }
                std::sort(order.begin() + lo, order.begin() + hi, [&](uint32_t a, uint32_t b) { return zx[a] > zx[b]; });
                lo = hi;
            }
            for (uint64_t e = 0; e < m; ++e) {
                const RegOut &r = raw[order[e]];
                lrge_overlap &o2 = rq->out[e];
                o2.query = r.qid; o2.target = r.rid; o2.rev = r.rev; o2.qs = r.qs; o2.qe = r.qe; o2.ts = r.rs; o2.te = r.re;
                o2.score = r.score; o2.cnt = r.cnt; o2.mlen = r.mlen; o2.blen = r.blen; o2.dv = dv[order[e]]; o2.rl = rl[order[e]];
            }
        }
    }
    c->prof.send(ST_TOTAL);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (use_smem_sort) {
        if (ctr_h[6]) throw std::runtime_error("internal: the per-query sort kernel exceeded a capacity it had planned for");
        S.n_pairs += ctr_h[4]; S.n_rounds = (uint32_t)ctr_h[5];
        S.n_deferred = (uint32_t)deferred.size();
        if (c->prof.kernels && ctr_h[14] && getenv("LRGE_SORT_PHASES")) {     // phase shares of the per-query sort kernel (SM cycles)
            const double tot = (double)ctr_h[14];
            fprintf(stderr, "[lrge] sort_query_kernel phases: plan %.3f stage %.3f sort %.3f groups %.3f (of CTA lifetime), rounds %llu\n",
                    ctr_h[10] / tot, ctr_h[11] / tot, ctr_h[12] / tot, ctr_h[13] / tot, (unsigned long long)ctr_h[5]);
        }
    }
    c->prof.resolve();
    S.ms_pack = (float)c->prof.ms[ST_PACK]; S.ms_sketch = (float)c->prof.ms[ST_SKETCH]; S.ms_index = (float)c->prof.ms[ST_INDEX];
    S.ms_seed = (float)c->prof.ms[ST_SEED]; S.ms_sort = (float)c->prof.ms[ST_SORT]; S.ms_chain = (float)c->prof.ms[ST_CHAIN];
    S.ms_d2h = (float)c->prof.ms[ST_D2H]; S.ms_comm = (float)c->prof.ms[ST_COMM]; S.ms_total = (float)c->prof.ms[ST_TOTAL];
    for (size_t i = 0; i < c->prof.names.size(); ++i)
        if (c->prof.names[i] == "chain_group_kernel" || c->prof.names[i] == "chain_pair_kernel") S.ms_chain_kernel += (float)c->prof.ms[ST_COUNT + i];
    if (host_dedupe)
        for (uint32_t q = 0; q < str->n; ++q) {
            auto &v = dedupe_sets[q];
            std::sort(v.begin(), v.end());
            counts[q] = (uint32_t)(std::unique(v.begin(), v.end()) - v.begin());
        }
    S.n_seeds = ctr_h[0]; S.n_hits = ctr_h[1];
    uint64_t sum = 0;
    for (uint32_t i = 0; i < n_counts; ++i) sum += counts[i];
    S.n_overlaps = mode == MODE_AVA ? sum / 2 : sum;
    S.no_mapping_count = 0;
    if (mode == MODE_TWOSET) { for (uint32_t q = 0; q < str->n; ++q) S.no_mapping_count += !had[q]; }
    else for (uint32_t i = 0; i < n_counts; ++i) S.no_mapping_count += counts[i] == 0;
    if (counts_out && n_counts) memcpy(counts_out, counts.data(), (size_t)n_counts * 4);
    if (had_out && mode == MODE_TWOSET && str->n) memcpy(had_out, had.data(), str->n);
    S.n_kernel_launches = c->launches - launches0;
    if (st) { const float h2d_ms = st->ms_h2d; *st = S; st->ms_h2d = h2d_ms; }
}

template <typename Fn> int guarded(lrge_ctx *ctx, Fn fn) {
    if (!ctx) return LRGE_ERR_INVALID_ARG;
    try {
        CUDA_CHECK(cudaSetDevice(ctx->device));
        fn();
        return LRGE_OK;
    } catch (const CudaError &e) { ctx->err = e.what(); cudaGetLastError(); return LRGE_ERR_CUDA; }
    catch (const NcclError &e) { ctx->err = e.what(); return LRGE_ERR_NCCL; }
    catch (const std::domain_error &e) { ctx->err = e.what(); return LRGE_ERR_DUPLICATE_READ_ID; }
    catch (const std::invalid_argument &e) { ctx->err = e.what(); return LRGE_ERR_MAP; }
    catch (const std::length_error &e) { ctx->err = e.what(); return LRGE_ERR_UNSUPPORTED; }
    catch (const std::bad_alloc &) { ctx->err = "host out of memory"; return LRGE_ERR_OUT_OF_MEMORY; }
    catch (const std::exception &e) { ctx->err = e.what(); return LRGE_ERR_INVALID_ARG; }
}

bool valid_opts(const lrge_opts *o) {
    return o && (o->platform == LRGE_PLATFORM_NANOPORE || o->platform == LRGE_PLATFORM_PACBIO) && o->shard_count >= 1 &&
           o->shard_index >= 0 && o->shard_index < o->shard_count;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

void lrge_opts_default(lrge_opts *o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->platform = LRGE_PLATFORM_NANOPORE;      // Platform::default (lib.rs:163-168)
    o->remove_internal = 0;                    // twoset/builder.rs:30
    o->max_overhang_ratio = 0.2f;              // twoset/builder.rs:31
    o->use_min_ref = 0;                        // twoset/builder.rs:32
    o->shard_index = 0; o->shard_count = 1;
}

int lrge_gpu_create(int device, lrge_ctx **out) {
    if (!out) return LRGE_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (liblrge_b200 has no CPU fallback)";
        cudaGetLastError();
        return LRGE_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return LRGE_ERR_INVALID_ARG; }
    try {
        CUDA_CHECK(cudaSetDevice(device));
        std::unique_ptr<lrge_ctx> c(new lrge_ctx());
        c->device = device;
        { const char *e = getenv("LRGE_CHAIN_V1"); c->chain_v1 = e && e[0] == '1'; }
        { const char *e = getenv("LRGE_MAX_ANCHORS"); c->max_anchors = e ? strtoull(e, nullptr, 10) : 0; }
        { const char *e = getenv("LRGE_FORCE_GLOBAL"); c->force_global = e && e[0] == '1'; }
        { const char *e = getenv("LRGE_INDEX_BITS_DELTA"); c->index_bits_delta = e ? atoi(e) : 0; }
        { const char *e = getenv("LRGE_MAP_CAP"); c->map_cap = e ? (uint32_t)strtoul(e, nullptr, 10) : 0; }
        { const char *e = getenv("LRGE_MAP_CTAS"); c->map_ctas_per_sm = e && e[0] == '2' ? 2 : 1; }
        CUDA_CHECK(cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        CUDA_CHECK(cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device));
        { const char *e = getenv("LRGE_SKETCH_TWO_PASS"); c->sketch_two_pass = e && e[0] == '1'; }
        { const char *e = getenv("LRGE_SKETCH_LEGACY"); c->sketch_legacy = (e && e[0] == '1') || c->sketch_two_pass; }
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->side_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        { const char *e = getenv("LRGE_NO_SIDE_STREAM"); c->no_side = e && e[0] == '1'; }
        { const char *e = getenv("LRGE_MIN_CHUNK"); if (e && atoll(e) > 0) c->min_chunk = (uint64_t)atoll(e); }
        { const char *e = getenv("LRGE_INDEX_HALVES"); c->no_halves = !(e && e[0] == '1'); }
        { const char *e = getenv("LRGE_TIMELINE"); c->prof.timeline = e && e[0] == '1'; }
        c->prof.s = c->stream;
        c->prof.reset();
        { size_t f = 0, t = 0; CUDA_CHECK(cudaMemGetInfo(&f, &t)); c->free_at_create = f; c->held_at_create = dev_bytes_in_use().load(); }
        CUDA_CHECK(cudaMemcpyToSymbol(c_nt4, kNt4, 256));
        CUDA_CHECK(cudaHostAlloc((void **)&c->pin, 64 + (OCC_BINS + 2) * 8 + 64, cudaHostAllocDefault));
        CUDA_CHECK(cudaHostAlloc((void **)&c->pin2, (33 + 32 * 32 + 16) * 8, cudaHostAllocDefault));
        *out = c.release();
        return LRGE_OK;
    } catch (const std::exception &ex) { g_create_error = ex.what(); cudaGetLastError(); return LRGE_ERR_CUDA; }
}

void lrge_gpu_destroy(lrge_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    peer_link_close(ctx);
    if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
    if (ctx->pin) { cudaFreeHost(ctx->pin); ctx->pin = nullptr; }
    if (ctx->pin2) { cudaFreeHost(ctx->pin2); ctx->pin2 = nullptr; }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    delete ctx;
}

const char *lrge_gpu_last_error(const lrge_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int lrge_gpu_stage(lrge_ctx *ctx, const lrge_readset *rs, lrge_staged **out) {
    if (!rs || !out) return LRGE_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded(ctx, [&] {
        std::unique_ptr<lrge_staged> s(new lrge_staged());
        stage_readset(ctx, rs, s.get());
        *out = s.release();
    });
}

void lrge_gpu_unstage(lrge_ctx *ctx, lrge_staged *s) {
    if (ctx) cudaSetDevice(ctx->device);
    delete s;
}

int lrge_gpu_twoset_counts_staged(lrge_ctx *ctx, lrge_staged *T, lrge_staged *Q, const lrge_opts *opts,
                                  uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats) {
    if (!T || !Q || !counts || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] {
        // twoset.rs:596-604: with use_min_ref and more target than query bases, index the query set
        if (opts->use_min_ref && T->total_bases > Q->total_bases) run_overlap(ctx, Q, T, MODE_INVERSE, *opts, counts, nullptr, stats);
        else run_overlap(ctx, T, Q, MODE_TWOSET, *opts, counts, had_mapping, stats);
    });
}

int lrge_gpu_ava_counts_staged(lrge_ctx *ctx, lrge_staged *R, const lrge_opts *opts, uint32_t *counts, lrge_stats *stats) {
    if (!R || !counts || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] { run_overlap(ctx, R, R, MODE_AVA, *opts, counts, nullptr, stats); });
}

int lrge_gpu_twoset_counts(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                           uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats) {
    if (!T || !Q || !counts || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] {
        const auto t0 = std::chrono::steady_clock::now();
        lrge_staged &sT = ctx->tmp_a, &sQ = ctx->tmp_b;
        const bool inverse = opts->use_min_ref && (T->n ? T->off[T->n] - T->off[0] : 0) > (Q->n ? Q->off[Q->n] - Q->off[0] : 0);
        struct Drain { lrge_ctx *c; ~Drain() { cudaStreamSynchronize(c->copy_stream); } } drain{ctx};   // caller buffers are free on return
        // the indexed set is needed first; the other copy overlaps its sketch
        uint32_t b0 = 0, b1 = 0xffffffffu;
        if (ctx->comm && opts->shard_count > 1) {                 // sharded index: only this rank's block of the indexed set
            uint32_t lo, hi;
            shard_range(inverse ? Q->n : T->n, opts->shard_index, opts->shard_count, &lo, &hi);
            b0 = lo; b1 = hi - lo;
        }
        stage_readset(ctx, inverse ? Q : T, inverse ? &sQ : &sT, true, b0, b1, true);
        stage_readset(ctx, inverse ? T : Q, inverse ? &sT : &sQ, true, 0, 0xffffffffu, true);
        if (inverse) run_overlap(ctx, &sQ, &sT, MODE_INVERSE, *opts, counts, nullptr, stats);
        else run_overlap(ctx, &sT, &sQ, MODE_TWOSET, *opts, counts, had_mapping, stats);
        if (stats) {   // host buffers in, counts out: wall time of the whole call
            stats->ms_h2d = 0.0f;
            stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    });
}

int lrge_gpu_twoset_counts_shard(lrge_ctx *ctx, const lrge_readset *T_block, const lrge_readset *Q, const lrge_opts *opts,
                                 uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats) {
    if (!T_block || !Q || !counts || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    if (!ctx->comm || opts->shard_count != ctx->world || opts->shard_index != ctx->rank || opts->use_min_ref) {
        ctx->err = "lrge_gpu_twoset_counts_shard needs a communicator, matching shard_index/shard_count and use_min_ref = 0";
        return LRGE_ERR_INVALID_ARG;
    }
    return guarded(ctx, [&] {
        const auto t0 = std::chrono::steady_clock::now();
        lrge_staged &sT = ctx->tmp_a, &sQ = ctx->tmp_b;
        struct Drain { lrge_ctx *c; ~Drain() { c->local_block = false; cudaStreamSynchronize(c->copy_stream); } } drain{ctx};
        ctx->local_block = true;
        stage_readset(ctx, T_block, &sT, true, 0, 0xffffffffu, true);
        stage_readset(ctx, Q, &sQ, true, 0, 0xffffffffu, true);
        run_overlap(ctx, &sT, &sQ, MODE_TWOSET, *opts, counts, had_mapping, stats);
        if (stats) {
            stats->ms_h2d = 0.0f;
            stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    });
}

int lrge_gpu_twoset_overlaps(lrge_ctx *ctx, const lrge_readset *T, const lrge_readset *Q, const lrge_opts *opts,
                             lrge_overlap *out, uint64_t cap, uint64_t *n_out, uint32_t *counts, uint8_t *had_mapping, lrge_stats *stats) {
    if (!T || !Q || !n_out || (!out && cap) || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] {
        lrge_staged &sT = ctx->tmp_a, &sQ = ctx->tmp_b;
        stage_readset(ctx, T, &sT);
        stage_readset(ctx, Q, &sQ);
        RegsRequest rq; rq.out = out; rq.cap = cap; rq.n_out = n_out;
        std::vector<uint32_t> tmp_counts;
        if (opts->use_min_ref && sT.total_bases > sQ.total_bases) {
            // inverse direction: records carry query = streamed (target-set) read, target = indexed (query-set) read
            if (!counts) { tmp_counts.resize(Q->n); counts = tmp_counts.data(); }
            run_overlap(ctx, &sQ, &sT, MODE_INVERSE, *opts, counts, nullptr, stats, &rq);
        } else {
            if (!counts) { tmp_counts.resize(Q->n); counts = tmp_counts.data(); }
            run_overlap(ctx, &sT, &sQ, MODE_TWOSET, *opts, counts, had_mapping, stats, &rq);
        }
    });
}

int lrge_gpu_ava_overlaps(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts, lrge_overlap *out, uint64_t cap, uint64_t *n_out,
                          uint32_t *counts, lrge_stats *stats) {
    if (!R || !n_out || (!out && cap) || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] {
        lrge_staged &sR = ctx->tmp_a;
        stage_readset(ctx, R, &sR);
        RegsRequest rq; rq.out = out; rq.cap = cap; rq.n_out = n_out;
        std::vector<uint32_t> tmp_counts;
        if (!counts) { tmp_counts.resize(R->n); counts = tmp_counts.data(); }
        run_overlap(ctx, &sR, &sR, MODE_AVA, *opts, counts, nullptr, stats, &rq);
    });
}

int lrge_gpu_ava_counts(lrge_ctx *ctx, const lrge_readset *R, const lrge_opts *opts, uint32_t *counts, lrge_stats *stats) {
    if (!R || !counts || !valid_opts(opts)) { if (ctx) ctx->err = "invalid argument"; return LRGE_ERR_INVALID_ARG; }
    return guarded(ctx, [&] {
        const auto t0 = std::chrono::steady_clock::now();
        lrge_staged &sR = ctx->tmp_a;
        struct Drain { lrge_ctx *c; ~Drain() { cudaStreamSynchronize(c->copy_stream); } } drain{ctx};
        stage_readset(ctx, R, &sR, true, 0, 0xffffffffu, true);
        run_overlap(ctx, &sR, &sR, MODE_AVA, *opts, counts, nullptr, stats);
        if (stats) {
            stats->ms_h2d = 0.0f;
            stats->ms_total = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    });
}

int lrge_gpu_comm_unique_id(uint8_t *out, size_t size) {
    if (!out || size < sizeof(ncclUniqueId)) return LRGE_ERR_INVALID_ARG;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return LRGE_ERR_NCCL;
    memset(out, 0, size);
    memcpy(out, &id, sizeof id);
    return LRGE_OK;
}

int lrge_gpu_comm_init(lrge_ctx *ctx, int rank, int world, const uint8_t *unique_id, size_t size) {
    if (!unique_id || size < sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) return LRGE_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        if (ctx->comm) { ncclCommDestroy(ctx->comm); ctx->comm = nullptr; }
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof id);
        ncclResult_t r = ncclCommInitRank(&ctx->comm, world, id, rank);
        if (r != ncclSuccess) { ctx->comm = nullptr; throw NcclError(std::string("ncclCommInitRank: ") + ncclGetErrorString(r)); }
        ctx->rank = rank; ctx->world = world;
        // peer-memory link for the key exchange: flags now, receive windows on first use.  Collective; any failure on
        // any rank (no IPC in this container, no peer access) leaves every rank on ncclSend / ncclRecv.
        peer_link_close(ctx);
        ctx->link.epoch = 0;
        if (world > 1 && world <= 16) {
            const char *e = getenv("LRGE_NO_PEER_PUSH");
            bool have = !(e && e[0] == '1');
            if (have && cudaMalloc((void **)&ctx->link.flags, 4096) != cudaSuccess) { cudaGetLastError(); ctx->link.flags = nullptr; have = false; }
            if (ctx->link.flags) CUDA_CHECK(cudaMemset(ctx->link.flags, 0, 4096));
            ctx->link.done.reserve(256);
            CUDA_CHECK(cudaMemset(ctx->link.done.p, 0, 256));
            void *mapped[16] = {};
            if (exchange_and_open(ctx, ctx->link.flags, have, mapped)) {
                for (int d = 0; d < world; ++d) ctx->link.flags_peer[d] = (unsigned long long *)mapped[d];
                ctx->link.ok = true;
            }
        }
    });
}

int lrge_gpu_set_profiling(lrge_ctx *ctx, int on) {
    if (!ctx) return LRGE_ERR_INVALID_ARG;
    ctx->prof.kernels = on != 0;
    return LRGE_OK;
}

int lrge_gpu_kernel_times(lrge_ctx *ctx, lrge_kernel_time *out, int cap) {
    if (!ctx || (!out && cap > 0)) return -1;
    const int n = (int)ctx->prof.names.size();
    for (int i = 0; i < n && i < cap; ++i) {
        memset(&out[i], 0, sizeof(out[i]));
        strncpy(out[i].name, ctx->prof.names[i].c_str(), sizeof(out[i].name) - 1);
        out[i].ms = ctx->prof.ms[ST_COUNT + i];
        out[i].algorithmic_bytes = ctx->prof.bytes[ST_COUNT + i];
        out[i].launches = ctx->prof.count[ST_COUNT + i];
        out[i].passes = ctx->prof.passes[ST_COUNT + i];
    }
    return n;
}

int lrge_gpu_host_alloc(size_t bytes, void **out) {
    if (!out) return LRGE_ERR_INVALID_ARG;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) { *out = nullptr; cudaGetLastError(); return LRGE_ERR_CUDA; }
    return LRGE_OK;
}
void lrge_gpu_host_free(void *p) { if (p) cudaFreeHost(p); }

int lrge_gpu_dump_minimizers_block(lrge_ctx *ctx, const lrge_readset *rs, int32_t platform, uint32_t first, uint32_t count,
                                   uint64_t *x, uint64_t *y, uint64_t cap, uint64_t *n_out, uint32_t *cnt) {
    if (!rs || !n_out) return LRGE_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        lrge_staged s;
        stage_readset(ctx, rs, &s, false, first, count);     // only the block's bases go to the device
        const MapOpt o = preset_for(platform);
        sketch_readset(ctx, &s, o, &ctx->sk_str, s.blk_first, s.blk_count);
        const uint64_t M = ctx->sk_str.M;
        *n_out = M;
        const uint64_t m = std::min(M, cap);
        if (x && m) CUDA_CHECK(cudaMemcpyAsync(x, ctx->sk_str.x.p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (y && m) CUDA_CHECK(cudaMemcpyAsync(y, ctx->sk_str.y.p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (cnt && s.blk_count) {
            std::vector<uint32_t> moff((size_t)s.blk_count + 1);
            CUDA_CHECK(cudaMemcpyAsync(moff.data(), ctx->sk_str.moff.p, ((size_t)s.blk_count + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            for (uint32_t i = 0; i < s.blk_count; ++i) cnt[i] = moff[i + 1] - moff[i];
        }
        CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    });
}

int lrge_gpu_dump_minimizers(lrge_ctx *ctx, const lrge_readset *rs, int32_t platform, uint64_t *x, uint64_t *y, uint64_t cap,
                             uint64_t *n_out, uint32_t *cnt) {
    return lrge_gpu_dump_minimizers_block(ctx, rs, platform, 0, 0xffffffffu, x, y, cap, n_out, cnt);
}

int lrge_gpu_dump_mid_occ(lrge_ctx *ctx, const lrge_readset *T, int32_t platform, int32_t *mid_occ, uint64_t *n_keys) {
    if (!T || !mid_occ) return LRGE_ERR_INVALID_ARG;
    return guarded(ctx, [&] {
        lrge_staged s;
        stage_readset(ctx, T, &s);
        const MapOpt o = preset_for(platform);
        sketch_readset(ctx, &s, o, &ctx->sk_idx);
        BuiltIndex B = build_index(ctx, ctx->sk_idx, o);
        *mid_occ = B.mid_occ;
        if (n_keys) *n_keys = B.n_keys;
    });
}

}  // extern "C"
