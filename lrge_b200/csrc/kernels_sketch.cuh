// kernels_sketch.cuh -- K0 pack (ASCII -> 2-bit + ambiguity mask), K0h homopolymer compression,
// K1 minimizer sketch.
//
// Replaces minimap2 mm_sketch (sketch.c) as called per target at index time and per query in
// collect_minimizers.  mm_sketch is a sequential state machine; the emitted
// set is nevertheless a pure function of a bounded history, which is what lets every position be
// evaluated independently here:
//   * info[t] (hash<<8|span, or MAX) depends on the k bases ending at step t and on l_t, the number of
//     unambiguous steps since the last ambiguous base, saturated at w+k;
//   * the running `min` of mm_sketch is always the RIGHTMOST minimum of the last w infos, so
//     m[t-1] and m[t] follow from info[t-w..t];
//   * the pushes done at step t (first-window ties, replaced minimum, minimum leaving the window,
//     ties after a rescan, final flush) depend only on info[t-w..t], m[t-1], m[t] and l_t.
// Steps are base positions, or homopolymer runs in HPC mode (ava-pb).  k must be odd (15 / 19): the
// "symmetric k-mer" skip of mm_sketch cannot occur.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace lrge {

constexpr int PAD_BASES = 128;            // every read starts at a multiple of this in padded coordinates
constexpr int TILE_POS = 8192;            // steps per CTA tile
constexpr int SK_THREADS = 256;           // 8 warps x 1024 steps
constexpr int WARP_POS = TILE_POS / (SK_THREADS / 32);
constexpr int TILE_WORDS = (TILE_POS + PAD_BASES) / 32;   // 260 u64 of bases / u32 of mask
constexpr int MAXW = 5;                   // minimizer window of both presets
constexpr uint32_t SK_CAP = 768;          // single-pass staging: minimizers per 1024-step sub-tile (density 1/3 expected)

__constant__ uint8_t c_nt4[256];

struct ReadLayout {
    const uint8_t  *asc;       // ASCII bases (HPC: run-compressed ASCII)
    const uint64_t *off;       // n   start offset of each entry in asc
    const uint32_t *rid_map;   // n   read id of each entry (NULL: entry index = read id); a rank sketches a subset of reads
    const uint64_t *poff;      // n   padded start (bases), multiple of PAD_BASES, >= PAD_BASES
    const uint32_t *len;       // n   number of steps (HPC: compressed length)
    const uint32_t *tile_off;  // n+1 tiles before read r (tiles computed from the uncompressed length)
    uint64_t *w2;              // 2-bit packed, 32 steps per u64, first step in the top bits
    uint32_t *nm;              // ambiguity mask, 32 steps per u32, step i at bit i
    uint32_t n;
};

__device__ __forceinline__ uint32_t find_read(const uint32_t *tile_off, uint32_t n, uint32_t tile) {
    uint32_t lo = 0, hi = n;   // largest r with tile_off[r] <= tile
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_off[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- K0: pack ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(SK_THREADS) pack_kernel(ReadLayout L) {
    __shared__ uint32_t s_r;
    if (threadIdx.x == 0) s_r = find_read(L.tile_off, L.n, blockIdx.x);
    __syncthreads();
    const uint32_t r = s_r;
    const uint32_t len = L.len[r];
    const uint64_t s = (uint64_t)(blockIdx.x - L.tile_off[r]) * TILE_POS;
    const uint64_t p0 = s + 32ull * threadIdx.x;
    const uint64_t plen = ((uint64_t)len + PAD_BASES - 1) / PAD_BASES * PAD_BASES;
    if (p0 >= plen) return;
    uint64_t word = 0;
    uint32_t mask = 0;
    if (p0 < len) {
        const uint8_t *src = L.asc + L.off[r] + p0;
        const uint32_t *a4 = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
        uint32_t v[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = __ldg(a4 + i);
        const uint32_t nvalid = (uint32_t)min((uint64_t)32, (uint64_t)len - p0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t q = __funnelshift_r(v[i], v[i + 1], sh);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                uint32_t idx = i * 4 + b;
                uint32_t c = c_nt4[(q >> (8 * b)) & 0xff];
                if (idx >= nvalid) c = 4;
                word |= (uint64_t)(c & 3) << (62 - 2 * idx);
                mask |= (uint32_t)(c >> 2) << idx;
            }
        }
    } else {
        mask = 0xffffffffu;
    }
    const uint64_t wi = (L.poff[r] + p0) >> 5;
    L.w2[wi] = word;
    L.nm[wi] = mask;
}

// ---- K0h: homopolymer compression (ava-pb) -------------------------------------------------------
// One warp per read.  Step = a run of equal unambiguous bases, or a single ambiguous base
// (sketch.c: the HPC branch of mm_sketch). Emits the run base as ASCII, the position of the run's last
// base, and the number of steps.
__global__ void __launch_bounds__(256) hpc_compress_kernel(const uint8_t *asc, const uint64_t *off, const uint32_t *len0, uint32_t n,
                                                           uint8_t *casc, uint32_t *cend, uint32_t *clen) {
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const uint32_t lane = lane_id();
    const uint64_t o = off[r];
    const uint32_t len = len0[r];
    const char kAsc[5] = {'A', 'C', 'G', 'T', 'N'};
    uint32_t count = 0;
    for (uint32_t c0 = 0; c0 < len; c0 += 32) {
        const uint32_t i = c0 + lane;
        uint32_t c = 255, cn = 254;
        if (i < len) c = c_nt4[asc[o + i]];
        if (i + 1 < len) cn = c_nt4[asc[o + i + 1]];
        const bool is_end = (i < len) && (c >= 4 || cn != c);
        const uint32_t b = __ballot_sync(0xffffffffu, is_end);
        if (is_end) {
            const uint32_t idx = count + __popc(b & ((1u << lane) - 1));
            casc[o + idx] = (uint8_t)kAsc[c];
            cend[o + idx] = i;
        }
        count += __popc(b);
    }
    if (lane == 0) clen[r] = count;
}

// ---- K1: sketch ---------------------------------------------------------------------------------

// the same mix in 32-bit lanes: for 2k <= 32 every step only needs the low 2k bits of its operands
__device__ __forceinline__ uint32_t hash32_masked(uint32_t key, uint32_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

__device__ __forceinline__ uint64_t hash64_masked(uint64_t key, uint64_t mask) {
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

template <int W, typename XT>
__device__ __forceinline__ XT pick_x(const XT (&xs)[W + 1], int d) {
    XT v = xs[0];
#pragma unroll
    for (int i = 1; i <= W; ++i) v = d == i ? xs[i] : v;
    return v;
}

struct SketchArgs {
    ReadLayout L;
    const uint32_t *cend;     // HPC: original position of the last base of each step (indexed off[r] + t)
    const uint32_t *woff;     // EMIT: exclusive scan of wcnt (one entry per warp sub-tile)
    uint32_t *wcnt;           // COUNT: emissions per warp sub-tile
    uint64_t *ox, *oy;        // EMIT: minimizers, (rid, position) order (MODE 2: the staging arrays)
    uint32_t cap;             // MODE 2: staging capacity per warp sub-tile
    uint32_t *overflow;       // MODE 2: set when a sub-tile does not fit its slot
    int k;
};

// X32: without HPC the span byte is the constant k, so windows compare the 2k <= 31 bit hash alone (32-bit lanes).
// MODE 0: count the pushes of every warp sub-tile (wcnt).  MODE 1: emit at the scanned offsets (woff).
// MODE 2: single pass -- emit into a fixed-capacity staging slot per sub-tile and record wcnt; compact_kernel
// then packs the slots.  A sub-tile that would exceed its slot raises *overflow (the host falls back to 0 + 1).
template <bool HPC, int MODE, bool X32>
__global__ void __launch_bounds__(SK_THREADS) sketch_kernel(SketchArgs A) {
    constexpr bool EMIT = MODE >= 1;
    static_assert(!(HPC && X32), "HPC spans vary: 64-bit compare");
    constexpr int W = MAXW;
    using XT = typename std::conditional<X32, uint32_t, uint64_t>::type;
    constexpr XT XMAX = (XT)~(XT)0;
    __shared__ alignas(16) uint64_t sW[TILE_WORDS + 4];
    __shared__ alignas(16) uint32_t sN[TILE_WORDS + 4];
    __shared__ alignas(8) uint64_t bar;
    __shared__ uint32_t s_r;

    const ReadLayout &L = A.L;
    const uint32_t tile = blockIdx.x;
    if (threadIdx.x == 0) {
        s_r = find_read(L.tile_off, L.n, tile);
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (threadIdx.x < 4) { sW[TILE_WORDS + threadIdx.x] = 0; sN[TILE_WORDS + threadIdx.x] = 0; }
    __syncthreads();
    const uint32_t r = s_r;
    const int64_t len = L.len[r];
    const int64_t s = (int64_t)(tile - L.tile_off[r]) * TILE_POS;
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const int64_t ws = s + (int64_t)warp * WARP_POS;          // first step of this warp's sub-tile
    if (s >= len) {                                            // HPC: tile lies past the compressed length
        if (MODE != 1 && lane == 0) A.wcnt[tile * (SK_THREADS / 32) + warp] = 0;
        return;
    }
    if (threadIdx.x == 0) {
        const uint64_t g = (L.poff[r] + (uint64_t)s - PAD_BASES) >> 5;   // first word of the tile
        mbar_expect_tx(&bar, TILE_WORDS * 12);
        tma_load_1d(sW, L.w2 + g, TILE_WORDS * 8, &bar);
        tma_load_1d(sN, L.nm + g, TILE_WORDS * 4, &bar);
    }
    mbar_wait(&bar, 0);

    const int k = A.k;
    const uint64_t kmask = (1ULL << (2 * k)) - 1;
    const int LW = W + k;                                      // l is only compared with k, w+k-1, w+k
    const uint32_t lw_mask = LW >= 32 ? 0xffffffffu : ((1u << LW) - 1);
    const uint64_t rbase = L.off[r];
    const uint32_t rid = L.rid_map ? L.rid_map[r] : r;

    XT xprev = XMAX;
    uint32_t zprev = 0;
    uint32_t total = 0;
    const uint64_t obase = MODE == 1 ? (uint64_t)A.woff[tile * (SK_THREADS / 32) + warp]
                         : MODE == 2 ? (uint64_t)(tile * (SK_THREADS / 32) + warp) * A.cap : 0;
    const int n_chunks = WARP_POS / 32 + 1;                    // chunk 0 is warm-up (steps ws-32 .. ws-1)
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t p = ws + 32 * (c - 1) + lane;            // step handled by this lane
        if (ws + 32 * (int64_t)(c - 1) >= len) break;          // warp-uniform
        const int32_t q = (int32_t)(p - s) + PAD_BASES;        // index inside the staged tile (>= 96)
        XT x = XMAX;
        uint32_t z = 0;
        int32_t l = 0;
        if (p >= 0 && p < len) {
            // l_t from the ambiguity mask: bit j of mr <-> step p-j
            const int32_t st2 = q + 1 - 32;
            const uint32_t wi2 = st2 >> 5, sh2 = st2 & 31;
            const uint32_t m = __funnelshift_r(sN[wi2], sN[wi2 + 1], sh2);
            const uint32_t mr = __brev(m) & lw_mask;
            const int32_t lrun = mr ? (__ffs(mr) - 1) : LW;
            l = (int32_t)min((int64_t)lrun, p + 1);
            if (l >= k) {
                const int32_t b0 = q - k + 1;
                const uint32_t wi = b0 >> 5, sh = (b0 & 31) * 2;
                const uint64_t hi = sW[wi], lo = sW[wi + 1];
                const uint64_t v = sh ? (hi << sh) | (lo >> (64 - sh)) : hi;
                const uint64_t fwd = v >> (64 - 2 * k);
                uint64_t h;
                if (X32) {
                    const uint32_t f32 = (uint32_t)fwd;
                    uint32_t t = __brev(~f32);
                    t = ((t & 0xAAAAAAAAu) >> 1) | ((t & 0x55555555u) << 1);
                    const uint32_t rc = t >> (32 - 2 * k);
                    z = f32 < rc ? 0 : 1;
                    h = hash32_masked(z ? rc : f32, (uint32_t)kmask);
                } else {
                    uint64_t t = __brevll(~fwd);
                    t = ((t & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((t & 0x5555555555555555ULL) << 1);
                    const uint64_t rc = t >> (64 - 2 * k);
                    z = fwd < rc ? 0 : 1;
                    h = hash64_masked(z ? rc : fwd, kmask);
                }
                uint32_t span = k;
                if (HPC) {
                    const int64_t e1 = A.cend[rbase + p];
                    const int64_t e0 = p - k >= 0 ? (int64_t)A.cend[rbase + p - k] : -1;
                    span = (uint32_t)(e1 - e0);
                }
                if (X32) x = (XT)h;
                else if (span < 256) x = (XT)(h << 8 | span);
            }
        }
        // infos of steps t-d, d = 0..W, through warp shuffles (previous chunk supplies lanes < d)
        XT xs[W + 1];
        xs[0] = x;
#pragma unroll
        for (int d = 1; d <= W; ++d) {
            const int src = ((int)lane - d) & 31;
            const XT vc = __shfl_sync(0xffffffffu, x, src);
            const XT vp = __shfl_sync(0xffffffffu, xprev, src);
            xs[d] = (int)lane - d >= 0 ? vc : vp;
        }
        const uint32_t zb = __ballot_sync(0xffffffffu, z);
        const uint64_t zcomb = (uint64_t)zb << 32 | zprev;

        uint64_t em = 0;       // emitted offsets d, 4 bits each, in push order
        uint32_t cnt = 0;
        if (c > 0 && p < len) {
            // m[t-1]: rightmost minimum of steps t-W .. t-1;  m[t]: of steps t-W+1 .. t
            XT xm1 = xs[W]; int dm1 = W;
#pragma unroll
            for (int d = W - 1; d >= 1; --d) if (xs[d] <= xm1) { xm1 = xs[d]; dm1 = d; }
            XT xm = xs[W - 1]; int dm = W - 1;
#pragma unroll
            for (int d = W - 2; d >= 0; --d) if (xs[d] <= xm) { xm = xs[d]; dm = d; }
            if (l == W + k - 1 && xm1 != XMAX) {             // first full window: ties of the old minimum
#pragma unroll
                for (int d = W - 1; d >= 1; --d)
                    if (xs[d] == xm1 && d != dm1) { em |= (uint64_t)d << (4 * cnt); ++cnt; }
            }
            if (x <= xm1) {                                    // new minimum replaces the old one
                if (l >= W + k && xm1 != XMAX) { em |= (uint64_t)dm1 << (4 * cnt); ++cnt; }
            } else if (dm1 == W) {                             // old minimum left the window
                if (l >= W + k - 1 && xm1 != XMAX) { em |= (uint64_t)W << (4 * cnt); ++cnt; }
                if (l >= W + k - 1 && xm != XMAX) {
#pragma unroll
                    for (int d = W - 1; d >= 0; --d)
                        if (xs[d] == xm && d != dm) { em |= (uint64_t)d << (4 * cnt); ++cnt; }
                }
            }
            if (p == len - 1 && xm != XMAX) { em |= (uint64_t)dm << (4 * cnt); ++cnt; }   // final flush
        }
        // ordered output: exclusive prefix of cnt over the warp
        uint32_t pre = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, pre, o);
            if ((int)lane >= o) pre += v;
        }
        const uint32_t chunk_total = __shfl_sync(0xffffffffu, pre, 31);
        if (EMIT) {
            uint64_t o = obase + total + pre - cnt;
            for (uint32_t e = 0; e < cnt; ++e, ++o) {
                if (MODE == 2 && o - obase >= A.cap) break;            // slot full: overflow is flagged at the end
                const int d = (int)((em >> (4 * e)) & 15);
                const int64_t pt = p - d;
                const uint64_t zz = (zcomb >> (32 + (int)lane - d)) & 1;
                const uint64_t pos = HPC ? (uint64_t)A.cend[rbase + pt] : (uint64_t)pt;
                const XT xv = pick_x<W, XT>(xs, d);
                A.ox[o] = X32 ? ((uint64_t)xv << 8 | (uint64_t)k) : (uint64_t)xv;
                A.oy[o] = (uint64_t)rid << 32 | pos << 1 | zz;
            }
        }
        total += chunk_total;
        xprev = x;
        zprev = zb;
    }
    if (MODE != 1 && lane == 0) {
        A.wcnt[tile * (SK_THREADS / 32) + warp] = total;
        if (MODE == 2 && total > A.cap) *A.overflow = 1;
    }
}

// ---- K1, lane-serial form (no HPC, 2k <= 32) ---------------------------------------------------------------------------
// Replaces pack_kernel + sketch_kernel<stage> for ava-ont:
//   * the CTA stages the raw ASCII bases of its tile (8 192 steps of one read, 32 bases of history in front) with one TMA
//     bulk copy (cp.async.bulk, 16-byte aligned source) -- no packed copy of the reads in HBM;
//   * every lane runs mm_sketch's state machine SEQUENTIALLY over 32 consecutive steps, after a warm-up of w + k steps:
//     rolling forward / reverse k-mers, hash, the w + 1 most recent infos in registers, m[t-1] carried over from the
//     previous step.  The push rules are the ones of sketch_kernel above, evaluated on registers instead of warp
//     shuffles (3x fewer instructions per base);
//   * pushes are recorded as a bit mask over the positions [first step - w, first step + 32): mm_sketch pushes in
//     ascending position and never twice ("these two loops make sure the output is sorted", sketch.c), so the mask IS
//     the output order; the infos of those positions sit in shared memory for the emission;
//   * output goes to the per-sub-tile staging slots of the single-pass path (SK_CAP minimizers per 1 024 steps, wcnt[]),
//     which compact_kernel packs.  (An ordered single-kernel variant with a decoupled look-back over the tiles was
//     measured slower: tiles of one read set differ in length, and every tile waits for its predecessors.)
constexpr int SK2_S = 32;                         // steps per lane
constexpr int SK2_HALO = 32;                      // history staged in front of a tile (>= w + k)
constexpr int SK2_STRIDE = SK2_S + MAXW;          // infos kept per lane (odd: conflict-free)
constexpr int SK2_RAW = TILE_POS + SK2_HALO + 32; // staged ASCII bytes incl. alignment slack

struct Sketch2Args {
    const uint8_t *asc;        // ASCII bases of the resident reads
    const uint64_t *off;       // start of each entry in asc
    const uint32_t *len;       // length of each entry
    const uint32_t *rid_map;   // read id of each entry (NULL: entry index)
    const uint32_t *tile_off;  // n + 1: tiles before entry r
    const uint32_t *tile_read; // entry of each tile
    uint32_t tile0;            // first tile of this launch (one launch per staged chunk)
    uint64_t *ox, *oy;         // staging slots: SK_CAP entries per warp sub-tile
    uint32_t *wcnt;            // minimizers per warp sub-tile
    uint32_t *overflow;        // set when a sub-tile does not fit its slot
};

// ASCII base -> 0..3, or 4 for anything that is not ACGTU in either case (the seq_nt4_table of minimap2), without a
// table lookup: (b >> 1) & 3 maps A C G T/U to 0 1 3 2, the xor swaps 2 and 3; valid letters are 0x40-0x7f with
// (b & 31) in {1, 3, 7, 20, 21}.
__device__ __forceinline__ uint32_t nt4_code(uint32_t b) {
    const uint32_t c = (b >> 1) & 3u;
    const bool ok = (b & 0xc0u) == 0x40u && ((0x0030008au >> (b & 31u)) & 1u);
    return ok ? (c ^ (c >> 1)) : b < 4u ? b : 4u;                       // bytes 0..3 are accepted as codes, like the table
}

template <int K>
__global__ void __launch_bounds__(SK_THREADS) sketch2_kernel(Sketch2Args A) {
    constexpr int W = MAXW;
    static_assert(2 * K <= 32 && (K & 1), "32-bit k-mers, odd k");
    static_assert(SK2_HALO >= W + K, "history");
    constexpr uint32_t KMASK = 2 * K == 32 ? 0xffffffffu : ((1u << (2 * K)) - 1);
    constexpr uint32_t XMAX = 0xffffffffu;
    __shared__ alignas(16) uint8_t sRaw[SK2_RAW];                        // the tile's ASCII bases (+ history, + alignment slack)
    __shared__ uint32_t sInfo[SK_THREADS * SK2_STRIDE];                  // infos of every lane's steps, for the emission
    __shared__ alignas(8) uint64_t bar;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t tile = A.tile0 + blockIdx.x, r = A.tile_read[tile];
    const int32_t len = (int32_t)A.len[r];
    const int32_t s = (int32_t)(tile - A.tile_off[r]) * TILE_POS;       // first step of the tile; < len unless len == 0
    const uint32_t rid = A.rid_map ? A.rid_map[r] : r;
    const int32_t halo = s ? SK2_HALO : 0;
    const int32_t nb = max(0, min(TILE_POS, len - s)) + halo;            // staged steps: [s - halo, s - halo + nb)
    const uint8_t *src = A.asc + A.off[r] + s - halo;
    const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15);
    if (tid == 0 && nb > 0) {
        const uint32_t bytes = (shift + (uint32_t)nb + 15u) & ~15u;
        mbar_expect_tx(&bar, bytes);
        tma_load_1d(sRaw, src - shift, bytes, &bar);
    }
    if (nb > 0) mbar_wait(&bar, 0);

    // ---- lane-serial state machine, straight from the staged bytes -----------------------------------------------------
    const int32_t seg = s + (int32_t)(warp * WARP_POS + lane * SK2_S);     // first step of this lane
    uint64_t pm = 0, zm = 0;                                               // pushed positions / strands, bit i = step seg - W + i
    uint32_t *myInfo = sInfo + tid * SK2_STRIDE;
    if (seg < len) {
        const uint8_t *base = sRaw + shift + (seg - (s - halo));           // byte of step seg
        uint32_t fwd = 0, rev = 0;
        int32_t l = 0;
        uint32_t xs[W + 1];
#pragma unroll
        for (int d = 0; d <= W; ++d) xs[d] = XMAX;
        uint32_t zwarm = 0, zcur = 0;
        // warm-up: the W + K steps before seg (ambiguous where the read has not started)
        const bool have_prev = seg - (W + K) >= s - halo;                  // false only at the start of a read
#pragma unroll
        for (int i = -(W + K); i < 0; ++i) {
            const uint32_t c = have_prev ? nt4_code(base[i]) : 4u;
            fwd = ((fwd << 2) | (c & 3u)) & KMASK;
            rev = (rev >> 2) | ((3u - (c & 3u)) << (2 * (K - 1)));
            l = c > 3u ? 0 : l + 1;
            if (i >= -W) {                                                 // infos of steps seg - W .. seg - 1
                uint32_t x = XMAX;
                if (l >= K) {
                    const uint32_t z = fwd < rev ? 0u : 1u;
                    x = hash32_masked(z ? rev : fwd, KMASK);
                    zwarm |= z << (i + W);
                }
                myInfo[i + W] = x;
#pragma unroll
                for (int d = W; d >= 1; --d) xs[d] = xs[d - 1];
                xs[0] = x;
            }
        }
        // m[t-1] for t = seg: rightmost minimum of steps seg - W .. seg - 1 (xs[W-1] .. xs[0] before t enters)
        uint32_t xm1 = xs[W - 1];
        int dm1 = W;
#pragma unroll
        for (int d = W - 1; d >= 1; --d) if (xs[d - 1] <= xm1) { xm1 = xs[d - 1]; dm1 = d; }
        const int32_t n_steps = min(SK2_S, len - seg);
#pragma unroll 4
        for (int i = 0; i < SK2_S; ++i) {
            if (i >= n_steps) break;
            const uint32_t c = nt4_code(base[i]);
            fwd = ((fwd << 2) | (c & 3u)) & KMASK;
            rev = (rev >> 2) | ((3u - (c & 3u)) << (2 * (K - 1)));
            l = c > 3u ? 0 : min(l + 1, W + K);
            uint32_t x = XMAX;
            if (l >= K) {
                const uint32_t z = fwd < rev ? 0u : 1u;
                x = hash32_masked(z ? rev : fwd, KMASK);
                zcur |= z << i;
            }
            myInfo[i + W] = x;
#pragma unroll
            for (int d = W; d >= 1; --d) xs[d] = xs[d - 1];
            xs[0] = x;
            // m[t]: rightmost minimum of steps t-W+1 .. t
            uint32_t xm = xs[W - 1]; int dm = W - 1;
#pragma unroll
            for (int d = W - 2; d >= 0; --d) if (xs[d] <= xm) { xm = xs[d]; dm = d; }
            // pushes of this step as a mask over d: bit (W - d) = step t - d; OR-ed into pm at bit i (= step seg - W + i)
            uint32_t m = 0;
            const bool full1 = l >= W + K - 1;
            if (l == W + K - 1 && xm1 != XMAX) {                           // first full window: ties of the old minimum
#pragma unroll
                for (int d = W - 1; d >= 1; --d) m |= (uint32_t)(xs[d] == xm1) << (W - d);
                m &= ~(1u << (W - dm1));
            }
            if (x <= xm1) {                                                // new minimum replaces the old one
                if (l >= W + K && xm1 != XMAX) m |= 1u << (W - dm1);
            } else if (dm1 == W) {                                         // old minimum left the window
                if (full1 && xm1 != XMAX) m |= 1u;
                if (full1 && xm != XMAX) {
                    uint32_t ties = 0;
#pragma unroll
                    for (int d = W - 1; d >= 0; --d) ties |= (uint32_t)(xs[d] == xm) << (W - d);
                    m |= ties & ~(1u << (W - dm));
                }
            }
            if (seg + i == len - 1 && xm != XMAX) m |= 1u << (W - dm);     // final flush
            pm |= (uint64_t)m << i;
            xm1 = xm; dm1 = dm + 1;                                        // m[t] is m[t-1] of the next step
        }
        zm = (uint64_t)zwarm | (uint64_t)zcur << W;
    }
    // ---- emission into the sub-tile's staging slot: ascending position = mm_sketch's push order ----------------------
    const uint32_t cnt = (uint32_t)__popcll(pm);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += u; }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t sub = tile * (SK_THREADS / 32) + warp;
    if (lane == 0) {
        A.wcnt[sub] = total;
        if (total > SK_CAP) *A.overflow = 1;
    }
    uint32_t o = incl - cnt;
    uint64_t *ox = A.ox + (uint64_t)sub * SK_CAP, *oy = A.oy + (uint64_t)sub * SK_CAP;
    while (pm && o < SK_CAP) {
        const int b = __ffsll((long long)pm) - 1;
        pm &= pm - 1;
        const uint32_t x = myInfo[b];
        const uint64_t z = (zm >> b) & 1ull;
        const uint64_t pos = (uint64_t)(seg - W + b);
        ox[o] = (uint64_t)x << 8 | (uint64_t)K;
        oy[o] = (uint64_t)rid << 32 | pos << 1 | z;
        ++o;
    }
}

// MODE 2 follow-up: pack the staging slots into the contiguous (rid, position)-ordered arrays. One warp per sub-tile.
__global__ void __launch_bounds__(256) compact_kernel(const uint64_t *__restrict__ sx, const uint64_t *__restrict__ sy,
                                                      const uint32_t *__restrict__ wcnt, const uint32_t *__restrict__ woff, uint32_t nw,
                                                      uint32_t cap, uint64_t *__restrict__ ox, uint64_t *__restrict__ oy,
                                                      uint32_t *__restrict__ k32 = nullptr) {
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nw) return;
    const uint32_t n = wcnt[w], o = woff[w];
    const uint64_t b = (uint64_t)w * cap;
    for (uint32_t e = lane_id(); e < n; e += 32) {
        const uint64_t x = sx[b + e];
        ox[o + e] = x; oy[o + e] = sy[b + e];
        if (k32) k32[o + e] = (uint32_t)(x >> 8);      // the index's 32-bit sort keys (2k <= 32), saving a pass over x
    }
}

}  // namespace lrge
