"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same seeded
inputs.  Integer work is compared bit-exactly."""
import numpy as np
import pytest

import lrge_b200
import oracle
from lrge_b200 import readsim
from lrge_b200.readset import ReadSet

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = lrge_b200.Context(0)
    yield c
    c.close()


def small_sets(nt=300, nq=120, mean=4000, gsize=300_000, err=readsim.ONT, n_rate=0.0):
    g = readsim.genome(gsize)
    T = readsim.reads(g, nt, mean, err, readsim.SEED_TARGET, "t", n_rate=n_rate)
    Q = readsim.reads(g, nq, mean, err, readsim.SEED_QUERY, "q", n_rate=n_rate)
    return T, Q


def check_minimizers(ctx, rs, platform):
    x, y, cnt = ctx.dump_minimizers(rs, platform)
    want, wcnt = oracle.sketch_readset(rs, oracle.options(platform))
    assert np.array_equal(cnt, wcnt)
    assert x.shape[0] == want.shape[0]
    assert np.array_equal(x, want["x"])
    assert np.array_equal(y, want["y"])


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_sketch_parity(ctx, platform):
    T, _ = small_sets(err=readsim.ONT if platform == "ont" else readsim.HIFI)
    check_minimizers(ctx, T, platform)


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_sketch_parity_ambiguous_and_short(ctx, platform):
    # N bases, lower case, reads shorter than k / shorter than w+k, homopolymers, tile-boundary lengths
    T, _ = small_sets(nt=60, mean=3000, n_rate=0.003)
    rng = np.random.default_rng(3)
    recs = [(T.name(i), T.seq(i)) for i in range(T.n)]
    recs += [("s%d" % L, bytes(rng.choice(list(b"ACGT"), size=L).tolist())) for L in (1, 5, 14, 15, 19, 20, 23, 24, 25, 31, 32, 33, 100)]
    recs += [("hp", b"A" * 400 + b"C" * 300 + b"ACGT" * 50 + b"G" * 260 + b"TTTTACGTAGCATGCATGACTGACTAGCTAGCATG" * 4)]
    recs += [("lc", T.seq(0).lower()), ("nn", b"N" * 50 + T.seq(1)[:500] + b"NNN" + T.seq(1)[500:900] + b"N")]
    recs += [("tile%d" % L, bytes(rng.choice(list(b"ACGT"), size=L).tolist())) for L in (1023, 1024, 1025, 8191, 8192, 8193, 16384, 20000)]
    recs += [("rep", b"ACGTTGCA" * 300)]
    check_minimizers(ctx, ReadSet.from_records(recs), platform)
    # periodic reads emit a minimizer at (almost) every step: the single-pass staging slots overflow and the
    # library must fall back to its count + emit passes
    recs += [("dense2", b"AC" * 3000), ("dense3", b"ACG" * 2500), ("polyA", b"A" * 4000), ("polyAC", (b"AAA" + b"CC") * 1500)]
    check_minimizers(ctx, ReadSet.from_records(recs), platform)


def test_sketch_legacy_single_pass_path(monkeypatch):
    """ava-ont defaults to the lane-serial single-kernel sketch; LRGE_SKETCH_LEGACY=1 selects the warp-shuffle kernel with
    staging slots + compaction (what ava-pb always uses)."""
    monkeypatch.setenv("LRGE_SKETCH_LEGACY", "1")
    c1 = lrge_b200.Context(0)
    try:
        T, _ = small_sets(n_rate=0.002)
        check_minimizers(c1, T, "ont")
    finally:
        c1.close()


def test_sketch_long_reads_and_many_tiles(ctx):
    """Reads spanning many 8 192-step tiles (look-back scan across tiles and reads), with N runs at tile borders."""
    rng = np.random.default_rng(5)
    recs = []
    for i, L in enumerate((70_000, 8192 * 3, 8192 * 3 + 1, 40_001, 123, 8192, 8191, 16_385, 1)):
        seq = bytearray(rng.choice(list(b"ACGT"), size=L).tolist())
        for at in (8190, 8192, 16383, 16384, 24576):
            if at + 3 < L:
                seq[at:at + 3] = b"NNN"
        recs.append(("long%d" % i, bytes(seq)))
    check_minimizers(ctx, ReadSet.from_records(recs), "ont")


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_sketch_two_pass_path(monkeypatch, platform):
    monkeypatch.setenv("LRGE_SKETCH_TWO_PASS", "1")
    c1 = lrge_b200.Context(0)
    try:
        T, _ = small_sets(err=readsim.ONT if platform == "ont" else readsim.HIFI)
        check_minimizers(c1, T, platform)
    finally:
        c1.close()


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_mid_occ_parity(ctx, platform):
    T, _ = small_sets(nt=400, gsize=100_000, err=readsim.ONT if platform == "ont" else readsim.HIFI)
    opt = oracle.options(platform)
    idx = oracle.Index(T, opt, threads=4)
    mid, nkeys = ctx.dump_mid_occ(T, platform)
    assert nkeys == idx.n_keys
    assert mid == opt.mid_occ


def test_twoset_counts_parity_ont(ctx):
    T, Q = small_sets()
    res = ctx.twoset_counts(T, Q, "ont")
    want, had, st = oracle.twoset_counts(T, Q, "ont", threads=8)
    assert res.stats["mid_occ"] == st["mid_occ"]
    assert res.stats["n_target_minimizers"] == st["n_t_minimizers"]
    assert res.stats["n_query_minimizers"] == st["n_q_minimizers"]
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want)
    assert np.array_equal(res.had_mapping, had)
    assert res.stats["no_mapping_count"] == st["no_mapping_count"]
    assert want.sum() > 0


def test_twoset_counts_parity_c1(ctx):
    _, platform, T, Q = readsim.make_config("C1")
    res = ctx.twoset_counts(T, Q, platform)
    want, had, st = oracle.twoset_counts(T, Q, platform, threads=8)
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want)
    assert np.array_equal(res.had_mapping, had)


def test_twoset_counts_parity_pb(ctx):
    T, Q = small_sets(nt=200, nq=80, mean=5000, gsize=200_000, err=readsim.HIFI)
    res = ctx.twoset_counts(T, Q, "pb")
    want, had, st = oracle.twoset_counts(T, Q, "pb", threads=8)
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want)
    assert np.array_equal(res.had_mapping, had)


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_ava_counts_parity(ctx, platform):
    g = readsim.genome(200_000)
    R = readsim.reads(g, 250, 4000, readsim.ONT if platform == "ont" else readsim.HIFI, readsim.SEED_AVA, "r")
    res = ctx.ava_counts(R, platform)
    want, st = oracle.ava_counts(R, platform, threads=8)
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want)
    assert res.stats["n_overlaps"] == st["n_overlaps"]
    assert res.stats["no_mapping_count"] == st["no_mapping_count"]


def test_remove_internal_parity(ctx):
    T, Q = small_sets(nt=300, nq=150, mean=3000)
    for ratio in (0.2, 0.05):
        res = ctx.twoset_counts(T, Q, "ont", remove_internal=True, max_overhang_ratio=ratio)
        want, had, _ = oracle.twoset_counts(T, Q, "ont", remove_internal=True, max_overhang_ratio=ratio, threads=8)
        assert np.array_equal(res.counts, want)
        assert np.array_equal(res.had_mapping, had)
    g = readsim.genome(150_000)
    R = readsim.reads(g, 200, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    res = ctx.ava_counts(R, "ont", remove_internal=True)
    want, _ = oracle.ava_counts(R, "ont", remove_internal=True, threads=8)
    assert np.array_equal(res.counts, want)


def test_use_min_ref_inverse_parity(ctx):
    T, Q = small_sets(nt=300, nq=100)
    assert T.total_bases > Q.total_bases
    for F in (False, True):
        res = ctx.twoset_counts(T, Q, "ont", use_min_ref=True, remove_internal=F)
        want, st = oracle.twoset_inverse_counts(T, Q, "ont", remove_internal=F, threads=8)
        assert np.array_equal(res.counts, want)
        assert res.stats["no_mapping_count"] == st["no_mapping_count"]


def test_target_sharding_is_additive(ctx):
    T, Q = small_sets()
    full = ctx.twoset_counts(T, Q, "ont")
    for world in (2, 3):
        parts = [ctx.twoset_counts(T, Q, "ont", shard_index=r, shard_count=world) for r in range(world)]
        assert np.array_equal(sum(p.counts for p in parts), full.counts)
        assert all(p.stats["mid_occ"] == full.stats["mid_occ"] for p in parts)
        assert sum(p.stats["n_anchors"] for p in parts) == full.stats["n_anchors"]


def test_shared_and_duplicate_names(ctx):
    # a query that is also a target (NO_DIAG path) and two targets sharing a name (HashSet<name> dedupe)
    T, Q = small_sets(nt=120, nq=40, gsize=100_000)
    recs_t = [(T.name(i), T.seq(i)) for i in range(T.n)]
    recs_q = [(Q.name(i), Q.seq(i)) for i in range(Q.n)]
    recs_q[0] = recs_t[5]
    recs_t[7] = (recs_t[6][0], recs_t[7][1])
    T2, Q2 = ReadSet.from_records(recs_t), ReadSet.from_records(recs_q)
    res = ctx.twoset_counts(T2, Q2, "ont")
    want, had, _ = oracle.twoset_counts(T2, Q2, "ont", threads=4)
    assert np.array_equal(res.counts, want)
    assert np.array_equal(res.had_mapping, had)


def test_error_paths(ctx):
    T, Q = small_sets(nt=20, nq=5, gsize=50_000)
    dup = ReadSet.from_records([(T.name(0), T.seq(0)), (T.name(0), T.seq(1)), (T.name(2), T.seq(2))])
    with pytest.raises(lrge_b200.LrgeError) as e:
        ctx.ava_counts(dup, "ont")
    assert e.value.kind == "DuplicateReadIdentifier"
    empty = ReadSet.from_records([("a", b"ACGT" * 100), ("b", b"")])
    with pytest.raises(lrge_b200.LrgeError) as e:
        ctx.twoset_counts(T, empty, "ont")
    assert e.value.kind == "MapError"
    # an empty read in the INDEXED set is not an error (minimap2's index build skips it): it still counts as a target read
    T, Q = small_sets(nt=60, nq=20, gsize=60_000)
    recs = [(T.name(i), T.seq(i)) for i in range(T.n)]
    recs.insert(7, ("empty_target", b""))
    for platform in ("ont", "pb"):
        T2 = ReadSet.from_records(recs)
        res = ctx.twoset_counts(T2, Q, platform)
        want, had, _ = oracle.twoset_counts(T2, Q, platform, threads=4)
        assert np.array_equal(res.counts, want) and np.array_equal(res.had_mapping, had)


def test_strategy_estimates_match_oracle(ctx):
    T, Q = small_sets()
    s = lrge_b200.TwoSetStrategy(T, Q, ctx=ctx)
    r = s.estimate(True, lrge_b200.LOWER_QUANTILE, lrge_b200.UPPER_QUANTILE)
    want, had, _ = oracle.twoset_counts(T, Q, "ont", threads=8)
    avg = np.float32(T.total_bases) / np.float32(T.n)
    est = [oracle.per_read_estimate(int(l), avg, T.n, int(c), 100) for l, c in zip(Q.lengths(), want)]
    lo, md, up = oracle.estimate(est, True, 0.15, 0.65)
    assert (r.lower, r.estimate, r.upper) == (lo, md, up)
    assert r.no_mapping_count == int((had == 0).sum())
    assert abs(r.estimate - 300_000) / 300_000 < 0.5


_V1_SNIPPET = """
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import lrge_b200
from lrge_b200 import readsim
from test_gpu_parity import small_sets
T, Q = small_sets(nt=400, nq=150, mean=5000, gsize=250_000, err=readsim.ONT if {platform!r} == "ont" else readsim.HIFI)
with lrge_b200.Context(0) as c:
    r = c.twoset_counts(T, Q, {platform!r}, remove_internal={F!r})
np.savez({out!r}, counts=r.counts, had=r.had_mapping)
"""


@pytest.mark.parametrize("platform,F", [("ont", True), ("ont", False), ("pb", False), ("pb", True)])
def test_chain_kernels_agree(ctx, tmp_path, platform, F):
    """The product's kernels (fused shared-memory path; early-exit counting when F is off, full DP + backtrack when on)
    against the literal thread-per-pair kernel, which only exists in the test build liblrge_b200_v1check.so
    (tests/native/kernels_chain_v1.cuh; always full DP + backtrack), run in a subprocess."""
    import os, subprocess, sys
    from lrge_b200.readset import csrc_dir
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    T, Q = small_sets(nt=400, nq=150, mean=5000, gsize=250_000, err=readsim.ONT if platform == "ont" else readsim.HIFI)
    a = ctx.twoset_counts(T, Q, platform, remove_internal=F)
    out = str(tmp_path / "v1.npz")
    env = dict(os.environ, LRGE_B200_LIB=os.path.join(csrc_dir(), "liblrge_b200_v1check.so"), LRGE_CHAIN_V1="1", LRGE_FORCE_GLOBAL="1")
    subprocess.run([sys.executable, "-c", _V1_SNIPPET.format(root=root, platform=platform, F=F, out=out)], check=True, env=env, timeout=600)
    b = np.load(out)
    assert np.array_equal(a.counts, b["counts"])
    assert np.array_equal(a.had_mapping, b["had"])
    assert a.counts.sum() > 0


def test_long_span_hpc_group_takes_literal_path(ctx):
    """ava-pb reads with long homopolymers: minimizer spans above (min_sc-1)/(min_cnt-1) disable the counting
    shortcut for their groups; results must still equal the oracle."""
    rng = np.random.default_rng(11)
    def hp_seq(n):
        out = []
        for _ in range(n):
            out.append(bytes([rng.choice(list(b"ACGT"))]) * int(rng.choice([1, 1, 1, 2, 3, 9, 14])))
        return b"".join(out)
    g = hp_seq(6000)
    recs_t = [("t%d" % i, g[s:s + 6000]) for i, s in enumerate(range(0, len(g) - 6000, 700))]
    recs_q = [("q%d" % i, g[s:s + 5000]) for i, s in enumerate(range(350, len(g) - 5000, 900))]
    T, Q = ReadSet.from_records(recs_t), ReadSet.from_records(recs_q)
    res = ctx.twoset_counts(T, Q, "pb")
    want, had, _ = oracle.twoset_counts(T, Q, "pb", threads=4)
    assert np.array_equal(res.counts, want)
    assert np.array_equal(res.had_mapping, had)
    assert want.sum() > 0


def test_query_batching(ctx, monkeypatch):
    """Global-memory path: anchors are processed in query batches sized from device memory; force many small batches."""
    T, Q = small_sets()
    full = ctx.twoset_counts(T, Q, "ont")
    monkeypatch.setenv("LRGE_FORCE_GLOBAL", "1")
    monkeypatch.setenv("LRGE_MAX_ANCHORS", "6000")
    c1 = lrge_b200.Context(0)
    try:
        b = c1.twoset_counts(T, Q, "ont")
        g = readsim.genome(200_000)
        R = readsim.reads(g, 250, 4000, readsim.ONT, readsim.SEED_AVA, "r")
        av = c1.ava_counts(R, "ont")
    finally:
        c1.close()
    assert b.stats["n_batches"] > 5
    assert np.array_equal(full.counts, b.counts)
    assert np.array_equal(full.had_mapping, b.had_mapping)
    want, _ = oracle.ava_counts(R, "ont", threads=8)
    assert av.stats["n_batches"] > 5
    assert np.array_equal(av.counts, want)


def _reg_rows(regs, qidx=None):
    """Rows as tuples incl. the PAF tags dv (bit pattern of the f32) and rl; product records or oracle regs."""
    if qidx is None:
        return list(zip(regs["query"].tolist(), regs["target"].tolist(), regs["rev"].tolist(), regs["qs"].tolist(), regs["qe"].tolist(),
                        regs["ts"].tolist(), regs["te"].tolist(), regs["score"].tolist(), regs["cnt"].tolist(), regs["mlen"].tolist(),
                        regs["blen"].tolist(), regs["dv"].view(np.uint32).tolist(), regs["rl"].tolist()))
    return list(zip(qidx.tolist(), regs["rid"].tolist(), regs["rev"].tolist(), regs["qs"].tolist(), regs["qe"].tolist(),
                    regs["rs"].tolist(), regs["re"].tolist(), regs["score"].tolist(), regs["cnt"].tolist(), regs["mlen"].tolist(),
                    regs["blen"].tolist(), regs["div"].view(np.uint32).tolist(), regs["rl"].tolist()))


def _reg_set(regs, qidx=None):
    return sorted(_reg_rows(regs, qidx))


def _assert_row_order(ov, regs, qidx):
    """Product rows come sorted by query and, within a query, in mm_map's order (mm_gen_regs: score descending, hash
    tie-break): compare the whole sequence with the oracle's, query by query (the oracle's flattened list is in
    thread-completion order of the queries, its `order` field is the row's position inside its query)."""
    key = np.lexsort((regs["order"], qidx))
    want = [_reg_rows(regs, qidx)[i] for i in key.tolist()]
    got = _reg_rows(ov)
    assert [r[0] for r in got] == sorted(r[0] for r in got)
    assert got == want


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_overlap_records_parity(ctx, platform):
    """Every kept chain (PAF row: coordinates, score, cm, match/block length) against the oracle's mm_gen_regs output."""
    T, Q = small_sets(nt=300, nq=120, mean=4000, gsize=250_000, err=readsim.ONT if platform == "ont" else readsim.HIFI)
    ov, cnt = ctx.twoset_overlaps(T, Q, platform)
    regs, qidx, _ = oracle.twoset_regs(T, Q, platform, threads=8, tie_mode=oracle.TIE_STABLE)
    assert ov.shape[0] == regs.shape[0] > 0
    assert _reg_set(ov) == _reg_set(regs, qidx)          # every field incl. dv (mm_est_err) and rl (rep_len)
    _assert_row_order(ov, regs, qidx)
    assert (ov["dv"] >= 0).all() and ov["dv"].max() > 0.001
    plain = ctx.twoset_counts(T, Q, platform)
    assert np.array_equal(cnt.counts, plain.counts) and np.array_equal(cnt.had_mapping, plain.had_mapping)
    # the literal (unstable ksort) tie order gives the same rows here
    regs2, qidx2, _ = oracle.twoset_regs(T, Q, platform, threads=8, tie_mode=oracle.TIE_KSORT)
    assert _reg_set(regs2, qidx2) == _reg_set(regs, qidx)
    # a small buffer reports how many records exist
    small, _ = ctx.twoset_overlaps(T, Q, platform, cap=8)
    assert small.shape[0] == ov.shape[0]
    line = lrge_b200.paf_line(ov[0], Q, T)
    assert line.count(b"\t") == 16 and b"tp:A:S" in line and b"dv:f:0." in line


def test_ava_overlap_records_parity(ctx):
    g = readsim.genome(150_000)
    R = readsim.reads(g, 200, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    ov, cnt = ctx.ava_overlaps(R, "ont")
    regs, qidx, _ = oracle.twoset_regs(R, R, "ont", threads=8, tie_mode=oracle.TIE_STABLE, dual=False)
    assert _reg_set(ov) == _reg_set(regs, qidx)
    _assert_row_order(ov, regs, qidx)                    # incl. the self-overlap rows (SEED_SELF enters the row hash)
    want, _ = oracle.ava_counts(R, "ont", threads=8)
    assert np.array_equal(cnt.counts, want)


def test_mid_occ_overflow_bin(ctx):
    """Keys occurring >= 4095 times leave the occurrence histogram's exact bins: mid_occ must still be the exact
    order statistic (mm_idx_cal_max_occ)."""
    rng = np.random.default_rng(21)
    unit = bytes(rng.choice(list(b"ACGT"), size=400).tolist())
    recs = [("t%05d" % i, unit) for i in range(4200)]
    recs += [("u%03d" % i, bytes(rng.choice(list(b"ACGT"), size=700).tolist())) for i in range(30)]
    T = ReadSet.from_records(recs)
    opt = oracle.options("ont")
    idx = oracle.Index(T, opt, threads=8)
    mid, nkeys = ctx.dump_mid_occ(T, "ont")
    assert nkeys == idx.n_keys
    assert mid == opt.mid_occ and mid > 4095


def test_ava_sharding_is_additive(ctx):
    g = readsim.genome(150_000)
    R = readsim.reads(g, 200, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    full = ctx.ava_counts(R, "ont")
    parts = [ctx.ava_counts(R, "ont", shard_index=r, shard_count=3) for r in range(3)]
    assert np.array_equal(sum(p.counts for p in parts), full.counts)


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_shard_block_staging_and_sketch(ctx, platform):
    """A rank stages and sketches only its contiguous block of targets; read ids stay those of the whole set."""
    T, _ = small_sets(nt=101, err=readsim.ONT if platform == "ont" else readsim.HIFI)
    want, wcnt = oracle.sketch_readset(T, oracle.options(platform))
    starts = np.concatenate([[0], np.cumsum(wcnt.astype(np.int64))])
    world = 3
    seen = 0
    for rank in range(world):
        lo, hi = T.n * rank // world, T.n * (rank + 1) // world
        x, y, cnt = ctx.dump_minimizers_block(T, lo, hi - lo, platform)
        a, b = int(starts[lo]), int(starts[hi])
        assert np.array_equal(cnt, wcnt[lo:hi])
        assert np.array_equal(x, want["x"][a:b]) and np.array_equal(y, want["y"][a:b])
        seen += x.shape[0]
    assert seen == want.shape[0]


# ---- the fused shared-memory path (default) vs the global-memory path vs the oracle -----------------------------

def _ctx_with(monkeypatch, **env):
    for k, v in env.items():
        monkeypatch.setenv(k, str(v))
    return lrge_b200.Context(0)


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_global_path_parity(ctx, monkeypatch, platform):
    """LRGE_FORCE_GLOBAL=1 replaces the per-query shared-memory sort by the device-wide radix sort + group kernels:
    counts, had_mapping, anchor totals and overlap records must equal the default path's and the oracle's."""
    T, Q = small_sets(nt=300, nq=120, mean=4000, gsize=250_000, err=readsim.ONT if platform == "ont" else readsim.HIFI)
    want, had, st = oracle.twoset_counts(T, Q, platform, threads=8)
    a = ctx.twoset_counts(T, Q, platform)
    ov_a, _ = ctx.twoset_overlaps(T, Q, platform)
    g = _ctx_with(monkeypatch, LRGE_FORCE_GLOBAL=1)
    try:
        b = g.twoset_counts(T, Q, platform)
        ov_b, _ = g.twoset_overlaps(T, Q, platform)
        R = readsim.reads(readsim.genome(150_000), 200, 3000, readsim.ONT if platform == "ont" else readsim.HIFI, readsim.SEED_AVA, "r")
        av_b = g.ava_counts(R, platform)
    finally:
        g.close()
    for r in (a, b):
        assert np.array_equal(r.counts, want) and np.array_equal(r.had_mapping, had)
        assert r.stats["n_anchors"] == st["n_anchors"]
    assert a.stats["n_deferred"] == 0 and a.stats["n_pairs"] == b.stats["n_pairs"]
    assert _reg_set(ov_a) == _reg_set(ov_b)
    av_want, _ = oracle.ava_counts(R, platform, threads=8)
    assert np.array_equal(av_b.counts, av_want)
    assert np.array_equal(ctx.ava_counts(R, platform).counts, av_want)


@pytest.mark.parametrize("cap", [2048, 256])
def test_fused_rounds_and_deferral(ctx, monkeypatch, cap):
    """LRGE_MAP_CAP shrinks the shared-memory capacity of the per-query sort kernel: queries are sorted in several rounds
    over disjoint target ranges, and a query holding one target bin above the capacity is deferred to the device-wide
    sort.  Results must not change."""
    T, Q = small_sets()
    want, had, st = oracle.twoset_counts(T, Q, "ont", threads=8)
    g = readsim.genome(200_000)
    R = readsim.reads(g, 250, 4000, readsim.HIFI, readsim.SEED_AVA, "r")
    av_want, av_st = oracle.ava_counts(R, "pb", threads=8)
    regs, qidx, _ = oracle.twoset_regs(T, Q, "ont", threads=8, tie_mode=oracle.TIE_STABLE)
    c1 = _ctx_with(monkeypatch, LRGE_MAP_CAP=cap)
    try:
        b = c1.twoset_counts(T, Q, "ont")
        bf = c1.twoset_counts(T, Q, "ont", remove_internal=True)
        av = c1.ava_counts(R, "pb")
        ov, _ = c1.twoset_overlaps(T, Q, "ont")
    finally:
        c1.close()
    assert np.array_equal(b.counts, want) and np.array_equal(b.had_mapping, had)
    assert b.stats["n_anchors"] == st["n_anchors"]
    assert b.stats["n_rounds"] > Q.n                        # several rounds per query
    if cap == 256:
        assert b.stats["n_deferred"] > 0                    # 4 kb overlaps hold > 256 anchors of one target
    else:
        assert b.stats["n_deferred"] == 0
    assert np.array_equal(bf.counts, ctx.twoset_counts(T, Q, "ont", remove_internal=True).counts)
    assert np.array_equal(av.counts, av_want) and av.stats["n_anchors"] == av_st["n_anchors"]
    assert _reg_set(ov) == _reg_set(regs, qidx)


def test_host_buffer_chunked_index_build(ctx, monkeypatch):
    """Host-buffer calls copy the indexed set in chunks; with LRGE_INDEX_HALVES=1 (an experiment kept off by default:
    measured slower on B200) the sketch, the compaction and the index sort of an ava-ont set run per HALF of the chunks and
    the two sorted runs are merged (sketch_index_halves).  LRGE_MIN_CHUNK makes small sets arrive in chunks.  The index (mid_occ, keys, hit order through the anchors) and every result must equal the
    unchunked path and the oracle -- including the overlap records, whose tie order depends on the hit-list order."""
    T, Q = small_sets()
    want, had, st = oracle.twoset_counts(T, Q, "ont", threads=8)
    g = readsim.genome(150_000)
    R = readsim.reads(g, 300, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    av_want, av_st = oracle.ava_counts(R, "ont", threads=8)
    c1 = _ctx_with(monkeypatch, LRGE_MIN_CHUNK=50_000, LRGE_INDEX_HALVES=1)
    try:
        a = c1.twoset_counts(T, Q, "ont")
        a2 = c1.twoset_counts(T, Q, "ont", remove_internal=True)
        av = c1.ava_counts(R, "ont")                     # the indexed set's minimizers (x, y, per-read offsets) are also the queries
    finally:
        c1.close()
    base = ctx.twoset_counts(T, Q, "ont")
    assert a.stats["n_index_runs"] == 2 and av.stats["n_index_runs"] == 2      # the chunked path did run
    assert base.stats["n_index_runs"] == 1
    assert np.array_equal(a.counts, want) and np.array_equal(a.had_mapping, had)
    for k in ("n_anchors", "mid_occ", "n_keys", "n_target_minimizers", "n_query_minimizers", "n_seeds", "n_hits", "n_overlaps"):
        assert a.stats[k] == base.stats[k], k
    assert a.stats["n_anchors"] == st["n_anchors"] and a.stats["mid_occ"] == st["mid_occ"]
    assert np.array_equal(a2.counts, ctx.twoset_counts(T, Q, "ont", remove_internal=True).counts)
    assert np.array_equal(av.counts, av_want) and av.stats["n_anchors"] == av_st["n_anchors"]


def _branch_report(st):
    return {k: int(st[k]) for k in ("mid_occ", "n_anchors", "n_tie_x_queries", "n_tie_x_anchors", "n_mz_flt_deleted",
                                    "n_occ_filtered", "n_max_iter_clamp", "n_skip_break")}


@pytest.mark.parametrize("platform", ["ont", "pb"])
def test_adversarial_repeats_parity(ctx, monkeypatch, platform):
    """Reads drawn from a genome with 2-5 kb tandem repeats, low-complexity tracts and dispersed duplications shared by
    T and Q.  These reach the order-dependent branches uniform genomes never do (reported below): mm_seed_mz_flt
    deletions, n > mid_occ seeds, the max_chain_iter clamp, the max_chain_skip break, equal-x anchors.
    Counts and had_mapping are compared with the oracle in its literal (ksort) tie order; overlap records with the
    oracle's stable tie order, which is what the CUDA sort implements."""
    import adversarial
    if platform == "ont":
        T, Q = adversarial.repeat_sets(errors=readsim.ONT)
    else:
        # HPC + HiFi: low-complexity tracts give groups of 10^4-10^5 anchors (mid_occ > 500); kept small because every
        # chain of such a group is extracted one by one in records mode
        T, Q = adversarial.repeat_sets(errors=readsim.HIFI, nt=60, nq=16, mean=3000, size=60_000, n_tandem=2, n_lowc=1, n_dup=1)
    want, had, st = oracle.twoset_counts(T, Q, platform, threads=8, tie_mode=oracle.TIE_KSORT)
    rep = _branch_report(st)
    print("adversarial", platform, rep)
    assert rep["n_tie_x_anchors"] > 1000 and rep["n_skip_break"] > 1000 and rep["n_occ_filtered"] > 100
    if platform == "ont":
        assert rep["n_mz_flt_deleted"] > 100
    else:
        assert rep["n_max_iter_clamp"] > 1000
    res = ctx.twoset_counts(T, Q, platform)
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want) and np.array_equal(res.had_mapping, had)
    resF = ctx.twoset_counts(T, Q, platform, remove_internal=True)
    wantF, hadF, _ = oracle.twoset_counts(T, Q, platform, remove_internal=True, threads=8, tie_mode=oracle.TIE_STABLE)
    assert np.array_equal(resF.counts, wantF) and np.array_equal(resF.had_mapping, hadF)
    ov, _ = ctx.twoset_overlaps(T, Q, platform)
    regs, qidx, _ = oracle.twoset_regs(T, Q, platform, threads=8, tie_mode=oracle.TIE_STABLE)
    assert _reg_set(ov) == _reg_set(regs, qidx)
    # the same inputs through the global-memory path and with the fused kernel forced into rounds
    for env in (dict(LRGE_FORCE_GLOBAL=1), dict(LRGE_MAP_CAP=1024)):
        c1 = _ctx_with(monkeypatch, **env)
        try:
            r1 = c1.twoset_counts(T, Q, platform)
        finally:
            c1.close()
        for k in env:
            monkeypatch.delenv(k)
        assert np.array_equal(r1.counts, want) and np.array_equal(r1.had_mapping, had)


def test_adversarial_repeats_ava(ctx):
    import adversarial
    g = adversarial.repeat_genome(120_000, seed=9)
    R = readsim.reads(g, 220, 3500, readsim.ONT, readsim.SEED_AVA, "r")
    want, st = oracle.ava_counts(R, "ont", threads=8, tie_mode=oracle.TIE_KSORT)
    print("adversarial ava", _branch_report(st))
    res = ctx.ava_counts(R, "ont")
    assert res.stats["n_anchors"] == st["n_anchors"]
    assert np.array_equal(res.counts, want)


def test_toy_real_reads(ctx):
    """The reference's only real-data fixture (lrge/tests/data/toy.bam: 500 ONT reads of M. tuberculosis, decoded by
    tests/golden/make_toy_fasta.py): all-vs-all and a two-set split, counts and records against the oracle."""
    import os
    R = lrge_b200.read_fastx(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "toy_reads.fa.gz"))
    assert R.n == 500 and R.total_bases == 1272286
    want, st = oracle.ava_counts(R, "ont", threads=8)
    print("toy ava", _branch_report(st), "overlaps", st["n_overlaps"])
    res = ctx.ava_counts(R, "ont")
    assert res.stats["n_anchors"] == st["n_anchors"] and res.stats["mid_occ"] == st["mid_occ"]
    assert np.array_equal(res.counts, want) and res.stats["n_overlaps"] == st["n_overlaps"] > 1000
    T, Q = R.subset(range(300, 500)), R.subset(range(0, 300))
    want2, had2, st2 = oracle.twoset_counts(T, Q, "ont", threads=8)
    r2 = ctx.twoset_counts(T, Q, "ont")
    assert np.array_equal(r2.counts, want2) and np.array_equal(r2.had_mapping, had2) and want2.sum() > 500
    ov, _ = ctx.twoset_overlaps(T, Q, "ont")
    regs, qidx, _ = oracle.twoset_regs(T, Q, "ont", threads=8, tie_mode=oracle.TIE_STABLE)
    assert _reg_set(ov) == _reg_set(regs, qidx)
    # the estimate the CLI would print for this split (twoset.rs:311-317, estimate.rs:55-77)
    s = lrge_b200.TwoSetStrategy(T, Q, ctx=ctx)
    r = s.estimate(True, lrge_b200.LOWER_QUANTILE, lrge_b200.UPPER_QUANTILE)
    avg = np.float32(T.total_bases) / np.float32(T.n)
    est = [oracle.per_read_estimate(int(l), avg, T.n, int(c), 100) for l, c in zip(Q.lengths(), want2)]
    assert (r.lower, r.estimate, r.upper) == oracle.estimate(est, True, 0.15, 0.65)
