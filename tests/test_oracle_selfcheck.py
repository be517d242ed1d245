"""Self-consistency of the (unpinned) oracle, SURVEY.md A.9: properties that hold for minimap2's algorithm whatever
the implementation, checked on the restatement."""
import numpy as np
import pytest

import oracle
from lrge_b200 import readsim
from lrge_b200.readset import ReadSet

COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(COMP)[::-1]


def hash64(key, mask):
    key = (~key + (key << 21)) & mask
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & mask
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & mask
    key = key ^ key >> 28
    key = (key + (key << 31)) & mask
    return key


def brute_minimizers(seq: bytes, w: int, k: int):
    """(w,k)-minimizers from the definition (N-free, odd k): a k-mer is reported iff its hash is the minimum of at
    least one full window of w consecutive k-mers (ties all reported); reads with fewer than w k-mers report the
    minimum of what they have (mm_sketch's final flush)."""
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    mask = (1 << 2 * k) - 1
    hs = []
    for i in range(k - 1, len(seq)):
        f = r = 0
        for j in range(i - k + 1, i + 1):
            c = code[seq[j]]
            f = (f << 2 | c) & mask
        for j in range(i, i - k, -1):
            r = (r << 2 | (3 - code[seq[j]])) & mask
        z = 0 if f < r else 1
        hs.append((hash64(min(f, r), mask) << 8 | k, i << 1 | z))
    out = set()
    if len(hs) < w:
        return out, hs
    for s in range(len(hs) - w + 1):
        win = hs[s:s + w]
        m = min(x for x, _ in win)
        out.update((x, y) for x, y in win if x == m)
    return out, hs


@pytest.mark.parametrize("k,w", [(15, 5), (19, 5), (11, 3)])
def test_sketch_matches_window_definition(k, w):
    rng = np.random.default_rng(k * 100 + w)
    for L in (w + k - 1, w + k, 60, 200, 1000):
        seq = bytes(rng.choice(list(b"ACGT"), size=L).tolist())
        got = oracle.sketch(seq, w, k, 0, False)
        want, _ = brute_minimizers(seq, w, k)
        assert set(zip(got["x"].tolist(), got["y"].tolist())) == want, (k, w, L)
        assert np.all(np.diff(got["y"].astype(np.int64)) > 0)   # emitted once each, in position order


def test_sketch_strand_symmetry():
    rng = np.random.default_rng(5)
    seq = bytes(rng.choice(list(b"ACGT"), size=3000).tolist())
    a, b = oracle.sketch(seq, 5, 15), oracle.sketch(revcomp(seq), 5, 15)
    # interior windows see the same k-mers from the other strand: same hash multiset away from the ends
    ha, hb = sorted(a["x"].tolist()), sorted(b["x"].tolist())
    assert len(set(ha) ^ set(hb)) <= 8


def test_hpc_sketch_ignores_homopolymer_lengths():
    rng = np.random.default_rng(9)
    base = bytes(rng.choice(list(b"ACGT"), size=1500).tolist())
    comp = bytes(c for i, c in enumerate(base) if i == 0 or c != base[i - 1])            # run-compressed
    stretched = b"".join(bytes([c]) * (1 + (i % 3)) for i, c in enumerate(comp))
    a, b = oracle.sketch(comp, 5, 19, 0, True), oracle.sketch(stretched, 5, 19, 0, True)
    assert (a["x"] >> 8).tolist() == (b["x"] >> 8).tolist()                               # same k-mers in the same order
    assert np.all((b["x"] & 0xFF) >= (a["x"] & 0xFF))                                     # spans grow with the runs


def test_identical_query_and_target_chain():
    rng = np.random.default_rng(1)
    seq = bytes(rng.choice(list(b"ACGT"), size=4000).tolist())
    T = ReadSet.from_records([("t0", seq)])
    opt = oracle.options("ont", dual=True)
    idx = oracle.Index(T, opt)
    r = idx.map(seq, b"q0")["regs"]
    n_min = oracle.sketch(seq, 5, 15).shape[0]
    best = r[np.argmax(r["score"])]
    assert best["rev"] == 0 and best["cnt"] == n_min
    assert best["qs"] == best["rs"] and best["qe"] == best["re"]
    assert best["mlen"] == best["blen"] == best["qe"] - best["qs"]
    # the reverse complement maps to the same place on the other strand
    r2 = idx.map(revcomp(seq), b"q0")["regs"]
    b2 = r2[np.argmax(r2["score"])]
    assert b2["rev"] == 1 and b2["cnt"] == n_min and b2["score"] == best["score"]


def test_self_diagonal_is_skipped_in_ava():
    rng = np.random.default_rng(2)
    seq = bytes(rng.choice(list(b"ACGT"), size=3000).tolist())
    T = ReadSet.from_records([("r0", seq)])
    opt = oracle.options("ont", dual=False)
    idx = oracle.Index(T, opt)
    assert idx.map(seq, b"r0")["regs"].shape[0] == 0          # same name + same length: diagonal anchors dropped
    assert idx.map(seq, b"r1")["regs"].shape[0] == 0          # "r1" > "r0": --dual=no skips the pair
    assert idx.map(seq, b"a")["regs"].shape[0] == 1


def test_counts_invariant_to_target_order():
    g = readsim.genome(120_000)
    T = readsim.reads(g, 150, 3000, readsim.ONT, readsim.SEED_TARGET, "t")
    Q = readsim.reads(g, 50, 3000, readsim.ONT, readsim.SEED_QUERY, "q")
    c1, h1, _ = oracle.twoset_counts(T, Q, "ont", threads=4)
    perm = np.random.default_rng(3).permutation(T.n)
    c2, h2, _ = oracle.twoset_counts(T.subset(perm), Q, "ont", threads=4)
    assert np.array_equal(c1, c2) and np.array_equal(h1, h2)


def test_ava_counts_are_symmetric_pair_counts():
    g = readsim.genome(100_000)
    R = readsim.reads(g, 120, 3000, readsim.ONT, readsim.SEED_AVA, "r")
    c, st = oracle.ava_counts(R, "ont", threads=4)
    assert int(c.sum()) == 2 * st["n_overlaps"]
    # two-set with distinct names over the same reads sees each unordered pair from both sides
    Q = ReadSet.from_records([(b"x" + R.name(i), R.seq(i)) for i in range(R.n)])
    c2, _, _ = oracle.twoset_counts(R, Q, "ont", threads=4)
    assert np.all(c2 >= c)          # plus the self overlap and pairs only seen from the larger name


def test_tie_modes_agree_on_counts():
    g = readsim.genome(150_000)
    T = readsim.reads(g, 200, 3000, readsim.ONT, readsim.SEED_TARGET, "t")
    Q = readsim.reads(g, 80, 3000, readsim.ONT, readsim.SEED_QUERY, "q")
    a, _, _ = oracle.twoset_counts(T, Q, "ont", threads=4, tie_mode=oracle.TIE_KSORT)
    b, _, _ = oracle.twoset_counts(T, Q, "ont", threads=4, tie_mode=oracle.TIE_STABLE)
    assert np.array_equal(a, b)
