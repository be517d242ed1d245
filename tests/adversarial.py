"""Adversarial read sets for the parity tests: genomes with tandem repeats and low-complexity tracts shared by the
target and query reads.  Uniform random genomes never reach minimap2's order-dependent branches; these do:
mm_seed_mz_flt deletions (a query minimizer occurring > mid_occ and > 1 % of the query), seeds dropped by n > mid_occ,
the max_chain_iter (5 000) window clamp, the max_chain_skip (25) break, and anchors with identical x (the tie order of
the anchor sort).  Test infrastructure only."""
from __future__ import annotations

import numpy as np

from lrge_b200 import readsim


def repeat_genome(size: int = 150_000, seed: int = 7, n_tandem: int = 6, n_lowc: int = 3, n_dup: int = 3,
                  mutate: float = 0.01) -> np.ndarray:
    """Random genome with (i) tandem repeats: a 11-180 bp unit repeated to 2-5 kb, each copy mutated at `mutate`
    per base; (ii) low-complexity tracts (di-/tri-nucleotide, 0.5-2 kb); (iii) dispersed duplications: a 2-4 kb
    segment copied to n other places (mutated)."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    g = readsim.genome(size, seed=seed).copy()

    def mut(a):
        a = a.copy()
        m = rng.random(a.shape[0]) < mutate
        a[m] = acgt[rng.integers(0, 4, int(m.sum()))]
        return a

    pos = 2000
    step = (size - 4000) // (n_tandem + n_lowc + n_dup + 1)
    for i in range(n_tandem):
        unit = acgt[rng.integers(0, 4, int(rng.choice([11, 17, 23, 37, 61, 97, 180])))]
        total = int(rng.integers(2000, 5000))
        tr = np.concatenate([mut(unit) for _ in range(total // unit.shape[0] + 1)])[:total]
        g[pos:pos + total] = tr
        pos += step
    for i in range(n_lowc):
        unit = acgt[rng.integers(0, 4, int(rng.choice([2, 3])))]
        total = int(rng.integers(500, 2000))
        g[pos:pos + total] = np.tile(unit, total // unit.shape[0] + 1)[:total]
        pos += step
    for i in range(n_dup):
        L = int(rng.integers(2000, 4000))
        seg = g[pos:pos + L].copy()
        for _ in range(3):
            at = int(rng.integers(1000, size - L - 1000))
            g[at:at + L] = mut(seg)
        pos += step
    return g


def repeat_sets(nt=260, nq=100, mean=4000, size=150_000, errors=readsim.ONT, seed=7, **kw):
    g = repeat_genome(size, seed=seed, **kw)
    T = readsim.reads(g, nt, mean, errors, readsim.SEED_TARGET, "t")
    Q = readsim.reads(g, nq, mean, errors, readsim.SEED_QUERY, "q")
    return T, Q
