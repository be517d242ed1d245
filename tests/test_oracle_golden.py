"""Pins the oracle against every golden vector the reference's own tests hold next to the hot path
(SURVEY.md section 8c): estimate.rs:159-343 (median / quantiles / per_read_estimate) and
mapping.rs:422-492 (is_internal).  The overlap path itself has no reference vector (parity unpinned)."""
import math

import numpy as np
import pytest

import oracle

INF = float("inf")


# estimate.rs:163-273
@pytest.mark.parametrize("data,expected", [
    ([1.0, 3.0, 5.0, 7.0, 9.0], 5.0),
    ([3.0, 1.0, 7.0, 5.0, 9.0], 5.0),
    ([1.0, 3.0, 5.0, 7.0], 4.0),
    ([10.0], 10.0),
    ([-3.0, 1.0, 0.0, 3.0, -1.0], 0.0),
    ([1.0, 2.0, 3.0, INF], 2.5),
    ([-INF, 1.0, 2.0, 3.0], 1.5),
    ([-INF, 1.0, 2.0, INF], 1.5),
    ([INF, INF], INF),
    ([-INF, -INF], -INF),
    ([-1.0, -INF, 0.0, 1.0, INF], 0.0),
])
def test_median_vectors(data, expected):
    assert oracle.median(data) == (None, expected, None)


def test_median_empty():  # estimate.rs:200-203
    assert oracle.median([]) == (None, None, None)


def test_median_with_quantiles():  # estimate.rs:275-282
    lo, md, up = oracle.median([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], 0.15, 0.65)
    assert (lo, md, up) == (np.float32(2.35), 5.5, np.float32(6.85))


def test_quantile_with_infinity():  # estimate.rs:284-303
    lo, md, up = oracle.median([1, 2, 3, 4, 5, 6, INF, INF, INF, INF], 0.15, 0.65)
    assert (lo, md, up) == (np.float32(2.35), 5.5, INF)


def test_per_read_estimate():  # estimate.rs:305-322
    assert oracle.per_read_estimate(100, 200.0, 1000, 100, 10) == 2910.0


def test_per_read_estimate_zero_ovlaps():  # estimate.rs:324-342
    assert oracle.per_read_estimate(100, 200.0, 1000, 0, 10) == INF


def test_estimate_filters_infinite():  # estimate.rs:63-69
    lo, md, up = oracle.estimate([1.0, 2.0, 3.0, INF], finite=True)
    assert md == 2.0
    lo, md, up = oracle.estimate([1.0, 2.0, 3.0, INF], finite=False)
    assert md == 2.5
    assert oracle.estimate([INF, INF], finite=True) == (None, None, None)


# mapping.rs:422-492
def test_is_internal_vectors():
    assert oracle.is_internal(390, 46, 317, 0, 278, 4, 275, 0.2)
    assert oracle.is_internal(298, 1, 297, 0, 398, 54, 350, 0.2)
    assert not oracle.is_internal(390, 0, 355, 0, 418, 39, 394, 0.05)


def test_preset_option_block():
    o = oracle.options("ont", dual=True)
    assert (o.k, o.w, o.is_hpc, o.bw, o.min_chain_score, o.max_chain_skip, o.occ_dist) == (15, 5, 0, 2000, 100, 25, 0)
    assert o.flag & oracle.F_NO_DIAG and not o.flag & oracle.F_NO_DUAL
    o = oracle.options("pb", dual=False)
    assert (o.k, o.w, o.is_hpc, o.bw, o.bw_long, o.min_chain_score) == (19, 5, 1, 500, 500, 100)
    assert o.flag & oracle.F_NO_DUAL


def test_gap_penalty_constants():
    # chn_pen_gap = (float)(0.8f*0.01*k): 0.12000000476837158 (k=15), 0.15199999511241913 (k=19)
    p15 = oracle.gap_penalty_table(15, 2001)
    p19 = oracle.gap_penalty_table(19, 501)
    assert p15[0] == 0 and p19[0] == 0
    g15 = np.float32(np.float64(np.float32(0.8)) * 0.01 * 15)
    assert float(g15) == 0.12000000476837158

    def mg_log2(x):
        z = np.array([x], dtype=np.float32).view(np.uint32)[0]
        log2 = np.float32(((int(z) >> 23) & 255) - 128)
        z = (int(z) & ~(255 << 23)) + (127 << 23)
        zf = np.array([z], dtype=np.uint32).view(np.float32)[0]
        poly = np.float32(np.float32(np.float32(np.float32(-0.34484843) * zf) + np.float32(2.02466578)) * zf) - np.float32(0.67487759)
        return np.float32(log2 + np.float32(poly))

    for dd in (1, 2, 3, 7, 50, 499, 1000, 2000):
        lin = np.float32(g15 * np.float32(dd))
        want = int(np.float32(lin + np.float32(np.float32(0.5) * mg_log2(np.float32(dd + 1)))))
        assert p15[dd] == want, dd
    assert all(np.diff(p15) >= 0)
