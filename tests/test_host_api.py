"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/lrge_b200.h declares, its host
arithmetic reproduces the reference's golden vectors (estimate.rs:159-343, mapping.rs:422-492), and it refuses
to run without a CUDA device (no CPU fallback).  No compute call needs a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

import lrge_b200
from conftest import ROOT, has_gpu

INF = float("inf")


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lrge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrge_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    L = lrge_b200.lib()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/lrge_b200.h but not exported"
    assert set(names) == set(lrge_b200.C_API)


def test_opts_default_match_builder():   # twoset/builder.rs:22-39
    o = lrge_b200.Opts()
    lrge_b200.lib().lrge_opts_default(ctypes.byref(o))
    assert o.platform == 0 and o.remove_internal == 0 and o.use_min_ref == 0
    assert abs(o.max_overhang_ratio - 0.2) < 1e-7
    assert (o.shard_index, o.shard_count) == (0, 1)


@pytest.mark.skipif(has_gpu(), reason="needs a box without a GPU")
def test_create_fails_without_gpu():
    with pytest.raises(lrge_b200.LrgeError) as e:
        lrge_b200.Context(0)
    assert e.value.kind == "CudaError"
    assert "no CPU fallback" in str(e.value)


def test_platform_parsing():   # lib.rs:149-180
    assert lrge_b200.platform_code("pacbio") == lrge_b200.platform_code("pb") == 1
    assert lrge_b200.platform_code("nanopore") == lrge_b200.platform_code("ONT") == 0
    with pytest.raises(lrge_b200.LrgeError) as e:
        lrge_b200.platform_code("illumina")
    assert e.value.kind == "InvalidPlatform"


@pytest.mark.parametrize("data,expected", [
    ([1.0, 3.0, 5.0, 7.0, 9.0], 5.0), ([3.0, 1.0, 7.0, 5.0, 9.0], 5.0), ([1.0, 3.0, 5.0, 7.0], 4.0), ([10.0], 10.0),
    ([-3.0, 1.0, 0.0, 3.0, -1.0], 0.0), ([1.0, 2.0, 3.0, INF], 2.5), ([-INF, 1.0, 2.0, 3.0], 1.5),
    ([-INF, 1.0, 2.0, INF], 1.5), ([INF, INF], INF), ([-INF, -INF], -INF), ([-1.0, -INF, 0.0, 1.0, INF], 0.0),
])
def test_median_vectors(data, expected):   # estimate.rs:163-273 (median() itself does not filter)
    assert lrge_b200.estimate_from(data, finite=False) == (None, expected, None)


def test_median_empty_and_quantiles():   # estimate.rs:200-203, 275-303
    assert lrge_b200.estimate_from([], finite=False) == (None, None, None)
    assert lrge_b200.estimate_from(list(range(1, 11)), False, 0.15, 0.65) == (np.float32(2.35), 5.5, np.float32(6.85))
    assert lrge_b200.estimate_from([1, 2, 3, 4, 5, 6, INF, INF, INF, INF], False, 0.15, 0.65) == (np.float32(2.35), 5.5, INF)
    with pytest.raises(ValueError):   # estimate.rs:305-310 panics
        lrge_b200.estimate_from([1.0, 2.0, 3.0, 4.0, 5.0], False, 1.1, None)


def test_estimate_finite_filter():   # estimate.rs:63-69
    assert lrge_b200.estimate_from([1.0, 2.0, 3.0, INF], finite=True)[1] == 2.0
    assert lrge_b200.estimate_from([INF, INF], finite=True) == (None, None, None)


def test_per_read_estimate_vectors():   # estimate.rs:305-342
    assert lrge_b200.per_read_estimate(100, 200.0, 1000, 100, 10) == 2910.0
    assert lrge_b200.per_read_estimate(100, 200.0, 1000, 0, 10) == INF


def test_is_internal_vectors():   # mapping.rs:422-492
    assert lrge_b200.is_internal(390, 46, 317, 0, 278, 4, 275, 0.2)
    assert lrge_b200.is_internal(298, 1, 297, 0, 398, 54, 350, 0.2)
    assert not lrge_b200.is_internal(390, 0, 355, 0, 418, 39, 394, 0.05)


def test_host_estimator_equals_oracle_bitwise():
    import oracle
    rng = np.random.default_rng(7)
    for _ in range(200):
        rl, avg = int(rng.integers(500, 60000)), np.float32(rng.uniform(500, 20000))
        nt, nov = int(rng.integers(1, 2_000_000)), int(rng.integers(0, 300))
        a, b = lrge_b200.per_read_estimate(rl, avg, nt, nov, 100), oracle.per_read_estimate(rl, avg, nt, nov, 100)
        assert a == b or (np.isinf(a) and np.isinf(b))
    v = rng.uniform(1e5, 1e7, size=4999).astype(np.float32)
    v[::17] = np.inf
    assert lrge_b200.estimate_from(v, True, 0.15, 0.65) == oracle.estimate(v, True, 0.15, 0.65)


def test_paf_format_golden():   # mapping.rs:197-420: the serialised record of the reference's own test
    ov = np.zeros(1, lrge_b200.OVERLAP_DTYPE)
    ov[0] = (0, 0, 1, 40, 237, 41, 238, 190, 59, 190, 197, 0.0022, 56)
    buf = ctypes.create_string_buffer(512)
    q, t = b"SRR28370649.1", b"SRR28370649.7311"
    n = lrge_b200.lib().lrge_paf_format(ov.ctypes.data, q, len(q), 4402, t, len(t), 5094, 0.0022, 56, buf, 512)
    assert buf.raw[:n] == (b"SRR28370649.1\t4402\t40\t237\t-\tSRR28370649.7311\t5094\t41\t238\t190\t197\t0\t"
                           b"tp:A:S\tcm:i:59\ts1:i:190\tdv:f:0.0022\trl:i:56")
    n = lrge_b200.lib().lrge_paf_format(ov.ctypes.data, q, len(q), 4402, t, len(t), 5094, 0.0, 56, buf, 512)
    assert buf.raw[:n].endswith(b"dv:f:0\trl:i:56")   # mapping.rs:415-419
