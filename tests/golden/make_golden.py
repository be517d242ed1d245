"""Regenerates tests/golden/*.json.

reference_vectors.json -- the golden vectors the REFERENCE's own tests hold next to the hot path, transcribed with
    their source lines (liblrge/src/estimate.rs:159-343, liblrge/src/minimap2/mapping.rs:197-492).
oracle_small.json      -- outputs of the CPU oracle (oracle/mm2_oracle.c) on small seeded read sets.  The overlap path
    has no reference-side vector (parity unpinned); this file pins the ORACLE against regressions
    and gives the GPU tests a fixed target that does not depend on the oracle library being rebuilt identically.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from lrge_b200 import readsim  # noqa: E402

INF = "inf"


def reference_vectors():
    return {
        "source": "mbhall88/lrge @ 9916ea3 (liblrge 0.3.0)",
        "median": [   # estimate.rs:163-273: (data, expected median)
            {"line": 163, "data": [1.0, 3.0, 5.0, 7.0, 9.0], "median": 5.0},
            {"line": 172, "data": [3.0, 1.0, 7.0, 5.0, 9.0], "median": 5.0},
            {"line": 181, "data": [1.0, 3.0, 5.0, 7.0], "median": 4.0},
            {"line": 190, "data": [10.0], "median": 10.0},
            {"line": 205, "data": [-3.0, 1.0, 0.0, 3.0, -1.0], "median": 0.0},
            {"line": 214, "data": [1.0, 2.0, 3.0, INF], "median": 2.5},
            {"line": 223, "data": ["-inf", 1.0, 2.0, 3.0], "median": 1.5},
            {"line": 232, "data": ["-inf", 1.0, 2.0, INF], "median": 1.5},
            {"line": 241, "data": [INF, INF], "median": INF},
            {"line": 250, "data": ["-inf", "-inf"], "median": "-inf"},
            {"line": 259, "data": [-1.0, "-inf", 0.0, 1.0, INF], "median": 0.0},
        ],
        "quantiles": [   # estimate.rs:275-303: q1 = 0.15, q3 = 0.65
            {"line": 275, "data": [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], "lower": 2.35, "median": 5.5, "upper": 6.85},
            {"line": 284, "data": [1, 2, 3, 4, 5, 6, INF, INF, INF, INF], "lower": 2.35, "median": 5.5, "upper": INF},
        ],
        "per_read_estimate": [   # estimate.rs:305-342: (read_len, avg_target_len, n_target_reads, n_ovlaps, thresh)
            {"line": 305, "args": [100, 200.0, 1000, 100, 10], "expected": 2910.0},
            {"line": 324, "args": [100, 200.0, 1000, 0, 10], "expected": INF},
        ],
        "is_internal": [   # mapping.rs:422-492: (qlen, qs, qe, rev, tlen, ts, te, ratio)
            {"line": 422, "args": [390, 46, 317, 0, 278, 4, 275, 0.2], "expected": True},
            {"line": 446, "args": [298, 1, 297, 0, 398, 54, 350, 0.2], "expected": True},
            {"line": 470, "args": [390, 0, 355, 0, 418, 39, 394, 0.05], "expected": False},
        ],
        "paf_line": {   # mapping.rs:197-199 and 415-419
            "line": 198,
            "record": {"query": "SRR28370649.1", "qlen": 4402, "qs": 40, "qe": 237, "rev": 1, "target": "SRR28370649.7311",
                       "tlen": 5094, "ts": 41, "te": 238, "mlen": 190, "blen": 197, "cnt": 59, "score": 190, "dv": 0.0022, "rl": 56},
            "text": "SRR28370649.1\t4402\t40\t237\t-\tSRR28370649.7311\t5094\t41\t238\t190\t197\t0\ttp:A:S\tcm:i:59\ts1:i:190\tdv:f:0.0022\trl:i:56",
        },
    }


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def oracle_small():
    out = {"generator": "tests/golden/make_golden.py", "status": "oracle-generated; the reference itself cannot run here (parity unpinned)",
           "cases": {}}
    for name, platform, err, nt, nq, mean, gsize in (("ont_small", "ont", readsim.ONT, 120, 40, 3000, 100_000),
                                                      ("pb_small", "pb", readsim.HIFI, 80, 30, 3000, 80_000)):
        g = readsim.genome(gsize)
        T = readsim.reads(g, nt, mean, err, readsim.SEED_TARGET, "t")
        Q = readsim.reads(g, nq, mean, err, readsim.SEED_QUERY, "q")
        mins, cnt = oracle.sketch_readset(T, oracle.options(platform))
        counts, had, st = oracle.twoset_counts(T, Q, platform, threads=2)
        inv, _ = oracle.twoset_inverse_counts(T, Q, platform, threads=2)
        cF, _, _ = oracle.twoset_counts(T, Q, platform, remove_internal=True, threads=2)
        R = T.subset(range(60))
        ava, ast = oracle.ava_counts(R, platform, threads=2)
        regs, qidx, _ = oracle.twoset_regs(T, Q, platform, threads=2, tie_mode=oracle.TIE_STABLE)
        cols = [qidx.astype(np.int64)] + [regs[f].astype(np.int64) for f in ("rid", "rev", "qs", "qe", "rs", "re", "score", "cnt", "mlen", "blen")]
        order = np.lexsort(tuple(reversed(cols)))       # total order over all fields
        avg = np.float32(T.total_bases) / np.float32(T.n)
        est = [oracle.per_read_estimate(int(l), avg, T.n, int(c), 100) for l, c in zip(Q.lengths(), counts)]
        lo, md, up = oracle.estimate(est, True, 0.15, 0.65)
        out["cases"][name] = {
            "platform": platform, "genome": gsize, "n_target": nt, "n_query": nq, "mean_len": mean,
            "target_bases_sha256": sha(T.bases), "query_bases_sha256": sha(Q.bases),
            "n_target_minimizers": int(mins.shape[0]), "minimizers_x_sha256": sha(mins["x"]), "minimizers_y_sha256": sha(mins["y"]),
            "mid_occ": int(st["mid_occ"]), "n_keys": int(st["n_keys"]), "n_anchors": int(st["n_anchors"]),
            "twoset_counts": counts.tolist(), "had_mapping": had.tolist(),
            "twoset_counts_filter_contained": cF.tolist(), "inverse_counts": inv.tolist(),
            "ava_counts_first60_targets": ava.tolist(), "ava_pairs": int(ast["n_overlaps"]),
            "n_overlap_records": int(regs.shape[0]),
            "overlap_records_sha256": sha(np.stack([c[order] for c in cols])),
            "estimate": {"lower": lo, "median": md, "upper": up},
        }
    return out


if __name__ == "__main__":
    json.dump(reference_vectors(), open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)
    json.dump(oracle_small(), open(os.path.join(HERE, "oracle_small.json"), "w"), indent=1)
    print("wrote", os.listdir(HERE))
