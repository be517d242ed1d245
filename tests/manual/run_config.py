"""Run one BASELINE.json configuration on the GPU and (optionally) the oracle; print timings and parity."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lrge_b200  # noqa: E402
from lrge_b200 import readsim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--oracle-queries", type=int, default=0, help="two-set: check only the first n queries on the oracle")
    ap.add_argument("--repeat", type=int, default=3)
    a = ap.parse_args()
    t0 = time.time()
    strat, platform, A, B = readsim.make_config(a.config, scale=a.scale)
    print(f"[{a.config}] generated in {time.time() - t0:.1f}s: {A.n} reads {A.total_bases / 1e6:.1f} Mb" +
          (f" + {B.n} reads {B.total_bases / 1e6:.1f} Mb" if B is not None else ""), flush=True)
    ctx = lrge_b200.Context(0)
    res = None
    for i in range(a.repeat):
        t0 = time.time()
        res = ctx.twoset_counts(A, B, platform) if strat == "twoset" else ctx.ava_counts(A, platform)
        print(f"[{a.config}] gpu pass {i}: wall {time.time() - t0:.3f}s", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in res.stats.items()}, flush=True)
    if strat == "twoset":
        s = lrge_b200.TwoSetStrategy(A, B, platform=platform, ctx=ctx)
    else:
        s = lrge_b200.AvaStrategy(A, platform=platform, ctx=ctx)
    r = s.estimate(True, lrge_b200.LOWER_QUANTILE, lrge_b200.UPPER_QUANTILE)
    print(f"[{a.config}] estimate {r.estimate:.0f} bp (IQR {r.lower:.0f} - {r.upper:.0f}), no_mapping {r.no_mapping_count}; "
          f"true genome {readsim.CONFIGS[a.config][4] * a.scale:.0f}", flush=True)
    if a.oracle:
        import oracle
        th = os.cpu_count() or 1
        t0 = time.time()
        if strat == "twoset":
            Q = B if not a.oracle_queries else B.subset(range(a.oracle_queries))
            want, had, st = oracle.twoset_counts(A, Q, platform, threads=th)
            got = res.counts[:Q.n]
        else:
            want, st = oracle.ava_counts(A, platform, threads=th)
            got = res.counts
        dt = time.time() - t0
        bad = int((got != want).sum())
        print(f"[{a.config}] oracle {dt:.1f}s on {th} threads; counts identical: {bad == 0} (mismatches {bad} of {want.shape[0]})", st, flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
