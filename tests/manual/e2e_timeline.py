"""LRGE_TIMELINE=1 python tests/manual/e2e_timeline.py: where the time of one host-buffer call (C2) goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lrge_b200
from lrge_b200 import readsim
strat, platform, T, Q = readsim.make_config("C2")
if strat != "twoset":
    raise SystemExit("two-set only")
with lrge_b200.Context(0) as c:
    pT, bT = lrge_b200.pinned_copy(T); pQ, bQ = lrge_b200.pinned_copy(Q)
    for _ in range(3):
        c.twoset_counts(pT, pQ, platform)
    c.set_profiling(True)
    c.twoset_counts(pT, pQ, platform)
    print("---- timed call ----", file=sys.stderr, flush=True)
    r = c.twoset_counts(pT, pQ, platform)
    print({k: round(v, 3) for k, v in r.stats.items() if k.startswith("ms_")})
