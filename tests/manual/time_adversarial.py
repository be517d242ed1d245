"""Times each product call on the adversarial ava-pb set (where the slow corners are)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lrge_b200, adversarial
from lrge_b200 import readsim

kw = dict(nt=60, nq=16, mean=3000, size=60_000, n_tandem=2, n_lowc=1, n_dup=1)
T, Q = adversarial.repeat_sets(errors=readsim.HIFI, **kw)
def timed(name, fn):
    t = time.time(); r = fn(); print(f"{name}: {time.time() - t:.2f}s", flush=True); return r
with lrge_b200.Context(0) as c:
    r = timed("counts", lambda: c.twoset_counts(T, Q, "pb"))
    print({k: r.stats[k] for k in ("n_anchors", "n_pairs", "n_batches", "n_deferred", "mid_occ", "ms_seed", "ms_chain", "ms_total")})
    timed("counts again", lambda: c.twoset_counts(T, Q, "pb"))
    timed("counts -F", lambda: c.twoset_counts(T, Q, "pb", remove_internal=True))
    timed("overlaps", lambda: c.twoset_overlaps(T, Q, "pb"))
for env in (dict(LRGE_FORCE_GLOBAL="1"), dict(LRGE_MAP_CAP="1024")):
    os.environ.update(env)
    with lrge_b200.Context(0) as c:
        r = timed(f"counts {env}", lambda: c.twoset_counts(T, Q, "pb"))
        print({k: r.stats[k] for k in ("n_batches", "n_deferred", "ms_seed", "ms_sort", "ms_chain", "ms_total")})
    for k in env: del os.environ[k]
