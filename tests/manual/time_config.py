"""Per-kernel times of one BASELINE.json configuration (resident inputs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import lrge_b200
from lrge_b200 import readsim
name = sys.argv[1]
strat, platform, A, B = readsim.make_config(name, scale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
with lrge_b200.Context(0) as c:
    sA = c.stage(A); sB = c.stage(B) if B is not None else None
    run = (lambda: c.twoset_counts(sA, sB, platform)) if strat == "twoset" else (lambda: c.ava_counts(sA, platform))
    run(); c.set_profiling(True); run(); r = run()
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.stats.items()})
    for kt in sorted(c.kernel_times(), key=lambda k: -k["ms"]):
        print(f'{kt["name"]:58s} {kt["ms"]:9.3f} ms  x{kt["launches"]}')
