#!/bin/bash
# compute-sanitizer passes over the product path (run under gpurun): memcheck + initcheck on smoke() and on a small
# run that exercises rounds, deferral to the device-wide sort, overlap records, -F and the ava-pb kernels; racecheck on
# the shared-memory kernels of a tiny case.  Logs: gpurun_out/sanitize_*.log (summaries are copied to profiles/).
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import lrge_b200
from lrge_b200 import readsim
g = readsim.genome(60_000)
T = readsim.reads(g, 60, 3000, readsim.ONT, readsim.SEED_TARGET, "t")
Q = readsim.reads(g, 24, 3000, readsim.ONT, readsim.SEED_QUERY, "q")
R = readsim.reads(g, 50, 2500, readsim.HIFI, readsim.SEED_AVA, "r")
with lrge_b200.Context(0) as c:
    a = c.twoset_counts(T, Q, "ont")
    b = c.twoset_counts(T, Q, "ont", remove_internal=True)
    ov, _ = c.twoset_overlaps(T, Q, "ont")
    d = c.ava_counts(R, "pb")
    print("sanitize case:", int(a.counts.sum()), int(b.counts.sum()), ov.shape[0], int(d.counts.sum()), a.stats["n_rounds"], a.stats["n_deferred"])
PY
for tool in memcheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; tail -2 gpurun_out/sanitize_${tool}_smoke.log
  LRGE_MAP_CAP=512 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitize_${tool}_case.log 2>&1
  echo "$tool case rc=$?"; tail -2 gpurun_out/sanitize_${tool}_case.log
done
LRGE_MAP_CAP=512 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san_case.py > gpurun_out/sanitize_racecheck_case.log 2>&1
echo "racecheck case rc=$?"; tail -3 gpurun_out/sanitize_racecheck_case.log
