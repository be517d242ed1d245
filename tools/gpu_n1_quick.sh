#!/bin/bash
# 1-GPU box: GPU tests, the bench line (no oracle), per-kernel times of C3 / C4
set -u
mkdir -p gpurun_out
tag=${1:-x}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 200 python bench.py --configs none --no-cpu-baseline > gpurun_out/n1_${tag}.json 2> gpurun_out/n1_${tag}.err; echo "n1 rc=$?"
python - <<PY
import json
for line in open("gpurun_out/n1_$tag.json"):
    if line.startswith("{"):
        d=json.loads(line); print(round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3))
        for k in d["roofline"].get("kernels",[])[:12]: print("   ", k["name"], k["ms_per_step"])
PY
for cfg in ${CFGS:-C3}; do timeout 300 python tests/manual/time_config.py $cfg 2>&1 | head -24; done
if [ "${TIMELINE:-0}" = 1 ]; then LRGE_TIMELINE=1 timeout 200 python tests/manual/e2e_timeline.py 2>&1 | grep -A60 "timed call" | cut -c1-160; LRGE_INDEX_HALVES=1 timeout 200 python bench.py --configs none --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('halves on:', d['ms_per_step'], d['e2e']['ms_per_step'])"; fi
