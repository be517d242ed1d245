#!/bin/bash
# 2-GPU box: GPU tests, N=1 and N=2 bench (N=2 also with more NCCL point-to-point channels), sanitizer re-run
set -u
mkdir -p gpurun_out
tag=${1:-x}
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for d in 0 1 2; do LRGE_INDEX_BITS_DELTA=$d timeout 200 python bench.py --configs none --no-cpu-baseline > gpurun_out/n1_${tag}$d.json 2> gpurun_out/n1_${tag}$d.err; echo "n1 delta $d rc=$?"; done
run2() { # name, extra env
  env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 10 --warmup 3 --configs none > gpurun_out/n2_$1.json 2> gpurun_out/n2_$1.err; echo "n2 $1 rc=$?"
}
run2 ${tag}a "LRGE_X=0" 29511
run2 ${tag}b "NCCL_MIN_P2P_NCHANNELS=8" 29512
[ "${SANITIZE:-0}" = 1 ] && bash tools/sanitize.sh 2>&1 | grep -E "rc=|SUMMARY"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/n1_$tag?.json")+glob.glob("gpurun_out/n2_$tag?.json")):
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), (d.get("cpu_baseline") or {}).get("parity"))
            for k in d["roofline"].get("kernels",[])[:14]: print("   ", k["name"], k["ms_per_step"])
            print("   ", d["roofline"].get("stage_ms"))
PY
