#!/bin/bash
# 2-GPU box: peer-memory key exchange vs ncclSend/Recv
set -u
mkdir -p gpurun_out
tag=${1:-x}; N=${2:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/manual/shard_entry_check.py 2>&1 | grep -E "identical|Error|error|timed out" | head
run2() { env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $N --steps 10 --warmup 3 --configs none > gpurun_out/n${N}_${tag}$1.json 2> gpurun_out/n${N}_${tag}$1.err; echo "n$N $1 ($2) rc=$?"; }
run2 a "LRGE_X=0" 29512
[ "${AB:-1}" = 1 ] && run2 b "LRGE_NO_PEER_PUSH=1" 29513
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/n${N}_$tag?.json")):
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), (d.get("cpu_baseline") or {}).get("parity"))
            for k in d["roofline"].get("kernels",[])[:20]: print("   ", k["name"], k["ms_per_step"])
            print("   ", d["roofline"].get("stage_ms"))
PY
grep -h -v "^W1017\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/n${N}_${tag}a.err | tail -6 | cut -c1-300
