#!/bin/bash
# N-GPU box (N = 4 or 8): multi-GPU parity (sharded index + per-shard entry), then the bench line at N
set -u
mkdir -p gpurun_out
N=${1:-4}; tag=${2:-x}
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/manual/shard_entry_check.py 2>&1 | grep -E "identical|Error|error" | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n${N}_$tag.json 2> gpurun_out/n${N}_$tag.err; echo "bench N=$N rc=$?"
python - <<PY
import json
for line in open("gpurun_out/n${N}_$tag.json"):
    if line.startswith("{"):
        d=json.loads(line); print(round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), (d.get("cpu_baseline") or {}).get("parity"), (d.get("cpu_baseline") or {}).get("configs"))
        for k in d["roofline"].get("kernels",[])[:16]: print("   ", k["name"], k["ms_per_step"])
        print("   ", d["roofline"].get("stage_ms"))
PY
tail -5 gpurun_out/n${N}_$tag.err | cut -c1-300
