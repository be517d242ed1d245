#!/bin/bash
# 2-GPU box: GPU tests, N=1 A/B of the side stream, N=2 bench, then the C5-shaped per-shard run (per-rank load of C5)
set -u
mkdir -p gpurun_out
tag=${1:-x}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do LRGE_NO_SIDE_STREAM=$v timeout 200 python bench.py --configs none --no-cpu-baseline > gpurun_out/n1_${tag}$v.json 2> gpurun_out/n1_${tag}$v.err; echo "n1 no_side=$v rc=$?"; done
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --configs none > gpurun_out/n2_${tag}a.json 2> gpurun_out/n2_${tag}a.err; echo "n2 rc=$?"
run2() { env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 10 --warmup 3 --configs none > gpurun_out/n2_${tag}$1.json 2> gpurun_out/n2_${tag}$1.err; echo "n2 $1 ($2) rc=$?"; }
run2 b "NCCL_P2P_USE_CUDA_MEMCPY=1" 29512
run2 c "NCCL_P2P_NVL_CHUNKSIZE=4194304 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32" 29513
run2 d "NCCL_BUFFSIZE=33554432" 29514
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/n1_$tag?.json")+glob.glob("gpurun_out/n2_$tag?.json")):
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); print(f, round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), (d.get("cpu_baseline") or {}).get("parity"))
            for k in d["roofline"].get("kernels",[])[:12]: print("   ", k["name"], k["ms_per_step"])
            print("   ", d["roofline"].get("stage_ms"))
PY
if [ "${C5:-1}" = 1 ]; then
  free -g | head -2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tests/manual/c5_shard_run.py ${C5_ARGS:-250000 25000 12000 3000000000} > gpurun_out/c5_shard_$tag.json 2> gpurun_out/c5_shard_$tag.err; echo "c5 rc=$?"
  cat gpurun_out/c5_shard_$tag.json | cut -c1-1500; grep -v "^W1017\|^\*\*\*\|OMP_NUM" gpurun_out/c5_shard_$tag.err | tail -8 | cut -c1-300
fi
