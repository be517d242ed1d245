#!/bin/bash
# One GPU call: tests, bench, ncu launch list + full captures of the top kernels, optional large config.
# Usage (under gpurun): bash tools/gpu_round_check.sh <tag> [C4]
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 120 python bench.py 2>&1 | tail -1 > gpurun_out/bench_${TAG}.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_${TAG}.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:"DeviceRadixSortOnesweep|sketch_kernel|seed_count_kernel|chain_group_kernel|anchor_fill_kernel|index_post_kernel" \
    -s 40 -c 24 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-120
if [ "${2:-}" = "C4" ]; then
    timeout 400 python tests/manual/run_config.py C4 --repeat 2 --oracle --oracle-queries 2000 2>&1 | tail -6 | cut -c1-900 | tee gpurun_out/c4_${TAG}.log
fi
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print({k:d[k] for k in ("value","ms_per_step","step_ms","stage_ms","e2e","cpu_baseline","clocks","gpu_launches")})
print(d["roofline"])
for k in d["kernels"][:8]: print(k["name"], k["ms_per_step"], k["frac"])
r=json.load(open("gpurun_out/bench_ref_${TAG}.json")); print("reference:", r["value"], r["ms_per_step"], r["cpu_baseline"])
PY
