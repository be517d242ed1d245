#!/bin/bash
# One GPU call that gathers the round's evidence: tests, bench (ours + reference arm), ncu launch list, ncu --set full of
# the hand-written kernels, compute-sanitizer passes.  Usage (under gpurun): bash tools/gpu_round_check.sh <tag>
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py 2> gpurun_out/bench_${TAG}.err | tail -1 > gpurun_out/bench_${TAG}.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref_${TAG}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"sort_query_kernel|sketch2_kernel|seed_lookup_kernel|anchor_fill_kernel|chain_group_kernel|index_post_kernel|compact_kernel|mz_flt_kernel|key32_kernel" \
    -s 27 -c 10 -o gpurun_out/prof_${TAG} python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_full_${TAG}.log | cut -c1-120
LRGE_TIMELINE=1 timeout 200 python tests/manual/e2e_timeline.py > gpurun_out/e2e_timeline_${TAG}.txt 2>&1
bash tools/sanitize.sh 2>&1 | grep -E "rc=|SUMMARY" | tee gpurun_out/sanitize_${TAG}.log
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e","clocks","gpu_launches")})
print({k:v for k,v in d["roofline"].items() if k not in ("kernels","counts","stage_ms","note")})
for k in d["roofline"]["kernels"][:10]: print(k["name"], k["ms_per_step"], k["frac"])
print(d["cpu_baseline"]["parity"], {k:(v["gpu_ms"], v["counts_identical"]) for k,v in d["cpu_baseline"]["configs"].items()})
r=json.load(open("gpurun_out/bench_ref_${TAG}.json")); print("reference:", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
PY
