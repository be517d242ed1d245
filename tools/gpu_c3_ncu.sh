#!/bin/bash
# 1-GPU box: timeline of one host-buffer call; ncu --set full of the seed-stage kernels of C3 (ava-pb)
set -u
mkdir -p gpurun_out
LRGE_TIMELINE=1 timeout 200 python tests/manual/e2e_timeline.py > gpurun_out/e2e_timeline.txt 2>&1; tail -3 gpurun_out/e2e_timeline.txt | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"anchor_fill_kernel|seed_count_kernel" -s 2 -c 3 -o gpurun_out/prof_c3_fill python tests/manual/time_config.py C3 > gpurun_out/ncu_c3.log 2>&1; tail -3 gpurun_out/ncu_c3.log | cut -c1-200
ls -la gpurun_out/prof_c3_fill.ncu-rep
