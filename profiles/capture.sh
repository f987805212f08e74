#!/bin/bash
# Round-1 profile capture (run under gpurun on one B200):  bash profiles/capture.sh
# 1. per-launch device times of one bench command (shares of the step), 2. ncu --set full of the path's kernels on the
# C2 batch (source-level import on), 3. instruction-cache request counters of K3a/K3b.  Outputs land in gpurun_out/ and
# are summarised into profiles/ with profiles/summarize_ncu.py.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_sweep|k3a_|k3b_' -c 7 -o gpurun_out/all_full_r01 \
    python profiles/ab_compare.py 1 > gpurun_out/all_full.log 2>&1
ncu --metrics gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__average_cache_request_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k3a_|k3b_' -c 2 python profiles/ab_compare.py 1 > gpurun_out/icache_r01.log 2>&1
ls -la gpurun_out
