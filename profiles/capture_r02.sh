#!/bin/bash
# Round-2 profile capture (run under gpurun on ONE B200):  bash profiles/capture_r02.sh
# 1. per-launch device times of one bench command (shares of the step); 2. ncu --set full of the path's kernels on the C2 batch
# (source import on) and of the two stand-alone GN kernels on 1 M hypotheses; 3. instruction-cache request counters of K3a / K3b.
# Outputs land in gpurun_out/ (scratch) and are summarised into profiles/r02_* by profiles/summarize_ncu.py (run here or there).
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02e_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02e_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_sweep|k3a_|k3b_' -c 7 -o gpurun_out/r02e_full \
    python profiles/ab_compare.py 1 > gpurun_out/r02e_full.log 2>&1
EG3D_K1_FULL=1 ncu --set full --clock-control none -k regex:'k1_sweep' -c 2 -o gpurun_out/r02e_full_k1sweep \
    python profiles/ab_compare.py 1 > gpurun_out/r02e_full_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gn32_|gn64_' -c 2 -o gpurun_out/r02e_full_gn \
    python profiles/run_configs.py c5 --n 1000000 > gpurun_out/r02e_full_gn.log 2>&1
ncu --metrics gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,gcc__average_cache_request_hit_rate.pct,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'k3a_|k3b_' -c 2 python profiles/ab_compare.py 1 > gpurun_out/r02e_icache.log 2>&1
for r in r02e_full r02e_full_k1sweep r02e_full_gn; do python profiles/summarize_ncu.py full gpurun_out/$r.ncu-rep gpurun_out/${r}_summary.txt; done
python profiles/summarize_ncu.py launches gpurun_out/r02e_launches.csv gpurun_out/r02e_launches_summary.txt
ls -la gpurun_out
