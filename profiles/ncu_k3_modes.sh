M=gcc__cache_requests_type_instruction.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio,smsp__average_warp_latency_issue_stalled_wait.ratio,smsp__average_warp_latency_issue_stalled_barrier.ratio,smsp__average_warp_latency_issue_stalled_branch_resolving.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
for mode in "EG3D_K3W_BARRIER=1" "EG3D_K3W_BARRIER=1 EG3D_K3W_MERGE=1" "EG3D_X=1" "EG3D_K3B_LEGACY=1"; do
  echo "== $mode"
  env $mode timeout 300 ncu --clock-control none -k regex:"k3w_|k3b_expand" --metrics $M python profiles/ab_compare.py 1 2>&1 | grep -E "gcc__|smsp__|gpu__time|sm__warps|libeg3d" 
done
