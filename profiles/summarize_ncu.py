#!/usr/bin/env python
"""Turn ncu reports (gpurun_out/*.ncu-rep, scratch) into the small text summaries committed under profiles/.
  python profiles/summarize_ncu.py launches <launches.csv> <out.txt>      # per-kernel time shares of one bench command
  python profiles/summarize_ncu.py full <report.ncu-rep> <out.txt>        # key --set full metrics per captured launch
"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[h]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = {}
    for r in rows[h + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"us": 1e-3, "ns": 1e-6, "s": 1e3}.get(r[ui], 1.0)
        agg.setdefault(r[ki][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# per-kernel device time (ms) from {path}: cold-cache, serialised launches — compare SHARES, not absolutes\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:72s} launches={len(v):3d} total_ms={sum(v):10.3f} share={sum(v) / tot:6.3f}\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    H, U = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n")
        for R in rows[2:]:
            f.write(f"== {R[H.index('Kernel Name')][:100]}\n")
            for k in KEYS:
                if k in H:
                    f.write(f"  {k:80s} {R[H.index(k)]} {U[H.index(k)]}\n")
            st = []
            for i, hn in enumerate(H):
                if hn.startswith("smsp__average_warps_issue_stalled") and hn.endswith("_per_issue_active.ratio"):
                    try:
                        st.append((float(R[i]), hn.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError:
                        pass
            f.write("  warp stall cycles per issued instruction: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
