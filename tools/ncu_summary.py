#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys, json
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__icc_request_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__shared_mem_per_block_dynamic",
        "dram__bytes.sum.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in d:
        print("%-70s %s %s" % (k, d[k][0], d[k][1]))
st = sorted(((float(v[0]), h) for h, v in d.items() if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and v[0] not in ("", "n/a")), reverse=True)
print("stall reasons (warps per issue-active cycle):")
for v, h in st[:8]:
    print("   %8.3f %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
