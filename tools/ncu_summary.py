#!/usr/bin/env python
"""Prints selected metrics of an .ncu-rep (raw page) as `name unit value` lines per kernel."""
import csv, subprocess, sys, re
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(
    r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__warps_active.avg.pct|launch__registers_per_thread|"
    r"launch__(grid|block)_size|launch__occupancy_limit|sm__inst_executed.sum$|smsp__inst_executed.avg.per_cycle_active|"
    r"smsp__issue_active.avg.pct|sm__throughput.avg.pct|gpu__dram_throughput.avg.pct|l1tex__t_sector_hit_rate.pct|"
    r"lts__t_sector_hit_rate.pct|smsp__thread_inst_executed_per_inst_executed.ratio|sm__inst_executed_pipe_(fma|alu|xu|lsu|fmaheavy|fp64).*sum$|"
    r"smsp__average_warps?_issue_stalled_.*_per_issue_active|smsp__warp_issue_stalled.*per_warp_active|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|"
    r"sm__pipe_(fma|alu|xu|fmaheavy|fp64)_cycles_active.avg.pct|smsp__inst_executed_pipe_.*pct|launch__shared_mem_per_block|"
    r"sm__cycles_elapsed.max|smsp__cycles_active.avg$|launch__waves_per_multiprocessor|sm__maximum_warps_per_active_cycle_pct|achieved_occupancy")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for h, u, v in zip(hdr, units, r):
        name = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1].startswith("Triage") else h
        if pat.search(h):
            print(f"  {h} [{u}] = {v}")
