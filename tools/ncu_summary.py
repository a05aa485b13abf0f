#!/usr/bin/env python
"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
seen = set()
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    if name in seen and "--all" not in sys.argv:
        continue
    seen.add(name)
    print("----", name, "grid", r[idx["Grid Size"]] if "Grid Size" in idx else "")
    for w in want:
        if w in idx:
            print(f"  {w:68s} {r[idx[w]]:>18s} {units[idx[w]]}")
    st = [(float(r[idx[h]]), h.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for h in hdr
          if "average_warps_issue_stalled" in h and "per_issue_active" in h and r[idx[h]]]
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
