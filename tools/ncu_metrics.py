#!/usr/bin/env python
"""Print the roofline-relevant raw metrics of every kernel in an .ncu-rep (reads `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----", d.get("Kernel Name", "")[:90], d.get("Grid Size"), d.get("Block Size"))
    for w in WANT:
        if w in d:
            print(f"   {w:70s} {d[w]:>16s} {units[hdr.index(w)]}")
    for w in ("smsp__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active"):
        if w in d:
            print(f"   {w:70s} {d[w]:>16s} {units[hdr.index(w)]}")
    stalls = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("_per_warp_active.pct"):
            try:
                stalls.append((float(v.replace(",", "")), k))
            except ValueError:
                pass
    for v, k in sorted(stalls, reverse=True)[:6]:
        print(f"   stall {k.split('issue_stalled_')[1].replace('_per_warp_active.pct', ''):62s} {v:16.2f} % of warp-active cycles")
