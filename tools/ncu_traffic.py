#!/usr/bin/env python
"""Roofline evidence from `ncu --set full` reports (read here with `ncu -i ... --page raw --csv`).

    python tools/ncu_traffic.py shapes <report.ncu-rep> <order.json>   -> per-shape table (duration, TFLOP/s, tensor-pipe %,
                                                                          DRAM bytes vs algorithmic)
    python tools/ncu_traffic.py bench <report.ncu-rep> <model> [json]  -> mean DRAM bytes per encoder-GEMM launch of a
                                                                          forward, merged into profiles/r02_gemm_traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        d["_units"] = dict(zip(hdr, units))
        res.append(d)
    return res


def num(d, key):
    v = float(d[key].replace(",", ""))
    u = d["_units"].get(key, "")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
             "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    return v * scale


def main():
    mode, rep = sys.argv[1], sys.argv[2]
    rows = rows_of(rep)
    if mode == "shapes":
        order = json.load(open(sys.argv[3]))
        gem = [d for d in rows if "gemm2_tcgen05" in d.get("Kernel Name", "")]
        assert len(gem) == len(order), (len(gem), len(order))
        print(f"{'GEMM':36s} {'M':>6s} {'N':>5s} {'K':>5s} {'us':>7s} {'TFLOP/s':>8s} {'tensor%':>8s} {'DRAM MB':>8s} "
              f"{'alg MB':>7s} {'DRAM TB/s':>9s} {'L2 hit%':>7s}")
        for d, o in zip(gem, order):
            us = num(d, "gpu__time_duration.sum")
            dram = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
            tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan")
            hit = d.get("lts__t_sector_hit_rate.pct", "nan")
            print(f"{o['name']:36s} {o['M']:6d} {o['N']:5d} {o['K']:5d} {us:7.1f} {o['flops'] / us / 1e6:8.1f} {tp:>8s} "
                  f"{dram / 1e6:8.1f} {o['algorithmic_bytes'] / 1e6:7.1f} {dram / us / 1e6:9.2f} {hit:>7s}")
        return
    model = sys.argv[3]
    dst = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    gem = [d for d in rows if "gemm2_tcgen05" in d.get("Kernel Name", "")]
    per = {}
    tot_b = tot_us = 0.0
    for d in gem:
        k = d["Kernel Name"].split("(")[0]
        b = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
        us = num(d, "gpu__time_duration.sum")
        e = per.setdefault(k, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += b
        e["us"] += us
        tot_b += b
        tot_us += us
    table = json.load(open(dst)) if os.path.exists(dst) else {}
    table[model] = {"source": os.path.basename(rep) + ": ncu --set full --clock-control none of the encoder GEMM launches "
                              "of one bench forward (cold-cache, serialised replays)",
                    "gemm_launches": len(gem), "mean_dram_bytes_per_gemm_launch": tot_b / max(len(gem), 1),
                    "mean_us_per_gemm_launch": tot_us / max(len(gem), 1), "per_kernel": per}
    json.dump(table, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(table[model], indent=1))


if __name__ == "__main__":
    main()
