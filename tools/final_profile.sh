#!/bin/bash
# Round-end evidence run (one GPU): default bench, ncu launch list of the same workload, the other BASELINE configs.
# Usage (from the repo root on a GPU box): bash tools/final_profile.sh <tag>     -> files under gpurun_out/
tag=${1:-final}
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline --no-e2e \
  > gpurun_out/ncu_${tag}.log 2>&1
python tools/launch_summary.py gpurun_out/launches_${tag}.csv > gpurun_out/launch_summary_${tag}.txt
timeout 200 python bench.py --model siglip_fusion --batch 256 > gpurun_out/bench_siglip_${tag}.json 2>/dev/null
timeout 200 python bench.py --model clip_mtl --batch 256 > gpurun_out/bench_mtl_${tag}.json 2>/dev/null
python - <<PY
import json
for f in ("bench_${tag}", "bench_siglip_${tag}", "bench_mtl_${tag}"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), d["parity"]["pass"],
          d["gpu_launches"], d["clocks"]["sm_mhz"],
          {k: round(v) for k, v in d["extras"].items() if k.startswith("value")})
PY
tail -4 gpurun_out/launch_summary_${tag}.txt
