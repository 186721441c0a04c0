#!/bin/bash
# Round-2 evidence run of the LAST build (one GPU), after the ring attention / prefetch changes.  The GEMM kernels are
# those of tools/final_profile_r2.sh, so its per-shape ncu table and traffic JSON are not repeated.
# Usage from the repo root on a GPU box: bash tools/final_profile_r2b.sh  -> gpurun_out/r02_*
mkdir -p gpurun_out
O=gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8) > $O/r02_gputests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
timeout 400 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 300 python bench.py --impl reference > $O/r02_bench_reference.json 2> /dev/null
timeout 300 python bench.py --model clip_mtl --batch 256 > $O/r02_bench_clip_mtl_b256.json 2> /dev/null
timeout 300 python bench.py --model siglip_fusion --batch 256 > $O/r02_bench_siglip_fusion_b256.json 2> /dev/null
timeout 300 python bench.py --batch 64 --no-cpu-baseline > $O/r02_bench_clip_fusion_b64.json 2> /dev/null
timeout 300 python bench.py --batch 256 --no-cpu-baseline > $O/r02_bench_clip_fusion_b256.json 2> /dev/null
timeout 300 python bench.py --batch 4096 --no-cpu-baseline > $O/r02_bench_clip_fusion_b4096.json 2> /dev/null
# launch list of one bench run (the kernel's SHARE of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches.csv \
  python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline --no-e2e > $O/r02_ncu_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches.csv > $O/r02_launch_summary.txt
# ncu --set full of the ring attention: launches 10, 11 = text tower (77 tokens, causal), 12, 13 = vision tower (50 tokens)
MMCM_NCU_RANGE=1 timeout 900 ncu --set full --import-source on --nvtx --nvtx-include "measure/" -k regex:attention_ring \
  --launch-skip 10 -c 4 --clock-control none -f -o $O/r02_attention_ring python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e \
  > $O/r02_attention_ring.log 2>&1
# B = 1 launch list + latency table
timeout 300 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" --clock-control none --csv \
  --log-file $O/r02_b1_launches.csv python tools/b1_launches.py 1 > $O/r02_b1.log 2>&1
python tools/launch_summary.py $O/r02_b1_launches.csv > $O/r02_b1_launch_summary.txt
timeout 200 python tools/latency.py > $O/r02_latency.txt 2>&1
tail -3 $O/r02_gputests.log; tail -1 $O/r02_smoke.log
python - <<PY
import json
for f in ("default", "clip_mtl_b256", "siglip_fusion_b256", "clip_fusion_b64", "clip_fusion_b256", "clip_fusion_b4096"):
    try:
        d = json.load(open(f"$O/r02_bench_{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["pipelined"]["value"]), round(d["roofline"]["frac"], 3),
              d["parity"] and d["parity"]["pass"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["extras"].get("latency_b1_ms"),
              (d["extras"].get("torch_gpu_bf16") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
head -12 $O/r02_launch_summary.txt
