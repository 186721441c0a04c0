mkdir -p gpurun_out
O=gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6) > $O/r02_gputests.log 2>&1
tail -3 $O/r02_gputests.log
timeout 400 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 300 python bench.py --impl reference > $O/r02_bench_reference.json 2> /dev/null
timeout 300 python bench.py --model clip_mtl --batch 256 > $O/r02_bench_clip_mtl_b256.json 2> /dev/null
timeout 300 python bench.py --model siglip_fusion --batch 256 > $O/r02_bench_siglip_fusion_b256.json 2> /dev/null
timeout 300 python bench.py --batch 64 --no-cpu-baseline > $O/r02_bench_clip_fusion_b64.json 2> /dev/null
timeout 300 python bench.py --batch 256 --no-cpu-baseline > $O/r02_bench_clip_fusion_b256.json 2> /dev/null
timeout 300 python bench.py --batch 4096 --no-cpu-baseline > $O/r02_bench_clip_fusion_b4096.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" --clock-control none --csv \
  --log-file $O/r02_b1_launches.csv python tools/b1_launches.py 1 > $O/r02_b1.log 2>&1
python tools/launch_summary.py $O/r02_b1_launches.csv > $O/r02_b1_launch_summary.txt
timeout 200 python tools/latency.py > $O/r02_latency.txt 2>&1
python - <<PY
import json
for f in ("default", "clip_mtl_b256", "siglip_fusion_b256", "clip_fusion_b64", "clip_fusion_b256", "clip_fusion_b4096"):
    try:
        d = json.load(open(f"$O/r02_bench_{f}.json"))
        print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["frac"], 3), d["parity"] and d["parity"]["pass"],
              d["gpu_launches"], d["clocks"]["sm_mhz"], d["extras"].get("latency_b1_ms"), (d["extras"].get("torch_gpu_bf16") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
