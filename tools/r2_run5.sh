mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_t5.log 2>&1
tail -4 gpurun_out/r2_t5.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench5.json')); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k: (round(x) if x > 100 else x) for k,x in d['extras'].items() if k.startswith('value') or k.startswith('latency')}, d['extras']['torch_gpu_bf16'].get('value'))"
timeout 200 python tools/latency.py > gpurun_out/r2_latency5.txt 2>&1; grep "ln_fold 1 head_cluster 1" gpurun_out/r2_latency5.txt
timeout 300 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" --clock-control none --csv \
  --log-file gpurun_out/r2_b1_launches.csv python tools/b1_launches.py 1 > gpurun_out/r2_b1.log 2>&1
python tools/launch_summary.py gpurun_out/r2_b1_launches.csv > gpurun_out/r2_b1_summary.txt; cat gpurun_out/r2_b1_summary.txt
timeout 300 python bench.py --no-cpu-baseline --no-e2e --model siglip_fusion --batch 256 > gpurun_out/r2_bench5_siglip.json 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline --no-e2e --model clip_mtl --batch 256 > gpurun_out/r2_bench5_mtl.json 2>/dev/null
for f in siglip mtl; do python -c "
import json
d=json.load(open('gpurun_out/r2_bench5_$f.json')); print('$f', round(d['value']), round(d['roofline']['frac'],3))"; done
