mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t7.log 2>&1
tail -6 gpurun_out/r2_t7.log
timeout 200 python tools/latency.py > gpurun_out/r2_latency7.txt 2>&1; grep "ln_fold 1 head_cluster 1" gpurun_out/r2_latency7.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench7.json')); print(round(d['value']), round(d['e2e']['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k: (round(x) if x > 100 else x) for k,x in d['extras'].items() if k.startswith('value') or k.startswith('latency')})"
