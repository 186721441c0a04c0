mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25) > gpurun_out/r2_t12.log 2>&1
tail -8 gpurun_out/r2_t12.log
timeout 400 python bench.py > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench12.json')); print(round(d['value']), round(d['e2e']['value']), d['e2e'].get('copy_share_of_call'), d['e2e'].get('frac_of_h2d_ceiling'), round(d['roofline']['frac'],3), d['roofline']['traffic'], d['clocks']['sm_mhz'], {k: (round(x) if x > 100 else x) for k,x in d['extras'].items() if k.startswith('value') or k.startswith('latency')}); print(json.dumps(d['roofline']['by_epilogue'], indent=1))"
