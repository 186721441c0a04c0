O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_forward.py -x -q -m gpu -k "prefetch or forward_host" 2>&1 | tail -5
timeout 400 python bench.py > $O/r2_bench36.json 2> $O/r2_bench36.err; python - <<PY
import json
d=json.load(open('$O/r2_bench36.json'))
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'pipelined', d['e2e'].get('pipelined'), 'u8', round(d['extras']['e2e_u8']['value']))
print('frac', d['roofline']['frac'], 'parity', d['parity'], 'clocks', d['clocks'])
print(json.dumps(d['roofline']['by_epilogue'].get('resid_stats_by_shape'), indent=0)[:1500])
print({k: v for k, v in d['extras'].items() if k.startswith('value_') or k.startswith('latency')})
PY
tail -3 $O/r2_bench36.err
