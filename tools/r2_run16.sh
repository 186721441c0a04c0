mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_randomized.py tests/test_gpu_abi.py -q -m gpu --tb=short 2>&1 | tail -40) > gpurun_out/r2_t16.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/r2_t16.log | head -30
timeout 300 python bench.py --no-cpu-baseline --varlen 1 > gpurun_out/r2_bench16_varlen.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/r2_bench16_varlen.json')); print(round(d['value']), round(d['e2e']['value']), d['extras'].get('value_varlen_text_0'))"
