mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "prep_rows or resid_stats or lnfold" 2>&1 | tail -30) > gpurun_out/r2_t1.log 2>&1
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -30) > gpurun_out/r2_t2.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -5 gpurun_out/r2_t1.log gpurun_out/r2_t2.log; cat gpurun_out/r2_bench1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['extras'], d['clocks'])"
