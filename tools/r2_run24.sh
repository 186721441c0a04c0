timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -8
for r in 1 0 1 0; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --attention-ring $r > gpurun_out/r2_ring$r.json 2> gpurun_out/r2_ring$r.err; python -c "
import json; d=json.load(open('gpurun_out/r2_ring$r.json')); print('ring=$r', round(d['value']), d['clocks']['sm_mhz'], d.get('parity'))"
done
for r in 1 0; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --varlen 1 --attention-ring $r > gpurun_out/r2_ringv$r.json 2> gpurun_out/r2_ringv$r.err; python -c "
import json; d=json.load(open('gpurun_out/r2_ringv$r.json')); print('varlen ring=$r', round(d['value']), d['clocks']['sm_mhz'])"
done
