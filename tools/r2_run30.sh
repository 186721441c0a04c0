O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --nvtx --nvtx-include "measure/" -k regex:"attention_ring" -c 4 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "duration|inst_executed" | head -8
for r in 1 0 1 0; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --attention-ring $r > $O/r2_ring$r.json 2> $O/r2_ring$r.err; python -c "
import json; d=json.load(open('$O/r2_ring$r.json')); print('ring=$r', round(d['value']), d['clocks']['sm_mhz'])"
done
