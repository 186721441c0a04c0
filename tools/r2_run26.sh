O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
for c in 0 1 0 1; do
MMCM_ATC_CONTIG=$c timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 > $O/r2_contig$c.json 2> $O/r2_contig$c.err; python -c "
import json; d=json.load(open('$O/r2_contig$c.json')); print('ring contiguous; tc contig=$c', round(d['value']), d['clocks']['sm_mhz'])"
done
MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --nvtx --nvtx-include "measure/" -k regex:"attention" -c 6 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "attention_|duration|dram__bytes" | head -40
