O=gpurun_out
for a in 3 0; do
echo "== attention_impl=$a"
MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --nvtx --nvtx-include "measure/" -k regex:"attention" -c 30 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --attention-impl $a 2>&1 | grep -E "attention_|duration" | paste - - | awk '{print $2, $NF}' | sort | uniq -c | sort -rn | head -12
done
for a in 3 0 3 0; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --attention-impl $a > $O/r2_impl$a.json 2> $O/r2_impl$a.err; python -c "
import json; d=json.load(open('$O/r2_impl$a.json')); print('attention_impl=$a', round(d['value']), d['clocks']['sm_mhz'])"
done
