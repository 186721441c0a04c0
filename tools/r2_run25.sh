O=gpurun_out
MMCM_NCU_RANGE=1 timeout 600 ncu --set full --import-source on --nvtx --nvtx-include "measure/" -k regex:"attention_ring" -c 3 \
  --clock-control none -f -o $O/r02_attention_ring python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r02_attention_ring.log 2>&1
tail -2 $O/r02_attention_ring.log
for a in 0 3; do
timeout 300 python bench.py --no-cpu-baseline --no-e2e --steps 20 --attention-impl $a > $O/r2_impl$a.json 2> $O/r2_impl$a.err; python -c "
import json; d=json.load(open('$O/r2_impl$a.json')); print('attention_impl=$a', round(d['value']), d['clocks']['sm_mhz'])"
done
