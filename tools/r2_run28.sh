O=gpurun_out
# dbg: bit0 consumer suspend-wait, bit1 producer suspend-wait, >>4 = hint ns
for d in 0 $((1+16*200)) $((1+16*1000)) $((3+16*1000)) 0; do
echo "== MMCM_ATR_DBG=$d"
MMCM_ATR_DBG=$d MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --nvtx --nvtx-include "measure/" -k regex:"attention_ring" -c 3 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "duration|inst_executed" | head -6
done
