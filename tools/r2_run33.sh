MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" -k regex:"attention" -c 80 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --attention-impl 3 2>&1 | grep -E "attention_|duration" | paste - - | awk '{print $2, $3, $NF}' | grep "<64" | head -14
