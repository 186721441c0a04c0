O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "attention" 2>&1 | tail -3
for a in 3; do
echo "== attention_impl=$a"
MMCM_NCU_RANGE=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --nvtx --nvtx-include "measure/" -k regex:"attention" -c 8 \
  --clock-control none python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --attention-impl $a 2>&1 | grep -E "attention_|duration|dram" | cut -c1-90
done
