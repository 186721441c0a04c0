mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_t4.log 2>&1
tail -4 gpurun_out/r2_t4.log
for v in default nstg1 att1b5 att1b4 att2b4; do
  if [ $v = default ]; then unset MMCM_LIB_PATH; else export MMCM_LIB_PATH=$PWD/build/libmmcm_$v.so; fi
  echo "=== $v" >> gpurun_out/r2_run4_bench.txt
  if [ $v = default ] || [ $v = nstg1 ]; then timeout 200 python tools/gemm_bench_fold.py 492 30 >> gpurun_out/r2_run4_bench.txt 2>&1; fi
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/r2_bench4_$v.json 2>> gpurun_out/r2_bench4_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench4_$v.json')); print('$v', round(d['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], {k: round(x) for k,x in d['extras'].items() if k.startswith('value')})" >> gpurun_out/r2_run4_bench.txt
done
unset MMCM_LIB_PATH
timeout 200 python tools/latency.py >> gpurun_out/r2_run4_bench.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/r2_launches_run4.csv python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline --no-e2e \
  > gpurun_out/r2_ncu_run4.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_run4.csv > gpurun_out/r2_launch_summary_run4.txt
cat gpurun_out/r2_run4_bench.txt; head -8 gpurun_out/r2_launch_summary_run4.txt
