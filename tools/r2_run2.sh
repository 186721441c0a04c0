mkdir -p gpurun_out
for v in default x3s3 early; do
  if [ $v = default ]; then unset MMCM_LIB_PATH; else export MMCM_LIB_PATH=$PWD/build/libmmcm_$v.so; fi
  echo "=== $v" >> gpurun_out/r2_fold_bench.txt
  timeout 200 python tools/gemm_bench_fold.py 492 30 >> gpurun_out/r2_fold_bench.txt 2>&1
  timeout 200 python tools/gemm_bench_fold.py 1024 30 >> gpurun_out/r2_fold_bench.txt 2>&1
  timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/r2_bench_$v.json 2>> gpurun_out/r2_bench_$v.err
done
unset MMCM_LIB_PATH
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
  --log-file gpurun_out/r2_launches_fold.csv python bench.py --steps 1 --warmup 3 --batch 1024 --no-cpu-baseline --no-e2e \
  > gpurun_out/r2_ncu_fold.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_fold.csv > gpurun_out/r2_launch_summary_fold.txt
timeout 120 python tools/latency.py > gpurun_out/r2_latency.txt 2>&1
cat gpurun_out/r2_fold_bench.txt
for v in default x3s3 early; do python -c "
import json,sys
d=json.load(open('gpurun_out/r2_bench_$v.json')); print('$v', round(d['value']), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"; done
head -20 gpurun_out/r2_launch_summary_fold.txt; cat gpurun_out/r2_latency.txt
