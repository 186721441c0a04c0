mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_t3.log 2>&1
timeout 400 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
timeout 300 python bench.py --impl reference > gpurun_out/r2_bench3_ref.json 2> gpurun_out/r2_bench3_ref.err
timeout 300 python tools/gemm_bench_fold.py 492 30 > gpurun_out/r2_fold_bench2.txt 2>&1
# per-shape ncu --set full of the encoder GEMMs (NVTX-ranged launches)
timeout 600 python tools/ncu_gemm_shapes.py > gpurun_out/r02_gemm_shapes_plain.log 2>&1
timeout 900 ncu --set full --import-source on --nvtx --nvtx-include "measure/" --clock-control none -f -o gpurun_out/r02_gemm_shapes \
  python tools/ncu_gemm_shapes.py > gpurun_out/r02_gemm_shapes.log 2>&1
# DRAM traffic of every encoder GEMM launch of one bench forward
MMCM_NCU_RANGE=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --nvtx --nvtx-include "measure/" -k regex:gemm2 --clock-control none -f -o gpurun_out/r02_bench_gemms \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_gemms.log 2>&1
tail -4 gpurun_out/r2_t3.log
python -c "
import json
d=json.load(open('gpurun_out/r2_bench3.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'roof',round(d['roofline']['frac'],3),'par',d['parity'],'cpu',d['cpu_baseline']['value'])
print(d['extras']); print(d['clocks']); print(d['e2e'])
r=json.load(open('gpurun_out/r2_bench3_ref.json')); print('ref', r['value'], r['cpu_baseline']['sample'])
"
cat gpurun_out/r2_fold_bench2.txt; ls -la gpurun_out/*.ncu-rep | tail -3
