mkdir -p gpurun_out
timeout 300 python tools/e2e_sweep.py 1024 f32 0,342 0,48,96,128,171 > gpurun_out/r2_e2e_first.txt 2>&1
grep -v Warn gpurun_out/r2_e2e_first.txt
timeout 200 python tools/e2e_sweep.py 256 f32 0 0 >> gpurun_out/r2_e2e_first.txt 2>&1; tail -1 gpurun_out/r2_e2e_first.txt
