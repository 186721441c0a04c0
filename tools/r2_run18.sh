mkdir -p gpurun_out
for B in 64 128 256 512; do echo "== B=$B fp32" >> gpurun_out/r2_e2e_sweep.txt; timeout 200 python tools/e2e_sweep.py $B f32 32,48,64,86,128,171,256,342 >> gpurun_out/r2_e2e_sweep.txt 2>&1; done
cat gpurun_out/r2_e2e_sweep.txt | grep -v Warn
