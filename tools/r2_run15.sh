mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40) > gpurun_out/r2_t15.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/r2_t15.log | head -30
timeout 200 python tools/latency.py > gpurun_out/r2_latency15.txt 2>&1; grep "head_cluster 1" gpurun_out/r2_latency15.txt | head -8
timeout 300 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" --clock-control none --csv \
  --log-file gpurun_out/r2_b1_launches15.csv python tools/b1_launches.py 1 > gpurun_out/r2_b1_15.log 2>&1
python tools/launch_summary.py gpurun_out/r2_b1_launches15.csv > gpurun_out/r2_b1_summary15.txt; head -10 gpurun_out/r2_b1_summary15.txt
