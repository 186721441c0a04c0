mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25) > gpurun_out/r2_t10.log 2>&1
tail -6 gpurun_out/r2_t10.log
timeout 200 python tools/latency.py > gpurun_out/r2_latency10.txt 2>&1; grep "head_cluster 1" gpurun_out/r2_latency10.txt | head -8
timeout 300 ncu --metrics gpu__time_duration.sum --nvtx --nvtx-include "measure/" --clock-control none --csv \
  --log-file gpurun_out/r2_b1_launches10.csv python tools/b1_launches.py 1 > gpurun_out/r2_b1_10.log 2>&1
python tools/launch_summary.py gpurun_out/r2_b1_launches10.csv | head -12
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
