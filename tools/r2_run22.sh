mkdir -p gpurun_out
for cfg in "0 2" "0 1" "64 2" "64 1"; do set -- $cfg; echo "== graph_max_batch $1 streams $2"; timeout 120 python tools/latency.py $1 $2 2>&1 | grep "ln_fold 1 head_cluster 1" | head -2; done
