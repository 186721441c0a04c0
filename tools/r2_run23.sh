for c in 0 1; do echo "== MMCM_CARVEOUT=$c"; MMCM_CARVEOUT=$c timeout 120 python tools/latency.py 0 2 2>&1 | grep "ln_fold 1 head_cluster 1" | head -3; done
MMCM_CARVEOUT=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench carveout=1', round(d['value']), d['clocks']['sm_mhz'])"
timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 20 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench carveout=0', round(d['value']), d['clocks']['sm_mhz'])"
