cat > /tmp/e2e_only.py <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
import mmcm_b200 as P
from mmcm_b200 import synthetic as syn, arch as A
a = A.CLIP_B32
torch.manual_seed(0)
m = P.MultiModalFusionClassifier(encoder_name="openai/clip-vit-base-patch32", num_labels=5).cuda().eval() if hasattr(P, "MultiModalFusionClassifier") else None
PY
for s in "" "342,682" "256,768" "400,624" "512,512" "300,724" "342,342,340"; do
MMCM_HOST_SPLIT=$s timeout 300 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r2_split.json 2> gpurun_out/r2_split.err; python -c "
import json; d=json.load(open('gpurun_out/r2_split.json')); print('split=[$s]', 'e2e', round(d['e2e']['value']), 'value', round(d['value']), d['clocks']['sm_mhz'])"
done
