mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -q -m gpu --tb=short 2>&1 | tail -40) > gpurun_out/r2_t21.log 2>&1
grep -n "^E \|FAILED\|passed\|failed" gpurun_out/r2_t21.log | head -30
python - <<PY
import os, sys, time, torch
sys.path.insert(0, ".")
from __graft_entry__ import load_package
P = load_package()
from mmcm_b200 import arch as A, synthetic as syn
a = A.CLIP_B32
m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
m.load_state_dict(syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0)); m = m.to("cuda:0").eval()
base = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, 1, seed=3).items()}
for name, tp, ip in (("text+image", 1., 1.), ("text only", 1., 0.), ("image only", 0., 1.)):
    b = dict(base); b["text_present"] = torch.full((1,), tp, device="cuda:0"); b["image_present"] = torch.full((1,), ip, device="cuda:0")
    for _ in range(10): m(**b)
    torch.cuda.synchronize(); lat = []
    for _ in range(40):
        t0 = time.perf_counter(); m(**b)["logits"].cpu(); lat.append((time.perf_counter() - t0) * 1e3)
    print(f"B=1 {name:11s} latency {sorted(lat)[20]:.3f} ms  launches {m._engine.last_launch_count()}")
PY
