#!/usr/bin/env python
"""Small-batch latency of the drop-in forward (the B=1 online path of scripts/inference.py / sagemaker predict_fn)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

a = A.CLIP_B32
sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0)
m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
m.load_state_dict(sd)
m = m.to("cuda:0").eval()
gmax = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m.set_option("graph_max_batch", gmax)
if len(sys.argv) > 2:
    m.set_option("streams", int(sys.argv[2]))
print("graph_max_batch", gmax)
# (gemm_impl, ln_fold, head_cluster): pair kernel + LN fold (default), single-CTA 128-row tiles, old single-CTA head
for impl, fold, hc in ((0, 1, 1), (0, 1, 0), (0, 0, 1), (2, 0, 1)):
    m.set_option("gemm_impl", impl)
    m.set_option("ln_fold", fold)
    m.set_option("head_cluster", hc)
    for B in (1, 8, 32, 64):
        batch = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, B, seed=3).items()}
        for _ in range(5):
            m(**batch)
        torch.cuda.synchronize()
        n = 30
        t0 = time.perf_counter()
        for _ in range(n):
            y = m(**batch)["logits"]
            y.cpu()                         # the callers' per-request D2H
        dt = (time.perf_counter() - t0) / n
        print(f"gemm_impl {impl} ln_fold {fold} head_cluster {hc}  B={B:4d}  latency {dt * 1e3:7.3f} ms  "
              f"({B / dt:8.0f} samples/s)  launches {m._engine.last_launch_count()}")
