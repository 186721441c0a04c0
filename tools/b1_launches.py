#!/usr/bin/env python
"""One B=1 forward inside an NVTX range (for `ncu --nvtx --nvtx-include "measure/" --metrics gpu__time_duration.sum`):
the per-kernel durations of the online path.  Dev tool."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

a = A.CLIP_B32
m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
m.load_state_dict(syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0))
m = m.to("cuda:0").eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
batch = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, B, seed=3).items()}
for _ in range(5):
    m(**batch)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("measure")
m(**batch)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
