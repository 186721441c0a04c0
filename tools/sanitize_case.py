#!/usr/bin/env python
"""Smallest end-to-end cases for compute-sanitizer (memcheck / racecheck): one CLIP-Fusion, one CLIP-MTL and one
SigLIP-Fusion forward at B=8 with the edge rows, packed and dense text, both GEMM kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "clip"
if which == "clip":
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0, hardened=True)
    m = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=5)
elif which == "mtl":
    a = A.CLIP_B32
    sd = syn.make_state_dict(A.mtl_spec(a, 5, 512, 256), a, seed=0, hardened=True)
    m = P.MultiTaskClassifier("openai/clip-vit-base-patch32", ["a", "b", "c", "d", "e"], head_hidden_dim=256)
else:
    a = A.SIGLIP2_B16
    sd = syn.make_state_dict(A.fusion_spec(a, 5, 512), a, seed=0, hardened=True)
    m = P.MultiModalFusionClassifier("google/siglip2-base-patch16-224", num_labels=5, backend="siglip")
m.load_state_dict(sd)
m = m.to("cuda:0").eval()
batch = {k: v.to("cuda:0") for k, v in syn.make_inputs(a, 9, seed=1, edge_rows=True).items()}
for varlen in (1, 0):
    for impl in (0, 2):
        m.set_option("varlen_text", varlen)
        m.set_option("gemm_impl", impl)
        y = m(**batch)["logits"]
        torch.cuda.synchronize()
        print(which, "varlen", varlen, "gemm_impl", impl, "finite", bool(torch.isfinite(y).all()), y[0].tolist())
