#!/usr/bin/env python
"""Measured logit error (as % of the batch logit std, hardened init) of the LN-fold and the separate-pass path against
the CPU oracle, over several input seeds: the numbers behind REL_GATE in tests/test_gpu_forward.py.  Dev tool."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TASKS, build_case, oracle_forward  # noqa: E402
import mmcm_b200 as P  # noqa: E402
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

for name in ("clip_fusion_hardened", "clip_mtl_h256_hardened", "clip_mtl_h0_hardened", "siglip_fusion_hardened"):
    kind, a, kw, sd, _, _ = build_case(name)
    enc = "openai/clip-vit-base-patch32" if a.backend == A.BACKEND_CLIP else "google/siglip2-base-patch16-224"
    m = P.MultiModalFusionClassifier(enc, num_labels=5, **kw) if kind == "fusion" else P.MultiTaskClassifier(enc, TASKS, **kw)
    m.load_state_dict(sd)
    m = m.to("cuda:0").eval()
    rows = []
    for seed in range(5 if a.backend == A.BACKEND_CLIP else 2):
        batch = syn.make_inputs(a, 24, seed=800 + seed, edge_rows=True)
        with torch.no_grad():
            ref = oracle_forward(kind, a, sd, batch)
        d = {k: v.to("cuda:0") for k, v in batch.items()}
        out = []
        for fold in (1, 0):
            m.set_option("ln_fold", fold)
            y = m(**d)["logits"].cpu()
            out.append(100 * (y - ref).abs().max().item() / ref.std().item())
        rows.append(out)
    print(name, "fold / separate, % of std per seed:", [f"{r[0]:.2f}/{r[1]:.2f}" for r in rows], flush=True)
    del m
