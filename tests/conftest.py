"""Shared fixtures.  `-m "not gpu"` runs in the CPU-only build container; `-m gpu` runs on a B200."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

load_package()  # registers multimodal-content-moderation_b200/ as `mmcm_b200`


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name: (kind, arch name, ctor kwargs, weight seed, hardened, input seed, batch) -- must match make_golden.py
GOLDEN_CASES = {
    "clip_fusion_hardened": ("fusion", "CLIP_B32", dict(backend="clip"), 0, True, 7, 8),
    "clip_fusion_default": ("fusion", "CLIP_B32", dict(backend="clip"), 0, False, 7, 8),
    "clip_mtl_h256_hardened": ("mtl", "CLIP_B32", dict(head_hidden_dim=256), 1, True, 8, 8),
    "clip_mtl_h0_hardened": ("mtl", "CLIP_B32", dict(head_hidden_dim=None), 2, True, 9, 8),
    "siglip_fusion_hardened": ("fusion", "SIGLIP2_B16", dict(backend="siglip"), 3, True, 10, 8),
    "clip_b16_fusion_hardened": ("fusion", "CLIP_B16", dict(backend="clip"), 4, True, 11, 8),
    # default random init: the official absolute gate (logits 2e-2, probabilities 5e-3, decisions) for every model
    "clip_mtl_h256_default": ("mtl", "CLIP_B32", dict(head_hidden_dim=256), 1, False, 8, 8),
    "clip_mtl_h0_default": ("mtl", "CLIP_B32", dict(head_hidden_dim=None), 2, False, 9, 8),
    "siglip_fusion_default": ("fusion", "SIGLIP2_B16", dict(backend="siglip"), 3, False, 10, 8),
}
TASKS = ["racist", "sexist", "homophobe", "religion", "otherhate"]


def build_case(name):
    """Rebuild (arch, state dict, batch, golden arrays) of one golden case from its seeds."""
    import numpy as np
    from mmcm_b200 import arch as A, synthetic as syn
    kind, arch_name, kw, wseed, hard, iseed, B = GOLDEN_CASES[name]
    a = getattr(A, arch_name)
    if kind == "fusion":
        spec = A.fusion_spec(a, 5, 512)
    else:
        spec = A.mtl_spec(a, 5, 512, kw.get("head_hidden_dim") or 0)
    sd = syn.make_state_dict(spec, a, seed=wseed, hardened=hard)
    batch = syn.make_inputs(a, B, seed=iseed, edge_rows=True)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return kind, a, kw, sd, batch, gold


def oracle_forward(kind, a, sd, batch, stages=None):
    from oracle import scoring_oracle as orc
    from mmcm_b200 import arch as A
    if kind == "fusion":
        return orc.fusion_forward(sd, batch, "clip" if a.backend == A.BACKEND_CLIP else "siglip", a.patch, a.eos_id,
                                  stages=stages)
    return orc.mtl_forward(sd, batch, a.patch, a.eos_id, stages=stages)
