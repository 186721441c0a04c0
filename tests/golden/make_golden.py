"""Generate golden vectors from the REAL reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

`/root/reference` does not travel to the GPU box, so this script imports the reference's own
`src.models` classes here (with the two adapters of SURVEY §8c: random-init-from-config instead of
`from_pretrained`, and `.pooler_output` for transformers>=5), loads the *seeded synthetic state dict*
(`mmcm_b200.synthetic.make_state_dict`, strict=True), runs the reference forward in fp32 on CPU on the
seeded edge-case batch and stores outputs.  Tests rebuild the same state dict and inputs from the seeds,
so the fixtures only need to hold outputs (a few KB each).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MMCM_REF_PATH", "/root/reference")
sys.path.insert(0, REF)

from __graft_entry__ import load_package  # noqa: E402

load_package()
from mmcm_b200 import arch as A, synthetic as syn  # noqa: E402

from transformers import (AutoModel, CLIPConfig, CLIPModel, CLIPTextModel, CLIPVisionModel,  # noqa: E402
                          SiglipConfig, SiglipModel)

# adapter 1: no network -> random init from config (the encoder name selects the vision patch size, like the hub would)
def _clip_config(name):
    return CLIPConfig(vision_config={"patch_size": 16}) if "patch16" in name else CLIPConfig()


CLIPModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(_clip_config(name)))
CLIPTextModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(_clip_config(name).text_config))
CLIPVisionModel.from_pretrained = classmethod(lambda cls, name, **kw: cls(_clip_config(name).vision_config))
AutoModel.from_pretrained = classmethod(
    lambda cls, name, **kw: SiglipModel(SiglipConfig(text_config={"vocab_size": 256000})))

from src.models import MultiModalFusionClassifier, MultiTaskClassifier  # noqa: E402  (the reference)

CASES = {
    # name: (kind, arch, ctor kwargs, weight seed, hardened, input seed, batch)
    "clip_fusion_hardened": ("fusion", A.CLIP_B32, dict(backend="clip"), 0, True, 7, 8),
    "clip_fusion_default": ("fusion", A.CLIP_B32, dict(backend="clip"), 0, False, 7, 8),
    "clip_mtl_h256_hardened": ("mtl", A.CLIP_B32, dict(head_hidden_dim=256), 1, True, 8, 8),
    "clip_mtl_h0_hardened": ("mtl", A.CLIP_B32, dict(head_hidden_dim=None), 2, True, 9, 8),
    "siglip_fusion_hardened": ("fusion", A.SIGLIP2_B16, dict(backend="siglip"), 3, True, 10, 8),
    "clip_b16_fusion_hardened": ("fusion", A.CLIP_B16, dict(backend="clip"), 4, True, 11, 8),
    # default random init = what BASELINE.json's absolute tolerances (2e-2 / 5e-3 / decisions) are calibrated for
    "clip_mtl_h256_default": ("mtl", A.CLIP_B32, dict(head_hidden_dim=256), 1, False, 8, 8),
    "clip_mtl_h0_default": ("mtl", A.CLIP_B32, dict(head_hidden_dim=None), 2, False, 9, 8),
    "siglip_fusion_default": ("fusion", A.SIGLIP2_B16, dict(backend="siglip"), 3, False, 10, 8),
}
TASKS = ["racist", "sexist", "homophobe", "religion", "otherhate"]


def run_case(name):
    kind, a, kw, wseed, hard, iseed, B = CASES[name]
    torch.manual_seed(0)
    feats = {}
    if kind == "fusion":
        m = MultiModalFusionClassifier("clip-patch%d" % a.patch, num_labels=5, **kw).eval()
        spec = A.fusion_spec(a, 5, 512)
        gt, gi = m.backbone.get_text_features, m.backbone.get_image_features

        def _t(**k):  # adapter 2 (transformers>=5 returns ModelOutput)
            o = gt(**k).pooler_output
            feats["text_feat"] = o.detach().clone()
            return o

        def _i(**k):
            o = gi(**k).pooler_output
            feats["vision_feat"] = o.detach().clone()
            return o
        m.backbone.get_text_features, m.backbone.get_image_features = _t, _i
    else:
        m = MultiTaskClassifier("x", TASKS, **kw).eval()
        spec = A.mtl_spec(a, 5, 512, kw.get("head_hidden_dim") or 0)
        et, ei = m._encode_text, m._encode_image

        def _t(i, mk):
            o = et(i, mk)
            feats["text_pooled"] = o.detach().clone()
            return o

        def _i(p):
            o = ei(p)
            feats["vision_pooled"] = o.detach().clone()
            return o
        m._encode_text, m._encode_image = _t, _i
    sd = syn.make_state_dict(spec, a, seed=wseed, hardened=hard)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    batch = syn.make_inputs(a, B, seed=iseed, edge_rows=True)
    labels = (torch.arange(B * 5).reshape(B, 5) % 3 == 0).float()
    with torch.no_grad():
        out = m(**batch, labels=labels)
    arrs = {"logits": out["logits"].numpy(), "loss": out["loss"].numpy(), "labels": labels.numpy()}
    for k, v in feats.items():
        arrs[k] = v.numpy()
    arrs["meta"] = np.array([wseed, int(hard), iseed, B])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, "logits std", float(out["logits"].std()), "range", float(out["logits"].min()),
          float(out["logits"].max()), "loss", float(out["loss"]))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or CASES):
        run_case(nm)
