"""Checkpoint directory -> scoring model.

The reference rebuilds its model from `inference_config.json` + `model.safetensors` in
`load_model_from_checkpoint` (R/scripts/evaluate.py:89-160; the same steps in R/scripts/inference.py:90-140 and
R/sagemaker/inference.py:60-130).  Two gaps at that boundary are documented in SURVEY 8b: `train.py` never writes
`"head"` into the config (so an MTL checkpoint is rebuilt as a Fusion model and `load_state_dict` fails), and the
callers never pass `head_hidden_dim` although `config/clip_mtl.yaml` trains with 256.  `load_checkpoint` keeps the
reference's file search order and error types but decides head type, task count, `head_hidden_dim` and
`learnable_task_weights` from the state-dict keys themselves, so every checkpoint the reference can write loads.

`PackedScorer` is the second half of the row: the repacked bf16 weight set as one blob (`mmcm_save_packed`), loaded
without the fp32 checkpoint, without repack kernels and without an fp32 master copy in the process.
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch

from . import arch as A
from .engine import Engine

CONFIG_NAMES = ("inference_config.json", "config.json")
PACKED_NAME = "mmcm_packed_sm100a.bin"


def find_config(checkpoint_dir) -> Dict:
    """evaluate.py:93-110: `<parent>/inference_config.json`, `<dir>/inference_config.json`, `<parent>/config.json`."""
    d = Path(checkpoint_dir)
    for cand in (d.parent / CONFIG_NAMES[0], d / CONFIG_NAMES[0], d.parent / CONFIG_NAMES[1]):
        if cand.exists():
            with open(cand, "r", encoding="utf-8") as f:
                return json.load(f)
    raise FileNotFoundError(f"Could not find inference_config.json or config.json in {checkpoint_dir} or parent")


def read_state_dict(checkpoint_dir) -> Dict[str, torch.Tensor]:
    """evaluate.py:139-151: `model.safetensors`, else `pytorch_model.bin`."""
    d = Path(checkpoint_dir)
    st = d / "model.safetensors"
    if st.exists():
        from safetensors.torch import load_file
        return load_file(str(st))
    pt = d / "pytorch_model.bin"
    if pt.exists():
        return torch.load(str(pt), map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"Could not find model weights in {checkpoint_dir}")


def infer_model_spec(state_dict: Dict[str, torch.Tensor], config: Dict) -> Dict:
    """Everything the constructors need, from the checkpoint itself.  `config["head"]`, when present, must agree."""
    keys = state_dict.keys()
    is_mtl = any(k.startswith(("tower_txt.", "tower_img.", "shared_head.", "heads.")) for k in keys)
    is_fusion = any(k.startswith(("backbone.", "cls.", "ln_fused.")) for k in keys)
    if is_mtl == is_fusion:
        raise ValueError("state dict is neither a MultiModalFusionClassifier nor a MultiTaskClassifier checkpoint")
    head = "mtl" if is_mtl else "fusion"
    declared = config.get("head")
    if declared is not None and declared != head:
        raise ValueError(f"inference config says head={declared!r} but the weights are a {head!r} checkpoint")
    spec = {"head": head, "encoder_name": config.get("encoder_name", "openai/clip-vit-base-patch32"),
            "backend": config.get("backend", "clip"), "fusion_dim": int(config.get("fusion_dim", 512)),
            "class_names": list(config.get("class_names", ["harmful"]))}
    if "proj_t.weight" in state_dict:
        spec["fusion_dim"] = int(state_dict["proj_t.weight"].shape[0])
    if head == "fusion":
        n = int(state_dict["cls.4.weight"].shape[0])
    else:
        idx = {int(m.group(1)) for m in (re.match(r"heads\.(\d+)\.", k) for k in keys) if m}
        n = len(idx)
        if idx != set(range(n)):
            raise ValueError("MTL checkpoint has non-contiguous task heads")
        hidden = state_dict.get("heads.0.0.weight")            # Linear(fd, h) -> GELU -> Dropout -> Linear(h, 1)
        spec["head_hidden_dim"] = int(hidden.shape[0]) if hidden is not None else None
        spec["learnable_task_weights"] = "log_vars" in state_dict
    # training-only buffers a strict load must find a home for (fusion.py:131-137, multitask.py:105-110)
    spec["has_pos_weight"] = "pos_weight" in state_dict
    spec["has_focal_alpha"] = "criterion.alpha" in state_dict
    if len(spec["class_names"]) != n:
        raise ValueError(f"config lists {len(spec['class_names'])} class names, the checkpoint has {n} outputs")
    return spec


def build_model(spec: Dict) -> torch.nn.Module:
    from .modules import MultiModalFusionClassifier, MultiTaskClassifier
    n = len(spec["class_names"])
    pw = torch.ones(n) if spec.get("has_pos_weight") else None          # placeholder, overwritten by load_state_dict
    if spec["head"] == "mtl":
        return MultiTaskClassifier(spec["encoder_name"], spec["class_names"], fusion_dim=spec["fusion_dim"],
                                   backend=spec["backend"], pos_weight=pw, head_hidden_dim=spec.get("head_hidden_dim"),
                                   learnable_task_weights=spec.get("learnable_task_weights", False))
    focal = spec.get("has_focal_alpha", False)
    return MultiModalFusionClassifier(spec["encoder_name"], num_labels=n, fusion_dim=spec["fusion_dim"],
                                      backend=spec["backend"], pos_weight=pw, loss_type="focal" if focal else "bce",
                                      alpha_focal=torch.ones(n) if focal else None)


def load_checkpoint(checkpoint_dir, device="cuda:0") -> Tuple[torch.nn.Module, Dict]:
    """-> (model on `device` in eval mode, config with "head" / "thresholds" filled in)."""
    config = dict(find_config(checkpoint_dir))
    sd = read_state_dict(checkpoint_dir)
    spec = infer_model_spec(sd, config)
    model = build_model(spec)
    model.load_state_dict(sd, strict=True)
    model = model.to(device).eval()
    config.update(head=spec["head"], class_names=spec["class_names"], fusion_dim=spec["fusion_dim"])
    config.setdefault("thresholds", [0.5] * len(spec["class_names"]))
    return model, config


class PackedScorer:
    """The forward of a checkpoint from its packed weight file: `{"loss": None, "logits": ...}` like the modules, no
    nn.Parameters behind it.  Write the file once with `PackedScorer.pack(model, path)` (or `model.save_packed`)."""

    def __init__(self, path, device="cuda:0"):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the B200 scoring path has no CPU fallback")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.engine = Engine.from_packed(str(path), self.device.index)

    @staticmethod
    def pack(model: torch.nn.Module, path) -> None:
        model.save_packed(str(path))

    def set_option(self, name: str, value: int) -> None:
        self.engine.set_option(name, value)

    @torch.no_grad()
    def __call__(self, input_ids, attention_mask, pixel_values, text_present, image_present, labels=None):
        if not input_ids.is_cuda or input_ids.device != self.device:
            raise RuntimeError("inputs must be CUDA tensors on the scorer's device: no CPU fallback")
        with torch.cuda.device(self.device):
            logits = self.engine.forward(input_ids, attention_mask, pixel_values, text_present, image_present)
        return {"loss": None, "logits": logits}

    forward = __call__
