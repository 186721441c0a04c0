"""Deterministic synthetic weights and inputs (there are no checkpoints or images offline).

BASELINE.json's configs are all "random-init, synthetic batch".  The reference gets its random
weights from Hugging Face's `_init_weights`; we reproduce the same *distributions*
(HF/models/clip/modeling_clip.py `CLIPPreTrainedModel._init_weights`, torch `nn.Linear.reset_parameters`)
with our own seeded generator so that the identical state dict can be rebuilt on the GPU box, where
`/root/reference` does not exist.  The state dicts produced here load with `strict=True` into the
reference's own classes (tests/golden/make_golden.py does exactly that).

Two flavours:
  * "default"  : biases 0, LayerNorm gamma=1/beta=0  -> logits almost input independent; the
                 official absolute tolerances (2e-2 / 5e-3) are calibrated for this.
  * "hardened" : every bias ~N(0,.02), LN gamma~N(1,.1), beta~N(0,.1), head matrices x4 -> O(1) logits that
                 straddle 0, so bias / affine / masking bugs become visible.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import arch as A


def _std_for(key: str, shape, a: A.ArchCfg) -> Optional[float]:
    """Init std of one backbone tensor, or None for 'not a normal-initialised matrix'."""
    tower = a.text if ".text_model." in key or key.startswith("tower_txt") else a.vision
    d, L, f = tower.hidden, tower.layers, tower.ffn
    clip = a.backend == A.BACKEND_CLIP
    if key.endswith("token_embedding.weight") or key.endswith("position_embedding.weight"):
        return 0.02 if clip else d ** -0.5
    if key.endswith("class_embedding"):
        return d ** -0.5
    if key.endswith("patch_embedding.weight"):
        return 0.02 if clip else (3 * a.patch * a.patch) ** -0.5
    if key.endswith("head.probe"):
        return d ** -0.5
    if key.endswith("in_proj_weight"):
        return math.sqrt(2.0 / (4 * d))
    if ".self_attn." in key and key.endswith(".weight"):
        if clip:
            in_std = (d ** -0.5) * ((2 * L) ** -0.5)
            return d ** -0.5 if ".out_proj." in key else in_std
        return d ** -0.5
    if key.endswith("attention.out_proj.weight"):
        return d ** -0.5
    if key.endswith("mlp.fc1.weight"):
        return (2 * d) ** -0.5 if clip else math.sqrt(2.0 / (d + f))
    if key.endswith("mlp.fc2.weight"):
        return (d ** -0.5) * ((2 * L) ** -0.5) if clip else math.sqrt(2.0 / (d + f))
    if key.endswith("visual_projection.weight") or key.endswith("text_projection.weight"):
        return shape[1] ** -0.5
    if key.endswith("text_model.head.weight"):
        return shape[1] ** -0.5
    return None


def make_state_dict(spec: A.Spec, a: A.ArchCfg, seed: int = 0, hardened: bool = False,
                    dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded CPU state dict for `spec` (arch.fusion_spec / arch.mtl_spec)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in spec:
        backbone = key.startswith(("backbone.", "tower_txt.", "tower_img."))
        n = 1
        for s_ in shape:
            n *= s_
        if key.endswith("logit_scale"):
            t = torch.full(shape, 2.6592 if a.backend == A.BACKEND_CLIP else 0.0)
        elif key.endswith("logit_bias"):
            t = torch.zeros(shape)
        elif _is_layernorm(key):
            if key.endswith(".weight"):
                t = torch.ones(shape)
                if hardened:
                    t = t + 0.1 * torch.randn(shape, generator=g)
            else:
                t = torch.zeros(shape)
                if hardened:
                    t = 0.1 * torch.randn(shape, generator=g)
        elif backbone:
            std = _std_for(key, shape, a)
            if std is not None:
                t = torch.randn(shape, generator=g) * std
            else:  # backbone biases
                t = torch.randn(shape, generator=g) * (0.02 if hardened else 0.0)
        else:
            # reference heads are plain nn.Linear: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
            fan_in = shape[-1] if len(shape) == 2 else None
            if fan_in is None:  # bias: bound from the sibling weight's fan_in -> use own length heuristically
                bound = 0.04
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
                if hardened:
                    t = t + 0.02 * torch.randn(shape, generator=g)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
                if hardened:
                    t = t * 4.0
        sd[key] = t.to(dtype).contiguous()
    return sd


def _is_layernorm(key: str) -> bool:
    parts = key.split(".")
    nm = parts[-2] if len(parts) >= 2 else ""
    if nm in ("layer_norm1", "layer_norm2", "final_layer_norm", "pre_layrnorm", "post_layernorm",
              "layernorm", "ln_fused"):
        return True
    return key in ("cls.0.weight", "cls.0.bias")


# --------------------------------------------------------------------------------------------
# inputs (SURVEY §8d "Synthetic inputs")
# --------------------------------------------------------------------------------------------
def make_inputs(a: A.ArchCfg, batch: int, seed: int = 1234, edge_rows: bool = False,
                device: str = "cpu") -> Dict[str, torch.Tensor]:
    """Batch dict with the reference's collate_fn keys (R/src/data/dataset.py:171-193).

    CLIP text: BOS, random tokens, EOS then EOS-valued padding (pad token == eos for the CLIP tokenizer);
    SigLIP text: random tokens right padded with 0.  `edge_rows=True` overwrites the first rows with the
    corner cases of SURVEY §3.6 (absent modalities, no-EOS row, EOS at position 1, full-length text).
    """
    g = torch.Generator().manual_seed(seed)
    S = a.max_pos
    if a.backend == A.BACKEND_CLIP:
        ids = torch.randint(1, a.eos_id - 1, (batch, S), generator=g)
        lens = torch.randint(3, S + 1, (batch,), generator=g)
        ids[:, 0] = a.eos_id - 1  # BOS = 49406
        pos = torch.arange(S)[None, :]
        ids = torch.where(pos >= (lens[:, None] - 1), torch.full_like(ids, a.eos_id), ids)
    else:
        ids = torch.randint(2, a.vocab, (batch, S), generator=g)
        lens = torch.randint(3, S + 1, (batch,), generator=g)
        pos = torch.arange(S)[None, :]
        ids = torch.where(pos >= lens[:, None], torch.zeros_like(ids), ids)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    px = torch.randn(batch, 3, a.image, a.image, generator=g)
    tp = torch.ones(batch)
    ip = torch.ones(batch)
    if edge_rows:
        if batch < 8:
            raise ValueError("edge_rows needs batch >= 8")
        tp[0] = 0.0                       # text absent
        ip[1] = 0.0                       # image absent
        tp[2] = 0.0; ip[2] = 0.0          # both absent
        if a.backend == A.BACKEND_CLIP:
            ids[3] = torch.randint(1, a.eos_id - 1, (S,), generator=g)   # no EOS anywhere -> pooled row 0
            mask[3] = 1
            ids[4, 1:] = a.eos_id; mask[4] = 0; mask[4, :2] = 1          # [BOS, EOS, pad...] -> row 1
            ids[5, 1:S - 1] = torch.randint(1, a.eos_id - 1, (S - 2,), generator=g)
            ids[5, S - 1] = a.eos_id; mask[5] = 1                        # full length, EOS at S-1
            mask[6] = 1                                                  # all-ones mask over padded ids
        else:
            mask[3] = 0                                                  # all-pad text: fully masked rows
            mask[4] = 1                                                  # no padding at all
            mask[5] = 0; mask[5, :1] = 1                                 # single valid token
    out = {"input_ids": ids, "attention_mask": mask, "pixel_values": px,
           "text_present": tp, "image_present": ip}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out
