"""CPU ORACLE (test infrastructure only -- never imported by the product package).

A plain fp32 restatement of the reference's scoring path, written as straight-line tensor algebra
over a *flat state dict with the reference's key names*.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import this file.

What it restates and where that arithmetic lives:

  * heads (in /root/reference):
      MultiModalFusionClassifier.forward     R/src/models/fusion.py:157-229
      MultiTaskClassifier.forward            R/src/models/multitask.py:156-227
  * encoders (third-party dependency `transformers`, requirement `>=4.35.0` at R/requirements.txt:4,
    un-pinned; restated from transformers 5.5.0 as installed in this image):
      CLIPVisionEmbeddings.forward           HF/models/clip/modeling_clip.py:202-218
      CLIPTextEmbeddings.forward             HF/models/clip/modeling_clip.py:234-258
      eager_attention_forward semantics      HF/models/clip/modeling_clip.py:261-279
      CLIPAttention / CLIPMLP / EncoderLayer HF/models/clip/modeling_clip.py:300-336,347-351,363-384
      CLIPTextTransformer tail (EOS pooling) HF/models/clip/modeling_clip.py:561-589
      CLIPVisionTransformer.forward          HF/models/clip/modeling_clip.py:667-691
      get_text_features/get_image_features   HF/models/clip/modeling_clip.py:793-863
      Siglip embeddings / text tail / MAP    HF/models/siglip/modeling_siglip.py:175-185,489-527,604-649
      activations                            HF/activations.py:45,122-123
      mask semantics                         HF/masking_utils.py:882-1085 as measured in SURVEY §3.3/§3.5/§3.6

PARITY PINNING: the reference ships no tests and no golden vectors.  The oracle is pinned
against outputs of the reference itself run in the build container: tests/golden/make_golden.py imports
`/root/reference/src/models` (+ transformers), loads the same seeded state dict and stores logits /
pooled features in tests/golden/*.npz; tests/test_oracle.py checks this file against those fixtures and,
because `transformers` is part of the image, also directly against HF's CLIP / SigLIP modules.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# ------------------------------------------------------------------------------------------ pieces
def _ln(x: Tensor, sd: SD, p: str, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _act(x: Tensor, kind: str) -> Tensor:
    if kind == "quick_gelu":          # HF/activations.py:122-123
        return x * torch.sigmoid(1.702 * x)
    if kind == "gelu_pytorch_tanh":   # HF/activations.py:45
        return F.gelu(x, approximate="tanh")
    raise ValueError(kind)


def _attention(q: Tensor, k: Tensor, v: Tensor, heads: int, allow: Optional[Tensor]) -> Tensor:
    """softmax_fp32(q k^T / sqrt(dh) + mask) v ; `allow` is bool [B,1,Tq,Tk] (True = may attend).

    A query row whose keys are all masked yields exactly 0 (torch>=2.5 SDPA safe-softmax).
    """
    B, Tq, D = q.shape
    Tk = k.shape[1]
    dh = D // heads
    qh = q.view(B, Tq, heads, dh).transpose(1, 2)
    kh = k.view(B, Tk, heads, dh).transpose(1, 2)
    vh = v.view(B, Tk, heads, dh).transpose(1, 2)
    s = torch.matmul(qh, kh.transpose(-1, -2)) * (dh ** -0.5)
    if allow is not None:
        s = s.masked_fill(~allow, float("-inf"))
        dead = ~allow.any(dim=-1, keepdim=True)
        s = s.masked_fill(dead, 0.0)
        p = torch.softmax(s, dim=-1).masked_fill(dead, 0.0)
    else:
        p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, vh)
    return o.transpose(1, 2).reshape(B, Tq, D)


def _encoder(x: Tensor, sd: SD, prefix: str, layers: int, heads: int, eps: float, act: str,
             allow: Optional[Tensor], stages: Optional[dict] = None, tag: str = "") -> Tensor:
    """Pre-LN residual blocks, HF/models/clip/modeling_clip.py:363-384 (SigLIP :340-361)."""
    for i in range(layers):
        p = f"{prefix}encoder.layers.{i}."
        h = _ln(x, sd, p + "layer_norm1", eps)
        q = _lin(h, sd, p + "self_attn.q_proj")
        k = _lin(h, sd, p + "self_attn.k_proj")
        v = _lin(h, sd, p + "self_attn.v_proj")
        a = _attention(q, k, v, heads, allow)
        x = x + _lin(a, sd, p + "self_attn.out_proj")
        h = _ln(x, sd, p + "layer_norm2", eps)
        h = _act(_lin(h, sd, p + "mlp.fc1"), act)
        x = x + _lin(h, sd, p + "mlp.fc2")
        if stages is not None:
            stages[f"{tag}layer{i}"] = x
    return x


def _num_layers(sd: SD, prefix: str) -> int:
    """Depth of the tower stored under `prefix` (so that truncated-depth test models need no extra arguments)."""
    n = 0
    while f"{prefix}encoder.layers.{n}.layer_norm1.weight" in sd:
        n += 1
    return n


# ------------------------------------------------------------------------------------------ CLIP
def clip_text_pooled(sd: SD, prefix: str, ids: Tensor, mask: Optional[Tensor], eos_id: int = 49407,
                     heads: int = 8, layers: Optional[int] = None, eps: float = 1e-5,
                     stages: Optional[dict] = None) -> Tensor:
    """CLIPTextTransformer.forward -> pooler_output (EOS row after final LN). prefix ends with 'text_model.'"""
    layers = _num_layers(sd, prefix) if layers is None else layers
    B, S = ids.shape
    pos_w = sd[prefix + "embeddings.position_embedding.weight"]
    if S > pos_w.shape[0]:  # HF/models/clip/modeling_clip.py:243-247
        raise ValueError(
            f"Sequence length must be less than max_position_embeddings (got `sequence length`: {S} and "
            f"max_position_embeddings: {pos_w.shape[0]}")
    x = sd[prefix + "embeddings.token_embedding.weight"][ids] + pos_w[:S][None]
    if stages is not None:
        stages["text_embed"] = x
    causal = torch.ones(S, S, dtype=torch.bool).tril()[None, None]
    allow = causal if mask is None else causal & (mask[:, None, None, :] != 0)
    x = _encoder(x, sd, prefix, layers, heads, eps, "quick_gelu", allow, stages, "text_")
    x = _ln(x, sd, prefix + "final_layer_norm", eps)
    if eos_id == 2:   # legacy branch, HF/models/clip/modeling_clip.py:564-574
        idx = ids.to(torch.int).argmax(dim=-1)
    else:             # first position equal to eos, 0 if none, :575-584
        idx = (ids.to(torch.int) == eos_id).int().argmax(dim=-1)
    return x[torch.arange(B), idx]


def clip_vision_pooled(sd: SD, prefix: str, px: Tensor, patch: int = 32, heads: int = 12,
                       layers: Optional[int] = None, eps: float = 1e-5, stages: Optional[dict] = None) -> Tensor:
    """CLIPVisionTransformer.forward -> pooler_output = post_layernorm(CLS). prefix ends with 'vision_model.'"""
    layers = _num_layers(sd, prefix) if layers is None else layers
    w = sd[prefix + "embeddings.patch_embedding.weight"]
    B, _, H, W = px.shape
    pos_w = sd[prefix + "embeddings.position_embedding.weight"]
    n_tok = (H // patch) * (W // patch) + 1
    if n_tok != pos_w.shape[0] or H != W:  # HF/models/clip/modeling_clip.py:204-207
        side = int(math.isqrt(pos_w.shape[0] - 1)) * patch
        raise ValueError(f"Input image size ({H}*{W}) doesn't match model ({side}*{side}).")
    pe = F.conv2d(px, w, None, stride=patch).flatten(2).transpose(1, 2)
    cls = sd[prefix + "embeddings.class_embedding"].expand(B, 1, -1)
    x = torch.cat([cls, pe], dim=1) + pos_w[None]
    if stages is not None:
        stages["vision_embed"] = x
    x = _ln(x, sd, prefix + "pre_layrnorm", eps)
    x = _encoder(x, sd, prefix, layers, heads, eps, "quick_gelu", None, stages, "vision_")
    return _ln(x[:, 0], sd, prefix + "post_layernorm", eps)


# ------------------------------------------------------------------------------------------ SigLIP
def siglip_text_pooled(sd: SD, prefix: str, ids: Tensor, mask: Optional[Tensor], heads: int = 12,
                       layers: Optional[int] = None, eps: float = 1e-6, stages: Optional[dict] = None) -> Tensor:
    """SiglipTextTransformer.forward -> pooler_output = head(final_LN(x)[:, -1]); bidirectional key mask."""
    layers = _num_layers(sd, prefix) if layers is None else layers
    B, S = ids.shape
    pos_w = sd[prefix + "embeddings.position_embedding.weight"]
    if S > pos_w.shape[0]:  # HF/models/siglip/modeling_siglip.py:211-215
        raise ValueError(
            f"Sequence length must be less than max_position_embeddings (got `sequence length`: {S} and "
            f"max_position_embeddings: {pos_w.shape[0]}")
    x = sd[prefix + "embeddings.token_embedding.weight"][ids] + pos_w[:S][None]
    allow = None
    if mask is not None and not bool((mask != 0).all()):
        allow = (mask[:, None, None, :] != 0).expand(B, 1, S, S)
    x = _encoder(x, sd, prefix, layers, heads, eps, "gelu_pytorch_tanh", allow, stages, "text_")
    x = _ln(x, sd, prefix + "final_layer_norm", eps)
    return _lin(x[:, -1], sd, prefix + "head")


def siglip_vision_pooled(sd: SD, prefix: str, px: Tensor, patch: int = 16, heads: int = 12,
                         layers: Optional[int] = None, eps: float = 1e-6, stages: Optional[dict] = None) -> Tensor:
    """SiglipVisionTransformer.forward -> MAP-pooled feature (HF/models/siglip/modeling_siglip.py:604-649)."""
    layers = _num_layers(sd, prefix) if layers is None else layers
    w = sd[prefix + "embeddings.patch_embedding.weight"]
    b = sd[prefix + "embeddings.patch_embedding.bias"]
    B = px.shape[0]
    pos_w = sd[prefix + "embeddings.position_embedding.weight"]
    x = F.conv2d(px, w, b, stride=patch).flatten(2).transpose(1, 2) + pos_w[None]
    x = _encoder(x, sd, prefix, layers, heads, eps, "gelu_pytorch_tanh", None, stages, "vision_")
    x = _ln(x, sd, prefix + "post_layernorm", eps)
    # SiglipMultiheadAttentionPoolingHead: packed nn.MultiheadAttention, one learned query
    h = prefix + "head."
    D = x.shape[-1]
    wi, bi = sd[h + "attention.in_proj_weight"], sd[h + "attention.in_proj_bias"]
    probe = sd[h + "probe"].expand(B, 1, D)
    q = F.linear(probe, wi[:D], bi[:D])
    k = F.linear(x, wi[D:2 * D], bi[D:2 * D])
    v = F.linear(x, wi[2 * D:], bi[2 * D:])
    a = _attention(q, k, v, heads, None)
    y = F.linear(a, sd[h + "attention.out_proj.weight"], sd[h + "attention.out_proj.bias"])
    r = y
    y = _ln(y, sd, h + "layernorm", eps)
    y = _lin(_act(_lin(y, sd, h + "mlp.fc1"), "gelu_pytorch_tanh"), sd, h + "mlp.fc2")
    return (r + y)[:, 0]


# ------------------------------------------------------------------------------------------ heads
def fusion_forward(sd: SD, batch: Dict[str, Tensor], backend: str = "clip", patch: int = 32,
                   eos_id: int = 49407, stages: Optional[dict] = None) -> Tensor:
    """MultiModalFusionClassifier.forward -> logits  (R/src/models/fusion.py:157-229)."""
    ids, mask, px = batch["input_ids"], batch.get("attention_mask"), batch["pixel_values"].float()
    tp, ip = batch["text_present"].float(), batch["image_present"].float()
    if backend == "clip":
        t = clip_text_pooled(sd, "backbone.text_model.", ids, mask, eos_id, stages=stages)
        v = clip_vision_pooled(sd, "backbone.vision_model.", px, patch, stages=stages)
        if stages is not None:
            stages["text_pooled"], stages["vision_pooled"] = t, v
        t = F.linear(t, sd["backbone.text_projection.weight"])      # HF clip :822-823
        v = F.linear(v, sd["backbone.visual_projection.weight"])    # HF clip :860-861
    else:
        t = siglip_text_pooled(sd, "backbone.text_model.", ids, mask, stages=stages)
        v = siglip_vision_pooled(sd, "backbone.vision_model.", px, patch, stages=stages)
    if stages is not None:
        stages["text_feat"], stages["vision_feat"] = t, v
    t = F.normalize(t, dim=-1) * tp[:, None]                        # fusion.py:188-189
    v = F.normalize(v, dim=-1) * ip[:, None]
    tpj = _lin(t, sd, "proj_t")                                     # :192-193
    vpj = _lin(v, sd, "proj_i")
    zt = torch.tanh(_lin(tpj, sd, "g_t"))                           # :196-199
    zi = torch.tanh(_lin(vpj, sd, "g_i"))
    g = torch.sigmoid(_lin(torch.cat([tpj, vpj, tp[:, None], ip[:, None]], 1), sd, "gate"))
    fused = torch.where((ip < 0.5)[:, None], zt,
                        torch.where((tp < 0.5)[:, None], zi, g * zt + (1.0 - g) * zi))   # :202-205
    fused = _ln(fused, sd, "ln_fused", 1e-5)                        # :206
    feat = torch.cat([fused, tpj, vpj, (tpj - vpj).abs(), tpj * vpj], 1)                # :209-215
    h = _ln(feat, sd, "cls.0", 1e-5)                                # :140-146
    h = F.gelu(_lin(h, sd, "cls.1"))
    return _lin(h, sd, "cls.4")


def mtl_forward(sd: SD, batch: Dict[str, Tensor], patch: int = 32, eos_id: int = 49407,
                stages: Optional[dict] = None) -> Tensor:
    """MultiTaskClassifier.forward (clip backend) -> logits [B,T]  (R/src/models/multitask.py:156-227)."""
    ids, mask, px = batch["input_ids"], batch.get("attention_mask"), batch["pixel_values"].float()
    tp, ip = batch["text_present"].float(), batch["image_present"].float()
    t = clip_text_pooled(sd, "tower_txt.text_model.", ids, mask, eos_id, stages=stages)    # :130-136
    v = clip_vision_pooled(sd, "tower_img.vision_model.", px, patch, stages=stages)        # :143-149
    if stages is not None:
        stages["text_pooled"], stages["vision_pooled"] = t, v
    tf = _lin(t, sd, "proj_t")                                       # :184-185
    vf = _lin(v, sd, "proj_i")
    zt = torch.tanh(_lin(tf, sd, "g_t"))                             # :188-191
    zi = torch.tanh(_lin(vf, sd, "g_i"))
    g = torch.sigmoid(_lin(torch.cat([tf, vf, tp[:, None], ip[:, None]], 1), sd, "gate"))
    fused = torch.where((ip < 0.5)[:, None], zt,
                        torch.where((tp < 0.5)[:, None], zi, g * zt + (1.0 - g) * zi))    # :194-197
    shared = F.gelu(_lin(fused, sd, "shared_head.1"))                # :98-103,200
    outs = []
    j = 0
    while f"heads.{j}.weight" in sd or f"heads.{j}.0.weight" in sd:  # :203-207
        if f"heads.{j}.weight" in sd:
            o = _lin(shared, sd, f"heads.{j}")
        else:
            o = _lin(F.gelu(_lin(shared, sd, f"heads.{j}.0")), sd, f"heads.{j}.3")
        outs.append(o.squeeze(-1))
        j += 1
    return torch.stack(outs, dim=1)


def bce_loss(logits: Tensor, labels: Tensor, pos_weight: Optional[Tensor] = None) -> Tensor:
    """R/src/models/fusion.py:223-226."""
    return F.binary_cross_entropy_with_logits(logits, labels, pos_weight=pos_weight)
