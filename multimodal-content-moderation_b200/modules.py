"""Drop-in replacements for the reference's two model classes.

    MultiModalFusionClassifier   R/src/models/fusion.py:55-229
    MultiTaskClassifier          R/src/models/multitask.py:16-227

Same constructor kwargs, same state-dict keys (a checkpoint written by the reference loads with
`load_state_dict(..., strict=True)`), same `forward(input_ids, attention_mask, pixel_values, text_present,
image_present, labels=None) -> {"loss", "logits"}` -- so `scripts/evaluate.py`, `scripts/inference.py` and
`sagemaker/inference.py` can import these instead of `src.models`.

The modules own fp32 master parameters exactly like the reference; the forward itself is the C-ABI call
`mmcm_forward` into hand-written sm_100a kernels.  Inference only: there is no autograd through the extension,
and no CPU fallback -- calling forward on CPU tensors raises.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import arch as A
from . import synthetic as syn
from .engine import Engine


class _Params(nn.Module):
    """A bare container; the nesting of these reproduces the reference's dotted state-dict keys."""


def _register_flat(root: nn.Module, spec: A.Spec, sd: Dict[str, torch.Tensor]) -> None:
    for key, _shape in spec:
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            nxt = mod._modules.get(p)
            if nxt is None:
                nxt = _Params()
                mod.add_module(p, nxt)
            mod = nxt
        mod.register_parameter(parts[-1], nn.Parameter(sd[key], requires_grad=False))


class _B200ScoringModule(nn.Module):
    """Shared plumbing: parameter tree, engine ownership, weight refresh, device checks."""

    _head: int = A.HEAD_FUSION

    def _init_common(self, a: A.ArchCfg, spec: A.Spec, num_outputs: int, fusion_dim: int, head_hidden_dim: int):
        self._arch = a
        self._spec = spec
        self._num_outputs = num_outputs
        self._fusion_dim = fusion_dim
        self._head_hidden_dim = head_hidden_dim
        # the reference initialises from a hub checkpoint (from_pretrained) and callers then load_state_dict;
        # offline we start from the same distributions, seeded from torch's global RNG state
        seed = int(torch.initial_seed() % (2 ** 31))
        _register_flat(self, spec, syn.make_state_dict(spec, a, seed=seed, hardened=False))
        self._engine: Optional[Engine] = None
        self._engine_dirty = True
        self._offloaded_to: Optional[int] = None   # CUDA device the engine lives on while the masters sit on the host

    # -- nn.Module hooks that can change parameter storage: mark the engine's repacked copy stale
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._engine_dirty = True
        self._offloaded_to = None
        return out

    def offload_master(self) -> "_B200ScoringModule":
        """Move the fp32 master parameters back to host memory and keep scoring on the GPU.

        The extension owns its own repacked copy (bf16 GEMM operands, fp32 norms / heads: 0.37 GB for CLIP-Fusion), so
        after the push the 0.62 GB of fp32 `nn.Parameter`s on the device are dead weight for inference.  `state_dict()`
        keeps working (host tensors); `.to(device)` / `load_state_dict` bring the usual behaviour back."""
        dev = self._device_index()
        self._ensure_engine(dev)
        super()._apply(lambda t: t.cpu())
        self._offloaded_to = dev
        self._engine_dirty = False
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self._engine_dirty = True
        return out

    def refresh_weights(self) -> None:
        """Re-push the master parameters into the extension (call after modifying parameters in place)."""
        self._engine_dirty = True

    def set_option(self, name: str, value: int) -> None:
        self._ensure_engine(self._device_index())
        self._engine.set_option(name, value)

    def save_packed(self, path: str) -> None:
        """Write the extension's repacked weight set as one blob (`mmcm_save_packed`; see checkpoint.PackedScorer)."""
        self._ensure_engine(self._device_index()).save_packed(path)

    def _device_index(self) -> int:
        if self._offloaded_to is not None:
            return self._offloaded_to
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("the B200 scoring path has no CPU fallback: move the model to a CUDA device first")
        return p.device.index if p.device.index is not None else torch.cuda.current_device()

    def _ensure_engine(self, dev: int) -> Engine:
        if self._engine is not None and self._engine.device != dev:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = Engine(self._arch, self._head, self._num_outputs, self._fusion_dim,
                                  self._head_hidden_dim, dev)
            self._engine_dirty = True
        if self._engine_dirty:
            self._engine.load_state_dict({k: v for k, v in self.state_dict().items()})
            self._engine_dirty = False
        return self._engine

    def _logits(self, input_ids, attention_mask, pixel_values, text_present, image_present) -> torch.Tensor:
        if not input_ids.is_cuda:
            raise RuntimeError("inputs must be CUDA tensors: the B200 scoring path has no CPU fallback")
        dev = input_ids.device.index if input_ids.device.index is not None else torch.cuda.current_device()
        if self._device_index() != dev:
            raise RuntimeError("model and inputs are on different devices")
        eng = self._ensure_engine(dev)
        with torch.cuda.device(dev):
            return eng.forward(input_ids, attention_mask, pixel_values, text_present, image_present)

    @torch.no_grad()
    def forward_u8(self, input_ids, attention_mask, images_u8, text_present, image_present, image_mean, image_std,
                   want_probs: bool = False):
        """Forward on raw uint8 [B,H,W,3] images (already resized / cropped): the eval transform's ToTensor +
        Normalize (R/src/data/dataset.py:106-111) run inside the patch im2col.  Returns logits (and probs)."""
        if not input_ids.is_cuda:
            raise RuntimeError("inputs must be CUDA tensors: the B200 scoring path has no CPU fallback")
        dev = input_ids.device.index if input_ids.device.index is not None else torch.cuda.current_device()
        if self._device_index() != dev:
            raise RuntimeError("model and inputs are on different devices")
        eng = self._ensure_engine(dev)
        with torch.cuda.device(dev):
            return eng.forward_u8(input_ids, attention_mask, images_u8, image_mean, image_std, text_present,
                                  image_present, want_probs=want_probs)

    @torch.no_grad()
    def predict_proba(self, **batch) -> torch.Tensor:
        """sigmoid(logits) fused into the head kernel (the callers' NumPy post-processing, inference.py:218)."""
        ids = batch["input_ids"]
        dev = ids.device.index if ids.device.index is not None else torch.cuda.current_device()
        eng = self._ensure_engine(dev)
        with torch.cuda.device(dev):
            _, probs = eng.forward(ids, batch.get("attention_mask"), batch["pixel_values"], batch["text_present"],
                                   batch["image_present"], want_probs=True)
        return probs


def focal_with_logits(logits: torch.Tensor, targets: torch.Tensor, alpha: Optional[torch.Tensor] = None,
                      gamma: float = 1.5, reduction: str = "mean") -> torch.Tensor:
    """Focal BCE on logits, the formula of R/src/models/fusion.py:39-52: ce * (1 - p_t)^gamma [* alpha_t] with
    p_t = p t + (1 - p)(1 - t) written as a lerp between the negative- and positive-class probabilities."""
    p = torch.sigmoid(logits)
    p_t = torch.lerp(1.0 - p, p, targets)
    per_elem = F.binary_cross_entropy_with_logits(logits, targets, reduction="none") * (1.0 - p_t).pow(gamma)
    if alpha is not None:
        per_elem = per_elem * torch.lerp(1.0 - alpha, alpha, targets)
    if reduction == "none":
        return per_elem
    return per_elem.mean() if reduction == "mean" else per_elem.sum()


class FocalWithLogitsLoss(nn.Module):
    """Module form of `focal_with_logits` (same constructor as the reference's class, fusion.py:16-37); only used to
    report an evaluation-time loss when `labels` are passed."""

    def __init__(self, alpha: Optional[torch.Tensor] = None, gamma: float = 1.5, reduction: str = "mean"):
        super().__init__()
        self.register_buffer("alpha", alpha)
        self.gamma, self.reduction = gamma, reduction

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return focal_with_logits(logits, targets, self.alpha, self.gamma, self.reduction)


class MultiModalFusionClassifier(_B200ScoringModule):
    """Gated late-fusion classifier over CLIP / SigLIP towers (R/src/models/fusion.py:55-229)."""

    _head = A.HEAD_FUSION

    def __init__(self, encoder_name: str, num_labels: int, fusion_dim: int = 512, backend: str = "clip",
                 freeze_text: bool = False, freeze_image: bool = False, loss_type: str = "bce",
                 focal_gamma: float = 1.5, pos_weight: Optional[torch.Tensor] = None,
                 alpha_focal: Optional[torch.Tensor] = None):
        super().__init__()
        self.backend = backend.lower()
        a = A.resolve_arch(encoder_name, self.backend)
        self._init_common(a, A.fusion_spec(a, num_labels, fusion_dim), num_labels, fusion_dim, 0)
        self.loss_type = loss_type
        self.register_buffer("pos_weight", pos_weight if pos_weight is not None else None)
        self.criterion = FocalWithLogitsLoss(alpha=alpha_focal, gamma=focal_gamma) if loss_type == "focal" else None

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                text_present: torch.Tensor, image_present: torch.Tensor,
                labels: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        logits = self._logits(input_ids, attention_mask, pixel_values, text_present, image_present)
        loss = None
        if labels is not None:  # R/src/models/fusion.py:218-227 (evaluate.py passes labels)
            if self.loss_type == "focal":
                loss = self.criterion(logits, labels)
            else:
                loss = F.binary_cross_entropy_with_logits(
                    logits, labels, pos_weight=self.pos_weight if self.pos_weight is not None else None)
        return {"loss": loss, "logits": logits}


class MultiTaskClassifier(_B200ScoringModule):
    """Shared towers + per-task heads (R/src/models/multitask.py:16-227)."""

    _head = A.HEAD_MTL

    def __init__(self, encoder_name: str, task_names: List[str], fusion_dim: int = 512, backend: str = "clip",
                 threshold: float = 0.5, freeze_text: bool = False, freeze_image: bool = False,
                 pos_weight: Optional[torch.Tensor] = None, head_hidden_dim: Optional[int] = None,
                 learnable_task_weights: bool = False):
        super().__init__()
        self.task_names = list(task_names)
        self.num_tasks = len(self.task_names)
        self.threshold = threshold
        self.backend = backend.lower()
        a = A.resolve_arch(encoder_name, self.backend)
        hh = int(head_hidden_dim) if head_hidden_dim and head_hidden_dim > 0 else 0
        # arch.mtl_spec raises the reference's AssertionError for non-clip backends (multitask.py:81-88)
        self._init_common(a, A.mtl_spec(a, self.num_tasks, fusion_dim, hh), self.num_tasks, fusion_dim, hh)
        if pos_weight is not None:
            assert pos_weight.dim() == 1 and pos_weight.shape[0] == self.num_tasks, "pos_weight must be [num_tasks]"
            self.register_buffer("pos_weight", pos_weight.float())
        else:
            self.pos_weight = None
        self.log_vars = nn.Parameter(torch.zeros(self.num_tasks)) if learnable_task_weights else None

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, pixel_values: torch.Tensor,
                text_present: torch.Tensor, image_present: torch.Tensor,
                labels: Optional[torch.Tensor] = None) -> Dict[str, Any]:
        logits = self._logits(input_ids, attention_mask, pixel_values, text_present, image_present)
        loss = None
        if labels is not None:  # R/src/models/multitask.py:209-225
            per_task = []
            for j in range(self.num_tasks):
                pw = self.pos_weight[j] if self.pos_weight is not None else None
                lj = F.binary_cross_entropy_with_logits(logits[:, j], labels[:, j], pos_weight=pw, reduction="mean")
                if self.log_vars is not None:
                    per_task.append(torch.exp(-self.log_vars[j]) * lj + 0.5 * self.log_vars[j])
                else:
                    per_task.append(lj)
            loss = torch.stack(per_task, dim=0).mean()
        return {"loss": loss, "logits": logits}
