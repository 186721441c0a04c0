"""Host-side mirror of the callers' pre-/post-processing around the forward, on the GPU.

    resize_crop_u8  Resize(size, antialias=True) + CenterCrop(size) of the same transform (dataset.py:106-108) on decoded
                    uint8 RGB images of different sizes -- Pillow's fixed-point bilinear resampler, byte-identical
    preprocess_u8   ToTensor + Normalize of `SocialHarmDataset.eval_tf` (R/src/data/dataset.py:106-111)
    postprocess     sigmoid -> per-class thresholds -> any_harmful (R/scripts/inference.py:218-232) and the per-class
                    confusion counts behind compute_detailed_metrics (R/src/training/metrics.py:164-215)
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import lib as L


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def resize_crop_u8(images: Sequence[torch.Tensor], size: int, device=None) -> torch.Tensor:
    """Decoded RGB images (uint8 [H_i, W_i, 3] tensors, CPU or CUDA, any sizes) -> CUDA uint8 [B, size, size, 3]:
    shorter side resized to `size` like `Image.resize(BILINEAR)`, then the centre crop.  Feed the result to
    `model.forward_u8` / `preprocess_u8`."""
    if len(images) == 0:
        raise ValueError("resize_crop_u8 needs at least one image")
    for im in images:
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[-1] != 3:
            raise ValueError("resize_crop_u8 expects uint8 [H, W, 3] tensors")
    if device is None:
        device = next((im.device for im in images if im.is_cuda), torch.device("cuda", torch.cuda.current_device()))
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("resize_crop_u8 runs on the GPU: no CPU fallback")
    B = len(images)
    flat = torch.cat([im.contiguous().reshape(-1) for im in images]).to(device, non_blocking=True)
    sizes = [int(im.shape[0]) * int(im.shape[1]) * 3 for im in images]
    offs, acc = [], 0
    for n in sizes:
        offs.append(acc)
        acc += n
    offsets = (C.c_int64 * B)(*offs)
    heights = (C.c_int32 * B)(*[int(im.shape[0]) for im in images])
    widths = (C.c_int32 * B)(*[int(im.shape[1]) for im in images])
    out = torch.empty((B, size, size, 3), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        L.check(L.load().mmcm_resize_crop_u8(flat.data_ptr(), offsets, heights, widths, B, int(size), out.data_ptr(),
                                             _stream(flat)))
    flat.record_stream(torch.cuda.current_stream(device))
    return out


def preprocess_u8(images_hwc: torch.Tensor, mean: Sequence[float], std: Sequence[float]) -> torch.Tensor:
    """uint8 CUDA tensor [B,H,W,3] (resized + centre-cropped) -> fp32 pixel_values [B,3,H,W]."""
    if not images_hwc.is_cuda or images_hwc.dtype != torch.uint8 or images_hwc.dim() != 4 or images_hwc.shape[-1] != 3:
        raise ValueError("preprocess_u8 expects a CUDA uint8 tensor [B,H,W,3]")
    x = images_hwc.contiguous()
    B, H, W, _ = x.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    m = (C.c_float * 3)(*[float(v) for v in mean])
    s = (C.c_float * 3)(*[float(v) for v in std])
    with torch.cuda.device(x.device):
        L.check(L.load().mmcm_preprocess_u8(x.data_ptr(), B, H, W, m, s, out.data_ptr(), _stream(x)))
    return out


def postprocess(logits: torch.Tensor, thresholds: torch.Tensor, labels: Optional[torch.Tensor] = None,
                confusion: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Returns probs [B,C] fp32, labels [B,C] bool, any_harmful [B] bool and (with `labels`) the running confusion
    counts [C,4] int64 = TP, FP, FN, TN (pass the tensor back in to accumulate over batches)."""
    if not logits.is_cuda:
        raise RuntimeError("postprocess runs on the GPU: no CPU fallback")
    lg = logits.to(torch.float32).contiguous()
    B, Cn = lg.shape
    thr = thresholds.to(device=lg.device, dtype=torch.float32).contiguous()
    if thr.numel() != Cn:
        raise ValueError("one threshold per class is required")
    probs = torch.empty_like(lg)
    dec = torch.empty((B, Cn), dtype=torch.uint8, device=lg.device)
    anyh = torch.empty((B,), dtype=torch.uint8, device=lg.device)
    lab = None
    if labels is not None:
        lab = labels.to(device=lg.device, dtype=torch.float32).contiguous()
        if confusion is None:
            confusion = torch.zeros((Cn, 4), dtype=torch.int64, device=lg.device)
    with torch.cuda.device(lg.device):
        L.check(L.load().mmcm_postprocess(lg.data_ptr(), thr.data_ptr(), None if lab is None else lab.data_ptr(), B, Cn,
                                          probs.data_ptr(), dec.data_ptr(), anyh.data_ptr(),
                                          None if lab is None else confusion.data_ptr(), _stream(lg)))
    out = {"probs": probs, "labels": dec.bool(), "any_harmful": anyh.bool()}
    if lab is not None:
        out["confusion"] = confusion
    return out


def metrics_from_confusion(confusion: torch.Tensor) -> Dict[str, object]:
    """f1 / precision / recall (macro, micro, per class; zero_division=0) from TP, FP, FN, TN counts -- the numbers
    sklearn's f1_score / precision_score / recall_score return in R/src/training/metrics.py:180-205."""
    c = confusion.detach().to("cpu", torch.float64)
    tp, fp, fn = c[:, 0], c[:, 1], c[:, 2]

    def _div(a, b):
        return torch.where(b > 0, a / b.clamp(min=1), torch.zeros_like(a))
    prec, rec = _div(tp, tp + fp), _div(tp, tp + fn)
    f1 = _div(2 * tp, 2 * tp + fp + fn)
    TP, FP, FN = tp.sum(), fp.sum(), fn.sum()
    return {"f1_macro": f1.mean().item(), "f1_micro": _div(2 * TP, 2 * TP + FP + FN).item(),
            "precision_macro": prec.mean().item(), "recall_macro": rec.mean().item(),
            "per_class": {"f1": f1.tolist(), "precision": prec.tolist(), "recall": rec.tolist(),
                          "support": (tp + fn).to(torch.int64).tolist()}}
