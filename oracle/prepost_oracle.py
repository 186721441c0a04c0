"""CPU ORACLE (test infrastructure only) for the callers' pre-/post-processing around the forward.

Restates, with the reference's own libraries (numpy / torchvision-free torch ops / sklearn):
  * eval transform tail   R/src/data/dataset.py:106-111     ToTensor (HWC uint8 -> CHW float / 255) then Normalize
  * post-processing       R/scripts/inference.py:218-232    probs = 1/(1+np.exp(-logits)); prob >= thresh; any(...)
  * detailed metrics      R/src/training/metrics.py:180-205 sklearn f1/precision/recall on (probs >= threshold)
Parity pinning: these are the reference's literal expressions; sklearn is the library the reference itself calls.
"""
from __future__ import annotations

import numpy as np
import torch


def to_tensor_normalize(images_hwc_u8: torch.Tensor, mean, std) -> torch.Tensor:
    x = images_hwc_u8.permute(0, 3, 1, 2).to(torch.float32).div(255)            # torchvision F.to_tensor
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return x.sub(m).div(s)                                                        # torchvision F.normalize


def postprocess(logits: np.ndarray, thresholds: np.ndarray):
    probs = 1 / (1 + np.exp(-logits))                                             # inference.py:218
    labels = probs >= thresholds[None, :]                                         # :222-225
    return probs, labels, labels.any(axis=1)                                      # :229-230


def detailed_metrics(probs: np.ndarray, y_true: np.ndarray, thresholds: np.ndarray):
    from sklearn.metrics import f1_score, precision_score, recall_score
    bin_preds = (probs >= thresholds[None, :]).astype(int)
    return {"f1_macro": float(f1_score(y_true, bin_preds, average="macro", zero_division=0)),
            "f1_micro": float(f1_score(y_true, bin_preds, average="micro", zero_division=0)),
            "precision_macro": float(precision_score(y_true, bin_preds, average="macro", zero_division=0)),
            "recall_macro": float(recall_score(y_true, bin_preds, average="macro", zero_division=0)),
            "per_class_f1": [float(f1_score(y_true[:, j], bin_preds[:, j], zero_division=0))
                             for j in range(probs.shape[1])]}


# ---------------------------------------------------------------------------------------------------------------
# Resize(size, antialias=True) + CenterCrop(size) of `eval_tf` (R/src/data/dataset.py:106-108) on a PIL RGB image.
# torchvision hands a PIL image to `Image.resize(..., BILINEAR)`, so the arithmetic is Pillow's two-pass fixed-point
# resampler (Pillow 12.2.0, src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
# ImagingResampleHorizontal_8bpc / Vertical_8bpc), restated here in numpy and pinned against Pillow itself in
# tests/test_prepost_cpu.py.
# ---------------------------------------------------------------------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def resized_hw(h: int, w: int, size: int):
    """torchvision.transforms.functional._compute_resized_output_size for an int `size` (shorter side -> size)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)            # (new_h, new_w)


def center_crop_origin(h: int, w: int, size: int):
    """torchvision F.center_crop: int(round((h - size) / 2.0)) with Python's round-half-to-even."""
    return int(round((h - size) / 2.0)), int(round((w - size) / 2.0))          # (top, left)


def bilinear_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1.0), box = the whole axis.
    Returns bounds [out_size, 2] (first source index, tap count) and integer coefficients [out_size, ksize]."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ws = []
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * ss)
            wgt = 1.0 - v if v < 1.0 else 0.0
            ws.append(wgt)
            ww += wgt
        for x in range(xmax):
            k = ws[x] / ww if ww != 0.0 else ws[x]
            kk[xx, x] = int(0.5 + k * (1 << PRECISION_BITS))                   # coefficients are >= 0 for bilinear
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_center_crop(img_hwc_u8: np.ndarray, size: int) -> np.ndarray:
    """uint8 [H,W,3] -> uint8 [size,size,3], bit-identical to T.CenterCrop(size)(T.Resize(size)(PIL image))."""
    h, w, _ = img_hwc_u8.shape
    new_h, new_w = resized_hw(h, w, size)
    src = img_hwc_u8.astype(np.int64)
    bh, kh = bilinear_coeffs(w, new_w)
    bv, kv = bilinear_coeffs(h, new_h)
    tmp = np.zeros((h, new_w, 3), dtype=np.uint8)                               # horizontal pass first, rounded to uint8
    for xx in range(new_w):
        x0, n = bh[xx]
        acc = (src[:, x0:x0 + n, :] * kh[xx, :n][None, :, None]).sum(axis=1) + (1 << (PRECISION_BITS - 1))
        tmp[:, xx, :] = _clip8(acc)
    tmp64 = tmp.astype(np.int64)
    out = np.zeros((new_h, new_w, 3), dtype=np.uint8)
    for yy in range(new_h):
        y0, n = bv[yy]
        acc = (tmp64[y0:y0 + n] * kv[yy, :n][:, None, None]).sum(axis=0) + (1 << (PRECISION_BITS - 1))
        out[yy] = _clip8(acc)
    top, left = center_crop_origin(new_h, new_w, size)
    return out[top:top + size, left:left + size]
