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
