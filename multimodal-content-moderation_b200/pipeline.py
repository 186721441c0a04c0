"""Batched scoring in the callers' place.

The reference scores one sample per forward: `MultiModalClassifier.predict_batch` is a Python loop over `predict`
(R/scripts/inference.py:256-270) and SageMaker's `predict_fn` loops over instances (R/sagemaker/inference.py:241-296).
`BatchedScorer` takes the already tokenised ids and the uint8 images of MANY requests -- either already resized /
cropped `[N,H,W,3]`, or a list of decoded images of any sizes, which `prepost.resize_crop_u8` resizes and crops on the
GPU byte-identically to Pillow -- and runs them as one batch: one forward on the raw uint8 images (ToTensor/Normalize inside the patch im2col,
`mmcm_forward_u8`), fused sigmoid / thresholds / any_harmful (prepost.postprocess).  Tokenisation and JPEG decode stay on the CPU exactly as in the
reference (R/src/data/dataset.py:106-165).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import prepost


class BatchedScorer:
    def __init__(self, model: torch.nn.Module, class_names: Sequence[str], thresholds: Sequence[float],
                 image_mean: Optional[Sequence[float]] = None, image_std: Optional[Sequence[float]] = None,
                 max_batch: int = 1024, tokenizer=None, max_text_length: Optional[int] = None):
        """image_mean / image_std default to the image processor constants of the model's encoder family (CLIP's
        dataset statistics, 0.5 / 0.5 for SigLIP) -- what `img_processor.image_mean` gives the reference's transform
        (R/src/data/dataset.py:100-110)."""
        from . import arch as A
        self.model = model.eval()
        fam_mean, fam_std = A.IMAGE_NORM[model._arch.backend]
        image_mean = fam_mean if image_mean is None else image_mean
        image_std = fam_std if image_std is None else image_std
        self.class_names = list(class_names)
        self.device = torch.device("cuda", model._device_index()) if hasattr(model, "_device_index") \
            else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("BatchedScorer needs the model on a CUDA device: no CPU fallback")
        self.thresholds = torch.tensor(list(thresholds), dtype=torch.float32, device=self.device)
        if self.thresholds.numel() != len(self.class_names):
            raise ValueError("one threshold per class is required")      # inference.py:118-121 pads/validates likewise
        self.mean, self.std, self.max_batch = list(image_mean), list(image_std), int(max_batch)
        # optional: anything callable like the reference's tokenizer (tokenizer.ClipTokenizer, or the HF object itself)
        self.tokenizer = tokenizer
        self.max_text_length = int(max_text_length or model._arch.max_pos)

    def score_texts(self, texts: Sequence[str], images_u8=None, image_present: Optional[torch.Tensor] = None
                    ) -> Dict[str, torch.Tensor]:
        """Raw request texts in: tokenised the way the reference does (pad to max_length, truncation,
        R/scripts/inference.py:168-180), `text_present` = non-blank text (inference.py:201), then `score`."""
        if self.tokenizer is None:
            raise RuntimeError("BatchedScorer was built without a tokenizer")
        enc = self.tokenizer(list(texts), padding="max_length", truncation=True, max_length=self.max_text_length,
                             return_attention_mask=True, return_tensors="pt")
        tp = torch.tensor([1.0 if (t and t.strip()) else 0.0 for t in texts], dtype=torch.float32)
        return self.score(enc["input_ids"], enc["attention_mask"], images_u8, text_present=tp, image_present=image_present)

    @torch.no_grad()
    def score(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, images_u8: Optional[torch.Tensor],
              text_present: Optional[torch.Tensor] = None, image_present: Optional[torch.Tensor] = None
              ) -> Dict[str, torch.Tensor]:
        """input_ids / attention_mask [N,S] int64; images_u8 [N,H,W,3] uint8 crops, or a list of N decoded uint8
        [H_i,W_i,3] images of any sizes (resized + centre-cropped on the GPU), or None: no images at all.
        Returns probs [N,C], labels [N,C] bool, any_harmful [N] bool -- the fields of inference.py:220-232."""
        N = input_ids.shape[0]
        dev = self.device
        a = self.model._arch
        tp = torch.ones(N, device=dev) if text_present is None else text_present.to(dev, torch.float32)
        if images_u8 is None:                                            # inference.py:142-166: zero image, flag 0
            ip = torch.zeros(N, device=dev)
        else:
            ip = torch.ones(N, device=dev) if image_present is None else image_present.to(dev, torch.float32)
        outs: List[Dict[str, torch.Tensor]] = []
        for s in range(0, N, self.max_batch):
            e = min(s + self.max_batch, N)
            ids = input_ids[s:e].to(dev, non_blocking=True)
            mask = attention_mask[s:e].to(dev, non_blocking=True)
            if images_u8 is None:
                px = torch.zeros(e - s, 3, a.image, a.image, device=dev)
                logits = self.model(input_ids=ids, attention_mask=mask, pixel_values=px, text_present=tp[s:e],
                                    image_present=ip[s:e])["logits"]
            else:
                if isinstance(images_u8, (list, tuple)):                 # raw decoded images: Resize + CenterCrop here
                    crops = prepost.resize_crop_u8(images_u8[s:e], a.image, device=dev)
                else:
                    crops = images_u8[s:e].to(dev, non_blocking=True)
                logits = self.model.forward_u8(ids, mask, crops, tp[s:e], ip[s:e], self.mean, self.std)
            outs.append(prepost.postprocess(logits, self.thresholds))
        return {k: torch.cat([o[k] for o in outs], dim=0) for k in ("probs", "labels", "any_harmful")}

    def as_records(self, result: Dict[str, torch.Tensor]) -> List[dict]:
        """The per-request dictionaries `predict` returns (inference.py:220-234)."""
        probs, labels, anyh = result["probs"].cpu(), result["labels"].cpu(), result["any_harmful"].cpu()
        thr = self.thresholds.cpu()
        recs = []
        for i in range(probs.shape[0]):
            preds = {n: {"label": bool(labels[i, j]), "probability": float(probs[i, j]), "threshold": float(thr[j])}
                     for j, n in enumerate(self.class_names)}
            recs.append({"predictions": preds, "any_harmful": bool(anyh[i])})
        return recs
