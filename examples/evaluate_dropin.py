#!/usr/bin/env python
"""The reference's evaluation loop (R/scripts/evaluate.py:163-183, :225-239) on the B200 path, with synthetic data.

    python examples/evaluate_dropin.py [--samples 4096] [--batch 512] [--head fusion|mtl]

Identical call pattern to the reference: build the model class, load a state dict, `.to(device).eval()`, loop
`outputs = model(**batch); logits = outputs["logits"]`.  The NumPy / sklearn post-processing of evaluate.py is replaced
by `prepost.postprocess` (probabilities, per-class thresholds, confusion counts stay on the device; one D2H at the end).
Under `torchrun --nproc-per-node N` every rank scores a contiguous shard and the scores are gathered once.
"""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package  # noqa: E402

P = load_package()
from mmcm_b200 import arch as A, prepost, sharding, synthetic as syn  # noqa: E402

CLASSES = ["racist", "sexist", "homophobe", "religion", "otherhate"]          # R/config/default.yaml:27-32
THRESHOLDS = [0.45, 0.2, 0.35, 0.3, 0.25]                                      # the range of R/runs/*/inference_config.json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4096)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--head", default="fusion", choices=["fusion", "mtl"])
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    a = A.CLIP_B32
    if args.head == "fusion":
        model = P.MultiModalFusionClassifier("openai/clip-vit-base-patch32", num_labels=len(CLASSES))
        sd = syn.make_state_dict(A.fusion_spec(a, len(CLASSES), 512), a, seed=0, hardened=True)
    else:
        model = P.MultiTaskClassifier("openai/clip-vit-base-patch32", CLASSES, head_hidden_dim=256)
        sd = syn.make_state_dict(A.mtl_spec(a, len(CLASSES), 512, 256), a, seed=0, hardened=True)
    model.load_state_dict(sd)                                   # evaluate.py:139-151
    model = model.to(dev).eval()                                # :153-154

    data = syn.make_inputs(a, args.samples, seed=7)             # stands in for SocialHarmDataset + collate_fn
    labels = (torch.rand(args.samples, len(CLASSES), generator=torch.Generator().manual_seed(1)) < 0.25).float()
    lo, hi = sharding.shard_range(args.samples, rank, world)
    thr = torch.tensor(THRESHOLDS, device=dev)
    with torch.no_grad():                                       # first call uploads / repacks the weights
        model(**{k: v[:8].to(dev) for k, v in data.items()})
    confusion = None
    all_logits = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for s in range(lo, hi, args.batch):                     # evaluate.py:170-178
            e = min(s + args.batch, hi)
            batch = {k: v[s:e].to(dev, non_blocking=True) for k, v in data.items()}
            outputs = model(**batch, labels=labels[s:e].to(dev))
            logits = outputs["logits"]
            post = prepost.postprocess(logits, thr, labels=labels[s:e].to(dev), confusion=confusion)
            confusion = post["confusion"]
            all_logits.append(logits)
    local_logits = torch.cat(all_logits) if all_logits else torch.zeros(0, len(CLASSES), device=dev)
    scores = sharding.gather_scores(local_logits, args.samples)
    if world > 1:
        dist.all_reduce(confusion)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        m = prepost.metrics_from_confusion(confusion)
        print(f"scored {args.samples} samples on {world} GPU(s) in {dt * 1e3:.1f} ms ({args.samples / dt:.0f} samples/s incl. H2D)")
        print(f"f1_macro {m['f1_macro']:.4f}  f1_micro {m['f1_micro']:.4f}  precision_macro {m['precision_macro']:.4f}  "
              f"recall_macro {m['recall_macro']:.4f}")
        print("per-class f1", [round(x, 4) for x in m["per_class"]["f1"]], "scores", tuple(scores.shape))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
