"""Batch-sharded scoring across the GPUs of one box.

Samples are independent (no cross-sample op anywhere in fusion.py / multitask.py / the encoders in eval mode), so
the N samples of a scoring job are split contiguously over the ranks, weights are replicated, and the only
collective is ONE gather of the `[N_r, C]` fp32 scores at the end (20 B/sample).  One process per GPU,
`torch.distributed` (NCCL on GPUs; gloo in the CPU tests) is plumbing only.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated afterwards
    (first touch) and the H2D copies of `mmcm_forward_host` stay on the GPU's side of the machine.  One process per
    GPU; a no-op (returns None) when the topology files are not readable."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank`: the first n % world ranks get one extra sample."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scores(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks receive the [n_total, C] scores in sample order.  Ragged shards are padded to the largest."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    C = local.shape[1]
    pad = torch.zeros((width, C), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, C), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


@torch.no_grad()
def score_sharded(score_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], batch: Dict[str, torch.Tensor],
                  micro_batch: int = 1024, group=None, device: Optional[torch.device] = None) -> torch.Tensor:
    """Score a whole job: every rank holds (or can index) the full `batch` dict, takes its contiguous shard, runs
    `score_fn` (e.g. ``lambda b: model(**b)["logits"]``) in chunks of `micro_batch`, then gathers the scores."""
    n = next(iter(batch.values())).shape[0]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    outs = []
    for s in range(lo, hi, micro_batch):
        e = min(s + micro_batch, hi)
        chunk = {k: (v[s:e].to(device, non_blocking=True) if device is not None else v[s:e]) for k, v in batch.items()}
        outs.append(score_fn(chunk).float())
    if outs:
        local = torch.cat(outs, dim=0)
    else:
        probe = score_fn({k: (v[:1].to(device) if device is not None else v[:1]) for k, v in batch.items()})
        local = probe.float()[:0]
    return gather_scores(local, n, group)
