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


def _parse_cpulist(spec: str) -> set:
    cpus = set()
    for part in spec.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _numa_from_sysfs(device_index: int):
    p = torch.cuda.get_device_properties(device_index)
    bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
        node = int(f.read().strip())
    if node < 0:
        return None
    with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
        return node, _parse_cpulist(f.read())


def _numa_from_nvidia_smi(device_index: int):
    """`nvidia-smi topo -m` prints, per GPU row, "CPU Affinity" and "NUMA Affinity" columns (the driver's own view;
    it works inside containers whose sysfs hides the PCI numa_node)."""
    import re
    import subprocess
    out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    out = re.sub(r"\x1b\[[0-9;]*m", "", out)
    header = None
    for line in out.splitlines():
        cols = [c for c in line.split("\t") if c.strip()]
        if header is None and "CPU Affinity" in line:
            header = [c.strip() for c in cols]
            continue
        if header and cols and cols[0].strip() == f"GPU{device_index}":
            vals = [c.strip() for c in cols]
            # the header has no entry above the row label: data column i sits under header[i - 1]
            row = dict(zip(header, vals[1:]))
            cpus = _parse_cpulist(row.get("CPU Affinity", ""))
            node_s = row.get("NUMA Affinity", "").split(",")[0].split("-")[0]
            node = int(node_s) if node_s.isdigit() else None
            if cpus:
                return node if node is not None else -1, cpus
    return None


def _prefer_numa_memory(node: int) -> bool:
    """set_mempolicy(MPOL_PREFERRED, {node}): pages this thread faults in afterwards -- the pinned staging buffers --
    come from the GPU's NUMA node even if the scheduler moves the thread.  Best effort (x86-64 / aarch64 syscall)."""
    import ctypes
    import platform
    nr = {"x86_64": 238, "aarch64": 237}.get(platform.machine())
    if nr is None or node < 0 or node >= 1024:
        return False
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(nr, 1, ctypes.byref(mask), ctypes.c_ulong(1025)) == 0
    except Exception:
        return False


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated afterwards
    (first touch) and the H2D copies of `mmcm_forward_host` stay on the GPU's side of the machine.  One process per
    GPU.  Topology comes from sysfs, else from `nvidia-smi topo -m`; returns the node (-1: CPU affinity known, node
    number not) or None when neither source answers."""
    import os
    found = None
    for probe in (_numa_from_sysfs, _numa_from_nvidia_smi):
        try:
            found = probe(device_index)
        except Exception:
            found = None
        if found:
            break
    if not found:
        return None
    node, cpus = found
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
    except Exception:
        return None
    if node >= 0:
        _prefer_numa_memory(node)
    return node


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
