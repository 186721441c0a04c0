"""World-size-2 gloo test of the multi-GPU host logic (contiguous shards + one final gather), on CPU.
The scoring function is a deterministic stand-in: the CUDA model itself has no CPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _standin(b):
    # depends on every input field and on the row only -> order errors are visible
    return torch.stack([b["input_ids"].float().sum(1), b["pixel_values"].flatten(1).sum(1),
                        b["text_present"] * 3 + b["image_present"]], dim=1)


def _worker(rank, world, port, n, mb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from __graft_entry__ import load_package
    load_package()
    from mmcm_b200 import sharding
    g = torch.Generator().manual_seed(0)
    batch = {"input_ids": torch.randint(0, 100, (n, 7), generator=g),
             "pixel_values": torch.randn(n, 3, 4, 4, generator=g),
             "text_present": torch.ones(n), "image_present": (torch.arange(n) % 2).float()}
    full = sharding.score_sharded(_standin, batch, micro_batch=mb)
    ok = torch.allclose(full, _standin(batch)) and full.shape == (n, 3)
    q.put((rank, bool(ok), sharding.shard_range(n, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,mb", [(10, 4), (11, 3), (1, 8)])
def test_sharded_scoring_two_ranks(n, mb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, mb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    (lo0, hi0), (lo1, hi1) = res[0][2], res[1][2]
    assert lo0 == 0 and hi0 == lo1 and hi1 == n and (hi0 - lo0) - (hi1 - lo1) in (0, 1)


def test_shard_range_partitions():
    from mmcm_b200 import sharding
    for n in (0, 1, 7, 22500):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
