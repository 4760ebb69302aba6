"""world_size-2 gloo test of the multi-GPU host logic (partition + pose all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepi2p_b200 import sharding


def test_partition_covers_everything():
    for n in (0, 1, 7, 512, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [sharding.partition(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.partition(4, 2, 2)


def _worker(rank, world, port, n_total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.partition(n_total, world, rank)
    ids = torch.arange(a, b, dtype=torch.float64)
    P = torch.eye(4, dtype=torch.float64).repeat(b - a, 1, 1)
    P[:, 0, 3] = ids                      # tag every record with its global sample id
    cost = ids * 10
    Pg, cg = sharding.gather_poses(P, cost, n_total=n_total)
    assert Pg.shape == (n_total, 4, 4)
    assert torch.equal(Pg[:, 0, 3], torch.arange(n_total, dtype=torch.float64))
    assert torch.equal(cg, torch.arange(n_total, dtype=torch.float64) * 10)
    assert torch.equal(Pg[:, 1, 1], torch.ones(n_total, dtype=torch.float64))
    # without n_total the shard sizes are exchanged first; same result
    Pg2, cg2 = sharding.gather_poses(P, cost)
    assert torch.equal(Pg2, Pg) and torch.equal(cg2, cg)
    # a shard that contradicts the block partition is an error, not a silent mis-gather
    if n_total % world == 0:
        try:
            sharding.gather_poses(P[:-1], cost[:-1], n_total=n_total)
            raise AssertionError("expected RuntimeError")
        except RuntimeError:
            pass
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_poses_gloo_world2(n_total):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n_total), nprocs=2, join=True)


def test_single_process_passthrough():
    P = torch.eye(4, dtype=torch.float64)[None]
    c = torch.tensor([1.0], dtype=torch.float64)
    Pg, cg = sharding.gather_poses(P, c)
    assert Pg is P and cg is c
    rec = sharding.pack_records(P, c)
    P2, c2 = sharding.unpack_records(rec)
    assert torch.equal(P2, P) and torch.equal(c2, c)
