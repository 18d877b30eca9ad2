"""CPU: batch sharding across ranks (host logic of the N>1 path), incl. a real
world_size-2 gloo run of the max-over-ranks / sum-over-ranks reductions bench.py uses."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def test_shard_bounds_partition():
    from gfla_b200.sharding import shard_bounds
    for n in (1, 7, 16, 17, 64):
        for ws in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, ws, r) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gfla_b200.sharding import reduce_max_time, reduce_sum, shard_batch
    full = torch.arange(10 * 3, dtype=torch.float32).reshape(10, 3)
    (mine,) = shard_batch([full], world, rank)
    ms = reduce_max_time(1.0 + rank)            # slowest rank defines the step time
    units = reduce_sum(float(mine.shape[0]))    # units processed by all ranks
    gathered = [torch.zeros(5, 3) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        torch.save({"ms": ms, "units": units, "cat": torch.cat(gathered)}, out)
    dist.destroy_process_group()


def test_two_rank_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ms"] == 2.0 and r["units"] == 10.0
    assert torch.equal(r["cat"], torch.arange(30, dtype=torch.float32).reshape(10, 3))
