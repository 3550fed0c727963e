"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- contiguous row shards that
tile the batch, ragged gather back into global order, max-over-ranks timing reduction, rank-independent sample offsets.
No kernels run here (no GPU); the device path of every shard is exactly the 1-GPU path tested in the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jammy_flows_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_rows, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        full = torch.randn(n_rows, 3, generator=g, dtype=torch.float64)      # same on every rank
        mine = sharding.shard_rows(full)
        lo, hi = sharding.shard_range(n_rows, rank, world)
        assert mine.shape[0] == hi - lo and torch.equal(mine, full[lo:hi])
        assert sharding.sample_seed_offset(n_rows, rank, world) == lo
        # stand-in for the per-shard device result: any row-wise function
        local = (mine * 2.0 + 1.0).sum(dim=1, keepdim=True)
        gathered = sharding.gather_rows(local, n_rows)
        assert torch.equal(gathered, (full * 2.0 + 1.0).sum(dim=1, keepdim=True))
        t = sharding.max_over_ranks(10.0 + rank)
        assert t == 10.0 + world - 1
        np.save(os.path.join(out_dir, "ok_%d.npy" % rank), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [7, 1])
def test_two_rank_sharding_over_gloo(tmp_path, n_rows):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rows, str(tmp_path)), nprocs=world, join=True)
    ranges = [np.load(tmp_path / ("ok_%d.npy" % r)) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_rows
    assert all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))


def test_shard_ranges_tile_any_batch():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 8, 10_000_001):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
