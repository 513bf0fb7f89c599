"""N > 1 host logic on CPU: two gloo processes shard a batch, 'trace' their slices (with the oracle,
on a tiny scene) and gather the RayHit slices onto rank 0 in rank order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from luxcore_b200 import shard


def test_rank_slices_cover_batch():
    for n, w in [(10, 3), (16, 4), (7, 8), (0, 2), (1 << 24, 8), (1000003, 7)]:
        t = shard.slice_table(n, w)
        assert t[0][0] == 0 and sum(c for _, c in t) == n
        for (f0, c0), (f1, _) in zip(t, t[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in t) - min(c for _, c in t) <= 1
    assert len({shard.rank_seed(2, r) for r in range(8)}) == 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import helpers as H
    from luxcore_b200 import rays as R, scenes as S
    from oracle import oracle as O
    desc = S.load_fixture("cornell")
    bvh = O.BVH(H.oracle_scene(desc))
    # every rank regenerates the WHOLE batch deterministically and keeps only its slice
    rays_all = R.to_numpy_rays(R.camera_rays(desc.cam, 64, n_total // 64, seed=11))
    first, count = shard.rank_slice(n_total, world, rank)
    hits = bvh.intersect(rays_all[first:first + count], nthreads=1)
    local = torch.from_numpy(hits.view(np.uint8).reshape(-1, 20).copy())
    gathered = shard.gather_hits(local, dst=0)
    # film tiles: every rank contributes its own samples, rank 0 ends up with the sum
    tile = torch.full((4, 8, 8), float(rank + 1), dtype=torch.float32)
    summed = shard.reduce_film_tiles(tile, dst=0)
    if rank == 0:
        assert torch.equal(summed, torch.full((4, 8, 8), float(world * (world + 1) // 2)))
    else:
        assert summed is None
    tmax = shard.max_over_ranks(1.0 + rank)
    total = shard.sum_over_ranks(count)
    if rank == 0:
        ref = bvh.intersect(rays_all, nthreads=1)
        ok = gathered.numpy().tobytes() == ref.view(np.uint8).tobytes()
        with open(out_path, "w") as f:
            f.write("%d %f %d" % (int(ok), tmax, int(total)))
    else:
        assert gathered is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 4096), (3, 4160)])
def test_gloo_shard_trace_gather(tmp_path, world, n_total):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    ok, tmax, total = open(out).read().split()
    assert int(ok) == 1
    assert float(tmax) == float(world)      # max over ranks of (1 + rank)
    assert int(total) == n_total
