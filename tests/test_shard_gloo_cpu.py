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


class _FileDevice:
    """Stand-in for capi.Device in the gloo test of shard.FilmMerger: alloc = a file-backed float32 array, the IPC
    handle = its path, film_reduce = the kernel's arithmetic (sequential binary32 sum in tile order) in numpy."""

    def __init__(self, directory, rank):
        self.dir, self.rank, self.maps, self.count = directory, rank, {}, 0

    def alloc(self, nbytes):
        self.count += 1
        path = os.path.join(self.dir, "film_r%d_%d.bin" % (self.rank, self.count))
        m = np.lib.format.open_memmap(path, mode="w+", dtype=np.float32, shape=(nbytes // 4,))
        m[:] = 0
        m.flush()
        key = 1000 * (self.rank + 1) + self.count
        self.maps[key] = (path, m)
        return key

    def ipc_get_handle(self, key):
        return self.maps[key][0].encode()

    def ipc_open_handle(self, handle):
        self.count += 1
        key = 1000 * (self.rank + 1) + self.count
        self.maps[key] = (handle.decode(), np.load(handle.decode(), mmap_mode="r+"))
        return key

    def ipc_close_handle(self, key):
        self.maps[key][1].flush()
        del self.maps[key]

    def view(self, key, n):
        return self.maps[key][1]

    def film_reduce(self, tiles, dst, first, count):
        acc = np.zeros(count, dtype=np.float32)
        for t in tiles:
            src = np.load(self.maps[t][0], mmap_mode="r")      # a fresh mapping sees the other ranks' flushed writes
            acc = acc + np.asarray(src[first:first + count], dtype=np.float32)
        out = self.maps[dst][1]
        out[first:first + count] = acc
        out.flush()

    def sync(self):
        for _, m in self.maps.values():
            m.flush()

    def free(self, key):
        self.maps.pop(key, None)


def test_film_slices_cover_the_film():
    for n in (0, 1, 3, 4, 5, 1023, 1024, 4 * 3840 * 2160):
        for world in (1, 2, 3, 8):
            cur = 0
            for r in range(world):
                first, count = shard.film_slice(n, world, r)
                assert first == cur and count >= 0
                if r < world - 1:
                    assert first % 4 == 0 and count % 4 == 0       # vector path of the kernel
                cur += count
            assert cur == n


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
    # film merge: host side of shard.FilmMerger (handle exchange, slices, ordering) over a stand-in device whose
    # "device memory" is a file every rank can map and whose film_reduce is the sequential float32 sum of the kernel
    n_floats = 4 * 8 * 8 * 4 + 3
    fm = shard.FilmMerger(_FileDevice(os.path.dirname(out_path), rank), n_floats, dst=0)
    rng = np.random.default_rng(100 + rank)
    film = fm.dev.view(fm.film, n_floats)
    film[:] = (rng.random(n_floats, dtype=np.float32) * np.float32(10.0 ** rng.integers(-3, 4))).astype(np.float32)
    film.flush()
    fm.merge()
    if rank == 0:
        want = np.zeros(n_floats, dtype=np.float32)
        for r in range(world):      # Film::AddFilm order: device 0, 1, ...
            rr = np.random.default_rng(100 + r)
            want = want + (rr.random(n_floats, dtype=np.float32) * np.float32(10.0 ** rr.integers(-3, 4))).astype(np.float32)
        got = np.array(fm.dev.view(fm.merged, n_floats))
        assert got.tobytes() == want.astype(np.float32).tobytes()
    fm.close()
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
