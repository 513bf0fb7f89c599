"""Multi-GPU bookkeeping for the intersection path.

The BVH is replicated on every GPU; a batch of rays is cut into contiguous per-rank slices
[rank*N/G, (rank+1)*N/G) and every rank traces its own slice -- there is no exchange during
traversal.  The only collective step is collecting the RayHit slices, in rank order, in one buffer
on the destination rank.  Everything here is plain torch.distributed and works with the gloo
backend on CPU tensors (tests) and with NCCL on CUDA tensors (bench.py --gather nccl); the NVLink
peer-memory push used by default on GPUs lives in the C ABI (lrb_trace_gather), and so does the film merge
(FilmMerger -> lrb_film_reduce).
"""
import torch
import torch.distributed as dist

HIT_BYTES = 20


def rank_slice(n_total, world, rank):
    """Contiguous slice of a batch owned by `rank`; sizes differ by at most one ray."""
    base, rem = divmod(int(n_total), int(world))
    first = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return first, count


def slice_table(n_total, world):
    return [rank_slice(n_total, world, r) for r in range(world)]


def rank_seed(base_seed, rank):
    """Every rank regenerates its own slice from a counter-based seed (no 48 B/ray scatter)."""
    return int(base_seed) + 7919 * int(rank)


def gather_hits(hits_local, dst=0, counts=None):
    """Collective gather of uint8 [n_r, 20] RayHit slices onto `dst`, concatenated in rank order.
    Returns the [sum n_r, 20] tensor on dst, None elsewhere.  Slices may have different lengths."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    assert hits_local.dtype == torch.uint8 and hits_local.dim() == 2 and hits_local.shape[1] == HIT_BYTES
    if counts is None:
        c = torch.tensor([hits_local.shape[0]], dtype=torch.int64, device=hits_local.device)
        allc = [torch.zeros_like(c) for _ in range(world)]
        dist.all_gather(allc, c)
        counts = [int(x.item()) for x in allc]
    if rank == dst:
        out = torch.empty((sum(counts), HIT_BYTES), dtype=torch.uint8, device=hits_local.device)
        offs = [0]
        for n in counts:
            offs.append(offs[-1] + n)
        views = [out[offs[r]:offs[r + 1]] for r in range(world)]
    else:
        out, views = None, None
    if len(set(counts)) == 1:
        dist.gather(hits_local, views, dst=dst)
    else:
        # uneven slices: point-to-point
        if rank == dst:
            views[dst].copy_(hits_local)
            reqs = [dist.irecv(views[r], src=r) for r in range(world) if r != dst]
            for q in reqs:
                q.wait()
        else:
            dist.send(hits_local.contiguous(), dst=dst)
    return out


def film_slice(n_floats, world, rank):
    """Slice [first, first + count) of a film of n_floats floats that `rank` sums in FilmMerger.merge: whole
    groups of four floats (16-byte vector loads) while n_floats allows it; the last rank takes the remainder."""
    groups = n_floats // 4
    first, count = rank_slice(groups, world, rank)
    first, count = first * 4, count * 4
    if rank == world - 1:
        count = n_floats - first
    return first, count


class FilmMerger:
    """Merge of per-GPU films on one GPU over NVLink peer memory -- replaces the reference's host-side
    PathOCLRenderEngine::MergeThreadFilms / Film::AddFilm (src/slg/engines/pathocl/pathocl.cpp:184-201,
    src/slg/film/film.cpp:707-760): merged[i] = (((0 + film_0[i]) + film_1[i]) + ...) in device order, binary32.

    Every rank owns a film of n_floats floats in device memory allocated through the C ABI (IPC-exportable);
    all ranks open each other's films and rank `dst`'s merged film (CUDA IPC, lrb_ipc_*).  merge() runs ONE kernel
    per rank (lrb_film_reduce) over that rank's slice of the film: peer loads pull the slice of every film, the sums
    are stored into the merged film on `dst` -- reduce-scatter and gather fused, no NCCL on the data path, and the
    result is bit-identical to the reference's sequential loop whatever the rank count."""

    def __init__(self, dev, n_floats, dst=0):
        self.dev, self.n, self.dst = dev, int(n_floats), dst
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.film = dev.alloc(self.n * 4)
        self.merged_local = dev.alloc(self.n * 4) if self.rank == dst else 0
        mine = (dev.ipc_get_handle(self.film), dev.ipc_get_handle(self.merged_local) if self.rank == dst else None)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine)
        self._opened = []
        self.tiles = []
        for r, (hf, hm) in enumerate(handles):
            if r == self.rank:
                self.tiles.append(self.film)
            else:
                p = dev.ipc_open_handle(hf)
                self._opened.append(p)
                self.tiles.append(p)
        if self.rank == dst:
            self.merged = self.merged_local
        else:
            self.merged = dev.ipc_open_handle(handles[dst][1])
            self._opened.append(self.merged)

    def merge(self):
        """Collective.  The caller's writes to its film must be complete on its stream (this synchronises it)."""
        self.dev.sync()
        dist.barrier()                  # every film is final
        first, count = film_slice(self.n, self.world, self.rank)
        self.dev.film_reduce(self.tiles, self.merged, first, count)
        self.dev.sync()
        dist.barrier()                  # every slice of the merged film has landed on dst

    def close(self):
        dist.barrier()
        for p in self._opened:
            self.dev.ipc_close_handle(p)
        self._opened = []
        dist.barrier()
        self.dev.free(self.film)
        if self.merged_local:
            self.dev.free(self.merged_local)


def max_over_ranks(value, device="cpu"):
    """Timing rule of the benchmark: a multi-GPU number is the MAX over ranks."""
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device="cpu"):
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
