"""Multi-GPU bookkeeping for the intersection path.

The BVH is replicated on every GPU; a batch of rays is cut into contiguous per-rank slices
[rank*N/G, (rank+1)*N/G) and every rank traces its own slice -- there is no exchange during
traversal.  The only collective step is collecting the RayHit slices, in rank order, in one buffer
on the destination rank.  Everything here is plain torch.distributed and works with the gloo
backend on CPU tensors (tests) and with NCCL on CUDA tensors (bench.py --gather nccl); the NVLink
peer-memory push used by default on GPUs lives in the C ABI (lrb_trace_gather).
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


def reduce_film_tiles(tile, dst=0):
    """Sum-reduce of per-rank film tiles (fp32 radiance/weight planes of identical shape) onto `dst` --
    the collective that replaces the reference's host-side Film::AddFilm merge of per-device films
    (src/slg/engines/pathocl/pathocl.cpp:184-206) when every GPU renders its own samples of the same
    tile.  In place; NCCL on CUDA tensors, gloo on CPU tensors.  Returns the tile on dst, None elsewhere."""
    assert tile.dtype == torch.float32
    dist.reduce(tile, dst=dst, op=dist.ReduceOp.SUM)
    return tile if dist.get_rank() == dst else None


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
