"""Size-independent properties of closest-hit tracing, checked with torch on whatever device the batch
lives on, so that they can run at BASELINE.json's full batch sizes (16 Mi rays) where the oracle would take
minutes.  `trace_fn(rays_u8 [n, 48]) -> hits_u8 [n, 20]` is the path under test (the CUDA device in the
GPU tests; the CPU emulation of the traversal body in the CPU test that keeps these checkers honest).

  determinism     the same batch twice gives the same bytes;
  shards          tracing the two halves separately gives the halves of the whole (what ray sharding
                  over GPUs relies on);
  permutation     tracing a shuffled batch gives the shuffled result (rays are independent: the
                  persistent kernel's dynamic ray assignment must not leak between rays);
  consistency     a hit lies on the triangle it names: o + t d == p0 + b1 e1 + b2 e2, barycentrics in
                  range, mint <= t <= maxt; a miss returns t = maxt and NULL indices;
  minimality      re-tracing every hit ray with maxt = its t must miss: the reference accepts a hit only
                  if t < rayHit->t, which starts at ray.maxt (bvhaccel.cpp:233), so nothing at or
                  beyond maxt is reported -- and nothing closer than the closest hit exists;
  reachability    re-tracing every hit ray with maxt = t (1 + 1e-3) returns the same record again -- for all
                  but a sliver of the rays: the reference's own boxes can cull a hit whose computed t lies
                  BEFORE the computed entry distance of its ancestors' boxes (kitchen: a needle triangle in the
                  plane x = 10.2 reports t = 4.116939 while its boxes are entered at 4.116984; the reference
                  library itself misses it for maxt up to 50 ulp above t), so at most 1e-3 of the rays may
                  answer differently.
"""
import torch

from luxcore_b200 import rays as R


def _same(a, b, what):
    if not torch.equal(a, b):
        bad = torch.nonzero((a != b).any(dim=1))[:, 0]
        raise AssertionError("%s: %d of %d records differ (first at %d)" % (what, bad.shape[0], a.shape[0], int(bad[0])))


def check_determinism(trace_fn, rays, hits):
    _same(trace_fn(rays), hits, "determinism")


def check_shards(trace_fn, rays, hits, parts=2):
    n = rays.shape[0]
    cuts = [n * k // parts for k in range(parts + 1)]
    got = torch.cat([trace_fn(rays[cuts[k]:cuts[k + 1]].contiguous()) for k in range(parts)])
    _same(got, hits, "shard invariance")


def check_permutation(trace_fn, rays, hits, seed=11):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    perm = torch.randperm(rays.shape[0], generator=g).to(rays.device)
    _same(trace_fn(rays[perm].contiguous()), hits[perm], "permutation invariance")


def check_consistency(rays, hits, p0, e1, e2, tri_offs, rel_tol=2e-4):
    """p0/e1/e2: world-space vertex 0 and edges of every triangle; tri_offs[mesh] = first triangle of a mesh."""
    r = R.rays_f32(rays)
    h = R.unpack_hits(hits)
    hit = h["mesh"] != -1
    miss = ~hit
    assert bool((h["tri"][miss] == -1).all()), "a miss must carry NULL indices"
    assert bool((h["t"][miss] == r[miss, 7]).all()), "a miss must return t = maxt"
    idx = torch.nonzero(hit)[:, 0]
    t, b1, b2 = h["t"][idx], h["b1"][idx], h["b2"][idx]
    assert bool((t >= r[idx, 6]).all() and (t <= r[idx, 7]).all()), "hit outside [mint, maxt]"
    assert bool((b1 >= 0).all() and (b2 >= 0).all() and (b1 + b2 <= 1.0 + 1e-6).all()), "barycentrics out of range"
    flat = tri_offs[h["mesh"][idx].long()] + h["tri"][idx].long()
    assert bool((flat >= 0).all() and (flat < p0.shape[0]).all()), "triangle index out of range"
    on_ray = r[idx, 0:3].double() + t.double()[:, None] * r[idx, 3:6].double()
    on_tri = p0[flat].double() + b1.double()[:, None] * e1[flat].double() + b2.double()[:, None] * e2[flat].double()
    scale = torch.maximum(on_tri.abs().amax(dim=1), (t.double() * r[idx, 3:6].double().norm(dim=1))).clamp(min=1.0)
    err = (on_ray - on_tri).norm(dim=1) / scale
    # Triangle::Intersect is plain float Moller-Trumbore: for a ray nearly parallel to its triangle the divisor
    # is small and t, b1, b2 lose digits (the reference's own values, reproduced bit for bit) -- so the bulk
    # must be tight and the tail bounded, not every ray tight.
    if idx.shape[0]:
        k = max(1, int(idx.shape[0] * 1e-4))
        tail = torch.topk(err, k).values
        worst, q9999 = float(tail[0]), float(tail[-1])
    else:
        worst = q9999 = 0.0
    assert q9999 <= rel_tol, "hit points off their triangles: 99.99th percentile of the relative distance %.3g" % q9999
    assert worst <= 5e-2, "hit point off its triangle: relative distance %.3g" % worst
    return {"hits": int(idx.shape[0]), "misses": int(miss.sum()), "worst_point_error": worst, "q9999_point_error": q9999}


def _with_maxt(rays, idx, maxt):
    sub = rays[idx].clone()
    R.rays_f32(sub)[:, 7] = maxt
    return sub


def check_minimality(trace_fn, rays, hits):
    h = R.unpack_hits(hits)
    idx = torch.nonzero(h["mesh"] != -1)[:, 0]
    if idx.shape[0] == 0:
        return 0
    got = R.unpack_hits(trace_fn(_with_maxt(rays, idx, h["t"][idx])))
    closer = got["mesh"] != -1
    assert not bool(closer.any()), "%d rays report a hit with maxt set to their closest hit's t" % int(closer.sum())
    return int(idx.shape[0])


def check_reachability(trace_fn, rays, hits, slack=1e-3, max_fraction=1e-3):
    h = R.unpack_hits(hits)
    idx = torch.nonzero(h["mesh"] != -1)[:, 0]
    if idx.shape[0] == 0:
        return 0
    t = h["t"][idx]
    again = trace_fn(_with_maxt(rays, idx, t * (1.0 + slack)))
    differ = int((again != hits[idx]).any(dim=1).sum())
    assert differ <= max_fraction * idx.shape[0], "re-trace with maxt just above t: %d of %d records differ" % (differ, idx.shape[0])
    return int(idx.shape[0])


def check_all(trace_fn, rays, p0, e1, e2, tri_offs, consistency=True):
    hits = trace_fn(rays)
    rep = {}
    check_determinism(trace_fn, rays, hits)
    check_shards(trace_fn, rays, hits)
    check_permutation(trace_fn, rays, hits)
    if consistency:
        rep.update(check_consistency(rays, hits, p0, e1, e2, tri_offs))
    rep["minimality_rays"] = check_minimality(trace_fn, rays, hits)
    rep["reachability_rays"] = check_reachability(trace_fn, rays, hits)
    return hits, rep
