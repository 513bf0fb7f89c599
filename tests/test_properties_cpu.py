"""The size-independent property checkers of tests/properties.py, exercised on the CPU: the path under test
is the emulation of the product's re-layout + traversal body (same traverse.h as the kernels), the batch is
the bench workload's ray kind at a size the CPU handles.  The GPU twin (test_gpu_zz_full_size_properties.py)
runs the same checkers through the C ABI at BASELINE.json's full batch sizes."""
import numpy as np
import pytest
import torch

import helpers as H
import properties as P
from luxcore_b200 import hostapi, rays as R, scenes as S
from oracle import oracle as O


def _emu_trace_fn(emu):
    def trace_fn(rays_u8):
        h = emu.trace(R.to_numpy_rays(rays_u8))
        return torch.from_numpy(h.view(np.uint8).reshape(-1, 20).copy())
    return trace_fn


def _tri_tables(desc):
    p0, e1, e2, offs = S.world_triangles(desc)
    return torch.from_numpy(p0), torch.from_numpy(e1), torch.from_numpy(e2), torch.from_numpy(offs)


@pytest.mark.parametrize("name,depth", [("kitchen", 2), ("cornell", 1)])
def test_properties_bvh_emulation(name, depth):
    import bench as B
    desc = S.load_fixture(name)
    s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    s.build_accelerator("BVH")
    nodes = s.bvh_nodes().copy()
    osc = H.oracle_scene(desc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    trace_fn = _emu_trace_fn(emu)
    rays = B.make_bounce_batch(trace_fn, desc, 60000, seed=2, device="cpu", depth=depth)
    hits, rep = P.check_all(trace_fn, rays, *_tri_tables(desc))
    print(name, rep)
    assert rep["hits"] > 0.5 * rays.shape[0] and rep["minimality_rays"] == rep["hits"] == rep["reachability_rays"]
    # and the batch agrees with the oracle, so the properties were checked on the right answers
    ref = O.BVH(osc, nodes=nodes).intersect(R.to_numpy_rays(rays))
    got = hits.numpy().reshape(-1).view(H.HIT_DTYPE)
    assert H.compare_hits(got, ref)["bit_exact_hits"] == rep["hits"]


def test_property_checkers_catch_wrong_answers():
    """A checker that cannot fail checks nothing: corrupt the path under test in four ways."""
    desc = S.load_fixture("cornell")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    good = _emu_trace_fn(emu)
    rays = R.camera_rays(desc.cam, 100, 100, seed=1)
    tables = _tri_tables(desc)
    hits = good(rays)
    P.check_all(good, rays, *tables)

    def farther(r):         # reports a hit that is not the closest: t stretched
        h = good(r).clone()
        f = h.view(torch.float32).view(-1, 5)
        f[:, 0] = torch.where(h.view(torch.int32).view(-1, 5)[:, 3] != -1, f[:, 0] * 1.01, f[:, 0])
        return h
    with pytest.raises(AssertionError):
        P.check_consistency(rays, farther(rays), *tables)

    state = {"calls": 0}
    def leaky(r):           # result depends on the position in the batch
        h = good(r).clone()
        h.view(torch.int32).view(-1, 5)[::977, 4] ^= (state["calls"] & 1)
        state["calls"] += 1
        return h
    with pytest.raises(AssertionError):
        P.check_determinism(leaky, rays, leaky(rays))

    def ignores_maxt(r):    # keeps reporting hits beyond maxt
        rr = r.clone()
        R.rays_f32(rr)[:, 7] = float("inf")
        return good(rr)
    with pytest.raises(AssertionError):
        P.check_minimality(ignores_maxt, rays, hits)

    def too_eager(r):       # drops hits in the last 0.2 % of the interval
        rr = r.clone()
        R.rays_f32(rr)[:, 7] *= (1.0 - 2e-3)
        h = good(rr).clone()
        f = h.view(torch.float32).view(-1, 5)
        missed = h.view(torch.int32).view(-1, 5)[:, 3] == -1
        f[:, 0] = torch.where(missed, R.rays_f32(r)[:, 7], f[:, 0])
        return h
    with pytest.raises(AssertionError):
        P.check_reachability(too_eager, rays, hits)
