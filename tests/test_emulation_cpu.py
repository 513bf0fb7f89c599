"""CPU-only checks of the product's re-layout + per-ray traversal body against the oracle.

The traversal body (luxcore_b200/csrc/traverse.h) is compiled for the host by tests/cpp only for
these tests; on the GPU box the same body runs inside the CUDA kernels (tests/test_gpu_*.py)."""
import numpy as np
import pytest
import torch

import helpers as H
from luxcore_b200 import rays as R
from luxcore_b200 import scenes as S
from oracle import oracle as O


def _rays_for(desc, n, seed):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1))
    return np.concatenate([a, b])


@pytest.mark.parametrize("name,tree_type", [("cornell", 4), ("cornell", 2), ("cornell", 8), ("bigmonkey", 4), ("kitchen", 4)])
def test_bvh_emulation_matches_oracle(name, tree_type):
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    nodes = bvh.nodes()
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    rays = _rays_for(desc, 20000 if name != "kitchen" else 40000, seed=11)
    ref = bvh.intersect(rays)
    got, st = emu.trace(rays, want_stats=True)
    rep = H.compare_hits(got, ref, rays, what="%s k=%d" % (name, tree_type))
    assert rep["hits"] > 0.3 * rep["n"]
    # single-level BVH: identical arithmetic => identical floats
    assert rep["bit_exact_hits"] == rep["hits"]
    assert st["max_stack"] <= emu.info()["stack_need"]


@pytest.mark.parametrize("name,tree_type,n", [("cornell", 4, 60000), ("bigmonkey", 4, 60000), ("kitchen", 4, 120000), ("kitchen", 8, 40000),
                                              ("luxball", 2, 40000)])
def test_bvh_emulation_grazing_rays(name, tree_type, n):
    """Rays that start on (or a few epsilons off) the surfaces, axis-parallel ones included: their
    hits lie at the very planes of the leaf boxes and of the quantized node grids."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n, seed=17))
    # also from far outside the scene (large |o| against small boxes)
    lo, hi = desc.bbox()
    far = R.to_numpy_rays(R.uniform_rays(lo - 40 * (hi - lo), hi + 40 * (hi - lo), n // 4, seed=19))
    c = 0.5 * (lo + hi)
    far["d"] = (c[None, :] + 0.3 * (hi - lo)[None, :] * (np.random.default_rng(5).random((far.shape[0], 3)).astype(np.float32) - 0.5)) - far["o"]
    rays = np.concatenate([rays, far])
    ref = bvh.intersect(rays)
    got = emu.trace(rays)
    rep = H.compare_hits(got, ref, rays, what="grazing %s k=%d" % (name, tree_type))
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
