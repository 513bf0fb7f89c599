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
