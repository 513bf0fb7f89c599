"""CPU checks of the GPU builder's ALGORITHM through its Python restatement (tests/builder_reference.py): the
ancestor-sum index formula reproduces a depth-first layout, the arrays obey the reference's rules and pass the
product's own upload validation, and -- like any tree -- give the reference's closest hits (the oracle walks the
array, the emulated traversal body must agree).  The CUDA implementation itself is tested in tests/test_gpu_builder.py
with the same checker."""
import importlib.util
import os

import numpy as np
import pytest

import builder_reference as BR
import helpers as H
from luxcore_b200 import rays as R, scenes as S
from oracle import oracle as O

_spec = importlib.util.spec_from_file_location("gpu_builder_tests", os.path.join(os.path.dirname(__file__), "test_gpu_builder.py"))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
check_array = _mod.check_array


@pytest.mark.parametrize("quality", [0, 1])
@pytest.mark.parametrize("tree_type", [2, 4, 8])
@pytest.mark.parametrize("n,kind", [(1, "uniform"), (2, "uniform"), (3, "uniform"), (7, "same"), (300, "uniform"), (200, "same"), (300, "line")])
def test_reference_builder_array_rules(n, kind, tree_type, quality):
    rng = np.random.default_rng(n * 8 + tree_type)
    if kind == "uniform":
        c = rng.random((n, 3), dtype=np.float32) * 10 - 5
    elif kind == "same":
        c = np.full((n, 3), 1.25, np.float32)
    else:
        c = np.zeros((n, 3), np.float32); c[:, 1] = np.linspace(-3, 3, n, dtype=np.float32)
    e = (rng.random((n, 3), dtype=np.float32) * 0.05).astype(np.float32)
    boxes = np.concatenate([c - e, c + e], axis=1).astype(np.float32)
    nodes = BR.build(boxes, tree_type, quality)
    assert n <= nodes.shape[0] <= max(1, 2 * n - 1)
    check_array(nodes, boxes, tree_type)
    assert H.Emu.lib().emu_validate_tree(nodes.ctypes.data, nodes.shape[0]) == 0


@pytest.mark.parametrize("quality", [0, 1])
def test_reference_builder_tree_gives_the_reference_hits(quality):
    desc = S.load_fixture("cornell")
    osc = H.oracle_scene(desc)
    ref_bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    nodes0 = ref_bvh.nodes()
    leaf = (nodes0["nodeData"] >> 31) == 1
    lw = nodes0["w"][leaf]                                          # v0 v1 v2 mesh tri
    V = verts.reshape(-1, 3)
    P = V[lw[:, :3].astype(np.int64) + offs[lw[:, 3]].astype(np.int64)[:, None]]
    boxes = np.concatenate([P.min(1) - 1e-4, P.max(1) + 1e-4], axis=1).astype(np.float32)
    arr = BR.build(boxes, 4, quality)
    L = (arr["nodeData"] >> 31) == 1
    arr["w"][L] = lw[arr["w"][L, 0]]                                 # leaf payload, as the host layer writes it in
    lo, hi = desc.bbox()
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo, hi, 20000, seed=3)), R.to_numpy_rays(R.camera_rays(desc.cam, 96, 96, seed=4))])
    want = O.BVH(osc, nodes=arr).intersect(rays)
    got = H.Emu.bvh(arr, verts, offs).trace(rays)
    rep = H.compare_hits(got, want, rays, what="reference builder, quality %d" % quality)
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    # and the same hits as the CLASSIC tree, t bit for bit (topology never changes a closest hit; exact ties may name another triangle)
    base = ref_bvh.intersect(rays)
    same = (base["meshIndex"] == want["meshIndex"]) & (base["triangleIndex"] == want["triangleIndex"])
    assert same.mean() > 0.999
    assert (base["t"][same].view(np.uint32) == want["t"][same].view(np.uint32)).all()
