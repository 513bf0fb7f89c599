"""GPU BVH builder (lrb_build_lbvh, luxcore_b200/csrc/build_kernels.cuh): the array it returns obeys the
reference's array rules (bvhclassicbuild.cpp:181-220), bounds what it must, and -- like any tree -- gives the
reference's closest hits: the oracle walks the very array the GPU built, the CUDA traversal must agree bit for bit.
Drop-in for BuildEmbreeBVHMorton (bvhembreebuild.cpp:218-336), reached through the host layer's EMBREE_MORTON."""
import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


def check_array(nodes, boxes, tree_type):
    """Array rules + bounds.  nodes: [m] 32-byte records; leaves carry their input index in word 0."""
    n = boxes.shape[0]
    m = nodes.shape[0]
    nd = nodes["nodeData"].astype(np.int64)
    leaf = (nd >> 31) == 1
    skip = nd & 0x7FFFFFFF
    assert skip[0] == m, "root skip index != node count"
    assert int(leaf.sum()) == n
    assert sorted(nodes["w"][leaf, 0].tolist()) == list(range(n)), "every input leaf exactly once"
    assert (skip[leaf] == np.nonzero(leaf)[0] + 1).all()
    fbox = nodes["w"].view(np.float32)          # inner nodes: min xyz, max xyz
    # children of every inner node, boxes bottom-up (children have larger indices: reverse sweep)
    lo = np.zeros((m, 3), np.float32); hi = np.zeros((m, 3), np.float32)
    li = nodes["w"][:, 0]
    lo[leaf] = boxes[li[leaf], :3]; hi[leaf] = boxes[li[leaf], 3:]
    lo[~leaf] = fbox[~leaf, :3]; hi[~leaf] = fbox[~leaf, 3:]
    max_kids = 0
    for i in np.nonzero(~leaf)[0][::-1]:
        c, kids = i + 1, []
        assert i + 1 < skip[i] <= m
        while c < skip[i]:
            kids.append(c)
            c = skip[c]
        assert c == skip[i], "children do not tile the parent's range"
        assert 1 <= len(kids) <= tree_type
        max_kids = max(max_kids, len(kids))
        with np.errstate(invalid="ignore"):
            assert np.array_equal(np.min(lo[kids], axis=0), lo[i]) and np.array_equal(np.max(hi[kids], axis=0), hi[i]), "box != union of the children"
    return max_kids


@pytest.mark.parametrize("quality", [0, 1])
@pytest.mark.parametrize("tree_type", [2, 4, 8])
@pytest.mark.parametrize("n,kind", [(1, "uniform"), (2, "uniform"), (3, "uniform"), (7, "same"), (1000, "uniform"), (1000, "same"),
                                    (50000, "uniform"), (50000, "line"), (50000, "same"), (200000, "clustered")])
def test_lbvh_array_rules(dev, n, kind, tree_type, quality):
    """quality 0 = radix tree (lrb_build_lbvh), 1 = PLOC; (50000, "same"): identical boxes must not degrade into one merge per iteration."""
    rng = np.random.default_rng(n * 8 + tree_type)
    if kind == "uniform":
        c = rng.random((n, 3), dtype=np.float32) * 10 - 5
    elif kind == "same":            # identical Morton codes: told apart by position
        c = np.full((n, 3), 1.25, np.float32)
    elif kind == "line":            # degenerate extent on two axes
        c = np.zeros((n, 3), np.float32); c[:, 1] = np.linspace(-3, 3, n, dtype=np.float32)
    else:
        k = rng.integers(0, 12, n)
        c = (rng.standard_normal((n, 3)) * 0.01 + rng.random((12, 3))[k] * 100).astype(np.float32)
    e = (rng.random((n, 3), dtype=np.float32) * 0.05).astype(np.float32)
    boxes = np.concatenate([c - e, c + e], axis=1).astype(np.float32)
    nodes, tm = dev.build_lbvh(boxes, tree_type, node_dtype=O.NODE_DTYPE, quality=quality)
    assert n <= nodes.shape[0] <= max(1, 2 * n - 1)
    kids = check_array(nodes, boxes, tree_type)
    if n >= 1000 and kind == "uniform":
        assert kids == tree_type            # the collapse really produces wide nodes
        assert nodes.shape[0] < (2 * n - 1 if tree_type > 2 else 2 * n)
        if tree_type == 4:
            assert nodes.shape[0] - n <= 0.6 * n        # greedy collapse: well under the n - 1 inner nodes of the binary tree
    if quality == 1 and n >= 1000:
        assert tm.kernels < 3 * 400 + 200, "PLOC needed %d launches" % tm.kernels
    assert H.Emu.lib().emu_validate_tree(nodes.ctypes.data, nodes.shape[0]) == 0        # the product's own upload check


@pytest.mark.parametrize("builder", ["EMBREE_MORTON", "B200_PLOC"])
@pytest.mark.parametrize("name,tree_type,n_rays", [("cornell", 4, 200000), ("kitchen", 4, 600000), ("kitchen", 8, 300000), ("bigmonkey", 2, 300000)])
def test_embree_morton_through_the_host_layer(name, tree_type, n_rays, builder):
    """accelerator.bvh.builder.type = EMBREE_MORTON / B200_PLOC: BVHAccel::Init hands its leaf list to the GPU builder; tracing
    the result on the GPU equals the oracle walking the same array."""
    desc = S.load_fixture(name)
    s = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": tree_type}, desc)
    s.start(0)
    nodes = s.bvh_nodes()
    n_tris = sum(t.shape[0] for _, t in desc.shapes) if all(m.kind == S.PLAIN for m in desc.meshes) else None
    nd = nodes["nodeData"]
    assert (nd[0] & 0x7FFFFFFF) == nodes.shape[0]
    if n_tris is not None:
        assert int(((nd >> 31) == 1).sum()) == sum(desc.shapes[m.shape][1].shape[0] for m in desc.meshes)
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n_rays, seed=5)),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 256, 256, seed=6))])
    got = s.trace_host(rays)
    ref = O.BVH(H.oracle_scene(desc), nodes=nodes).intersect(rays)
    rep = H.compare_hits(got, ref, rays, what="gpu-built/%s/k=%d" % (name, tree_type))
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    # same hits as a tree from another builder (topology never changes a closest hit): t bit for bit
    ref2 = O.BVH(H.oracle_scene(desc), tree_type=4).intersect(rays)
    same = (ref2["meshIndex"] == ref["meshIndex"]) & (ref2["triangleIndex"] == ref["triangleIndex"])
    assert same.mean() > 0.9999
    assert (ref2["t"][same].view(np.uint32) == ref["t"][same].view(np.uint32)).all()
    s.stop()
    s.close()


@pytest.mark.parametrize("builder", ["EMBREE_MORTON", "B200_PLOC"])
def test_embree_morton_mbvh_root_and_leaves(builder):
    """Two-level scenes: the root tree over the instances and every leaf tree come from the GPU builder."""
    desc = Z.instances_scene(20)
    s = hostapi.Session({"accelerator.bvh.builder.type": builder}, desc)
    s.start(0)
    assert s.accelerator_type() == hostapi.ACCEL_MBVH
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    rays = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, 200000, seed=15))
    got = s.trace_host(rays)
    # the oracle's own (CLASSIC) trees: another topology, the same closest hits
    ref = O.MBVH(H.oracle_scene(desc)).intersect(rays)
    rep = H.compare_hits(got, ref, rays, what="gpu-built/mbvh")
    assert rep["hits"] > 0 and rep["index_mismatch"] == 0 and rep["value_mismatch"] == 0
    s.stop()
    s.close()
