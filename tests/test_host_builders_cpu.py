"""Host-layer builders (product code, luxcore_b200/host/bvhbuild.cpp) against the oracle -- no GPU.

* CLASSIC must reproduce the oracle's (= the reference's) BVHArrayNode arrays bit for bit: two
  independent restatements of bvhclassicbuild.cpp agreeing is one of the pins of the oracle.
* The SAH builder (stand-in for Embree) must emit a well-formed array over every triangle; parity on
  such trees is then checked by walking the SAME array with the oracle's Intersect.
"""
import numpy as np
import pytest

import helpers as H
from luxcore_b200 import hostapi, rays as R, scenes as S
from oracle import oracle as O


def _session(desc, builder, tree_type=4, extra=None):
    cfg = {"accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": tree_type}
    cfg.update(extra or {})
    return hostapi.Session(cfg, desc)


@pytest.mark.parametrize("name,tree_type", [("cornell", 2), ("cornell", 4), ("cornell", 8), ("bigmonkey", 4), ("luxball", 8), ("kitchen", 4)])
def test_classic_builder_bit_identical(name, tree_type):
    desc = S.load_fixture(name)
    s = _session(desc, "CLASSIC", tree_type)
    assert s.build_accelerator("BVH") == hostapi.ACCEL_BVH
    got = s.bvh_nodes()
    ref = O.BVH(H.oracle_scene(desc), tree_type=tree_type).nodes()
    assert got.shape == ref.shape
    assert got.tobytes() == ref.tobytes()


def test_classic_builder_costsamples():
    desc = S.load_fixture("bigmonkey")
    s = _session(desc, "CLASSIC", 4, {"accelerator.bvh.costsamples": 8})
    s.build_accelerator("BVH")
    ref = O.BVH(H.oracle_scene(desc), tree_type=4, cost_samples=8).nodes()
    assert s.bvh_nodes().tobytes() == ref.tobytes()


def _check_tree(nodes, n_tris_expected, max_children):
    nd = nodes["nodeData"]
    n = nodes.shape[0]
    assert H.Emu.lib().emu_validate_tree(nodes.ctypes.data, n) == 0
    leaf = (nd & 0x80000000) != 0
    assert int(leaf.sum()) == n_tris_expected
    # arity
    skip = nd & 0x7FFFFFFF
    for i in np.nonzero(~leaf)[0][:2000]:
        c, k = i + 1, 0
        while c < skip[i]:
            k += 1
            c = skip[c]
        assert 2 <= k <= max_children


@pytest.mark.parametrize("name,tree_type", [("cornell", 4), ("bigmonkey", 2), ("kitchen", 4), ("classroom", 8)])
def test_sah_builder_tree_and_oracle_walk(name, tree_type):
    desc = S.load_fixture(name)
    s = _session(desc, "EMBREE_BINNED_SAH", tree_type)
    s.build_accelerator("BVH")
    nodes = s.bvh_nodes()
    _check_tree(nodes, desc.triangle_count(), tree_type)
    # every (mesh, triangle) appears exactly once
    leaf = (nodes["nodeData"] & 0x80000000) != 0
    keys = nodes["w"][leaf][:, 3].astype(np.uint64) << np.uint64(32) | nodes["w"][leaf][:, 4].astype(np.uint64)
    assert np.unique(keys).shape[0] == desc.triangle_count()

    osc = H.oracle_scene(desc)
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo, hi, 3000, seed=9))
    walk = O.BVH(osc, nodes=nodes).intersect(rays)
    brute, second = osc.brute(rays, want_second=True)
    # topology-free pin: same closest hit as testing every triangle, except exact/near ties
    with np.errstate(invalid="ignore"):
        tie = np.abs(second - brute["t"]) <= 1e-5 * np.maximum(1.0, np.abs(brute["t"]))
    same = (walk["meshIndex"] == brute["meshIndex"]) & ((walk["triangleIndex"] == brute["triangleIndex"]) | (brute["meshIndex"] == H.NULL))
    assert (same | tie).all()
    assert (walk["t"][same & (brute["meshIndex"] != H.NULL)] == brute["t"][same & (brute["meshIndex"] != H.NULL)]).all()

    # and the product's wide re-layout of that tree reproduces the oracle's walk of it exactly
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    rep = H.compare_hits(emu.trace(rays), walk, rays, what="sah/" + name)
    assert rep["bit_exact_hits"] == rep["hits"]


def test_host_geometry_matches_oracle():
    rng = np.random.default_rng(1)
    for v in [0.0, 1.0, -3.5, 1e-7, 123456.0, 1e20]:
        assert hostapi.machine_epsilon(v) == O.machine_epsilon(v)
    for _ in range(50):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        m[3] = [0, 0, 0, 1]
        assert hostapi.matrix_inverse(m).tobytes() == O.matrix_inverse(m).tobytes()
    with pytest.raises(hostapi.HostError):
        hostapi.matrix_inverse(np.zeros((4, 4), np.float32))


def _expected_node_visits(nodes):
    """Sum over inner nodes of area(node) / area(root): the expected number of inner-node visits of a long
    random ray, the quantity the SAH builder minimises (one triangle per leaf: the leaf term is constant)."""
    leaf = (nodes["nodeData"] & 0x80000000) != 0
    box = nodes["w"][~leaf][:, :6].copy().view(np.float32).astype(np.float64)
    d = box[:, 3:] - box[:, :3]
    area = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    return float(area.sum() / area[0])


@pytest.mark.parametrize("name,tree_type", [("kitchen", 4), ("classroom", 4), ("bigmonkey", 8)])
def test_sah_builder_optimised_tree_is_cheaper_and_deterministic(name, tree_type, monkeypatch):
    """The default builder (binary SAH -> insertion-based optimisation -> optimal k-ary collapse) must beat
    the plain top-down k-ary builder it replaced on the cost both minimise, and give the same array twice."""
    desc = S.load_fixture(name)

    def build():
        s = _session(desc, "EMBREE_BINNED_SAH", tree_type)
        s.build_accelerator("BVH")
        return s.bvh_nodes().copy()
    a = build()
    b = build()
    assert a.tobytes() == b.tobytes()
    _check_tree(a, desc.triangle_count(), tree_type)
    monkeypatch.setenv("LRB_BVH_OPT", "0")
    legacy = build()
    _check_tree(legacy, desc.triangle_count(), tree_type)
    assert _expected_node_visits(a) < 0.97 * _expected_node_visits(legacy), (_expected_node_visits(a), _expected_node_visits(legacy))
    assert a.shape[0] <= legacy.shape[0]


@pytest.mark.parametrize("name", ["cornell", "bigmonkey-instances", "bigmonkey-motion"])
def test_dataset_bbox_and_bsphere(name):
    """DataSet::GetBBox = union of the mesh boxes (instances / motion included, dataset.cpp:91-104);
    GetBSphere = BBox::BoundingSphere (bbox.cpp:65-75): centre of the box, radius to its max corner."""
    desc = S.load_fixture(name)
    s = _session(desc, "CLASSIC")
    lo, hi, c, rad = s.dataset_bounds()
    boxes = np.stack([s.mesh_bbox(i) for i in range(len(desc.meshes))])
    assert np.array_equal(lo, boxes[:, :3].min(axis=0)) and np.array_equal(hi, boxes[:, 3:].max(axis=0))
    osc = H.oracle_scene(desc)
    for i in range(len(desc.meshes)):
        assert np.array_equal(boxes[i], np.asarray(osc.mesh_bbox(i), dtype=np.float32).reshape(6))
    cc = (lo + hi) * np.float32(0.5)
    assert np.array_equal(c, cc)
    d = (cc - hi).astype(np.float32)
    assert rad == float(np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))


def _one_mesh_scene(name, verts, tris):
    d = S.SceneDesc(name)
    d.add_shape(np.asarray(verts, np.float32), np.asarray(tris, np.uint32))
    d.add_plain(0)
    return d


@pytest.mark.parametrize("kind", ["identical", "nested", "collinear", "points"])
@pytest.mark.parametrize("tree_type", [2, 4, 8])
def test_sah_builder_keeps_degenerate_geometry_shallow(kind, tree_type):
    """Coincident, nested, collinear and zero-size triangles: equal costs must not let the optimisation turn
    the tree into a chain (the traversal stack, and the kernel's spill buffers, are sized by the worst-case
    depth), searches through boxes that all overlap are bounded, and results stay those of the oracle."""
    rng = np.random.default_rng(3)
    n = 3000
    if kind == "identical":
        verts = np.tile(np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (n, 1))
    elif kind == "nested":
        s = (1.002 ** np.arange(n)).astype(np.float32)
        verts = np.stack([np.stack([-s, -s, 0 * s], 1), np.stack([s, -s, 0 * s], 1), np.stack([0 * s, s, 0 * s], 1)], 1).reshape(-1, 3)
    elif kind == "collinear":
        x = rng.random(3 * n).astype(np.float32)
        verts = np.stack([x, np.zeros_like(x), np.zeros_like(x)], 1)
    else:
        g = np.stack(np.meshgrid(np.arange(15), np.arange(20), np.arange(10), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
        verts = np.repeat(g, 3, axis=0)
    desc = _one_mesh_scene(kind, verts, np.arange(3 * n).reshape(-1, 3))
    s = _session(desc, "EMBREE_BINNED_SAH", tree_type)
    s.build_accelerator("BVH")
    nodes = s.bvh_nodes().copy()
    _check_tree(nodes, n, tree_type)
    osc = H.oracle_scene(desc)
    verts2, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts2, offs)
    assert emu.info()["stack_need"] <= 8 * int(np.ceil(np.log2(n))) + 32, emu.info()
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo - 0.5, hi + 0.5, 3000, seed=3))
    ref = O.BVH(osc, nodes=nodes).intersect(rays)
    rep = H.compare_hits_tie_aware(emu.trace(rays), ref, rays, osc, what=kind, max_ties=3000)
    assert rep["index_mismatch"] == 0 and rep["value_mismatch"] == 0
