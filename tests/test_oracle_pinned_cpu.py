"""Pins the oracle (oracle/lux_oracle.cpp, a restatement) against the REFERENCE ITSELF: the reference's
own C++ sources compiled from /root/reference into oracle/_ref/libluxrays_ref.so (oracle/ref/Makefile).
Everything is compared bit for bit: MachineEpsilon, Matrix4x4::Inverse, Triangle::Intersect,
BBox::IntersectP, mesh bounding boxes, the CLASSIC builder's BVHArrayNode arrays, BVHAccel::Intersect
and MBVHAccel::Intersect (instances, motion blur, Update) on the fixtures and on stress batches.

Where oracle/_ref is neither prebuilt nor buildable (no /root/reference), the same comparisons run
against the committed vectors the reference produced (tests/golden/ref_vectors.npz, written by
tools/make_ref_vectors.py) -- see test_golden_vectors_from_the_reference."""
import os

import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import rays as R, scenes as S
from oracle import oracle as O
from oracle import refapi as RF

needs_ref = pytest.mark.skipif(not RF.available(), reason="oracle/_ref is not built and /root/reference is absent")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors.npz")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _same_hits(got, ref):
    """RayHit arrays of the oracle and of the reference must be identical in every defined field."""
    assert np.array_equal(got["meshIndex"], ref["meshIndex"])
    hit = ref["meshIndex"] != H.NULL
    assert np.array_equal(got["triangleIndex"][hit], ref["triangleIndex"][hit])
    for f in ("t", "b1", "b2"):
        assert np.array_equal(_bits(got[f][hit]), _bits(ref[f][hit])), f
    assert np.array_equal(_bits(got["t"][~hit]), _bits(ref["t"][~hit]))      # miss: t = ray.maxt
    return int(hit.sum())


def _same_nodes(a, b, root_tree=False):
    """BVHArrayNode arrays: every DEFINED word must be identical (the reference leaves pad0 and the
    unused tail of the leaf payload uninitialised)."""
    a = np.ascontiguousarray(a).view(np.uint32).reshape(-1, 8)
    b = np.ascontiguousarray(b).view(np.uint32).reshape(-1, 8)
    assert a.shape == b.shape
    assert np.array_equal(a[:, 6], b[:, 6])                     # nodeData: leaf flag + skip index
    leaf = (a[:, 6] >> 31) == 1
    assert np.array_equal(a[~leaf, :6], b[~leaf, :6])           # inner: bboxMin, bboxMax
    k = 4 if root_tree else 5                                   # bvhLeaf: 4 indices; triangleLeaf: v[3], mesh, triangle
    assert np.array_equal(a[leaf, :k], b[leaf, :k])
    return a.shape[0]


def _stress_rays(desc, n, seed, time_range=None):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed, time_range=time_range))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1, time_range=time_range))
    p0, e1, e2, _ = S.world_triangles(desc)
    c = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n // 2, seed=seed + 2, axis_fraction=0.3))
    if time_range is not None:
        c["time"] = np.random.default_rng(seed).random(c.shape[0]).astype(np.float32) * (time_range[1] - time_range[0]) + time_range[0]
    return np.concatenate([a, b, c])


@needs_ref
def test_machine_epsilon_bit_identical():
    rng = np.random.default_rng(1)
    vals = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-30, 30, 2000),
                           [0.0, -0.0, 1.0, -1.0, 1e-45, 3.4e38, -3.4e38, np.inf, -np.inf, 0.1, 1e-5, 1e-1]]).astype(np.float32)
    for v in vals:
        a, b = np.float32(RF.epsilon(float(v))), np.float32(O.machine_epsilon(float(v)))
        assert a.tobytes() == b.tobytes() or (np.isnan(a) and np.isnan(b)), v


@needs_ref
def test_matrix_inverse_bit_identical():
    rng = np.random.default_rng(2)
    for k in range(300):
        m = rng.standard_normal((4, 4)).astype(np.float32)
        if k % 3 == 0:      # affine, like instance transforms
            m[3] = [0, 0, 0, 1]
        assert RF.matrix_inverse(m).tobytes() == O.matrix_inverse(m).tobytes()


@needs_ref
def test_triangle_and_box_tests_bit_identical():
    rng = np.random.default_rng(3)
    n = 20000
    rays = R.to_numpy_rays(R.uniform_rays([-1, -1, -1], [1, 1, 1], n, seed=4))
    rays["d"][::7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, rays[::7].shape[0])]      # axis-parallel: 1/d = inf
    rays["mint"][::5] = 0.0
    rays["maxt"][::3] = rng.random(rays[::3].shape[0]).astype(np.float32) * 3
    tri = (rng.random((n, 3, 3)).astype(np.float32) - 0.5) * 2
    tri[::11, 2] = tri[::11, 1]                     # degenerate
    tri[::13, :, 2] = rays["o"][::13, None, 2]      # ray origin in the triangle's plane
    lib_o, lib_r = O.lib(), RF.lib()
    tb_o, tb_r = np.zeros(3, np.float32), np.zeros(3, np.float32)
    n_hit = 0
    for i in range(n):
        r = rays[i:i + 1]
        p = np.ascontiguousarray(tri[i])
        ho = lib_o.orc_triangle_intersect(r.ctypes.data, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data, tb_o.ctypes.data)
        hr = lib_r.ref_triangle_intersect(r.ctypes.data, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data, tb_r.ctypes.data)
        assert ho == hr, i
        if hr:
            assert tb_o.tobytes() == tb_r.tobytes(), i
            n_hit += 1
        lo, hi = np.ascontiguousarray(p.min(axis=0)), np.ascontiguousarray(p.max(axis=0))
        if i % 9 == 0:
            lo[i % 3] = r["o"][0][i % 3]            # origin exactly on a slab plane
        assert lib_o.orc_bbox_intersectp(r.ctypes.data, lo.ctypes.data, hi.ctypes.data) == \
            lib_r.ref_bbox_intersectp(r.ctypes.data, lo.ctypes.data, hi.ctypes.data), i
    assert n_hit > 500


@needs_ref
@pytest.mark.parametrize("name,tree_type,cost_samples", [("cornell", 4, 0), ("cornell", 2, 0), ("cornell", 8, 0), ("bigmonkey", 4, 0),
                                                         ("bigmonkey", 4, 8), ("kitchen", 4, 0), ("luxball", 8, 0)])
def test_classic_builder_and_bvh_intersect_bit_identical(name, tree_type, cost_samples):
    desc = S.load_fixture(name)
    osc, rsc = H.oracle_scene(desc), H.reference_scene(desc)
    ob = O.BVH(osc, tree_type=tree_type, cost_samples=cost_samples)
    rb = RF.BVH(rsc, tree_type=tree_type, cost_samples=cost_samples)
    assert _same_nodes(ob.nodes(), rb.nodes()) > 0              # BVHAccel::Init + CLASSIC builder
    rays = _stress_rays(desc, 40000 if name == "kitchen" else 20000, seed=7)
    assert _same_hits(ob.intersect(rays), rb.intersect(rays)) > 0.3 * rays.shape[0]


@needs_ref
def test_bvh_intersect_on_the_products_sah_tree_bit_identical():
    """The tree bench.py traces (product's binned-SAH builder) walked by the reference's BVHAccel::Intersect
    and by the oracle's restatement."""
    from luxcore_b200 import hostapi
    desc = S.load_fixture("kitchen")
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    nodes = sess.bvh_nodes()
    ob = O.BVH(H.oracle_scene(desc), nodes=nodes)
    rb = RF.BVH(H.reference_scene(desc), nodes=nodes)
    rays = _stress_rays(desc, 40000, seed=9)
    assert _same_hits(ob.intersect(rays), rb.intersect(rays)) > 0.3 * rays.shape[0]


@needs_ref
@pytest.mark.parametrize("which,tree_type,time_range", [("zoo-inst", 4, None), ("zoo-inst", 2, None), ("bigmonkey-instances", 4, None),
                                                        ("lightinstances", 4, None), ("zoo-motion", 4, (-0.1, 1.1)),
                                                        ("bigmonkey-motion", 4, (0.0, 1.0))])
def test_mbvh_bit_identical(which, tree_type, time_range):
    desc = {"zoo-inst": lambda: Z.instances_scene(), "zoo-motion": lambda: Z.motion_scene(),
            "lightinstances": lambda: S.load_fixture("lightinstances", max_objects=300)}.get(which, lambda: S.load_fixture(which))()
    osc, rsc = H.oracle_scene(desc), H.reference_scene(desc)
    for i in range(len(desc.meshes)):       # Mesh::GetBBox: instance corners / motion time samples
        assert osc.mesh_bbox(i).tobytes() == rsc.mesh_bbox(i).tobytes(), i
    om, rm = O.MBVH(osc, tree_type=tree_type), RF.MBVH(rsc, tree_type=tree_type)
    _same_nodes(om.root_nodes(), rm.root_nodes(), root_tree=True)
    assert om.leaf_count() == rm.leaf_count()
    for i in range(om.leaf_count()):
        _same_nodes(om.leaf_nodes(i), rm.leaf_nodes(i))
    assert om.transforms_minv().tobytes() == rm.transforms_minv().tobytes()
    if time_range is not None:
        for i in range(rm.motion_count()):
            for t in np.linspace(time_range[0], time_range[1], 23):
                assert om.motion_sample(i, float(t)).tobytes() == rm.motion_sample(i, float(t)).tobytes(), (i, t)
    rays = _stress_rays(desc, 15000, seed=11, time_range=time_range)
    assert _same_hits(om.intersect(rays), rm.intersect(rays)) > 0.1 * rays.shape[0]


@needs_ref
def test_mbvh_update_bit_identical():
    desc = Z.instances_scene(12)
    osc, rsc = H.oracle_scene(desc), H.reference_scene(desc)
    om, rm = O.MBVH(osc), RF.MBVH(rsc)
    inst = [i for i, m in enumerate(desc.meshes) if m.kind == S.INSTANCE]
    for k, i in enumerate(inst[:3]):
        m = Z.translate(1.5 * (k + 1), -2.0, 0.5) @ Z.rot_z(33.0 * (k + 1))
        osc.set_instance_transform(i, m)
        rsc.set_instance_transform(i, m)
    om.update()
    rm.update()
    _same_nodes(om.root_nodes(), rm.root_nodes(), root_tree=True)
    assert om.transforms_minv().tobytes() == rm.transforms_minv().tobytes()
    rays = _stress_rays(desc, 8000, seed=13)
    _same_hits(om.intersect(rays), rm.intersect(rays))


def test_golden_vectors_from_the_reference():
    """Runs everywhere: the oracle against outputs the reference itself produced (tools/make_ref_vectors.py)."""
    z = np.load(GOLDEN)
    for v, e in zip(z["eps_in"], z["eps_out"]):
        assert np.float32(O.machine_epsilon(float(v))).tobytes() == np.float32(e).tobytes()
    for m, inv in zip(z["minv_in"], z["minv_out"]):
        assert O.matrix_inverse(m).tobytes() == inv.tobytes()
    rays = np.ascontiguousarray(z["tri_rays"]).reshape(-1).view(RF.RAY_DTYPE)
    tri = np.ascontiguousarray(z["tri_verts"])
    lib = O.lib()
    tb = np.zeros(3, np.float32)
    for i in range(rays.shape[0]):
        r, p = rays[i:i + 1], tri[i]
        h = lib.orc_triangle_intersect(r.ctypes.data, p[0].ctypes.data, p[1].ctypes.data, p[2].ctypes.data, tb.ctypes.data)
        assert h == int(z["tri_hit"][i]), i
        if h:
            assert tb.tobytes() == z["tri_tb"][i].tobytes(), i
        lo, hi = np.ascontiguousarray(p.min(axis=0)), np.ascontiguousarray(p.max(axis=0))
        assert lib.orc_bbox_intersectp(r.ctypes.data, lo.ctypes.data, hi.ctypes.data) == int(z["box_hit"][i]), i
    desc = S.load_fixture("cornell")
    for k in (4, 8):
        ob = O.BVH(H.oracle_scene(desc), tree_type=k)
        _same_nodes(ob.nodes(), np.ascontiguousarray(z["cornell%d_nodes" % k]).reshape(-1).view(RF.NODE_DTYPE))
        r = np.ascontiguousarray(z["cornell%d_rays" % k]).reshape(-1).view(RF.RAY_DTYPE)
        ref = np.ascontiguousarray(z["cornell%d_hits" % k]).reshape(-1).view(RF.HIT_DTYPE)
        assert _same_hits(ob.intersect(r), ref) > 0.3 * r.shape[0]
    for name, mk in (("zooinst", Z.instances_scene), ("zoomotion", Z.motion_scene)):
        d = mk()
        osc = H.oracle_scene(d)
        for i in range(len(d.meshes)):
            assert osc.mesh_bbox(i).tobytes() == z[name + "_bboxes"][i].tobytes()
        r = np.ascontiguousarray(z[name + "_rays"]).reshape(-1).view(RF.RAY_DTYPE)
        ref = np.ascontiguousarray(z[name + "_hits"]).reshape(-1).view(RF.HIT_DTYPE)
        assert _same_hits(O.MBVH(osc).intersect(r), ref) > 0.1 * r.shape[0]
