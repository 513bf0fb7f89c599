"""lrb_bvh_build_scene (luxcore_b200/csrc/relayout_kernels.cuh + build_kernels.cuh; SURVEY.md 8f "GPU BVH builder"): triangles
in, traceable scene out, everything on the device.  The scene must be, byte for byte, what lrb_bvh_upload's HOST re-layout
(relayout.cpp) makes of the reference array the same call returns; that array must obey the reference's array rules
(bvhclassicbuild.cpp:181-220) and carry every triangle's payload; and tracing it must give the reference's closest hits
(the oracle walks the very array).  Replaces BVHAccel::Init + builder + BVHKernel upload (bvhaccel.cpp:72-168,
bvhembreebuild.cpp:218-336, bvhaccelhw.cpp:38-257) for the builder types EMBREE_MORTON / B200_PLOC of the host layer."""
import numpy as np
import pytest

import helpers as H
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


def _inputs(desc):
    osc = H.oracle_scene(desc)
    verts, voff = H.flattened_from_oracle(desc, osc)
    tri, toff = H.flattened_triangles(desc)
    return osc, verts, voff, tri, toff


def _rays(desc, n, seed):
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    return np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed)),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 200, 200, seed=seed + 1))])


def check_payload(nodes, tri, toff, tree_type):
    """Array rules of the product's own upload check + every triangle exactly once, with the payload a host builder writes
    (vertex indices of THAT triangle, its mesh, its index inside the mesh)."""
    assert H.Emu.lib().emu_validate_tree(nodes.ctypes.data, nodes.shape[0]) == 0
    nd = nodes["nodeData"].astype(np.int64)
    leaf = (nd >> 31) == 1
    assert (nd[0] & 0x7FFFFFFF) == nodes.shape[0]
    assert int(leaf.sum()) == tri.shape[0]
    w = nodes["w"][leaf]
    g = np.asarray(toff, dtype=np.int64)[w[:, 3]] + w[:, 4]
    assert np.array_equal(np.sort(g), np.arange(tri.shape[0])), "every triangle exactly once"
    assert (w[:, 4] < (np.asarray(toff)[w[:, 3] + 1] - np.asarray(toff)[w[:, 3]])).all(), "triangle index inside its mesh"
    assert np.array_equal(w[:, :3], tri[g]), "leaf vertex indices"
    assert (nodes["w"][leaf, 5] == 0).all() and (nodes["pad0"] == 0).all()
    # arity
    skip = nd & 0x7FFFFFFF
    for i in np.nonzero(~leaf)[0][:20000]:
        c, k = i + 1, 0
        while c < skip[i]:
            c = skip[c]
            k += 1
        assert 1 <= k <= tree_type


def check_scene(dev, scene, nodes, osc, verts, voff, rays, what):
    """Scene bytes == host re-layout of the array; hits == the oracle on the array == lrb_bvh_upload of the array."""
    emu = H.Emu.bvh(nodes, verts, voff)
    wide, tris, ids = emu.arrays()
    dw, dt, di = scene.download()
    info = scene.info()
    assert info.n_ref_nodes == nodes.shape[0] and info.n_wide_nodes == wide.shape[0] and info.n_triangles == tris.shape[0]
    assert info.two_level == 0 and info.n_instances == 0
    assert dw.tobytes() == wide.tobytes(), what + ": wide nodes differ from the host re-layout"
    assert dt.tobytes() == tris.tobytes(), what + ": triangle records differ from the host re-layout"
    assert di.tobytes() == ids.tobytes(), what + ": triangle ids differ from the host re-layout"
    assert info.stack_need == emu.info()["stack_need"], what + ": stack bound"
    got = scene.trace_host(rays)
    ref = O.BVH(osc, nodes=nodes).intersect(rays)
    rep = H.compare_hits(got, ref, rays, what=what)
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    up = dev.upload_bvh(nodes, verts, voff)
    assert up.info().device_bytes == info.device_bytes
    got2 = up.trace_host(rays)
    assert got2.tobytes() == got.tobytes(), what + ": lrb_bvh_upload of the same array traces differently"
    up.free()


@pytest.mark.parametrize("name,tree_type,quality,n_rays", [("cornell", 4, 1, 100000), ("kitchen", 4, 1, 400000), ("kitchen", 4, 0, 200000),
                                                           ("kitchen", 8, 1, 200000), ("bigmonkey", 2, 0, 200000), ("classroom", 4, 1, 200000)])
def test_scene_built_on_the_device_is_the_host_layout_of_its_array(dev, name, tree_type, quality, n_rays):
    desc = S.load_fixture(name)
    osc, verts, voff, tri, toff = _inputs(desc)
    before = dev.counters().device_bytes_in_use
    scene, tm, nodes = dev.build_scene(verts, voff, tri, toff, tree_type, quality, want_nodes=True, node_dtype=O.NODE_DTYPE)
    assert tm.kernels > 8 and tm.relayout_ms > 0 and tm.tree_ms > 0
    assert dev.counters().device_bytes_in_use - before == scene.info().device_bytes     # scratch released, the scene accounted
    check_payload(nodes, tri, toff, tree_type)
    check_scene(dev, scene, nodes, osc, verts, voff, _rays(desc, n_rays, 31), "%s k=%d q=%d" % (name, tree_type, quality))
    scene.free()
    assert dev.counters().device_bytes_in_use == before


def test_scene_build_on_a_soup(dev):
    n = 1000000
    desc = S.random_soup(n, seed=4, size=0.002 * (50e6 / n) ** (1.0 / 3.0), name="soup")
    osc, verts, voff, tri, toff = _inputs(desc)
    scene, tm, nodes = dev.build_scene(verts, voff, tri, toff, 4, 1, want_nodes=True, node_dtype=O.NODE_DTYPE)
    check_payload(nodes, tri, toff, 4)
    rays = R.to_numpy_rays(R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 300000, seed=9))
    check_scene(dev, scene, nodes, osc, verts, voff, rays, "soup-1M")
    # without asking for the array: nothing of the tree comes back to the host
    d2h = dev.counters().d2h_bytes
    scene2, tm2, none = dev.build_scene(verts, voff, tri, toff, 4, 1)
    assert none is None and tm2.d2h_ms == 0.0
    assert dev.counters().d2h_bytes - d2h < 4096
    assert scene2.trace_host(rays).tobytes() == scene.trace_host(rays).tobytes()       # the build is deterministic
    scene.free()
    scene2.free()


def test_scene_build_small_and_ragged_inputs(dev):
    rng = np.random.default_rng(3)
    # 1, 2, 3, 5 triangles in one mesh; then several meshes with empty ones in between
    for n in (1, 2, 3, 5):
        verts = (rng.random((3 * n, 3)) * 4 - 2).astype(np.float32)
        tri = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
        scene, tm, nodes = dev.build_scene(verts, [0], tri, [0, n], 4, 1, want_nodes=True, node_dtype=O.NODE_DTYPE)
        assert nodes.shape[0] == (1 if n == 1 else n + 1) or (n == 5 and n + 1 <= nodes.shape[0] <= 2 * n - 1)
        check_payload(nodes, tri, np.asarray([0, n]), 4)
        rays = R.to_numpy_rays(R.uniform_rays([-3, -3, -3], [3, 3, 3], 20000, seed=n))
        got = scene.trace_host(rays)
        up = dev.upload_bvh(nodes, verts, [0])
        assert got.tobytes() == up.trace_host(rays).tobytes()
        assert (got["meshIndex"] != 0xFFFFFFFF).sum() > 0
        up.free()
        scene.free()
    verts = (rng.random((40, 3)) * 4 - 2).astype(np.float32)
    voff = np.asarray([0, 10, 10, 10, 25], dtype=np.uint32)             # meshes 1 and 2 are empty
    toff = np.asarray([0, 6, 6, 6, 13, 20], dtype=np.uint32)
    tri = np.concatenate([rng.integers(0, 10, (6, 3)), rng.integers(0, 15, (7, 3)), rng.integers(0, 15, (7, 3))]).astype(np.uint32)
    scene, tm, nodes = dev.build_scene(verts, voff, tri, toff, 2, 0, want_nodes=True, node_dtype=O.NODE_DTYPE)
    check_payload(nodes, tri, toff, 2)
    leaf = (nodes["nodeData"] >> 31) == 1
    assert set(nodes["w"][leaf, 3].tolist()) == {0, 3, 4}
    rays = R.to_numpy_rays(R.uniform_rays([-3, -3, -3], [3, 3, 3], 50000, seed=77))
    got = scene.trace_host(rays)
    up = dev.upload_bvh(nodes, verts, voff)
    assert got.tobytes() == up.trace_host(rays).tobytes()
    hit = got["meshIndex"] != 0xFFFFFFFF
    assert hit.sum() > 0 and set(np.unique(got["meshIndex"][hit]).tolist()) <= {0, 3, 4}
    up.free()
    scene.free()


def test_scene_build_rejects_bad_input(dev):
    verts = np.zeros((4, 3), np.float32)
    tri = np.asarray([[0, 1, 2], [1, 2, 9]], dtype=np.uint32)          # vertex 9 does not exist
    before = dev.counters().device_bytes_in_use
    with pytest.raises(capi.LrbError, match="vertex outside"):
        dev.build_scene(verts, [0], tri, [0, 2])
    with pytest.raises(capi.LrbError, match="tree type"):
        dev.build_scene(verts, [0], tri[:1], [0, 1], tree_type=3)
    assert dev.counters().device_bytes_in_use == before


@pytest.mark.parametrize("builder", ["EMBREE_MORTON", "B200_PLOC"])
def test_host_layer_resident_scene_equals_the_two_step_path(builder):
    """BVHAccel::Init with a GPU builder keeps the scene on the device and BVHKernel adopts it; with
    accelerator.b200.resident = 0 the leaf boxes go up, the array comes down and lrb_bvh_upload lays it out on the host.
    Same array (=> the device's build boxes are BVHAccel::Init's), same hits, same memory accounting."""
    desc = S.load_fixture("kitchen")
    cfg = {"accelerator.type": "BVH", "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}
    a = hostapi.Session(dict(cfg), desc)
    a.start(0)
    b = hostapi.Session(dict(cfg, **{"accelerator.b200.resident": 0}), desc)
    b.start(0)
    na, nb = a.bvh_nodes(), b.bvh_nodes()
    assert na.tobytes() == nb.tobytes()
    rays = _rays(desc, 300000, 41)
    ga, gb = a.trace_host(rays), b.trace_host(rays)
    assert ga.tobytes() == gb.tobytes()
    ref = O.BVH(H.oracle_scene(desc), nodes=na).intersect(rays)
    rep = H.compare_hits(ga, ref, rays, what="resident/" + builder)
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    ia, ib = a.native_scene().info(), b.native_scene().info()
    assert (ia.n_wide_nodes, ia.n_triangles, ia.stack_need, ia.device_bytes) == (ib.n_wide_nodes, ib.n_triangles, ib.stack_need, ib.device_bytes)
    # a restarted device has no resident scene left to adopt: it uploads the array, with the same result
    a.stop()
    a.start(0)
    assert a.trace_host(rays).tobytes() == ga.tobytes()
    for s in (a, b):
        s.stop()
        s.close()


def test_resident_scene_of_a_flattened_instance_scene():
    """accelerator.instances.enable = 0: the BVH is built over world-space copies of the instances (dataset.h:43); the device
    build gets those vertices (Mesh::GetVertex) and the base meshes' triangle indices.  Same array and hits as the two-step
    path, and the oracle's flattened BVH walking that array agrees bit for bit."""
    import scene_zoo as Z
    desc = Z.instances_scene(10)
    cfg = {"accelerator.instances.enable": 0, "accelerator.bvh.builder.type": "B200_PLOC"}
    a = hostapi.Session(dict(cfg), desc)
    a.start(0)
    b = hostapi.Session(dict(cfg, **{"accelerator.b200.resident": 0}), desc)
    b.start(0)
    assert a.accelerator_type() == hostapi.ACCEL_BVH and b.accelerator_type() == hostapi.ACCEL_BVH
    assert a.bvh_nodes().tobytes() == b.bvh_nodes().tobytes()
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    rays = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, 200000, seed=97))
    ga, gb = a.trace_host(rays), b.trace_host(rays)
    assert ga.tobytes() == gb.tobytes()
    ref = O.BVH(H.oracle_scene(desc), nodes=a.bvh_nodes()).intersect(rays)
    rep = H.compare_hits(ga, ref, rays, what="resident/flattened")
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    for s in (a, b):
        s.stop()
        s.close()
