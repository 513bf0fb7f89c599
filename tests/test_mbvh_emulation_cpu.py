"""Two-level (MBVH) parity on the CPU: product re-layout + traversal body vs the oracle's
MBVHAccel::Intersect restatement, on instanced and motion-blurred scenes."""
import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import rays as R, scenes as S
from oracle import oracle as O


def _rays(desc, n, seed, time_range=None):
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed, time_range=time_range))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1, time_range=time_range))
    return np.concatenate([a, b])


def _check(desc, tree_type, n, time_range=None, motion=False):
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=tree_type)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    rays = _rays(desc, n, 31, time_range)
    ref, cnt = mb.intersect(rays, count=True)
    got, st = emu.trace(rays, want_stats=True)
    # instances: identical arithmetic -> bit-exact; motion: sinf/acosf may differ between libm and
    # CUDA on the GPU, but the CPU emulation uses the same libm as the oracle
    rep = H.compare_hits(got, ref, rays, what="%s k=%d" % (desc.name, tree_type))
    assert rep["hits"] > 0.1 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    assert st["max_stack"] <= emu.info()["stack_need"]
    if motion:
        assert st["motion_samples"] > 0
    # topology-free pin of the oracle itself
    brute, second = osc.brute(rays[:3000], two_level=True, want_second=True)
    with np.errstate(invalid="ignore"):
        tie = np.abs(second - brute["t"]) <= 1e-5 * np.maximum(1.0, np.abs(brute["t"]))
    r3 = ref[:3000]
    same = (r3["meshIndex"] == brute["meshIndex"]) & ((r3["triangleIndex"] == brute["triangleIndex"]) | (brute["meshIndex"] == H.NULL))
    assert (same | tie).all()
    return rep


@pytest.mark.parametrize("tree_type", [2, 4, 8])
def test_instances_zoo(tree_type):
    _check(Z.instances_scene(), tree_type, 15000)


def test_motion_zoo():
    _check(Z.motion_scene(), 4, 15000, time_range=(-0.1, 1.1), motion=True)


def test_bigmonkey_instances():
    _check(S.load_fixture("bigmonkey-instances"), 4, 20000)


def test_bigmonkey_motion():
    _check(S.load_fixture("bigmonkey-motion"), 4, 20000, time_range=(0.0, 1.0), motion=True)


def test_lightinstances_subset():
    _check(S.load_fixture("lightinstances", max_objects=300), 4, 20000)


def test_mbvh_update_root_only():
    desc = Z.instances_scene(12)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    rays = _rays(desc, 8000, 41)
    H.compare_hits(emu.trace(rays), mb.intersect(rays), rays, what="before update")
    # move two instances (scene edit), Update() rebuilds only the root tree
    inst = [i for i, m in enumerate(desc.meshes) if m.kind == S.INSTANCE]
    for k, i in enumerate(inst[:2]):
        m = Z.translate(1.5 * (k + 1), -2.0, 0.5) @ Z.rot_z(33.0 * (k + 1))
        osc.set_instance_transform(i, m)
    mb.update()
    emu.update(mb.root_nodes(), mb.transforms_minv())
    ref = mb.intersect(rays)
    rep = H.compare_hits(emu.trace(rays), ref, rays, what="after update")
    assert rep["bit_exact_hits"] == rep["hits"]


@pytest.mark.parametrize("which", ["zoo-inst", "bigmonkey-instances", "lightinstances"])
def test_instances_grazing_rays(which):
    """Two-level scenes with rays that start on (or a few epsilons off) the instanced surfaces, a
    third of them axis-parallel: hits at the planes of the quantized grids of leaf trees (in instance
    space) and of the root tree (whole-grid instance slots)."""
    desc = {"zoo-inst": lambda: Z.instances_scene(), "bigmonkey-instances": lambda: S.load_fixture("bigmonkey-instances"),
            "lightinstances": lambda: S.load_fixture("lightinstances", max_objects=300)}[which]()
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = R.to_numpy_rays(R.surface_rays(p0, e1, e2, 40000, seed=71, axis_fraction=0.33))
    ref = mb.intersect(rays)
    got = emu.trace(rays)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="grazing " + which, two_level=True)
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]


@pytest.mark.parametrize("kind", ["instance", "motion", "plain"])
def test_single_mesh_dataset_root_is_a_leaf(kind):
    """A dataset with ONE mesh: the MBVH root tree is a single leaf without any box (the node takes an
    unbounded grid), for an instanced, a motion-blurred and a plain mesh."""
    s = S.SceneDesc("single-" + kind)
    shape = s.add_shape(*Z.blob(12, 9))
    if kind == "instance":
        s.add_instance(shape, Z.translate(0.5, -0.25, 0.75) @ Z.rot_z(25.0))
    elif kind == "motion":
        s.add_motion(shape, [0.0, 1.0], [np.linalg.inv(Z.translate(0, 0, 0)).astype(np.float32), np.linalg.inv(Z.translate(1.0, 0.5, 0.0)).astype(np.float32)])
    else:
        s.add_plain(shape)
    s.cam = np.asarray([0.5, -6.0, 0.75, 0.5, 0.0, 0.75, 0, 0, 1, 45], dtype=np.float32)
    tr = (0.0, 1.0) if kind == "motion" else None
    _check(s, 4, 12000, time_range=tr, motion=(kind == "motion"))


@pytest.mark.parametrize("fixture,kw", [("bigmonkey-instances", {}), ("bigmonkey-motion", {"time_range": (0.0, 1.0)}),
                                        ("lightinstances", {"max_objects": 300})])
def test_host_layer_sah_trees_two_level(fixture, kw):
    """The product host layer's MBVHAccel::Init with the default builder (binary SAH -> optimisation ->
    k-ary collapse, bvhbuild.cpp) emits root and leaf trees of its own topology; payloads (leaf /
    transform / motion / mesh-offset indices) follow mbvhaccel.cpp:58-250 like the oracle's, so its
    arrays can be dropped into the oracle's array set.  The re-layout + traversal body over THOSE trees
    must reproduce the oracle's MBVHAccel::Intersect bit for bit (results do not depend on topology)."""
    from luxcore_b200 import hostapi
    time_range = kw.pop("time_range", None)
    desc = S.load_fixture(fixture, **kw)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    arr = H.mbvh_arrays(desc, mb)
    s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    assert s.build_accelerator("MBVH") == hostapi.ACCEL_MBVH
    assert s.mbvh_leaf_count() == len(arr["leaf_nodes"])
    root = s.mbvh_root_nodes().copy()
    # same multiset of root-leaf payloads as the oracle's root tree
    def payloads(nodes):
        leaf = (nodes["nodeData"] & 0x80000000) != 0
        return sorted(map(tuple, nodes["w"][leaf][:, :4].tolist()))
    assert payloads(root) == payloads(arr["root_nodes"])
    for i in range(s.mbvh_leaf_count()):
        ln = s.mbvh_leaf_nodes(i).copy()
        assert int(((ln["nodeData"] & 0x80000000) != 0).sum()) == int(((arr["leaf_nodes"][i]["nodeData"] & 0x80000000) != 0).sum())
        assert H.Emu.lib().emu_validate_tree(ln.ctypes.data, ln.shape[0]) == 0
        arr["leaf_nodes"][i] = ln
    arr["root_nodes"] = root
    emu = H.Emu.mbvh(arr)
    rays = _rays(desc, 20000, 47, time_range)
    ref = mb.intersect(rays)
    got, st = emu.trace(rays, want_stats=True)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="host sah mbvh " + fixture, two_level=True)
    assert rep["hits"] > 0.1 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    assert st["max_stack"] <= emu.info()["stack_need"]


@pytest.mark.parametrize("inst_bias", [0, 2, 8, 64])
def test_two_level_scheduling_model_matches_per_ray_traversal(inst_bias):
    """The warp-level replay of the persistent kernel's loop (three-way vote: node / triangle phase, entering
    instances) must give the per-ray traversal's hits and do the same traversal work for any vote setting:
    lanes that wait for an instance phase must neither be dropped nor enter twice."""
    desc = S.load_fixture("lightinstances", max_objects=300)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    rays = _rays(desc, 20000, 53)
    rays["flags"][::11] = 1
    ref, st = emu.trace(rays, want_stats=True)
    for n_warps, refill in ((1, 24), (32, 24), (8, 1)):
        hits, c = H.warp_sim(emu, rays, n_warps=n_warps, refill_below=refill, tri_bias=8, inst_bias=inst_bias)
        assert hits.tobytes() == ref.tobytes()
        assert c["rays"] == st["rays"] and c["node_lanes"] == st["wide_nodes"] and c["tri_lanes"] == st["triangles"]
        assert c["instance_lanes"] == st["instances"]
    live = rays["flags"] == 0           # masked rays are a device-side notion (bvh.cl:242-244): RayHit untouched
    got = H.compare_hits(ref[live], mb.intersect(rays[live]), rays[live], what="two-level model")
    assert got["bit_exact_hits"] == got["hits"] > 0


def test_motion_bounds_cover_large_rotations():
    """Instance slots of motion-blurred meshes hold world-space bounds over all times, sampled at a few
    times per segment and grown by the inter-sample displacement (relayout.cpp MotionWorldBox).  Stress:
    objects far from the rotation centre swinging through up to 175 degrees in ONE segment, two of them
    sharing a root node with each other -- every hit the oracle finds at any time must still be found."""
    s = S.SceneDesc("orbit")
    a = s.add_shape(*Z.blob(10, 9, radius=1.2))
    for k, (deg, r) in enumerate(((175.0, 6.0), (120.0, 3.0), (-160.0, 9.0), (60.0, 1.5))):
        m0 = Z.rot_z(20.0 * k) @ Z.translate(r, 0, 0.3 * k)
        m1 = Z.rot_z(20.0 * k + deg) @ Z.translate(r, 0, 0.3 * k) @ Z.rot_x(35.0)
        s.add_motion(a, [0.0, 1.0], [Z.inv(m0), Z.inv(m1)])
    s.add_plain(s.add_shape(*S.grid_mesh(4, 4, z=-1.0, size=12.0)))
    s.cam = np.asarray([0, -20, 8, 0, 0, 0, 0, 0, 1, 60], dtype=np.float32)
    osc = H.oracle_scene(s)
    mb = O.MBVH(osc, tree_type=4)
    emu = H.Emu.mbvh(H.mbvh_arrays(s, mb))
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays([-10, -10, -1], [10, 10, 2], 60000, seed=8, time_range=(-0.05, 1.05))),
                           R.to_numpy_rays(R.camera_rays(s.cam, 200, 200, seed=9, time_range=(0.0, 1.0)))])
    ref = mb.intersect(rays)
    got, st = emu.trace(rays, want_stats=True)
    rep = H.compare_hits(got, ref, rays, what="orbit")
    moving = (ref["meshIndex"] < 4)
    assert int(moving.sum()) > 2000, int(moving.sum())      # the swinging objects are actually hit, at all times
    t_hit = rays["time"][moving]
    assert t_hit.min() < 0.05 and t_hit.max() > 0.95 and ((t_hit > 0.4) & (t_hit < 0.6)).any()
    assert rep["bit_exact_hits"] == rep["hits"]
    # and the bounds do cull: far fewer instance entries than "every motion instance of every visited root node"
    assert st["instances"] / st["rays"] < 2.0


def test_nan_mint_with_a_lone_instance():
    """Regression (tools/fuzz_parity.py): a dataset of ONE instance has a root tree that is a single leaf; its
    slot covers the whole grid, i.e. all three slabs decode to NaN.  With ray.mint = NaN as well the entry
    distance was NaN and read as "missed", while the reference (NaN mint never rejects) enters the instance."""
    desc = S.SceneDesc("lone")
    desc.add_instance(desc.add_shape(*Z.blob(12, 4)), Z.translate(30, -20, 50) @ Z.rot_z(25) @ Z.scale(20, 20, 20))
    desc.cam = np.asarray([30, -120, 50, 30, -20, 50, 0, 0, 1, 40], dtype=np.float32)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = np.concatenate([R.to_numpy_rays(R.surface_rays(p0, e1, e2, 4000, seed=3)), R.to_numpy_rays(R.camera_rays(desc.cam, 50, 50, seed=2))])
    for mint in (np.nan, 0.0, -np.inf, -3.0):
        r = rays.copy()
        r["mint"] = mint
        rep = H.compare_hits_tie_aware(emu.trace(r), mb.intersect(r), r, osc, what="lone instance, mint %r" % mint, two_level=True)
        assert rep["hits"] > 2000 and rep["bit_exact_hits"] == rep["hits"]
        got = H.Lockstep.trace(emu, r[:600])
        assert got.tobytes() == emu.trace(r[:600]).tobytes()
