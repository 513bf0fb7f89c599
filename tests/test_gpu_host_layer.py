"""The drop-in sequence through the C++ host layer (luxrays::Context / DataSet / CUDAIntersectionDevice,
SURVEY.md 3.5) on the GPU, checked against the oracle walking the very trees the host layer built."""
import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _rays(desc, n, seed, time_range=None):
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed, time_range=time_range))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1, time_range=time_range))
    return np.concatenate([a, b])


@pytest.mark.parametrize("builder", ["CLASSIC", "EMBREE_BINNED_SAH"])
@pytest.mark.parametrize("name,n", [("cornell", 100000), ("kitchen", 600000)])
def test_bvh_drop_in_sequence(name, n, builder):
    desc = S.load_fixture(name)
    s = hostapi.Session({"accelerator.bvh.builder.type": builder}, desc)
    s.start(0)
    assert s.accelerator_type() == hostapi.ACCEL_BVH        # AUTO -> BVH without instances / motion
    rays = _rays(desc, n, 81)
    got = s.trace_host(rays)                                 # AllocBufferRW / Enqueue / Read / Finish
    ref = O.BVH(H.oracle_scene(desc), nodes=s.bvh_nodes()).intersect(rays)
    rep = H.compare_hits(got, ref, rays, what="host/%s/%s" % (name, builder))
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    assert s.total_rays() == rays.shape[0]
    assert s.used_memory() == rays.shape[0] * (48 + 20)
    # serial interface: one ray, traced on the GPU
    for i in (0, 17, 4242):
        hit, h = s.trace_ray(rays[i])
        assert hit == (ref[i]["meshIndex"] != H.NULL)
        if hit:
            assert h["t"] == ref[i]["t"] and h["triangleIndex"] == ref[i]["triangleIndex"]
    s.stop()
    s.close()


def test_mbvh_auto_selection_update_and_masked():
    desc = Z.instances_scene(20)
    s = hostapi.Session({"accelerator.bvh.builder.type": "CLASSIC"}, desc)
    s.start(0)
    assert s.accelerator_type() == hostapi.ACCEL_MBVH       # AUTO -> MBVH because of instances
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    # the host layer's CLASSIC trees are the oracle's trees
    assert s.mbvh_root_nodes().tobytes() == mb.root_nodes().tobytes()
    assert s.mbvh_leaf_count() == mb.leaf_count()
    for i in range(mb.leaf_count()):
        assert s.mbvh_leaf_nodes(i).tobytes() == mb.leaf_nodes(i).tobytes()
    rays = _rays(desc, 150000, 91)
    rep = H.compare_hits(s.trace_host(rays), mb.intersect(rays), rays, what="host/mbvh")
    assert rep["bit_exact_hits"] == rep["hits"]

    # scene edit through the plugin surface: SetTransformation + Context::UpdateDataSet
    inst = [i for i, m in enumerate(desc.meshes) if m.kind == S.INSTANCE][:3]
    for k, i in enumerate(inst):
        m = Z.translate(2.0 - k, 1.0 + k, 0.5) @ Z.rot_x(25.0 * (k + 1))
        s.set_instance_transform(i, m)
        osc.set_instance_transform(i, m)
    s.update()
    mb.update()
    assert s.mbvh_root_nodes().tobytes() == mb.root_nodes().tobytes()
    rep = H.compare_hits(s.trace_host(rays), mb.intersect(rays), rays, what="host/mbvh/updated")
    assert rep["bit_exact_hits"] == rep["hits"]

    # masked rays keep the caller's RayHit (hits uploaded with AllocBufferRW(src = hostHits))
    mask = np.random.default_rng(3).random(rays.shape[0]) < 0.5
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    hits = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
    hits["t"] = -5.0
    hits["meshIndex"] = 123
    s.trace_host(rays, hits)
    assert (hits["t"][mask] == -5.0).all() and (hits["meshIndex"][mask] == 123).all()
    H.compare_hits(hits[~mask], mb.intersect(rays)[~mask], what="host/mbvh/masked")
    s.stop()
    s.close()


def test_motion_scene_sah_trees():
    desc = S.load_fixture("bigmonkey-motion")
    s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH"}, desc)
    s.start(0)
    assert s.accelerator_type() == hostapi.ACCEL_MBVH
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    mb.set_root_nodes(s.mbvh_root_nodes())                  # oracle walks the product's SAH trees
    for i in range(s.mbvh_leaf_count()):
        mb.set_leaf_nodes(i, s.mbvh_leaf_nodes(i))
    rays = _rays(desc, 200000, 95, time_range=(0.0, 1.0))
    _, second = osc.brute(rays, two_level=True, want_second=True)
    rep = H.compare_hits(s.trace_host(rays), mb.intersect(rays), rays, second_t=second, what="host/motion/sah", libm_outlier_frac=1e-4)
    assert rep["hits"] > 0 and rep["tie_exempt"] <= 1e-4 * rep["n"]
    s.stop()
    s.close()


def test_instances_disabled_flattens_to_bvh():
    desc = Z.instances_scene(10)
    s = hostapi.Session({"accelerator.instances.enable": 0, "accelerator.bvh.builder.type": "CLASSIC"}, desc)
    s.start(0)
    assert s.accelerator_type() == hostapi.ACCEL_BVH        # world-space flattened copy (dataset.h:43)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    assert s.bvh_nodes().tobytes() == bvh.nodes().tobytes()
    rays = _rays(desc, 100000, 97)
    rep = H.compare_hits(s.trace_host(rays), bvh.intersect(rays), rays, what="host/flattened")
    assert rep["bit_exact_hits"] == rep["hits"]
    s.stop()
    s.close()


def test_error_behaviour():
    desc = S.load_fixture("cornell")
    s = hostapi.Session({}, desc)
    with pytest.raises(hostapi.HostError):
        s.trace_host(np.zeros(4, dtype=capi.RAY_DTYPE))      # not started
    with pytest.raises(hostapi.HostError):
        hostapi.Session({"accelerator.type": "NOPE"})
    s2 = hostapi.Session({"accelerator.bvh.builder.type": "WHATEVER"}, desc)
    with pytest.raises(hostapi.HostError):
        s2.build_accelerator("BVH")
    s.close()
    s2.close()
