"""GPU parity for the two-level path (instances, motion blur, Update) through the C ABI."""
import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import capi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


def _rays(desc, n, seed, time_range=None):
    lo, hi = desc.bbox()
    pad = 0.1 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed, time_range=time_range))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1, time_range=time_range))
    return np.concatenate([a, b])


def _upload(dev, desc, mb):
    a = H.mbvh_arrays(desc, mb)
    return dev.upload_mbvh(a["root_nodes"], a["leaf_nodes"], a["leaf_verts"], a["transforms_minv"], a["motion_table"], a["interps"])


@pytest.mark.parametrize("kernel", ["persistent", "simple"])
@pytest.mark.parametrize("which,tree_type,n", [("zoo-inst", 4, 200000), ("zoo-inst", 2, 100000), ("zoo-inst", 8, 100000),
                                               ("bigmonkey-instances", 4, 300000), ("lightinstances", 4, 400000)])
def test_instances_bit_exact(dev, kernel, which, tree_type, n):
    dev.set_option("kernel", kernel)
    desc = Z.instances_scene() if which == "zoo-inst" else S.load_fixture(which, max_objects=1500 if which == "lightinstances" else None)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=tree_type)
    scene = _upload(dev, desc, mb)
    assert scene.info().two_level == 1
    rays = _rays(desc, n, 51)
    ref = mb.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits(got, ref, rays, what="%s k=%d %s" % (which, tree_type, kernel))
    assert rep["hits"] > 0.1 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]     # same operations in the same order -> same floats
    scene.free()
    dev.set_option("kernel", "persistent")


@pytest.mark.parametrize("which,n", [("zoo-motion", 200000), ("bigmonkey-motion", 300000)])
def test_motion_within_tolerance(dev, which, n):
    desc = Z.motion_scene() if which == "zoo-motion" else S.load_fixture(which)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    scene = _upload(dev, desc, mb)
    rays = _rays(desc, n, 61, time_range=(-0.05, 1.05))
    ref = mb.intersect(rays)
    got = scene.trace_host(rays)
    # CUDA sinf/acosf vs libm: t may move by an ulp, so index flips are tolerated only on near-ties
    _, second = osc.brute(rays, two_level=True, want_second=True)
    rep = H.compare_hits(got, ref, rays, second_t=second, rel_tol=1e-5, what=which, libm_outlier_frac=1e-4)
    assert rep["hits"] > 0.1 * rep["n"]
    assert rep["tie_exempt"] <= 1e-4 * rep["n"]
    assert rep["bit_exact_hits"] >= 0.995 * rep["hits"]     # double-rounded sin/acos reproduce glibc almost always
    scene.free()


def test_update_root_and_transforms(dev):
    desc = Z.instances_scene(16)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    scene = _upload(dev, desc, mb)
    rays = _rays(desc, 100000, 71)
    H.compare_hits(scene.trace_host(rays), mb.intersect(rays), rays, what="before update")
    inst = [i for i, m in enumerate(desc.meshes) if m.kind == S.INSTANCE]
    for step in range(3):
        for k, i in enumerate(inst[:4]):
            osc.set_instance_transform(i, Z.translate(0.7 * step + k, -1.0 * k, 0.3 * step) @ Z.rot_z(20.0 * step + 11 * k))
        mb.update()
        scene.update(mb.root_nodes(), mb.transforms_minv())
        rep = H.compare_hits(scene.trace_host(rays), mb.intersect(rays), rays, what="after update %d" % step)
        assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()


@pytest.mark.parametrize("which", ["zoo-inst", "bigmonkey-instances", "lightinstances"])
def test_instances_grazing_rays(dev, which):
    """Rays starting on / a few epsilons off the instanced surfaces, a third of them axis-parallel."""
    desc = {"zoo-inst": lambda: Z.instances_scene(), "bigmonkey-instances": lambda: S.load_fixture("bigmonkey-instances"),
            "lightinstances": lambda: S.load_fixture("lightinstances", max_objects=300)}[which]()
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    scene = _upload(dev, desc, mb)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = R.to_numpy_rays(R.surface_rays(p0, e1, e2, 200000, seed=73, axis_fraction=0.33))
    ref = mb.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="grazing " + which, two_level=True)
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()
