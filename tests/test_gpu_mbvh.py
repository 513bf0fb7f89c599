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


def _reference_mbvh(desc):
    from oracle import refapi
    if not refapi.available():
        pytest.skip("oracle/_ref (the reference compiled from /root/reference) is not present on this machine")
    return refapi.MBVH(H.reference_scene(desc))


def test_full_lightinstances_against_the_reference_library(dev):
    """BASELINE.json configs[2] at full size -- 4 502 objects, 4 500 instances -- against the REFERENCE's own
    MBVHAccel::Intersect (oracle/_ref), not the oracle port: bit for bit (camera, uniform and bounce-like surface rays)."""
    desc = S.load_fixture("lightinstances")
    assert len(desc.meshes) == 4502
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    scene = _upload(dev, desc, mb)
    assert scene.info().n_instances == 4502
    ref_accel = _reference_mbvh(desc)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = np.concatenate([_rays(desc, 400000, 91), R.to_numpy_rays(R.surface_rays(p0, e1, e2, 200000, seed=93, axis_fraction=0.2))])
    ref = ref_accel.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="full lightinstances vs reference library", two_level=True)
    assert rep["hits"] > 0.2 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()


@pytest.mark.parametrize("which", ["zoo-inst", "bigmonkey-instances"])
def test_instances_against_the_reference_library(dev, which):
    desc = Z.instances_scene() if which == "zoo-inst" else S.load_fixture(which)
    osc = H.oracle_scene(desc)
    scene = _upload(dev, desc, O.MBVH(osc))
    ref_accel = _reference_mbvh(desc)
    rays = _rays(desc, 300000, 95)
    rep = H.compare_hits(scene.trace_host(rays), ref_accel.intersect(rays), rays, what=which + " vs reference library")
    assert rep["hits"] > 0.1 * rep["n"] and rep["bit_exact_hits"] == rep["hits"]
    scene.free()


@pytest.mark.parametrize("which,n", [("zoo-motion", 300000), ("bigmonkey-motion", 300000)])
def test_motion_against_the_reference_library(dev, which, n):
    """Rotating (zoo-motion: Slerp) and translating (bigmonkey-motion) motion blur against the REFERENCE's own
    MotionSystem::Sample + MBVHAccel::Intersect.  The reference's Slerp calls the host libm's sinf / acosf; the device
    evaluates them in double and rounds once, which reproduces glibc in all but a small fraction of the calls -- the
    observed fraction of hits that are not bit-identical is printed and bounded, none may be off by more than 1e-3."""
    desc = Z.motion_scene() if which == "zoo-motion" else S.load_fixture(which)
    osc = H.oracle_scene(desc)
    scene = _upload(dev, desc, O.MBVH(osc))
    ref_accel = _reference_mbvh(desc)
    rays = _rays(desc, n, 97, time_range=(-0.05, 1.05))
    ref = ref_accel.intersect(rays)
    got = scene.trace_host(rays)
    _, second = osc.brute(rays, two_level=True, want_second=True)
    rep = H.compare_hits(got, ref, rays, second_t=second, rel_tol=1e-5, what=which + " vs reference library", libm_outlier_frac=1e-4)
    frac_inexact = 1.0 - rep["bit_exact_hits"] / max(1, rep["hits"])
    print("%s: %d hits, %.4f %% not bit-identical to the reference library, %d beyond 1e-5 (libm outliers), %d tie-exempt"
          % (which, rep["hits"], 100.0 * frac_inexact, rep["libm_outliers"], rep["tie_exempt"]))
    assert rep["hits"] > 0.1 * rep["n"]
    assert rep["tie_exempt"] <= 1e-4 * rep["n"]
    assert frac_inexact <= 0.005
    if which == "bigmonkey-motion":
        assert rep["libm_outliers"] == 0       # translation only: no trigonometry involved
    scene.free()
