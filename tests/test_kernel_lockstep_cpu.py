"""The REAL kernel source on the CPU.  luxcore_b200/csrc/trace_kernels.cuh is compiled by g++ against a
stand-in <cuda_runtime.h> (tests/cpp/fakecuda) and executed with one OS thread per lane; warp collectives
are rendezvous of a warp's 32 threads (tests/cpp/kernel_lockstep.cpp).  This covers what the per-ray
emulation cannot see: TracePersistent's own loop (bulk re-fill by ballot + popc prefix, masked rays, deferred
RayHit stores, Resolve / instance entry / phase votes as written in the kernel), the shared-memory stack and
its global spill, TraceStatic's local stack and counters, the prefetching twin.  The arithmetic is the host
variant of traverse.h, the same as the emulation's, so hits must be bit-identical to the emulation's --
which the other CPU tests pin to the oracle and the reference."""
import numpy as np
import pytest

import helpers as H
import scene_zoo as Z
from luxcore_b200 import rays as R, scenes as S
from oracle import oracle as O


# A kernel loop that fails to terminate would block its 32 lane threads for ever: give up after 15 minutes
# (the whole file takes about one).  method="thread": the main thread sits inside the C call, a signal cannot fire.
pytestmark = pytest.mark.timeout(900, method="thread")


def _preloaded(n, seed):
    """A RayHit buffer with recognisable garbage: masked rays must leave it untouched."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 255, size=n * 20, dtype=np.uint8).view(H.HIT_DTYPE).copy()


def _expect(emu, rays, pre):
    ref = emu.trace(rays)
    masked = (rays["flags"] & 1) != 0
    ref[masked] = pre[masked]
    return ref


@pytest.fixture(scope="module")
def kitchen():
    desc = S.load_fixture("kitchen")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    lo, hi = desc.bbox()
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo, hi, 2500, seed=5)), R.to_numpy_rays(R.camera_rays(desc.cam, 30, 30, seed=6))])
    rays["flags"][::9] = 1
    return emu, rays


@pytest.mark.parametrize("kw", [dict(), dict(n_warps=4), dict(smem_depth=4), dict(smem_depth=64), dict(prefetch=True),
                                dict(refill_below=1, tri_bias=64), dict(refill_below=32, tri_bias=1), dict(kernel="static", n_warps=4)],
                         ids=lambda kw: ",".join("%s=%s" % kv for kv in kw.items()) or "default")
def test_bvh_kernels_in_lockstep(kitchen, kw):
    emu, rays = kitchen
    assert emu.info()["stack_need"] > 16        # the default shared depth spills on this scene
    pre = _preloaded(rays.shape[0], 3)
    got = H.Lockstep.trace(emu, rays, hits=pre, **kw)
    assert got.tobytes() == _expect(emu, rays, pre).tobytes()


def test_static_kernel_counters_match_the_emulation(kitchen):
    emu, rays = kitchen
    ref, st = emu.trace(rays, want_stats=True)
    got, ks = H.Lockstep.trace(emu, rays, kernel="static", n_warps=4, want_stats=True)
    assert {k: ks[k] for k in ("rays", "wide_nodes", "triangles")} == {k: st[k] for k in ("rays", "wide_nodes", "triangles")}
    assert 0 < ks["max_stack"] <= emu.info()["stack_need"]


@pytest.mark.parametrize("which", ["zoo-instances", "zoo-motion", "lightinstances", "single-instance"])
def test_two_level_kernels_in_lockstep(which):
    if which == "zoo-instances":
        desc = Z.instances_scene()
    elif which == "zoo-motion":
        desc = Z.motion_scene()
    elif which == "lightinstances":
        desc = S.load_fixture("lightinstances", max_objects=300)
    else:
        desc = S.SceneDesc("one")
        desc.add_instance(desc.add_shape(*Z.blob(8, 2)), Z.translate(0.5, 0, 0) @ Z.rot_z(30))
        desc.cam = np.asarray([0, -5, 1, 0, 0, 0, 0, 0, 1, 50], dtype=np.float32)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    lo, hi = desc.bbox()
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - 0.2, hi + 0.2, 1500, seed=5, time_range=(-0.1, 1.1))),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 36, 36, seed=6, time_range=(0.0, 1.0)))])
    rays["flags"][::13] = 1
    pre = _preloaded(rays.shape[0], 4)
    want = _expect(emu, rays, pre)
    assert ((want["meshIndex"] != H.NULL) & ((rays["flags"] & 1) == 0)).sum() > 100
    for kw in (dict(inst_bias=0), dict(inst_bias=8), dict(inst_bias=64, n_warps=4), dict(smem_depth=4, n_warps=2),
               dict(refill_below=1), dict(kernel="static", n_warps=4)):
        got = H.Lockstep.trace(emu, rays, hits=pre, **kw)
        assert got.tobytes() == want.tobytes(), kw
    # and the emulation these are compared with is the oracle's answer
    live = (rays["flags"] & 1) == 0
    rep = H.compare_hits(want[live], mb.intersect(rays[live]), rays[live], what=which)
    assert rep["bit_exact_hits"] == rep["hits"]


def test_empty_scene_and_tiny_batches():
    desc = S.load_fixture("cornell")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    rays = R.to_numpy_rays(R.camera_rays(desc.cam, 16, 16, seed=1))
    for n in (1, 2, 31, 32, 33, 100):
        got = H.Lockstep.trace(emu, rays[:n], n_warps=2)
        assert got.tobytes() == emu.trace(rays[:n]).tobytes(), n
    empty = H.Emu.bvh(np.zeros(0, dtype=bvh.nodes().dtype), np.zeros((0, 3), np.float32), np.zeros(1, np.uint32))
    got = H.Lockstep.trace(empty, rays)
    assert (got["meshIndex"] == H.NULL).all() and (got["t"] == rays["maxt"]).all()


@pytest.mark.parametrize("n_warps,chunk_shift", [(4, 5), (4, 7), (4, 10)])
def test_signalled_gather_kernel_in_lockstep(kitchen, n_warps, chunk_shift):
    """TracePersistent<*, *, SIGNAL>: warp 0 watches the other warps' watermarks and raises a chunk's flag when
    every ray index below the chunk's end is finished; the copy stream (here: a watcher thread) may then read
    the chunk.  Hits (local and peer copy) must equal the plain kernel's, every flag raised during the run must
    have been raised only after all of its chunk's records were final, and flags must not be raised early."""
    emu, rays = kitchen
    pre = _preloaded(rays.shape[0], 6)
    want = _expect(emu, rays, pre)
    got, peer, flags, raised, order_ok = H.Lockstep.trace_signal(emu, rays, n_warps=n_warps, chunk_shift=chunk_shift, epoch=9, hits=pre)
    assert got.tobytes() == want.tobytes()
    assert peer.tobytes() == want.tobytes()         # dual stores + forwarded masked records
    assert order_ok
    assert set(np.unique(flags).tolist()) <= {0, 9}
    assert raised == int((flags == 9).sum()) == len(flags)     # the detector itself raised every chunk (on the GPU the host's memset is only the safety net)


@pytest.mark.parametrize("tree_type", [2, 4, 8])
def test_non_finite_rays_stay_within_the_stack_bound(tree_type):
    """Regression (tools/fuzz_parity.py --lockstep under AddressSanitizer): for a ray with a NaN origin or
    direction every comparison of the slab test is false, so EVERY slot of every visited node passes -- the
    unused ones too -- and the stack grows by three entries per level whatever the arity of the tree.  The
    worst-case depth computed at upload (it sizes the kernels' spill buffers and selects the non-spilling
    kernel) counted the used slots only: 17 against 29 reached.  Such rays must stay inside the bound, and
    the real kernel source must survive them at any shared-memory depth (results: whatever the emulation says;
    they are outside the parity contract)."""
    desc = S.load_fixture("bigmonkey")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo, hi, 40, seed=3))
    rays["o"][0::4] = np.nan            # all three slabs NaN: nothing is ever culled
    rays["d"][1::4, 2] = np.nan
    rays["d"][2::4, 1] = np.inf
    rays["mint"][3::4] = np.nan
    rays["maxt"][3::8] = np.nan
    got, st = emu.trace(rays, want_stats=True)
    need = emu.info()["stack_need"]
    assert st["max_stack"] <= need, (st["max_stack"], need)
    assert st["max_stack"] > need // 4          # the all-NaN rays do go deep (continuation nodes of 8-ary trees add slack)
    for depth in (1, 16):
        assert H.Lockstep.trace(emu, rays, smem_depth=depth, n_warps=2).tobytes() == got.tobytes()
    assert H.Lockstep.trace(emu, rays, kernel="static", n_warps=4).tobytes() == got.tobytes()


@pytest.mark.parametrize("kw", [dict(), dict(n_warps=4, smem_depth=4), dict(kernel="static", n_warps=4)],
                         ids=lambda kw: ",".join("%s=%s" % kv for kv in kw.items()) or "default")
def test_anyhit_kernels_in_lockstep(kitchen, kw):
    """Shadow-ray kernels (ANYHIT): same hit / miss as the closest-hit traversal, every reported hit is a real one
    (the reference's own triangle arithmetic, bit for bit), masked rays untouched."""
    emu, rays = kitchen
    desc = S.load_fixture("kitchen")
    pre = _preloaded(rays.shape[0], 5)
    got = H.Lockstep.trace(emu, rays, hits=pre, anyhit=True, **kw)
    masked = (rays["flags"] & 1) != 0
    assert got[masked].tobytes() == pre[masked].tobytes()
    rep = H.check_anyhit(got[~masked], emu.trace(rays)[~masked], rays[~masked], desc, what="lockstep any-hit")
    assert rep["hits"] > 500


def test_anyhit_two_level_in_lockstep():
    desc = Z.instances_scene()
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    emu = H.Emu.mbvh(H.mbvh_arrays(desc, mb))
    lo, hi = desc.bbox()
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - 0.2, hi + 0.2, 1500, seed=5)),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 36, 36, seed=6))])
    want = emu.trace(rays)
    for kw in (dict(), dict(inst_bias=0, n_warps=4), dict(kernel="static", n_warps=4)):
        got = H.Lockstep.trace(emu, rays, anyhit=True, **kw)
        rep = H.check_anyhit(got, want, rays, None, what="lockstep any-hit two-level %s" % kw)
        assert rep["hits"] > 100
