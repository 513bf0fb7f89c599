"""GPU parity of the rows SURVEY.md 8(f) calls "next": shadow rays (any-hit), the pass-through re-trace loop of
Scene::Intersect (src/slg/scene/scene.cpp:556-690) and dead-lane compaction between launches
(pathoclbase_kernels_micro.cl:34-106), all through the C ABI and all against the oracle."""
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


def _rays_for(desc, n, seed, grazing=False):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1))
    out = [a, b]
    if grazing:
        # shadow-ray like: finite maxt, half of them short
        c = a.copy()
        c["maxt"] = np.random.default_rng(seed).random(c.shape[0]).astype(np.float32) * np.float32(np.linalg.norm(hi - lo))
        out.append(c)
    return np.concatenate(out)


def _upload_bvh(dev, desc, tree_type=4):
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    return bvh, dev.upload_bvh(bvh.nodes(), verts, offs)


class DevBuf:
    def __init__(self, dev, arr):
        self.dev, self.n = dev, arr.nbytes
        self.p = dev.alloc(max(arr.nbytes, 4))
        if arr.nbytes:
            dev.h2d(self.p, arr, blocking=True)

    def read(self, dtype, count):
        out = np.zeros(count, dtype=dtype)
        self.dev.d2h(out, self.p)
        return out

    def free(self):
        self.dev.free(self.p)


@pytest.mark.parametrize("kernel", ["persistent", "simple"])
@pytest.mark.parametrize("name,tree_type,n", [("cornell", 4, 100000), ("kitchen", 4, 600000), ("kitchen", 2, 200000),
                                              ("classroom", 8, 300000), ("bigmonkey", 4, 200000)])
def test_anyhit_hit_miss_equals_closest_hit(dev, kernel, name, tree_type, n):
    dev.set_option("kernel", kernel)
    desc = S.load_fixture(name)
    bvh, scene = _upload_bvh(dev, desc, tree_type)
    rays = _rays_for(desc, n, seed=71, grazing=True)
    m = rays.shape[0]
    mask = np.random.default_rng(3).random(m) < 0.15
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    d_rays = DevBuf(dev, rays)
    pre = np.full(m * 20, 9, dtype=np.uint8)
    d_hits = DevBuf(dev, pre)
    scene.trace_anyhit(d_rays.p, d_hits.p, m)
    dev.sync()
    got = d_hits.read(capi.HIT_DTYPE, m)
    assert (got.view(np.uint8).reshape(m, 20)[mask] == 9).all()         # masked rays: RayHit untouched
    ref = bvh.intersect(rays[~mask])
    rep = H.check_anyhit(got[~mask], ref, rays[~mask], desc, what="%s k=%d %s" % (name, tree_type, kernel))
    assert rep["hits"] > 0.1 * rep["rays"]
    d_rays.free(); d_hits.free(); scene.free()
    dev.set_option("kernel", "persistent")


@pytest.mark.parametrize("which", ["zoo-inst", "lightinstances", "zoo-motion"])
def test_anyhit_two_level(dev, which):
    desc = {"zoo-inst": Z.instances_scene, "zoo-motion": Z.motion_scene}.get(which, lambda: S.load_fixture(which, max_objects=800))()
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc)
    a = H.mbvh_arrays(desc, mb)
    scene = dev.upload_mbvh(a["root_nodes"], a["leaf_nodes"], a["leaf_verts"], a["transforms_minv"], a["motion_table"], a["interps"])
    lo, hi = desc.bbox()
    tr = (0.0, 1.0) if which == "zoo-motion" else None
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - 0.1, hi + 0.1, 200000, seed=81, time_range=tr)),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 400, 400, seed=82, time_range=tr))])
    m = rays.shape[0]
    d_rays, d_hits = DevBuf(dev, rays), DevBuf(dev, np.zeros(m * 20, np.uint8))
    scene.trace_anyhit(d_rays.p, d_hits.p, m)
    dev.sync()
    got = d_hits.read(capi.HIT_DTYPE, m)
    ref = mb.intersect(rays)
    if which == "zoo-motion":
        # hit / miss of a ray grazing a moving silhouette can depend on the last bit of sinf / acosf (helpers.compare_hits)
        diff = (got["meshIndex"] == H.NULL) != (ref["meshIndex"] == H.NULL)
        assert diff.mean() < 1e-4
    else:
        rep = H.check_anyhit(got, ref, rays, None, what=which)
        assert rep["hits"] > 0.05 * rep["rays"]
    d_rays.free(); d_hits.free(); scene.free()


def test_advance_rays_follows_the_reference_rule(dev):
    """lrb_advance_rays == scene.cpp:646-680 applied to a traced batch (mint = t + MachineEpsilon::E(t), the
    "not enough precision" exit, RAY_FLAGS_MASKED for the finished rays), bit for bit."""
    desc = S.load_fixture("kitchen")
    bvh, scene = _upload_bvh(dev, desc)
    rays = _rays_for(desc, 300000, seed=91, grazing=True)
    m = rays.shape[0]
    rng = np.random.default_rng(7)
    rays["flags"][rng.random(m) < 0.1] = capi.RAY_FLAGS_MASKED
    hits = bvh.intersect(rays)
    # adversarial records: hits at t where t + E(t) >= maxt, and t so large that t + E(t) == t
    k = np.flatnonzero(hits["meshIndex"] != H.NULL)
    tight = k[:2000]
    rays["maxt"][tight] = hits["t"][tight] + H.machine_epsilon_np(hits["t"][tight]) * np.float32(0.5)
    huge = k[2000:2500]
    hits["t"][huge] = np.float32(3e7)
    rays["maxt"][huge] = np.float32(np.inf)
    n_mesh = len(desc.meshes)
    pass_mesh = rng.random(n_mesh) < 0.5
    words = np.zeros((n_mesh + 31) // 32, dtype=np.uint32)
    for i in np.flatnonzero(pass_mesh):
        words[i >> 5] |= np.uint32(1 << (i & 31))
    cont_flags = (rng.random(m) < 0.05).astype(np.uint8)
    cont_flags[tight] = 1; cont_flags[huge] = 1            # the adversarial records do continue ...
    rays["flags"][tight] = 0; rays["flags"][huge] = 0      # ... and are live

    d_rays, d_hits, d_words, d_flags = DevBuf(dev, rays), DevBuf(dev, hits), DevBuf(dev, words), DevBuf(dev, cont_flags)
    n_cont = scene.advance_rays(d_rays.p, d_hits.p, m, d_words.p, words.shape[0], d_flags.p)
    got_rays, got_hits = d_rays.read(capi.RAY_DTYPE, m), d_hits.read(capi.HIT_DTYPE, m)

    want_rays, want_hits = rays.copy(), hits.copy()
    live = (rays["flags"] & 1) == 0
    hit = hits["meshIndex"] != H.NULL
    cont = live & hit & (pass_mesh[np.minimum(hits["meshIndex"], n_mesh - 1)] | (cont_flags != 0))
    t = hits["t"]
    mint = (t + H.machine_epsilon_np(t)).astype(np.float32)
    dead_end = cont & ((mint == t) | (mint >= rays["maxt"]))
    armed = cont & ~dead_end
    want_rays["mint"][armed] = mint[armed]
    want_rays["flags"][live & ~armed] |= 1
    want_hits["t"][dead_end] = rays["maxt"][dead_end]
    want_hits["b1"][dead_end] = 0; want_hits["b2"][dead_end] = 0
    want_hits["meshIndex"][dead_end] = H.NULL; want_hits["triangleIndex"][dead_end] = H.NULL
    assert dead_end.sum() >= 2000 and armed.sum() > 1000
    assert n_cont == int(armed.sum())
    assert got_rays.tobytes() == want_rays.tobytes()
    assert got_hits.tobytes() == want_hits.tobytes()
    for b in (d_rays, d_hits, d_words, d_flags):
        b.free()
    scene.free()


@pytest.mark.parametrize("name,frac", [("kitchen", 0.5), ("classroom", 0.3), ("bigmonkey", 0.7)])
def test_passthrough_loop_matches_the_reference_loop(dev, name, frac):
    """lrb_trace_passthrough against the loop of Scene::Intersect run with the oracle's closest hit: a random
    subset of the meshes is "pass-through" (camera-invisible / fully transparent objects)."""
    desc = S.load_fixture(name)
    bvh, scene = _upload_bvh(dev, desc)
    rays = _rays_for(desc, 300000, seed=101)
    m = rays.shape[0]
    rng = np.random.default_rng(11)
    masked = rng.random(m) < 0.1
    rays["flags"][masked] = capi.RAY_FLAGS_MASKED
    n_mesh = len(desc.meshes)
    pass_mesh = rng.random(n_mesh) < frac
    words = np.zeros((n_mesh + 31) // 32, dtype=np.uint32)
    for i in np.flatnonzero(pass_mesh):
        words[i >> 5] |= np.uint32(1 << (i & 31))

    # the reference loop, batch-wise
    want = np.zeros(m, dtype=capi.HIT_DTYPE)
    work = rays.copy()
    live = np.flatnonzero(~masked)
    rounds_ref, traced_ref = 0, 0
    while live.size:
        h = bvh.intersect(work[live])
        want[live] = h
        rounds_ref += 1; traced_ref += live.size
        cont = (h["meshIndex"] != H.NULL) & pass_mesh[np.minimum(h["meshIndex"], n_mesh - 1)]
        mint = (h["t"] + H.machine_epsilon_np(h["t"])).astype(np.float32)
        dead = cont & ((mint == h["t"]) | (mint >= work["maxt"][live]))
        di = live[dead]
        want["t"][di] = work["maxt"][di]; want["b1"][di] = 0; want["b2"][di] = 0
        want["meshIndex"][di] = H.NULL; want["triangleIndex"][di] = H.NULL
        go = cont & ~dead
        work["mint"][live[go]] = mint[go]
        live = live[go]
        assert rounds_ref < 200

    pre = np.full(m * 20, 5, dtype=np.uint8)
    d_rays, d_hits, d_words = DevBuf(dev, rays), DevBuf(dev, pre), DevBuf(dev, words)
    rounds, traced = scene.trace_passthrough(d_rays.p, d_hits.p, m, d_words.p, words.shape[0], max_rounds=200)
    got = d_hits.read(capi.HIT_DTYPE, m)
    assert (got.view(np.uint8).reshape(m, 20)[masked] == 5).all()
    assert rounds_ref > 2
    rep = H.compare_hits(got[~masked], want[~masked], what="passthrough/" + name)
    assert rep["bit_exact_hits"] >= rep["hits"] - rep["tie_exempt"]
    # no surviving hit lies on a pass-through mesh
    hm = got["meshIndex"][~masked]
    assert not pass_mesh[hm[hm != H.NULL]].any()
    if rep["tie_exempt"] == 0:
        assert rounds == rounds_ref
    for b in (d_rays, d_hits, d_words):
        b.free()
    scene.free()


@pytest.mark.parametrize("frac", [0.0, 0.3, 0.7, 0.97, 1.0])
@pytest.mark.parametrize("n", [1, 1000, 1024, 1025, 700001])
def test_compaction_lists_exactly_the_live_rays(dev, n, frac):
    desc = S.load_fixture("cornell")
    bvh, scene = _upload_bvh(dev, desc)
    rays = _rays_for(desc, max(n, 1000), seed=111)[:n]
    masked = np.random.default_rng(5).random(n) < frac
    if frac == 1.0:
        masked[:] = True
    rays["flags"][masked] = capi.RAY_FLAGS_MASKED
    d_rays = DevBuf(dev, rays)
    idx_p, cnt_p, count = dev.compact_rays(d_rays.p, n)
    want = np.flatnonzero(~masked).astype(np.uint32)
    assert count == want.shape[0]
    got = np.zeros(max(count, 1), dtype=np.uint32)
    if count:
        dev.d2h(got[:count], idx_p)
        assert (got[:count] == want).all()
    # tracing through the list == tracing the batch (masked rays skipped inside the kernel), byte for byte
    pre = np.full(n * 20, 3, dtype=np.uint8)
    a, b = DevBuf(dev, pre), DevBuf(dev, pre)
    scene.trace(d_rays.p, a.p, n)
    scene.trace_indexed(d_rays.p, b.p, n, idx_p, cnt_p)
    dev.sync()
    assert a.read(np.uint8, n * 20).tobytes() == b.read(np.uint8, n * 20).tobytes()
    dev.set_option("compact", "1")
    c = DevBuf(dev, pre)
    scene.trace(d_rays.p, c.p, n)
    dev.sync()
    dev.set_option("compact", "0")
    assert a.read(np.uint8, n * 20).tobytes() == c.read(np.uint8, n * 20).tobytes()
    for x in (d_rays, a, b, c):
        x.free()
    scene.free()


def test_compaction_with_the_static_kernel_and_shadow_rays(dev):
    desc = S.load_fixture("kitchen")
    bvh, scene = _upload_bvh(dev, desc)
    rays = _rays_for(desc, 200000, seed=121)
    n = rays.shape[0]
    masked = np.random.default_rng(6).random(n) < 0.6
    rays["flags"][masked] = capi.RAY_FLAGS_MASKED
    d_rays = DevBuf(dev, rays)
    idx_p, cnt_p, count = dev.compact_rays(d_rays.p, n)
    ref = bvh.intersect(rays[~masked])
    for kernel in ("persistent", "simple"):
        dev.set_option("kernel", kernel)
        pre = np.full(n * 20, 3, dtype=np.uint8)
        h = DevBuf(dev, pre)
        scene.trace_indexed(d_rays.p, h.p, n, idx_p, cnt_p, any_hit=True)
        dev.sync()
        got = h.read(capi.HIT_DTYPE, n)
        assert (got.view(np.uint8).reshape(n, 20)[masked] == 3).all()
        H.check_anyhit(got[~masked], ref, rays[~masked], desc, what="indexed any-hit " + kernel)
        h.free()
    dev.set_option("kernel", "persistent")
    d_rays.free(); scene.free()


def test_host_layer_shadow_rays_and_pass_through(dev):
    """The same two extensions through the luxrays:: plugin surface (CUDAIntersectionDevice::
    EnqueueTraceShadowRayBuffer / AdvancePassThroughRayBuffer, include/luxrays/devices/cudaintersectiondevice.h)."""
    import torch
    from luxcore_b200 import hostapi
    desc = S.load_fixture("bigmonkey")
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH"}, desc)
    sess.build_accelerator("BVH")
    sess.start(0)
    rays = _rays_for(desc, 200000, seed=131, grazing=True)
    n = rays.shape[0]
    t_rays = torch.from_numpy(rays.view(np.uint8).reshape(n, 48)).cuda()
    t_hits = torch.zeros((n, 20), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    sess.trace_device_shadow(t_rays.data_ptr(), t_hits.data_ptr(), n)
    sess.finish()
    got = t_hits.cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
    orc = O.BVH(H.oracle_scene(desc), nodes=sess.bvh_nodes())
    H.check_anyhit(got, orc.intersect(rays), rays, desc, what="host layer shadow rays")
    assert sess.total_rays() == n
    sess.stop(); sess.close()


@pytest.mark.parametrize("n_floats,n_tiles", [(4 * 512 * 512, 8), (4 * 333 * 77 + 3, 5), (17, 16), (4 * 1024, 1)])
def test_film_reduce_is_the_reference_sum(dev, n_floats, n_tiles):
    """lrb_film_reduce == Film::AddFilm applied film by film in device order (film.cpp:707-760): binary32 adds in tile
    order, bit for bit; here every "rank" is a buffer of the one GPU and the slices are summed by separate launches."""
    from luxcore_b200 import shard
    rng = np.random.default_rng(n_floats)
    tiles = [(rng.standard_normal(n_floats).astype(np.float32) * np.float32(10.0 ** rng.integers(-4, 5))) for _ in range(n_tiles)]
    tiles[0][:7] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3e38][:min(7, n_floats)]
    bufs = [DevBuf(dev, t) for t in tiles]
    dst = DevBuf(dev, np.full(n_floats + 4, 7.0, dtype=np.float32))
    world = 3
    for r in range(world):
        first, count = shard.film_slice(n_floats, world, r)
        dev.film_reduce([b.p for b in bufs], dst.p, first, count)
    dev.sync()
    got = dst.read(np.float32, n_floats + 4)
    want = np.zeros(n_floats, dtype=np.float32)
    with np.errstate(all="ignore"):
        for t in tiles:
            want = (want + t).astype(np.float32)
    # NaN payloads are exempt: the device's add returns the canonical 0x7fffffff, SSE addss propagates the operand's
    nan = np.isnan(want)
    assert (np.isnan(got[:n_floats]) == nan).all()
    assert got[:n_floats][~nan].tobytes() == want[~nan].tobytes()
    assert (got[n_floats:] == 7.0).all()
    for b in bufs + [dst]:
        b.free()
