"""GPU parity: the CUDA device (through the C ABI) against the oracle on identical ray batches."""
import numpy as np
import pytest

import helpers as H
from luxcore_b200 import capi
from luxcore_b200 import rays as R
from luxcore_b200 import scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = capi.Device(0)
    yield d
    d.close()


def _rays_for(desc, n, seed):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1))
    return np.concatenate([a, b])


@pytest.mark.parametrize("kernel", ["persistent", "simple"])
@pytest.mark.parametrize("name,tree_type,n", [("cornell", 4, 100000), ("cornell", 2, 50000), ("cornell", 8, 50000),
                                              ("bigmonkey", 4, 200000), ("kitchen", 4, 1000000), ("kitchen", 2, 300000),
                                              ("classroom", 4, 500000), ("luxball", 8, 300000)])
def test_bvh_matches_oracle(dev, kernel, name, tree_type, n):
    dev.set_option("kernel", kernel)
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = _rays_for(desc, n, seed=21)
    ref = bvh.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits(got, ref, rays, what="%s k=%d %s" % (name, tree_type, kernel))
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()
    dev.set_option("kernel", "persistent")


def test_empty_scene_all_miss(dev):
    nodes = np.zeros(0, dtype=capi.NODE_DTYPE)
    scene = dev.upload_bvh(nodes, np.zeros((0, 3), np.float32), np.zeros(0, np.uint32))
    rays = R.to_numpy_rays(R.uniform_rays([-1, -1, -1], [1, 1, 1], 1000, seed=5))
    rays["maxt"][:500] = 7.5
    got = scene.trace_host(rays)
    assert (got["meshIndex"] == H.NULL).all() and (got["triangleIndex"] == H.NULL).all()
    assert (got["t"] == rays["maxt"]).all()
    scene.free()


def test_masked_rays_leave_hits_untouched(dev):
    desc = S.load_fixture("cornell")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = _rays_for(desc, 10000, seed=3)
    mask = np.random.default_rng(0).random(rays.shape[0]) < 0.4
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    hits = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
    hits["t"] = 123.0
    hits["meshIndex"] = 77
    hits["triangleIndex"] = 88
    scene.trace_host(rays, hits)
    assert (hits["t"][mask] == 123.0).all() and (hits["meshIndex"][mask] == 77).all() and (hits["triangleIndex"][mask] == 88).all()
    ref = bvh.intersect(rays)
    H.compare_hits(hits[~mask], ref[~mask], rays[~mask], what="masked")
    scene.free()


@pytest.mark.parametrize("kernel", ["persistent", "simple"])
@pytest.mark.parametrize("chunks", [0, -1, 1, 5])
@pytest.mark.parametrize("n", [1, 128, 129, 1000, 32768 * 3 + 1234, 500000])
def test_trace_gather_pushes_every_hit(dev, chunks, n, kernel):
    """lrb_trace_gather: chunks == 0 -> one kernel + copy-engine pushes triggered by the kernel's
    chunk-completion flags (here with 16 Ki-ray chunks so that several are in flight); chunks == -1 ->
    one kernel with dual-destination RayHit stores; chunks >= 1 -> chunked launches + copy engine.
    The gather slice must end up byte-identical to the local RayHit buffer, masked rays' records included."""
    if kernel == "simple" and n > 1000 and chunks != 0:
        pytest.skip("the static kernel's gather path is covered by the small batches")
    dev.set_option("kernel", kernel)
    dev.set_option("gather_stores", "1" if chunks < 0 else "0")
    dev.set_option("gather_chunk_shift", "14")
    chunks = max(chunks, 0)
    desc = S.load_fixture("kitchen")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = _rays_for(desc, max(n, 1000), seed=33)[:n]
    mask = np.random.default_rng(1).random(n) < (0.1 if n > 1 else 0.0)
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    d_rays = dev.alloc(n * 48)
    d_hits = dev.alloc(n * 20)
    d_dst = dev.alloc(n * 20 + 64)
    dev.h2d(d_rays, rays, blocking=True)
    pre = np.full(n, 7, dtype=np.uint8).repeat(20)
    dev.h2d(d_hits, pre, blocking=True)          # masked records keep this content
    dev.h2d(d_dst, np.zeros(n * 20 + 64, np.uint8), blocking=True)
    scene.trace_gather(d_rays, d_hits, n, d_dst, chunks)
    dev.sync()
    local = np.zeros(n, dtype=capi.HIT_DTYPE)
    pushed = np.zeros(n * 20 + 64, dtype=np.uint8)
    dev.d2h(local, d_hits)
    dev.d2h(pushed, d_dst)
    assert pushed[:n * 20].tobytes() == local.tobytes()
    assert (pushed[n * 20:] == 0).all()          # nothing written past the slice
    ref = bvh.intersect(rays)
    H.compare_hits(local[~mask], ref[~mask], what="trace_gather")
    assert (local.view(np.uint8).reshape(n, 20)[mask] == 7).all()
    for p in (d_rays, d_hits, d_dst):
        dev.free(p)
    scene.free()
    dev.set_option("gather_stores", "0")
    dev.set_option("gather_chunk_shift", "19")
    dev.set_option("kernel", "persistent")


def test_trace_host_without_hit_buffer_zeroes_masked_records(dev):
    """Scene.trace_host(rays) with no caller buffer: the records of masked rays read back as zeros, never as
    what an earlier trace left in the staging buffer."""
    desc = S.load_fixture("cornell")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = _rays_for(desc, 20000, seed=5)
    scene.trace_host(rays)                      # fills the staging buffer with real hits
    mask = np.random.default_rng(2).random(rays.shape[0]) < 0.5
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    hits = scene.trace_host(rays)
    assert (hits.view(np.uint8).reshape(-1, 20)[mask] == 0).all()
    H.compare_hits(hits[~mask], bvh.intersect(rays)[~mask], rays[~mask], what="trace_host/no buffer")
    scene.free()


@pytest.mark.parametrize("name,n,bits", [("kitchen", 600000, 5), ("cornell", 300000, 3), ("bigmonkey", 400000, 9)])
def test_sorted_ray_order_gives_identical_hits(dev, name, n, bits):
    """The coherence pre-pass (sort_rays) only changes the order rays are processed in: the RayHit
    buffer must be byte-identical to the unsorted trace, masked rays included."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = _rays_for(desc, n, seed=41)
    m = rays.shape[0]
    mask = np.random.default_rng(2).random(m) < 0.15
    rays["flags"][mask] = capi.RAY_FLAGS_MASKED
    base = np.zeros(m, dtype=capi.HIT_DTYPE)
    base["t"] = 5.0
    plain = base.copy()
    scene.trace_host(rays, plain)
    dev.set_option("sort_rays", "1")
    dev.set_option("sort_bits", str(bits))
    dev.set_option("sort_min_rays", "1024")
    try:
        srt = base.copy()
        scene.trace_host(rays, srt)
    finally:
        dev.set_option("sort_rays", "0")
        dev.set_option("sort_bits", "5")
    assert srt.tobytes() == plain.tobytes()
    ref = bvh.intersect(rays)
    H.compare_hits(srt[~mask], ref[~mask], rays[~mask], what="sorted " + name)
    scene.free()


@pytest.mark.parametrize("name,tree_type,n", [("cornell", 4, 300000), ("bigmonkey", 4, 300000), ("kitchen", 4, 1000000), ("classroom", 8, 300000)])
def test_bvh_grazing_rays_match_oracle(dev, name, tree_type, n):
    """Rays starting on / a few epsilons off the surfaces (axis-parallel ones included) and rays from
    far outside the scene: hits at the very planes of the leaf boxes and quantized node grids."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n, seed=23))
    lo, hi = desc.bbox()
    far = R.to_numpy_rays(R.uniform_rays(lo - 40 * (hi - lo), hi + 40 * (hi - lo), n // 4, seed=29))
    c = 0.5 * (lo + hi)
    far["d"] = (c[None, :] + 0.3 * (hi - lo)[None, :] * (np.random.default_rng(6).random((far.shape[0], 3)).astype(np.float32) - 0.5)) - far["o"]
    rays = np.concatenate([rays, far])
    ref = bvh.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="grazing %s k=%d" % (name, tree_type))
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()


@pytest.mark.parametrize("offset", [0, 4, 8, 12])
def test_hit_buffer_alignment(dev, offset):
    """RayHit records are written with alignment-dependent vector stores when the buffer is 16-byte
    aligned and with scalar stores otherwise: every byte offset must give the same records."""
    desc = S.load_fixture("bigmonkey")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    n = 50001
    rays = _rays_for(desc, n, seed=51)[:n]
    d_rays = dev.alloc(n * 48)
    d_hits = dev.alloc(n * 20 + 64)
    dev.h2d(d_rays, rays, blocking=True)
    dev.h2d(d_hits, np.zeros(n * 20 + 64, np.uint8), blocking=True)
    dev.set_option("wide_stores", "3")
    try:
        scene.trace(d_rays, d_hits + 16 + offset, n)
        dev.sync()
    finally:
        dev.set_option("wide_stores", "2")
    raw = np.zeros(n * 20 + 64, dtype=np.uint8)
    dev.d2h(raw, d_hits)
    assert (raw[:16 + offset] == 0).all() and (raw[16 + offset + n * 20:] == 0).all()
    got = raw[16 + offset:16 + offset + n * 20].view(capi.HIT_DTYPE)
    ref = bvh.intersect(rays)
    rep = H.compare_hits(got, ref, rays, what="offset %d" % offset)
    assert rep["bit_exact_hits"] == rep["hits"]
    dev.free(d_rays)
    dev.free(d_hits)
    scene.free()


@pytest.mark.parametrize("name,tree_type", [("kitchen", 4), ("cornell", 8)])
def test_cuda_path_against_the_reference_library(dev, name, tree_type):
    """The CUDA device against the REFERENCE ITSELF (oracle/_ref: BVHAccel built and intersected by the
    reference's own sources), not only against the restatement."""
    from oracle import refapi as RF
    if not RF.available():
        pytest.skip("oracle/_ref is not built")
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    rb = RF.BVH(H.reference_scene(desc), tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(rb.nodes().view(capi.NODE_DTYPE), verts, offs)
    rays = _rays_for(desc, 400000, seed=61)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = np.concatenate([rays, R.to_numpy_rays(R.surface_rays(p0, e1, e2, 200000, seed=62))])
    ref = rb.intersect(rays)
    got = scene.trace_host(rays)
    rep = H.compare_hits_tie_aware(got, ref.view(capi.HIT_DTYPE), rays, osc, what="vs reference library " + name)
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]
    scene.free()


@pytest.mark.parametrize("name,n", [("cornell", 400000), ("kitchen", 400000), ("bigmonkey", 200000)])
def test_in_plane_rays_bound_the_known_residual(dev, name, n):
    """GPU twin of tests/test_emulation_cpu.py::test_in_plane_rays_bound_the_known_residual, against the reference
    library itself where it is present: rays lying in the plane of a triangle make the reference's triangle test return
    rounding noise whose survival depends on its traversal order; fewer than 1 in 1 000 of such rays answer differently
    and every one of them is coplanar with the triangle one side names."""
    from oracle import refapi
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    rays = H.in_plane_rays(desc, n, seed=2)
    checker = refapi.BVH(H.reference_scene(desc), nodes=bvh.nodes()) if refapi.available() else bvh
    ref = checker.intersect(rays)
    got = scene.trace_host(rays)
    hit = ref["meshIndex"] != H.NULL
    diff = (got["meshIndex"] != ref["meshIndex"]) | (hit & (got["triangleIndex"] != ref["triangleIndex"]))
    print("%s: %d of %d in-plane rays answer differently (%s)" % (name, int(diff.sum()), n, "reference library" if refapi.available() else "oracle"))
    assert diff.sum() <= n * 1e-3
    assert (H.coplanar_with_reported(desc, rays, ref) | H.coplanar_with_reported(desc, rays, got))[diff].all()
    same = ~diff & hit
    assert (got["t"][same].view(np.uint32) == ref["t"][same].view(np.uint32)).all()
    scene.free()
