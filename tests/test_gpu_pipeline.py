"""The reference-facing call sequence -- AllocBufferRW(src) / EnqueueTraceRayBuffer / EnqueueReadBuffer / FinishQueue =
lrb_h2d / lrb_trace / lrb_d2h / lrb_sync -- is pipelined chunk by chunk behind its own interface (include/luxrays_b200.h,
"memory + queue").  The pipelining must be invisible: the same bytes as the plain sequence (device option pipeline = 0),
for pinned and pageable host memory, with pre-loaded RayHit buffers and masked rays, and for every call pattern that
does NOT match the pipeline's shape (the queue stays in order)."""
import numpy as np
import pytest
import torch

import helpers as H
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

N = 3 << 20         # 144 MB of rays: three chunks of 2^20 rays


@pytest.fixture(scope="module")
def world():
    desc = S.load_fixture("kitchen")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    dev = capi.Device(0)
    scene = dev.upload_bvh(bvh.nodes(), verts, offs)
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo, hi, N + 12345, seed=71))
    yield dev, scene, bvh, rays
    scene.free()
    dev.close()


def _sequence(dev, scene, rays, hits_in=None, blocking_read=False):
    n = rays.shape[0]
    d_r = dev.alloc(n * 48)
    d_h = dev.alloc(n * 20)
    dev.h2d(d_r, rays)                                     # AllocBufferRW(&rays, hostRays)
    if hits_in is not None:
        dev.h2d(d_h, hits_in)                              # AllocBufferRW(&hits, hostHits)
    scene.trace(d_r, d_h, n)                               # EnqueueTraceRayBuffer
    out = np.empty(n, dtype=capi.HIT_DTYPE)
    dev.d2h(out, d_h, blocking=blocking_read)              # EnqueueReadBuffer
    dev.sync()                                             # FinishQueue
    dev.free(d_r); dev.free(d_h)
    return out


def test_pipelined_sequence_equals_the_plain_one(world):
    dev, scene, bvh, rays = world
    dev.set_option("pipeline", 0)
    plain = _sequence(dev, scene, rays)
    dev.set_option("pipeline", 1)
    c0 = dev.counters().trace_launches
    piped = _sequence(dev, scene, rays)
    assert dev.counters().trace_launches - c0 == len(H.chunk_ends(rays.shape[0])) == 7      # really chunked: 2 full chunks, then the tail halves down to 64 Ki rays (host_chunks.h)
    assert piped.tobytes() == plain.tobytes()
    assert _sequence(dev, scene, rays, blocking_read=True).tobytes() == plain.tobytes()
    # and both are the reference's answer
    sel = np.random.default_rng(1).choice(rays.shape[0], 100000, replace=False)
    rep = H.compare_hits(piped[sel], bvh.intersect(rays[sel]), rays[sel], what="pipelined plugin sequence")
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    # pinned host memory (the case the pipeline is for)
    pr = torch.from_numpy(rays.view(np.uint8).reshape(-1, 48)).pin_memory()
    pinned = _sequence(dev, scene, pr.numpy().view(capi.RAY_DTYPE).reshape(-1))
    assert pinned.tobytes() == plain.tobytes()


def test_preloaded_hits_and_masked_rays(world):
    dev, scene, bvh, rays = world
    r = rays.copy()
    mask = np.random.default_rng(2).random(r.shape[0]) < 0.3
    r["flags"][mask] = capi.RAY_FLAGS_MASKED
    pre = np.zeros(r.shape[0], dtype=capi.HIT_DTYPE)
    pre["t"] = -7.0
    pre["meshIndex"] = 99
    dev.set_option("pipeline", 0)
    plain = _sequence(dev, scene, r, hits_in=pre)
    dev.set_option("pipeline", 1)
    piped = _sequence(dev, scene, r, hits_in=pre)
    assert piped.tobytes() == plain.tobytes()
    assert (piped["t"][mask] == -7.0).all() and (piped["meshIndex"][mask] == 99).all()      # masked rays keep the caller's record


def test_calls_that_break_the_pattern_stay_in_order(world):
    dev, scene, bvh, rays = world
    n = rays.shape[0]
    dev.set_option("pipeline", 1)
    d_r = dev.alloc(n * 48); d_h = dev.alloc(n * 20); d_h2 = dev.alloc(n * 20)
    # upload, then read the SAME buffer back without a trace in between
    dev.h2d(d_r, rays)
    back = np.empty(n, dtype=capi.RAY_DTYPE)
    dev.d2h(back, d_r, blocking=True)
    assert back.tobytes() == rays.tobytes()
    # upload, trace only a PART of it, trace again into another buffer, read both (neither read matches a chunked trace)
    dev.h2d(d_r, rays)
    scene.trace(d_r, d_h, n // 2)
    scene.trace(d_r, d_h2, n)
    a = np.empty(n // 2, dtype=capi.HIT_DTYPE); b = np.empty(n, dtype=capi.HIT_DTYPE)
    dev.d2h(a, d_h, blocking=False)
    dev.d2h(b, d_h2, blocking=False)
    dev.sync()
    assert a.tobytes() == b[:n // 2].tobytes()
    # chunked trace, then a second upload into the ray buffer before the read: the read still sees the first trace
    dev.h2d(d_r, rays)
    scene.trace(d_r, d_h, n)
    other = rays[::-1].copy()
    dev.h2d(d_r, other)
    c = np.empty(n, dtype=capi.HIT_DTYPE)
    dev.d2h(c, d_h, blocking=False)
    dev.sync()
    assert c.tobytes() == b.tobytes()
    # ... and the next trace sees the second upload
    scene.trace(d_r, d_h, n)
    d = np.empty(n, dtype=capi.HIT_DTYPE)
    dev.d2h(d, d_h, blocking=True)
    assert d.tobytes() == b[::-1].tobytes()
    for p in (d_r, d_h, d_h2):
        dev.free(p)


def test_host_layer_sequence_uses_the_pipeline():
    """hostapi.Session.trace_host = AllocBufferRW / EnqueueTraceRayBuffer / EnqueueReadBuffer / FinishQueue of the C++ mirror classes."""
    desc = S.load_fixture("kitchen")
    s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH"}, desc)
    s.start(0)
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo, hi, 2 << 20, seed=72))
    c0 = s.counters().trace_launches
    got = s.trace_host(rays)
    assert s.counters().trace_launches - c0 == len(H.chunk_ends(rays.shape[0])) == 6
    sel = np.random.default_rng(3).choice(rays.shape[0], 60000, replace=False)
    ref = O.BVH(H.oracle_scene(desc), nodes=s.bvh_nodes()).intersect(rays[sel])
    rep = H.compare_hits(got[sel], ref, rays[sel], what="host layer, pipelined")
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
    s.stop()
    s.close()
