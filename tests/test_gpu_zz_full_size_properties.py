"""BASELINE.json's full batch sizes on the GPU, through size-independent properties (tests/properties.py:
determinism, shard and permutation invariance, hit-point consistency, minimality, reachability) plus a
sampled bit-exact comparison with the oracle.  At 16 Mi rays the oracle alone would take minutes; the
properties run on the device with torch.  (This file sorts last: it was written after the round's GPU
budget was spent; the checkers themselves are exercised on the CPU by test_properties_cpu.py.)"""
import numpy as np
import pytest
import torch

import helpers as H
import properties as P
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _session(desc, accel, builder="EMBREE_BINNED_SAH"):
    s = hostapi.Session({"accelerator.type": accel, "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}, desc)
    s.build_accelerator(accel)
    s.start(0)
    return s


def _device_trace_fn(sess, dev):
    def trace_fn(rays_u8):
        assert rays_u8.is_cuda and rays_u8.is_contiguous()
        hits = torch.empty((rays_u8.shape[0], 20), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()        # the batch was written on torch's stream, the device traces on its own queue
        sess.trace_device(rays_u8.data_ptr(), hits.data_ptr(), rays_u8.shape[0])
        sess.finish()
        return hits
    return trace_fn


def _tables(desc, dev):
    p0, e1, e2, offs = S.world_triangles(desc)
    return torch.from_numpy(p0).to(dev), torch.from_numpy(e1).to(dev), torch.from_numpy(e2).to(dev), torch.from_numpy(offs).to(dev)


def _sample_against_oracle(orc, rays, hits, k, what, **kw):
    g = torch.Generator(device="cpu")
    g.manual_seed(5)
    pick = torch.randperm(rays.shape[0], generator=g)[:k].to(rays.device)
    rn = R.to_numpy_rays(rays[pick].cpu())
    got = hits[pick].cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
    ref = orc.intersect(rn, nthreads=O.hardware_threads())
    return H.compare_hits(got, ref, rn, what=what, **kw)


@pytest.mark.parametrize("name,n_rays", [("kitchen", 16 << 20), ("classroom", 16 << 20)])
def test_interiors_16mi_bounce_rays(name, n_rays):
    """configs[3]: scenes/kitchen and scenes/classroom, 16 Mi incoherent (bounce depth 2) rays per batch."""
    import bench as B
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    desc = S.load_fixture(name)
    sess = _session(desc, "BVH")
    try:
        trace_fn = _device_trace_fn(sess, dev)
        rays = B.make_bounce_batch(trace_fn, desc, n_rays, seed=2, device=dev, depth=2)
        assert rays.shape[0] == n_rays
        hits, rep = P.check_all(trace_fn, rays, *_tables(desc, dev))
        assert rep["hits"] > 0.9 * n_rays and rep["minimality_rays"] == rep["hits"] == rep["reachability_rays"]
        orc = O.BVH(H.oracle_scene(desc), nodes=sess.bvh_nodes())
        par = _sample_against_oracle(orc, rays, hits, 200000, name + " 16Mi sample")
        assert par["bit_exact_hits"] == par["hits"] > 0
    finally:
        sess.stop()
        sess.close()


def test_luxball_4mi_camera_and_bounce4():
    """configs[1]: coherent camera rays vs incoherent 4-bounce path rays, 4 Mi-ray batches."""
    import bench as B
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    desc = S.load_fixture("luxball")
    sess = _session(desc, "BVH")
    try:
        trace_fn = _device_trace_fn(sess, dev)
        tables = _tables(desc, dev)
        orc = O.BVH(H.oracle_scene(desc), nodes=sess.bvh_nodes())
        for depth in (0, 4):
            rays = B.make_bounce_batch(trace_fn, desc, 4 << 20, seed=3, device=dev, depth=depth)
            hits, rep = P.check_all(trace_fn, rays, *tables)
            assert rep["hits"] > 0
            par = _sample_against_oracle(orc, rays, hits, 100000, "luxball depth %d sample" % depth)
            assert par["bit_exact_hits"] == par["hits"] > 0
    finally:
        sess.stop()
        sess.close()


def test_lightinstances_4mi_two_level():
    """configs[2]: 4 500 instances + 2 static meshes through the MBVH two-level traversal, 4 Mi primary and
    4 Mi bounce rays.  Instances use identical arithmetic to the reference: bit-exact on the sample."""
    import bench as B
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    desc = S.load_fixture("lightinstances")
    sess = _session(desc, "MBVH")
    try:
        assert sess.accelerator_type() == hostapi.ACCEL_MBVH
        trace_fn = _device_trace_fn(sess, dev)
        tables = _tables(desc, dev)
        orc = O.MBVH(H.oracle_scene(desc))         # the oracle's own (CLASSIC) trees: results do not depend on topology
        for depth in (0, 1):
            rays = B.make_bounce_batch(trace_fn, desc, 4 << 20, seed=4, device=dev, depth=depth)
            hits, rep = P.check_all(trace_fn, rays, *tables)
            assert rep["hits"] > 0
            par = _sample_against_oracle(orc, rays, hits, 50000, "lightinstances depth %d sample" % depth)
            assert par["bit_exact_hits"] == par["hits"] > 0
    finally:
        sess.stop()
        sess.close()
