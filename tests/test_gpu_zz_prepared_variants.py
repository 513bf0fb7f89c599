"""GPU parity of kernel variants that are prepared behind options and OFF by default (written after
the round's GPU budget was spent; this file sorts last on purpose).  Same bar as every other variant:
bit-exact against the oracle."""
import numpy as np
import pytest

import helpers as H
from luxcore_b200 import capi, hostapi
from luxcore_b200 import rays as R
from luxcore_b200 import scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_prefetch_variant_matches_oracle():
    """`prefetch` = 1: TracePersistent<false, true, *, PREFETCH> requests the records of pushed children
    into L2 (meant for scenes larger than L2).  Prefetches are hints: results must not change."""
    dev = capi.Device(0)
    try:
        desc = S.load_fixture("kitchen")
        s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
        s.build_accelerator("BVH")
        nodes = s.bvh_nodes().copy()
        osc = H.oracle_scene(desc)
        bvh = O.BVH(osc, nodes=nodes)
        verts, offs = H.flattened_from_oracle(desc, osc)
        scene = dev.upload_bvh(nodes, verts, offs)
        assert scene.info().stack_need > 16          # the spilling variant is the one that has the prefetching twin
        lo, hi = desc.bbox()
        rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo, hi, 400000, seed=61)),
                               R.to_numpy_rays(R.camera_rays(desc.cam, 600, 600, seed=62))])
        ref = bvh.intersect(rays)
        base = scene.trace_host(rays)
        dev.set_option("prefetch", 1)
        got = scene.trace_host(rays)
        dev.set_option("prefetch", 0)
        assert got.tobytes() == base.tobytes()
        rep = H.compare_hits(got, ref, rays, what="kitchen prefetch=1")
        assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0.3 * rep["n"]
        with pytest.raises(Exception):
            dev.set_option("prefetch", 3)
        scene.free()
    finally:
        dev.close()


@pytest.mark.parametrize("inst_bias", [0, 4, 64])
def test_instance_vote_settings_match_oracle(inst_bias):
    """Two-level kernel: entering instances is a voted phase (`inst_bias`, default 8).  Any setting -- enter at
    once (0), eager (4), wait until nothing else is left (64) -- must give the oracle's hits bit for bit."""
    import scene_zoo as Z
    dev = capi.Device(0)
    try:
        for desc in (Z.instances_scene(), S.load_fixture("lightinstances", max_objects=800)):
            osc = H.oracle_scene(desc)
            mb = O.MBVH(osc, tree_type=4)
            a = H.mbvh_arrays(desc, mb)
            scene = dev.upload_mbvh(a["root_nodes"], a["leaf_nodes"], a["leaf_verts"], a["transforms_minv"], a["motion_table"], a["interps"])
            assert scene.info().two_level == 1
            lo, hi = desc.bbox()
            rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo, hi, 200000, seed=71)),
                                   R.to_numpy_rays(R.camera_rays(desc.cam, 400, 400, seed=72))])
            ref = mb.intersect(rays)
            dev.set_option("inst_bias", inst_bias)
            got = scene.trace_host(rays)
            dev.set_option("inst_bias", 8)
            rep = H.compare_hits(got, ref, rays, what="%s inst_bias=%d" % (desc.name, inst_bias))
            assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0.1 * rep["n"]
            scene.free()
    finally:
        dev.close()
