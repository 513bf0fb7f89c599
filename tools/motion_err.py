import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import helpers as H, scene_zoo as Z
from luxcore_b200 import capi, rays as R, scenes as S
from oracle import oracle as O
dev = capi.Device(0)
for which in ["zoo-motion", "bigmonkey-motion"]:
    desc = Z.motion_scene() if which == "zoo-motion" else S.load_fixture(which)
    osc = H.oracle_scene(desc); mb = O.MBVH(osc)
    a = H.mbvh_arrays(desc, mb)
    scene = dev.upload_mbvh(a["root_nodes"], a["leaf_nodes"], a["leaf_verts"], a["transforms_minv"], a["motion_table"], a["interps"])
    lo, hi = desc.bbox(); pad = 0.1 * (hi - lo)
    rays = np.concatenate([R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, 200000, seed=61, time_range=(-0.05, 1.05))),
                           R.to_numpy_rays(R.camera_rays(desc.cam, 447, 447, seed=62, time_range=(-0.05, 1.05)))])
    ref = mb.intersect(rays); got = scene.trace_host(rays)
    both = (ref["meshIndex"] != H.NULL) & (got["meshIndex"] != H.NULL) & (ref["triangleIndex"] == got["triangleIndex"]) & (ref["meshIndex"] == got["meshIndex"])
    idxbad = ((ref["meshIndex"] != got["meshIndex"]) | ((ref["triangleIndex"] != got["triangleIndex"]) & (ref["meshIndex"] != H.NULL))).sum()
    for f in ["t", "b1", "b2"]:
        e = np.abs(got[f][both] - ref[f][both]) / np.maximum(1.0, np.abs(ref[f][both]))
        print(which, f, "hits", int(both.sum()), "exact", int((e == 0).sum()), ">1e-6", int((e > 1e-6).sum()), ">1e-5", int((e > 1e-5).sum()), ">1e-4", int((e > 1e-4).sum()), "max", float(e.max()))
    print(which, "index mismatches", int(idxbad))
