#!/usr/bin/env python3
"""Writes tests/golden/ref_vectors.npz: inputs and the outputs THE REFERENCE ITSELF produced for them
(oracle/_ref = the reference's C++ sources compiled from /root/reference, oracle/ref/Makefile).
Run in the container that holds /root/reference:  python tools/make_ref_vectors.py
The vectors let tests/test_oracle_pinned_cpu.py::test_golden_vectors_from_the_reference pin the
oracle on machines where the reference sources (and the prebuilt oracle/_ref) are absent."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import helpers as H  # noqa: E402
import scene_zoo as Z  # noqa: E402
from luxcore_b200 import rays as R, scenes as S  # noqa: E402
from oracle import refapi as RF  # noqa: E402


def stress_rays(desc, n, seed, time_range=None):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed, time_range=time_range))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1, time_range=time_range))
    p0, e1, e2, _ = S.world_triangles(desc)
    c = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n // 2, seed=seed + 2, axis_fraction=0.3))
    if time_range is not None:
        c["time"] = np.random.default_rng(seed).random(c.shape[0]).astype(np.float32) * (time_range[1] - time_range[0]) + time_range[0]
    return np.concatenate([a, b, c])


def main():
    assert RF.available(), "oracle/_ref cannot be built here"
    out = {}
    rng = np.random.default_rng(101)
    eps_in = np.concatenate([rng.standard_normal(500) * 10.0 ** rng.integers(-30, 30, 500), [0.0, 1.0, -1.0, 1e-45, 3.4e38, 0.1]]).astype(np.float32)
    out["eps_in"] = eps_in
    out["eps_out"] = np.asarray([RF.epsilon(float(v)) for v in eps_in], dtype=np.float32)
    mats = rng.standard_normal((64, 4, 4)).astype(np.float32)
    mats[::2, 3] = [0, 0, 0, 1]
    out["minv_in"] = mats
    out["minv_out"] = np.stack([RF.matrix_inverse(m) for m in mats])
    # Triangle::Intersect / BBox::IntersectP
    n = 4000
    rays = R.to_numpy_rays(R.uniform_rays([-1, -1, -1], [1, 1, 1], n, seed=102))
    rays["d"][::7] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, rays[::7].shape[0])]
    rays["mint"][::5] = 0.0
    tri = (rng.random((n, 3, 3)).astype(np.float32) - 0.5) * 2
    tri[::11, 2] = tri[::11, 1]
    tri[::13, :, 2] = rays["o"][::13, None, 2]
    hit = np.zeros(n, np.uint8)
    tb = np.zeros((n, 3), np.float32)
    boxhit = np.zeros(n, np.uint8)
    for i in range(n):
        h, v = RF.triangle_intersect(rays[i:i + 1], tri[i, 0], tri[i, 1], tri[i, 2])
        hit[i], tb[i] = h, v
        boxhit[i] = RF.bbox_intersectp(rays[i:i + 1], tri[i].min(axis=0), tri[i].max(axis=0))
    out.update(tri_rays=rays.view(np.uint8).reshape(n, 48), tri_verts=tri, tri_hit=hit, tri_tb=tb, box_hit=boxhit)
    # BVHAccel (CLASSIC builder) on cornell, arity 4 and 8
    desc = S.load_fixture("cornell")
    for k in (4, 8):
        b = RF.BVH(H.reference_scene(desc), tree_type=k)
        r = stress_rays(desc, 3000, seed=103 + k)
        nodes = b.nodes()
        nodes["pad0"] = 0                                  # uninitialised in the reference
        words = nodes.view(np.uint32).reshape(-1, 8)
        words[(words[:, 6] >> 31) == 1, 5] = 0             # unused tail of the triangle-leaf payload
        out["cornell%d_nodes" % k] = nodes.view(np.uint8).reshape(-1, 32)
        out["cornell%d_rays" % k] = r.view(np.uint8).reshape(-1, 48)
        out["cornell%d_hits" % k] = b.intersect(r).view(np.uint8).reshape(-1, 20)
    # MBVHAccel: instances and motion blur (scene zoo)
    for name, mk, tr in (("zooinst", Z.instances_scene, None), ("zoomotion", Z.motion_scene, (-0.1, 1.1))):
        d = mk()
        m = RF.MBVH(H.reference_scene(d))
        r = stress_rays(d, 3000, seed=107, time_range=tr)
        out[name + "_rays"] = r.view(np.uint8).reshape(-1, 48)
        out[name + "_hits"] = m.intersect(r).view(np.uint8).reshape(-1, 20)
        out[name + "_bboxes"] = np.stack([m.scene.mesh_bbox(i) for i in range(len(d.meshes))])
    path = os.path.join(ROOT, "tests", "golden", "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", RF.lib().ref_describe().decode())


if __name__ == "__main__":
    main()
