#!/usr/bin/env python3
"""lrb_bvh_build_scene on the GPU box: triangles in, traceable scene out -- stage times (CUDA events) and wall time for the
kitchen and for soups, next to the two-step path of the same builder (lrb_build_bvh from host boxes + lrb_bvh_upload with
the host re-layout).  Writes gpurun_out/r02_scene_build_bench.json.  Development aid, not a bench line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as H
from luxcore_b200 import capi, scenes as S
from oracle import oracle as O

sizes = [int(a) for a in sys.argv[1:]] or [0, 1000000, 4000000]
rows = []
dev = capi.Device(0)
for n in sizes:
    desc = S.load_fixture("kitchen") if n == 0 else S.random_soup(n, seed=4, size=0.002 * (50e6 / n) ** (1.0 / 3.0), name="soup")
    tri, toff = H.flattened_triangles(desc)
    if n == 0:
        verts, voff = H.flattened_from_oracle(desc, H.oracle_scene(desc))
    else:
        verts, voff = np.ascontiguousarray(desc.shapes[0][0], dtype=np.float32), np.zeros(1, np.uint32)
    for quality in (0, 1):
        sc, _, _ = dev.build_scene(verts[:300], [0], np.arange(300, dtype=np.uint32).reshape(-1, 3), [0, 100], 4, quality)     # warm-up (context, CUB)
        sc.free()
        t0 = time.perf_counter()
        scene, tm, _ = dev.build_scene(verts, voff, tri, toff, 4, quality)
        wall = time.perf_counter() - t0
        info = scene.info()
        row = {"scene": "kitchen" if n == 0 else "soup", "triangles": int(tri.shape[0]), "binary_tree": "PLOC" if quality else "radix",
               "path": "lrb_bvh_build_scene (device boxes + tree + payload + lay-out)", "wall_s": round(wall, 4),
               **{k: (round(v, 3) if isinstance(v, float) else int(v)) for k, v in tm.as_dict().items()},
               "ref_nodes": int(info.n_ref_nodes), "wide_nodes": int(info.n_wide_nodes), "device_bytes": int(info.device_bytes)}
        rows.append(row); print(json.dumps(row), flush=True)
        scene.free()
        if n <= 4000000:
            # the two-step path on the same input: boxes on the host, array down, host re-layout, lay-out up
            t0 = time.perf_counter()
            p = verts[(tri + voff[np.searchsorted(toff, np.arange(tri.shape[0]), side="right") - 1][:, None]).reshape(-1)].reshape(-1, 3, 3)
            lo, hi = p.min(axis=1), p.max(axis=1)
            boxes = np.concatenate([lo - 1e-5, hi + 1e-5], axis=1).astype(np.float32)
            t1 = time.perf_counter()
            nodes, btm = dev.build_lbvh(boxes, 4, node_dtype=O.NODE_DTYPE, quality=quality)
            t2 = time.perf_counter()
            leaf = (nodes["nodeData"] >> 31) == 1
            g = nodes["w"][leaf, 0].astype(np.int64)
            m = np.searchsorted(toff, g, side="right") - 1
            w = nodes["w"]
            w[leaf, 0:3] = tri[g]; w[leaf, 3] = m; w[leaf, 4] = g - toff[m]
            t3 = time.perf_counter()
            up = dev.upload_bvh(nodes, verts, voff)
            t4 = time.perf_counter()
            row = {"scene": row["scene"], "triangles": row["triangles"], "binary_tree": row["binary_tree"], "path": "two-step (numpy boxes, lrb_build_bvh, numpy payload, lrb_bvh_upload)",
                   "wall_s": round(t4 - t0, 4), "boxes_s": round(t1 - t0, 4), "build_s": round(t2 - t1, 4), "payload_s": round(t3 - t2, 4), "relayout_upload_s": round(t4 - t3, 4)}
            rows.append(row); print(json.dumps(row), flush=True)
            up.free()
dev.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_scene_build_bench.json"), "w"), indent=1)
