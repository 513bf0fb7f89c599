#!/usr/bin/env python3
"""GPU BVH builder (lrb_build_lbvh) on the GPU box: build times by stage for the kitchen's and for soup-sized leaf sets,
against the host SAH builder of the same host layer, and the traversal cost of the trees (node visits per ray).
Writes gpurun_out/r02_builder_bench.json.  Development aid, not a bench line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench as B
from luxcore_b200 import capi, hostapi, rays as R, scenes as S

rows = []
dev = capi.Device(0)
for n, quality in ((86032, 0), (86032, 1), (1000000, 0), (1000000, 1), (10000000, 0), (10000000, 1), (50000000, 0), (50000000, 1)):
    rng = np.random.default_rng(n)
    c = rng.random((n, 3), dtype=np.float32)
    e = np.float32(0.001)
    boxes = np.concatenate([c - e, c + e], axis=1)
    dev.build_lbvh(boxes[:1000], 4, quality=quality)       # warm-up (context, CUB)
    t0 = time.perf_counter()
    nodes, tm = dev.build_lbvh(boxes, 4, quality=quality)
    wall = time.perf_counter() - t0
    row = {"leaves": n, "binary_tree": "PLOC" if quality else "radix", "launches": int(tm.kernels), "nodes": int(nodes.shape[0]), "wall_s": round(wall, 4), "h2d_ms": round(tm.h2d_ms, 3), "sort_ms": round(tm.sort_ms, 3),
           "tree_ms": round(tm.tree_ms, 3), "emit_ms": round(tm.emit_ms, 3), "d2h_ms": round(tm.d2h_ms, 3),
           "device_ms_without_copies": round(tm.sort_ms + tm.tree_ms + tm.emit_ms, 3),
           "mleaves_per_s_device": round(n / (tm.sort_ms + tm.tree_ms + tm.emit_ms) / 1e3, 1)}
    rows.append(row); print(json.dumps(row), flush=True)
    del nodes, boxes, c
dev.close()

# the kitchen through the host layer: SAH (host) vs Morton (GPU) -- build time and what the tree costs to walk
desc = S.load_fixture("kitchen")
tdev = torch.device("cuda", 0)
for builder in ("EMBREE_BINNED_SAH", "EMBREE_MORTON", "B200_PLOC"):
    t0 = time.perf_counter()
    s = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}, desc)
    s.build_accelerator("BVH")
    build_s = time.perf_counter() - t0
    s.start(0)
    scene = s.native_scene()
    cam = R.camera_rays(desc.cam, 1024, 1024, seed=3, device=tdev)
    hits = torch.empty((cam.shape[0], 20), dtype=torch.uint8, device=tdev)
    st = scene.trace_stats(cam.data_ptr(), 0, cam.shape[0])
    for _ in range(3):
        s.trace_device(cam.data_ptr(), hits.data_ptr(), cam.shape[0])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.trace_device(cam.data_ptr(), hits.data_ptr(), cam.shape[0]); b.record(); torch.cuda.synchronize()
    row = {"scene": "kitchen", "builder": builder, "build_s": round(build_s, 4), "ref_nodes": int(s.bvh_nodes().shape[0]),
           "camera_rays": int(cam.shape[0]), "nodes_per_ray": round(st.wide_nodes / max(1, st.rays), 2), "tris_per_ray": round(st.triangles / max(1, st.rays), 2),
           "camera_mrays_per_s": round(cam.shape[0] / a.elapsed_time(b) / 1e3, 1)}
    rows.append(row); print(json.dumps(row), flush=True)
    s.stop(); s.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_builder_bench.json"), "w"), indent=1)
