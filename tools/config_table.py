#!/usr/bin/env python3
"""Throughput + spot parity of the BASELINE.json configurations that are NOT the bench line (GPU box).
Writes one JSON object per configuration to stdout / gpurun_out/config_table.json.  Development
aid: the numbers are CUDA-event timings of lrb_trace over device-resident batches, like bench.py's
`value`; parity is checked against the oracle on a sample of each batch."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
import bench as B
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

import argparse
_ap = argparse.ArgumentParser()
_ap.add_argument("--only", action="append", default=[], help="run only these scenes")
_ap.add_argument("--opt", action="append", default=[], help="device option key=value")
ARGS = _ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
out = []

def run(name, accel, builder, kinds, n, max_objects=None, time_range=None, sample=100000):
    if ARGS.only and name not in ARGS.only:
        return
    try:
        desc = S.load_fixture(name, max_objects=max_objects) if max_objects else S.load_fixture(name)
        t0 = time.perf_counter()
        sess = hostapi.Session({"accelerator.type": accel, "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}, desc)
        sess.build_accelerator(accel)
        build_s = time.perf_counter() - t0
        sess.start(0); sess.set_stream(stream.cuda_stream)
        for kv in ARGS.opt:
            k, v = kv.split("=", 1)
            sess.set_option(k, v)
        scene = sess.native_scene(); info = scene.info()
        def trace_fn(r):
            h = torch.empty((r.shape[0], 20), dtype=torch.uint8, device=dev)
            sess.trace_device(r.data_ptr(), h.data_ptr(), r.shape[0]); return h
        osc = H.oracle_scene(desc)
        if accel == "BVH":
            orc = O.BVH(osc, nodes=sess.bvh_nodes())
        else:
            orc = O.MBVH(osc)       # oracle's own trees (CLASSIC): results are topology independent
        for kind in kinds:
            if kind == "camera":
                side = int(n ** 0.5)
                rays = R.camera_rays(desc.cam, side, side, seed=1, device=dev, time_range=time_range)
            else:
                depth = int(kind.split("-")[1])
                rays = B.make_bounce_batch(trace_fn, desc, n, seed=2, device=dev, depth=depth)
            m = rays.shape[0]
            hits = torch.empty((m, 20), dtype=torch.uint8, device=dev)
            for _ in range(3):
                sess.trace_device(rays.data_ptr(), hits.data_ptr(), m)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                sess.trace_device(rays.data_ptr(), hits.data_ptr(), m)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            st = scene.trace_stats(rays.data_ptr(), 0, m)
            k = min(m, sample)
            rn = R.to_numpy_rays(rays[:k])
            ref = orc.intersect(rn)
            got = hits[:k].cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
            try:
                rep = H.compare_hits(got, ref, rn, what=name)
                par = {k2: int(v) for k2, v in rep.items()}
            except AssertionError as e:
                par = {"FAILED": str(e)[:300]}
            row = {"scene": name, "accelerator": accel, "builder": builder, "rays": kind, "n": m, "ms": round(ms, 4),
                   "mrays_per_s": round(m / ms / 1e3, 1), "nodes_per_ray": round(st.wide_nodes / max(1, st.rays), 2),
                   "tris_per_ray": round(st.triangles / max(1, st.rays), 2), "instances_per_ray": round(st.instances / max(1, st.rays), 2),
                   "triangles": int(info.n_triangles), "wide_nodes": int(info.n_wide_nodes), "instances": int(info.n_instances),
                   "device_MB": round(info.device_bytes / 1e6, 1), "host_build_s": round(build_s, 2), "parity_sample": par}
            out.append(row); print(json.dumps(row), flush=True)
            del rays, hits
        sess.stop(); sess.close()
    except Exception as e:
        row = {"scene": name, "accelerator": accel, "error": repr(e)[:300]}
        out.append(row); print(json.dumps(row), flush=True)

M = 1 << 20
run("cornell", "BVH", "CLASSIC", ["camera", "bounce-1"], 1 * M)
run("luxball", "BVH", "EMBREE_BINNED_SAH", ["camera", "bounce-4"], 4 * M)
run("bigmonkey", "BVH", "EMBREE_BINNED_SAH", ["camera", "bounce-4"], 4 * M)
run("lightinstances", "MBVH", "EMBREE_BINNED_SAH", ["camera", "bounce-1"], 4 * M)
run("bigmonkey-instances", "MBVH", "EMBREE_BINNED_SAH", ["camera", "bounce-1"], 4 * M)
run("bigmonkey-motion", "MBVH", "EMBREE_BINNED_SAH", ["camera"], 4 * M, time_range=(0.0, 1.0))
run("classroom", "BVH", "EMBREE_BINNED_SAH", ["bounce-2"], 16 * M)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
suffix = ("_" + "_".join(ARGS.only) if ARGS.only else "") + ("_" + "_".join(o.replace("=", "-") for o in ARGS.opt) if ARGS.opt else "")
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_table%s.json" % suffix), "w"), indent=1)
