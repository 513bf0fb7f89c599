#!/usr/bin/env python3
"""GPU probe: random-soup BVH that does not fit L2 -- ray order (index order vs. sorted) against kernel time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from luxcore_b200 import hostapi, rays as R, scenes as S
n_tris = int(sys.argv[1]) if len(sys.argv) > 1 else 16000000
n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 4194304
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
desc = S.random_soup(n_tris, seed=4, size=0.002 * (50e6 / n_tris) ** (1.0 / 3.0))
sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
sess.build_accelerator("BVH"); sess.start(0); sess.set_stream(stream.cuda_stream)
info = sess.native_scene().info()
print("soup", n_tris, "device MB", info.device_bytes / 1e6, "stack_need", info.stack_need, flush=True)
rays = R.uniform_rays([0, 0, 0], [1, 1, 1], n_rays, seed=5, device=dev)
hits = torch.empty((n_rays, 20), dtype=torch.uint8, device=dev)
ref = None
variants = [("index order", {"sort_rays": "0"}), ("sorted 6 bits/axis", {"sort_rays": "1", "sort_bits": "6"}),
            ("sorted 8 bits/axis", {"sort_rays": "1", "sort_bits": "8"})]
if len(sys.argv) > 3:
    variants = [variants[int(i)] for i in sys.argv[3].split(",")]
for label, opts in variants:
    for k, v in opts.items():
        sess.set_option(k, v)
    sess.trace_device(rays.data_ptr(), hits.data_ptr(), n_rays); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.trace_device(rays.data_ptr(), hits.data_ptr(), n_rays); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = None
    if ref is None: ref = hits.clone()
    else: same = bool(torch.equal(ref, hits))
    print("%-20s %10.2f ms  %9.1f Mrays/s  identical_to_index_order=%s" % (label, ms, n_rays / ms / 1e3, same), flush=True)
