"""CPU-only probe of tree quality: wide-node visits / triangle tests per ray of the product's
re-layout + traversal body (tests/cpp/wide_emulation.cpp = traverse.h compiled for the host) on the
bench workload's ray kind.  The kernel is issue-bound on L2-resident scenes (profiles/), so visits per
ray are the first-order cost; this tool measures them without a GPU.

    python tools/tree_stats.py [scene] [n_rays] [depth] [builder]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import bench as B
import helpers as H
from luxcore_b200 import hostapi, rays as R
from oracle import oracle as O


def bounce_batch_cpu(desc, nodes, n, depth, seed=2):
    _, bvh = B.oracle_for(desc, nodes)
    th = O.hardware_threads()

    def trace_fn(rays_u8):
        h = bvh.intersect(R.to_numpy_rays(rays_u8), nthreads=th)
        return torch.from_numpy(h.view(np.uint8).reshape(-1, 20).copy())
    return R.to_numpy_rays(B.make_bounce_batch(trace_fn, desc, n, seed=seed, device="cpu", depth=depth)), bvh


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    builder = sys.argv[4] if len(sys.argv) > 4 else "EMBREE_BINNED_SAH"
    desc = B.build_scene_arrays(scene)
    t0 = time.perf_counter()
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    nodes = sess.bvh_nodes()
    t_build = time.perf_counter() - t0
    rays, obvh = bounce_batch_cpu(desc, nodes, n, depth)
    osc = H.oracle_scene(desc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    t0 = time.perf_counter()
    emu = H.Emu.bvh(nodes, verts, offs)
    t_relayout = time.perf_counter() - t0
    hits, st = emu.trace(rays, want_stats=True)
    ref = obvh.intersect(rays, nthreads=O.hardware_threads())
    rep = H.compare_hits(hits, ref, rays, what="tree_stats")
    info = emu.info()
    print("scene %s builder %s: ref nodes %d, wide nodes %d, tris %d, stack need %d, build %.2fs relayout %.2fs" % (
        scene, builder, nodes.shape[0], info["wide"], info["tris"], info["stack_need"], t_build, t_relayout))
    print("rays %d depth %d: wide nodes/ray %.3f  triangles/ray %.3f  max stack %d  parity %r" % (
        st["rays"], depth, st["wide_nodes"] / st["rays"], st["triangles"] / st["rays"], st["max_stack"], rep))


if __name__ == "__main__":
    main()
