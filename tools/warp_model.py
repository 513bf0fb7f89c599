"""CPU-only issue-slot model of the persistent trace kernel.

The kernel is issue-bound on L2-resident scenes (profiles/r02c6: issue-active 82 %, 276 warp
instructions per ray at 17.9 active threads).  tests/cpp/wide_emulation.cpp::WarpSim replays the
kernel's scheduling (bulk re-fill, Resolve, node / triangle phase vote) with the real traverse.h
bodies and counts the phases a warp issues; this tool weights them with the SASS instruction counts
of the sections of TracePersistent<0,1,0> (cuobjdump -sass) and prints warp instructions per ray --
a GPU-free figure of merit for comparing trees (builder, re-layout) and scheduling policies.

    python tools/warp_model.py [scene] [n_rays] [depth] [refill_below] [tri_bias]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np

import bench as B
import helpers as H
import tree_stats as TS
from luxcore_b200 import hostapi

# Warp instructions per EXECUTED section, from the ncu SOURCE view of the benched kernel (round 2, commit c8a58a5,
# profiles/r02c6_kitchen_ncu_summary.txt; per-SASS-instruction execution counts aggregated by section and divided by
# the executions of the section -- DESIGN.md section 4, item 6): node phase 122.7 warp instructions per ray over 0.73
# phases per ray, triangle phase + in-record gate 52.5 over 0.46, Resolve's loop 5 per trip, the rest per iteration.
# tools/sass_costs.py attributes the static SASS to source functions automatically -- run it after touching the kernel.
COST = {"inner_fixed": 55,      # Resolve entry / exit around the loop, __syncwarp, two ballots, vote, loop branch
        "pop_trip": 5,          # one trip of Resolve's loop (longest lane decides): LDS, two address updates, compare, branch
        "node_phase": 168,      # fetch + decode + 4 slab tests + network + predicated 64-bit pushes
        "tri_phase": 114,       # fetch + Moller-Trumbore + the gate box from the same record + commit
        "gate": 0,              # (round 1: a separate side path per accepted hit; now part of the triangle phase)
        "outer_fixed": 10,      # store / re-fill section when nothing to do
        "store": 40,            # RayHit stores of the finished lanes
        "refill": 150}          # atomic, ray fetch, 1/d, root box


def model(c):
    instr = (c["inner_iters"] * COST["inner_fixed"] + c["pop_trips"] * COST["pop_trip"] + c["node_phases"] * COST["node_phase"] +
             c["tri_phases"] * COST["tri_phase"] + c["gate_phases"] * COST["gate"] + c["outer_iters"] * COST["outer_fixed"] +
             c["store_phases"] * COST["store"] + c["refills"] * COST["refill"])
    lanes = (c["node_lanes"] * COST["node_phase"] + c["tri_lanes"] * COST["tri_phase"] + c["gate_lanes"] * COST["gate"] +
             c["pop_lanes"] * COST["pop_trip"])
    return instr, lanes


ENTER_INSTANCE = 110    # EnterInstance: record + matrix fetch, ray transform, three IEEE reciprocals, sentinel push

MBVH_SCENES = ("lightinstances", "bigmonkey-instances", "bigmonkey-motion")


def main_two_level(scene, n, refill, bias):
    """Two-level scenes: host-layer SAH trees (root + leaves) inside the oracle's array set, camera and
    first-bounce rays, instance entry as a voted phase (inst_bias sweep)."""
    import torch
    from luxcore_b200 import rays as R, scenes as S
    from oracle import oracle as O
    desc = S.load_fixture(scene)
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=4)
    arr = H.mbvh_arrays(desc, mb)
    s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    s.build_accelerator("MBVH")
    arr["root_nodes"] = s.mbvh_root_nodes().copy()
    for i in range(s.mbvh_leaf_count()):
        arr["leaf_nodes"][i] = s.mbvh_leaf_nodes(i).copy()
    emu = H.Emu.mbvh(arr)
    th = O.hardware_threads()
    time_range = (0.0, 1.0) if "motion" in scene else None

    def trace_fn(rays_u8):
        h = mb.intersect(R.to_numpy_rays(rays_u8), nthreads=th)
        return torch.from_numpy(h.view(np.uint8).reshape(-1, 20).copy())
    for kind, depth in (("camera", 0), ("bounce-1", 1)):
        if time_range is not None:
            if depth:
                continue
            side = int(n ** 0.5)
            rays = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=1, time_range=time_range))
        else:
            rays = R.to_numpy_rays(B.make_bounce_batch(trace_fn, desc, n, seed=2, device="cpu", depth=depth))
        ref, st = emu.trace(rays, want_stats=True)
        r = st["rays"]
        print("scene %s %s: per ray nodes %.2f tris %.2f instance entries %.2f motion samples %.2f" % (
            scene, kind, st["wide_nodes"] / r, st["triangles"] / r, st["instances"] / r, st["motion_samples"] / r))
        for ib in (0, 4, 8, 16):
            hits, c = H.warp_sim(emu, rays, n_warps=128, refill_below=refill, tri_bias=bias, inst_bias=ib)
            assert hits.tobytes() == ref.tobytes(), "scheduling model and per-ray emulation disagree"
            instr, _ = model(c)
            enter = ENTER_INSTANCE * c["instance_trips"] / r
            print("  inst_bias %2d: MODEL warp instructions / ray %.1f (+ %.1f entering instances = %.1f) | entry executions %.3f per ray at %.1f lanes"
                  " | iterations in which a lane leaves an instance %.3f per ray (%.1f lanes)" % (
                      ib, instr / r, enter, instr / r + enter, c["instance_trips"] / r, c["instance_lanes"] / max(1, c["instance_trips"]),
                      c["leave_trips"] / r, c["leave_lanes"] / max(1, c["leave_trips"])))


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
    if scene in MBVH_SCENES:
        return main_two_level(scene, int(sys.argv[2]) if len(sys.argv) > 2 else 100000,
                              int(sys.argv[4]) if len(sys.argv) > 4 else 24, int(sys.argv[5]) if len(sys.argv) > 5 else 8)
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
    depth = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    refill = int(sys.argv[4]) if len(sys.argv) > 4 else 24
    bias = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    desc = B.build_scene_arrays(scene)
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    nodes = sess.bvh_nodes()
    rays, obvh = TS.bounce_batch_cpu(desc, nodes, n, depth)
    osc = H.oracle_scene(desc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    ref, st = emu.trace(rays, want_stats=True)
    spec = os.environ.get("LRB_MODEL_SPEC_POP", "1") != "0"
    hits, c = H.warp_sim(emu, rays, n_warps=256, refill_below=refill, tri_bias=bias, spec_pop=spec, tri_min=int(os.environ.get("LRB_MODEL_TRI_MIN", "0")))
    assert hits.tobytes() == ref.tobytes(), "scheduling model and per-ray emulation disagree"
    instr, lanes = model(c)
    print("speculative pop in the phases:", spec, {k: round(v / c["rays"], 3) for k, v in c.items()})
    r = c["rays"]
    print("scene %s: ref nodes %d wide %d | per ray: nodes %.2f tris %.2f | node phases %.2f (%.1f lanes) tri phases %.2f (%.1f lanes) "
          "pop trips %.2f inner %.2f" % (scene, nodes.shape[0], emu.info()["wide"], st["wide_nodes"] / r, st["triangles"] / r,
                                       c["node_phases"] / r, c["node_lanes"] / max(1, c["node_phases"]), c["tri_phases"] / r,
                                       c["tri_lanes"] / max(1, c["tri_phases"]), c["pop_trips"] / r, c["inner_iters"] / r))
    print("MODEL warp instructions / ray: %.1f   (lane-useful fraction of phase work %.2f; refill_below %d tri_bias %d)" % (
        instr / r, lanes / (32.0 * max(1, instr)), refill, bias))


if __name__ == "__main__":
    main()
