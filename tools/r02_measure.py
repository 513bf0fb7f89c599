#!/usr/bin/env python3
"""Round-2 measurement tool (GPU box): option sweeps of the traversal kernels inside ONE process per
scene, so that a 50 M-triangle soup is built once for all its variants.  CUDA-event timings of
lrb_trace over device-resident batches (like bench.py's `value`); every variant is compared byte for
byte with the first one, and a sample of each batch with the oracle.

    python tools/r02_measure.py kitchen | mbvh | soup[:NTRIS] [--rays N]

Writes gpurun_out/r02_measure_<section>.json.  Development aid, not a bench line."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
import bench as B
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("section")
ap.add_argument("--rays", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--no-parity", action="store_true")
ARGS = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rows = []
DEFAULTS = {"smem_depth": 16, "refill_below": 24, "tri_bias": 8, "inst_bias": 8, "sort_rays": 2, "prefetch": 0,
            "blocks_per_sm": 0, "carveout": -1, "sort_bits": 5, "kernel": "persistent"}


def emit(row):
    rows.append(row)
    print(json.dumps(row), flush=True)


def time_trace(sess, rays, hits, reps):
    m = rays.shape[0]
    for _ in range(2):
        sess.trace_device(rays.data_ptr(), hits.data_ptr(), m)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); sess.trace_device(rays.data_ptr(), hits.data_ptr(), m); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def sweep(sess, name, kind, rays, variants, orc=None, sample=50000):
    m = rays.shape[0]
    hits = torch.empty((m, 20), dtype=torch.uint8, device=dev)
    ref = None
    scene = sess.native_scene()
    st = scene.trace_stats(rays.data_ptr(), 0, m)
    base = {"scene": name, "rays": kind, "n": m, "nodes_per_ray": round(st.wide_nodes / max(1, st.rays), 2),
            "tris_per_ray": round(st.triangles / max(1, st.rays), 2), "instances_per_ray": round(st.instances / max(1, st.rays), 3),
            "max_stack": int(st.max_stack)}
    for label, opts in variants:
        full = dict(DEFAULTS); full.update(opts)
        try:
            for k, v in full.items():
                sess.set_option(k, v)
            hits.zero_()
            med, best = time_trace(sess, rays, hits, ARGS.reps)
            same = None
            if ref is None:
                ref = hits.clone()
                if orc is not None and not ARGS.no_parity:
                    k = min(m, sample)
                    rn = R.to_numpy_rays(rays[:k])
                    want = orc.intersect(rn)
                    got = hits[:k].cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
                    try:
                        rep = H.compare_hits(got, want, rn, what=name)
                        base["parity_sample"] = {k2: int(v) for k2, v in rep.items()}
                    except AssertionError as e:
                        base["parity_sample"] = {"FAILED": str(e)[:300]}
            else:
                same = bool(torch.equal(ref, hits))
            row = dict(base); row.update({"variant": label, "opts": opts, "ms": round(med, 4), "ms_best": round(best, 4),
                                          "mrays_per_s": round(m / med / 1e3, 1), "identical_to_first": same})
        except Exception as e:
            row = dict(base); row.update({"variant": label, "opts": opts, "error": repr(e)[:300]})
        emit(row)
    for k, v in DEFAULTS.items():
        sess.set_option(k, v)
    del hits, ref


def open_session(desc, accel, builder="EMBREE_BINNED_SAH"):
    t0 = time.perf_counter()
    sess = hostapi.Session({"accelerator.type": accel, "accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator(accel)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    sess.start(0); sess.set_stream(stream.cuda_stream)
    upload_s = time.perf_counter() - t0
    info = sess.native_scene().info()
    emit({"scene": desc.name if hasattr(desc, "name") else "?", "accelerator": accel, "host_build_s": round(build_s, 2),
          "start_upload_relayout_s": round(upload_s, 2), "triangles": int(info.n_triangles), "wide_nodes": int(info.n_wide_nodes),
          "instances": int(info.n_instances), "device_MB": round(info.device_bytes / 1e6, 1), "stack_need": int(info.stack_need)})
    return sess


def trace_fn_of(sess):
    def trace_fn(r):
        h = torch.empty((r.shape[0], 20), dtype=torch.uint8, device=dev)
        sess.trace_device(r.data_ptr(), h.data_ptr(), r.shape[0]); return h
    return trace_fn


def section_kitchen():
    for name in ("kitchen", "classroom"):
        desc = S.load_fixture(name)
        sess = open_session(desc, "BVH")
        n = ARGS.rays or (16 << 20)
        rays = B.make_bounce_batch(trace_fn_of(sess), desc, n, seed=2, device=dev, depth=2)
        orc = O.BVH(H.oracle_scene(desc), nodes=sess.bvh_nodes())
        if name == "kitchen":
            v = [("default", {})]
            v += [("smem_depth=%d" % d, {"smem_depth": d}) for d in (8, 12, 24)]
            v += [("refill_below=%d" % d, {"refill_below": d}) for d in (16, 20, 28)]
            v += [("tri_bias=%d" % d, {"tri_bias": d}) for d in (4, 6, 12)]
            v += [("blocks_per_sm=%d" % d, {"blocks_per_sm": d}) for d in (4, 6, 7)]
            v += [("carveout=%d" % d, {"carveout": d}) for d in (25, 50, 75, 100)]
            v += [("smem8+carveout=%d" % d, {"smem_depth": 8, "carveout": d}) for d in (25, 50)]
            v += [("sort_rays=1 bits=%d" % b, {"sort_rays": 1, "sort_bits": b}) for b in (4, 5, 6)]
            v += [("prefetch=1", {"prefetch": 1}), ("kernel=simple", {"kernel": "simple"}), ("default again", {})]
        else:
            v = [("default", {}), ("sort_rays=1", {"sort_rays": 1})]
        sweep(sess, name, "bounce-2", rays, v, orc)
        del rays
        sess.stop(); sess.close()


def section_mbvh():
    for name, kinds, tr in (("lightinstances", ["camera", "bounce-1"], None), ("bigmonkey-instances", ["bounce-1"], None),
                            ("bigmonkey-motion", ["camera"], (0.0, 1.0))):
        desc = S.load_fixture(name)
        sess = open_session(desc, "MBVH")
        orc = O.MBVH(H.oracle_scene(desc))
        n = ARGS.rays or (4 << 20)
        for kind in kinds:
            if kind == "camera":
                side = int(n ** 0.5)
                rays = R.camera_rays(desc.cam, side, side, seed=1, device=dev, time_range=tr)
            else:
                rays = B.make_bounce_batch(trace_fn_of(sess), desc, n, seed=2, device=dev, depth=int(kind.split("-")[1]))
            v = [("default", {})]
            if name == "lightinstances":
                v += [("inst_bias=%d" % d, {"inst_bias": d}) for d in (0, 2, 4, 16)]
                v += [("refill_below=%d" % d, {"refill_below": d}) for d in (16, 20, 28)]
                v += [("tri_bias=%d" % d, {"tri_bias": d}) for d in (4, 12)]
                v += [("smem_depth=%d" % d, {"smem_depth": d}) for d in (8, 24)]
                v += [("blocks_per_sm=%d" % d, {"blocks_per_sm": d}) for d in (4, 5)]
                v += [("sort_rays=1", {"sort_rays": 1}), ("kernel=simple", {"kernel": "simple"})]
            sweep(sess, name, kind, rays, v, orc, sample=30000)
            del rays
        # Update() cost: move every instance a little, time the host-side refit + re-layout + upload
        if name == "lightinstances":
            inst = [i for i, m in enumerate(desc.meshes) if m.kind == S.INSTANCE]
            t0 = time.perf_counter()
            for i in inst:
                m = np.array(desc.meshes[i].xform, dtype=np.float32).reshape(4, 4).copy()
                m[0, 3] += 0.01
                sess.set_instance_transform(i, m)
            t1 = time.perf_counter()
            sess.update(); torch.cuda.synchronize()
            t2 = time.perf_counter()
            emit({"scene": name, "update": {"instances_moved": len(inst), "set_transform_s": round(t1 - t0, 4), "update_s": round(t2 - t1, 4)}})
        sess.stop(); sess.close()


def section_soup(n_tris):
    desc = S.random_soup(n_tris, seed=4, size=0.002 * (50e6 / n_tris) ** (1.0 / 3.0), name="soup")
    sess = open_session(desc, "BVH")
    n = ARGS.rays or (32 << 20)
    # guard: one 2 Mi-ray launch first; a pathologically slow kernel must not eat the GPU budget
    probe = R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 2 << 20, seed=9, device=dev)
    ph = torch.empty((probe.shape[0], 20), dtype=torch.uint8, device=dev)
    sess.trace_device(probe.data_ptr(), ph.data_ptr(), probe.shape[0]); torch.cuda.synchronize()
    t0 = time.perf_counter(); sess.trace_device(probe.data_ptr(), ph.data_ptr(), probe.shape[0]); torch.cuda.synchronize()
    rate = probe.shape[0] / (time.perf_counter() - t0) / 1e6
    emit({"scene": "soup:%d" % n_tris, "probe_2Mi_mrays_per_s": round(rate, 1)})
    if rate < 100.0:
        n = min(n, 4 << 20); ARGS.reps = 2
    del probe, ph
    rays = R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], n, seed=7, device=dev)
    orc = None
    if not ARGS.no_parity:
        orc = O.BVH(H.oracle_scene(desc), nodes=sess.bvh_nodes())
    v = [("index order", {"sort_rays": 0}), ("sorted 5 bits", {"sort_rays": 1, "sort_bits": 5}), ("sorted 7 bits", {"sort_rays": 1, "sort_bits": 7}),
         ("sorted 9 bits", {"sort_rays": 1, "sort_bits": 9}),
         ("index order + prefetch", {"sort_rays": 0, "prefetch": 1}), ("sorted 7 bits + prefetch", {"sort_rays": 1, "sort_bits": 7, "prefetch": 1}),
         ("sorted 7 bits, smem_depth 8", {"sort_rays": 1, "sort_bits": 7, "smem_depth": 8}),
         ("sorted 7 bits, refill 16", {"sort_rays": 1, "sort_bits": 7, "refill_below": 16}),
         ("sorted 7 bits, carveout 25", {"sort_rays": 1, "sort_bits": 7, "carveout": 25}),
         ("sorted 7 bits, simple kernel", {"sort_rays": 1, "sort_bits": 7, "kernel": "simple"})]
    sweep(sess, "soup:%d" % n_tris, "uniform", rays, v, orc, sample=20000)
    # probes for the roofline denominators, same process / same clocks
    dv = capi.Device.borrow(sess.native_device())
    emit({"probe": {"l2_read_gbs_32MiB": round(dv.measure_read_bandwidth(32 << 20, 50), 1),
                    "l2_read_gbs_64MiB": round(dv.measure_read_bandwidth(64 << 20, 30), 1),
                    "hbm_read_gbs_4GiB": round(dv.measure_read_bandwidth(4 << 30, 5), 1)}})
    sess.stop(); sess.close()


sec = ARGS.section
if sec == "kitchen":
    section_kitchen()
elif sec == "mbvh":
    section_mbvh()
elif sec.startswith("soup"):
    section_soup(int(sec.split(":")[1]) if ":" in sec else 50000000)
else:
    raise SystemExit("unknown section")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_measure_%s.json" % sec.replace(":", "_")), "w"), indent=1)
