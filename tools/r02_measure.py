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
ap.add_argument("--quick", action="store_true", help="default options only (A/B of differently compiled libraries, LRB_LIB_DIR)")
ap.add_argument("--opt", action="append", default=[], help="device option key=value applied to every variant")
ap.add_argument("--tag", default="")
ap.add_argument("--prefetch-modes", action="store_true", help="soup: sweep of what the prefetching kernel fetches ahead (device option prefetch_mode)")
ARGS = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
rows = []
DEFAULTS0 = {"smem_depth": 16, "refill_below": 24, "tri_bias": 8, "inst_bias": 8, "sort_rays": 2, "prefetch": 0, "prefetch_mode": 1,
            "blocks_per_sm": 0, "carveout": -1, "sort_bits": 5, "kernel": "persistent"}
DEFAULTS = dict(DEFAULTS0)
for kv in ARGS.opt:
    k, v = kv.split("=", 1)
    DEFAULTS[k] = v


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
    if ARGS.quick:
        variants = variants[:1]
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
        if ARGS.quick and name != "kitchen":
            sess.stop(); sess.close()
            continue
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


def section_ingest():
    """What the gathering GPU of an 8-GPU run sees: 7 x 336 MB of RayHit records arrive per step while its own kernel
    traces an L2-resident scene.  Emulated on ONE GPU: a copy stream streams the same amount through the device
    (device-to-device copies by the copy engine) while the trace kernel runs; with and without the L2 persistence
    window over the scene (device option l2_persist)."""
    desc = S.load_fixture("kitchen")
    sess = open_session(desc, "BVH")
    n = ARGS.rays or (16 << 20)
    rays = B.make_bounce_batch(trace_fn_of(sess), desc, n, seed=2, device=dev, depth=2)
    hits = torch.empty((n, 20), dtype=torch.uint8, device=dev)
    src = torch.empty(7 * n * 20, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    copy_stream = torch.cuda.Stream()
    for persist in (0, 2, 1):
        sess.set_option("l2_persist", persist)
        for ingest in (False, True):
            for _ in range(2):
                sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
            torch.cuda.synchronize()
            ts = []
            for _ in range(ARGS.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if ingest:
                    with torch.cuda.stream(copy_stream):
                        dst.copy_(src, non_blocking=True)
                a.record(); sess.trace_device(rays.data_ptr(), hits.data_ptr(), n); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            emit({"scene": "kitchen", "rays": "bounce-2", "n": n, "l2_persist": persist, "concurrent_copy_of_2.35GB": ingest,
                  "ms": round(ts[len(ts) // 2], 4), "ms_best": round(ts[0], 4)})
    sess.set_option("l2_persist", 2)
    sess.stop(); sess.close()


def section_mbvh():
    for name, kinds, tr in (("lightinstances", ["camera", "bounce-1"], None), ("bigmonkey-instances", ["bounce-1"], None),
                            ("bigmonkey-motion", ["camera"], (0.0, 1.0))):
        if ARGS.quick and name != "lightinstances":
            continue
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
        if name == "lightinstances" and not ARGS.quick:
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
    if ARGS.quick:
        v = [("sorted 5 bits (what sort_rays = auto picks for a scene larger than L2)", {"sort_rays": 1, "sort_bits": 5})]
    if ARGS.prefetch_modes:
        v = [("sorted 5 bits", {"sort_rays": 1, "sort_bits": 5})]
        v += [("sorted 5 bits + prefetch mode %d" % m, {"sort_rays": 1, "sort_bits": 5, "prefetch": 1, "prefetch_mode": m}) for m in (2, 4, 6, 3, 1)]
        v += [("index order + prefetch mode %d" % m, {"sort_rays": 0, "prefetch": 1, "prefetch_mode": m}) for m in (2, 4)]
        v += [("sorted 5 bits again", {"sort_rays": 1, "sort_bits": 5})]
    sweep(sess, "soup:%d" % n_tris, "uniform", rays, v, orc, sample=20000)
    # probes for the roofline denominators, same process / same clocks
    dv = capi.Device.borrow(sess.native_device())
    emit({"probe": {"l2_read_gbs_32MiB": round(dv.measure_read_bandwidth(32 << 20, 50), 1),
                    "l2_read_gbs_64MiB": round(dv.measure_read_bandwidth(64 << 20, 30), 1),
                    "hbm_read_gbs_4GiB": round(dv.measure_read_bandwidth(4 << 30, 5), 1)}})
    sess.stop(); sess.close()


def section_masked():
    """f1: dead lanes between bounces -- masked rays skipped inside the trace kernel vs compacted before it, and
    f3: shadow rays through the any-hit kernels vs the closest-hit kernels on the same batch."""
    desc = S.load_fixture("kitchen")
    sess = open_session(desc, "BVH")
    scene = sess.native_scene()
    n = ARGS.rays or (16 << 20)
    rays = B.make_bounce_batch(trace_fn_of(sess), desc, n, seed=2, device=dev, depth=2)
    hits = torch.empty((n, 20), dtype=torch.uint8, device=dev)
    flags = rays.view(torch.int32)[:, 9]
    g = torch.Generator(device=dev); g.manual_seed(3)
    u = torch.rand(n, generator=g, device=dev)
    for frac in (0.0, 0.3, 0.5, 0.7, 0.9):
        flags.copy_((u < frac).to(torch.int32))
        row = {"scene": "kitchen", "rays": "bounce-2", "n": n, "masked_fraction": frac}
        ref = None
        for label, opt in (("skip_in_kernel", 0), ("compacted", 1), ("auto", 2)):
            sess.set_option("compact", opt)
            hits.zero_()
            med, best = time_trace(sess, rays, hits, ARGS.reps)
            row[label + "_ms"] = round(med, 4)
            row[label + "_live_mrays_per_s"] = round(n * (1.0 - frac) / med / 1e3, 1)
            if ref is None:
                ref = hits.clone()
            else:
                row["identical"] = row.get("identical", True) and bool(torch.equal(ref, hits))
        sess.set_option("compact", 0)
        emit(row)
    flags.zero_()
    # shadow rays: from the surface points of the bounce batch towards uniform points of the scene box (d = target - o,
    # maxt = 1 - 1e-4: the segment up to the target, the way the reference builds its shadow rays)
    lo, hi = desc.bbox()
    lo_t, hi_t = torch.from_numpy(lo).to(dev), torch.from_numpy(hi).to(dev)
    rf = rays.view(torch.float32).clone()
    target = lo_t + (hi_t - lo_t) * torch.rand((n, 3), generator=g, device=dev)
    rf[:, 3:6] = target - rf[:, 0:3]
    rf[:, 7] = 1.0 - 1e-4
    shadow = rf.view(torch.uint8).reshape(n, 48).contiguous()
    def run(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ARGS.reps)]
        for a, b in ev:
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return sorted(a.elapsed_time(b) for a, b in ev)[len(ev) // 2]
    for label, batch in (("bounce-2 (infinite maxt)", rays), ("shadow segments", shadow)):
        h1 = torch.empty((n, 20), dtype=torch.uint8, device=dev); h2 = torch.empty((n, 20), dtype=torch.uint8, device=dev)
        ms_c = run(lambda: scene.trace(batch.data_ptr(), h1.data_ptr(), n))
        ms_a = run(lambda: scene.trace_anyhit(batch.data_ptr(), h2.data_ptr(), n))
        m1 = h1.view(torch.int32)[:, 3] == -1; m2 = h2.view(torch.int32)[:, 3] == -1
        emit({"scene": "kitchen", "rays": label, "n": n, "closest_ms": round(ms_c, 4), "closest_mrays_per_s": round(n / ms_c / 1e3, 1),
              "anyhit_ms": round(ms_a, 4), "anyhit_mrays_per_s": round(n / ms_a / 1e3, 1), "hit_fraction": round(1.0 - float(m1.float().mean()), 4),
              "hit_miss_identical": bool(torch.equal(m1, m2))})
    sess.stop(); sess.close()


sec = ARGS.section
if sec == "masked":
    section_masked()
elif sec == "ingest":
    section_ingest()
elif sec == "kitchen":
    section_kitchen()
elif sec == "mbvh":
    section_mbvh()
elif sec.startswith("soup"):
    section_soup(int(sec.split(":")[1]) if ":" in sec else 50000000)
else:
    raise SystemExit("unknown section")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_measure_%s%s.json" % (sec.replace(":", "_"), ARGS.tag)), "w"), indent=1)
