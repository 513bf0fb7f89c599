#!/usr/bin/env python3
"""bench.py -- closest-hit Mrays/s of the B200 intersection device (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU

Workload (config.workload = "kitchen-16M-bounce"): BASELINE.json configs[3] -- the reference's
`scenes/kitchen` interior (86 032 triangles, committed as tests/golden/scenes/kitchen.npz), BVH
accelerator built by the product's host layer with the reference's defaults (builder
EMBREE_BINNED_SAH -> this tree's binned-SAH builder, 4-ary, one triangle per leaf), and batches of
16 Mi INCOHERENT rays: second-bounce diffuse path rays generated on the GPU the way the reference's
path tracer emits them (camera ray -> hit -> cosine-weighted bounce -> hit -> bounce).
One step = one EnqueueTraceRayBuffer over one batch per GPU (weak scaling: every rank owns a 16 Mi
batch of its own seed; the BVH is replicated); with N > 1 the step also gathers the RayHit buffers
onto rank 0 (the only exchange the path has).

The JSON line follows the driver contract; see DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

RAYS_PER_BATCH = 16 * 1024 * 1024
METRIC = "closest_hit_mrays_per_s"
UNIT = "Mrays/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the traversal kernel on this workload,
    from the committed `ncu --set full` capture (profiles/traffic.json); None if no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(workload)
        if e:
            return float(e["dram_bytes_per_launch"]), e.get("source")
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """Summary of the samples that arrived inside [t_begin, t_end] (the timed region); when the region
        was shorter than the sampling period, the samples closest to it."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        lines = list(self.lines)
        if t_begin is not None and t_end is not None and lines:
            inside = [l for l in lines if t_begin - 0.005 <= l[0] <= t_end + 0.03]
            if not inside:
                mid = 0.5 * (t_begin + t_end)
                inside = sorted(lines, key=lambda l: abs(l[0] - mid))[:2]
            lines = inside
        for _, line in lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------

def is_soup(scene_name):
    return scene_name.startswith("soup")


def build_scene_arrays(scene_name):
    """A committed fixture of one of the reference's scenes, or `soup[:NTRIS]` = BASELINE.json configs[4]:
    the synthetic random triangle soup (50 M triangles by default, size scaled so
    that the triangle density per unit volume stays the one of the 50 M / 0.002 soup)."""
    from luxcore_b200 import scenes as S
    if is_soup(scene_name):
        n_tris = int(scene_name.split(":")[1]) if ":" in scene_name else 50000000
        return S.random_soup(n_tris, seed=4, size=0.002 * (50e6 / n_tris) ** (1.0 / 3.0), name="soup")
    return S.load_fixture(scene_name)


def make_batch(trace_fn, desc, args, n_rays, seed, device):
    """The ray batch of the workload: incoherent bounce rays for the reference's scenes, uniform
    origins x uniform directions inside the unit cube for the soup."""
    if is_soup(args.scene):
        from luxcore_b200 import rays as R
        return R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], n_rays, seed=5 + seed, device=device)
    return make_bounce_batch(trace_fn, desc, n_rays, seed=seed, device=device, depth=args.depth)


def make_bounce_batch(trace_fn, desc, n_rays, seed, device, depth=2):
    """n_rays rays at path depth `depth` (depth 0 = camera rays).  trace_fn(rays_u8) -> hits_u8."""
    from luxcore_b200 import rays as R, scenes as S
    p0, e1, e2, offs = S.world_triangles(desc)
    p0 = torch.from_numpy(p0).to(device); e1 = torch.from_numpy(e1).to(device); e2 = torch.from_numpy(e2).to(device)
    offs = torch.from_numpy(offs).to(device)
    out, have, attempt = [], 0, 0
    while have < n_rays:
        side = int(math.ceil(math.sqrt((n_rays - have) * 1.15))) + 8
        rays = R.camera_rays(desc.cam, side, side, seed=seed * 1000 + attempt, device=device)
        for b in range(depth):
            hits = trace_fn(rays)
            h = R.unpack_hits(hits)
            hit = h["mesh"] != -1
            flat = torch.where(hit, offs[h["mesh"].clamp(min=0).long()] + h["tri"].clamp(min=0).long(), torch.zeros_like(h["tri"], dtype=torch.long))
            rays, _ = R.bounce_rays(rays, hits, p0[flat], e1[flat], e2[flat], seed=seed * 1000 + attempt * 10 + b + 1)
            del hits, h, flat
        out.append(rays)
        have += rays.shape[0]
        attempt += 1
        if attempt > 8:
            break
    rays = torch.cat(out)[:n_rays].contiguous()
    if rays.shape[0] < n_rays:      # pathological scene: pad by repetition
        reps = int(math.ceil(n_rays / max(1, rays.shape[0])))
        rays = rays.repeat(reps, 1)[:n_rays].contiguous()
    return rays


def oracle_for(desc, nodes):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    from oracle import oracle as O
    osc = H.oracle_scene(desc)
    return O, O.BVH(osc, nodes=nodes)


def reference_for(desc, nodes):
    """The reference's own BVHAccel (oracle/_ref: its C++ sources compiled from /root/reference) walking the
    same BVHArrayNode array, or None where the prebuilt library is absent (then the oracle port is used)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import helpers as H
        from oracle import refapi as RF
        if not RF.available():
            return None
        return RF.BVH(H.reference_scene(desc), nodes=nodes)
    except Exception as e:      # a broken checker must not take the benchmark down
        log("reference library unavailable:", repr(e))
        return None


def cpu_time_sample(O, bvh, rays_np, target_s, threads):
    """Times the oracle (reference algorithm, CPU) on a bounded prefix of the batch."""
    probe = min(rays_np.shape[0], 200000)
    t0 = time.perf_counter(); bvh.intersect(rays_np[:probe], nthreads=threads); dt = time.perf_counter() - t0
    rate = probe / max(dt, 1e-9)
    n = int(min(rays_np.shape[0], max(probe, rate * target_s)))
    t0 = time.perf_counter(); bvh.intersect(rays_np[:n], nthreads=threads); dt = time.perf_counter() - t0
    return n, dt


# ---------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="kitchen")
    ap.add_argument("--rays", type=int, default=RAYS_PER_BATCH, help="rays per batch per GPU")
    ap.add_argument("--depth", type=int, default=2, help="bounce depth of the ray batch")
    ap.add_argument("--builder", default="EMBREE_BINNED_SAH")
    ap.add_argument("--gather", default="p2p", choices=["nccl", "p2p", "direct", "none"])
    ap.add_argument("--chunks", type=int, default=0, help="0 = fused in-kernel push; >= 1 = launches per batch for the copy-engine push")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="device option key=value")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    kind = "uniform" if is_soup(args.scene) else "bounce%d" % args.depth
    workload = "%s-%dM-%s" % (args.scene, args.rays >> 20, kind) if args.rays >= (1 << 20) else "%s-%d-%s" % (args.scene, args.rays, kind)
    config = {"workload": workload,
              "scene": ("synthetic random triangle soup, seed 4 (BASELINE.json configs[4])" if is_soup(args.scene)
                        else "scenes/%s (fixture tests/golden/scenes/%s.npz)" % (args.scene, args.scene)),
              "accelerator": "BVH", "builder": args.builder, "treetype": 4, "rays_per_batch_per_gpu": args.rays,
              "ray_kind": ("incoherent: origins uniform in the unit cube, directions uniform on the sphere" if is_soup(args.scene)
                           else "incoherent diffuse bounce, path depth %d" % args.depth),
              "parallelism": "replicated BVH, one %d-ray batch per GPU x%d, RayHit gathered on rank 0 (%s)" % (args.rays, n_gpus, args.gather if n_gpus > 1 else "n/a"),
              "l2_policy": "inputs larger than L2 (48 B x rays + 20 B x rays per step >> 126 MB)"}

    if args.impl == "reference":
        return run_reference(args, rank, world, config)

    from luxcore_b200 import capi, hostapi, rays as R

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)

    desc = build_scene_arrays(args.scene)
    t0 = time.perf_counter()
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": args.builder, "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    build_s = time.perf_counter() - t0
    sess.start(local_rank)
    sess.set_stream(stream.cuda_stream)
    for kv in args.opt:
        k, v = kv.split("=", 1)
        sess.set_option(k, v)
    scene = sess.native_scene()
    info = scene.info()

    def trace_fn(rays_u8):
        hits = torch.empty((rays_u8.shape[0], 20), dtype=torch.uint8, device=device)
        sess.trace_device(rays_u8.data_ptr(), hits.data_ptr(), rays_u8.shape[0])
        return hits

    n = args.rays
    rays = make_batch(trace_fn, desc, args, n, seed=2 + rank, device=device)
    torch.cuda.synchronize()
    hits = torch.empty((n, 20), dtype=torch.uint8, device=device)

    # ---- multi-GPU: RayHit gather onto rank 0 ----
    #   p2p  : rank 0 owns the gather buffer; the other ranks map it over NVLink (CUDA IPC) and
    #          lrb_trace_gather pushes each traced chunk into it while the next chunk is traced
    #   nccl : trace, then torch.distributed gather (the baseline way)
    from luxcore_b200 import shard
    dev_view = capi.Device.borrow(sess.native_device())
    gather_mode = args.gather if world > 1 else "none"
    gbuf_local, gbuf_peer, my_dst, glist = 0, 0, 0, None
    if gather_mode in ("p2p", "direct"):
        import torch.distributed as dist
        if rank == 0:
            gbuf_local = dev_view.alloc(world * n * 20)
            handle = [dev_view.ipc_get_handle(gbuf_local)]
        else:
            handle = [None]
        dist.broadcast_object_list(handle, src=0)
        if rank == 0:
            my_dst = gbuf_local
        else:
            gbuf_peer = dev_view.ipc_open_handle(handle[0])
            my_dst = gbuf_peer + rank * n * 20
        flag = torch.zeros(1, dtype=torch.int32, device=device)
    elif gather_mode == "nccl":
        gathered = torch.empty((world * n, 20), dtype=torch.uint8, device=device) if rank == 0 else None
        glist = [gathered[i * n:(i + 1) * n] for i in range(world)] if rank == 0 else None

    def step():
        if gather_mode == "direct":
            import torch.distributed as dist
            # every rank's tracer lanes store their RayHit records straight into rank 0's buffer (NVLink stores)
            sess.trace_device(rays.data_ptr(), my_dst, n)
            dist.all_reduce(flag)
        elif gather_mode == "p2p":
            import torch.distributed as dist
            # rank 0 traces straight into its slice of the gather buffer (no copy at all)
            scene.trace_gather(rays.data_ptr(), my_dst if rank == 0 else hits.data_ptr(), n, my_dst, args.chunks)
            dist.all_reduce(flag)           # 4-byte "batch complete" signal, ordered after the pushes
        else:
            sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
            if gather_mode == "nccl":
                import torch.distributed as dist
                dist.gather(hits, glist, dst=0)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()         # started before the warm-up so that nvidia-smi is already streaming samples
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: device-resident inputs ----
    c0 = sess.counters()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e_start.record()
    for i in range(args.steps):
        step()
    e_stop.record()
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    c1 = sess.counters()
    total_ms = shard.max_over_ranks(e_start.elapsed_time(e_stop), device)
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    launches = int(shard.sum_over_ranks(c1.trace_launches - c0.trace_launches, device)) if world > 1 else int(c1.trace_launches - c0.trace_launches)

    # kernel-only time of one whole-batch launch (roofline numerator), measured live with CUDA events
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a_, b_ in kev:
        a_.record()
        sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
        b_.record()
    torch.cuda.synchronize()
    kern_ms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in kev]))

    # ---- gather verification (not timed): rank 0's buffer == every rank's local hits ----
    gather_ok = None
    if gather_mode in ("p2p", "nccl", "direct"):
        import torch.distributed as dist
        sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
        torch.cuda.synchronize()
        ref_all = shard.gather_hits(hits, dst=0, counts=[n] * world)
        if rank == 0:
            if gather_mode in ("p2p", "direct"):
                got = np.empty(world * n * 20, dtype=np.uint8)
                dev_view.d2h(got, gbuf_local, blocking=True)
                gather_ok = bool(got.tobytes() == ref_all.cpu().numpy().tobytes())
            else:
                gather_ok = bool(torch.equal(gathered, ref_all))
        del ref_all

    # ---- end to end: pinned host buffers in, pinned host buffers out, through the C ABI ----
    h_rays = torch.empty((n, 48), dtype=torch.uint8, pin_memory=True)
    h_hits = torch.empty((n, 20), dtype=torch.uint8, pin_memory=True)
    h_rays.copy_(rays)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        scene.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)      # synchronises inside
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n / e2e_s / 1e6
    # same batch through the reference-facing plugin sequence (AllocBufferRW/Enqueue/Read/Finish)
    sess.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)
    t0 = time.perf_counter()
    sess.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)
    plugin_s = time.perf_counter() - t0

    # ---- roofline of the traversal kernel ----
    st = scene.trace_stats(rays.data_ptr(), 0, n)
    nodes_per_ray = st.wide_nodes / max(1, st.rays)
    tris_per_ray = st.triangles / max(1, st.rays)
    a_impl = 48 + 20 + 64 * nodes_per_ray + 64 * tris_per_ray      # 64-B quantized wide nodes, 64-B triangle records
    peak, peak_src = read_peaks()
    traffic, traffic_src = read_traffic(workload)

    out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic rays (seeded, generated on the GPU) over " + ("a synthetic triangle soup" if is_soup(args.scene) else "the reference's %s scene geometry" % args.scene),
           "config": config, "gpu_launches": launches,
           "gather": {"mode": gather_mode, "chunks": args.chunks if gather_mode == "p2p" else None, "verified": gather_ok,
                      "bytes_per_step_into_rank0": (world - 1) * n * 20 if world > 1 else 0},
           "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 20,
                   "api": "lrb_trace_host (C ABI, pinned host buffers, chunked copy/trace overlap)",
                   "plugin_sequence_mrays_per_s": round(n / plugin_s / 1e6, 2)}}

    if rank == 0:
        cpu = None
        a_ref = None
        if not args.no_cpu_baseline and world == 1:
            O, bvh = oracle_for(desc, sess.bvh_nodes())
            rays_np = R.to_numpy_rays(rays)
            threads = O.hardware_threads()
            refbvh = reference_for(desc, sess.bvh_nodes())
            cn, cdt = cpu_time_sample(O, refbvh or bvh, rays_np, args.cpu_seconds, threads)
            cpu = {"value": round(cn / cdt / 1e6, 3), "unit": UNIT, "cores": threads, "kind": "reference" if refbvh else "port",
                   "sample": "first %d rays of the same batch, %s walking the same BVHArrayNode array, %d threads, %.1f s" % (
                       cn, "the reference's own BVHAccel::Intersect (oracle/_ref, compiled from the reference sources)" if refbvh
                       else "oracle BVHAccel::Intersect restatement", threads, cdt)}
            # reference-traversal visit counts on the same tree (canonical algorithmic bytes)
            k = min(rays_np.shape[0], 200000)
            _, cnt = bvh.intersect(rays_np[:k], nthreads=threads, count=True)
            a_ref = 48 + 20 + 32.0 * cnt[0] / k + 68.0 * cnt[1] / k
            # spot parity of the timed batch (not timed): device hits == oracle hits on a slice
            ref = (refbvh or bvh).intersect(rays_np[:k], nthreads=threads)
            got = hits[:k].cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
            same = (got["meshIndex"] == ref["meshIndex"]) & ((got["triangleIndex"] == ref["triangleIndex"]) | (ref["meshIndex"] == 0xFFFFFFFF))
            out["parity_check"] = {"rays": int(k), "index_mismatch": int((~same).sum()),
                                   "t_bit_exact": bool((got["t"][same] == ref["t"][same]).all()),
                                   "against": "reference library (oracle/_ref)" if refbvh else "oracle port"}
        alg = a_ref if a_ref is not None else a_impl
        achieved = alg * n / (kern_ms * 1e-3) / 1e9
        try:
            l2_bw = dev_view.measure_read_bandwidth(32 << 20, 20)
        except Exception:
            l2_bw = None
        out["roofline"] = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                           "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                           "note": (("scene (nodes + triangles = %.1f MB) is L2-resident: the algorithmic bytes are served by L1/L2, DRAM only sees "
                                     "the compulsory ray/hit streams (see traffic); frac is algorithmic bytes / HBM peak as the contract defines it "
                                     "and can exceed 1" % (info.device_bytes / 1e6)) if info.device_bytes < 100e6 else
                                    ("scene (%.1f GB on the device) does not fit L2: node / triangle fetches of incoherent rays go to HBM"
                                     % (info.device_bytes / 1e9))),
                           "algorithmic_bytes_per_ray": round(alg, 1),
                           "algorithmic_bytes_definition": ("A_ref = 68 + 32*N_inner + 68*N_leaf of the REFERENCE traversal on the same tree (SURVEY 8d)"
                                                            if a_ref is not None else "A_impl (reference visit counts unavailable at N>1)"),
                           "kernel_ms": round(kern_ms, 4), "units_per_launch": n,
                           "impl_bytes_per_ray": round(a_impl, 1), "impl_wide_nodes_per_ray": round(nodes_per_ray, 2),
                           "impl_triangles_per_ray": round(tris_per_ray, 2),
                           "impl_requested_gbs": round(a_impl * n / (kern_ms * 1e-3) / 1e9, 1),
                           "l2_read_peak_gbs_measured": round(l2_bw, 1) if l2_bw else None,
                           "impl_frac_of_l2_peak": round(a_impl * n / (kern_ms * 1e-3) / 1e9 / l2_bw, 4) if l2_bw else None}
        if cpu:
            out["cpu_baseline"] = cpu
        out["clocks"] = clocks
        out["scene"] = {"triangles": int(info.n_triangles), "ref_nodes": int(info.n_ref_nodes), "wide_nodes": int(info.n_wide_nodes),
                        "device_bytes": int(info.device_bytes), "host_build_s": round(build_s, 3)}
        print(json.dumps(out), flush=True)

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        if gbuf_peer:
            dev_view.ipc_close_handle(gbuf_peer)
        dist.barrier()
        if gbuf_local:
            dev_view.free(gbuf_local)
    sess.stop()
    sess.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args, rank, world, config):
    """The reference's own CPU algorithm (oracle port: the reference cannot be compiled here) on the
    host cores, same scene / ray kind / metric; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from luxcore_b200 import hostapi, rays as R
    desc = build_scene_arrays(args.scene)
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": args.builder, "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")       # host-only build, no GPU involved
    nodes = sess.bvh_nodes()
    O, bvh = oracle_for(desc, nodes)
    threads = O.hardware_threads()
    refbvh = reference_for(desc, nodes)
    if refbvh is not None:
        bvh = refbvh        # the reference's own code; same intersect(rays, nthreads=) call
    sample = int(os.environ.get("LRB_REF_SAMPLE", "1048576"))

    def trace_fn(rays_u8):
        r = R.to_numpy_rays(rays_u8)
        h = bvh.intersect(r, nthreads=threads)
        return torch.from_numpy(h.view(np.uint8).reshape(-1, 20).copy())

    rays = make_batch(trace_fn, desc, args, sample, seed=2, device="cpu")
    rays_np = R.to_numpy_rays(rays)
    for _ in range(max(1, min(args.warmup, 2))):
        bvh.intersect(rays_np, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bvh.intersect(rays_np, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    v = round(sample / dt / 1e6, 3)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic rays (seeded) over " + ("a synthetic triangle soup" if is_soup(args.scene) else "the reference's %s scene geometry" % args.scene),
           "config": config, "gpu_launches": 0,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference" if refbvh is not None else "port",
                            "sample": "%d rays of the same ray kind per step (bounded sample of the %d-ray batch), %s, %d threads"
                                      % (sample, args.rays, "the reference's own BVHAccel::Intersect (oracle/_ref) on the product's binned-SAH BVHArrayNode array"
                                         if refbvh is not None else "oracle restatement of BVHAccel::Intersect", threads)},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
