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
One step = one EnqueueTraceRayBuffer over one batch per GPU (weak scaling, the default: every rank owns a
16 Mi batch of its own seed; `--scaling strong`: ONE 16 Mi batch cut into the contiguous per-rank slices of
SURVEY.md 8e; the BVH is replicated); with N > 1 the step also gathers the RayHit buffers onto rank 0 (the only
exchange the path has).  `--scene soup` is BASELINE.json configs[4] (50 M-triangle soup, HBM-resident),
`--scene lightinstances --accel MBVH --depth 1 --rays 4194304` configs[2] (two-level traversal).

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
    """Per-launch counters of the traversal kernel on this workload from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_roofline.py); None if no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


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


def bind_to_gpu_numa_node(device_index):
    """One process per GPU: keep this rank's host threads -- and, by first touch, its pinned Ray / RayHit buffers -- on the
    NUMA node its GPU hangs off, so that eight ranks streaming 70 GB/s each do not all go through one socket's memory
    controllers and the inter-socket link (the host-buffer `e2e` of round 1 scaled 2.15x at 8 GPUs).  Reads the GPU's node
    from sysfs; a no-op (reported as such) when the platform does not tell.  -> dict for the JSON line."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return {"bound": False, "gpu": bdf, "why": "the platform reports no NUMA node for the GPU"}
        cpus = set()
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            for part in f.read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return {"bound": False, "gpu": bdf, "node": node, "why": "none of the node's CPUs is available to this process"}
        os.sched_setaffinity(0, allowed)
        return {"bound": True, "gpu": bdf, "node": node, "cpus": len(allowed)}
    except Exception as e:      # no sysfs, no permission, ...: run unbound
        return {"bound": False, "why": ("%s: %s" % (type(e).__name__, e))[:120]}


def oracle_for(desc, nodes, accel="BVH"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    from oracle import oracle as O
    osc = H.oracle_scene(desc)
    if accel == "MBVH":
        return O, O.MBVH(osc)      # the oracle's own (CLASSIC) trees: closest-hit results do not depend on the topology
    return O, O.BVH(osc, nodes=nodes)


def reference_for(desc, nodes, accel="BVH"):
    """The reference's own BVHAccel / MBVHAccel (oracle/_ref: its C++ sources compiled from /root/reference) --
    BVH: walking the same BVHArrayNode array; MBVH: its own CLASSIC-built trees -- or None where the prebuilt
    library is absent (then the oracle port is used)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    try:
        import helpers as H
        from oracle import refapi as RF
        if not RF.available():
            return None
        if accel == "MBVH":
            return RF.MBVH(H.reference_scene(desc))
        return RF.BVH(H.reference_scene(desc), nodes=nodes)
    except Exception as e:      # a broken checker must not take the benchmark down
        log("reference library unavailable:", repr(e))
        return None


def embree_status():
    """the Embree CPU baseline is only timed where libembree3 exists on the box."""
    import ctypes.util
    return "available" if ctypes.util.find_library("embree3") else "unavailable (no libembree3 on this box; the CPU arm is the reference's native BVH/MBVH Intersect)"


def cpu_time_sample(O, bvh, rays_np, target_s, threads):
    """Times the oracle (reference algorithm, CPU) on a bounded prefix of the batch."""
    probe = min(rays_np.shape[0], 200000)
    t0 = time.perf_counter(); bvh.intersect(rays_np[:probe], nthreads=threads); dt = time.perf_counter() - t0
    rate = probe / max(dt, 1e-9)
    n = int(min(rays_np.shape[0], max(probe, rate * target_s)))
    t0 = time.perf_counter(); bvh.intersect(rays_np[:n], nthreads=threads); dt = time.perf_counter() - t0
    return n, dt


def accel_config(args):
    return {"accelerator.type": args.accel, "accelerator.bvh.builder.type": args.builder, "accelerator.bvh.treetype": 4}


# ---------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene", default="kitchen")
    ap.add_argument("--accel", default="BVH", choices=["BVH", "MBVH"])
    ap.add_argument("--rays", type=int, default=RAYS_PER_BATCH, help="rays per batch per GPU (weak) / per batch (strong)")
    ap.add_argument("--depth", type=int, default=2, help="bounce depth of the ray batch")
    ap.add_argument("--builder", default="EMBREE_BINNED_SAH")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--gather", default="p2p", choices=["nccl", "p2p", "direct", "none"])
    ap.add_argument("--chunks", type=int, default=1, help="p2p gather: >= 1 = launches per batch, each followed by a copy-engine push of its RayHit "
                    "slice (default 1: plain kernel + one push, measured fastest); 0 = ONE kernel that signals finished chunks to the copy stream")
    ap.add_argument("--no-pipeline", action="store_true", help="p2p gather: every step waits for its own pushes and completion signal "
                    "(default: the tail of step k's pushes and its signal overlap the trace of step k + 1; two RayHit buffers)")
    ap.add_argument("--completion", default="flag", choices=["flag", "nccl"], help="pipelined p2p gather: completion signal = a step counter written by the "
                    "copy engine into rank 0's memory (default) or a 4-byte ncclAllReduce on a side stream")
    ap.add_argument("--timeline", action="store_true", help="N > 1: CUDA-event breakdown of un-pipelined steps per rank (kernel / tail of the pushes / signal)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="device option key=value")
    ap.add_argument("--numa", default="auto", choices=["auto", "on", "off"], help="bind the rank's host threads (and so its pinned buffers) to the NUMA node "
                    "of its GPU: auto = when there is more than one rank (at N = 1 the CPU baseline keeps every host core)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = args.gpus
    strong = args.scaling == "strong" and world > 1
    kind = "uniform" if is_soup(args.scene) else "bounce%d" % args.depth
    workload = "%s-%dM-%s" % (args.scene, args.rays >> 20, kind) if args.rays >= (1 << 20) else "%s-%d-%s" % (args.scene, args.rays, kind)
    config = {"workload": workload,
              "scene": ("synthetic random triangle soup, seed 4 (BASELINE.json configs[4])" if is_soup(args.scene)
                        else "scenes/%s (fixture tests/golden/scenes/%s.npz)" % (args.scene, args.scene)),
              "accelerator": args.accel, "builder": args.builder, "treetype": 4,
              "rays_per_batch_per_gpu": args.rays if not strong else None, "rays_per_batch_total": args.rays if strong else args.rays * max(1, n_gpus),
              "ray_kind": ("incoherent: origins uniform in the unit cube, directions uniform on the sphere" if is_soup(args.scene)
                           else "incoherent diffuse bounce, path depth %d" % args.depth),
              "parallelism": ("replicated BVH, ONE %d-ray batch cut into %d contiguous slices, RayHit gathered on rank 0 (%s)" % (args.rays, n_gpus, args.gather)
                              if strong else
                              "replicated BVH, one %d-ray batch per GPU x%d, RayHit gathered on rank 0 (%s)" % (args.rays, n_gpus, args.gather if n_gpus > 1 else "n/a")),
              "l2_policy": "inputs larger than L2 (48 B x rays + 20 B x rays per step >> 126 MB)"}

    if args.impl == "reference":
        return run_reference(args, rank, world, config)

    from luxcore_b200 import capi, hostapi, rays as R, shard

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)
    numa = bind_to_gpu_numa_node(local_rank) if (args.numa == "on" or (args.numa == "auto" and world > 1)) else {"bound": False, "why": "one rank: not requested"}

    desc = build_scene_arrays(args.scene)
    # GPU builders (EMBREE_MORTON / B200_PLOC) build on the GPU the rank traces on, so that the scene stays where it was built
    os.environ.setdefault("LRB_BUILDER_DEVICE", str(local_rank))
    t0 = time.perf_counter()
    sess = hostapi.Session(accel_config(args), desc)
    sess.build_accelerator(args.accel)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    sess.start(local_rank)
    upload_s = time.perf_counter() - t0         # host-side re-layout (relayout.cpp) + H2D of the scene
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

    if strong:
        # ONE batch (the rank-0 batch of the weak run), every rank regenerates it and keeps its slice [g N/G, (g+1) N/G)
        full = make_batch(trace_fn, desc, args, args.rays, seed=2, device=device)
        first, n = shard.rank_slice(args.rays, world, rank)
        rays = full[first:first + n].clone()
        del full
        counts = [shard.rank_slice(args.rays, world, r)[1] for r in range(world)]
    else:
        n = args.rays
        rays = make_batch(trace_fn, desc, args, n, seed=2 + rank, device=device)
        counts = [n] * world
    offsets = [sum(counts[:r]) for r in range(world)]
    total_rays = sum(counts)
    torch.cuda.synchronize()
    hits = torch.empty((n, 20), dtype=torch.uint8, device=device)

    # ---- multi-GPU: RayHit gather onto rank 0 ----
    #   p2p  : rank 0 owns the gather buffer; the other ranks map it over NVLink (CUDA IPC) and lrb_trace_gather
    #          pushes each finished chunk into it with the copy engine while the kernel keeps tracing.  Pipelined
    #          (default): the pushes of step k and its completion signal are not waited for before step k + 1 traces.
    #   nccl : trace, then torch.distributed gather (the baseline way)
    dev_view = capi.Device.borrow(sess.native_device())
    gather_mode = args.gather if world > 1 else "none"
    pipeline = gather_mode == "p2p" and not args.no_pipeline
    n_buf = 2 if pipeline else 1
    gbuf_local, gbuf_peer, my_dst, glist = [], [], [], None
    flags_local, flags_peer, my_flag = 0, 0, 0
    hits_buf = [hits] + ([torch.empty((n, 20), dtype=torch.uint8, device=device)] if pipeline else [])
    side = torch.cuda.Stream(device=device) if pipeline else None
    if gather_mode in ("p2p", "direct"):
        import torch.distributed as dist
        if rank == 0:
            gbuf_local = [dev_view.alloc(total_rays * 20) for _ in range(n_buf)]
            # one completion flag per rank (step counters), written by the ranks' copy engines, waited for on this GPU's queue
            flags_local = dev_view.alloc(4 * world)
            dev_view.h2d(flags_local, np.zeros(world, dtype=np.uint32), blocking=True)
            handle = [[dev_view.ipc_get_handle(p) for p in gbuf_local + [flags_local]]]
        else:
            handle = [None]
        dist.broadcast_object_list(handle, src=0)
        if rank == 0:
            my_dst = list(gbuf_local)
            my_flag = flags_local
        else:
            opened = [dev_view.ipc_open_handle(h) for h in handle[0]]
            gbuf_peer = opened[:n_buf]
            flags_peer = opened[n_buf]
            my_dst = [p + offsets[rank] * 20 for p in gbuf_peer]
            my_flag = flags_peer + 4 * rank
        flag = torch.zeros(1, dtype=torch.int32, device=device)
        if pipeline:
            sess.set_option("gather_defer", 1)
    elif gather_mode == "nccl":
        gathered = torch.empty((total_rays, 20), dtype=torch.uint8, device=device) if rank == 0 else None
        glist = [gathered[offsets[i]:offsets[i] + counts[i]] for i in range(world)] if rank == 0 else None

    pending = []        # async completion signals of the pipelined gather
    step_no = [0]
    probe = [None]      # when a list: (start, stop) CUDA events around every step's trace on the device queue

    def step():
        k = step_no[0]
        step_no[0] += 1
        if gather_mode == "direct":
            import torch.distributed as dist
            # every rank's tracer lanes store their RayHit records straight into rank 0's buffer (NVLink stores)
            sess.trace_device(rays.data_ptr(), my_dst[0], n)
            dist.all_reduce(flag)
        elif gather_mode == "p2p":
            import torch.distributed as dist
            b = k % n_buf
            if probe[0] is not None:
                probe[0].append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                probe[0][-1][0].record(stream)
            # rank 0 traces straight into its slice of the gather buffer (no copy at all)
            scene.trace_gather(rays.data_ptr(), my_dst[b] if rank == 0 else hits_buf[b].data_ptr(), n, my_dst[b], args.chunks)
            if probe[0] is not None:
                probe[0][-1][1].record(stream)
            if not pipeline:
                dist.all_reduce(flag)           # 4-byte "batch complete" signal, ordered after the pushes
            elif args.completion == "flag":
                # completion signal of step k: this rank's step counter, written into rank 0's flag word by the copy engine
                # behind the push -- no kernel, no NCCL (an all-reduce kernel cannot run next to a persistent trace kernel
                # that fills every SM: measured +0.25 ms per step at >= 4 GPUs)
                dev_view.gather_signal(my_flag, step_no[0] & 0xFFFF)
            else:
                # the signal of step k rides on a side stream behind this step's kernel and pushes; the queue goes on
                done = torch.cuda.Event()
                done.record(stream)
                side.wait_event(done)
                dev_view.gather_wait(side.cuda_stream, 0)
                with torch.cuda.stream(side):
                    pending.append(dist.all_reduce(flag, async_op=True))
        else:
            sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
            if gather_mode == "nccl":
                import torch.distributed as dist
                if len(set(counts)) == 1:
                    dist.gather(hits, glist, dst=0)
                else:
                    shard.gather_hits(hits, dst=0, counts=counts)

    def drain():
        """Everything the steps left in flight joins the queue (inside the timed region)."""
        if pipeline:
            dev_view.gather_wait(0, -1)
            if args.completion == "flag":
                if rank == 0:       # the gathered batch is complete when every rank's counter has reached this step
                    for r in range(world):
                        dev_view.wait_value(flags_local + 4 * r, step_no[0] & 0xFFFF)
            else:
                for w in pending:
                    w.wait()
                del pending[:]
                stream.wait_stream(side)

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
    drain()
    barrier()

    # ---- timed region: device-resident inputs ----
    c0 = sess.counters()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin = time.time()
    e_start.record()
    for i in range(args.steps):
        step()
    drain()
    e_stop.record()
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    c1 = sess.counters()
    total_ms = shard.max_over_ranks(e_start.elapsed_time(e_stop), device)
    ms_per_step = total_ms / args.steps
    value = total_rays / (ms_per_step * 1e-3) / 1e6
    launches = int(shard.sum_over_ranks(c1.trace_launches - c0.trace_launches, device)) if world > 1 else int(c1.trace_launches - c0.trace_launches)
    last_buf = (step_no[0] - 1) % n_buf

    # kernel-only time of one whole-batch launch (roofline numerator), measured live with CUDA events
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a_, b_ in kev:
        a_.record()
        sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
        b_.record()
    torch.cuda.synchronize()
    kern_ms = float(np.mean([a_.elapsed_time(b_) for a_, b_ in kev]))

    # ---- N > 1, pipelined gather: every rank's trace time inside the running pipeline (not part of the timed region): the
    # kernel of step k + 1 shares the GPU with the push of step k -- and rank 0's with the slices arriving from every rank
    pipe_kernel_ms = None
    if world > 1 and pipeline:
        import torch.distributed as dist
        barrier()
        probe[0] = []
        for _ in range(10):
            step()
        drain()
        torch.cuda.synchronize()
        mine = float(np.median([a_.elapsed_time(b_) for a_, b_ in probe[0][2:]]))
        probe[0] = None
        allk = [None] * world
        dist.all_gather_object(allk, mine)
        pipe_kernel_ms = [round(x, 4) for x in allk]
        barrier()

    # ---- N > 1: where a step's time goes (CUDA events, un-pipelined steps; not part of the timed region) ----
    timeline = None
    if args.timeline and gather_mode == "p2p" and args.chunks == 0:
        import torch.distributed as dist
        sess.set_option("gather_defer", 1)
        rows = []
        for it in range(6):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            barrier()
            ev[0].record()
            scene.trace_gather(rays.data_ptr(), my_dst[0] if rank == 0 else hits_buf[0].data_ptr(), n, my_dst[0], 0)
            ev[1].record()                      # the kernel (and the flag memset behind it) has finished
            dev_view.gather_wait(0, -1)
            ev[2].record()                      # ... and so have this rank's pushes
            dist.all_reduce(flag)
            ev[3].record()                      # ... and every rank's (completion signal)
            torch.cuda.synchronize()
            if it:
                rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
        med = [float(np.median([r[i] for r in rows])) for i in range(3)]
        allr = [None] * world
        dist.all_gather_object(allr, med)
        timeline = {"what": "median over 5 un-pipelined steps, per rank: ms of [trace kernel incl. in-flight pushes, tail of the pushes after the kernel, completion all-reduce incl. waiting for the slowest rank]",
                    "per_rank_ms": [[round(x, 4) for x in r] for r in allr], "kernel_alone_ms_rank0": round(kern_ms, 4)}
        if not pipeline:
            sess.set_option("gather_defer", 0)

    # ---- gather verification (not timed): rank 0's buffer == every rank's local hits ----
    gather_ok = None
    if gather_mode in ("p2p", "nccl", "direct"):
        import torch.distributed as dist
        if gather_mode == "p2p":
            # one more gathered step into buffer 0, waited for, so that the check sees a complete, known state
            scene.trace_gather(rays.data_ptr(), my_dst[0] if rank == 0 else hits_buf[0].data_ptr(), n, my_dst[0], args.chunks)
            dev_view.gather_wait(0, -1)
            torch.cuda.synchronize()
            dist.barrier()
        sess.trace_device(rays.data_ptr(), hits.data_ptr(), n)
        torch.cuda.synchronize()
        ref_all = shard.gather_hits(hits, dst=0, counts=counts)
        if rank == 0:
            if gather_mode in ("p2p", "direct"):
                got = np.empty(total_rays * 20, dtype=np.uint8)
                dev_view.d2h(got, gbuf_local[0], blocking=True)
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
    e2e_value = total_rays / e2e_s / 1e6
    # same batch through the reference-facing plugin sequence (AllocBufferRW/Enqueue/Read/Finish)
    sess.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)
    t0 = time.perf_counter()
    sess.trace_host_ptr(h_rays.data_ptr(), h_hits.data_ptr(), n)
    plugin_s = time.perf_counter() - t0
    # (not timed) what came back through the host pipelines is, byte for byte, what a device-resident trace of the batch gives
    h_hits2 = torch.empty_like(h_hits)
    scene.trace_host_ptr(h_rays.data_ptr(), h_hits2.data_ptr(), n)
    d_chk = torch.empty((n, 20), dtype=torch.uint8, device=device)
    sess.trace_device(rays.data_ptr(), d_chk.data_ptr(), n)
    torch.cuda.synchronize()
    e2e_verified = bool(torch.equal(d_chk, h_hits.to(device))) and bool(torch.equal(d_chk, h_hits2.to(device)))
    del d_chk, h_hits2

    # ---- what the implementation itself asks memory for (instrumented kernel) ----
    st = scene.trace_stats(rays.data_ptr(), 0, n)
    nodes_per_ray = st.wide_nodes / max(1, st.rays)
    tris_per_ray = st.triangles / max(1, st.rays)
    inst_per_ray = st.instances / max(1, st.rays)
    # 64-B quantized wide nodes, 64-B triangle records, 32-B instance record + 64-B matrix per instance entry
    a_impl = 48 + 20 + 64 * nodes_per_ray + 64 * tris_per_ray + 96 * inst_per_ray

    # ---- spot parity of the timed batch (not timed) on every rank: device hits == reference / oracle hits ----
    parity = None
    a_ref = None
    cpu = None
    if not args.no_cpu_baseline:
        k = min(n, 200000 if world == 1 else 50000)
        nodes = sess.bvh_nodes() if args.accel == "BVH" else None
        O, bvh = oracle_for(desc, nodes, args.accel)
        threads = max(1, O.hardware_threads() // max(1, world))
        refbvh = reference_for(desc, nodes, args.accel) if rank == 0 else None
        rays_np = R.to_numpy_rays(rays[:k]) if world > 1 or rank != 0 else R.to_numpy_rays(rays)
        checker = refbvh or bvh
        ref = checker.intersect(rays_np[:k], nthreads=threads)
        got = hits[:k].cpu().numpy().reshape(-1).view(capi.HIT_DTYPE)
        same = (got["meshIndex"] == ref["meshIndex"]) & ((got["triangleIndex"] == ref["triangleIndex"]) | (ref["meshIndex"] == 0xFFFFFFFF))
        mism = int((~same).sum())
        t_exact = bool((got["t"][same] == ref["t"][same]).all())
        if world > 1:
            mism = int(shard.sum_over_ranks(mism, device))
            t_exact = shard.sum_over_ranks(0 if t_exact else 1, device) == 0
        parity = {"rays": int(k) * world, "ranks_checked": world, "index_mismatch": mism, "t_bit_exact": bool(t_exact),
                  "against": "reference library (oracle/_ref)" if refbvh else "oracle port"}
        if rank == 0 and world == 1:
            cn, cdt = cpu_time_sample(O, checker, rays_np, args.cpu_seconds, threads)
            what = ("the reference's own %sAccel::Intersect (oracle/_ref, compiled from the reference sources)" % args.accel) if refbvh \
                else "oracle %sAccel::Intersect restatement" % args.accel
            cpu = {"value": round(cn / cdt / 1e6, 3), "unit": UNIT, "cores": threads, "kind": "reference" if refbvh else "port",
                   "embree": embree_status(),
                   "sample": "first %d rays of the same batch, %s%s, %d threads, %.1f s" % (
                       cn, what, " walking the same BVHArrayNode array" if args.accel == "BVH" else " on its own CLASSIC-built trees", threads, cdt)}
        if rank == 0:
            # reference-traversal visit counts (canonical algorithmic bytes)
            _, cnt = bvh.intersect(rays_np[:k], nthreads=threads, count=True)
            a_ref = 48 + 20 + 32.0 * cnt[0] / k + 68.0 * cnt[1] / k

    out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic rays (seeded, generated on the GPU) over " + ("a synthetic triangle soup" if is_soup(args.scene) else "the reference's %s scene geometry" % args.scene),
           "config": config, "gpu_launches": launches,
           "gather": {"mode": gather_mode, "chunks": args.chunks if gather_mode == "p2p" else None, "pipelined": pipeline,
                      "signal": (args.completion if pipeline else ("nccl" if gather_mode == "p2p" else None)), "verified": gather_ok,
                      "bytes_per_step_into_rank0": (total_rays - counts[0]) * 20 if world > 1 else 0},
           "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": n * 48, "d2h_bytes_per_step": n * 20,
                   "verified_against_device_trace": e2e_verified,
                   "api": "lrb_trace_host (C ABI, pinned host buffers, chunked copy/trace overlap)",
                   "plugin_sequence_mrays_per_s": round(n / plugin_s / 1e6, 2)}}
    out["numa"] = numa
    if parity:
        out["parity_check"] = parity
    if timeline:
        out["timeline"] = timeline
    if pipe_kernel_ms:
        out["gather"]["per_rank_trace_ms_in_pipeline"] = pipe_kernel_ms

    if rank == 0:
        peak, peak_src = read_peaks()
        cap = read_traffic(workload)
        sm_clock = (clocks or {}).get("sm_mhz") or 1965.0
        sms = dev_view.props().sm_count
        try:
            l2_bw = max(dev_view.measure_read_bandwidth(64 << 20, 30), dev_view.measure_read_bandwidth(32 << 20, 30))
        except Exception:
            l2_bw = None
        kern_s = kern_ms * 1e-3
        resident = info.device_bytes < 100e6
        roof = {}
        if cap and cap.get("dram_bytes_per_launch"):
            scale = n / float(cap.get("rays_in_launch") or n)        # capture and bench launch of the same size: 1
            dram = cap["dram_bytes_per_launch"] * scale
            hbm_gbs = dram / kern_s / 1e9
            memory = {"hbm": {"achieved_gbs": round(hbm_gbs, 1), "peak_gbs": peak, "frac": round(hbm_gbs / peak, 4),
                              "bytes_per_ray": round(dram / n, 1)}}
            if cap.get("lts_bytes_per_launch"):
                l2_gbs = cap["lts_bytes_per_launch"] * scale / kern_s / 1e9
                memory["l2"] = {"achieved_gbs": round(l2_gbs, 1), "peak_gbs_measured": round(l2_bw, 1) if l2_bw else None,
                                "frac": round(l2_gbs / l2_bw, 4) if l2_bw else None, "ncu_lts_throughput_pct": cap.get("lts_throughput_pct"),
                                "hit_rate_pct": cap.get("lts_hit_rate_pct")}
            if cap.get("l1_global_load_bytes_per_launch"):
                l1b = cap["l1_global_load_bytes_per_launch"] * scale
                memory["l1_requests"] = {"ncu_sector_bytes_per_ray": round(l1b / n, 1), "a_impl_bytes_per_ray": round(a_impl, 1),
                                         "a_impl_over_ncu": round(a_impl * n / l1b, 3),
                                         "note": "A_impl counts every lane's 64-B node / triangle fetch; ncu counts the distinct 32-B sectors of a warp's "
                                                 "request, so lanes of a warp that fetch the same node (all of them near the root) are counted once",
                                         "l1_data_pipe_pct": cap.get("l1_data_pipe_pct"), "hit_rate_pct": cap.get("l1_hit_rate_pct")}
            if resident and cap.get("warp_instructions_per_ray"):
                # L2-resident scene: the kernel is bound by the SMs' issue slots, not by a memory level
                ginst = cap["warp_instructions_per_ray"] * n / kern_s / 1e9
                ipeak = sms * 4 * sm_clock * 1e6 / 1e9
                roof = {"bound": "issue", "achieved": round(ginst, 1), "peak": round(ipeak, 1), "unit": "Gwarp-inst/s", "frac": round(ginst / ipeak, 4),
                        "traffic": dram, "peak_source": "%d SMs x 4 schedulers x %.0f MHz (SM clock sampled during the timed region)" % (sms, sm_clock),
                        "warp_instructions_per_ray": cap["warp_instructions_per_ray"], "threads_per_instruction": cap.get("threads_per_instruction"),
                        "ncu_issue_active_pct": cap.get("issue_active_pct"), "ncu_alu_pipe_pct": cap.get("alu_pipe_pct"),
                        "note": "scene (%.1f MB) is L2-resident; ncu: issue slots %.0f %%, ALU pipe %.0f %%, L1 data pipe %.0f %% busy, DRAM %.1f %% "
                                "and L2 %.1f %% of their peaks -- the bound is instruction issue at %.1f of 32 lanes per instruction" % (
                                    info.device_bytes / 1e6, cap.get("issue_active_pct") or 0, cap.get("alu_pipe_pct") or 0, cap.get("l1_data_pipe_pct") or 0,
                                    100.0 * hbm_gbs / peak, cap.get("lts_throughput_pct") or 0, cap.get("threads_per_instruction") or 0)}
            else:
                roof = {"bound": "hbm", "achieved": round(hbm_gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(hbm_gbs / peak, 4), "traffic": dram,
                        "peak_source": peak_src,
                        "note": "scene (%.1f GB on the device) does not fit L2; achieved = ncu-measured DRAM bytes of one launch / live kernel time. "
                                "ncu: L2 hit rate %.0f %%, issue slots %.0f %% busy (%.0f warp instructions per ray at %.1f of 32 lanes), top stall "
                                "long_scoreboard %.1f warps per issue: the walk sits between the SMs' issue bound and the latency of dependent node "
                                "fetches; HBM bandwidth itself is not the limit (more resident warps spill registers and run slower, an L2 prefetch of the "
                                "pushed children costs more traffic than it hides: both measured, DESIGN.md)" % (
                                    info.device_bytes / 1e9, cap.get("lts_hit_rate_pct") or 0, cap.get("issue_active_pct") or 0,
                                    cap.get("warp_instructions_per_ray") or 0, cap.get("threads_per_instruction") or 0,
                                    (cap.get("top_stalls_warps_per_issue") or {}).get("long_scoreboard", 0))}
            roof["memory"] = memory
            roof["traffic_source"] = cap.get("source")
            roof["captured_at_commit"] = cap.get("captured_at_commit")
        else:
            # no capture of this workload under profiles/: only the implementation's own requested bytes can be stated
            req = a_impl * n / kern_s / 1e9
            roof = {"bound": "hbm" if not resident else "issue", "achieved": round(req, 1), "peak": peak, "unit": "GB/s", "frac": round(req / peak, 4),
                    "traffic": None, "peak_source": peak_src,
                    "note": "no ncu capture of this workload is committed: achieved = bytes requested by the implementation (A_impl) / kernel time"}
        alg = a_ref if a_ref is not None else a_impl
        roof.update({"kernel_ms": round(kern_ms, 4), "units_per_launch": n,
                     "impl_bytes_per_ray": round(a_impl, 1), "impl_wide_nodes_per_ray": round(nodes_per_ray, 2),
                     "impl_triangles_per_ray": round(tris_per_ray, 2), "impl_instance_entries_per_ray": round(inst_per_ray, 2),
                     "impl_requested_gbs": round(a_impl * n / kern_s / 1e9, 1),
                     "canonical": {"algorithmic_bytes_per_ray": round(alg, 1),
                                   "definition": ("A_ref = 68 + 32*N_inner + 68*N_leaf of the REFERENCE traversal (SURVEY 8d)" if a_ref is not None
                                                  else "A_impl (reference visit counts not computed in this run)"),
                                   "achieved_gbs": round(alg * n / kern_s / 1e9, 1), "frac_of_hbm_peak": round(alg * n / kern_s / 1e9 / peak, 4),
                                   "note": "SURVEY 8d's implementation-independent figure: bytes the REFERENCE's stackless walk would fetch for these rays "
                                           "per second of this kernel; it is served by L1/L2 and says how much less this traversal touches, not how busy HBM is"}})
        out["roofline"] = roof
        if cpu:
            out["cpu_baseline"] = cpu
        out["clocks"] = clocks
        out["scene"] = {"triangles": int(info.n_triangles), "ref_nodes": int(info.n_ref_nodes), "wide_nodes": int(info.n_wide_nodes),
                        "instances": int(info.n_instances), "device_bytes": int(info.device_bytes), "host_build_s": round(build_s, 3),
                        "relayout_and_upload_s": round(upload_s, 3)}
        print(json.dumps(out), flush=True)

    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()
        for p_ in gbuf_peer + ([flags_peer] if flags_peer else []):
            dev_view.ipc_close_handle(p_)
        dist.barrier()
        for p_ in gbuf_local + ([flags_local] if flags_local else []):
            dev_view.free(p_)
    sess.stop()
    sess.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args, rank, world, config):
    """The reference's own CPU implementation of the path (oracle/_ref: BVHAccel / MBVHAccel::Intersect compiled from the
    reference's sources; the oracle port where that library is absent) on the host cores, same scene / ray kind / metric;
    each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from luxcore_b200 import hostapi, rays as R
    desc = build_scene_arrays(args.scene)
    sess = hostapi.Session(accel_config(args), desc)
    sess.build_accelerator(args.accel)       # host-only build, no GPU involved
    nodes = sess.bvh_nodes() if args.accel == "BVH" else None
    O, bvh = oracle_for(desc, nodes, args.accel)
    threads = O.hardware_threads()
    refbvh = reference_for(desc, nodes, args.accel)
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
    what = ("the reference's own %sAccel::Intersect (oracle/_ref)%s" % (args.accel, " on the product's binned-SAH BVHArrayNode array" if args.accel == "BVH" else "")
            if refbvh is not None else "oracle restatement of %sAccel::Intersect" % args.accel)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": max(1, min(args.warmup, 2)), "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong" if args.scaling == "strong" and world > 1 else "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic rays (seeded) over " + ("a synthetic triangle soup" if is_soup(args.scene) else "the reference's %s scene geometry" % args.scene),
           "config": config, "gpu_launches": 0,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "reference" if refbvh is not None else "port", "embree": embree_status(),
                            "sample": "%d rays of the same ray kind per step (bounded sample of the %d-ray batch), %s, %d threads" % (sample, args.rays, what, threads)},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
