#!/usr/bin/env python3
"""Multi-GPU check + timing of the film merge over NVLink peer memory (shard.FilmMerger -> lrb_film_reduce).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/film_merge_check.py
Every rank fills a 3840 x 2160 x (RGB + weight) film with seeded values; rank 0 checks the merged film bit for bit
against the reference's device-order sum of the films (collected with NCCL for the check only) and prints timings
next to ncclReduce of the same planes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from luxcore_b200 import capi, shard

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = capi.Device(local)
n = 3840 * 2160 * 4
fm = shard.FilmMerger(dev, n, dst=0)
g = torch.Generator(device="cuda"); g.manual_seed(1234 + rank)
film = (torch.rand(n, generator=g, device="cuda") * (10.0 ** (rank % 5 - 2))).float()
host = film.cpu().numpy()
dev.h2d(fm.film, host, blocking=True)
fm.merge()
times = []
for _ in range(5):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); fm.merge(); times.append(time.perf_counter() - t0)
# NCCL reduce of the same planes for comparison (sum order is NCCL's own: not bit-comparable)
nt = []
for _ in range(5):
    x = film.clone(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); dist.reduce(x, dst=0); torch.cuda.synchronize(); nt.append(time.perf_counter() - t0)
allf = [torch.empty_like(film) for _ in range(world)] if rank == 0 else None
dist.gather(film, allf, dst=0)
if rank == 0:
    want = torch.zeros_like(film)
    for t in allf:
        want = want + t
    got = np.empty(n, dtype=np.float32)
    dev.d2h(got, fm.merged, blocking=True)
    ok = bool(got.tobytes() == want.cpu().numpy().tobytes())
    out = {"film": "3840x2160x4 float32 (%.1f MB)" % (n * 4 / 1e6), "gpus": world, "bit_exact_vs_device_order_sum": ok,
           "merge_ms_incl_two_barriers": round(1e3 * float(np.median(times)), 3), "nccl_reduce_ms": round(1e3 * float(np.median(nt)), 3)}
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "film_merge_%dgpu.json" % world), "w"), indent=1)
fm.close()
dev.close()
dist.destroy_process_group()
