"""Ad-hoc kernel timing on the GPU box (development aid, not the benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import helpers as H
from luxcore_b200 import capi, rays as R, scenes as S
from oracle import oracle as O

name = sys.argv[1] if len(sys.argv) > 1 else "kitchen"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4 << 20
opts = sys.argv[3:]
desc = S.load_fixture(name)
osc = H.oracle_scene(desc)
bvh = O.BVH(osc, tree_type=4)
verts, offs = H.flattened_from_oracle(desc, osc)
dev = capi.Device(0)
scene = dev.upload_bvh(bvh.nodes(), verts, offs)
info = scene.info()
print("scene", name, "wide", info.n_wide_nodes, "tris", info.n_triangles, "stack", info.stack_need)
torch.cuda.set_device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
assert stream.cuda_stream != 0
dev.set_stream(stream.cuda_stream)
lo, hi = desc.bbox()
batches = {
    "uniform": R.uniform_rays(lo, hi, n, seed=3, device="cuda"),
    "camera": R.camera_rays(desc.cam, int(n ** 0.5), int(n ** 0.5), seed=1, device="cuda"),
}
def run(label, rays, reps=5):
    m = rays.shape[0]
    hits = torch.empty((m, 20), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        scene.trace(rays.data_ptr(), hits.data_ptr(), m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        scene.trace(rays.data_ptr(), hits.data_ptr(), m)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = scene.trace_stats(rays.data_ptr(), 0, m)
    print("%-28s %8.3f ms  %8.1f Mrays/s   nodes/ray %.1f tris/ray %.1f maxstack %d" % (
        label, ms, m / ms / 1e3, st.wide_nodes / max(1, st.rays), st.triangles / max(1, st.rays), st.max_stack))
    return hits

configs = [dict(kernel="persistent"), dict(kernel="simple")]
for o in opts:
    configs.append(dict(kv.split("=") for kv in o.split(",")))
for cfg in configs:
    for k, v in cfg.items():
        dev.set_option(k, v)
    for bname, rays in batches.items():
        run("%s %s" % (bname, cfg), rays)
