"""Randomised differential test of the product's re-layout + traversal body (CPU emulation of the same
traverse.h / relayout.cpp the kernels use) against the oracle, on scenes and rays the fixed tests do not
enumerate: random scales (1e-3 ... 1e4) and offsets (up to 1e5) of the geometry, slivers, degenerate and
duplicated triangles, coplanar sheets, every builder and arity, one- and two-level scenes with random
(rotating, mirroring, non-uniformly scaling) instances, and ray batches that mix uniform rays, rays starting
on / within 2 eps of surfaces, axis-parallel directions, finite and tiny maxt, zero and negative mint.

    python tools/fuzz_parity.py [seconds] [seed] [--lockstep] [--reference]

--lockstep also runs the REAL kernel source (trace_kernels.cuh compiled for the host, one OS thread per lane)
on a slice of every batch with random warp counts, stack depths and vote settings; --reference pins the oracle
against the reference's own code (oracle/_ref) on every one-level scene, bit for bit.

Prints one line per scene; any disagreement that is not a t-tie within the stated epsilon raises.
CPU only (test infrastructure: imports oracle/)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import helpers as H
import scene_zoo as Z
from luxcore_b200 import hostapi, rays as R, scenes as S
from oracle import oracle as O


def random_mesh(rng, n_tris, scale, offset):
    kind = rng.integers(0, 5)
    if kind == 0:       # soup of small triangles
        c = rng.uniform(-1, 1, (n_tris, 1, 3))
        v = c + rng.uniform(-0.15, 0.15, (n_tris, 3, 3))
    elif kind == 1:     # slivers: long and thin
        c = rng.uniform(-1, 1, (n_tris, 1, 3))
        d = rng.normal(size=(n_tris, 1, 3))
        v = c + d * rng.uniform(-1, 1, (n_tris, 3, 1)) + rng.normal(scale=1e-4, size=(n_tris, 3, 3))
    elif kind == 2:     # coplanar sheets (axis-aligned planes: zero-thickness boxes)
        c = rng.uniform(-1, 1, (n_tris, 1, 3))
        v = c + rng.uniform(-0.2, 0.2, (n_tris, 3, 3))
        ax = rng.integers(0, 3)
        v[:, :, ax] = np.round(v[:, :1, ax] * 4) / 4
    elif kind == 3:     # a grid surface plus duplicates of some triangles
        gv, gt = S.grid_mesh(max(2, int(np.sqrt(n_tris / 2))), max(2, int(np.sqrt(n_tris / 2))), z=0.1, size=1.0)
        v = gv[gt].astype(np.float64)
        v = np.concatenate([v, v[rng.integers(0, v.shape[0], max(1, v.shape[0] // 10))]])
    else:               # mixed sizes over five decades, some degenerate (two equal vertices / a point)
        c = rng.uniform(-1, 1, (n_tris, 1, 3))
        v = c + rng.normal(size=(n_tris, 3, 3)) * (10.0 ** rng.uniform(-5, -0.5, (n_tris, 1, 1)))
        deg = rng.random(n_tris) < 0.05
        v[deg, 1] = v[deg, 0]
        pt = rng.random(n_tris) < 0.02
        v[pt, 1] = v[pt, 0]
        v[pt, 2] = v[pt, 0]
    v = (v * scale + offset).astype(np.float32).reshape(-1, 3)
    return v, np.arange(v.shape[0], dtype=np.uint32).reshape(-1, 3)


def random_rays(rng, desc, n, seed):
    lo, hi = desc.bbox()
    ext = np.maximum(hi - lo, 1e-6)
    a = R.to_numpy_rays(R.uniform_rays(lo - 0.3 * ext, hi + 0.3 * ext, n, seed=seed))
    p0, e1, e2, _ = S.world_triangles(desc)
    b = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n, seed=seed + 1, axis_fraction=0.3))
    rays = np.concatenate([a, b])
    m = rays.shape[0]
    r = rng.random(m)
    diag = float(np.linalg.norm(ext))
    rays["maxt"] = np.where(r < 0.15, rng.uniform(0, diag, m), rays["maxt"]).astype(np.float32)
    rays["maxt"] = np.where((r >= 0.15) & (r < 0.2), rays["mint"] * 4, rays["maxt"]).astype(np.float32)
    r2 = rng.random(m)
    rays["mint"] = np.where(r2 < 0.1, 0.0, rays["mint"]).astype(np.float32)
    rays["mint"] = np.where((r2 >= 0.1) & (r2 < 0.15), -rng.uniform(0, diag, m), rays["mint"]).astype(np.float32)
    # unnormalised directions (the reference does not require unit length, ray.h)
    s = np.where(rng.random(m) < 0.2, 10.0 ** rng.uniform(-3, 3, m), 1.0).astype(np.float32)
    rays["d"] = (rays["d"] * s[:, None]).astype(np.float32)
    if "time" in rays.dtype.names:
        rays["time"] = rng.uniform(-0.1, 1.1, m).astype(np.float32)
    # a few non-finite rays: NaN / infinite components in the origin, the direction or the interval
    weird = np.nonzero(rng.random(m) < 0.01)[0]
    for i in weird:
        v = [np.nan, np.inf, -np.inf, 0.0][int(rng.integers(0, 4))]
        f = ["o", "d", "mint", "maxt"][int(rng.integers(0, 4))]
        if f in ("o", "d"):
            rays[f][i, int(rng.integers(0, 3))] = v
        else:
            rays[f][i] = v
    return rays


SKIP_BELOW = int(os.environ.get("FUZZ_SKIP_BELOW", "0"))      # replay aid: scenes below this index only consume random numbers
CURRENT = [0]
REFERENCE = False       # --reference: also pin the oracle against the reference library on every one-level scene
LOCKSTEP = False        # --lockstep: also run the REAL kernel source (tests/cpp/kernel_lockstep.cpp) on a slice of every batch


def lockstep_check(rng, emu, rays, got, what):
    if not LOCKSTEP:
        return
    pick = np.sort(rng.choice(rays.shape[0], size=min(500, rays.shape[0]), replace=False))
    sub = np.ascontiguousarray(rays[pick])
    kw = dict(n_warps=int(rng.integers(1, 5)), smem_depth=int(rng.choice([2, 4, 16, 64])), refill_below=int(rng.choice([1, 16, 24, 32])),
              tri_bias=int(rng.choice([1, 8, 64])), inst_bias=int(rng.choice([0, 8, 64])))
    if rng.random() < 0.25:
        kw = dict(kernel="static", n_warps=4)
    if SKIP_BELOW > CURRENT[0]:
        return
    if os.environ.get("FUZZ_VERBOSE"):
        print("      lockstep", what, kw, "stack_need", emu.info()["stack_need"], "max stack of the per-ray emulation",
              emu.trace(sub, want_stats=True)[1]["max_stack"], flush=True)
    ks = H.Lockstep.trace(emu, sub, **kw)
    if ks.tobytes() != got[pick].tobytes():
        bad = np.nonzero(ks != got[pick])[0]
        raise AssertionError("%s: kernel source in lockstep %r differs from the emulation at rays %r" % (what, kw, pick[bad[:5]]))


def comparable(rays):
    """Rays with a NaN / infinite origin or direction component are outside the parity contract: the reference's
    answer for them is an artefact of its arithmetic (e.g. d.z = inf makes Triangle::Intersect report t = 0 with NaN
    barycentrics for every triangle of every node whose x-y footprint holds the origin), and the product culls
    triangles by their own boxes, which such "hits" need not lie in.  They are traced (no crash, no hang) but not
    compared.  Non-finite mint / maxt ARE compared."""
    return np.isfinite(rays["o"]).all(axis=1) & np.isfinite(rays["d"]).all(axis=1)


def one_level(rng, it):
    scale = 10.0 ** rng.uniform(-3, 4)
    offset = rng.uniform(-1, 1, 3) * (10.0 ** rng.uniform(-2, 5)) * (rng.random() < 0.6)
    desc = S.SceneDesc("fuzz%d" % it)
    for _ in range(int(rng.integers(1, 4))):
        desc.add_plain(desc.add_shape(*random_mesh(rng, int(rng.integers(1, 900)), scale, offset)))
    builder = ["CLASSIC", "EMBREE_BINNED_SAH"][int(rng.integers(0, 2))]
    tree_type = [2, 4, 8][int(rng.integers(0, 3))]
    s = hostapi.Session({"accelerator.bvh.builder.type": builder, "accelerator.bvh.treetype": tree_type}, desc)
    s.build_accelerator("BVH")
    nodes = s.bvh_nodes().copy()
    osc = H.oracle_scene(desc)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    if nodes.shape[0] > 1:
        # the DEVICE re-layout's per-record code (relayout_kernels.cuh, driven in the device pipeline's order on the host):
        # leaf payload, indices by exclusive sum, fill, bottom-up stack bound -- the bytes of the host re-layout
        tri, toff = H.flattened_triangles(desc)
        dv = H.RelayoutDev.run(H.to_builder_format(nodes, toff), verts, offs, tri, toff)
        w, t, i = emu.arrays()
        assert dv["ref_nodes"].tobytes() == nodes.tobytes(), ("device leaf payload", it)
        assert dv["wide"].tobytes() == w.tobytes() and dv["tris"].tobytes() == t.tobytes() and dv["ids"].tobytes() == i.tobytes(), ("device re-layout", it)
        assert dv["stack_need"] == emu.info()["stack_need"], ("device stack bound", it, dv["stack_need"], emu.info())
    rays = random_rays(rng, desc, 3000, int(rng.integers(1, 1 << 30)))
    ref = O.BVH(osc, nodes=nodes).intersect(rays)
    if REFERENCE and SKIP_BELOW <= CURRENT[0]:
        # the reference's own BVHAccel::Intersect (oracle/_ref, compiled from /root/reference) on the same array:
        # the oracle must reproduce it bit for bit, non-finite rays included
        from oracle import refapi as RF
        theirs = RF.BVH(H.reference_scene(desc), nodes=nodes).intersect(rays)
        same = (theirs["meshIndex"] == ref["meshIndex"]) & ((theirs["triangleIndex"] == ref["triangleIndex"]) | (ref["meshIndex"] == H.NULL))
        hit = same & (ref["meshIndex"] != H.NULL)
        same &= (theirs["t"].view(np.uint32) == ref["t"].view(np.uint32)) | ~hit
        same[hit] &= (theirs["b1"][hit].view(np.uint32) == ref["b1"][hit].view(np.uint32)) & (theirs["b2"][hit].view(np.uint32) == ref["b2"][hit].view(np.uint32))
        if not same.all():
            i = int(np.nonzero(~same)[0][0])
            raise AssertionError("fuzz %d: oracle %r != reference %r for ray %r" % (it, ref[i], theirs[i], rays[i]))
    got, st = emu.trace(rays, want_stats=True)
    assert st["max_stack"] <= emu.info()["stack_need"], ("stack bound", it, st["max_stack"], emu.info())    # non-finite rays included
    lockstep_check(rng, emu, rays, got, "fuzz %d" % it)
    ok = comparable(rays)
    rep = H.compare_hits_tie_aware(got[ok], ref[ok], rays[ok], osc, what="fuzz %d %s k=%d" % (it, builder, tree_type), max_ties=40)
    return "BVH  %-18s k=%d scale %8.2e |offset| %8.2e tris %5d: hits %5d bit-exact %5d ties %d" % (
        builder, tree_type, scale, float(np.abs(offset).max()), desc.triangle_count(), rep["hits"], rep["bit_exact_hits"], rep["tie_exempt"])


def two_level(rng, it):
    scale = 10.0 ** rng.uniform(-2, 3)
    desc = S.SceneDesc("fuzz2l%d" % it)
    shapes = [desc.add_shape(*random_mesh(rng, int(rng.integers(1, 300)), 1.0, np.zeros(3))) for _ in range(int(rng.integers(1, 4)))]
    if rng.random() < 0.5:
        desc.add_plain(shapes[0])
    for _ in range(int(rng.integers(1, 25))):
        m = Z.translate(*(rng.uniform(-4, 4, 3) * scale)) @ Z.rot_z(rng.uniform(0, 360)) @ Z.rot_x(rng.uniform(0, 360)) @ \
            Z.scale(*(rng.uniform(0.3, 2.0, 3) * scale))
        if rng.random() < 0.2:
            m = m @ Z.scale(-1, 1, 1)
        desc.add_instance(shapes[int(rng.integers(0, len(shapes)))], m)
    motion = rng.random() < 0.5
    if motion:
        for _ in range(int(rng.integers(1, 5))):
            m0 = Z.translate(*(rng.uniform(-4, 4, 3) * scale)) @ Z.rot_z(rng.uniform(0, 360)) @ Z.scale(scale, scale, scale)
            m1 = m0 @ Z.translate(*rng.uniform(-0.5, 0.5, 3)) @ Z.rot_x(rng.uniform(-170, 170))
            desc.add_motion(shapes[int(rng.integers(0, len(shapes)))], [0.0, 1.0], [Z.inv(m0), Z.inv(m1)])
    desc.cam = np.asarray([0, -12 * scale, 3 * scale, 0, 0, 0, 0, 0, 1, 50], dtype=np.float32)
    tree_type = [2, 4, 8][int(rng.integers(0, 3))]
    osc = H.oracle_scene(desc)
    mb = O.MBVH(osc, tree_type=tree_type)
    arr = H.mbvh_arrays(desc, mb)
    if rng.random() < 0.5:      # the host layer's SAH trees instead of the oracle's CLASSIC ones
        s = hostapi.Session({"accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": tree_type}, desc)
        s.build_accelerator("MBVH")
        arr["root_nodes"] = s.mbvh_root_nodes().copy()
        for i in range(s.mbvh_leaf_count()):
            arr["leaf_nodes"][i] = s.mbvh_leaf_nodes(i).copy()
        # the oracle walks the SAME arrays: exact ties (duplicated triangles) go to the first leaf in array order
        mb.set_root_nodes(arr["root_nodes"])
        for i in range(s.mbvh_leaf_count()):
            mb.set_leaf_nodes(i, arr["leaf_nodes"][i])
    emu = H.Emu.mbvh(arr)
    rays = random_rays(rng, desc, 2500, int(rng.integers(1, 1 << 30)))      # incl. rays starting on the instanced surfaces
    ref = mb.intersect(rays)
    kw = dict(libm_outlier_frac=0.0)
    got, st = emu.trace(rays, want_stats=True)
    assert st["max_stack"] <= emu.info()["stack_need"], ("stack bound", it, st["max_stack"], emu.info())    # non-finite rays included
    lockstep_check(rng, emu, rays, got, "fuzz2l %d" % it)
    ok = comparable(rays)
    got, ref, rays = got[ok], ref[ok], rays[ok]
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="fuzz2l %d k=%d" % (it, tree_type), two_level=True, max_ties=40, **kw)
    return "MBVH k=%d scale %8.2e objects %3d motion %d: hits %5d bit-exact %5d ties %d" % (
        tree_type, scale, len(desc.meshes), int(motion), rep["hits"], rep["bit_exact_hits"], rep["tie_exempt"])


def main():
    global LOCKSTEP, REFERENCE
    if "--reference" in sys.argv:
        REFERENCE = True
        sys.argv.remove("--reference")
    if "--lockstep" in sys.argv:
        LOCKSTEP = True
        sys.argv.remove("--lockstep")
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    t0 = time.time()
    it = 0
    while time.time() - t0 < seconds:
        CURRENT[0] = it
        line = (one_level if it % 3 else two_level)(rng, it)
        print("%4d %s" % (it, line), flush=True)
        it += 1
    print("fuzz: %d scenes, no disagreement" % it)


if __name__ == "__main__":
    main()
