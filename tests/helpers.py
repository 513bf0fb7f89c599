"""Shared test plumbing: SceneDesc -> oracle scene, emulation binding, hit comparison."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from luxcore_b200 import scenes as S  # noqa: E402
from oracle import oracle as O  # noqa: E402

HIT_DTYPE = O.HIT_DTYPE
RAY_DTYPE = O.RAY_DTYPE
NULL = 0xFFFFFFFF


def oracle_scene(desc):
    """SceneDesc -> oracle Scene: same shapes, same dataset order (index == meshIndex)."""
    sc = O.Scene()
    for v, t in desc.shapes:
        sc.add_shape(v, t)
    for m in desc.meshes:
        if m.kind == S.PLAIN:
            sc.add_plain(m.shape)
        elif m.kind == S.INSTANCE:
            sc.add_instance(m.shape, m.xform)
        else:
            sc.add_motion(m.shape, m.times, m.motion_xforms)
    return sc


def flattened_from_oracle(desc, osc):
    """World-space vertex buffer + per-mesh offsets exactly as BVHKernel builds them
    (bvhaccelhw.cpp:68-92), taken from the oracle's GetVertex so both sides see the same floats."""
    verts, offs, total = [], [], 0
    for i, m in enumerate(desc.meshes):
        n = desc.shapes[m.shape][0].shape[0]
        verts.append(osc.world_vertices(i, n))
        offs.append(total)
        total += n
    return np.concatenate(verts) if verts else np.zeros((0, 3), np.float32), np.asarray(offs, dtype=np.uint32)


TIE_EPS = 1e-5     # "two nearest hits within epsilon": |t1 - t2| <= TIE_EPS * max(1, t)   


def compare_hits(got, ref, rays=None, second_t=None, rel_tol=1e-5, what="", libm_outlier_frac=0.0, libm_outlier_tol=1e-3):
    """Parity rule of BASELINE.json: hit/miss, meshIndex and triangleIndex bit-exact; t, b1, b2
    within rel_tol relative (they are expected to be bit-identical wherever the arithmetic is
    identical).  When `second_t` (the second-nearest hit distance per ray, from the oracle's brute
    force) is given, an index mismatch is tolerated for rays whose two nearest hits lie within
    TIE_EPS in t -- the only exemption BASELINE.json allows; it is needed where libm and CUDA
    sinf/acosf differ by an ulp (motion blur).

    libm_outlier_frac > 0 (rotating motion-blur leaves only): the reference's Slerp calls the HOST
    libm's sinf/acosf, whose last bit differs between libm builds (glibc's are not correctly rounded
    in 1-5% of calls) and from any device implementation; a one-ulp change of the interpolated
    rotation moves b1/b2 of small distant triangles by more than 1e-5.  At most that fraction of the
    hits may exceed rel_tol, and none may exceed libm_outlier_tol.

    Returns a dict of counts; raises AssertionError with a readable report otherwise."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape
    miss_g = got["meshIndex"] == NULL
    miss_r = ref["meshIndex"] == NULL
    bad = miss_g != miss_r
    both = ~miss_g & ~miss_r
    bad |= both & ((got["meshIndex"] != ref["meshIndex"]) | (got["triangleIndex"] != ref["triangleIndex"]))
    n_tie_exempt = 0
    if second_t is not None and bad.any():
        with np.errstate(invalid="ignore"):
            tref = np.where(miss_r, got["t"], ref["t"])
            near = np.abs(second_t - tref) <= TIE_EPS * np.maximum(1.0, np.abs(tref))
            # a hit/miss flip is a tie only when the single hit sits within eps of the ray's maxt/mint limits
            exempt = bad & both & near & (np.abs(got["t"] - ref["t"]) <= TIE_EPS * np.maximum(1.0, np.abs(ref["t"])))
        n_tie_exempt = int(exempt.sum())
        bad &= ~exempt
        both = both & ~exempt

    def close(a, b):
        with np.errstate(invalid="ignore"):
            return np.abs(a - b) <= rel_tol * np.maximum(1.0, np.abs(b))
    val_bad = both & ~bad & ~(close(got["t"], ref["t"]) & close(got["b1"], ref["b1"]) & close(got["b2"], ref["b2"]))
    n_outliers = 0
    if libm_outlier_frac > 0 and val_bad.any():
        def close2(a, b):
            with np.errstate(invalid="ignore"):
                return np.abs(a - b) <= libm_outlier_tol * np.maximum(1.0, np.abs(b))
        within = val_bad & close2(got["t"], ref["t"]) & close2(got["b1"], ref["b1"]) & close2(got["b2"], ref["b2"])
        if within.sum() <= libm_outlier_frac * max(1, int(both.sum())):
            n_outliers = int(within.sum())
            val_bad &= ~within
    miss_t_bad = miss_g & miss_r & (got["t"] != ref["t"]) & ~(np.isnan(got["t"]) & np.isnan(ref["t"]))
    exact = both & ~bad & (got["t"] == ref["t"]) & (got["b1"] == ref["b1"]) & (got["b2"] == ref["b2"])
    rep = {"n": int(got.shape[0]), "hits": int(both.sum()), "index_mismatch": int(bad.sum()),
           "value_mismatch": int(val_bad.sum()), "miss_t_mismatch": int(miss_t_bad.sum()),
           "bit_exact_hits": int(exact.sum()), "tie_exempt": n_tie_exempt,
           "libm_outliers": n_outliers}
    if rep["index_mismatch"] or rep["value_mismatch"] or rep["miss_t_mismatch"]:
        idx = np.nonzero(bad | val_bad | miss_t_bad)[0][:8]
        lines = ["%s parity FAILED: %r" % (what, rep)]
        for i in idx:
            lines.append("  ray %d: got %r  ref %r%s" % (i, got[i], ref[i], ("  ray %r" % (rays[i],)) if rays is not None else ""))
        raise AssertionError("\n".join(lines))
    return rep


def compare_hits_tie_aware(got, ref, rays, osc, what="", two_level=False, max_ties=8, **kw):
    """compare_hits with north_star's exemption evaluated only where it is needed: for the rays whose
    indices differ, the oracle's brute-force pass supplies the distance of the second-nearest hit;
    rays whose two nearest hits lie within TIE_EPS of each other are exempt (and counted: at most
    `max_ties` of them may occur)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    diff = (got["meshIndex"] != ref["meshIndex"]) | ((got["triangleIndex"] != ref["triangleIndex"]) & (ref["meshIndex"] != NULL))
    second = None
    if diff.any():
        idx = np.nonzero(diff)[0]
        if idx.shape[0] > 2000:
            raise AssertionError("%s: %d index mismatches" % (what, idx.shape[0]))
        _, sec = osc.brute(np.ascontiguousarray(rays[idx]), two_level=two_level, want_second=True)
        second = np.full(got.shape[0], np.nan, dtype=np.float32)
        second[idx] = sec
    rep = compare_hits(got, ref, rays, what=what, second_t=second, **kw)
    assert rep["tie_exempt"] <= max_ties, "%s: %d tie-exempt rays" % (what, rep["tie_exempt"])
    return rep


def mbvh_arrays(desc, ombvh):
    """The arrays MBVHKernel hands to lrb_mbvh_upload, taken from the oracle's MBVHAccel restatement."""
    n = ombvh.leaf_count()
    leaf_nodes = [ombvh.leaf_nodes(i) for i in range(n)]
    leaf_verts = [desc.shapes[ombvh.leaf_mesh(i)][0] for i in range(n)]
    table, interps = ombvh.motion_systems()
    return {"root_nodes": ombvh.root_nodes(), "leaf_nodes": leaf_nodes, "leaf_verts": leaf_verts,
            "transforms_minv": ombvh.transforms_minv(), "motion_table": table, "interps": interps}


def fill_mbvh_desc(arr):
    """-> (capi.MBVHDesc, keepalive list)"""
    from luxcore_b200 import capi
    root = np.ascontiguousarray(arr["root_nodes"])
    ln = [np.ascontiguousarray(a) for a in arr["leaf_nodes"]]
    lv = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in arr["leaf_verts"]]
    n = len(ln)
    d = capi.MBVHDesc()
    d.root_nodes = root.ctypes.data
    d.n_root_nodes = root.shape[0]
    d.n_leaves = n
    a1 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in ln])
    a2 = (C.c_uint32 * max(n, 1))(*[a.shape[0] for a in ln])
    a3 = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in lv])
    a4 = (C.c_uint32 * max(n, 1))(*[a.shape[0] for a in lv])
    d.leaf_nodes = C.cast(a1, C.POINTER(C.c_void_p))
    d.leaf_n_nodes = C.cast(a2, C.POINTER(C.c_uint32))
    d.leaf_vertices = C.cast(a3, C.POINTER(C.c_void_p))
    d.leaf_n_vertices = C.cast(a4, C.POINTER(C.c_uint32))
    keep = [root, ln, lv, a1, a2, a3, a4]
    tm = np.ascontiguousarray(arr["transforms_minv"], dtype=np.float32).reshape(-1, 16)
    if tm.shape[0]:
        d.transforms_minv = tm.ctypes.data
        d.n_transforms = tm.shape[0]
        keep.append(tm)
    mt = np.ascontiguousarray(arr["motion_table"], dtype=np.uint32).reshape(-1, 4)
    if mt.shape[0]:
        it = np.ascontiguousarray(arr["interps"], dtype=np.uint8)
        d.motion_systems = mt.ctypes.data
        d.n_motion_systems = mt.shape[0]
        d.interpolated_transforms = it.ctypes.data
        d.n_interpolated_transforms = it.shape[0] // 576
        keep += [mt, it]
    return d, keep


class Emu:
    """CPU emulation of the product's re-layout + traversal body (tests/cpp/wide_emulation.cpp)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            so = os.path.join(ROOT, "tests", "cpp", "libwide_emulation.so")
            src = [os.path.join(ROOT, "tests", "cpp", "wide_emulation.cpp"),
                   os.path.join(ROOT, "luxcore_b200", "csrc", "relayout.cpp")]
            deps = src + [os.path.join(ROOT, "luxcore_b200", "csrc", f) for f in ("traverse.h", "layout.h", "relayout.h")]
            if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-msse", "-msse2", "-mfma", "-ffp-contract=off"] + (["-DLRB_EXIT_ORDER=" + os.environ["LRB_EXIT_ORDER"]] if "LRB_EXIT_ORDER" in os.environ else []) + [
                                       "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "luxcore_b200", "csrc"),
                                       "-o", so] + src)
            L = C.CDLL(so)
            L.emu_last_error.restype = C.c_char_p
            L.emu_bvh_create.restype = C.c_void_p
            L.emu_bvh_create.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
            L.emu_mbvh_create.restype = C.c_void_p
            L.emu_mbvh_create.argtypes = [C.c_void_p]
            L.emu_mbvh_update.restype = C.c_int
            L.emu_mbvh_update.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
            L.emu_free.argtypes = [C.c_void_p]
            L.emu_info.argtypes = [C.c_void_p, C.c_void_p]
            L.emu_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
            for f in (L.emu_copy_nodes, L.emu_copy_tris, L.emu_copy_ids):
                f.argtypes = [C.c_void_p, C.c_void_p]
            L.emu_warp_sim.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
            L.emu_scene_view_size.restype = C.c_uint32
            L.emu_scene_view.argtypes = [C.c_void_p, C.c_void_p]
            L.emu_validate_tree.restype = C.c_int
            L.emu_validate_tree.argtypes = [C.c_void_p, C.c_uint32]
            cls._lib = L
        return cls._lib

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(self.lib().emu_last_error().decode())
        self.h = handle

    @classmethod
    def bvh(cls, nodes, verts, offs):
        nodes = np.ascontiguousarray(nodes)
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        offs = np.ascontiguousarray(offs, dtype=np.uint32)
        return cls(cls.lib().emu_bvh_create(nodes.ctypes.data, nodes.shape[0], verts.ctypes.data, verts.shape[0],
                                            offs.ctypes.data, offs.shape[0]))

    @classmethod
    def mbvh(cls, arr):
        d, keep = fill_mbvh_desc(arr)
        e = cls(cls.lib().emu_mbvh_create(C.byref(d)))
        del keep
        return e

    def update(self, root_nodes, minv):
        root_nodes = np.ascontiguousarray(root_nodes)
        minv = np.ascontiguousarray(minv, dtype=np.float32).reshape(-1, 16)
        if self.lib().emu_mbvh_update(self.h, root_nodes.ctypes.data, root_nodes.shape[0], minv.ctypes.data, minv.shape[0]) != 0:
            raise RuntimeError(self.lib().emu_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().emu_free(self.h)
            self.h = None

    def info(self):
        a = np.zeros(6, dtype=np.uint32)
        self.lib().emu_info(self.h, a.ctypes.data)
        return dict(zip(["wide", "tris", "insts", "stack_need", "root", "two_level"], [int(x) for x in a]))

    WIDE_DTYPE = np.dtype([("org", "<f4", 3), ("exps", "<u4"), ("child", "<u4", 4), ("qlo", "<u4", 3), ("qhi", "<u4", 3),
                           ("next", "<u4"), ("flags", "<u4")])
    TRI_DTYPE = np.dtype([("p0", "<f4", 3), ("p1", "<f4", 3), ("p2", "<f4", 3), ("gateLo", "<f4", 3), ("gateHi", "<f4", 3), ("order", "<u4")])
    IDS_DTYPE = np.dtype([("meshIndex", "<u4"), ("triangleIndex", "<u4")])

    def arrays(self):
        """The re-laid-out arrays exactly as they would be uploaded: (wide nodes, triangle records, triangle ids)."""
        assert self.WIDE_DTYPE.itemsize == 64 and self.TRI_DTYPE.itemsize == 64 and self.IDS_DTYPE.itemsize == 8
        i = self.info()
        wide = np.zeros(i["wide"], dtype=self.WIDE_DTYPE)
        tris = np.zeros(i["tris"], dtype=self.TRI_DTYPE)
        ids = np.zeros(i["tris"], dtype=self.IDS_DTYPE)
        if i["wide"]:
            self.lib().emu_copy_nodes(self.h, wide.ctypes.data)
        if i["tris"]:
            self.lib().emu_copy_tris(self.h, tris.ctypes.data)
            self.lib().emu_copy_ids(self.h, ids.ctypes.data)
        return wide, tris, ids

    def trace(self, rays, want_stats=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = np.zeros(6, dtype=np.uint64)
        self.lib().emu_trace(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], st.ctypes.data)
        if want_stats:
            return hits, dict(zip(["rays", "wide_nodes", "triangles", "instances", "motion_samples", "max_stack"], [int(x) for x in st]))
        return hits


def chunk_ends(n, chunk=1 << 20, min_chunk=1 << 16, taper=True):
    """Python restatement of luxcore_b200/csrc/host_chunks.h ChunkEnds (held to it by tests/test_host_chunks_cpu.py): the
    pieces a batch from host memory is cut into by lrb_trace_host and by the pipelined plugin sequence."""
    chunk = max(1, chunk)
    min_chunk = min(max(1, min_chunk), chunk)
    ends, pos = [], 0
    while pos < n:
        rem = n - pos
        c = chunk
        if taper and n > chunk and rem <= 2 * chunk:
            c = (rem // 2 + 1023) & ~1023
            c = min(max(c, min_chunk), chunk)
            if rem <= min_chunk:
                c = rem
        c = min(c, rem)
        pos += c
        ends.append(pos)
    return ends


def flattened_triangles(desc):
    """Triangle index buffer + per-mesh triangle offsets (n_meshes + 1 entries) in dataset order: what
    lrb_bvh_build_scene takes next to flattened_from_oracle's vertices (mesh-local indices, luxrays::Triangle::v)."""
    tris, offs, total = [], [0], 0
    for m in desc.meshes:
        t = np.ascontiguousarray(desc.shapes[m.shape][1], dtype=np.uint32).reshape(-1, 3)
        tris.append(t)
        total += t.shape[0]
        offs.append(total)
    return (np.concatenate(tris) if tris else np.zeros((0, 3), np.uint32)), np.asarray(offs, dtype=np.uint32)


def to_builder_format(nodes, mesh_tri_offsets):
    """A reference array as the GPU builder kernels emit it BEFORE the leaf payload is written: leaves carry the number of
    their triangle in the concatenated triangle buffer in the first word and zeros elsewhere (build_kernels.cuh EmitKernel)."""
    out = np.array(nodes, copy=True)
    leaf = (out["nodeData"] & 0x80000000) != 0
    w = out["w"]
    g = np.asarray(mesh_tri_offsets, dtype=np.uint32)[w[leaf, 3]] + w[leaf, 4]
    w[leaf] = 0
    w[leaf, 0] = g
    out["pad0"] = 0
    return out


class RelayoutDev:
    """The per-record bodies of the DEVICE re-layout (luxcore_b200/csrc/relayout_kernels.cuh, relayout_shared.h) compiled for
    the host and driven in the device pipeline's order by tests/cpp/relayout_device_emulation.cpp."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            d = os.path.join(ROOT, "tests", "cpp")
            csrc = os.path.join(ROOT, "luxcore_b200", "csrc")
            so = os.path.join(d, "librelayout_device_emulation.so")
            src = os.path.join(d, "relayout_device_emulation.cpp")
            deps = [src] + [os.path.join(csrc, f) for f in ("relayout_kernels.cuh", "relayout_shared.h", "layout.h")]
            if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-msse", "-msse2", "-mfma", "-ffp-contract=off",
                                       "-I" + os.path.join(ROOT, "include"), "-I" + csrc, "-o", so, src])
            L = C.CDLL(so)
            L.rde_last_error.restype = C.c_char_p
            L.rde_run.restype = C.c_void_p
            L.rde_run.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
            L.rde_free.argtypes = [C.c_void_p]
            for f in (L.rde_info, L.rde_copy_nodes, L.rde_copy_tris, L.rde_copy_ids, L.rde_copy_ref_nodes, L.rde_copy_boxes, L.rde_copy_entry_box):
                f.argtypes = [C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    @classmethod
    def run(cls, builder_nodes, verts, mesh_vert_offsets, tri_idx, mesh_tri_offsets):
        """-> dict(wide, tris, ids, ref_nodes, boxes, stack_need, entry_box)"""
        L = cls.lib()
        nodes = np.ascontiguousarray(builder_nodes)
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        voff = np.ascontiguousarray(mesh_vert_offsets, dtype=np.uint32)
        toff = np.ascontiguousarray(mesh_tri_offsets, dtype=np.uint32)
        tri = np.ascontiguousarray(tri_idx, dtype=np.uint32).reshape(-1, 3)
        h = L.rde_run(nodes.ctypes.data, nodes.shape[0], verts.ctypes.data, verts.shape[0], voff.ctypes.data, toff.ctypes.data, voff.shape[0], tri.ctypes.data)
        if not h:
            raise RuntimeError(L.rde_last_error().decode())
        try:
            info = np.zeros(3, dtype=np.uint32)
            L.rde_info(h, info.ctypes.data)
            wide = np.zeros(int(info[0]), dtype=Emu.WIDE_DTYPE)
            tris = np.zeros(int(info[1]), dtype=Emu.TRI_DTYPE)
            ids = np.zeros(int(info[1]), dtype=Emu.IDS_DTYPE)
            ref = np.zeros(nodes.shape[0], dtype=nodes.dtype)
            boxes = np.zeros((tri.shape[0], 6), dtype=np.float32)
            entry = np.zeros(6, dtype=np.float32)
            L.rde_copy_nodes(h, wide.ctypes.data)
            L.rde_copy_tris(h, tris.ctypes.data)
            L.rde_copy_ids(h, ids.ctypes.data)
            L.rde_copy_ref_nodes(h, ref.ctypes.data)
            L.rde_copy_boxes(h, boxes.ctypes.data)
            L.rde_copy_entry_box(h, entry.ctypes.data)
        finally:
            L.rde_free(h)
        return {"wide": wide, "tris": tris, "ids": ids, "ref_nodes": ref, "boxes": boxes, "stack_need": int(info[2]), "entry_box": entry}


class Lockstep:
    """The product's REAL kernel source (luxcore_b200/csrc/trace_kernels.cuh) compiled for the host against a
    stand-in <cuda_runtime.h> and run with one OS thread per lane (tests/cpp/kernel_lockstep.cpp)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            d = os.path.join(ROOT, "tests", "cpp")
            so = os.path.join(d, "libkernel_lockstep.so")
            csrc = os.path.join(ROOT, "luxcore_b200", "csrc")
            deps = [os.path.join(d, "kernel_lockstep.cpp"), os.path.join(d, "fakecuda", "cuda_runtime.h")] + \
                   [os.path.join(csrc, f) for f in ("trace_kernels.cuh", "traverse.h", "layout.h")]
            if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in deps):
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-msse", "-msse2", "-mfma", "-ffp-contract=off",
                                       "-D__CUDACC__", "-I" + os.path.join(d, "fakecuda"), "-I" + os.path.join(ROOT, "include"), "-I" + csrc,
                                       "-o", so, deps[0]])
            L = C.CDLL(so)
            L.ks_trace.restype = C.c_int
            L.ks_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_uint32,
                                   C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
            L.ks_trace_signal.restype = C.c_int
            L.ks_trace_signal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32,
                                          C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    @classmethod
    def trace_signal(cls, emu, rays, n_warps=4, smem_depth=16, chunk_shift=7, epoch=5, hits=None):
        """The SIGNAL kernel (trace of the multi-GPU gather): -> (hits, peer copy, chunk flags, flags raised while the
        kernel ran, snapshots-at-flag == final records)."""
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE) if hits is None else np.ascontiguousarray(hits).copy()
        peer = out.copy()
        E = Emu.lib()
        view = (C.c_uint8 * int(E.emu_scene_view_size()))()
        E.emu_scene_view(emu.h, view)
        n_chunks = (rays.shape[0] + (1 << chunk_shift) - 1) >> chunk_shift
        flags = np.zeros(max(1, n_chunks), dtype=np.uint32)
        res = np.zeros(2, dtype=np.uint32)
        rc = cls.lib().ks_trace_signal(view, rays.ctypes.data, out.ctypes.data, peer.ctypes.data, rays.shape[0], n_warps, smem_depth,
                                       emu.info()["stack_need"], chunk_shift, epoch, flags.ctypes.data, res.ctypes.data)
        if rc != 0:
            raise RuntimeError("ks_trace_signal failed: %d" % rc)
        return out, peer, flags[:n_chunks], int(res[1]), bool(res[0])

    @classmethod
    def trace(cls, emu, rays, kernel="persistent", n_warps=1, smem_depth=16, refill_below=24, tri_bias=8, inst_bias=8,
              prefetch=False, hits=None, want_stats=False, anyhit=False):
        """-> hits (and TraceStatic's counters with want_stats).  `hits` pre-loads the RayHit buffer (masked rays)."""
        rays = np.ascontiguousarray(rays)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE) if hits is None else np.ascontiguousarray(hits).copy()
        E = Emu.lib()
        view = (C.c_uint8 * int(E.emu_scene_view_size()))()
        E.emu_scene_view(emu.h, view)
        st = np.zeros(6, dtype=np.uint64)
        rc = cls.lib().ks_trace(view, rays.ctypes.data, out.ctypes.data, rays.shape[0], 0 if kernel == "persistent" else 1, n_warps,
                                smem_depth, emu.info()["stack_need"], refill_below, tri_bias, inst_bias, (1 if prefetch else 0) | (2 if anyhit else 0), st.ctypes.data)
        if rc != 0:
            raise RuntimeError("ks_trace failed: %d" % rc)
        if want_stats:
            return out, dict(zip(["rays", "wide_nodes", "triangles", "instances", "motion_samples", "max_stack"], [int(x) for x in st]))
        return out


WARP_SIM_FIELDS = ["rays", "outer_iters", "inner_iters", "node_phases", "tri_phases", "node_lanes", "tri_lanes", "pop_trips",
                   "pop_lanes", "gate_phases", "gate_lanes", "store_phases", "refills", "waiting_lanes", "slow_push_phases", "slow_push_lanes", "instance_trips", "instance_lanes", "leave_trips", "leave_lanes"]


def warp_sim(emu, rays, n_warps=64, refill_below=24, tri_bias=8, inst_bias=8, spec_pop=True, tri_min=0):
    """Scheduling model of TracePersistent (tests/cpp/wide_emulation.cpp WarpSim): -> (hits, counts).
    spec_pop=False: the loop without the speculative pop at the end of the phases (the round-1 kernel)."""
    tri_bias = (tri_bias & 0xFFFF) | ((inst_bias & 0x7FFF) << 16) | (0 if spec_pop else 1 << 31)
    refill_below = (refill_below & 0xFF) | ((tri_min & 0xFF) << 8)       # experimental "both phases per iteration" policy
    rays = np.ascontiguousarray(rays)
    hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    out = np.zeros(20, dtype=np.uint64)
    emu.lib().emu_warp_sim(emu.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], n_warps, refill_below, tri_bias, out.ctypes.data)
    return hits, dict(zip(WARP_SIM_FIELDS, [int(x) for x in out]))


def triangle_intersect_np(rays, p0, e1, e2):
    """Triangle::Intersect (include/luxrays/core/geometry/triangle.h:55-89) for ray i against triangle i, in
    float32 numpy arithmetic (every operation rounded to float32, none fused: the reference's own floats).
    p0 / e1 / e2: [n,3] float32 (e1 = p1 - p0, e2 = p2 - p0 as float32 differences).  -> (hit, t, b1, b2)."""
    f = np.float32
    o, d = np.asarray(rays["o"], f), np.asarray(rays["d"], f)
    p0 = np.asarray(p0, f); e1 = np.asarray(e1, f); e2 = np.asarray(e2, f)

    def cross(a, b):
        return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2], a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1)

    def dot(a, b):
        return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]

    with np.errstate(all="ignore"):
        s1 = cross(d, e2)
        div = dot(s1, e1)
        inv = f(1.0) / div
        dd = o - p0
        b1 = dot(dd, s1) * inv
        s2 = cross(dd, e1)
        b2 = dot(d, s2) * inv
        b0 = (f(1.0) - b1) - b2
        t = dot(e2, s2) * inv
        mint = np.where(np.isnan(rays["mint"]), f(-np.inf), rays["mint"]).astype(f)
        hit = (div != 0) & ~(b1 < 0) & ~(b2 < 0) & ~(b0 < 0) & ~(t < mint) & ~(t > rays["maxt"])
    return hit, t.astype(f), b1.astype(f), b2.astype(f)


def machine_epsilon_np(v):
    """MachineEpsilon::E(float) (include/luxrays/core/epsilon.h:48-53,75-82) on a float32 array."""
    v = np.asarray(v, np.float32)
    with np.errstate(all="ignore"):
        nxt = (v.view(np.uint32) + np.uint32(0x80)).view(np.float32)
        e = np.abs(nxt - v)
    return np.where(e < np.float32(1e-5), np.float32(1e-5), np.where(e > np.float32(1e-1), np.float32(1e-1), e)).astype(np.float32)


def reference_scene(desc):
    """SceneDesc -> the REFERENCE's own mesh objects (oracle/_ref, oracle/refapi.py), same dataset order."""
    from oracle import refapi as RF
    sc = RF.Scene()
    for v, t in desc.shapes:
        sc.add_shape(v, t)
    for m in desc.meshes:
        if m.kind == S.PLAIN:
            sc.add_plain(m.shape)
        elif m.kind == S.INSTANCE:
            sc.add_instance(m.shape, m.xform)
        else:
            sc.add_motion(m.shape, m.times, m.motion_xforms)
    return sc


def check_anyhit(got, closest, rays, desc=None, what=""):
    """Contract of lrb_trace_anyhit: hit / miss identical to the closest-hit answer; a miss carries the closest-hit
    miss payload; a hit names a triangle the ray really hits inside [mint, maxt] at exactly the reported t / b1 / b2
    (checked with the reference's own arithmetic when `desc` is a one-level scene), never closer than the closest hit."""
    miss_g, miss_c = got["meshIndex"] == NULL, closest["meshIndex"] == NULL
    assert (miss_g == miss_c).all(), "%s: any-hit hit/miss differs from closest-hit on %d rays" % (what, int((miss_g != miss_c).sum()))
    # miss payload: t = ray.maxt, meshIndex = NULL (b1 / b2 / triangleIndex of a miss are unspecified in the reference)
    assert (got["t"][miss_g].view(np.uint32) == closest["t"][miss_c].view(np.uint32)).all(), "%s: miss payload" % what
    h = ~miss_g
    assert (got["t"][h] >= closest["t"][h]).all(), "%s: any-hit closer than the closest hit" % what
    if desc is not None and h.any():
        p0, e1, e2, offs = S.world_triangles(desc)
        flat = offs[got["meshIndex"][h].astype(np.int64)] + got["triangleIndex"][h].astype(np.int64)
        ok, t, b1, b2 = triangle_intersect_np(rays[h], p0[flat], e1[flat], e2[flat])
        assert ok.all(), "%s: %d any-hit records name a triangle the ray does not hit" % (what, int((~ok).sum()))
        assert (t.view(np.uint32) == got["t"][h].view(np.uint32)).all() and (b1.view(np.uint32) == got["b1"][h].view(np.uint32)).all() \
            and (b2.view(np.uint32) == got["b2"][h].view(np.uint32)).all(), "%s: any-hit t / b1 / b2 not bit-exact" % what
    return {"rays": int(got.shape[0]), "hits": int(h.sum()), "same_as_closest": int((got[h] == closest[h]).sum())}


def in_plane_rays(desc, n, seed):
    """Adversarial batch for the one known parity residual (DESIGN.md "Parity"): rays lying IN THE PLANE of a scene
    triangle, starting outside it, running along in-plane directions (a third of them exactly along an edge).  For a
    triangle in general position Triangle::Intersect then divides rounding noise by rounding noise and the reference can
    report a hit the ray misses by far -- whenever it gets to test the triangle."""
    p0, e1, e2, _ = S.world_triangles(desc)
    rng = np.random.default_rng(seed)
    tri = rng.integers(0, p0.shape[0], n)
    P0, E1, E2 = p0[tri].astype(np.float64), e1[tri].astype(np.float64), e2[tri].astype(np.float64)
    a, b = rng.uniform(-1.5, 2.5, n), rng.uniform(-1.5, 2.5, n)
    inside = (a >= -0.05) & (b >= -0.05) & (a + b <= 1.05)
    a[inside] += 1.5
    o = P0 + a[:, None] * E1 + b[:, None] * E2
    d = rng.standard_normal(n)[:, None] * E1 + rng.standard_normal(n)[:, None] * E2
    mode = rng.integers(0, 3, n)
    d[mode == 1] = E1[mode == 1]
    d[mode == 2] = E2[mode == 2]
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-30
    rays = np.zeros(n, dtype=RAY_DTYPE)
    rays["o"] = o.astype(np.float32)
    rays["d"] = d.astype(np.float32)
    rays["mint"] = 1e-5
    rays["maxt"] = np.inf
    return rays


def index_differences(got, ref):
    hit = ref["meshIndex"] != NULL
    return int(((got["meshIndex"] != ref["meshIndex"]) | (hit & (got["triangleIndex"] != ref["triangleIndex"]))).sum())


def coplanar_with_reported(desc, rays, rec, tol=1e-3):
    """True where the ray lies in the plane of the triangle `rec` names, to within `tol` radians and `tol` x scene size
    (float64 geometry).  1e-3: on the kitchen's needle triangles (0.3 long, 2e-3 wide) a ray 3e-4 rad off the plane still
    gets a t from Triangle::Intersect that is 3 % off its float64 value (observed: round-2 GPU run, seed 2)."""
    p0, e1, e2, offs = S.world_triangles(desc)
    ok = rec["meshIndex"] != NULL
    flat = np.where(ok, offs[np.minimum(rec["meshIndex"], len(offs) - 1).astype(np.int64)] + rec["triangleIndex"].astype(np.int64), 0)
    n = np.cross(e1[flat].astype(np.float64), e2[flat].astype(np.float64))
    n /= np.linalg.norm(n, axis=1, keepdims=True) + 1e-300
    d = rays["d"].astype(np.float64)
    d /= np.linalg.norm(d, axis=1, keepdims=True) + 1e-300
    lo, hi = desc.bbox()
    scale = float(np.linalg.norm(hi - lo))
    sin_to_plane = np.abs((n * d).sum(1))
    dist = np.abs((n * (rays["o"].astype(np.float64) - p0[flat])).sum(1))
    return ok & (sin_to_plane < tol) & (dist < tol * scale)
