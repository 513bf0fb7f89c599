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


def compare_hits(got, ref, rays=None, brute_second=None, rel_tol=1e-5, what=""):
    """Parity rule of BASELINE.json: hit/miss, meshIndex and triangleIndex bit-exact; t, b1, b2
    within rel_tol relative (they are expected to be bit-identical for the single-level BVH).
    Returns a dict of mismatch counts; raises AssertionError with a readable report otherwise."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape
    miss_g = got["meshIndex"] == NULL
    miss_r = ref["meshIndex"] == NULL
    bad = miss_g != miss_r
    both = ~miss_g & ~miss_r
    bad |= both & ((got["meshIndex"] != ref["meshIndex"]) | (got["triangleIndex"] != ref["triangleIndex"]))

    def close(a, b):
        return np.abs(a - b) <= rel_tol * np.maximum(1.0, np.abs(b))
    val_bad = both & ~bad & ~(close(got["t"], ref["t"]) & close(got["b1"], ref["b1"]) & close(got["b2"], ref["b2"]))
    miss_t_bad = miss_g & miss_r & (got["t"] != ref["t"]) & ~(np.isnan(got["t"]) & np.isnan(ref["t"]))
    exact = both & ~bad & (got["t"] == ref["t"]) & (got["b1"] == ref["b1"]) & (got["b2"] == ref["b2"])
    rep = {"n": int(got.shape[0]), "hits": int(both.sum()), "index_mismatch": int(bad.sum()),
           "value_mismatch": int(val_bad.sum()), "miss_t_mismatch": int(miss_t_bad.sum()),
           "bit_exact_hits": int(exact.sum())}
    if rep["index_mismatch"] or rep["value_mismatch"] or rep["miss_t_mismatch"]:
        idx = np.nonzero(bad | val_bad | miss_t_bad)[0][:8]
        lines = ["%s parity FAILED: %r" % (what, rep)]
        for i in idx:
            lines.append("  ray %d: got %r  ref %r%s" % (i, got[i], ref[i], ("  ray %r" % (rays[i],)) if rays is not None else ""))
        raise AssertionError("\n".join(lines))
    return rep


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
                subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-msse", "-msse2", "-ffp-contract=off",
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

    def __del__(self):
        if getattr(self, "h", None):
            self.lib().emu_free(self.h)
            self.h = None

    def info(self):
        a = np.zeros(6, dtype=np.uint32)
        self.lib().emu_info(self.h, a.ctypes.data)
        return dict(zip(["wide", "tris", "insts", "stack_need", "root", "two_level"], [int(x) for x in a]))

    def trace(self, rays, want_stats=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        st = np.zeros(6, dtype=np.uint64)
        self.lib().emu_trace(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], st.ctypes.data)
        if want_stats:
            return hits, dict(zip(["rays", "wide_nodes", "triangles", "instances", "motion_samples", "max_stack"], [int(x) for x in st]))
        return hits
