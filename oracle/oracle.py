"""ctypes binding of the CPU oracle (oracle/lux_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never from luxcore_b200/.  See the header of lux_oracle.cpp for what it restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "build", "liblux_oracle.so")

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("mint", "<f4"), ("maxt", "<f4"), ("time", "<f4"),
                      ("flags", "<u4"), ("pad", "<f4", 2)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("b1", "<f4"), ("b2", "<f4"), ("meshIndex", "<u4"), ("triangleIndex", "<u4")])
NODE_DTYPE = np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])
assert RAY_DTYPE.itemsize == 48 and HIT_DTYPE.itemsize == 20 and NODE_DTYPE.itemsize == 32
NULL_INDEX = 0xFFFFFFFF


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "lux_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_set_epsilon": (None, [f32, f32]),
            "orc_epsilon": (f32, [f32]),
            "orc_matrix_inverse": (i32, [vp, vp]),
            "orc_triangle_intersect": (i32, [vp, vp, vp, vp, vp]),
            "orc_bbox_intersectp": (i32, [vp, vp, vp]),
            "orc_scene_create": (vp, []),
            "orc_scene_free": (None, [vp]),
            "orc_scene_mesh_count": (i32, [vp]),
            "orc_scene_add_shape": (i32, [vp, vp, u32, vp, u32]),
            "orc_scene_add_plain": (i32, [vp, i32]),
            "orc_scene_add_instance": (i32, [vp, i32, vp]),
            "orc_scene_add_motion": (i32, [vp, i32, u32, vp, vp]),
            "orc_scene_set_instance_transform": (i32, [vp, i32, vp]),
            "orc_scene_mesh_bbox": (i32, [vp, i32, vp]),
            "orc_scene_mesh_world_vertices": (i32, [vp, i32, vp]),
            "orc_bvh_build": (vp, [vp, i32, i32, i32, i32, f32]),
            "orc_bvh_from_nodes": (vp, [vp, vp, u32]),
            "orc_bvh_free": (None, [vp]),
            "orc_bvh_node_count": (u32, [vp]),
            "orc_bvh_nodes": (vp, [vp]),
            "orc_bvh_intersect": (i32, [vp, vp, vp, u64, i32, vp]),
            "orc_mbvh_build": (vp, [vp, i32, i32, i32, i32, f32]),
            "orc_mbvh_free": (None, [vp]),
            "orc_mbvh_update": (i32, [vp]),
            "orc_mbvh_root_node_count": (u32, [vp]),
            "orc_mbvh_root_nodes": (vp, [vp]),
            "orc_mbvh_set_root_nodes": (None, [vp, vp, u32]),
            "orc_mbvh_leaf_count": (u32, [vp]),
            "orc_mbvh_leaf_node_count": (u32, [vp, u32]),
            "orc_mbvh_leaf_nodes": (vp, [vp, u32]),
            "orc_mbvh_set_leaf_nodes": (None, [vp, u32, vp, u32]),
            "orc_mbvh_leaf_mesh": (i32, [vp, u32]),
            "orc_mbvh_transform_count": (u32, [vp]),
            "orc_mbvh_transform_minv": (None, [vp, u32, vp]),
            "orc_mbvh_motion_count": (u32, [vp]),
            "orc_mbvh_motion_interp_count": (u32, [vp, u32]),
            "orc_mbvh_motion_interps": (vp, [vp, u32]),
            "orc_motion_sample": (i32, [vp, u32, f32, vp]),
            "orc_mbvh_intersect": (i32, [vp, vp, vp, u64, i32, vp]),
            "orc_brute": (i32, [vp, i32, vp, vp, vp, u64, i32]),
            "orc_hardware_threads": (i32, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _err():
    return lib().orc_last_error().decode()


def hardware_threads():
    return max(1, lib().orc_hardware_threads())


def copy_nodes(ptr, n):
    if n == 0:
        return np.zeros(0, dtype=NODE_DTYPE)
    buf = (C.c_char * (32 * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=NODE_DTYPE).copy()


class Scene:
    """DataSet restatement: meshes in Add() order; index == meshIndex reported in RayHit."""

    def __init__(self):
        self.h = lib().orc_scene_create()
        self.kinds = []
        self.base = []

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_scene_free(self.h)
            self.h = None

    def add_shape(self, verts, tris):
        """A TriangleMesh (geometry only)."""
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
        return lib().orc_scene_add_shape(self.h, _ptr(v), v.shape[0], _ptr(t), t.shape[0])

    def add_plain(self, shape):
        if lib().orc_scene_add_plain(self.h, shape) != 0:
            raise RuntimeError(_err())
        self.kinds.append(0)
        self.base.append(shape)
        return len(self.kinds) - 1

    def add_mesh(self, verts, tris):
        """Convenience: new shape + plain dataset entry."""
        return self.add_plain(self.add_shape(verts, tris))

    def add_instance(self, shape, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        if lib().orc_scene_add_instance(self.h, shape, _ptr(m)) != 0:
            raise RuntimeError(_err())
        self.kinds.append(1)
        self.base.append(shape)
        return len(self.kinds) - 1

    def add_motion(self, shape, times, mats):
        t = np.ascontiguousarray(times, dtype=np.float32)
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 4, 4)
        assert m.shape[0] == t.shape[0]
        if lib().orc_scene_add_motion(self.h, shape, t.shape[0], _ptr(t), _ptr(m)) != 0:
            raise RuntimeError(_err())
        self.kinds.append(2)
        self.base.append(shape)
        return len(self.kinds) - 1

    def set_instance_transform(self, mesh, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        if lib().orc_scene_set_instance_transform(self.h, mesh, _ptr(m)) != 0:
            raise RuntimeError(_err())

    def mesh_count(self):
        return lib().orc_scene_mesh_count(self.h)

    def mesh_bbox(self, i):
        out = np.zeros(6, dtype=np.float32)
        if lib().orc_scene_mesh_bbox(self.h, i, _ptr(out)) != 0:
            raise RuntimeError(_err())
        return out

    def world_vertices(self, i, nverts):
        out = np.zeros((nverts, 3), dtype=np.float32)
        n = lib().orc_scene_mesh_world_vertices(self.h, i, _ptr(out))
        assert n == nverts
        return out

    def brute(self, rays, two_level=False, nthreads=None, want_second=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        second = np.zeros(rays.shape[0], dtype=np.float32) if want_second else None
        lib().orc_brute(self.h, 1 if two_level else 0, _ptr(rays), _ptr(hits), _ptr(second) if want_second else None,
                        rays.shape[0], nthreads or hardware_threads())
        return (hits, second) if want_second else hits


class BVH:
    """BVHAccel restatement (CLASSIC builder, or wrapping an externally built node array)."""

    def __init__(self, scene, tree_type=4, cost_samples=0, isect_cost=80, trav_cost=10, empty_bonus=0.5, nodes=None):
        self.scene = scene
        if nodes is None:
            self.h = lib().orc_bvh_build(scene.h, tree_type, cost_samples, isect_cost, trav_cost, empty_bonus)
        else:
            nodes = np.ascontiguousarray(nodes)
            assert nodes.dtype.itemsize == 32
            self.h = lib().orc_bvh_from_nodes(scene.h, _ptr(nodes), nodes.shape[0])
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_bvh_free(self.h)
            self.h = None

    def nodes(self):
        return copy_nodes(lib().orc_bvh_nodes(self.h), lib().orc_bvh_node_count(self.h))

    def intersect(self, rays, nthreads=None, count=False):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        cnt = np.zeros(4, dtype=np.uint64) if count else None
        lib().orc_bvh_intersect(self.h, _ptr(rays), _ptr(hits), rays.shape[0], nthreads or hardware_threads(),
                                _ptr(cnt) if count else None)
        return (hits, cnt) if count else hits


class MBVH:
    """MBVHAccel restatement."""

    def __init__(self, scene, tree_type=4, cost_samples=0, isect_cost=80, trav_cost=10, empty_bonus=0.5):
        self.scene = scene
        self.h = lib().orc_mbvh_build(scene.h, tree_type, cost_samples, isect_cost, trav_cost, empty_bonus)
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_mbvh_free(self.h)
            self.h = None

    def update(self):
        if lib().orc_mbvh_update(self.h) != 0:
            raise RuntimeError(_err())

    def root_nodes(self):
        return copy_nodes(lib().orc_mbvh_root_nodes(self.h), lib().orc_mbvh_root_node_count(self.h))

    def set_root_nodes(self, nodes):
        nodes = np.ascontiguousarray(nodes)
        lib().orc_mbvh_set_root_nodes(self.h, _ptr(nodes), nodes.shape[0])

    def leaf_count(self):
        return lib().orc_mbvh_leaf_count(self.h)

    def leaf_nodes(self, i):
        return copy_nodes(lib().orc_mbvh_leaf_nodes(self.h, i), lib().orc_mbvh_leaf_node_count(self.h, i))

    def set_leaf_nodes(self, i, nodes):
        nodes = np.ascontiguousarray(nodes)
        lib().orc_mbvh_set_leaf_nodes(self.h, i, _ptr(nodes), nodes.shape[0])

    def leaf_mesh(self, i):
        return lib().orc_mbvh_leaf_mesh(self.h, i)

    def transforms_minv(self):
        n = lib().orc_mbvh_transform_count(self.h)
        out = np.zeros((n, 4, 4), dtype=np.float32)
        for i in range(n):
            lib().orc_mbvh_transform_minv(self.h, i, C.c_void_p(out[i].ctypes.data))
        return out

    def motion_systems(self):
        """-> (first/last index table uint32 [n,4], flat uint8 [k*576] of ocl::InterpolatedTransform)."""
        n = lib().orc_mbvh_motion_count(self.h)
        table = np.full((n, 4), NULL_INDEX, dtype=np.uint32)
        blobs = []
        k = 0
        for i in range(n):
            c = lib().orc_mbvh_motion_interp_count(self.h, i)
            buf = (C.c_char * (576 * c)).from_address(lib().orc_mbvh_motion_interps(self.h, i))
            blobs.append(np.frombuffer(buf, dtype=np.uint8).copy())
            table[i, 0] = k
            table[i, 1] = k + c - 1
            k += c
        flat = np.concatenate(blobs) if blobs else np.zeros(0, dtype=np.uint8)
        return table, flat

    def motion_sample(self, i, time):
        out = np.zeros((4, 4), dtype=np.float32)
        lib().orc_motion_sample(self.h, i, time, _ptr(out))
        return out

    def intersect(self, rays, nthreads=None, count=False):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype == RAY_DTYPE
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        cnt = np.zeros(4, dtype=np.uint64) if count else None
        lib().orc_mbvh_intersect(self.h, _ptr(rays), _ptr(hits), rays.shape[0], nthreads or hardware_threads(),
                                 _ptr(cnt) if count else None)
        return (hits, cnt) if count else hits


def machine_epsilon(v):
    return float(lib().orc_epsilon(C.c_float(v)))


def matrix_inverse(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
    out = np.zeros((4, 4), dtype=np.float32)
    if lib().orc_matrix_inverse(_ptr(m), _ptr(out)) != 0:
        raise RuntimeError("Singular matrix in MatrixInvert")
    return out
