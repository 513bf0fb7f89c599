"""ctypes access to oracle/_ref/libluxrays_ref.so: the REFERENCE's own C++ sources (Triangle::Intersect,
BBox::IntersectP, CLASSIC builder, BVHAccel, MBVHAccel, Transform, Matrix4x4::Inverse, MotionSystem,
MachineEpsilon, the mesh classes) compiled from /root/reference by oracle/ref/Makefile.

TEST INFRASTRUCTURE: pins oracle/lux_oracle.cpp (tests/test_oracle_pinned_cpu.py, tools/make_ref_vectors.py)
and serves as the `reference` CPU baseline of bench.py.  The library is built in the container that
holds /root/reference and travels to the GPU box as a prebuilt file; available() says whether it is there.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libluxrays_ref.so")
REFERENCE_ROOT = os.environ.get("LUX_REFERENCE_ROOT", "/root/reference")
_lib = None

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("mint", "<f4"), ("maxt", "<f4"), ("time", "<f4"), ("flags", "<u4"), ("pad", "<f4", 2)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("b1", "<f4"), ("b2", "<f4"), ("meshIndex", "<u4"), ("triangleIndex", "<u4")])
NODE_DTYPE = np.dtype([("data", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])


def can_build():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "luxrays", "accelerators"))


def build(force=False):
    """Compile the reference sources (only possible where /root/reference exists)."""
    if not can_build():
        return os.path.exists(_LIB_PATH)
    cmd = ["make", "-C", os.path.join(_HERE, "ref"), "-j8", "REF=" + REFERENCE_ROOT]
    if force:
        subprocess.check_call(cmd + ["clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return True


def available():
    if os.path.exists(_LIB_PATH):
        return True
    try:
        return build() and os.path.exists(_LIB_PATH)
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        if can_build():
            build()
        L = C.CDLL(_LIB_PATH)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        sig = {
            "ref_last_error": (C.c_char_p, []), "ref_describe": (C.c_char_p, []),
            "ref_set_epsilon": (None, [f32, f32]), "ref_epsilon": (f32, [f32]),
            "ref_matrix_inverse": (i32, [vp, vp]),
            "ref_triangle_intersect": (i32, [vp, vp, vp, vp, vp]), "ref_bbox_intersectp": (i32, [vp, vp, vp]),
            "ref_scene_create": (vp, []), "ref_scene_free": (None, [vp]), "ref_scene_mesh_count": (i32, [vp]),
            "ref_scene_add_shape": (i32, [vp, vp, u32, vp, u32]), "ref_scene_add_plain": (i32, [vp, i32]),
            "ref_scene_add_instance": (i32, [vp, i32, vp]), "ref_scene_add_motion": (i32, [vp, i32, u32, vp, vp]),
            "ref_scene_set_instance_transform": (i32, [vp, i32, vp]), "ref_scene_mesh_bbox": (i32, [vp, i32, vp]),
            "ref_bvh_build": (vp, [vp, i32, i32, i32, i32, f32]), "ref_bvh_set_nodes": (i32, [vp, vp, u32]),
            "ref_accel_free": (None, [vp]),
            "ref_bvh_node_count": (u32, [vp]), "ref_bvh_nodes": (vp, [vp]),
            "ref_accel_intersect": (i32, [vp, vp, vp, u64, i32]),
            "ref_mbvh_build": (vp, [vp, i32, i32, i32, i32, f32]), "ref_mbvh_update": (i32, [vp]),
            "ref_mbvh_root_node_count": (u32, [vp]), "ref_mbvh_root_nodes": (vp, [vp]),
            "ref_mbvh_leaf_count": (u32, [vp]), "ref_mbvh_leaf_node_count": (u32, [vp, u32]), "ref_mbvh_leaf_nodes": (vp, [vp, u32]),
            "ref_mbvh_transform_count": (u32, [vp]), "ref_mbvh_transform_minv": (None, [vp, u32, vp]),
            "ref_mbvh_motion_count": (u32, [vp]), "ref_motion_sample": (i32, [vp, u32, f32, vp]),
            "ref_hardware_threads": (i32, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _err():
    return RuntimeError(lib().ref_last_error().decode())


def _nodes(ptr, n):
    if not n:
        return np.zeros(0, dtype=NODE_DTYPE)
    buf = (C.c_char * (int(n) * 32)).from_address(ptr)
    return np.frombuffer(buf, dtype=NODE_DTYPE).copy()


def epsilon(v):
    return float(lib().ref_epsilon(C.c_float(v)))


def matrix_inverse(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
    out = np.zeros((4, 4), dtype=np.float32)
    if lib().ref_matrix_inverse(m.ctypes.data, out.ctypes.data) != 0:
        raise _err()
    return out


def triangle_intersect(ray, p0, p1, p2):
    ray = np.ascontiguousarray(ray)
    ps = [np.ascontiguousarray(p, dtype=np.float32) for p in (p0, p1, p2)]
    tb = np.zeros(3, dtype=np.float32)
    hit = lib().ref_triangle_intersect(ray.ctypes.data, ps[0].ctypes.data, ps[1].ctypes.data, ps[2].ctypes.data, tb.ctypes.data)
    return bool(hit), tb


def bbox_intersectp(ray, bmin, bmax):
    ray = np.ascontiguousarray(ray)
    a = np.ascontiguousarray(bmin, dtype=np.float32)
    b = np.ascontiguousarray(bmax, dtype=np.float32)
    return bool(lib().ref_bbox_intersectp(ray.ctypes.data, a.ctypes.data, b.ctypes.data))


class Scene:
    """luxrays::TriangleMesh / InstanceTriangleMesh / MotionTriangleMesh objects in dataset order."""

    def __init__(self):
        self.h = lib().ref_scene_create()

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ref_scene_free(self.h)
        except Exception:       # interpreter shutdown
            pass
        self.h = None

    def add_shape(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
        r = lib().ref_scene_add_shape(self.h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0])
        if r < 0:
            raise _err()
        return r

    def add_plain(self, shape):
        return lib().ref_scene_add_plain(self.h, shape)

    def add_instance(self, shape, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        r = lib().ref_scene_add_instance(self.h, shape, m.ctypes.data)
        if r < 0:
            raise _err()
        return r

    def add_motion(self, shape, times, mats):
        t = np.ascontiguousarray(times, dtype=np.float32)
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 4, 4)
        r = lib().ref_scene_add_motion(self.h, shape, t.shape[0], t.ctypes.data, m.ctypes.data)
        if r < 0:
            raise _err()
        return r

    def set_instance_transform(self, mesh, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        if lib().ref_scene_set_instance_transform(self.h, mesh, m.ctypes.data) != 0:
            raise _err()

    def mesh_count(self):
        return lib().ref_scene_mesh_count(self.h)

    def mesh_bbox(self, i):
        out = np.zeros(6, dtype=np.float32)
        lib().ref_scene_mesh_bbox(self.h, i, out.ctypes.data)
        return out


class _Accel:
    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().ref_accel_free(self.h)
        except Exception:       # interpreter shutdown
            pass
        self.h = None

    def intersect(self, rays, nthreads=None):
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 48
        hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        if lib().ref_accel_intersect(self.h, rays.ctypes.data, hits.ctypes.data, rays.shape[0], nthreads or 0) != 0:
            raise _err()
        return hits


class BVH(_Accel):
    """luxrays::BVHAccel built by the reference's CLASSIC builder (the Embree builders need Embree);
    nodes= replaces the tree with a caller-supplied BVHArrayNode array before intersecting."""

    def __init__(self, scene, tree_type=4, cost_samples=0, isect_cost=80, trav_cost=10, empty_bonus=0.5, nodes=None):
        self.scene = scene
        self.h = lib().ref_bvh_build(scene.h, tree_type, cost_samples, isect_cost, trav_cost, C.c_float(empty_bonus))
        if not self.h:
            raise _err()
        if nodes is not None:
            nodes = np.ascontiguousarray(nodes)
            assert nodes.dtype.itemsize == 32
            if lib().ref_bvh_set_nodes(self.h, nodes.ctypes.data, nodes.shape[0]) != 0:
                raise _err()

    def nodes(self):
        return _nodes(lib().ref_bvh_nodes(self.h), lib().ref_bvh_node_count(self.h))


class MBVH(_Accel):
    def __init__(self, scene, tree_type=4, cost_samples=0, isect_cost=80, trav_cost=10, empty_bonus=0.5):
        self.scene = scene
        self.h = lib().ref_mbvh_build(scene.h, tree_type, cost_samples, isect_cost, trav_cost, C.c_float(empty_bonus))
        if not self.h:
            raise _err()

    def update(self):
        if lib().ref_mbvh_update(self.h) != 0:
            raise _err()

    def root_nodes(self):
        return _nodes(lib().ref_mbvh_root_nodes(self.h), lib().ref_mbvh_root_node_count(self.h))

    def leaf_count(self):
        return lib().ref_mbvh_leaf_count(self.h)

    def leaf_nodes(self, i):
        return _nodes(lib().ref_mbvh_leaf_nodes(self.h, i), lib().ref_mbvh_leaf_node_count(self.h, i))

    def transforms_minv(self):
        n = lib().ref_mbvh_transform_count(self.h)
        out = np.zeros((n, 16), dtype=np.float32)
        for i in range(n):
            lib().ref_mbvh_transform_minv(self.h, i, out[i].ctypes.data)
        return out

    def motion_count(self):
        return lib().ref_mbvh_motion_count(self.h)

    def motion_sample(self, i, time):
        out = np.zeros((4, 4), dtype=np.float32)
        if lib().ref_motion_sample(self.h, i, C.c_float(time), out.ctypes.data) != 0:
            raise _err()
        return out


def hardware_threads():
    return lib().ref_hardware_threads()
