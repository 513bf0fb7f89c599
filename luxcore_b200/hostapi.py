"""ctypes binding of the C++ host layer (libluxrays_b200_host.so, luxcore_b200/host/shim.cpp).

A Session drives the luxrays:: classes in the order a LuxCore application does:
Context -> DataSet.Add/Preprocess -> SetDataSet/Start -> AllocBufferRW -> EnqueueTraceRayBuffer ->
EnqueueReadBuffer -> FinishQueue.  Accelerators are built on the host by the product's own builders
(luxcore_b200/host/bvhbuild.cpp); tracing always happens on the GPU.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from . import scenes as S

_HERE = os.path.dirname(os.path.abspath(__file__))
# LRB_LIB_DIR: development switch -- load a differently compiled build of the two libraries (A/B of kernel variants)
LIB_PATH = os.path.join(os.environ.get("LRB_LIB_DIR") or os.path.join(_HERE, "lib"), "libluxrays_b200_host.so")

EXPORTS = [
    "lrh_last_error", "lrh_create", "lrh_destroy", "lrh_device_description_count", "lrh_add_shape", "lrh_add_plain",
    "lrh_add_instance", "lrh_add_motion", "lrh_preprocess", "lrh_build_accelerator", "lrh_bvh_node_count",
    "lrh_bvh_nodes", "lrh_mbvh_root_node_count", "lrh_mbvh_root_nodes", "lrh_mbvh_leaf_count",
    "lrh_mbvh_leaf_node_count", "lrh_mbvh_leaf_nodes", "lrh_mesh_bbox", "lrh_dataset_bounds", "lrh_start", "lrh_stop", "lrh_native_device", "lrh_native_scene",
    "lrh_accelerator_type", "lrh_trace_host", "lrh_trace_device", "lrh_trace_device_shadow", "lrh_finish", "lrh_trace_ray",
    "lrh_set_instance_transform", "lrh_update", "lrh_stats_total_rays", "lrh_used_memory", "lrh_machine_epsilon",
    "lrh_matrix_inverse",
]

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libluxrays_b200_host.so is missing (%s); run __graft_entry__.build()" % LIB_PATH)
        capi.lib()      # dependency (RPATH=$ORIGIN also resolves it)
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32, f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_float
        sig = {
            "lrh_last_error": (C.c_char_p, []),
            "lrh_create": (vp, [C.c_char_p]),
            "lrh_destroy": (None, [vp]),
            "lrh_device_description_count": (i32, [vp]),
            "lrh_add_shape": (i32, [vp, vp, u32, vp, u32]),
            "lrh_add_plain": (i32, [vp, i32]),
            "lrh_add_instance": (i32, [vp, i32, vp]),
            "lrh_add_motion": (i32, [vp, i32, u32, vp, vp]),
            "lrh_preprocess": (i32, [vp]),
            "lrh_build_accelerator": (i32, [vp, C.c_char_p]),
            "lrh_bvh_node_count": (u32, [vp]),
            "lrh_bvh_nodes": (vp, [vp]),
            "lrh_mbvh_root_node_count": (u32, [vp]),
            "lrh_mbvh_root_nodes": (vp, [vp]),
            "lrh_mbvh_leaf_count": (u32, [vp]),
            "lrh_mbvh_leaf_node_count": (u32, [vp, u32]),
            "lrh_mbvh_leaf_nodes": (vp, [vp, u32]),
            "lrh_mesh_bbox": (i32, [vp, i32, vp]),
            "lrh_dataset_bounds": (i32, [vp, vp]),
            "lrh_start": (i32, [vp, i32]),
            "lrh_stop": (i32, [vp]),
            "lrh_native_device": (vp, [vp]),
            "lrh_native_scene": (vp, [vp]),
            "lrh_accelerator_type": (i32, [vp]),
            "lrh_trace_host": (i32, [vp, vp, vp, u32, i32]),
            "lrh_trace_device": (i32, [vp, vp, vp, u32]),
            "lrh_trace_device_shadow": (i32, [vp, vp, vp, u32]),
            "lrh_finish": (i32, [vp]),
            "lrh_trace_ray": (i32, [vp, vp, vp]),
            "lrh_set_instance_transform": (i32, [vp, i32, vp]),
            "lrh_update": (i32, [vp]),
            "lrh_stats_total_rays": (C.c_double, [vp]),
            "lrh_used_memory": (u64, [vp]),
            "lrh_machine_epsilon": (f32, [f32]),
            "lrh_matrix_inverse": (i32, [vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _nodes(ptr, n):
    if not n:
        return np.zeros(0, dtype=capi.NODE_DTYPE)
    buf = (C.c_char * (32 * n)).from_address(ptr)
    return np.frombuffer(buf, dtype=capi.NODE_DTYPE).copy()


class HostError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise HostError(lib().lrh_last_error().decode())


ACCEL_BVH, ACCEL_MBVH = 1, 2


class Session:
    """Context + DataSet + (after start) one CUDAIntersectionDevice."""

    def __init__(self, config=None, desc=None):
        """config: dict of luxrays properties, e.g. {"accelerator.bvh.builder.type": "CLASSIC"}."""
        text = "\n".join("%s = %s" % (k, v) for k, v in (config or {}).items())
        self.h = lib().lrh_create(text.encode())
        if not self.h:
            raise HostError(lib().lrh_last_error().decode())
        self.started = False
        if desc is not None:
            self.add_scene(desc)

    def close(self):
        if getattr(self, "h", None):
            lib().lrh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- DataSet ----
    def add_shape(self, verts, tris):
        v = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        t = np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)
        i = lib().lrh_add_shape(self.h, _ptr(v), v.shape[0], _ptr(t), t.shape[0])
        if i < 0:
            raise HostError(lib().lrh_last_error().decode())
        return i

    def add_plain(self, shape):
        _check(lib().lrh_add_plain(self.h, shape))

    def add_instance(self, shape, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        _check(lib().lrh_add_instance(self.h, shape, _ptr(m)))

    def add_motion(self, shape, times, mats):
        t = np.ascontiguousarray(times, dtype=np.float32)
        m = np.ascontiguousarray(mats, dtype=np.float32).reshape(-1, 4, 4)
        _check(lib().lrh_add_motion(self.h, shape, t.shape[0], _ptr(t), _ptr(m)))

    def add_scene(self, desc):
        for v, t in desc.shapes:
            self.add_shape(v, t)
        for m in desc.meshes:
            if m.kind == S.PLAIN:
                self.add_plain(m.shape)
            elif m.kind == S.INSTANCE:
                self.add_instance(m.shape, m.xform)
            else:
                self.add_motion(m.shape, m.times, m.motion_xforms)

    def preprocess(self):
        _check(lib().lrh_preprocess(self.h))

    def build_accelerator(self, kind="AUTO"):
        r = lib().lrh_build_accelerator(self.h, kind.encode())
        if r < 0:
            raise HostError(lib().lrh_last_error().decode())
        return r

    def mesh_bbox(self, i):
        out = np.zeros(6, dtype=np.float32)
        _check(lib().lrh_mesh_bbox(self.h, i, _ptr(out)))
        return out

    def dataset_bounds(self):
        """DataSet::GetBBox / GetBSphere -> (min xyz, max xyz, centre xyz, radius)."""
        out = np.zeros(10, dtype=np.float32)
        _check(lib().lrh_dataset_bounds(self.h, _ptr(out)))
        return out[:3].copy(), out[3:6].copy(), out[6:9].copy(), float(out[9])

    # ---- builder output (for parity tests) ----
    def bvh_nodes(self):
        return _nodes(lib().lrh_bvh_nodes(self.h), lib().lrh_bvh_node_count(self.h))

    def mbvh_root_nodes(self):
        return _nodes(lib().lrh_mbvh_root_nodes(self.h), lib().lrh_mbvh_root_node_count(self.h))

    def mbvh_leaf_count(self):
        return lib().lrh_mbvh_leaf_count(self.h)

    def mbvh_leaf_nodes(self, i):
        return _nodes(lib().lrh_mbvh_leaf_nodes(self.h, i), lib().lrh_mbvh_leaf_node_count(self.h, i))

    # ---- device ----
    def start(self, device_index=0):
        _check(lib().lrh_start(self.h, device_index))
        self.started = True

    def stop(self):
        if self.started:
            _check(lib().lrh_stop(self.h))
            self.started = False

    def accelerator_type(self):
        return lib().lrh_accelerator_type(self.h)

    def native_device(self):
        return lib().lrh_native_device(self.h)

    def native_scene(self):
        """capi.Scene view of the running kernel's C-ABI scene (not owned)."""
        h = lib().lrh_native_scene(self.h)
        if not h:
            raise HostError("session not started")
        sc = capi.Scene(None, C.c_void_p(h))
        return sc

    def set_stream(self, cuda_stream_handle):
        capi._check(capi.lib().lrb_device_set_stream(C.c_void_p(self.native_device()), C.c_void_p(cuda_stream_handle or 0)))

    def set_option(self, key, value):
        capi._check(capi.lib().lrb_device_set_option(C.c_void_p(self.native_device()), key.encode(), str(value).encode()))

    def counters(self):
        c = capi.Counters()
        capi._check(capi.lib().lrb_get_counters(C.c_void_p(self.native_device()), C.byref(c)))
        return c

    def trace_host(self, rays, hits=None):
        """AllocBufferRW + EnqueueTraceRayBuffer + EnqueueReadBuffer + FinishQueue on host arrays."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 48
        preload = hits is not None
        if hits is None:
            hits = np.zeros(rays.shape[0], dtype=capi.HIT_DTYPE)
        _check(lib().lrh_trace_host(self.h, _ptr(rays), _ptr(hits), rays.shape[0], 1 if preload else 0))
        return hits

    def trace_host_ptr(self, rays_ptr, hits_ptr, n):
        _check(lib().lrh_trace_host(self.h, C.c_void_p(rays_ptr), C.c_void_p(hits_ptr), n, 0))

    def trace_device(self, rays_devptr, hits_devptr, n):
        """EnqueueTraceRayBuffer on caller-owned device memory (asynchronous)."""
        _check(lib().lrh_trace_device(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n))

    def trace_device_shadow(self, rays_devptr, hits_devptr, n):
        """EnqueueTraceShadowRayBuffer (any-hit) on caller-owned device memory (asynchronous)."""
        _check(lib().lrh_trace_device_shadow(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n))

    def finish(self):
        _check(lib().lrh_finish(self.h))

    def trace_ray(self, ray):
        ray = np.ascontiguousarray(ray).reshape(1)
        hit = np.zeros(1, dtype=capi.HIT_DTYPE)
        r = lib().lrh_trace_ray(self.h, _ptr(ray), _ptr(hit))
        if r < 0:
            raise HostError(lib().lrh_last_error().decode())
        return bool(r), hit[0]

    def set_instance_transform(self, mesh, m):
        m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
        _check(lib().lrh_set_instance_transform(self.h, mesh, _ptr(m)))

    def update(self):
        _check(lib().lrh_update(self.h))

    def total_rays(self):
        return lib().lrh_stats_total_rays(self.h)

    def used_memory(self):
        return lib().lrh_used_memory(self.h)


def machine_epsilon(v):
    return float(lib().lrh_machine_epsilon(C.c_float(v)))


def matrix_inverse(m):
    m = np.ascontiguousarray(m, dtype=np.float32).reshape(4, 4)
    out = np.zeros((4, 4), dtype=np.float32)
    _check(lib().lrh_matrix_inverse(_ptr(m), _ptr(out)))
    return out
