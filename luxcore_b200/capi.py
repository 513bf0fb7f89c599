"""ctypes binding of the C ABI declared in include/luxrays_b200.h (libluxrays_b200.so).

This is the same boundary a LuxCore maintainer would bind from C++; Python uses it
for the parity tests and the benchmark.  There is no fallback of any kind: a missing library or a
missing CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# LRB_LIB_DIR: development switch -- load a differently compiled build of the two libraries (A/B of kernel variants)
LIB_PATH = os.path.join(os.environ.get("LRB_LIB_DIR") or os.path.join(_HERE, "lib"), "libluxrays_b200.so")

LRB_OK = 0
LRB_ERR_INVALID, LRB_ERR_CUDA, LRB_ERR_NO_DEVICE, LRB_ERR_OOM, LRB_ERR_INTERNAL = 1, 2, 3, 4, 5
NULL_INDEX = 0xFFFFFFFF
RAY_FLAGS_MASKED = 1

RAY_DTYPE = np.dtype([("o", "<f4", 3), ("d", "<f4", 3), ("mint", "<f4"), ("maxt", "<f4"), ("time", "<f4"),
                      ("flags", "<u4"), ("pad", "<f4", 2)])
HIT_DTYPE = np.dtype([("t", "<f4"), ("b1", "<f4"), ("b2", "<f4"), ("meshIndex", "<u4"), ("triangleIndex", "<u4")])
NODE_DTYPE = np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])

# every symbol include/luxrays_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "lrb_device_count", "lrb_device_create", "lrb_device_destroy", "lrb_device_get_props",
    "lrb_device_set_stream", "lrb_device_get_stream", "lrb_device_set_option",
    "lrb_alloc", "lrb_free", "lrb_h2d", "lrb_d2h", "lrb_flush", "lrb_sync",
    "lrb_bvh_upload", "lrb_mbvh_upload", "lrb_mbvh_update", "lrb_scene_free", "lrb_scene_get_info",
    "lrb_trace", "lrb_trace_host", "lrb_trace_stats",
    "lrb_trace_anyhit", "lrb_compact_rays", "lrb_trace_indexed", "lrb_advance_rays", "lrb_trace_passthrough",
    "lrb_last_error_string", "lrb_get_counters", "lrb_reset_counters", "lrb_version_string",
    "lrb_measure_read_bandwidth",
    "lrb_ipc_get_handle", "lrb_ipc_open_handle", "lrb_ipc_close_handle", "lrb_trace_gather", "lrb_gather_wait", "lrb_film_reduce",
    "lrb_build_lbvh", "lrb_build_bvh", "lrb_gather_signal", "lrb_wait_value",
    "lrb_bvh_build_scene", "lrb_scene_adopt", "lrb_scene_download",
]


class DeviceProps(C.Structure):
    _fields_ = [("cuda_ordinal", C.c_int), ("cc_major", C.c_int), ("cc_minor", C.c_int), ("sm_count", C.c_int),
                ("l2_bytes", C.c_int), ("total_mem_bytes", C.c_uint64), ("name", C.c_char * 128)]


class MBVHDesc(C.Structure):
    _fields_ = [("root_nodes", C.c_void_p), ("n_root_nodes", C.c_uint32), ("n_leaves", C.c_uint32),
                ("leaf_nodes", C.POINTER(C.c_void_p)), ("leaf_n_nodes", C.POINTER(C.c_uint32)),
                ("leaf_vertices", C.POINTER(C.c_void_p)), ("leaf_n_vertices", C.POINTER(C.c_uint32)),
                ("transforms_minv", C.c_void_p), ("n_transforms", C.c_uint32),
                ("motion_systems", C.c_void_p), ("n_motion_systems", C.c_uint32),
                ("interpolated_transforms", C.c_void_p), ("n_interpolated_transforms", C.c_uint32)]


class SceneInfo(C.Structure):
    _fields_ = [("n_ref_nodes", C.c_uint32), ("n_wide_nodes", C.c_uint32), ("n_triangles", C.c_uint32),
                ("n_instances", C.c_uint32), ("stack_need", C.c_uint32), ("two_level", C.c_uint32),
                ("device_bytes", C.c_uint64)]


class Counters(C.Structure):
    _fields_ = [("rays_traced", C.c_uint64), ("trace_launches", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("device_bytes_in_use", C.c_uint64)]


class TraceStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("wide_nodes", C.c_uint64), ("triangles", C.c_uint64),
                ("instances", C.c_uint64), ("motion_samples", C.c_uint64), ("max_stack", C.c_uint64)]


class BuildTimings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("sort_ms", C.c_double), ("tree_ms", C.c_double), ("emit_ms", C.c_double),
                ("d2h_ms", C.c_double), ("kernels", C.c_uint32)]


class SceneBuildTimings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("leafbox_ms", C.c_double), ("sort_ms", C.c_double), ("tree_ms", C.c_double),
                ("emit_ms", C.c_double), ("relayout_ms", C.c_double), ("d2h_ms", C.c_double), ("kernels", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class LrbError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "luxrays_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the CUDA library.  Raises if it has not been built -- there is no other path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libluxrays_b200.so is missing (%s); run __graft_entry__.build() -- the B200 device "
                              "has no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
        pvp = C.POINTER(C.c_void_p)
        sig = {
            "lrb_device_count": (i32, [C.POINTER(C.c_int)]),
            "lrb_device_create": (i32, [i32, pvp]),
            "lrb_device_destroy": (i32, [vp]),
            "lrb_device_get_props": (i32, [vp, C.POINTER(DeviceProps)]),
            "lrb_device_set_stream": (i32, [vp, vp]),
            "lrb_device_get_stream": (i32, [vp, pvp]),
            "lrb_device_set_option": (i32, [vp, C.c_char_p, C.c_char_p]),
            "lrb_alloc": (i32, [vp, sz, pvp]),
            "lrb_free": (i32, [vp, vp]),
            "lrb_h2d": (i32, [vp, vp, vp, sz, i32]),
            "lrb_d2h": (i32, [vp, vp, vp, sz, i32]),
            "lrb_flush": (i32, [vp]),
            "lrb_sync": (i32, [vp]),
            "lrb_bvh_upload": (i32, [vp, vp, u32, vp, u64, vp, u32, pvp]),
            "lrb_mbvh_upload": (i32, [vp, C.POINTER(MBVHDesc), pvp]),
            "lrb_mbvh_update": (i32, [vp, vp, u32, vp, u32]),
            "lrb_scene_free": (i32, [vp]),
            "lrb_scene_get_info": (i32, [vp, C.POINTER(SceneInfo)]),
            "lrb_trace": (i32, [vp, vp, vp, u32]),
            "lrb_trace_host": (i32, [vp, vp, vp, u32, i32]),
            "lrb_trace_anyhit": (i32, [vp, vp, vp, u32]),
            "lrb_compact_rays": (i32, [vp, vp, u32, pvp, pvp, C.POINTER(u32)]),
            "lrb_trace_indexed": (i32, [vp, vp, vp, u32, vp, vp, i32]),
            "lrb_advance_rays": (i32, [vp, vp, vp, u32, vp, u32, vp, C.POINTER(u32)]),
            "lrb_trace_passthrough": (i32, [vp, vp, vp, u32, vp, u32, u32, C.POINTER(u32), C.POINTER(u64)]),
            "lrb_trace_stats": (i32, [vp, vp, vp, u32, C.POINTER(TraceStats)]),
            "lrb_last_error_string": (C.c_char_p, []),
            "lrb_get_counters": (i32, [vp, C.POINTER(Counters)]),
            "lrb_reset_counters": (i32, [vp]),
            "lrb_version_string": (C.c_char_p, []),
            "lrb_measure_read_bandwidth": (i32, [vp, sz, i32, C.POINTER(C.c_double)]),
            "lrb_ipc_get_handle": (i32, [vp, vp, C.c_char_p]),
            "lrb_ipc_open_handle": (i32, [vp, C.c_char_p, pvp]),
            "lrb_ipc_close_handle": (i32, [vp, vp]),
            "lrb_trace_gather": (i32, [vp, vp, vp, u32, vp, u32]),
            "lrb_gather_wait": (i32, [vp, vp, i32]),
            "lrb_film_reduce": (i32, [vp, C.POINTER(vp), u32, vp, u64, u64]),
            "lrb_build_lbvh": (i32, [vp, vp, u32, u32, vp, u32, C.POINTER(u32), C.POINTER(BuildTimings)]),
            "lrb_gather_signal": (i32, [vp, vp, u32]),
            "lrb_wait_value": (i32, [vp, vp, u32, vp]),
            "lrb_build_bvh": (i32, [vp, vp, u32, u32, u32, vp, u32, C.POINTER(u32), C.POINTER(BuildTimings)]),
            "lrb_bvh_build_scene": (i32, [vp, vp, u64, vp, vp, u32, vp, u32, u32, pvp, vp, u32, C.POINTER(u32), C.POINTER(SceneBuildTimings)]),
            "lrb_scene_adopt": (i32, [vp, vp]),
            "lrb_scene_download": (i32, [vp, vp, vp, vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _check(rc):
    if rc != LRB_OK:
        raise LrbError(rc, lib().lrb_last_error_string().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def device_count():
    n = C.c_int(0)
    rc = lib().lrb_device_count(C.byref(n))
    if rc == LRB_ERR_NO_DEVICE:
        return 0
    _check(rc)
    return n.value


class Device:
    """One B200 behind the C ABI (the object CUDAIntersectionDevice wraps on the C++ side)."""

    def __init__(self, ordinal=0, _borrow=None):
        if _borrow is not None:
            self.h = C.c_void_p(_borrow)
            self.owned = False
        else:
            h = C.c_void_p()
            _check(lib().lrb_device_create(ordinal, C.byref(h)))
            self.h = h
            self.owned = True
        self.ordinal = ordinal

    @classmethod
    def borrow(cls, native_handle):
        """Wrap an lrb_device* owned by someone else (e.g. a hostapi.Session)."""
        return cls(_borrow=native_handle)

    def close(self):
        if getattr(self, "h", None) and self.owned:
            lib().lrb_device_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def props(self):
        p = DeviceProps()
        _check(lib().lrb_device_get_props(self.h, C.byref(p)))
        return p

    def set_stream(self, cuda_stream_handle):
        _check(lib().lrb_device_set_stream(self.h, C.c_void_p(cuda_stream_handle or 0)))

    def set_option(self, key, value):
        _check(lib().lrb_device_set_option(self.h, key.encode(), str(value).encode()))

    def alloc(self, nbytes):
        p = C.c_void_p()
        _check(lib().lrb_alloc(self.h, nbytes, C.byref(p)))
        return p.value or 0

    def free(self, devptr):
        _check(lib().lrb_free(self.h, C.c_void_p(devptr)))

    def h2d(self, devptr, arr, blocking=False):
        arr = np.ascontiguousarray(arr)
        _check(lib().lrb_h2d(self.h, C.c_void_p(devptr), _ptr(arr), arr.nbytes, 1 if blocking else 0))

    def d2h(self, arr, devptr, blocking=True):
        assert arr.flags["C_CONTIGUOUS"]
        _check(lib().lrb_d2h(self.h, _ptr(arr), C.c_void_p(devptr), arr.nbytes, 1 if blocking else 0))

    def sync(self):
        _check(lib().lrb_sync(self.h))

    def flush(self):
        _check(lib().lrb_flush(self.h))

    def counters(self):
        c = Counters()
        _check(lib().lrb_get_counters(self.h, C.byref(c)))
        return c

    def measure_read_bandwidth(self, nbytes, iters=20):
        """GB/s of a streaming 256-bit read kernel over an nbytes buffer (L2-resident if it fits)."""
        g = C.c_double(0)
        _check(lib().lrb_measure_read_bandwidth(self.h, nbytes, iters, C.byref(g)))
        return g.value

    def reset_counters(self):
        _check(lib().lrb_reset_counters(self.h))

    def compact_rays(self, rays_devptr, n, want_count=True):
        """Dense list of the non-masked rays -> (device pointer of the index list, device pointer of its length, length or None)."""
        idx, cnt, host = C.c_void_p(), C.c_void_p(), C.c_uint32(0)
        _check(lib().lrb_compact_rays(self.h, C.c_void_p(rays_devptr), n, C.byref(idx), C.byref(cnt), C.byref(host) if want_count else None))
        return idx.value or 0, cnt.value or 0, (int(host.value) if want_count else None)

    def gather_wait(self, cuda_stream_handle=0, which=-1):
        """Deferred gathers (option gather_defer): make a stream wait for the pushes of the last (0) / previous (1) / both (-1) calls."""
        _check(lib().lrb_gather_wait(self.h, C.c_void_p(cuda_stream_handle or 0), which))

    def gather_signal(self, flag_devptr, value):
        """Write `value` into the 32-bit flag word (local or peer-mapped) behind this device's pushes: copy engine only."""
        _check(lib().lrb_gather_signal(self.h, C.c_void_p(flag_devptr), value))

    def wait_value(self, flag_devptr, value, cuda_stream_handle=0):
        """Make a stream (0 = the device's queue) wait until the flag word (this device's memory) is >= value."""
        _check(lib().lrb_wait_value(self.h, C.c_void_p(flag_devptr), value, C.c_void_p(cuda_stream_handle or 0)))

    def film_reduce(self, tile_devptrs, dst_devptr, first, count):
        """dst[first:first+count] = sum of the tiles' float planes in list order (Film::AddFilm order), asynchronous."""
        arr = (C.c_void_p * len(tile_devptrs))(*[C.c_void_p(p) for p in tile_devptrs])
        _check(lib().lrb_film_reduce(self.h, arr, len(tile_devptrs), C.c_void_p(dst_devptr), first, count))

    # ---- multi-GPU gather buffer sharing ----
    def ipc_get_handle(self, devptr):
        buf = C.create_string_buffer(64)
        _check(lib().lrb_ipc_get_handle(self.h, C.c_void_p(devptr), buf))
        return buf.raw

    def ipc_open_handle(self, handle_bytes):
        p = C.c_void_p()
        _check(lib().lrb_ipc_open_handle(self.h, C.c_char_p(handle_bytes), C.byref(p)))
        return p.value

    def ipc_close_handle(self, devptr):
        _check(lib().lrb_ipc_close_handle(self.h, C.c_void_p(devptr)))

    # ---- BVH construction on the device ----
    def build_lbvh(self, leaf_boxes, tree_type=4, node_dtype=None, quality=0):
        """GPU BVH builder (lrb_build_bvh; quality 0 = radix tree = lrb_build_lbvh, 1 = PLOC): leaf_boxes [n, 6] float32
        (min xyz, max xyz) -> (BVHArrayNode array as a [n_nodes] array of 32-byte records, BuildTimings).  Leaf records
        carry the input index of their leaf in their first word; the caller writes the leaf payload in."""
        boxes = np.ascontiguousarray(leaf_boxes, dtype=np.float32).reshape(-1, 6)
        n = boxes.shape[0]
        dt = node_dtype or np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])
        assert dt.itemsize == 32
        out = np.zeros(max(1, 2 * n), dtype=dt)
        total = C.c_uint32()
        tm = BuildTimings()
        _check(lib().lrb_build_bvh(self.h, _ptr(boxes), n, tree_type, quality, _ptr(out), out.shape[0], C.byref(total), C.byref(tm)))
        return out[:total.value].copy(), tm

    def build_scene(self, verts, mesh_vertex_offsets, triangles, mesh_triangle_offsets, tree_type=4, quality=1, want_nodes=False, node_dtype=None):
        """lrb_bvh_build_scene: triangles in, traceable scene out, everything on the device (leaf boxes, tree, leaf payload,
        re-layout).  verts [n, 3] float32 + first vertex per mesh, triangles [t, 3] uint32 (mesh-local indices) + first
        triangle per mesh (n_meshes + 1 entries).  -> (Scene, SceneBuildTimings, reference array or None)."""
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        voff = np.ascontiguousarray(mesh_vertex_offsets, dtype=np.uint32)
        tri = np.ascontiguousarray(triangles, dtype=np.uint32).reshape(-1, 3)
        toff = np.ascontiguousarray(mesh_triangle_offsets, dtype=np.uint32)
        assert toff.shape[0] == voff.shape[0] + 1 and int(toff[-1]) == tri.shape[0]
        dt = node_dtype or np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")])
        out = np.zeros(max(1, 2 * tri.shape[0]), dtype=dt) if want_nodes else None
        total = C.c_uint32()
        tm = SceneBuildTimings()
        s = C.c_void_p()
        _check(lib().lrb_bvh_build_scene(self.h, _ptr(verts), verts.shape[0], _ptr(voff), _ptr(toff), voff.shape[0], _ptr(tri), tree_type, quality,
                                         C.byref(s), _ptr(out) if want_nodes else None, out.shape[0] if want_nodes else 0, C.byref(total), C.byref(tm)))
        return Scene(self, s), tm, (out[:total.value].copy() if want_nodes else None)

    # ---- scenes ----
    def upload_bvh(self, nodes, verts, mesh_vertex_offsets):
        nodes = np.ascontiguousarray(nodes)
        assert nodes.dtype.itemsize == 32
        verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        offs = np.ascontiguousarray(mesh_vertex_offsets, dtype=np.uint32)
        s = C.c_void_p()
        _check(lib().lrb_bvh_upload(self.h, _ptr(nodes), nodes.shape[0], _ptr(verts), verts.shape[0],
                                    _ptr(offs), offs.shape[0], C.byref(s)))
        return Scene(self, s)

    def upload_mbvh(self, root_nodes, leaf_nodes, leaf_verts, transforms_minv=None, motion_table=None, interps=None):
        """root_nodes: BVHArrayNode[]; leaf_nodes / leaf_verts: one array per unique leaf;
        transforms_minv: [n,4,4]; motion_table: uint32 [m,4] (ocl::MotionSystem); interps: uint8 [k*576]."""
        root_nodes = np.ascontiguousarray(root_nodes)
        leaf_nodes = [np.ascontiguousarray(a) for a in leaf_nodes]
        leaf_verts = [np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 3) for a in leaf_verts]
        n = len(leaf_nodes)
        d = MBVHDesc()
        d.root_nodes = root_nodes.ctypes.data
        d.n_root_nodes = root_nodes.shape[0]
        d.n_leaves = n
        ln = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in leaf_nodes])
        lc = (C.c_uint32 * max(n, 1))(*[a.shape[0] for a in leaf_nodes])
        lv = (C.c_void_p * max(n, 1))(*[a.ctypes.data for a in leaf_verts])
        lvc = (C.c_uint32 * max(n, 1))(*[a.shape[0] for a in leaf_verts])
        d.leaf_nodes = C.cast(ln, C.POINTER(C.c_void_p))
        d.leaf_n_nodes = C.cast(lc, C.POINTER(C.c_uint32))
        d.leaf_vertices = C.cast(lv, C.POINTER(C.c_void_p))
        d.leaf_n_vertices = C.cast(lvc, C.POINTER(C.c_uint32))
        keep = [root_nodes, leaf_nodes, leaf_verts, ln, lc, lv, lvc]
        if transforms_minv is not None and len(transforms_minv):
            tm = np.ascontiguousarray(transforms_minv, dtype=np.float32).reshape(-1, 16)
            d.transforms_minv = tm.ctypes.data
            d.n_transforms = tm.shape[0]
            keep.append(tm)
        if motion_table is not None and len(motion_table):
            mt = np.ascontiguousarray(motion_table, dtype=np.uint32).reshape(-1, 4)
            it = np.ascontiguousarray(interps, dtype=np.uint8)
            d.motion_systems = mt.ctypes.data
            d.n_motion_systems = mt.shape[0]
            d.interpolated_transforms = it.ctypes.data
            d.n_interpolated_transforms = it.shape[0] // 576
            keep += [mt, it]
        s = C.c_void_p()
        _check(lib().lrb_mbvh_upload(self.h, C.byref(d), C.byref(s)))
        del keep
        return Scene(self, s)


class Scene:
    """A re-laid-out BVH / MBVH resident in HBM (what BVHKernel / MBVHKernel own)."""

    def __init__(self, dev, handle):
        self.dev = dev
        self.h = handle

    def free(self):
        if getattr(self, "h", None) and self.dev is not None:
            lib().lrb_scene_free(self.h)
        self.h = None

    def __del__(self):
        try:
            # dev is None for borrowed handles (hostapi.Session.native_scene)
            if self.dev is not None and getattr(self.dev, "h", None):
                self.free()
        except Exception:
            pass

    def info(self):
        i = SceneInfo()
        _check(lib().lrb_scene_get_info(self.h, C.byref(i)))
        return i

    def adopt(self, dev):
        """lrb_scene_adopt: hand the scene to another Device handle of the same CUDA device."""
        _check(lib().lrb_scene_adopt(dev.h, self.h))
        self.dev = dev

    def download(self):
        """lrb_scene_download -> (wide nodes [n, 64] uint8, triangle records [t, 64] uint8, triangle ids [t, 2] uint32)."""
        i = self.info()
        wide = np.zeros((i.n_wide_nodes, 64), dtype=np.uint8)
        tris = np.zeros((i.n_triangles, 64), dtype=np.uint8)
        ids = np.zeros((i.n_triangles, 2), dtype=np.uint32)
        _check(lib().lrb_scene_download(self.h, _ptr(wide) if wide.size else None, _ptr(tris) if tris.size else None, _ptr(ids) if ids.size else None))
        return wide, tris, ids

    def update(self, root_nodes, transforms_minv):
        root_nodes = np.ascontiguousarray(root_nodes)
        tm = np.ascontiguousarray(transforms_minv, dtype=np.float32).reshape(-1, 16)
        _check(lib().lrb_mbvh_update(self.h, _ptr(root_nodes), root_nodes.shape[0], _ptr(tm) if tm.shape[0] else None, tm.shape[0]))

    def trace(self, rays_devptr, hits_devptr, n):
        """Asynchronous EnqueueTraceRayBuffer on device pointers."""
        _check(lib().lrb_trace(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n))

    def trace_anyhit(self, rays_devptr, hits_devptr, n):
        """Shadow rays: first hit found, hit / miss identical to trace()."""
        _check(lib().lrb_trace_anyhit(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n))

    def trace_indexed(self, rays_devptr, hits_devptr, n, live_idx_devptr, live_count_devptr=0, any_hit=False):
        _check(lib().lrb_trace_indexed(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n, C.c_void_p(live_idx_devptr),
                                       C.c_void_p(live_count_devptr or 0), 1 if any_hit else 0))

    def advance_rays(self, rays_devptr, hits_devptr, n, pass_mesh_bits_devptr=0, n_pass_words=0, continue_flags_devptr=0):
        """One round of the pass-through loop; returns the number of rays that continue."""
        c = C.c_uint32(0)
        _check(lib().lrb_advance_rays(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n, C.c_void_p(pass_mesh_bits_devptr or 0),
                                      n_pass_words, C.c_void_p(continue_flags_devptr or 0), C.byref(c)))
        return int(c.value)

    def trace_passthrough(self, rays_devptr, hits_devptr, n, pass_mesh_bits_devptr, n_pass_words, max_rounds=0):
        """-> (rounds traced, rays traced over all rounds)."""
        r, t = C.c_uint32(0), C.c_uint64(0)
        _check(lib().lrb_trace_passthrough(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n, C.c_void_p(pass_mesh_bits_devptr or 0),
                                           n_pass_words, max_rounds, C.byref(r), C.byref(t)))
        return int(r.value), int(t.value)

    def trace_host(self, rays, hits=None):
        """Host arrays in, host array out (H2D + trace + D2H inside)."""
        rays = np.ascontiguousarray(rays)
        assert rays.dtype.itemsize == 48
        preload = hits is not None
        if hits is None:
            hits = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        assert hits.dtype.itemsize == 20 and hits.shape[0] == rays.shape[0]
        _check(lib().lrb_trace_host(self.h, _ptr(rays), _ptr(hits), rays.shape[0], 1 if preload else 0))
        return hits

    def trace_host_ptr(self, rays_hostptr, hits_hostptr, n, preload_hits=False):
        _check(lib().lrb_trace_host(self.h, C.c_void_p(rays_hostptr), C.c_void_p(hits_hostptr), n, 1 if preload_hits else 0))

    def trace_gather(self, rays_devptr, hits_devptr, n, gather_dst_devptr, n_chunks=0):
        """Trace + overlapped push of the RayHit slice into the (possibly peer-mapped) gather buffer."""
        _check(lib().lrb_trace_gather(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr), n,
                                      C.c_void_p(gather_dst_devptr), n_chunks))

    def trace_stats(self, rays_devptr, hits_devptr, n):
        st = TraceStats()
        _check(lib().lrb_trace_stats(self.h, C.c_void_p(rays_devptr), C.c_void_p(hits_devptr or 0), n, C.byref(st)))
        return st
