"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
and fails loudly (no fallback) without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from luxcore_b200 import capi, hostapi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "luxrays_b200.h")).read()
    declared = set(re.findall(r"LRB_API\s+[\w\s\*]+?\b(lrb_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS)
    L = capi.lib()
    for name in declared:
        assert getattr(L, name) is not None


def test_host_layer_symbols_exported():
    L = hostapi.lib()
    for name in hostapi.EXPORTS:
        assert getattr(L, name) is not None


def test_wire_struct_sizes():
    assert capi.RAY_DTYPE.itemsize == 48 and capi.HIT_DTYPE.itemsize == 20 and capi.NODE_DTYPE.itemsize == 32
    assert C.sizeof(capi.SceneInfo) == 32 and C.sizeof(capi.Counters) == 48 and C.sizeof(capi.TraceStats) == 48


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.device_count() == 0
    with pytest.raises(capi.LrbError) as e:
        capi.Device(0)
    assert e.value.code == capi.LRB_ERR_NO_DEVICE
    # the host layer refuses to start as well -- no CPU fallback anywhere
    s = hostapi.Session({}, None)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    s.add_plain(s.add_shape(v, np.array([[0, 1, 2]], np.uint32)))
    with pytest.raises(hostapi.HostError):
        s.start(0)
    s.close()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under luxcore_b200/ or include/ may reference it."""
    bad = []
    for base in ("luxcore_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(d, f), errors="replace").read()
                    if re.search(r"(^|\s)(from|import)\s+oracle\b|lux_oracle|liblux_oracle|orc_\w+\(", txt):
                        bad.append(os.path.join(d, f))
    assert not bad, bad


def test_scene_build_argument_checks_need_no_device():
    """lrb_bvh_build_scene / lrb_scene_adopt / lrb_scene_download refuse malformed input with LRB_ERR_INVALID before any CUDA
    call (the same checks run on the GPU box in tests/test_gpu_zz_scene_build.py; here: no device at all)."""
    L = capi.lib()
    verts = np.zeros((3, 3), np.float32)
    tri = np.array([[0, 1, 2]], np.uint32)
    voff = np.array([0], np.uint32)
    scene = C.c_void_p(0xdead)
    n_nodes = C.c_uint32(7)

    def call(dev=None, v=verts, vo=voff, to=np.array([0, 1], np.uint32), t=tri, n_meshes=1, tree_type=4, quality=1, out=scene):
        return L.lrb_bvh_build_scene(dev, capi._ptr(v), v.shape[0] if v is not None else 0, capi._ptr(vo), capi._ptr(to), n_meshes, capi._ptr(t), tree_type, quality,
                                     C.byref(out) if out is not None else None, None, 0, C.byref(n_nodes), None)
    for kw, msg in ((dict(out=None), "null out"), (dict(tree_type=3), "tree type"), (dict(quality=2), "quality"), (dict(v=None), "needs vertices"),
                    (dict(n_meshes=0), "needs vertices"), (dict(to=np.array([1, 2], np.uint32)), "start at 0"),
                    (dict(to=np.array([0, 3, 2], np.uint32), vo=np.array([0, 0], np.uint32), n_meshes=2), "must not decrease"),
                    (dict(vo=np.array([9], np.uint32)), "outside the vertex buffer"), (dict(t=None), "null triangle"), (dict(), "null device")):
        assert call(**kw) == capi.LRB_ERR_INVALID, kw
        assert msg in L.lrb_last_error_string().decode(), (kw, L.lrb_last_error_string())
        if kw.get("out", scene) is not None:
            assert scene.value is None and n_nodes.value == 0       # outputs cleared on every failure
            scene.value, n_nodes.value = 0xdead, 7
    assert L.lrb_scene_adopt(None, None) == capi.LRB_ERR_INVALID
    assert L.lrb_scene_download(None, None, None, None) == capi.LRB_ERR_INVALID
