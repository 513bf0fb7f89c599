"""The GPU tests that were written without a GPU at hand (tests/test_gpu_zz_*.py) are executed here, on the CPU,
against a stand-in device: `capi.Device` / `hostapi.Session` keep their real method signatures (checked with
inspect.signature.bind) and their real host-side parts (builders, node arrays), while uploads and traces go to
the CPU emulation of the traversal body.  Batch sizes are scaled down.  This does not test the CUDA path -- it
keeps the test code itself honest (API misuse, false properties, thresholds) before it first meets a GPU."""
import ctypes
import inspect
import os
import types

import numpy as np
import pytest
import torch

import helpers as H
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


class _FakeScene:
    def __init__(self, emu):
        self.emu = emu

    def info(self):
        i = self.emu.info()
        return types.SimpleNamespace(stack_need=i["stack_need"], two_level=i["two_level"], n_wide_nodes=i["wide"], n_triangles=i["tris"])

    def trace_host(self, rays):
        return self.emu.trace(rays)

    def free(self):
        pass


def _fake_device_class():
    sigs = {n: inspect.signature(getattr(capi.Device, n)) for n in ("upload_bvh", "upload_mbvh", "set_option", "close")}

    class FakeDevice:
        def __init__(self, ordinal):
            pass

        def upload_bvh(self, *a, **k):
            sigs["upload_bvh"].bind(self, *a, **k)
            return _FakeScene(H.Emu.bvh(*a))

        def upload_mbvh(self, *a, **k):
            b = sigs["upload_mbvh"].bind(self, *a, **k)
            b.apply_defaults()
            arr = {k2: b.arguments[k2] for k2 in ("root_nodes", "leaf_nodes", "leaf_verts", "transforms_minv", "motion_table", "interps")}
            return _FakeScene(H.Emu.mbvh(arr))

        def set_option(self, *a, **k):
            b = sigs["set_option"].bind(self, *a, **k)
            if b.arguments["key"] == "prefetch" and not 0 <= int(b.arguments["value"]) <= 2:
                raise RuntimeError("prefetch must be 0..2")

        def close(self):
            pass
    return FakeDevice


class _FakeSession(hostapi.Session):
    """Host-only parts are the real ones (builders, arrays); the device is the emulation."""

    def start(self, device_index=0):
        desc = self._desc
        osc = H.oracle_scene(desc)
        if self.accelerator_type() == hostapi.ACCEL_BVH:
            verts, offs = H.flattened_from_oracle(desc, osc)
            self._emu = H.Emu.bvh(self.bvh_nodes(), verts, offs)
        else:
            arr = H.mbvh_arrays(desc, O.MBVH(osc, tree_type=4))
            arr["root_nodes"] = self.mbvh_root_nodes().copy()
            for i in range(self.mbvh_leaf_count()):
                arr["leaf_nodes"][i] = self.mbvh_leaf_nodes(i).copy()
            self._emu = H.Emu.mbvh(arr)

    def trace_device(self, rays_ptr, hits_ptr, n):
        rays = np.frombuffer((ctypes.c_uint8 * (n * 48)).from_address(rays_ptr), dtype=H.RAY_DTYPE)
        hits = self._emu.trace(rays.copy())
        ctypes.memmove(hits_ptr, hits.ctypes.data, n * 20)

    def finish(self):
        pass

    def stop(self):
        pass


def _load(name, replacements, monkeypatch):
    src = open(os.path.join(HERE, name)).read()
    for a, b in replacements:
        assert a in src, a
        src = src.replace(a, b)
    mod = types.ModuleType(name[:-3])
    mod.__file__ = os.path.join(HERE, name)
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


@pytest.fixture
def small_batches(monkeypatch):
    ou, oc = R.uniform_rays, R.camera_rays
    monkeypatch.setattr(R, "uniform_rays", lambda lo, hi, n, **k: ou(lo, hi, min(n, 3000), **k))
    monkeypatch.setattr(R, "camera_rays", lambda cam, w, h, **k: oc(cam, min(w, 160), min(h, 160), **k))


def test_prepared_variant_tests_run(monkeypatch, small_batches):
    monkeypatch.setattr(capi, "Device", _fake_device_class())
    mod = _load("test_gpu_zz_prepared_variants.py", [], monkeypatch)
    mod.test_prefetch_variant_matches_oracle()
    mod.test_instance_vote_settings_match_oracle(0)


def test_full_size_property_tests_run(monkeypatch):
    def make_session(cfg=None, desc=None):
        s = _FakeSession(cfg, desc)
        s._desc = desc
        return s
    monkeypatch.setattr(hostapi, "Session", make_session)
    mod = _load("test_gpu_zz_full_size_properties.py", [
        ("assert rays_u8.is_cuda and rays_u8.is_contiguous()", "assert rays_u8.is_contiguous()"),
        ('torch.device("cuda", 0)', 'torch.device("cpu")'), ("torch.cuda.set_device(0)", "pass"), ("torch.cuda.synchronize()", "pass"),
        ("4 << 20", "20000"), ("200000", "5000"), ("100000", "5000"), ("50000", "3000"),
        ('S.load_fixture("lightinstances")', 'S.load_fixture("lightinstances", max_objects=200)')], monkeypatch)
    mod.test_interiors_16mi_bounce_rays("kitchen", 40000)
    mod.test_luxball_4mi_camera_and_bounce4()
    mod.test_lightinstances_4mi_two_level()


# ---- tests/test_gpu_zz_scene_build.py: lrb_bvh_build_scene ---------------------------------------------------------------

class _FakeBuiltScene(_FakeScene):
    """Scene of the fake build_scene: the arrays are those of the DEVICE re-layout's bodies run on the host (H.RelayoutDev)."""

    def __init__(self, emu, dev_arrays, n_ref, counters):
        _FakeScene.__init__(self, emu)
        self.arr = dev_arrays
        self.n_ref = n_ref
        self.counters = counters
        self.bytes = 64 * dev_arrays["wide"].shape[0] + 72 * dev_arrays["tris"].shape[0]
        counters.device_bytes_in_use += self.bytes

    def info(self):
        return types.SimpleNamespace(stack_need=self.arr["stack_need"], two_level=0, n_instances=0, n_ref_nodes=self.n_ref,
                                     n_wide_nodes=self.arr["wide"].shape[0], n_triangles=self.arr["tris"].shape[0], device_bytes=self.bytes)

    def download(self):
        return self.arr["wide"], self.arr["tris"], self.arr["ids"]

    def free(self):
        self.counters.device_bytes_in_use -= self.bytes


def _fake_build_device_class():
    sig_build = inspect.signature(capi.Device.build_scene)
    sig_upload = inspect.signature(capi.Device.upload_bvh)

    class FakeDevice:
        def __init__(self, ordinal):
            self._c = types.SimpleNamespace(device_bytes_in_use=0, d2h_bytes=0)

        def counters(self):
            return types.SimpleNamespace(**vars(self._c))

        def build_scene(self, *a, **k):
            b = sig_build.bind(self, *a, **k)
            b.apply_defaults()
            g = b.arguments
            verts = np.ascontiguousarray(g["verts"], dtype=np.float32).reshape(-1, 3)
            voff = np.ascontiguousarray(g["mesh_vertex_offsets"], dtype=np.uint32)
            tri = np.ascontiguousarray(g["triangles"], dtype=np.uint32).reshape(-1, 3)
            toff = np.ascontiguousarray(g["mesh_triangle_offsets"], dtype=np.uint32)
            if g["tree_type"] not in (2, 4, 8):
                raise capi.LrbError(1, "tree type must be 2, 4 or 8 (bvhaccel.cpp:51)")
            if tri.shape[0] and (tri + voff[np.searchsorted(toff, np.arange(tri.shape[0]), side="right") - 1][:, None] >= verts.shape[0]).any():
                raise capi.LrbError(1, "triangle leaf references a vertex outside the vertex buffer")
            # the tree: the host SAH builder stands in for the builder kernels (any tree exercises the re-layout)
            s = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": g["tree_type"]})
            for m in range(voff.shape[0]):
                v1 = voff[m + 1] if m + 1 < voff.shape[0] else verts.shape[0]
                s.add_plain(s.add_shape(verts[voff[m]:v1].reshape(-1, 3) if v1 > voff[m] else np.zeros((1, 3), np.float32), tri[toff[m]:toff[m + 1]]))
            s.build_accelerator("BVH")
            nodes = s.bvh_nodes().copy()
            s.close()
            arr = H.RelayoutDev.run(H.to_builder_format(nodes, toff), verts, voff, tri, toff) if tri.shape[0] > 1 else None
            emu = H.Emu.bvh(nodes, verts, voff)
            if arr is None:
                w, t, i = emu.arrays()
                arr = {"wide": w, "tris": t, "ids": i, "stack_need": emu.info()["stack_need"]}
            arr = dict(arr, wide=arr["wide"].view(np.uint8).reshape(-1, 64), tris=arr["tris"].view(np.uint8).reshape(-1, 64))
            tm = types.SimpleNamespace(kernels=20, relayout_ms=1.0, tree_ms=1.0, d2h_ms=1.0 if g["want_nodes"] else 0.0)
            return _FakeBuiltScene(emu, arr, nodes.shape[0], self._c), tm, (nodes.astype(g["node_dtype"]) if g["want_nodes"] else None)

        def upload_bvh(self, *a, **k):
            sig_upload.bind(self, *a, **k)
            emu = H.Emu.bvh(*a)
            w, t, i = emu.arrays()
            sc = _FakeScene(emu)
            nbytes = 64 * w.shape[0] + 72 * t.shape[0]
            sc.info = lambda: types.SimpleNamespace(device_bytes=nbytes)
            return sc

        def close(self):
            pass
    return FakeDevice


def test_scene_build_tests_run(monkeypatch, small_batches):
    monkeypatch.setattr(capi, "Device", _fake_build_device_class())
    mod = _load("test_gpu_zz_scene_build.py", [("n = 1000000", "n = 20000")], monkeypatch)
    d = capi.Device(0)
    mod.test_scene_built_on_the_device_is_the_host_layout_of_its_array(d, "cornell", 4, 1, 100000)
    mod.test_scene_built_on_the_device_is_the_host_layout_of_its_array(d, "kitchen", 8, 1, 100000)
    mod.test_scene_build_on_a_soup(d)
    mod.test_scene_build_small_and_ragged_inputs(d)
    mod.test_scene_build_rejects_bad_input(d)
