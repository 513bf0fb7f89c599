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
