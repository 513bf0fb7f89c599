"""The chunk schedule of the host pipelines (luxcore_b200/csrc/host_chunks.h: lrb_trace_host and the pipelined plugin
sequence cut a batch that arrives from host memory into pieces whose tail shrinks geometrically, so that the part of
the pipeline nothing overlaps -- trace + read-back of the last piece -- is short)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ends():
    d = os.path.join(ROOT, "tests", "cpp")
    so = os.path.join(d, "libhost_chunks_shim.so")
    src = os.path.join(d, "host_chunks_shim.cpp")
    hdr = os.path.join(ROOT, "luxcore_b200", "csrc", "host_chunks.h")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.dirname(hdr), "-o", so, src])
    L = C.CDLL(so)
    L.hc_chunk_ends.restype = C.c_uint32
    L.hc_chunk_ends.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32]

    def f(n, chunk=1 << 20, min_chunk=1 << 16, taper=True):
        out = np.zeros(n // max(1, chunk) + 64, dtype=np.uint64)
        k = L.hc_chunk_ends(n, chunk, min_chunk, 1 if taper else 0, out.ctypes.data, out.shape[0])
        assert k <= out.shape[0]
        return [int(x) for x in out[:k]]
    return f


def sizes(e):
    return [b - a for a, b in zip([0] + e[:-1], e)]


def test_the_benched_batch(ends):
    e = ends(1 << 24)
    s = sizes(e)
    assert s == [1 << 20] * 15 + [1 << 19, 1 << 18, 1 << 17, 1 << 16, 1 << 16]
    assert sizes(ends(1 << 24, taper=False)) == [1 << 20] * 16


@pytest.mark.parametrize("taper", [False, True])
def test_pieces_tile_the_batch(ends, taper):
    rng = np.random.default_rng(7)
    cases = [0, 1, 2, 1023, 1024, 1025, 65535, 65536, 65537, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, (2 << 20) + 3, 33554432, 4294967295]
    cases += [int(x) for x in rng.integers(1, 1 << 26, 200)]
    for n in cases:
        for chunk, mn in ((1 << 20, 1 << 16), (1 << 18, 1 << 18), (4096, 1024), (1 << 20, 1 << 22), (rng.integers(1024, 1 << 21), rng.integers(1024, 1 << 18))):
            if n // int(chunk) > 100000:
                continue
            e = ends(n, int(chunk), int(mn), taper)
            assert e == H.chunk_ends(n, int(chunk), int(mn), taper)     # the restatement the GPU tests count launches with
            s = sizes(e)
            assert (n == 0 and e == []) or (e[-1] == n and all(x > 0 for x in s)), (n, chunk, mn)
            assert all(x <= chunk for x in s)
            if not taper or n <= chunk:
                assert len(e) == (n + chunk - 1) // chunk       # uniform pieces, as before
            else:
                # full pieces, then a tail that never grows; the drain (last piece) is at most min(chunk, minChunk) rays
                assert all(a >= b for a, b in zip(s, s[1:]))
                assert s[-1] <= min(chunk, mn)
                assert len(e) <= (n + chunk - 1) // chunk + 40
