"""BASELINE.json configs[4] in the small: a 4 M-triangle random soup (0.5 GB on the device, four times the L2), the
HBM-resident regime of the 50 M-triangle workload.  The CUDA path is compared bit for bit with the reference library
(oracle/_ref; the oracle port where it is absent) walking the same BVHArrayNode array, on uniform rays AND on exactly
axis-parallel ones (d = +-e_k: 1/d = +-inf on two axes -- the rays that once walked half the tree),
in index order and in the sorted order the device picks for scenes larger than L2; the reference's 31-bit skip-index /
8-page limits (bvhaccelhw.cpp:38-257) do not apply to the flat arrays used here."""
import numpy as np
import pytest
import torch

import helpers as H
from luxcore_b200 import capi, hostapi, rays as R, scenes as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

N_TRIS = 4000000


@pytest.fixture(scope="module")
def soup():
    desc = S.random_soup(N_TRIS, seed=4, size=0.002 * (50e6 / N_TRIS) ** (1.0 / 3.0), name="soup")
    s = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    s.build_accelerator("BVH")
    s.start(0)
    yield desc, s
    s.stop()
    s.close()


def _axis_parallel(n, seed):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["o"] = rng.random((n, 3), dtype=np.float32)
    k = rng.integers(0, 3, n)
    sgn = np.where(rng.random(n) < 0.5, -1.0, 1.0).astype(np.float32)
    d = np.zeros((n, 3), np.float32)
    d[np.arange(n), k] = sgn
    rays["d"] = d
    rays["mint"] = 1e-5
    rays["maxt"] = np.inf
    return rays


def test_soup_parity_uniform_and_axis_parallel(soup):
    from oracle import refapi
    desc, s = soup
    info = s.native_scene().info()
    assert info.n_triangles == N_TRIS and info.device_bytes > 3 * 126e6        # does not fit L2
    nodes = s.bvh_nodes()
    uni = R.to_numpy_rays(R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 150000, seed=31))
    axis = _axis_parallel(50000, seed=32)
    rays = np.concatenate([uni, axis])
    checker = refapi.BVH(H.reference_scene(desc), nodes=nodes) if refapi.available() else O.BVH(H.oracle_scene(desc), nodes=nodes)
    ref = checker.intersect(rays)
    for opt in (0, 1):          # index order / sorted (octant + Morton key of the origin cell)
        s.set_option("sort_rays", opt)
        s.set_option("sort_min_rays", 0)
        got = s.trace_host(rays)
        rep = H.compare_hits(got, ref, rays, what="soup-4M sort_rays=%d" % opt)
        assert rep["hits"] > 0.5 * rep["n"]
        assert rep["bit_exact_hits"] == rep["hits"] and rep["index_mismatch"] == 0
    s.set_option("sort_rays", 2)


def test_soup_axis_parallel_rays_are_not_slow(soup):
    """An axis-parallel ray must keep its box culling (clamped reciprocal in the plane decode): the instrumented kernel
    counts the nodes it visits."""
    desc, s = soup
    dev = torch.device("cuda", 0)
    axis = torch.from_numpy(_axis_parallel(200000, seed=33).view(np.uint8).reshape(-1, 48)).to(dev)
    uni = R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 200000, seed=34, device=dev)
    torch.cuda.synchronize()
    scene = s.native_scene()
    st_a = scene.trace_stats(axis.data_ptr(), 0, axis.shape[0])
    st_u = scene.trace_stats(uni.data_ptr(), 0, uni.shape[0])
    per_a = st_a.wide_nodes / max(1, st_a.rays)
    per_u = st_u.wide_nodes / max(1, st_u.rays)
    print("soup-4M: wide nodes per ray: axis-parallel %.1f, uniform %.1f" % (per_a, per_u))
    assert per_a < 3.0 * per_u


def test_soup_full_batch_sorted_equals_index_order(soup):
    """8 Mi rays: the sorted launch (what sort_rays = auto does for this scene) returns the same bytes as index order."""
    desc, s = soup
    dev = torch.device("cuda", 0)
    rays = R.uniform_rays([0.0, 0.0, 0.0], [1.0, 1.0, 1.0], 8 << 20, seed=35, device=dev)
    a = torch.empty((rays.shape[0], 20), dtype=torch.uint8, device=dev)
    b = torch.empty_like(a)
    torch.cuda.synchronize()
    s.set_option("sort_rays", 0)
    s.trace_device(rays.data_ptr(), a.data_ptr(), rays.shape[0])
    s.set_option("sort_rays", 1)
    s.trace_device(rays.data_ptr(), b.data_ptr(), rays.shape[0])
    s.finish()
    s.set_option("sort_rays", 2)
    assert torch.equal(a, b)
    hit = a.view(torch.int32)[:, 3] != -1
    assert 0.5 < hit.float().mean().item() < 1.0
