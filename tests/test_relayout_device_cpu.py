"""The DEVICE re-layout's per-record code (luxcore_b200/csrc/relayout_kernels.cuh + relayout_shared.h: what the kernels of
lrb_bvh_build_scene run, one thread per record) compiled for the host and driven in the device pipeline's order, held to the
BYTES of the host re-layout (relayout.cpp BuildWideBVH, what lrb_bvh_upload runs) of the same reference array.  On the GPU
box tests/test_gpu_builder.py repeats the comparison with the real kernels."""
import numpy as np
import pytest

import helpers as H
from luxcore_b200 import hostapi
from luxcore_b200 import scenes as S
from oracle import oracle as O


def _arrays(desc):
    osc = H.oracle_scene(desc)
    verts, voff = H.flattened_from_oracle(desc, osc)
    tri, toff = H.flattened_triangles(desc)
    return osc, verts, voff, tri, toff


def _check(nodes, verts, voff, tri, toff, what):
    emu = H.Emu.bvh(nodes, verts, voff)
    wide, tris, ids = emu.arrays()
    dev = H.RelayoutDev.run(H.to_builder_format(nodes, toff), verts, voff, tri, toff)
    # the leaf payload written on the device gives back the array a host builder would have made
    assert dev["ref_nodes"].tobytes() == np.ascontiguousarray(nodes).tobytes(), what + ": reference array"
    assert dev["wide"].shape == wide.shape and dev["wide"].tobytes() == wide.tobytes(), what + ": wide nodes"
    assert dev["tris"].tobytes() == tris.tobytes(), what + ": triangle records"
    assert dev["ids"].tobytes() == ids.tobytes(), what + ": triangle ids"
    assert dev["stack_need"] == emu.info()["stack_need"], what + ": stack bound"
    # exact root box = the reference's node 0
    root = np.ascontiguousarray(nodes)[0]["w"].view(np.float32)
    assert np.array_equal(dev["entry_box"], root), what + ": root box"
    return dev


@pytest.mark.parametrize("name", ["cornell", "bigmonkey", "kitchen", "luxball"])
@pytest.mark.parametrize("tree_type", [2, 4, 8])
def test_device_relayout_bodies_equal_the_host_relayout_classic_trees(name, tree_type):
    desc = S.load_fixture(name)
    osc, verts, voff, tri, toff = _arrays(desc)
    nodes = O.BVH(osc, tree_type=tree_type).nodes()
    _check(nodes, verts, voff, tri, toff, "%s CLASSIC k=%d" % (name, tree_type))


@pytest.mark.parametrize("name,tree_type", [("kitchen", 4), ("kitchen", 8), ("classroom", 4)])
def test_device_relayout_bodies_equal_the_host_relayout_sah_trees(name, tree_type):
    desc = S.load_fixture(name)
    osc, verts, voff, tri, toff = _arrays(desc)
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": tree_type}, desc)
    sess.build_accelerator("BVH")
    _check(sess.bvh_nodes(), verts, voff, tri, toff, "%s SAH k=%d" % (name, tree_type))


def test_device_relayout_bodies_on_a_soup_and_its_build_boxes():
    n = 200000
    desc = S.random_soup(n, seed=4, size=0.002 * (50e6 / n) ** (1.0 / 3.0), name="soup")
    osc, verts, voff, tri, toff = _arrays(desc)
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    dev = _check(sess.bvh_nodes(), verts, voff, tri, toff, "soup")
    # LeafBoxKernel's boxes: bounds of the three vertices grown by MachineEpsilon::E of the bounds (bvhaccel.cpp:116-122),
    # recomputed here with numpy from the definition in epsilon.h:48-86
    p = verts[tri.reshape(-1)].reshape(-1, 3, 3)
    lo, hi = p.min(axis=1), p.max(axis=1)

    def eps(v):
        bumped = (v.view(np.uint32) + np.uint32(0x80)).view(np.float32)
        return np.clip(np.abs(bumped - v), np.float32(1e-5), np.float32(1e-1)).astype(np.float32)
    e = np.maximum(eps(lo).max(axis=1), eps(hi).max(axis=1))[:, None]
    assert np.array_equal(dev["boxes"][:, :3], lo - e) and np.array_equal(dev["boxes"][:, 3:], hi + e)


def test_mesh_lookup_skips_empty_meshes():
    """Leaf payload with empty meshes in the table: triangle g belongs to the LAST mesh whose offset is <= g."""
    desc = S.load_fixture("cornell")
    osc, verts, voff, tri, toff = _arrays(desc)
    nodes = O.BVH(osc, tree_type=4).nodes()
    # insert an empty mesh after mesh 0: mesh indices >= 1 shift by one
    voff2 = np.insert(voff, 1, voff[1]).astype(np.uint32)
    toff2 = np.insert(toff, 1, toff[1]).astype(np.uint32)
    shifted = np.array(nodes, copy=True)
    leaf = (shifted["nodeData"] & 0x80000000) != 0
    w = shifted["w"]
    w[leaf, 3] = np.where(w[leaf, 3] >= 1, w[leaf, 3] + 1, w[leaf, 3])
    _check(shifted, verts, voff2, tri, toff2, "cornell + empty mesh")


def test_threaded_host_relayout_is_the_serial_one(monkeypatch):
    """Arrays of 400 000 nodes and more are laid out on several threads (index pass: counts per range + exclusive prefix;
    fill pass: ranges of the array) into buffers that are not zero-filled first (relayout.h RawVector): the bytes must not
    depend on the thread count, and equal the device re-layout's."""
    import hashlib
    n = 300000
    desc = S.random_soup(n, seed=11, size=0.002 * (50e6 / n) ** (1.0 / 3.0), name="soup")
    osc, verts, voff, tri, toff = _arrays(desc)
    monkeypatch.setenv("LRB_BVH_OPT", "0")          # the plain top-down builder: any tree will do, this one is quick
    sess = hostapi.Session({"accelerator.type": "BVH", "accelerator.bvh.builder.type": "EMBREE_BINNED_SAH", "accelerator.bvh.treetype": 4}, desc)
    sess.build_accelerator("BVH")
    nodes = sess.bvh_nodes()
    assert nodes.shape[0] >= 400000
    digests = set()
    for threads in ("1", "2", "5", "16"):
        monkeypatch.setenv("LRB_RELAYOUT_THREADS", threads)
        emu = H.Emu.bvh(nodes, verts, voff)
        w, t, i = emu.arrays()
        digests.add((hashlib.sha256(w.tobytes() + t.tobytes() + i.tobytes()).hexdigest(), emu.info()["stack_need"]))
    assert len(digests) == 1
    monkeypatch.setenv("LRB_RELAYOUT_THREADS", "7")
    _check(nodes, verts, voff, tri, toff, "soup 300k, 7 threads")


@pytest.mark.parametrize("quality,tree_type", [(0, 4), (1, 4), (1, 8), (0, 2)])
def test_whole_device_pipeline_restated_on_the_host(quality, tree_type):
    """lrb_bvh_build_scene end to end without a GPU: LeafBoxBody's boxes -> the Python restatement of the builder kernels
    (tests/builder_reference.py: leaves carry the input triangle number, as EmitKernel writes them) -> leaf payload, indices,
    fill and stack bound by the device bodies.  The result must be the host re-layout of the payload-carrying array, and the
    emulated traversal of it must give the oracle's closest hits."""
    import builder_reference as BR
    from luxcore_b200 import rays as R
    desc = S.load_fixture("cornell")
    rng = np.random.default_rng(5)
    c = rng.random((700, 1, 3)) * 4 - 2
    v = (c + rng.normal(scale=0.15, size=(700, 3, 3))).astype(np.float32).reshape(-1, 3)
    desc.add_plain(desc.add_shape(v, np.arange(2100, dtype=np.uint32).reshape(-1, 3)))      # cornell + a 700-triangle cloud
    osc, verts, voff, tri, toff = _arrays(desc)
    # the boxes the device computes (compared with the definition in the soup test above)
    probe = H.RelayoutDev.run(H.to_builder_format(O.BVH(osc, tree_type=4).nodes(), toff), verts, voff, tri, toff)
    built = BR.build(probe["boxes"], tree_type, quality)        # builder output format
    dev = H.RelayoutDev.run(built, verts, voff, tri, toff)
    nodes = dev["ref_nodes"].astype(O.NODE_DTYPE)
    # payload: every triangle once, with its own vertex indices
    leaf = (nodes["nodeData"] >> 31) == 1
    w = nodes["w"][leaf]
    g = toff[w[:, 3]].astype(np.int64) + w[:, 4]
    assert np.array_equal(np.sort(g), np.arange(tri.shape[0])) and np.array_equal(w[:, :3], tri[g])
    emu = H.Emu.bvh(nodes, verts, voff)
    wide, tris, ids = emu.arrays()
    assert dev["wide"].tobytes() == wide.tobytes() and dev["tris"].tobytes() == tris.tobytes() and dev["ids"].tobytes() == ids.tobytes()
    assert dev["stack_need"] == emu.info()["stack_need"]
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), 20000, seed=3))
    rep = H.compare_hits(emu.trace(rays), O.BVH(osc, nodes=nodes).intersect(rays), rays, what="device pipeline restated")
    assert rep["bit_exact_hits"] == rep["hits"] and rep["hits"] > 0
