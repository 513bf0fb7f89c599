"""CPU-only checks of the product's re-layout + per-ray traversal body against the oracle.

The traversal body (luxcore_b200/csrc/traverse.h) is compiled for the host by tests/cpp only for
these tests; on the GPU box the same body runs inside the CUDA kernels (tests/test_gpu_*.py)."""
import numpy as np
import pytest
import torch

import helpers as H
from luxcore_b200 import rays as R
from luxcore_b200 import scenes as S
from oracle import oracle as O


def _rays_for(desc, n, seed):
    lo, hi = desc.bbox()
    pad = 0.05 * (hi - lo)
    a = R.to_numpy_rays(R.uniform_rays(lo - pad, hi + pad, n, seed=seed))
    side = int(np.sqrt(n))
    b = R.to_numpy_rays(R.camera_rays(desc.cam, side, side, seed=seed + 1))
    return np.concatenate([a, b])


@pytest.mark.parametrize("name,tree_type", [("cornell", 4), ("cornell", 2), ("cornell", 8), ("bigmonkey", 4), ("kitchen", 4)])
def test_bvh_emulation_matches_oracle(name, tree_type):
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    nodes = bvh.nodes()
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    rays = _rays_for(desc, 20000 if name != "kitchen" else 40000, seed=11)
    ref = bvh.intersect(rays)
    got, st = emu.trace(rays, want_stats=True)
    rep = H.compare_hits(got, ref, rays, what="%s k=%d" % (name, tree_type))
    assert rep["hits"] > 0.3 * rep["n"]
    # single-level BVH: identical arithmetic => identical floats
    assert rep["bit_exact_hits"] == rep["hits"]
    assert st["max_stack"] <= emu.info()["stack_need"]


@pytest.mark.parametrize("name,tree_type,n", [("cornell", 4, 60000), ("bigmonkey", 4, 60000), ("kitchen", 4, 120000), ("kitchen", 8, 40000),
                                              ("luxball", 2, 40000)])
def test_bvh_emulation_grazing_rays(name, tree_type, n):
    """Rays that start on (or a few epsilons off) the surfaces, axis-parallel ones included: their
    hits lie at the very planes of the leaf boxes and of the quantized node grids."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    p0, e1, e2, _ = S.world_triangles(desc)
    rays = R.to_numpy_rays(R.surface_rays(p0, e1, e2, n, seed=17))
    # also from far outside the scene (large |o| against small boxes)
    lo, hi = desc.bbox()
    far = R.to_numpy_rays(R.uniform_rays(lo - 40 * (hi - lo), hi + 40 * (hi - lo), n // 4, seed=19))
    c = 0.5 * (lo + hi)
    far["d"] = (c[None, :] + 0.3 * (hi - lo)[None, :] * (np.random.default_rng(5).random((far.shape[0], 3)).astype(np.float32) - 0.5)) - far["o"]
    rays = np.concatenate([rays, far])
    ref = bvh.intersect(rays)
    got = emu.trace(rays)
    rep = H.compare_hits_tie_aware(got, ref, rays, osc, what="grazing %s k=%d" % (name, tree_type))
    assert rep["hits"] > 0.3 * rep["n"]
    assert rep["bit_exact_hits"] == rep["hits"]


@pytest.mark.parametrize("name,tree_type", [("cornell", 4), ("kitchen", 4), ("kitchen", 8), ("bigmonkey", 2)])
def test_quantized_nodes_contain_the_reference_boxes(name, tree_type):
    """Layout invariants of the 64-byte nodes (layout.h): every slot box, decoded in exact arithmetic
    as org + q * step, contains the reference box of that child (inner node: its BVHArrayNode box;
    triangle: the bounds of its vertices) with a margin on both sides, unused slots are inverted, every
    reference leaf appears exactly once, and every gate is the box of the leaf's parent."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=tree_type)
    nodes = bvh.nodes()
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(nodes, verts, offs)
    wide, tris, ids = emu.arrays()
    raw = nodes.view(np.uint8).reshape(len(nodes), 32)
    ref_box = raw[:, :24].copy().view(np.float32).reshape(-1, 6).astype(np.float64)
    nd = nodes["nodeData"]
    is_leaf = (nd >> 31) == 1
    assert len(tris) == int(is_leaf.sum())
    # triangle records are stored in the order of the leaves in the reference array (the record index is the tie-break order)
    assert tris["order"].tolist() == np.nonzero(is_leaf)[0].tolist()
    leaf_raw = raw[tris["order"]]
    assert np.array_equal(ids["meshIndex"], leaf_raw[:, 12:16].copy().view(np.uint32).ravel())
    assert np.array_equal(ids["triangleIndex"], leaf_raw[:, 16:20].copy().view(np.uint32).ravel())

    # parent of every reference node (depth-first array, skip index = end of the subtree)
    parent = np.full(len(nodes), -1, dtype=np.int64)
    stack = []
    for i in range(len(nodes)):
        while stack and (nd[stack[-1]] & 0x7FFFFFFF) <= i:
            stack.pop()
        if stack:
            parent[i] = stack[-1]
        if not is_leaf[i]:
            stack.append(i)
    # gates: exact box of the parent
    par = parent[tris["order"]]
    assert (par >= 0).all()
    assert np.array_equal(tris["gateLo"].astype(np.float64), ref_box[par, :3]) and np.array_equal(tris["gateHi"].astype(np.float64), ref_box[par, 3:])

    org = wide["org"].astype(np.float64)
    eb = np.stack([(wide["exps"] >> (8 * a)) & 0xFF for a in range(3)], axis=1).astype(np.int64)
    step = np.ldexp(1.0, eb - 15 - 127)
    n_slots = wide["exps"] >> 24
    tri_lo = np.minimum(np.minimum(tris["p0"], tris["p1"]), tris["p2"]).astype(np.float64)
    tri_hi = np.maximum(np.maximum(tris["p0"], tris["p1"]), tris["p2"]).astype(np.float64)
    # wide index -> reference inner node: the k-th inner node in array order owns wide node 1 + (nodes before it)
    checked_inner = checked_tri = 0
    for k in range(4):
        ql = np.stack([(wide["qlo"][:, a] >> (8 * k)) & 0xFF for a in range(3)], axis=1).astype(np.float64)
        qh = np.stack([(wide["qhi"][:, a] >> (8 * k)) & 0xFF for a in range(3)], axis=1).astype(np.float64)
        lo = org + ql * step
        hi = org + qh * step
        used = n_slots > k
        ref = wide["child"][:, k]
        assert (ref[~used] == 0xFFFFFFFF).all()
        assert ((ql[~used] == 255) & (qh[~used] == 0)).all()         # inverted: nothing passes
        is_tri = used & ((ref & 0xC0000000) == 0x40000000)
        t = ref[is_tri] & 0x3FFFFFFF
        m = step[is_tri] / 64.0
        assert (lo[is_tri] <= tri_lo[t] - m).all() and (hi[is_tri] >= tri_hi[t] + m).all()
        assert (ql[is_tri] >= 1).all() and (qh[is_tri] <= 254).all()
        checked_tri += int(is_tri.sum())
        is_inner = used & ((ref & 0xC0000000) == 0)
        checked_inner += int(is_inner.sum())
    assert checked_tri == len(tris)
    # inner children: rebuild the reference-node -> wide-node map (one wide node per inner node, plus
    # continuation nodes, in array order after the entry node) and check every inner child's own box
    skip = (nd & 0x7FFFFFFF).astype(np.int64)
    kids_of = {}
    for i in np.nonzero(~is_leaf)[0]:
        ks, c = [], i + 1
        while c < skip[i]:
            ks.append(c)
            c = skip[c] if not is_leaf[c] else c + 1
        kids_of[int(i)] = ks
    wide_of, idx = {}, 1
    for i in sorted(kids_of):
        wide_of[i] = idx
        idx += max(1, (len(kids_of[i]) + 3) // 4)
    assert idx == len(wide)
    assert wide["child"][0, 0] == wide_of[0] and (wide["flags"][0] & 1) == 1      # entry node -> root
    # slot order inside a node: ascending size of the slot's box (relayout.cpp SlotSizeKey): inner child = its
    # reference box, triangle = bounds of the vertices grown by MachineEpsilon::E (bvhaccel.cpp:116-122)
    def eps32(v):
        v = np.asarray(v, dtype=np.float32)
        e = np.abs((v.view(np.int32) + 0x80).view(np.float32) - v)
        return np.clip(e, np.float32(1e-5), np.float32(1e-1))

    def half_area(lo, hi):
        d = np.abs(hi.astype(np.float64) - lo.astype(np.float64))
        return d[0] * d[1] + d[1] * d[2] + d[2] * d[0]

    ref_box32 = raw[:, :24].copy().view(np.float32).reshape(-1, 6)
    tlo32 = np.minimum(np.minimum(tris["p0"], tris["p1"]), tris["p2"])
    thi32 = np.maximum(np.maximum(tris["p0"], tris["p1"]), tris["p2"])
    inner_of_wide = {w: i for i, w in wide_of.items()}
    n_checked = 0
    for i, ks in kids_of.items():
        n_w = max(1, (len(ks) + 3) // 4)
        seen, sizes = [], []
        for j in range(n_w):
            w = wide_of[i] + j
            assert n_slots[w] == min(4, len(ks) - 4 * j)
            for slot in range(int(n_slots[w])):
                ref = int(wide["child"][w, slot])
                if (ref & 0xC0000000) == 0x40000000:
                    t = ref & 0x3FFFFFFF
                    c = int(tris["order"][t])
                    assert is_leaf[c]
                    e = np.float32(max(eps32(tlo32[t]).max(), eps32(thi32[t]).max()))
                    sizes.append(half_area(tlo32[t] - e, thi32[t] + e))
                else:
                    c = inner_of_wide[ref]
                    assert not is_leaf[c]
                    ql = np.array([(wide["qlo"][w, a] >> (8 * slot)) & 0xFF for a in range(3)], dtype=np.float64)
                    qh = np.array([(wide["qhi"][w, a] >> (8 * slot)) & 0xFF for a in range(3)], dtype=np.float64)
                    lo, hi = org[w] + ql * step[w], org[w] + qh * step[w]
                    assert (lo <= ref_box[c, :3] - step[w] / 64.0).all() and (hi >= ref_box[c, 3:] + step[w] / 64.0).all()
                    sizes.append(half_area(ref_box32[c, :3], ref_box32[c, 3:]))
                    n_checked += 1
                seen.append(c)
            assert wide["next"][w] == (w + 1 if j + 1 < n_w else 0xFFFFFFFF)
        assert sorted(seen) == sorted(int(c) for c in ks)          # every child of the reference node, once
        assert all(a <= b for a, b in zip(sizes, sizes[1:])), (i, sizes)
    assert n_checked == int((~is_leaf).sum()) - 1


@pytest.mark.parametrize("name", ["kitchen", "cornell"])
def test_axis_parallel_rays_keep_their_box_culling(name):
    """Zero direction components (1/d = inf) must not switch a slab test off: an axis-parallel ray
    has to visit about as few nodes as a generic one, not every box in its column (regression:
    15 000 node visits per ray on the kitchen, seconds per ray on a 50 M-triangle scene)."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    lo, hi = desc.bbox()
    base = R.to_numpy_rays(R.uniform_rays(lo, hi, 3000, seed=77))
    _, st0 = emu.trace(base, want_stats=True)
    generic = st0["wide_nodes"] / st0["rays"]
    for d in ([0, 0, 1], [0, -1, 0], [1, 0, 0], [0.6, 0.8, 0.0], [0.0, -0.6, 0.8]):
        r = base.copy()
        r["d"][:] = np.asarray(d, dtype=np.float32)
        got, st = emu.trace(r, want_stats=True)
        rep = H.compare_hits_tie_aware(got, bvh.intersect(r), r, osc, what="axis %r" % (d,))
        assert rep["bit_exact_hits"] == rep["hits"]
        assert st["wide_nodes"] / st["rays"] < 4 * generic + 8, (d, st["wide_nodes"] / st["rays"], generic)


def _tiny_tree():
    """Three triangles under one inner node, in the reference array format."""
    verts = np.asarray([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 1], [3, 0, 1], [2, 1, 1], [0, 2, 2], [1, 2, 2], [0, 3, 2]], dtype=np.float32)
    nodes = np.zeros(4, dtype=H.Emu.NODE_DTYPE if hasattr(H.Emu, "NODE_DTYPE") else np.dtype([("w", "<u4", 6), ("nodeData", "<u4"), ("pad0", "<i4")]))
    w = nodes.view(np.uint32).reshape(-1, 8)
    w[0, :6] = np.asarray([-0.1, -0.1, -0.1, 3.1, 3.1, 2.1], dtype=np.float32).view(np.uint32)
    w[0, 6] = 4
    for k in range(3):
        w[1 + k, :5] = [3 * k, 3 * k + 1, 3 * k + 2, 0, k]
        w[1 + k, 6] = 0x80000000 | (k + 2)
    return nodes, verts, np.zeros(1, dtype=np.uint32)


def test_relayout_rejects_malformed_input():
    """The device refuses arrays it could walk off: bad skip indices, vertex / mesh indices out of range,
    non-finite boxes -- with an error, not a crash (lrb_bvh_upload returns LRB_ERR_INVALID for the same input)."""
    nodes, verts, offs = _tiny_tree()
    emu = H.Emu.bvh(nodes, verts, offs)                      # the well-formed tree is accepted ...
    rays = R.to_numpy_rays(R.uniform_rays([-1, -1, 3], [3, 3, 4], 64, seed=1))
    rays["d"][:] = [0, 0, -1]
    assert (emu.trace(rays)["meshIndex"] != H.NULL).any()     # ... and hit from above

    def broken(mutate, what):
        n, v, o = _tiny_tree()
        mutate(n.view(np.uint32).reshape(-1, 8), v)
        with pytest.raises(RuntimeError) as e:
            H.Emu.bvh(n, v, o)
        assert what in str(e.value), str(e.value)

    broken(lambda w, v: w.__setitem__((0, 6), 5), "root skip index")
    broken(lambda w, v: w.__setitem__((2, 6), 0x80000000 | 7), "leaf skip index")
    broken(lambda w, v: w.__setitem__((1, 0), 1000), "vertex outside")
    broken(lambda w, v: w.__setitem__((1, 3), 5), "mesh outside")
    # boxes the reference tolerates are tolerated the same way: a NaN plane never rejects, corners given in
    # the wrong order behave like the sorted box (BBox::IntersectP swaps the slab distances)
    n2, v2, o2 = _tiny_tree()
    w2 = n2.view(np.uint32).reshape(-1, 8)
    w2[0, 0] = np.asarray([np.nan], np.float32).view(np.uint32)[0]
    w2[0, 1], w2[0, 4] = w2[0, 4], w2[0, 1]
    osc = O.Scene()
    osc.add_plain(osc.add_shape(v2, np.arange(9, dtype=np.uint32).reshape(3, 3)))
    rep = H.compare_hits(H.Emu.bvh(n2, v2, o2).trace(rays), O.BVH(osc, nodes=n2).intersect(rays), rays, what="tolerated boxes")
    assert rep["hits"] > 0 and rep["bit_exact_hits"] == rep["hits"]
    assert H.Emu.lib().emu_validate_tree(nodes.ctypes.data, 3) != 0      # truncated array


def test_scheduling_model_is_consistent_with_the_per_ray_emulation():
    """tools/warp_model.py's engine (WarpSim: the persistent kernel's re-fill / Resolve / phase-vote control
    flow around the real traverse.h bodies) must produce the same hits and do the same amount of traversal
    work as the plain per-ray emulation, for any policy setting -- it only regroups the work into warps."""
    desc = S.load_fixture("kitchen")
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    lo, hi = desc.bbox()
    rays = R.to_numpy_rays(R.uniform_rays(lo, hi, 20000, seed=5))
    rays["flags"][::7] = 1      # masked rays are skipped, their RayHit stays untouched (zero here)
    ref, st = emu.trace(rays, want_stats=True)
    for n_warps, refill, bias in ((1, 24, 8), (64, 24, 8), (16, 32, 2), (16, 1, 64)):
        hits, c = H.warp_sim(emu, rays, n_warps=n_warps, refill_below=refill, tri_bias=bias)
        assert hits.tobytes() == ref.tobytes()
        assert c["rays"] == st["rays"] and c["node_lanes"] == st["wide_nodes"] and c["tri_lanes"] == st["triangles"]
        assert c["node_phases"] * 32 >= c["node_lanes"] and c["tri_phases"] * 32 >= c["tri_lanes"]


@pytest.mark.parametrize("name,n", [("cornell", 400000), ("kitchen", 400000), ("bigmonkey", 200000)])
def test_in_plane_rays_bound_the_known_residual(name, n):
    """DESIGN.md "Parity", known residual -- BUILT.  Rays lying in the plane of a triangle in general position make
    Triangle::Intersect (triangle.h:55-89) divide rounding noise by rounding noise: its t / b1 / b2 are arbitrary, it
    "hits" triangles the ray misses by far, and which of those artefacts survives depends on the order in which the
    reference's fixed depth-first walk meets them.  A traversal in another order (ours: near to far, triangles culled
    by their own boxes) cannot reproduce that order dependence.  This test bounds the class: on a batch made ONLY of such
    rays fewer than 1 in 1 000 answers differ, and every differing ray is coplanar with the triangle one of the two
    sides names (i.e. belongs to the artefact class, not to ordinary geometry)."""
    desc = S.load_fixture(name)
    osc = H.oracle_scene(desc)
    bvh = O.BVH(osc, tree_type=4)
    verts, offs = H.flattened_from_oracle(desc, osc)
    emu = H.Emu.bvh(bvh.nodes(), verts, offs)
    rays = H.in_plane_rays(desc, n, seed=1)
    ref = bvh.intersect(rays)
    got = emu.trace(rays)
    hit = ref["meshIndex"] != H.NULL
    diff = (got["meshIndex"] != ref["meshIndex"]) | (hit & (got["triangleIndex"] != ref["triangleIndex"]))
    print("%s: %d of %d in-plane rays answer differently" % (name, int(diff.sum()), n))
    assert diff.sum() <= n * 1e-3       # observed: 0.8e-5 (kitchen), 3.5e-5 (cornell), 5.9e-4 (bigmonkey: every triangle in general position)
    artefact = H.coplanar_with_reported(desc, rays, ref) | H.coplanar_with_reported(desc, rays, got)
    assert artefact[diff].all()
    # everywhere else the records are the reference's bit for bit
    same = ~diff & hit
    assert (got["t"][same].view(np.uint32) == ref["t"][same].view(np.uint32)).all()
