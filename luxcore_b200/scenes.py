"""Benchmark / test geometry: the reference's scenes as committed fixtures plus synthetic generators.

A SceneDesc mirrors what slg::Scene::Preprocess hands to luxrays::DataSet::Add
(src/slg/scene/scenepreprocess.cpp:52-56): an ordered list of meshes -- plain TriangleMesh,
InstanceTriangleMesh(base, Transform) or MotionTriangleMesh(base, MotionSystem) -- whose position in
the list is the meshIndex reported in RayHit.
"""
import os

import numpy as np

FIXTURE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes")

PLAIN, INSTANCE, MOTION = 0, 1, 2


class MeshDesc:
    __slots__ = ("kind", "shape", "xform", "times", "motion_xforms")

    def __init__(self, kind, shape, xform=None, times=None, motion_xforms=None):
        self.kind = kind
        self.shape = shape                  # index into SceneDesc.shapes
        self.xform = xform                  # [4,4] row-major local->world (instances)
        self.times = times                  # [k] (motion)
        self.motion_xforms = motion_xforms  # [k,4,4] row-major, as stored by the MotionSystem (world->local)


class SceneDesc:
    def __init__(self, name):
        self.name = name
        self.shapes = []    # list of (verts float32 [V,3], tris uint32 [T,3])
        self.meshes = []    # list of MeshDesc, dataset order
        self.cam = None
        self.skipped = []

    def add_shape(self, verts, tris):
        self.shapes.append((np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3),
                            np.ascontiguousarray(tris, dtype=np.uint32).reshape(-1, 3)))
        return len(self.shapes) - 1

    def add_plain(self, shape):
        self.meshes.append(MeshDesc(PLAIN, shape))
        return len(self.meshes) - 1

    def add_instance(self, shape, xform):
        self.meshes.append(MeshDesc(INSTANCE, shape, xform=np.asarray(xform, dtype=np.float32).reshape(4, 4)))
        return len(self.meshes) - 1

    def add_motion(self, shape, times, xforms):
        self.meshes.append(MeshDesc(MOTION, shape, times=np.asarray(times, dtype=np.float32),
                                    motion_xforms=np.asarray(xforms, dtype=np.float32).reshape(-1, 4, 4)))
        return len(self.meshes) - 1

    @property
    def has_instances(self):
        return any(m.kind == INSTANCE for m in self.meshes)

    @property
    def has_motion(self):
        return any(m.kind == MOTION for m in self.meshes)

    def triangle_count(self):
        return int(sum(self.shapes[m.shape][1].shape[0] for m in self.meshes))

    def flattened(self):
        """World-space copy of every mesh (what a single-level BVH sees: Mesh::GetVertex with
        TRANS_IDENTITY, include/luxrays/core/trianglemesh.h:96,214-216).  Only static content:
        motion meshes are frozen untransformed like the reference (trianglemesh.h:319-321).
        -> verts [V,3], mesh_vertex_offsets [M], list of per-mesh tris."""
        verts, offs, tris = [], [], []
        total = 0
        for m in self.meshes:
            v, t = self.shapes[m.shape]
            if m.kind == INSTANCE:
                v = transform_points(m.xform, v)
            verts.append(v)
            offs.append(total)
            tris.append(t)
            total += v.shape[0]
        return (np.concatenate(verts).astype(np.float32) if verts else np.zeros((0, 3), np.float32),
                np.asarray(offs, dtype=np.uint32), tris)

    def bbox(self):
        v, _, _ = self.flattened()
        return v.min(axis=0), v.max(axis=0)


def transform_points(m, p):
    """Transform * Point (include/luxrays/core/geometry/transform.h:117-130) in float32, same
    operation order, divide by w only when w != 1."""
    m = np.asarray(m, dtype=np.float32)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    out = np.empty_like(p, dtype=np.float32)
    for r in range(3):
        out[:, r] = ((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3]
    w = ((m[3, 0] * x + m[3, 1] * y) + m[3, 2] * z) + m[3, 3]
    ne = w != np.float32(1.0)
    if ne.any():
        inv = (np.float32(1.0) / w[ne]).astype(np.float32)
        out[ne] = out[ne] * inv[:, None]
    return out


def available_fixtures():
    if not os.path.isdir(FIXTURE_DIR):
        return []
    return sorted(f[:-4] for f in os.listdir(FIXTURE_DIR) if f.endswith(".npz"))


def load_fixture(name, max_objects=None):
    """Load tests/golden/scenes/<name>.npz (written by tools/import_scenes.py)."""
    z = np.load(os.path.join(FIXTURE_DIR, name + ".npz"), allow_pickle=False)
    s = SceneDesc(name)
    vo, to = z["shape_vert_off"], z["shape_tri_off"]
    for i in range(len(vo) - 1):
        s.add_shape(z["verts"][vo[i]:vo[i + 1]], z["tris"][to[i]:to[i + 1]])
    kinds, shapes, xf = z["obj_kind"], z["obj_shape"], z["obj_xform"]
    mo, mt, mx = z["motion_obj"], z["motion_time"], z["motion_xform"]
    n = len(kinds) if max_objects is None else min(len(kinds), max_objects)
    for i in range(n):
        if kinds[i] == PLAIN:
            s.add_plain(int(shapes[i]))
        elif kinds[i] == INSTANCE:
            s.add_instance(int(shapes[i]), xf[i])
        else:
            sel = mo == i
            # the file stores local->world; scene objects keep world->local in the MotionSystem
            # (src/slg/scene/parseobjects.cpp:155-157)
            inv = np.stack([np.linalg.inv(m.astype(np.float64)).astype(np.float32) for m in mx[sel]])
            s.add_motion(int(shapes[i]), mt[sel], inv)
    s.cam = z["cam"].astype(np.float32)
    s.skipped = [str(x) for x in z["skipped"]]
    return s


def random_soup(n_tris, seed=4, size=0.002, name=None):
    """Config 5 generator: triangle k has vertices c + 0.5*size*u, c ~ U[0,1)^3, u ~ U[-1,1)^3."""
    rng = np.random.Generator(np.random.Philox(seed))
    c = rng.random((n_tris, 1, 3), dtype=np.float32)
    u = rng.random((n_tris, 3, 3), dtype=np.float32) * 2.0 - 1.0
    v = (c + np.float32(0.5 * size) * u).astype(np.float32).reshape(-1, 3)
    t = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    s = SceneDesc(name or ("soup%d" % n_tris))
    s.add_plain(s.add_shape(v, t))
    s.cam = np.asarray([0.5, -2.0, 0.5, 0.5, 0.5, 0.5, 0, 0, 1, 45], dtype=np.float32)
    return s


def grid_mesh(nx, ny, z=0.0, size=1.0):
    """Small regular test mesh: nx x ny quads split in two triangles."""
    xs = np.linspace(-size, size, nx + 1, dtype=np.float32)
    ys = np.linspace(-size, size, ny + 1, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys, indexing="xy")
    v = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, z, dtype=np.float32)], axis=1)
    tris = []
    for j in range(ny):
        for i in range(nx):
            a = j * (nx + 1) + i
            b, c, d = a + 1, a + nx + 1, a + nx + 2
            tris.append((a, b, d))
            tris.append((a, d, c))
    return v.astype(np.float32), np.asarray(tris, dtype=np.uint32)


def world_triangles(desc):
    """World-space (p0, e1, e2) of every triangle of every dataset mesh, plus the per-mesh offset of
    its first triangle -- what a path tracer's shading stage would look up from a RayHit.  Motion
    meshes are taken at their first key.  Used only to synthesise bounce-ray batches."""
    p0s, e1s, e2s, offs = [], [], [], []
    total = 0
    for m in desc.meshes:
        v, t = desc.shapes[m.shape]
        if m.kind == INSTANCE:
            v = transform_points(m.xform, v)
        elif m.kind == MOTION:
            v = transform_points(np.linalg.inv(m.motion_xforms[0].astype(np.float64)).astype(np.float32), v)
        a = v[t[:, 0]]
        p0s.append(a)
        e1s.append(v[t[:, 1]] - a)
        e2s.append(v[t[:, 2]] - a)
        offs.append(total)
        total += t.shape[0]
    return (np.concatenate(p0s).astype(np.float32), np.concatenate(e1s).astype(np.float32),
            np.concatenate(e2s).astype(np.float32), np.asarray(offs, dtype=np.int64))
