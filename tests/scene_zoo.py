"""Small synthetic two-level scenes for the MBVH tests (instances with rotation / non-uniform
scale / mirroring, motion with translation, rotation and scale, shared and unshared base meshes,
one-triangle meshes)."""
import math

import numpy as np

from luxcore_b200 import scenes as S


def rot_z(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0, 0], [s, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float32)


def rot_x(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=np.float32)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = [x, y, z]
    return m


def scale(x, y, z):
    return np.diag([x, y, z, 1]).astype(np.float32)


def blob(n, seed, radius=0.5):
    """A bumpy closed-ish surface: icosphere-like point cloud triangulated by lat/long."""
    rng = np.random.default_rng(seed)
    nu, nv = n, n // 2
    v = []
    for j in range(nv + 1):
        th = math.pi * j / nv
        for i in range(nu):
            ph = 2 * math.pi * i / nu
            r = radius * (1.0 + 0.15 * rng.standard_normal())
            v.append((r * math.sin(th) * math.cos(ph), r * math.sin(th) * math.sin(ph), r * math.cos(th)))
    t = []
    for j in range(nv):
        for i in range(nu):
            a = j * nu + i
            b = j * nu + (i + 1) % nu
            c = (j + 1) * nu + i
            d = (j + 1) * nu + (i + 1) % nu
            t.append((a, b, d))
            t.append((a, d, c))
    return np.asarray(v, dtype=np.float32), np.asarray(t, dtype=np.uint32)


def inv(m):
    return np.linalg.inv(m.astype(np.float64)).astype(np.float32)


def instances_scene(n_inst=40, seed=5):
    rng = np.random.default_rng(seed)
    s = S.SceneDesc("zoo-instances")
    a = s.add_shape(*blob(16, 1))
    b = s.add_shape(*blob(10, 2, radius=0.3))
    floor = s.add_shape(*S.grid_mesh(8, 8, z=-1.0, size=6.0))
    one = s.add_shape(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32))
    s.add_plain(floor)
    s.add_plain(a)                      # the base mesh itself is also a dataset entry
    for i in range(n_inst):
        m = translate(*(rng.uniform(-4, 4, 3))) @ rot_z(rng.uniform(0, 360)) @ rot_x(rng.uniform(0, 360)) @ \
            scale(*rng.uniform(0.4, 1.8, 3))
        if i % 7 == 0:
            m = m @ scale(-1, 1, 1)     # mirrored instance (swaps handedness)
        s.add_instance(a if i % 3 else b, m.astype(np.float32))
    s.add_instance(one, translate(0.5, 0.5, 2.0))   # leaf tree that is a single triangle
    s.add_plain(one)
    s.add_plain(a)                      # same TriangleMesh added twice as a plain mesh
    s.cam = np.asarray([0, -12, 3, 0, 0, 0, 0, 0, 1, 50], dtype=np.float32)
    return s


def motion_scene(seed=6):
    rng = np.random.default_rng(seed)
    s = S.SceneDesc("zoo-motion")
    a = s.add_shape(*blob(14, 3))
    floor = s.add_shape(*S.grid_mesh(6, 6, z=-1.2, size=6.0))
    s.add_plain(floor)
    # translation-only, 2 keys
    s.add_motion(a, [0.0, 1.0], [inv(translate(-3, 0, 0)), inv(translate(-2, 0.5, 0))])
    # rotation + translation, 3 keys, not starting at 0
    k = [translate(0, 0, 0) @ rot_z(0), translate(0.5, 0, 0.2) @ rot_z(40), translate(1.0, 0.3, 0.2) @ rot_z(95) @ rot_x(20)]
    s.add_motion(a, [0.2, 0.6, 0.9], [inv(m) for m in k])
    # scale + rotation
    k = [translate(3, 0, 0) @ scale(1, 1, 1), translate(3, 0, 0.5) @ rot_x(30) @ scale(1.5, 0.8, 1.2)]
    s.add_motion(a, [0.0, 1.0], [inv(m) for m in k])
    # static motion system (identical keys) and an instance next to it
    s.add_motion(a, [0.0, 1.0], [inv(translate(0, 3, 0)), inv(translate(0, 3, 0))])
    s.add_instance(a, translate(0, -3, 0) @ rot_z(10))
    for i in range(6):
        m0 = translate(*rng.uniform(-4, 4, 3)) @ rot_z(rng.uniform(0, 360))
        m1 = m0 @ translate(*rng.uniform(-0.5, 0.5, 3)) @ rot_x(rng.uniform(-40, 40))
        s.add_motion(a, [0.0, 1.0], [inv(m0), inv(m1)])
    s.cam = np.asarray([0, -12, 3, 0, 0, 0, 0, 0, 1, 50], dtype=np.float32)
    return s
