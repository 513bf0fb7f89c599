#!/usr/bin/env python3
"""Convert the reference's benchmark scenes into compact geometry fixtures.

Runs ONLY in the build container (it reads /root/reference/scenes, which does not exist on the GPU
box).  For every scene named in BASELINE.json it parses the `.scn` property file the way
slg::Scene::CreateObject does (src/slg/scene/parseobjects.cpp:87-181):

  * `scene.objects.<name>.ply`            -> a shape, loaded once per file name (mesh cache)
  * `.transformation` (16 floats, COLUMN-major, src/luxrays/utils/properties.cpp:606-614)
                                          -> InstanceTriangleMesh(shape, Transform(m))
  * `.motion.N.time/.transformation`      -> MotionTriangleMesh(shape, MotionSystem(times, Inverse(T)))
  * neither                               -> the plain TriangleMesh
  * `.appliedtransformation`              -> metadata only (vertices are already transformed)

PLY files are read like ExtTriangleMesh::Load (src/luxrays/core/exttrianglemeshfile.cpp:207-239):
only x/y/z and vertex_indices matter here, quads (a,b,c,d) become (a,b,c) + (a,c,d), polygons with
another vertex count are dropped.  Missing PLY files (.MISSING_LARGE_BLOBS) are skipped and listed.

Output: tests/golden/scenes/<scene>.npz with
  shape_vert_off / shape_tri_off : prefix offsets into verts / tris per unique shape
  verts float32 [V,3], tris uint32 [T,3]
  obj_shape int32 [O], obj_kind int8 [O] (0 plain, 1 instance, 2 motion)
  obj_xform float32 [O,4,4]  row-major local->world (instances; identity otherwise)
  motion_obj int32 [K], motion_time float32 [K], motion_xform float32 [K,4,4]  (row-major, as
      written in the file = local->world; the loader inverts them like parseobjects.cpp:155-157)
  cam float32 [10] = orig(3) target(3) up(3) fov
  skipped: names of objects whose PLY is missing
"""
import os
import re
import shlex
import struct
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes")

PLY_TYPES = {
    "char": ("b", 1), "int8": ("b", 1), "uchar": ("B", 1), "uint8": ("B", 1),
    "short": ("h", 2), "int16": ("h", 2), "ushort": ("H", 2), "uint16": ("H", 2),
    "int": ("i", 4), "int32": ("i", 4), "uint": ("I", 4), "uint32": ("I", 4),
    "float": ("f", 4), "float32": ("f", 4), "double": ("d", 8), "float64": ("d", 8),
}


def read_ply(path):
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header")
    end = data.index(b"\n", end) + 1
    header = data[:end].decode("ascii", "replace").splitlines()
    fmt = None
    elements = []
    for line in header:
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "format":
            fmt = tok[1]
        elif tok[0] == "element":
            elements.append({"name": tok[1], "count": int(tok[2]), "props": []})
        elif tok[0] == "property":
            if tok[1] == "list":
                elements[-1]["props"].append(("list", tok[2], tok[3], tok[4]))
            else:
                elements[-1]["props"].append(("scalar", tok[1], tok[2]))
    verts = None
    tris = []
    if fmt == "ascii":
        toks = data[end:].split()
        pos = 0
        for el in elements:
            for _ in range(el["count"]):
                row = {}
                for p in el["props"]:
                    if p[0] == "scalar":
                        row[p[2]] = float(toks[pos]); pos += 1
                    else:
                        n = int(toks[pos]); pos += 1
                        row[p[3]] = [int(float(x)) for x in toks[pos:pos + n]]; pos += n
                if el["name"] == "vertex":
                    if verts is None:
                        verts = []
                    verts.append((row["x"], row["y"], row["z"]))
                elif el["name"] == "face":
                    idx = row.get("vertex_indices", row.get("vertex_index"))
                    if idx is None:
                        continue
                    if len(idx) == 3:
                        tris.append(tuple(idx))
                    elif len(idx) == 4:
                        tris.append((idx[0], idx[1], idx[2]))
                        tris.append((idx[0], idx[2], idx[3]))
        verts = np.asarray(verts, dtype=np.float32)
    else:
        endian = "<" if fmt == "binary_little_endian" else ">"
        pos = end
        for el in elements:
            if all(p[0] == "scalar" for p in el["props"]):
                dt = np.dtype([(p[2], endian + PLY_TYPES[p[1]][0]) for p in el["props"]])
                arr = np.frombuffer(data, dtype=dt, count=el["count"], offset=pos)
                pos += dt.itemsize * el["count"]
                if el["name"] == "vertex":
                    verts = np.stack([arr["x"], arr["y"], arr["z"]], axis=1).astype(np.float32)
            else:
                for _ in range(el["count"]):
                    row = {}
                    for p in el["props"]:
                        if p[0] == "scalar":
                            c, sz = PLY_TYPES[p[1]]
                            row[p[2]] = struct.unpack_from(endian + c, data, pos)[0]; pos += sz
                        else:
                            c, sz = PLY_TYPES[p[1]]
                            n = struct.unpack_from(endian + c, data, pos)[0]; pos += sz
                            c2, sz2 = PLY_TYPES[p[2]]
                            row[p[3]] = struct.unpack_from(endian + str(n) + c2, data, pos); pos += sz2 * n
                    if el["name"] == "face":
                        idx = row.get("vertex_indices", row.get("vertex_index"))
                        if idx is None:
                            continue
                        if len(idx) == 3:
                            tris.append(tuple(idx))
                        elif len(idx) == 4:
                            tris.append((idx[0], idx[1], idx[2]))
                            tris.append((idx[0], idx[2], idx[3]))
    return verts, np.asarray(tris, dtype=np.uint32).reshape(-1, 3)


def parse_props(path):
    props = {}
    order = []
    with open(path, "r", errors="replace") as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            if "=" not in line:
                continue
            k, v = line.split("=", 1)
            k = k.strip()
            try:
                vals = shlex.split(v.strip())
            except ValueError:
                vals = v.strip().split()
            if k not in props:
                order.append(k)
            props[k] = vals
    return props, order


def mat_from_prop(vals):
    v = np.asarray([float(x) for x in vals], dtype=np.float32)
    assert v.size == 16
    # column-major in the file: m[row][col] = v[col * 4 + row]
    return v.reshape(4, 4).T.copy()


def convert(scene_name, scn_rel, scene_dir_relative_ply=False):
    scn = os.path.join(REF, scn_rel)
    props, order = parse_props(scn)
    objs = []
    for k in order:
        m = re.match(r"scene\.objects\.([^.]+)\.", k)
        if m and m.group(1) not in objs:
            objs.append(m.group(1))
    shapes = {}
    shape_list = []
    verts_all, tris_all = [], []
    shape_vert_off, shape_tri_off = [0], [0]
    obj_shape, obj_kind, obj_xform = [], [], []
    motion_obj, motion_time, motion_xform = [], [], []
    skipped = []
    for name in objs:
        pre = "scene.objects." + name
        # `.ply = <file>` (old syntax), or `.shape = <name>` / `.ply = <name>` naming a
        # `scene.shapes.<name>.type = mesh` definition (parseobjects.cpp:103-134, parseshapes.cpp)
        ref = props.get(pre + ".ply", props.get(pre + ".shape", [None]))[0]
        if ref is None:
            skipped.append(name + " (no shape)")
            continue
        if "scene.shapes." + ref + ".ply" in props:
            if props.get("scene.shapes." + ref + ".type", ["mesh"])[0] != "mesh":
                skipped.append(name + " (procedural shape)")
                continue
            ref = props["scene.shapes." + ref + ".ply"][0]
        ply = ref
        path = os.path.join(os.path.dirname(scn), ply) if scene_dir_relative_ply else os.path.join(REF, ply)
        if ply not in shapes:
            if not os.path.exists(path):
                skipped.append(name + " (" + ply + " missing)")
                continue
            v, t = read_ply(path)
            assert t.max() < v.shape[0], path
            shapes[ply] = len(shape_list)
            shape_list.append(ply)
            verts_all.append(v)
            tris_all.append(t)
            shape_vert_off.append(shape_vert_off[-1] + v.shape[0])
            shape_tri_off.append(shape_tri_off[-1] + t.shape[0])
        si = shapes[ply]
        oi = len(obj_shape)
        obj_shape.append(si)
        if pre + ".motion.0.time" in props:
            obj_kind.append(2)
            obj_xform.append(np.eye(4, dtype=np.float32))
            i = 0
            while pre + ".motion.%d.time" % i in props:
                motion_obj.append(oi)
                motion_time.append(float(props[pre + ".motion.%d.time" % i][0]))
                key = pre + ".motion.%d.transformation" % i
                motion_xform.append(mat_from_prop(props[key]) if key in props else np.eye(4, dtype=np.float32))
                i += 1
        elif pre + ".transformation" in props:
            obj_kind.append(1)
            obj_xform.append(mat_from_prop(props[pre + ".transformation"]))
        else:
            obj_kind.append(0)
            obj_xform.append(np.eye(4, dtype=np.float32))

    cam = np.zeros(10, dtype=np.float32)
    if "scene.camera.lookat" in props:
        cam[0:6] = [float(x) for x in props["scene.camera.lookat"]]
    else:
        cam[0:3] = [float(x) for x in props.get("scene.camera.lookat.orig", ["0", "10", "0"])]
        cam[3:6] = [float(x) for x in props.get("scene.camera.lookat.target", ["0", "0", "0"])]
    cam[6:9] = [float(x) for x in props.get("scene.camera.up", ["0", "0", "1"])]     # parsecamera.cpp:70
    cam[9] = float(props.get("scene.camera.fieldofview", ["45"])[0])

    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, scene_name + ".npz")
    np.savez_compressed(
        out,
        verts=np.concatenate(verts_all).astype(np.float32),
        tris=np.concatenate(tris_all).astype(np.uint32),
        shape_vert_off=np.asarray(shape_vert_off, dtype=np.int64),
        shape_tri_off=np.asarray(shape_tri_off, dtype=np.int64),
        obj_shape=np.asarray(obj_shape, dtype=np.int32),
        obj_kind=np.asarray(obj_kind, dtype=np.int8),
        obj_xform=np.asarray(obj_xform, dtype=np.float32).reshape(-1, 4, 4),
        motion_obj=np.asarray(motion_obj, dtype=np.int32),
        motion_time=np.asarray(motion_time, dtype=np.float32),
        motion_xform=np.asarray(motion_xform, dtype=np.float32).reshape(-1, 4, 4),
        cam=cam,
        skipped=np.asarray(skipped, dtype=object).astype(str),
        source=np.asarray([scn_rel]),
    )
    ntri = sum(int(shape_tri_off[s + 1] - shape_tri_off[s]) for s in obj_shape)
    print("%-22s objects %5d  shapes %3d  tris(unique) %8d  tris(instanced) %9d  skipped %s  -> %s (%.1f KB)" % (
        scene_name, len(obj_shape), len(shape_list), shape_tri_off[-1], ntri, skipped, out, os.path.getsize(out) / 1024.0))


SCENES = [
    ("cornell", "scenes/cornell/cornell.scn", False),
    ("kitchen", "scenes/kitchen/kitchen.scn", False),
    ("classroom", "scenes/classroom/classroom.scn", False),
    ("bigmonkey", "scenes/bigmonkey/bigmonkey.scn", False),
    ("bigmonkey-motion", "scenes/bigmonkey/bigmonkey-motion.scn", False),
    ("bigmonkey-instances", "scenes/bigmonkey/bigmonkey-instances.scn", False),
    ("luxball", "scenes/luxball/luxball-hdr.scn", False),
    ("lightinstances", "scenes/lightinstances/scene.scn", True),
]

if __name__ == "__main__":
    want = sys.argv[1:]
    for name, rel, local in SCENES:
        if want and name not in want:
            continue
        convert(name, rel, local)
