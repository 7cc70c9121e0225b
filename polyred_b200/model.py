"""Wavefront .obj/.mtl ingest (reference: model/load.go:40-260, model/obj/obj.go). Input-side
only: produces the Geometry arrays that prc_scene_upload consumes."""
from __future__ import annotations

import os

import numpy as np

from . import imageutil
from .material import BlinnPhong, Default, Texture, color_from_value
from .scene import Geometry, Group

f32 = np.float32


def _parse_mtl(path):
    mats, cur = {}, None
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if not f or f[0].startswith("#"):
                continue
            k = f[0]
            if k == "newmtl":
                cur = mats.setdefault(f[1], {"name": f[1], "illum": 0, "Ns": f32(0), "Kd": (0, 0, 0, 0), "Ks": (0, 0, 0, 0), "map_Kd": ""})
            elif cur is None:
                continue
            elif k == "Kd":
                cur["Kd"] = color_from_value(f32(f[1]), f32(f[2]), f32(f[3]), 1)
            elif k == "Ks":
                cur["Ks"] = color_from_value(f32(f[1]), f32(f[2]), f32(f[3]), 1)
            elif k == "Ns":
                cur["Ns"] = f32(f[1])
            elif k == "illum":
                cur["illum"] = int(f[1])
            elif k == "map_Kd":
                cur["map_Kd"] = f[1]
    return mats


def Load(path: str) -> Group:
    """model.Load for .obj (model/load.go:40-151). One Geometry per `o`/`g` object, materials
    local to the geometry in first-use order, vertex colour opaque white."""
    V, N, T = [], [], []
    objs, cur, curmat, mtllib = [], None, None, ""
    used = {}
    lineno = 0
    with open(path) as fh:
        for line in fh:
            lineno += 1
            f = line.split()
            if not f or f[0].startswith("#"):
                continue
            k = f[0]
            if k == "mtllib":
                mtllib = f[1]
            elif k in ("o", "g"):
                cur = {"name": f[1], "faces": []}
                objs.append(cur)
            elif k == "v":
                w = f32(1)
                if len(f) - 1 == 4:
                    w = f32(1) / f32(f[4])
                V.append([f32(x) if w == 1 else f32(x) * w for x in f[1:4]])
            elif k == "vn":
                N.append([f32(x) for x in f[1:4]])
            elif k == "vt":
                T.append([f32(x) for x in f[1:3]])
            elif k == "usemtl":
                if cur is None:
                    cur = {"name": f"unnamed{lineno}", "faces": []}
                    objs.append(cur)
                curmat = f[1]
                used.setdefault(curmat, None)
            elif k == "f":
                if cur is None:
                    cur = {"name": f"unnamed{lineno}", "faces": []}
                    objs.append(cur)
                vi, ti, ni = [], [], []
                for part in f[1:]:
                    p = part.split("/")
                    a = int(p[0])
                    vi.append(a - 1 if a > 0 else len(V) + a)
                    if len(p) > 1 and p[1]:
                        b = int(p[1])
                        ti.append(b - 1 if b > 0 else len(T) + b)
                    else:
                        ti.append(-1)
                    if len(p) >= 3:
                        c = int(p[2])
                        ni.append(c - 1 if c > 0 else len(N) + c)
                    else:
                        ni.append(-1)
                cur["faces"].append((vi, ti, ni, curmat if curmat is not None else "polyred_default"))
    V = np.array(V, np.float32).reshape(-1, 3)
    N = np.array(N, np.float32).reshape(-1, 3)
    T = np.array(T, np.float32).reshape(-1, 2)

    mtl = {}
    if mtllib:
        mtl = _parse_mtl(os.path.join(os.path.dirname(path), mtllib))
    all_mats = {}
    for name, m in mtl.items():
        if m["illum"] not in (0, 2):
            raise ValueError("unsupported illumination model")
        if m["map_Kd"] == "":
            tex = Texture.uniform((0, 0, 255, 255))
        else:
            tex = Texture(imageutil.load_image(os.path.join(os.path.dirname(path), m["map_Kd"]), gamma_correct=True), use_mipmap=True)
        all_mats[name] = BlinnPhong(texture=tex, diffuse=m["Kd"], specular=m["Ks"], shininess=m["Ns"], name=name)

    g = Group()
    for ob in objs:
        geom_mats, local_of = [], {}

        def local_index(m):
            if id(m) not in local_of:
                local_of[id(m)] = len(geom_mats)
                geom_mats.append(m)
            return local_of[id(m)]

        tris, quads, polys = [], [], []
        for vi, ti, ni, mname in ob["faces"]:
            m = all_mats.get(mname) or Default()
            mid = local_index(m)
            (tris if len(vi) == 3 else quads if len(vi) == 4 else polys).append((vi, ti, ni, mid))
        if not ob["faces"]:
            continue

        def vert(vi, ti, ni, k):
            p = V[vi[k]]
            n = N[ni[k]] if len(N) > 0 else np.zeros(3, np.float32)
            t = T[ti[k]] if len(T) > 0 else np.zeros(2, np.float32)
            return p, n, t

        P, Nn, U, M = [], [], [], []

        def emit(face, order, mid):
            vi, ti, ni, _ = face
            vs = [vert(vi, ti, ni, k) for k in order]
            P.append([v[0] for v in vs]); Nn.append([v[1] for v in vs]); U.append([v[2] for v in vs]); M.append(mid)

        if tris and not quads:      # TriangleMesh (model/load.go:124-125)
            for fc in tris:
                emit(fc, (0, 1, 2), fc[3])
            # newTrianglePrimitive (load.go:171-179): zero normals replaced by the face normal — and by
            # a typo zero V2/V3 normals overwrite V1. Not reproduced: no reference fixture has zero normals.
        elif quads and not tris:    # QuadMesh: V1V2V3, V1V3V4 (geometry/primitive/quad.go:55-63)
            for fc in quads:
                emit(fc, (0, 1, 2), fc[3]); emit(fc, (0, 2, 3), fc[3])
        else:                        # PolygonMesh: NewPolygon resets MaterialID to -1 (primitive/polygon.go:21-25)
            for fc in tris + quads + polys:
                n = len(fc[0])
                for i in range(n - 2):
                    emit(fc, (0, i + 1, i + 2), -1)
        geo = Geometry(np.array(P, np.float32), np.array(Nn, np.float32), np.array(U, np.float32), None, np.array(M, np.int32), geom_mats)
        g.Add(geo)
    return g
