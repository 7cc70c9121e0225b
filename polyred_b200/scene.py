"""Scene graph on the input side of the render pass (reference: scene/core.go:44-219,
scene/scene.go:35-52, scene/group.go:21-44, geometry/geometry.go:26-59).

Only what `Render()` consumes is mirrored: traversal order (= draw order), the model-matrix
chain handed to the iterator, `Lights()` and `Center()`.
"""
from __future__ import annotations

import numpy as np

from . import gomath as gm
from .light import Ambient, Directional, Point

f32 = np.float32
_FMAX = np.finfo(np.float32).max


class _Tracked(list):
    """A Group's object list: every mutation bumps the scene-graph epoch (gm.EPOCH), so Scene.signature() can tell in O(1)
    that nothing changed since it last walked the graph."""


def _bumping(name):
    base = getattr(list, name)

    def f(self, *a, **k):
        gm.touch()
        return base(self, *a, **k)
    f.__name__ = name
    return f


for _n in ("append", "extend", "insert", "remove", "pop", "clear", "sort", "reverse", "__setitem__", "__delitem__", "__iadd__", "__imul__"):
    setattr(_Tracked, _n, _bumping(_n))


class Geometry(gm.TransformContext):
    """geometry.Geometry: a triangle soup + the materials it owns (geometry/geometry.go:26-59).

    pos/nor: [n,3,3] float32 (Pos.W=1, Nor.W=0), uv: [n,3,2], col: [n,3] packed RGBA8,
    mat: [n] geometry-LOCAL material index (negative = vertex colour)."""

    def __init__(self, pos, nor=None, uv=None, col=None, mat=None, materials=()):
        super().__init__()
        self.pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3, 3)
        n = self.pos.shape[0]
        self.nor = np.zeros((n, 3, 3), np.float32) if nor is None else np.ascontiguousarray(nor, dtype=np.float32).reshape(n, 3, 3)
        self.uv = np.zeros((n, 3, 2), np.float32) if uv is None else np.ascontiguousarray(uv, dtype=np.float32).reshape(n, 3, 2)
        self.col = np.full((n, 3), 0xFFFFFFFF, np.uint32) if col is None else np.ascontiguousarray(col, dtype=np.uint32).reshape(n, 3)
        self.mat = np.zeros(n, np.int32) if mat is None else np.ascontiguousarray(mat, dtype=np.int32).reshape(n)
        self.materials = list(materials)

    def __setattr__(self, name, value):
        if name in ("pos", "nor", "uv", "col", "mat", "materials"):  # a new array object = new geometry for a renderer's cache
            gm.touch()
        object.__setattr__(self, name, value)

    def Triangles(self):
        return self.pos

    def Materials(self):
        return self.materials

    def aabb(self):
        """mesh AABB in MODEL space (geometry/mesh/mesh_triangle.go:44-48)."""
        c = getattr(self, "_aabb_cache", None)
        if c is None or c[0] is not self.pos:  # the vertex array is immutable once attached; cached per array object
            p = self.pos.reshape(-1, 3)
            c = (self.pos, p.min(axis=0).astype(np.float32), p.max(axis=0).astype(np.float32))
            self._aabb_cache = c
        return c[1].copy(), c[2].copy()


class Group(gm.TransformContext):
    """scene.Group (scene/core.go:114-219)."""

    def __init__(self, *objects):
        super().__init__()
        self.objects = _Tracked(objects)

    def __setattr__(self, name, value):
        if name == "objects":
            gm.touch()
            if not isinstance(value, _Tracked):
                value = _Tracked(value)
        object.__setattr__(self, name, value)

    def Add(self, *objects):
        self.objects.extend(objects)
        return self

    def _iter(self, fn):
        """Group.iterObjects (scene/core.go:197-219): leaf objects get THIS group's model
        matrix; nested groups chain g.ModelMatrix().MulM(nested)."""
        for o in self.objects:
            if isinstance(o, Group):
                o._iter(lambda obj, m, g=self: fn(obj, gm.mulm(g.ModelMatrix(), m)))
            else:
                fn(o, self.ModelMatrix())

    def leaves(self):
        out = []
        self._iter(lambda o, m: out.append(o))
        return out

    def aabb(self):
        """Group.AABB (scene/group.go:21-44): union of the leaves' model-space AABBs."""
        mn = mx = None
        for o in self.leaves():
            a, b = o.aabb()
            if mn is None:
                mn, mx = a.copy(), b.copy()
            else:
                mn, mx = np.minimum(mn, a), np.maximum(mx, b)
        if mn is None:
            return gm.v3(0, 0, 0), gm.v3(0, 0, 0)
        return mn, mx

    def Normalize(self):
        """Group.Normalize (scene/group.go:47-64)."""
        m = self.ModelMatrix()
        a, b = self.aabb()
        mn = gm.mulv(m, np.append(a, f32(1)))[:3]
        mx = gm.mulv(m, np.append(b, f32(1)))[:3]
        center = ((mn + mx).astype(np.float32) * f32(0.5)).astype(np.float32)
        radius = gm.v3_len((mx - mn).astype(np.float32)) / f32(2)
        fac = f32(1) / radius
        self.Translate(-center[0], -center[1], -center[2])
        self.Scale(fac, fac, fac)


class Scene:
    """scene.Scene (scene/core.go:30-111)."""

    def __init__(self, *objects):
        self.root = Group()
        self.Add(*objects)

    def Add(self, *objects):
        self.root.Add(*objects)
        return self.root

    def IterObjects(self, fn):
        """Scene.IterObjects (scene/core.go:86-111): root-level leaves get root.ModelMatrix();
        groups get root.ModelMatrix().MulM(chain)."""
        r = self.root
        for o in r.objects:
            if isinstance(o, Group):
                o._iter(lambda obj, m: fn(obj, gm.mulm(r.ModelMatrix(), m)))
            else:
                fn(o, r.ModelMatrix())

    def geometries(self):
        """[(Geometry, modelMatrix)] in draw order; modelMatrix is what the iterator hands to
        cpuForwardPass, BEFORE .MulM(g.ModelMatrix()) (render/raster.go:241-242)."""
        out = []
        self.IterObjects(lambda o, m: out.append((o, m)) if isinstance(o, Geometry) else None)
        return out

    def Lights(self):
        """Scene.Lights (scene/scene.go:35-47): (sources, environments) in traversal order."""
        src, env = [], []
        def visit(o, m):
            if isinstance(o, (Point, Directional)):
                src.append(o)
            elif isinstance(o, Ambient):
                env.append(o)
        self.IterObjects(visit)
        return src, env

    def signature(self):
        """(membership, transforms): what a renderer caches against. `membership` changes when an object is
        added to / removed from the graph or a geometry gets new vertex arrays (=> re-flatten and re-upload); `transforms`
        changes when any TransformContext on the way to a leaf was scaled / translated / rotated (=> only the per-object
        matrices are recomputed; the triangle soup stays resident)."""
        cached = getattr(self, "_sig_cache", None)
        if cached is not None and cached[0] == gm.EPOCH[0]:
            return cached[1]  # nothing in ANY scene graph was mutated since the last walk
        epoch = gm.EPOCH[0]
        ids, ver = [], 0

        def walk(g):
            nonlocal ver
            ver += g._version
            ids.append(id(g))
            for o in g.objects:
                if isinstance(o, Group):
                    walk(o)
                else:
                    ids.append(id(o))
                    ver += getattr(o, "_version", 0)
                    if isinstance(o, Geometry):
                        ids.append(id(o.pos))
        walk(self.root)
        sig = (hash(tuple(ids)), ver)
        self._sig_cache = (epoch, sig)
        return sig

    def Center(self):
        """Scene.Center (scene/scene.go:49-52): centre of the root AABB — model-space AABBs of
        ALL root objects, lights included."""
        a, b = self.root.aabb()
        return ((a + b).astype(np.float32) * f32(0.5)).astype(np.float32)
