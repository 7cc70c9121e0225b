"""Host-side float32 math with the exact rounding behaviour of poly.red/math.

The render pass consumes matrices that the reference computes on the host with plain
float32 arithmetic (``Mat4.MulM/Inv/Det/T``: unfused, left to right — math/mat4.go:201-283)
and with ``math.FMA`` = float64 fma rounded to float32 (``Vec3.Dot/Cross`` —
math/vec3.go:78-120). This module restates those few functions on numpy float32 so the
uniforms handed to ``prc_render`` are bit-identical to what the Go host would compute.

All functions broadcast over leading dimensions (one call builds the uniforms of every
object of a scene); numpy evaluates ``a*b + c*d`` as ``(a*b) + (c*d)`` in float32, the same
order and rounding as Go on amd64.

Reference: math/mat4.go, math/vec3.go, math/math.go, math/quaternion.go, math/context.go.
"""
from __future__ import annotations

import math as _m

import numpy as np

f32 = np.float32
PI32 = f32(_m.pi)  # float32(math.Pi) (math/math.go:16)


def _a(x):
    return np.asarray(x, dtype=np.float32)


# ----------------------------------------------------------------------------- scalars
def fma(x, y, z):
    """math.FMA[float32] (math/math.go:265-267): float64 fma, rounded to float32.

    The product of two float32 is exact in float64, so x*y+z in float64 has ONE rounding."""
    return (np.asarray(x, np.float64) * np.asarray(y, np.float64) + np.asarray(z, np.float64)).astype(np.float32)


def sqrt(x):
    """math.Sqrt (math/math.go:231-233)."""
    return np.sqrt(np.asarray(x, np.float64)).astype(np.float32)


def tan(x):
    """math.Tan (math/math.go:176-178) — float64 tan rounded to float32."""
    return f32(_m.tan(float(x)))


# ----------------------------------------------------------------------------- Vec3
def v3(x, y, z):
    return np.array([x, y, z], dtype=np.float32)


def v3_dot(v, u):
    """Vec3.Dot (math/vec3.go:78-82)."""
    v, u = _a(v), _a(u)
    return fma(v[..., 0], u[..., 0], fma(v[..., 1], u[..., 1], v[..., 2] * u[..., 2]))


def v3_len(v):
    return sqrt(v3_dot(v, v))


def v3_unit(v):
    """Vec3.Unit (math/vec3.go:90-93): n = 1/len, components * n."""
    v = _a(v)
    n = f32(1.0) / v3_len(v)
    return (v * n[..., None]).astype(np.float32) if np.ndim(n) else (v * n).astype(np.float32)


def v3_cross(v, u):
    """Vec3.Cross (math/vec3.go:113-120)."""
    v, u = _a(v), _a(u)
    x = fma(v[..., 1], u[..., 2], -v[..., 2] * u[..., 1])
    y = fma(v[..., 2], u[..., 0], -v[..., 0] * u[..., 2])
    z = fma(v[..., 0], u[..., 1], -v[..., 1] * u[..., 0])
    return np.stack([x, y, z], axis=-1).astype(np.float32)


# ----------------------------------------------------------------------------- Vec4
def v4_apply(v, m):
    """Vec4.Apply (math/vec4.go:108-116): nested FMA per row."""
    v, m = _a(v), _a(m)
    rows = []
    for r in range(4):
        rows.append(fma(m[..., r, 0], v[..., 0], fma(m[..., r, 1], v[..., 1], fma(m[..., r, 2], v[..., 2], m[..., r, 3] * v[..., 3]))))
    return np.stack(rows, axis=-1).astype(np.float32)


def v4_pos(v):
    """Vec4.Pos (math/vec4.go:140-146)."""
    v = _a(v)
    out = v.copy()
    w = v[..., 3]
    if np.ndim(w) == 0:
        if w == 1 or w == 0:
            out[3] = 1
            return out
        inv = f32(1.0) / w
        return np.array([v[0] * inv, v[1] * inv, v[2] * inv, 1], dtype=np.float32)
    raise NotImplementedError


# ----------------------------------------------------------------------------- Mat4
def mat4(*vals):
    return np.array(vals, dtype=np.float32).reshape(4, 4)


def identity():
    return np.eye(4, dtype=np.float32)


def mulm(m, n):
    """Mat4.MulM (math/mat4.go:201-220): each element ((a*b + c*d) + e*f) + g*h in float32."""
    m, n = _a(m), _a(n)
    if m.ndim == 2 and n.ndim == 2:
        # one matrix: all 16 elements at once (element [i, j] of every numpy operation is exactly that element's scalar operation)
        return ((m[:, 0, None] * n[None, 0, :] + m[:, 1, None] * n[None, 1, :]) + m[:, 2, None] * n[None, 2, :]) + m[:, 3, None] * n[None, 3, :]
    shape = np.broadcast_shapes(m.shape, n.shape)
    out = np.empty(shape, dtype=np.float32)
    for i in range(4):  # batches: 16 passes over contiguous-ish columns beat 7 broadcasts over 4x4 inner blocks
        for j in range(4):
            out[..., i, j] = m[..., i, 0] * n[..., 0, j] + m[..., i, 1] * n[..., 1, j] + m[..., i, 2] * n[..., 2, j] + m[..., i, 3] * n[..., 3, j]
    return out


def mulv(m, v):
    """Mat4.MulV (math/mat4.go:224-230)."""
    m, v = _a(m), _a(v)
    rows = [m[..., r, 0] * v[..., 0] + m[..., r, 1] * v[..., 1] + m[..., r, 2] * v[..., 2] + m[..., r, 3] * v[..., 3] for r in range(4)]
    return np.stack(rows, axis=-1).astype(np.float32)


def transpose(m):
    """Mat4.T (math/mat4.go:250-257)."""
    return np.swapaxes(_a(m), -1, -2).copy()


# The cofactor expansions below are written as term tables: each term is a signed product
# of matrix elements "rc" (row, column) multiplied left to right, and the terms are summed
# left to right — the order fixes the float32 rounding (math/mat4.go:233-283).
_DET_TERMS = (
    "+00.11.22.33 -00.11.23.32 +00.12.23.31 -00.12.21.33 +00.13.21.32 -00.13.22.31 "
    "-01.12.23.30 +01.12.20.33 -01.13.20.32 +01.13.22.30 -01.10.22.33 +01.10.23.32 "
    "+02.13.20.31 -02.13.21.30 +02.10.21.33 -02.10.23.31 +02.11.23.30 -02.11.20.33 "
    "-03.10.21.32 +03.10.22.31 -03.11.22.30 +03.11.20.32 -03.12.20.31 +03.12.21.30"
)
_INV_TERMS = {
    (0, 0): "+12.23.31 -13.22.31 +13.21.32 -11.23.32 -12.21.33 +11.22.33",
    (0, 1): "+03.22.31 -02.23.31 -03.21.32 +01.23.32 +02.21.33 -01.22.33",
    (0, 2): "+02.13.31 -03.12.31 +03.11.32 -01.13.32 -02.11.33 +01.12.33",
    (0, 3): "+03.12.21 -02.13.21 -03.11.22 +01.13.22 +02.11.23 -01.12.23",
    (1, 0): "+13.22.30 -12.23.30 -13.20.32 +10.23.32 +12.20.33 -10.22.33",
    (1, 1): "+02.23.30 -03.22.30 +03.20.32 -00.23.32 -02.20.33 +00.22.33",
    (1, 2): "+03.12.30 -02.13.30 -03.10.32 +00.13.32 +02.10.33 -00.12.33",
    (1, 3): "+02.13.20 -03.12.20 +03.10.22 -00.13.22 -02.10.23 +00.12.23",
    (2, 0): "+11.23.30 -13.21.30 +13.20.31 -10.23.31 -11.20.33 +10.21.33",
    (2, 1): "+03.21.30 -01.23.30 -03.20.31 +00.23.31 +01.20.33 -00.21.33",
    (2, 2): "+01.13.30 -03.11.30 +03.10.31 -00.13.31 -01.10.33 +00.11.33",
    (2, 3): "+03.11.20 -01.13.20 -03.10.21 +00.13.21 +01.10.23 -00.11.23",
    (3, 0): "+12.21.30 -11.22.30 -12.20.31 +10.22.31 +11.20.32 -10.21.32",
    (3, 1): "+01.22.30 -02.21.30 +02.20.31 -00.22.31 -01.20.32 +00.21.32",
    (3, 2): "+02.11.30 -01.12.30 -02.10.31 +00.12.31 +01.10.32 -00.11.32",
    (3, 3): "+01.12.20 -02.11.20 +02.10.21 -00.12.21 -01.10.22 +00.11.22",
}


def _parse(terms):
    """'+00.11.22 -...' -> (signs [T] of +-1, rows [T, F], cols [T, F])"""
    sg, rr, cc = [], [], []
    for t in terms.split():
        sg.append(1.0 if t[0] == "+" else -1.0)
        facs = t[1:].split(".")
        rr.append([int(f[0]) for f in facs])
        cc.append([int(f[1]) for f in facs])
    return np.array(sg, np.float32), np.array(rr), np.array(cc)


_DET_P = _parse(_DET_TERMS)
# the 16 cofactor sums evaluated together: term k of every entry, entries in row-major order -> index arrays [6, 16, 3]
_INV_P = [_parse(_INV_TERMS[(i, j)]) for i in range(4) for j in range(4)]
_INV_SG = np.stack([p[0] for p in _INV_P], axis=1)   # [6, 16]
_INV_R = np.stack([p[1] for p in _INV_P], axis=1)    # [6, 16, 3]
_INV_C = np.stack([p[2] for p in _INV_P], axis=1)


def det(m):
    """Mat4.Det (math/mat4.go:233-247): 24 signed products of four elements, each multiplied left to right, summed in the
    reference's order. (x - p and x + (-p) are the same IEEE operation, so a term's sign is applied to its product.)"""
    m = _a(m)
    sg, rr, cc = _DET_P
    p = m[..., rr[:, 0], cc[:, 0]]
    for f in range(1, 4):
        p = p * m[..., rr[:, f], cc[:, f]]
    p = p * sg  # exact: multiplication by +-1
    acc = p[..., 0]
    for k in range(1, 24):
        acc = acc + p[..., k]
    return acc


def inv(m):
    """Mat4.Inv (math/mat4.go:260-283). Raises ZeroDivisionError where the reference panics."""
    m = _a(m)
    d = det(m)
    if np.any(d == 0):
        raise ZeroDivisionError("zero determinant")
    dinv = f32(1) / d
    acc = None
    for k in range(6):  # term k of all 16 entries
        p = m[..., _INV_R[k, :, 0], _INV_C[k, :, 0]] * m[..., _INV_R[k, :, 1], _INV_C[k, :, 1]] * m[..., _INV_R[k, :, 2], _INV_C[k, :, 2]]
        p = p * _INV_SG[k]
        acc = p if acc is None else acc + p
    out = (np.asarray(dinv)[..., None] * acc).astype(np.float32)
    return out.reshape(m.shape)


def viewport_matrix(w, h):
    """math.ViewportMatrix (math/math.go:270-277)."""
    w, h = f32(w), f32(h)
    return mat4(w / f32(2), 0, 0, w / f32(2), 0, h / f32(2), 0, h / f32(2), 0, 0, 1, 0, 0, 0, 0, 1)


# ----------------------------------------------------------------------------- scene-graph mutation epoch
EPOCH = [0]  # bumped by every mutation of a TransformContext, a Group's object list or a Geometry's arrays


def touch():
    EPOCH[0] += 1


# ----------------------------------------------------------------------------- quaternion / TransformContext
class Quaternion:
    """math.Quaternion (math/quaternion.go)."""

    def __init__(self, a, b, c, d):
        self.A = f32(a)
        self.V = v3(b, c, d)

    def mul(self, p: "Quaternion") -> "Quaternion":
        aa = self.A * p.A - v3_dot(self.V, p.V)
        vv = (p.V * self.A + self.V * p.A).astype(np.float32) + v3_cross(self.V, p.V)
        return Quaternion(aa, vv[0], vv[1], vv[2])

    def to_romat(self):
        w, (x, y, z) = self.A, self.V
        two, one = f32(2), f32(1)
        return mat4(
            one - two * y * y - two * z * z, two * x * y - two * z * w, two * x * z + two * y * w, 0,
            two * x * y + two * z * w, one - two * x * x - two * z * z, two * y * z - two * x * w, 0,
            two * x * z - two * y * w, two * y * z + two * x * w, one - two * x * x - two * y * y, 0,
            0, 0, 0, 1,
        )


class TransformContext:
    """math.TransformContext (math/context.go:18-175): scale/translate accumulate in a matrix,
    rotations in a quaternion; ModelMatrix = internal.MulM(rotation.ToRoMat())."""

    def __init__(self):
        self.ResetContext()

    def _touch(self):
        """Every mutation bumps this context's version (a renderer sees WHICH matrices moved) and the process-wide epoch (a
        renderer sees in O(1) that nothing moved at all, Scene.signature)."""
        self._version = getattr(self, "_version", 0) + 1
        touch()

    def ResetContext(self):
        self._context = identity()
        self._rotation = Quaternion(1, 0, 0, 0)
        self._internal = identity()
        self._need = False
        self._touch()

    def ModelMatrix(self):
        if self._need:
            self._context = mulm(self._internal, self._rotation.to_romat())
            self._need = False
        return self._context

    def Scale(self, sx, sy, sz):
        self._internal = mulm(mat4(sx, 0, 0, 0, 0, sy, 0, 0, 0, 0, sz, 0, 0, 0, 0, 1), self._internal)
        self._need = True
        self._touch()

    def Translate(self, tx, ty, tz):
        self._internal = mulm(mat4(1, 0, 0, tx, 0, 1, 0, ty, 0, 0, 1, tz, 0, 0, 0, 1), self._internal)
        self._need = True
        self._touch()

    def Rotate(self, direction, angle):
        u = v3_unit(direction)
        half = f32(angle) * f32(0.5)
        cosa = f32(_m.cos(float(half)))
        sina = f32(_m.sin(float(half)))
        q = Quaternion(cosa, sina * u[0], sina * u[1], sina * u[2])
        self._rotation = q.mul(self._rotation)
        self._need = True
        self._touch()

    def RotateX(self, angle):
        self.Rotate(v3(1, 0, 0), angle)

    def RotateY(self, angle):
        self.Rotate(v3(0, 1, 0), angle)

    def RotateZ(self, angle):
        self.Rotate(v3(0, 0, 1), angle)
