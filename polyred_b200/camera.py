"""Cameras: host-side producers of the View / Proj matrices (reference: camera/camera.go:42-55,
camera/perspective.go:100-111, camera/orthographic.go:106-119)."""
from __future__ import annotations

import numpy as np

from . import gomath as gm

f32 = np.float32


def ViewMatrix(pos, target, up):
    """camera.ViewMatrix (camera/camera.go:42-55)."""
    pos, target, up = gm._a(pos), gm._a(target), gm._a(up)
    l = gm.v3_unit((target - pos).astype(np.float32))
    lxu = gm.v3_unit(gm.v3_cross(l, up))
    u = gm.v3_unit(gm.v3_cross(lxu, l))
    return gm.mat4(
        lxu[0], lxu[1], lxu[2], -gm.v3_dot(lxu, pos),
        u[0], u[1], u[2], -gm.v3_dot(u, pos),
        -l[0], -l[1], -l[2], gm.v3_dot(l, pos),
        0, 0, 0, 1,
    )


class _Camera:
    def Position(self):
        return self.position

    def ViewMatrix(self):
        return ViewMatrix(self.position, self.target, self.up)


class Perspective(_Camera):
    """camera.NewPerspective (camera/perspective.go:31-47); ViewFrustum = (fov, aspect, near, far)."""

    perspect = True

    def __init__(self, position=(0, 0, 1), target=(0, 0, 0), up=(0, 1, 0), fov=60, aspect=16.0 / 9, near=0.01, far=1000):
        self.position, self.target, self.up = gm._a(position), gm._a(target), gm._a(up)
        self.fov, self.aspect, self.near, self.far = f32(fov), f32(aspect), f32(near), f32(far)

    def ProjMatrix(self):
        """Perspective.ProjMatrix (camera/perspective.go:100-111): w_clip = z_view, +1 = near."""
        fov = (self.fov * gm.PI32) / f32(180)
        n, f = self.near, self.far
        t = gm.tan(fov / f32(2))
        return gm.mat4(
            f32(-1) / (self.aspect * t), 0, 0, 0,
            0, f32(-1) / t, 0, 0,
            0, 0, (n + f) / (n - f), (f32(2) * n * f) / (n - f),
            0, 0, 1, 0,
        )


class Orthographic(_Camera):
    """camera.NewOrthographic (camera/orthographic.go:32-50); ViewFrustum = (l, r, b, t, near, far)."""

    perspect = False

    def __init__(self, position=(0, 0, 1), target=(0, 0, 0), up=(0, 1, 0), left=-1, right=1, bottom=-1, top=1, near=1, far=-1):
        self.position, self.target, self.up = gm._a(position), gm._a(target), gm._a(up)
        self.left, self.right, self.bottom, self.top, self.near, self.far = (f32(x) for x in (left, right, bottom, top, near, far))

    def ProjMatrix(self):
        """Orthographic.ProjMatrix (camera/orthographic.go:106-119)."""
        l, r, t, b, n, f = self.left, self.right, self.top, self.bottom, self.near, self.far
        two = f32(2)
        return gm.mat4(
            two / (r - l), 0, 0, (l + r) / (l - r),
            0, two / (t - b), 0, (b + t) / (b - t),
            0, 0, two / (n - f), (f + n) / (f - n),
            0, 0, 0, 1,
        )
