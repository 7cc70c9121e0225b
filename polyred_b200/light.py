"""Light sources (reference: light/point.go, light/directional.go, light/ambient.go)."""
from __future__ import annotations

import numpy as np

from . import gomath as gm

f32 = np.float32


class Point:
    """light.NewPoint (light/point.go:36-49)."""

    kind = 0

    def __init__(self, intensity=1, color=(255, 255, 255, 255), position=(1, 1, 1), cast_shadow=False):
        self.intensity, self.color, self.position, self.cast_shadow = f32(intensity), tuple(color), gm._a(position), bool(cast_shadow)

    def Position(self):
        return self.position

    def aabb(self):  # light/point.go:70
        return self.position.copy(), self.position.copy()


class Directional:
    """light.NewDirectional (light/directional.go:38-53); direction is normalised at construction."""

    kind = 1

    def __init__(self, intensity=1, color=(255, 255, 255, 255), direction=(0, -1, 0), position=(0, 0, 0), cast_shadow=False):
        self.intensity, self.color, self.cast_shadow = f32(intensity), tuple(color), bool(cast_shadow)
        self.position = gm._a(position)
        self.direction = gm.v3_unit(gm._a(direction))

    def Position(self):
        return self.position

    def aabb(self):  # light/directional.go:63
        return gm.v3(0, 0, 0), gm.v3(0, 0, 0)


class Ambient:
    """light.NewAmbient (light/ambient.go:37-48)."""

    def __init__(self, intensity=0.1, color=(255, 255, 255, 255)):
        self.intensity, self.color = f32(intensity), tuple(color)

    def aabb(self):  # light/ambient.go:55
        return gm.v3(0, 0, 0), gm.v3(0, 0, 0)
