"""Materials and textures (reference: material/material.go:44-86, material/pool.go:15-29,
buffer/texture.go:15-69)."""
from __future__ import annotations

import numpy as np

from . import imageutil

f32 = np.float32


def color_from_value(r, g, b, a):
    """color.FromValue (color/color.go:33-44): uint8(Round(v*255)), half away from zero."""
    def q(v):
        x = float(f32(v) * f32(255))
        return int(np.floor(abs(x) + 0.5)) if x >= 0 else -int(np.floor(abs(x) + 0.5))
    return (q(r), q(g), q(b), q(a))


class Texture:
    """buffer.Texture (buffer/texture.go:28-69): RGBA8 image + mip chain built with imageutil.Resize."""

    def __init__(self, image: np.ndarray | None = None, use_mipmap: bool = True, mipmap: list | None = None):
        if image is None:
            image = np.full((1, 1, 4), 255, dtype=np.uint8)
        self.image = np.ascontiguousarray(image, dtype=np.uint8)
        self.use_mipmap = use_mipmap
        self.mipmap = mipmap if mipmap is not None else imageutil.build_mipmap(self.image)

    @staticmethod
    def uniform(rgba):
        """buffer.NewUniformTexture (buffer/texture.go:15-23)."""
        return Texture(np.array(rgba, dtype=np.uint8).reshape(1, 1, 4))

    def Size(self):
        return self.image.shape[1]


class BlinnPhong:
    """material.NewBlinnPhong (material/material.go:54-70)."""

    def __init__(self, texture: Texture | None = None, diffuse=None, specular=None, shininess=1,
                 flat_shading=False, ambient_occlusion=False, receive_shadow=False, name=""):
        self.texture = texture
        self.diffuse = tuple(diffuse) if diffuse is not None else color_from_value(0.5, 0.5, 0.5, 1.0)
        self.specular = tuple(specular) if specular is not None else color_from_value(0.5, 0.5, 0.5, 1.0)
        self.shininess = f32(shininess)
        self.flat_shading, self.ambient_occlusion, self.receive_shadow = flat_shading, ambient_occlusion, receive_shadow
        self.name = name


_default = None


def Default():
    """material.Default (material/pool.go:15-29): blue 1x1 texture, Kd .7, Ks .5, shininess 30."""
    global _default
    if _default is None:
        _default = BlinnPhong(texture=Texture.uniform((0, 0, 255, 255)), diffuse=color_from_value(0.7, 0.7, 0.7, 1.0),
                              specular=color_from_value(0.5, 0.5, 0.5, 1.0), shininess=30.0, name="default")
    return _default
