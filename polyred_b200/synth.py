"""Deterministic synthetic scenes for the BASELINE.json configurations (all seeds fixed).

The reference ships no scene of the benchmark sizes, so the workloads are generated here and
fed identically to the CUDA path, the CPU oracle and the CPU baseline. Shapes follow
SURVEY.md §8(d): a heightfield ground plus many instanced closed meshes, several materials
with mip-mapped procedural textures, point lights on a ring (every other one casting).
"""
from __future__ import annotations

import math

import numpy as np

from . import camera, light, material, scene

f32 = np.float32


def _rng(seed):
    return np.random.default_rng(seed)


def checker_texture(size=256, seed=2, tint=(255, 255, 255)):
    """Checker + hash-noise RGBA8 texture (linear space), with the reference's mip chain."""
    r = _rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    chk = (((x // max(1, size // 8)) + (y // max(1, size // 8))) & 1).astype(np.float32)
    noise = r.random((size, size)).astype(np.float32)
    v = (0.35 + 0.5 * chk + 0.15 * noise)
    img = np.empty((size, size, 4), np.uint8)
    for c in range(3):
        img[..., c] = np.clip(v * tint[c], 0, 255).astype(np.uint8)
    img[..., 3] = 255
    return material.Texture(img, use_mipmap=True)


def quads(t1, t2):
    """A quad mesh as a triangle list: the two triangles (V1,V2,V3), (V1,V3,V4) of each quad are consecutive, the
    order Quad.Triangles emits them in (geometry/primitive/quad.go:55-63)."""
    return np.stack([t1, t2], axis=1).reshape((-1,) + t1.shape[1:]).astype(np.float32)


def sphere_mesh(stacks=50, slices=50, seed=1, bump=0.15):
    """Closed lat-long mesh, 2*stacks*slices triangles (pole triangles are degenerate and are
    rejected by Triangle.IsValid, as in any lat-long .obj). Radius displaced by smooth noise."""
    r = _rng(seed)
    i = np.arange(stacks + 1, dtype=np.float64)
    j = np.arange(slices + 1, dtype=np.float64)
    th = np.pi * i / stacks
    ph = 2 * np.pi * j / slices
    TH, PH = np.meshgrid(th, ph, indexing="ij")
    k = r.integers(1, 4, size=(3, 2))
    a = r.random(3) * 2 * np.pi
    rad = 1.0 + bump * (np.sin(k[0, 0] * TH * 2 + a[0]) * np.cos(k[0, 1] * PH + a[1]) * 0.5 + np.sin(k[1, 0] * PH + a[2]) * np.sin(TH) * 0.5)
    rad[:, -1] = rad[:, 0]
    P = np.stack([rad * np.sin(TH) * np.cos(PH), rad * np.cos(TH), rad * np.sin(TH) * np.sin(PH)], axis=-1)
    N = np.stack([np.sin(TH) * np.cos(PH), np.cos(TH), np.sin(TH) * np.sin(PH)], axis=-1)
    UV = np.stack([PH / (2 * np.pi), 1.0 - TH / np.pi], axis=-1)
    ii, jj = np.meshgrid(np.arange(stacks), np.arange(slices), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)

    def tri(a0, b0, c0):
        return np.stack([a0, b0, c0], axis=1)

    def at(A, di, dj):
        return A[ii + di, jj + dj]

    # counter-clockwise seen from outside
    pos = quads(tri(at(P, 0, 0), at(P, 0, 1), at(P, 1, 1)), tri(at(P, 0, 0), at(P, 1, 1), at(P, 1, 0)))
    nor = quads(tri(at(N, 0, 0), at(N, 0, 1), at(N, 1, 1)), tri(at(N, 0, 0), at(N, 1, 1), at(N, 1, 0)))
    uv = quads(tri(at(UV, 0, 0), at(UV, 0, 1), at(UV, 1, 1)), tri(at(UV, 0, 0), at(UV, 1, 1), at(UV, 1, 0)))
    return pos, nor, uv


def ground_mesh(cells=1000, half=1.0, seed=3, amp=0.02, uv_tiles=1.0):
    """Heightfield of cells x cells quads (2 triangles each) over [-half, half]^2 in xz, y up."""
    r = _rng(seed)
    g = np.linspace(-half, half, cells + 1)
    X, Z = np.meshgrid(g, g, indexing="ij")
    k = r.random(4) * 6 + 2
    Y = amp * (np.sin(k[0] * X + k[1] * Z) * 0.5 + np.sin(k[2] * X - k[3] * Z + 1.0) * 0.5)
    P = np.stack([X, Y, Z], axis=-1)
    dYdx = amp * (np.cos(k[0] * X + k[1] * Z) * 0.5 * k[0] + np.cos(k[2] * X - k[3] * Z + 1.0) * 0.5 * k[2])
    dYdz = amp * (np.cos(k[0] * X + k[1] * Z) * 0.5 * k[1] - np.cos(k[2] * X - k[3] * Z + 1.0) * 0.5 * k[3])
    N = np.stack([-dYdx, np.ones_like(X), -dYdz], axis=-1)
    N /= np.linalg.norm(N, axis=-1, keepdims=True)
    UV = np.stack([(X + half) / (2 * half) * uv_tiles, (Z + half) / (2 * half) * uv_tiles], axis=-1)
    ii, jj = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    ii, jj = ii.reshape(-1), jj.reshape(-1)

    def at(A, di, dj):
        return A[ii + di, jj + dj]

    def tri(a0, b0, c0):
        return np.stack([a0, b0, c0], axis=1)

    # counter-clockwise seen from +y
    pos = quads(tri(at(P, 0, 0), at(P, 0, 1), at(P, 1, 1)), tri(at(P, 0, 0), at(P, 1, 1), at(P, 1, 0)))
    nor = quads(tri(at(N, 0, 0), at(N, 0, 1), at(N, 1, 1)), tri(at(N, 0, 0), at(N, 1, 1), at(N, 1, 0)))
    uv = quads(tri(at(UV, 0, 0), at(UV, 0, 1), at(UV, 1, 1)), tri(at(UV, 0, 0), at(UV, 1, 1), at(UV, 1, 0)))
    return pos, nor, uv


def make_materials(n=8, tex_size=256, receive_shadow=True, ambient_occlusion=False, seed=5):
    r = _rng(seed)
    tints = [(255, 220, 200), (200, 255, 210), (210, 220, 255), (255, 255, 200)]
    texs = [checker_texture(tex_size, seed=20 + t, tint=tints[t % 4]) for t in range(min(4, n))]
    mats = []
    for i in range(n):
        kd = int(150 + r.integers(0, 80))
        ks = int(80 + r.integers(0, 100))
        mats.append(material.BlinnPhong(texture=texs[i % len(texs)], diffuse=(kd, kd, kd, 255), specular=(ks, ks, ks, 255),
                                        shininess=[8, 16, 32, 64][i % 4], receive_shadow=receive_shadow, ambient_occlusion=ambient_occlusion))
    return mats


def city_scene(n_objects=1600, obj_stacks=50, obj_slices=50, ground_cells=1000, n_lights=8, casting_every=2, n_materials=8,
               tex_size=256, seed=4, receive_shadow=True, ambient_occlusion=False, aspect=16.0 / 9.0, cam_angle=0.0, cam_radius=2.5,
               cam_height=1.3, directional=False):
    """The C3/C4/C5 generator: ground heightfield + n_objects instanced bumpy spheres.

    Defaults give 2*1000*1000 + 1600*5000 = 10.0 M triangles, 8 point lights (lights 0,2,4,6 cast)."""
    r = _rng(seed)
    mats = make_materials(n_materials, tex_size, receive_shadow, ambient_occlusion)
    s = scene.Scene()
    lights = []
    for i in range(n_lights):
        a = 2 * math.pi * i / max(1, n_lights)
        lights.append(light.Point(intensity=12, color=(255, 255, 255, 255), position=(1.2 * math.cos(a), 0.8, 1.2 * math.sin(a)),
                                  cast_shadow=(casting_every > 0 and i % casting_every == 0)))
    if directional:
        lights.append(light.Directional(intensity=0.9, direction=(-1, -1, -1)))
    s.Add(*lights)
    s.Add(light.Ambient(intensity=0.3))
    gp, gn, gu = ground_mesh(ground_cells, seed=seed + 10)
    s.Add(scene.Geometry(gp, gn, gu, None, np.zeros(gp.shape[0], np.int32), [mats[0]]))
    if n_objects:
        op, on, ou = sphere_mesh(obj_stacks, obj_slices, seed=seed + 20)
        side = int(math.ceil(math.sqrt(n_objects)))
        cell = 2.0 / side
        for k in range(n_objects):
            gi, gj = divmod(k, side)
            sc = cell * 0.5 * float(np.exp(r.uniform(math.log(0.35), math.log(0.8))))
            cx = -1.0 + (gi + 0.5 + r.uniform(-0.2, 0.2)) * cell
            cz = -1.0 + (gj + 0.5 + r.uniform(-0.2, 0.2)) * cell
            g = scene.Geometry(op, on, ou, None, np.zeros(op.shape[0], np.int32), [mats[(k + 1) % len(mats)]])
            g.Rotate(r.normal(size=3).astype(np.float32), float(r.uniform(0, 2 * math.pi)))
            g.Scale(sc, sc, sc)
            g.Translate(cx, sc * 0.9 + 0.02, cz)
            s.Add(g)
    return s, orbit_camera(cam_angle, cam_radius, cam_height, aspect)


def orbit_camera(cam_angle=0.0, cam_radius=2.5, cam_height=1.3, aspect=16.0 / 9.0):
    """The city_scene camera: on a circle of `cam_radius` around the scene centre at `cam_height`, looking at the origin
    (C5 = 256 of these, angle 2 pi k / 256)."""
    return camera.Perspective(position=(cam_radius * math.sin(cam_angle), cam_height, cam_radius * math.cos(cam_angle)), target=(0, 0, 0),
                              up=(0, 1, 0), fov=45, aspect=aspect, near=0.5, far=6.0)


def mesh_scene(subdiv=187, with_ground=False, shadows=False, ao=False, aspect=1.6):
    """C1/C2: one bunny-scale closed mesh (2*187*187 = 69 938 triangles), textured Blinn-Phong;
    newscene()-style lights (render/raster_test.go:32-53): point I=5 at (-2,2.5,6) + ambient 0.5.
    C2 adds a ground quad, a non-casting directional light, a casting point light, AO (see
    SURVEY bug-list 4 for why the caster is a point light)."""
    mats = make_materials(2, 256, receive_shadow=shadows, ambient_occlusion=ao, seed=7)
    s = scene.Scene()
    if shadows:
        s.Add(light.Directional(intensity=0.9, direction=(-1, -1, -1)), light.Point(intensity=3, position=(4, 4, 2), cast_shadow=True))
    else:
        s.Add(light.Point(intensity=5, position=(-2, 2.5, 6)))
    s.Add(light.Ambient(intensity=0.5))
    p, n, u = sphere_mesh(subdiv, subdiv, seed=1)
    g = scene.Geometry(p, n, u, None, np.zeros(p.shape[0], np.int32), [mats[0]])
    g.Scale(0.5, 0.5, 0.5)
    g.Translate(0, 0.55, 0)
    s.Add(g)
    if with_ground:
        gp, gn, gu = ground_mesh(1, half=2.0, amp=0.0, uv_tiles=1.0)
        s.Add(scene.Geometry(gp, gn, gu, None, np.zeros(gp.shape[0], np.int32), [mats[1]]))
    cam = camera.Perspective(position=(0, 1.0, 2.2), target=(0, 0.4, 0), up=(0, 1, 0), fov=45, aspect=aspect, near=0.1, far=10)
    return s, cam
