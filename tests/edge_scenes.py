"""Edge-case scenes for the render pass (SURVEY 8a-9' bug list and the corners of the input domain): the same list is rendered
by the CPU oracle alone (tests/test_oracle_edge_cases.py, invariants) and by the CUDA path against the oracle
(tests/test_gpu_edge_cases.py, bit-exact). Each entry: name -> (scene, camera, width, height, renderer options)."""
from __future__ import annotations

import numpy as np

from polyred_b200 import camera, light, material, scene, synth

f32 = np.float32


def _cam(aspect=1.6, pos=(0, 0.3, 2.5), target=(0, 0, 0), near=0.1, far=10):
    return camera.Perspective(position=pos, target=target, up=(0, 1, 0), fov=45, aspect=aspect, near=near, far=far)


def _tex(size=64, seed=3, mip=True):
    t = synth.checker_texture(size, seed=seed)
    t.use_mipmap = mip
    return t


def _tri_geo(tris, mat=None, uv=None, col=None, nor=None, matid=None):
    tris = np.asarray(tris, np.float32).reshape(-1, 3, 3)
    n = tris.shape[0]
    if nor is None:
        e1, e2 = tris[:, 1] - tris[:, 0], tris[:, 2] - tris[:, 0]
        fn = np.cross(e1, e2)
        ln = np.linalg.norm(fn, axis=1, keepdims=True)
        fn = np.where(ln > 0, fn / np.where(ln > 0, ln, 1), np.array([0, 0, 1], np.float32))
        nor = np.repeat(fn[:, None, :], 3, axis=1).astype(np.float32)
    if uv is None:
        uv = np.tile(np.array([[0, 0], [1, 0], [0.5, 1]], np.float32), (n, 1, 1))
    mats = [mat] if mat is not None else []
    if matid is None:
        matid = np.zeros(n, np.int32) if mat is not None else np.full(n, -1, np.int32)
    return scene.Geometry(tris, nor, uv, col, matid, mats)


def _sphere(mat, subdiv=24, scale=0.6, move=(0, 0, 0)):
    p, n, u = synth.sphere_mesh(subdiv, subdiv, seed=1)
    g = scene.Geometry(p, n, u, None, np.zeros(p.shape[0], np.int32), [mat])
    g.Scale(scale, scale, scale)
    g.Translate(*move)
    return g


def _lights(s, shadows=False):
    s.Add(light.Point(intensity=4, position=(-2, 2.5, 4), cast_shadow=shadows), light.Ambient(intensity=0.4))
    return s


def edge_scenes():
    out = {}
    # no geometry at all: every pixel is the background
    out["empty"] = (_lights(scene.Scene()), _cam(), 64, 40, dict(background=(9, 8, 7, 6)))
    # only triangles Triangle.IsValid rejects (zero-length edge, collinear): nothing is drawn
    bad = [[[0, 0, 0], [0, 0, 0], [1, 0, 0]], [[0, 0, 0], [1, 0, 0], [2, 0, 0]], [[0.5, 0.5, 0], [0.5, 0.5, 0], [0.5, 0.5, 0]]]
    out["all_invalid"] = (_lights(scene.Scene(_tri_geo(bad, material.BlinnPhong(texture=_tex())))), _cam(), 64, 40, dict(background=(1, 2, 3, 255)))
    # A collinear triangle that IsValid lets through (float32: cos = 1.0000001, outside the 1e-7 window): a sliver whose screen
    # area is a rounding residue — the exact-safe pruning must decline it (prc_prune.h) and the literal pixel loop decide
    out["sliver_collinear"] = (_lights(scene.Scene(_tri_geo([[[0, 0, 0], [1, 1, 0], [2, 2, 0]]], material.BlinnPhong(texture=_tex())))), _cam(), 64, 40,
                               dict(background=(1, 2, 3, 255)))
    # A triangle in the plane x = 0 seen edge-on by a camera in that plane: all three screen x are equal, Sabc is exactly 0,
    # the barycentrics of the pixel centres on that line are 0/0 = NaN and pass the `< -eps` test, and the NaN depth wins over an empty pixel (bug-list 8): reproduced by
    # the CUDA path's NaN mode (prc_kernels.cuh nan_first / k_nan_fix).
    out["nan_depth_degenerate"] = (_lights(scene.Scene(_tri_geo([[[0, -0.5, 0], [0, 0.5, -1], [0, 0.2, 1]]], material.BlinnPhong(texture=_tex())))), _cam(aspect=65 / 40), 65, 40,
                                   dict(background=(1, 2, 3, 255), nan_depth=True))  # odd width: x = 32.5 is a pixel centre
    # negative material id: vertex colours pass through shade() untouched (raster.go:330-333), persp-correct interpolation
    col = np.array([[0xFF0000FF, 0xFF00FF00, 0xFFFF0000], [0xFF00FFFF, 0x80FFFFFF, 0xFF102030]], np.uint32)
    vc = [[[-1, -0.6, 0], [1, -0.7, -1.5], [0.1, 0.9, 0.4]], [[-0.9, 0.8, -0.5], [-0.2, 0.2, 0.8], [-1.2, -0.1, 0.2]]]
    out["vertex_colours"] = (_lights(scene.Scene(_tri_geo(vc, None, col=col))), _cam(), 96, 60, {})
    # FlatShading(true): the face normal replaces the interpolated one (blinn_cpu.go)
    out["flat_shading"] = (_lights(scene.Scene(_sphere(material.BlinnPhong(texture=_tex(), shininess=16, flat_shading=True)))), _cam(), 96, 60, dict(gamma=True))
    # a texture without mip chain (Texture.useMipmap == false) minified strongly, and UVs far outside [0, 1] incl. negative
    uv = np.array([[[-2.3, -1.7], [3.9, -0.4], [0.2, 4.6]], [[-0.25, 0.5], [-7.0, 8.0], [1.5, -3.0]]], np.float32)
    big = [[[-1.5, -1, -1], [1.5, -1, 0.5], [0, 1.2, -0.3]], [[-1.4, 1.0, 0.2], [-0.3, 0.1, 0.9], [-1.6, -0.4, 0.1]]]
    out["no_mipmap_wild_uv"] = (_lights(scene.Scene(_tri_geo(big, material.BlinnPhong(texture=_tex(64, 5, mip=False)), uv=uv))), _cam(), 96, 60, {})
    out["mipmap_wild_uv"] = (_lights(scene.Scene(_tri_geo(big, material.BlinnPhong(texture=_tex(64, 5, mip=True)), uv=uv))), _cam(), 96, 60, {})
    # vertices behind the eye are perspective-divided like any other (bug-list 7): mirrored positions, then the screen-space clipper
    be = [[[-0.8, -0.5, 0.5], [0.9, -0.4, 0.2], [0.1, 0.4, 4.0]], [[-2.0, 0.1, 3.5], [0.3, -0.8, -1.0], [1.1, 0.9, -0.5]]]
    out["behind_the_eye"] = (_lights(scene.Scene(_tri_geo(be, material.BlinnPhong(texture=_tex())))), _cam(), 96, 60, {})
    # triangles much larger than the screen (clip path + tile path) with a small one in front and a coplanar duplicate (depth tie:
    # the first drawn wins, buffer.go:279)
    huge = [[[-40, -30, -2], [40, -30, -2], [0, 50, -2]], [[-40, -30, -2], [40, -30, -2], [0, 50, -2]], [[-0.3, -0.3, 0.5], [0.3, -0.3, 0.5], [0, 0.3, 0.5]]]
    g1 = _tri_geo(huge[:1], material.BlinnPhong(texture=_tex(64, 6)))
    g2 = _tri_geo(huge[1:], material.BlinnPhong(texture=_tex(64, 7)))
    out["screen_filling_and_depth_tie"] = (_lights(scene.Scene(g1, g2)), _cam(pos=(0, 0, 2.5)), 100, 64, dict(gamma=True))  # camera axis = plane normal
    # thousands of random COPLANAR triangles in two objects: almost every covered pixel is an exact depth tie between several
    # triangles, resolved by draw order (buffer.go:279) — on the CUDA path by the sequence half of the 64-bit visibility key
    rng = np.random.default_rng(3)

    def coplanar(n):
        c = rng.uniform(-1, 1, size=(n, 1, 2)).astype(np.float32)
        d = rng.uniform(-0.15, 0.15, size=(n, 3, 2)).astype(np.float32)
        p = np.concatenate([c + d, np.zeros((n, 3, 1), np.float32)], axis=2)
        e1, e2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        cw = (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]) < 0
        p[cw] = p[cw][:, [0, 2, 1]]
        return p

    tmat = material.BlinnPhong(texture=_tex())
    out["coplanar_depth_ties"] = (_lights(scene.Scene(_tri_geo(coplanar(3000), tmat), _tri_geo(coplanar(3000), tmat))),
                                  camera.Perspective(position=(0, 0, 3), fov=45, aspect=1, near=0.1, far=10), 256, 256, {})
    # pixel (0,0) covered, most pixels not: every uncovered pixel is shaded from G(0,0) with matTable[0] (bug-list 3)
    q = [[[-0.2, -0.2, 0], [0.2, -0.2, 0], [0, 0.2, 0]], [[-3, -3, 0], [1, -3, 0], [-3, 1, 0]]]
    out["pixel00_quirk"] = (scene.Scene(light.Ambient(intensity=1), _tri_geo(q, material.BlinnPhong(texture=material.Texture.uniform((10, 200, 30, 255))))),
                            camera.Perspective(position=(0, 0, 3), fov=45, aspect=1, near=0.1, far=10), 64, 64, dict(background=(1, 2, 3, 4)))
    # frame sizes that are not multiples of anything, down to one pixel
    for w, h in ((1, 1), (3, 2), (33, 17), (7, 301), (130, 5)):
        out[f"size_{w}x{h}"] = (_lights(scene.Scene(_sphere(material.BlinnPhong(texture=_tex(), shininess=8), 12))), _cam(aspect=w / h), w, h, dict(gamma=True))
    # lights: none at all, ambient only, directional only, several ambient terms
    mat = material.BlinnPhong(texture=_tex(), shininess=8)
    out["no_lights"] = (scene.Scene(_sphere(mat, 12)), _cam(), 64, 40, {})
    out["ambient_only_two_terms"] = (scene.Scene(light.Ambient(intensity=0.3), light.Ambient(intensity=0.45), _sphere(mat, 12)), _cam(), 64, 40, {})
    out["directional_only"] = (scene.Scene(light.Directional(intensity=0.9, direction=(-1, -1, -1)), _sphere(mat, 12)), _cam(), 64, 40, {})
    # shadows: caster and receiver in different groups with their own transforms; a second casting light; orthographic camera
    rs = material.BlinnPhong(texture=_tex(), shininess=8, receive_shadow=True)
    grp = scene.Group(_sphere(rs, 16, 0.35, (0.2, 0.45, 0.1)))
    grp.RotateY(0.4)
    gp, gn, gu = synth.ground_mesh(4, half=1.5, amp=0.0, uv_tiles=2.0)
    s = scene.Scene(light.Point(intensity=3, position=(1.5, 3, 2), cast_shadow=True), light.Point(intensity=2, position=(-2, 2.5, 1), cast_shadow=True),
                    light.Point(intensity=1, position=(0, 2, -2)), light.Ambient(intensity=0.3), grp,
                    scene.Geometry(gp, gn, gu, None, np.zeros(gp.shape[0], np.int32), [rs]))
    out["two_casters_grouped"] = (s, _cam(pos=(0, 1.2, 2.6), target=(0, 0.2, 0)), 120, 75, dict(shadow=True, gamma=True))
    ortho = camera.Orthographic(position=(0, 1.0, 2.0), target=(0, 0.2, 0), up=(0, 1, 0), left=-1.6, right=1.6, bottom=-1, top=1, near=-0.1, far=-6)
    out["orthographic_shadow"] = (s, ortho, 120, 75, dict(shadow=True))
    return out


def options(render, s, cam, w, h, o):
    return [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(o.get("shadow", False)), render.GammaCorrection(o.get("gamma", False)),
            render.Background(o.get("background", (0, 0, 0, 0))), render.MSAA(o.get("msaa", 1))]
