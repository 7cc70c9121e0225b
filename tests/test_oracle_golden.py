"""CPU: the oracle (driven through the host mirror) against the MSAA(1) renders committed in the
reference (soft goldens: produced by an unknown commit with Workers>1). These pin
the whole path top-down: transform, coverage, perspective-correct UV, mip LOD, texture filtering,
gamma, Blinn-Phong (within the reference's own 1-2 LSB band) and shadow-map rasterisation."""
import math
import os

import numpy as np
import pytest
from PIL import Image

import oracle_binding as ob
from polyred_b200 import gomath as gm
from polyred_b200 import camera, light, material, model, render, scene

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
A = os.path.join(G, "assets")


def _golden(name):
    return np.asarray(Image.open(os.path.join(G, "ref_renders", name)).convert("RGBA"))


def _cmp(img, gold):
    cov = (img[..., 3] > 0) ^ (gold[..., 3] > 0)
    d = np.abs(img.astype(int) - gold.astype(int)).max(axis=2)
    return int(cov.sum()), int((d > 0).sum()), int((d > 1).sum()), int((d > 8).sum()), int(d.max()), int((gold[..., 3] > 0).sum())


def test_ground_png():
    """internal/examples/ground_test.go:15-37 -> examples/out/ground.png (ambient only, textured)."""
    s = scene.Scene(light.Ambient(intensity=1))
    g = model.Load(os.path.join(A, "ground.obj"))
    g.Scale(2, 2, 2)
    s.Add(g)
    cam = camera.Perspective(position=(0, 3, 3), fov=45, aspect=1, near=0.1, far=10)
    r = render.NewRenderer(render.Camera(cam), render.Size(500, 500), render.Scene(s), render.MSAA(1), render.ShadowMap(False), render._Backend(ob.OracleBackend()))
    xor, ndiff, gt1, gt8, mx, covered = _cmp(r.Render(), _golden("ground.png"))
    assert covered == 9672 and xor == 0
    assert ndiff <= 1 and mx <= 1  # 9671/9672 pixels identical, one off by 1 LSB


def test_perspect_png():
    """internal/examples/perspect_test.go:19-42 -> perspect.png (quad -> 2 tris, UV grid, gamma)."""
    s = scene.Scene()
    g = model.Load(os.path.join(A, "perspect.obj"))
    g.Scale(2, 2, 2)
    s.Add(g)
    cam = camera.Perspective(position=(0, 3, 3), fov=45, aspect=1, near=0.1, far=10)
    r = render.NewRenderer(render.Camera(cam), render.Size(500, 500), render.Scene(s), render.GammaCorrection(True), render._Backend(ob.OracleBackend()))
    xor, ndiff, gt1, gt8, mx, covered = _cmp(r.Render(), _golden("perspect.png"))
    assert covered == 60602 and xor == 0
    assert ndiff <= 11 and mx <= 2


def _gopher_group():
    z = np.load(os.path.join(G, "scene_gopher.npz"))
    grp = scene.Group()
    for i in range(int(z["n_geometries"])):
        mats = []
        for row in z[f"materials{i}"]:
            mats.append(material.BlinnPhong(texture=material.Texture.uniform([int(x) for x in row[9:13]]), diffuse=[int(x) for x in row[0:4]],
                                            specular=[int(x) for x in row[4:8]], shininess=row[8]))
        grp.Add(scene.Geometry(z[f"pos{i}"], z[f"nor{i}"], z[f"uv{i}"], None, z[f"mat{i}"], mats))
    return grp


def test_gopher_png():
    """internal/examples/gopher_test.go:20-51 -> gopher.png: 12 objects, 49 792 triangles, point + ambient,
    un-normalised WordPos lighting (bug-list 1), MaterialID -1 white pixels (PolygonMesh quirk)."""
    s = scene.Scene(light.Point(intensity=5, color=(255, 255, 255, 255), position=(0, 0, 5)), light.Ambient(intensity=0.7))
    m = _gopher_group()
    m.RotateY(-np.float32(math.pi) / np.float32(2))
    m.Normalize()
    s.Add(m)
    cam = camera.Perspective(position=(1, 1, 2), fov=45, aspect=np.float32(500) / np.float32(500), near=0.01, far=600)
    r = render.NewRenderer(render.Camera(cam), render.Size(500, 500), render.Scene(s), render.ShadowMap(True), render._Backend(ob.OracleBackend()))
    img = r.Render()
    gold = _golden("gopher.png")
    xor, ndiff, gt1, gt8, mx, covered = _cmp(img, gold)
    assert covered == 75993 and xor == 0
    assert gt8 == 0 and mx <= 5
    assert gt1 <= 0.012 * covered  # >= 98.8 % of covered pixels within 1 LSB (the golden predates the FMA dot products)
    white_g = int(((gold[..., :3] == 255).all(axis=2) & (gold[..., 3] == 255)).sum())
    white_o = int(((img[..., :3] == 255).all(axis=2) & (img[..., 3] == 255)).sum())
    assert white_g > 4000 and abs(white_g - white_o) <= 0.02 * white_g  # MaterialID -1 faces pass the white vertex colour through


def test_bunny_png_msaa2():
    """internal/examples/bunny_test.go:21-60 -> examples/out/bunny.png: MSAA(2) (render at 1920x1080 with the double-MSAA
    cull box, then imageutil.Resize to 960x540), textured Blinn-Phong with shininess 250, point + ambient light.
    Alpha (coverage after the fixed-point bilinear downsample, texture alpha included) is identical in every pixel; every
    fully covered pixel is within 1 LSB (99.4 % identical). The golden's partially covered pixels carry un-darkened RGB
    (rendered by an older downsample), so their colour is not compared."""
    s = scene.Scene(light.Point(intensity=200, color=(255, 255, 255, 255), position=(-200, 250, 600)), light.Ambient(intensity=0.7))
    m = model.Load(os.path.join(A, "bunny_textured", "bunny.obj"))
    m.Scale(1500, 1500, 1500)
    m.Translate(-700, -5, 350)
    s.Add(m)
    cam = camera.Perspective(position=(-550, 194, 734), target=(-1000, 0, 0), up=(0, 1, 1), fov=45, aspect=np.float32(960) / np.float32(540), near=100, far=600)
    r = render.NewRenderer(render.Camera(cam), render.Size(960, 540), render.Scene(s), render.MSAA(2), render.ShadowMap(True), render._Backend(ob.OracleBackend()))
    img = r.Render()
    gold = _golden("bunny_msaa2.png")
    assert img.shape == gold.shape == (540, 960, 4)
    assert np.array_equal(img[..., 3], gold[..., 3]) and int((gold[..., 3] > 0).sum()) == 133101
    full = gold[..., 3] == 255
    d = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(axis=2)[full]
    assert int(full.sum()) == 129733 and int(d.max()) <= 1 and int((d > 0).sum()) <= 800


def test_testrender_png_msaa2_newscene():
    """render/raster_test.go:32-89 (newscene + TestRender) -> internal/testdata/render.png: the scene BASELINE configs[0] is
    modelled on (textured bunny, point light I=5 with a BLACK colour + ambient 0.5, opaque background), rendered by the
    reference at 1920x1080 with MSAA(2) — a 3840x2160 G-buffer, the size of the headline workload. The background mask is
    identical in all 2 073 600 pixels; of the 231 631 bunny pixels at most a few dozen differ by more than 1 LSB."""
    s = scene.Scene(light.Point(intensity=5, color=(0, 0, 0, 255), position=(-2, 2.5, 6)), light.Ambient(intensity=0.5))
    m = model.Load(os.path.join(A, "bunny_textured", "bunny.obj"))
    m.Rotate(gm.v3(0, 1, 0), -np.float32(np.pi) / np.float32(6))
    m.Scale(4, 4, 4)
    m.Translate(0.1, 0, -0.2)
    s.Add(m)
    w, h = 1920, 1080
    cam = camera.Perspective(position=(0, 1.5, 1), target=(0, 0, -0.5), up=(0, 1, 0), fov=45, aspect=np.float32(w) / np.float32(h), near=0.1, far=3)
    r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.MSAA(2), render.Background((0, 127, 255, 255)),
                           render._Backend(ob.OracleBackend(threads=4)))
    img = r.Render()
    gold = _golden("testrender_msaa2.png")
    assert img.shape == gold.shape == (h, w, 4)
    bg_gold = (gold[..., :3] == np.array([0, 127, 255])).all(axis=2)
    bg_mine = (img[..., :3] == np.array([0, 127, 255])).all(axis=2)
    assert int((bg_gold ^ bg_mine).sum()) == 0 and int((~bg_gold).sum()) == 231631
    d = np.abs(img.astype(int) - gold.astype(int)).max(axis=2)
    assert int(d.max()) <= 12 and int((d > 1).sum()) <= 60 and float((d == 0).mean()) >= 0.95


def _new_plane():
    """model.NewPlane(1, 1) (model/plane.go:17-46): two triangles (v1,v2,v3), (v1,v3,v4) with per-vertex colours, MaterialID -1."""
    v = {1: ((-0.5, 0, -0.5), (0, 1), 0xFF0000FF), 2: ((-0.5, 0, 0.5), (0, 0), 0xFF00FF00), 3: ((0.5, 0, 0.5), (1, 0), 0xFFFF0000), 4: ((0.5, 0, -0.5), (1, 1), 0xFF000000)}
    tris = [(1, 2, 3), (1, 3, 4)]
    pos = np.array([[v[i][0] for i in t] for t in tris], np.float32)
    uv = np.array([[v[i][1] for i in t] for t in tris], np.float32)
    col = np.array([[v[i][2] for i in t] for t in tris], np.uint32)
    nor = np.tile(np.array([0, 1, 0], np.float32), (2, 3, 1))
    return scene.Geometry(pos, nor, uv, col, np.full(2, -1, np.int32), [])


def test_plane_and_normalize_png_vertex_colours():
    """model/plane_test.go:21-38 -> examples/out/plane.png and scene/group_test.go:20-36 -> normalize.png: the only reference
    renders of the VERTEX-COLOUR path (MaterialID -1: no material, shade() passes the interpolated colour through,
    raster.go:330-333,546-551), once with a perspective camera (perspective-correct colour interpolation) and once with an
    orthographic one after Scene.Normalize(). Background masks identical in every pixel; no channel differs by more than 1 LSB
    (the colour is uint8(float32) truncated, raster.go:546-551 — the last-ulp differences of a render made on another
    architecture flip 254.99998 / 255.0)."""
    bg = (0x18, 0x18, 0x18, 255)  # color.FromHex("#181818")
    s = scene.Scene(light.Point(intensity=1, color=(0, 128, 255, 255), position=(2, 2, 2)), _new_plane())
    cam = camera.Perspective(position=(2, 2, 2), fov=45, aspect=1, near=0.1, far=10)
    img = render.NewRenderer(render.Camera(cam), render.Size(500, 500), render.MSAA(2), render.Scene(s), render.Background(bg), render._Backend(ob.OracleBackend())).Render()
    s2 = scene.Scene(_new_plane())
    s2.root.Normalize()
    cam2 = camera.Orthographic(position=(0, 1, 0), target=(0, 0, 0), up=(0, 0, -1), left=-1, right=1, bottom=-1, top=1, near=1, far=-1)
    img2 = render.NewRenderer(render.Camera(cam2), render.Size(500, 500), render.MSAA(2), render.Scene(s2), render.Background(bg), render._Backend(ob.OracleBackend())).Render()
    for name, mine, fg_px in (("plane_msaa2.png", img, 18612), ("normalize_msaa2.png", img2, 126736)):
        gold = _golden(name)
        d = np.abs(mine.astype(int) - gold.astype(int)).max(axis=2)
        bg_gold, bg_mine = (gold[..., :3] == 0x18).all(axis=2), (mine[..., :3] == 0x18).all(axis=2)
        assert int((bg_gold ^ bg_mine).sum()) == 0 and int((~bg_gold).sum()) == fg_px, name
        assert int(d.max()) <= 1, name
        assert len(np.unique(mine[~bg_mine].reshape(-1, 4), axis=0)) > 1000, name   # a colour gradient, not a flat fill


def test_shadow_png_msaa2_two_casting_lights():
    """internal/examples/shadow_test.go:20-85 -> examples/out/shadow.png: textured bunny on the textured ground, TWO
    shadow-casting point lights, every material ReceiveShadow, MSAA(2). Alpha is identical in every pixel; of the 170 248
    fully covered pixels 97.6 % are within 2 LSB and 6 differ by more than 8 (a shadow mismatch would halve the colour):
    the golden predates the FMA dot products, like gopher.png, so it pins shadow-map generation, lookup and the
    ReceiveShadow combine softly."""
    s = scene.Scene(light.Point(intensity=3, position=(4, 4, 2), cast_shadow=True), light.Point(intensity=3, position=(-6, 4, 2), cast_shadow=True),
                    light.Ambient(intensity=0.7))
    m = model.Load(os.path.join(A, "bunny_textured", "bunny.obj"))
    m.Scale(2, 2, 2)
    s.Add(m)
    g = model.Load(os.path.join(A, "ground.obj"))
    g.Scale(2, 2, 2)
    s.Add(g)
    for geo, _ in s.geometries():
        for mt in geo.materials:
            mt.receive_shadow = True
    cam = camera.Perspective(position=(0, 0.6, 0.9), fov=45, aspect=np.float32(960) / np.float32(540), near=0.1, far=2)
    r = render.NewRenderer(render.Camera(cam), render.Size(960, 540), render.Scene(s), render.MSAA(2), render.ShadowMap(True),
                           render._Backend(ob.OracleBackend(threads=4)))
    img = r.Render()
    gold = _golden("shadow_msaa2.png")
    assert np.array_equal(img[..., 3], gold[..., 3]) and int((gold[..., 3] > 0).sum()) == 175149
    full = gold[..., 3] == 255
    d = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(axis=2)[full]
    assert int(full.sum()) == 170248
    assert int((d > 2).sum()) <= 4500 and int((d > 4).sum()) <= 200 and int((d > 8).sum()) <= 10


def test_dragon_png_msaa2_normalized_group():
    """internal/examples/dragon_test.go:18-60 -> examples/out/dragon.png: 9 266-face mesh without a .mtl (default
    material), Scale/Translate then Group.Normalize (scene/group.go:47-64), 30-degree camera, MSAA(2). Alpha identical in
    every pixel; every fully covered pixel within 3 LSB, 75 % identical."""
    s = scene.Scene(light.Point(intensity=2, color=(255, 255, 255, 255), position=(-1.5, -1, 1)), light.Ambient(intensity=0.5))
    m = model.Load(os.path.join(A, "dragon.obj"))
    m.Scale(1.5, 1.5, 1.5)
    m.Translate(0, -0.1, -0.15)
    m.Normalize()
    s.Add(m)
    cam = camera.Perspective(position=(-3, 1.25, -2), target=(0, -0.1, -0.1), up=(0, 1, 0), fov=30, aspect=1, near=0.01, far=1000)
    r = render.NewRenderer(render.Camera(cam), render.Size(500, 500), render.Scene(s), render.MSAA(2), render.ShadowMap(False),
                           render._Backend(ob.OracleBackend(threads=4)))
    img = r.Render()
    gold = _golden("dragon_msaa2.png")
    assert np.array_equal(img[..., 3], gold[..., 3]) and int((gold[..., 3] > 0).sum()) == 60133
    full = gold[..., 3] == 255
    d = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(axis=2)[full]
    assert int(full.sum()) == 56081 and int(d.max()) <= 3 and int((d > 1).sum()) <= 1000 and int((d > 0).sum()) <= 15000


def test_resize_restatements_agree():
    """imageutil.Resize: the oracle's C++ restatement (orc_resize, used for the MSAA downsample) against the numpy
    restatement that is pinned by the reference's mip-chain goldens (test_oracle_kat.py), on random images: integer
    MSAA factors, a mip-style 2^k reduction, non-integer scales and the identity."""
    import ctypes as C
    from polyred_b200 import imageutil
    L = ob.load()
    rng = np.random.default_rng(5)
    for (iw, ih, ow, oh) in ((64, 48, 32, 24), (96, 60, 32, 20), (128, 128, 16, 16), (100, 70, 37, 29), (50, 40, 50, 40), (33, 17, 11, 5), (40, 30, 80, 45)):
        img = rng.integers(0, 256, size=(ih, iw, 4), dtype=np.uint8)
        want = imageutil.resize(ow, oh, img)
        got = np.zeros((oh, ow, 4), np.uint8)
        assert L.orc_resize(img.ctypes.data_as(C.c_void_p), iw, ih, got.ctypes.data_as(C.c_void_p), ow, oh) == 0
        assert np.array_equal(got, want), (iw, ih, ow, oh)


def _benchmark_scene(center=None):
    s = scene.Scene(light.Point(intensity=7, color=(0, 0, 0, 255), position=(4, 4, 2), cast_shadow=True), light.Ambient(intensity=0.5))
    m1 = model.Load(os.path.join(A, "bunny.obj"))
    m1.Scale(2, 2, 2)
    s.Add(m1)
    m2 = model.Load(os.path.join(A, "ground.obj"))
    m2.Scale(2, 2, 2)
    s.Add(m2)
    return s, (m1.objects[0], m2.objects[0])


def test_benchmark_coverage_and_shadow_map_dump():
    """internal/examples/benchmark/benchmark.go:69-101 (bunny + ground, 960x540, casting point light):
    forward coverage vs benchmark.png and the shadow map vs the committed Debug dump shadow-0.png.
    The dump was rendered when the light camera looked at the world-space centre of the GEOMETRY alone
    (SURVEY App. C); the light-camera matrices are host inputs of the path, so the override below only
    reproduces that older host-side choice — the depth raster it pins is the path's own (XOR 0, max diff 0)."""
    s, geos = _benchmark_scene()
    mn = np.minimum(*[g.aabb()[0] for g in geos]) * np.float32(2)
    mx = np.maximum(*[g.aabb()[1] for g in geos]) * np.float32(2)
    c = ((mn + mx) * np.float32(0.5)).astype(np.float32)
    s.Center = lambda: c
    cam = camera.Perspective(position=(0, 0.6, 0.9), fov=45, aspect=np.float32(960) / np.float32(540), near=0.1, far=2)
    be = ob.OracleBackend()
    r = render.NewRenderer(render.Camera(cam), render.Size(960, 540), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render._Backend(be))
    img = r.Render()
    z = np.load(os.path.join(G, "ref_renders", "benchmark_coverage.npz"))
    gold_cov = np.unpackbits(z["covered"])[:960 * 540].reshape(540, 960).astype(bool)
    assert int(((img[..., 3] > 0) ^ gold_cov).sum()) == 0 and int(gold_cov.sum()) == 172805
    dump = _golden("benchmark_shadow-0.png")[..., 0]
    mine = r.shadow_map_image(0)  # what render.Debug(true) saves as shadow-0.png (shadow.go:98-118)
    assert np.array_equal(mine[..., 0], mine[..., 1]) and np.array_equal(mine[..., 0], mine[..., 2]) and (mine[..., 3] == 255).all()
    mine = mine[..., 0]
    assert int((dump > 0).sum()) == 29590
    assert int(((dump > 0) ^ (mine > 0)).sum()) == 0 and int(np.abs(dump.astype(int) - mine.astype(int)).max()) == 0


def test_debug_option_dumps_shadow_maps(tmp_path, monkeypatch, capsys):
    """render.Debug(true) (render/options.go:101-107): timings are printed and passShadows saves shadow-<i>.png for every
    casting light into the working directory (render/shadow.go:98-118)."""
    from PIL import Image
    s, _ = _benchmark_scene()
    cam = camera.Perspective(position=(0, 0.6, 0.9), fov=45, aspect=16 / 9, near=0.1, far=2)
    be = ob.OracleBackend()
    r = render.NewRenderer(render.Camera(cam), render.Size(192, 108), render.Scene(s), render.ShadowMap(True), render.Debug(True), render._Backend(be))
    monkeypatch.chdir(tmp_path)
    r.Render()
    assert "saving (shadow map)... ./shadow-0.png" in capsys.readouterr().out
    dump = np.asarray(Image.open(tmp_path / "shadow-0.png").convert("RGBA"))
    assert dump.shape == (108, 192, 4) and np.array_equal(dump, r.shadow_map_image(0)) and int((dump[..., 0] > 0).sum()) > 500
    z = be.read_shadowmap(0, 192, 108)
    assert dump[10, 20, 0] == int(np.float32(z[108 - 1 - 10, 20]) * np.float32(255))  # pixel (i, j) = depths[i + (H-j-1)*W]


def test_shadow_maps_persist_and_reset():
    s, _ = _benchmark_scene()
    cam = camera.Perspective(position=(0, 0.6, 0.9), fov=45, aspect=16 / 9, near=0.1, far=2)
    be = ob.OracleBackend()
    r = render.NewRenderer(render.Camera(cam), render.Size(192, 108), render.Scene(s), render.ShadowMap(True), render._Backend(be))
    a = r.Render()
    m1 = be.read_shadowmap(0, 192, 108).copy()
    b = r.Render()
    assert np.array_equal(a, b) and np.array_equal(m1, be.read_shadowmap(0, 192, 108))
    be.shadow_reset()
    assert be.read_shadowmap(0, 192, 108).max() == 0


def test_uncovered_pixels_shade_pixel00_quirk():
    """bug-list 3 (render/raster.go:326-332): when G(0,0) is covered, EVERY uncovered pixel is shaded from
    G(0,0)'s attributes with matTable[0]; the background colour is used only if G(0,0) is empty."""
    tex = material.Texture.uniform((10, 200, 30, 255))
    mat = material.BlinnPhong(texture=tex)
    def build(cover00):
        s = scene.Scene(light.Ambient(intensity=1))
        # a small triangle in the middle, plus optionally one over the bottom-left corner (screen pixel (0,0))
        tris = [[[-0.2, -0.2, 0], [0.2, -0.2, 0], [0, 0.2, 0]]]
        if cover00:
            tris.append([[-3, -3, 0], [1, -3, 0], [-3, 1, 0]])
        s.Add(scene.Geometry(np.array(tris, np.float32), materials=[mat]))
        cam = camera.Perspective(position=(0, 0, 3), fov=45, aspect=1, near=0.1, far=10)
        be = ob.OracleBackend()
        r = render.NewRenderer(render.Camera(cam), render.Size(64, 64), render.Scene(s), render.Background((1, 2, 3, 4)), render._Backend(be))
        return r.Render(keep_gbuffer=True), be.read_gbuffer(64, 64)
    img, g = build(False)
    assert not g["ok"][0, 0] and tuple(img[0, 63]) == (1, 2, 3, 4)       # top-right corner: background
    img, g = build(True)
    assert g["ok"][0, 0] and not g["ok"][63, 63]
    assert tuple(img[0, 63]) == (10, 200, 30, 255)                         # uncovered, yet shaded from G(0,0)
