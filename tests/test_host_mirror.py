"""CPU: host-side mirror of the renderer front end (scene traversal, flat material table, uniforms)."""
import numpy as np
import pytest

import oracle_binding as ob
from polyred_b200 import _abi as A
from polyred_b200 import camera, gomath as gm, light, material, render, scene, synth


def _tri(n=1, z=0.0):
    return np.tile(np.array([[[0, 0, z], [1, 0, z], [0, 1, z]]], np.float32), (n, 1, 1))


def test_traversal_order_and_matrix_chain():
    """scene/core.go:86-111, 157-219: root leaves get root.ModelMatrix(); grouped leaves get root*group chain;
    the renderer then multiplies by the geometry's own matrix (render/raster.go:242)."""
    a, b, c = scene.Geometry(_tri()), scene.Geometry(_tri()), scene.Geometry(_tri())
    g = scene.Group(b)
    g.Translate(1, 2, 3)
    inner = scene.Group(c)
    inner.Scale(2, 2, 2)
    g.Add(inner)
    s = scene.Scene(a, light.Point(), g)
    geos = s.geometries()
    assert [x[0] for x in geos] == [a, b, c]
    assert np.array_equal(geos[0][1], gm.identity())
    assert np.array_equal(geos[1][1], g.ModelMatrix())
    assert np.array_equal(geos[2][1], gm.mulm(g.ModelMatrix(), inner.ModelMatrix()))


def test_reference_scene_graph_kats():
    """scene/scene_test.go:19-193 (TestScene) and :331-363 (TestAddMultiple) replayed: the model matrix the iterator hands to
    the renderer for every leaf as the graph is built up — root scaled and translated, a nested group with its own
    transform, a group without one, a leaf added to the root later — and one object added several times."""
    def mats(s):
        out = []
        s.IterObjects(lambda o, m: out.append((o, np.asarray(m, np.float32))))
        return out

    def m4(*v):
        return np.array(v, np.float32).reshape(4, 4)

    s = scene.Scene()
    p1 = scene.Geometry(_tri())
    g = s.Add(p1)                                     # Scene.Add returns the root group
    assert [(o is p1, m.tolist()) for o, m in mats(s)] == [(True, np.eye(4, dtype=np.float32).tolist())]
    g.Scale(2, 2, 2)
    assert np.array_equal(mats(s)[0][1], m4(2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 2, 0, 0, 0, 0, 1))
    g.Translate(1, 2, 3)
    root = m4(2, 0, 0, 1, 0, 2, 0, 2, 0, 0, 2, 3, 0, 0, 0, 1)
    assert np.array_equal(mats(s)[0][1], root)
    p2 = scene.Geometry(_tri())
    g2 = scene.Group()
    g2.Add(p2)
    g.Add(g2)
    g2.Scale(2, 2, 2)
    g2.Translate(1, 1, 1)
    got = mats(s)
    assert [o for o, _ in got] == [p1, p2]
    assert np.array_equal(got[0][1], root) and np.array_equal(got[1][1], m4(4, 0, 0, 3, 0, 4, 0, 4, 0, 0, 4, 5, 0, 0, 0, 1))
    p3 = scene.Geometry(_tri())
    g3 = scene.Group()
    g3.Add(p3)
    s.Add(g3)
    got = mats(s)
    assert [o for o, _ in got] == [p1, p2, p3] and np.array_equal(got[2][1], root)   # an untransformed group: the root's matrix
    p4 = scene.Geometry(_tri())
    s.Add(p4)
    got = mats(s)
    assert [o for o, _ in got] == [p1, p2, p3, p4] and np.array_equal(got[3][1], root)
    assert g3.leaves() == [p3]
    # TestAddMultiple: the same object added six times is drawn six times, a seventh time through a group
    p = scene.Geometry(_tri())
    s2 = scene.Scene()
    for _ in range(6):
        s2.Add(p)
    assert len(mats(s2)) == 6
    gg = scene.Group()
    gg.Add(p)
    s2.Add(gg)
    assert len(mats(s2)) == 7 and len(s2.geometries()) == 7


def test_reference_material_kats():
    """material/pool_test.go:10-32 (TestDefault, TestNewBlinnPhong): the default material is the 1x1 blue texture with
    Kd .7 / Ks .5 / shininess 30 (material/pool.go:15-29); NewBlinnPhong keeps the diffuse colour it is given."""
    m = material.Default()
    assert m.texture.image.shape == (1, 1, 4) and tuple(m.texture.image[0, 0]) == (0, 0, 255, 255)
    assert m.diffuse == (179, 179, 179, 255) and m.specular == (128, 128, 128, 255) and float(m.shininess) == 30.0
    want = (11, 22, 33, 255)
    assert material.BlinnPhong(diffuse=want).diffuse == want


def test_flat_material_table_matches_cpuForwardPass():
    """render/raster.go:252-262: flat id = base + local, negative ids stay negative, table = concatenation."""
    m0, m1, m2 = (material.BlinnPhong(texture=material.Texture.uniform((i, i, i, 255))) for i in (1, 2, 3))
    g0 = scene.Geometry(_tri(3), mat=[0, 1, -1], materials=[m0, m1])
    g1 = scene.Geometry(_tri(2), mat=[0, -1], materials=[m2])
    sd = render.SceneDesc(scene.Scene(g0, g1))
    assert sd.mat.tolist() == [0, 1, -1, 2, -1]
    assert sd.obj_start.tolist() == [0, 3, 5]
    assert sd.struct.n_materials == 3 and sd.struct.n_textures == 3


def test_lights_order_and_center():
    p, d, amb = light.Point(position=(1, 2, 3)), light.Directional(direction=(0, -2, 0)), light.Ambient(intensity=0.5)
    s = scene.Scene(amb, d, scene.Geometry(_tri()), p)
    src, env = s.Lights()
    assert src == [d, p] and env == [amb]
    assert np.allclose(d.direction, [0, -1, 0])
    # Scene.Center (scene/scene.go:49-52): model-space AABB of ALL root objects, lights included
    assert np.allclose(s.Center(), [(0 + 1) / 2, (0 + 2) / 2, (0 + 3) / 2])


def test_frame_uniforms_follow_the_reference_formulas():
    s, cam = synth.city_scene(n_objects=3, obj_stacks=4, obj_slices=4, ground_cells=2, tex_size=8)
    r = render.NewRenderer(render.Camera(cam), render.Size(64, 36), render.Scene(s), render.ShadowMap(True), render._Backend(ob.OracleBackend()))
    fd = r.frame_desc()
    f = fd.struct
    assert f.flags & A.PRC_FRAME_PERSPECT and f.flags & A.PRC_FRAME_SHADOWMAP
    view, proj, vp = cam.ViewMatrix(), cam.ProjMatrix(), gm.viewport_matrix(64, 36)
    geos = s.geometries()
    model = gm.mulm(geos[2][1], geos[2][0].ModelMatrix())
    want_trans = gm.mulm(gm.mulm(proj, view), model)
    assert np.array_equal(np.array(f.objects[2].trans[:], np.float32).reshape(4, 4), want_trans)
    assert np.array_equal(np.array(f.objects[2].normal[:], np.float32).reshape(4, 4), gm.transpose(gm.inv(model)))
    assert np.array_equal(np.array(f.viewport_to_world[:], np.float32).reshape(4, 4), gm.mulm(gm.mulm(gm.inv(view), gm.inv(proj)), gm.inv(vp)))
    # lights 0,2,4,6 cast; each casting point light has an orthographic camera fitted to the view frustum (shadow.go:41-86)
    assert [f.lights[i].cast_shadow for i in range(8)] == [1, 0, 1, 0, 1, 0, 1, 0]
    assert isinstance(r._light_cams[0], camera.Orthographic) and r._light_cams[1] is None


def test_casting_directional_light_is_an_error_not_a_guess():
    """render/shadow.go:67-86,123: only *light.Point gets a light camera; a casting Directional panics."""
    s = scene.Scene(light.Directional(cast_shadow=True), scene.Geometry(_tri(), materials=[material.Default()]))
    with pytest.raises(NotImplementedError):
        render.NewRenderer(render.Camera(camera.Perspective()), render.Scene(s), render.ShadowMap(True), render._Backend(ob.OracleBackend()))


def test_unsupported_options_raise():
    s = scene.Scene(scene.Geometry(_tri(), materials=[material.Default()]))
    with pytest.raises(ValueError):
        render.NewRenderer(render.Camera(camera.Perspective()), render.Scene(s), render.PixelFormat(2), render._Backend(ob.OracleBackend()))
    with pytest.raises(NotImplementedError):  # not on this path
        render.NewRenderer(render.Camera(camera.Perspective()), render.Scene(s), render.Blending(lambda d, s: s), render._Backend(ob.OracleBackend()))
    with pytest.raises(ValueError):
        render.NewRenderer(render.Camera(camera.Perspective()), render.Scene(s), render.MSAA(0), render._Backend(ob.OracleBackend()))
    # MSAA itself is supported: the frame comes back at render.Size
    r = render.NewRenderer(render.Camera(camera.Perspective()), render.Size(40, 30), render.Scene(s), render.MSAA(2), render._Backend(ob.OracleBackend()))
    assert r.Render().shape == (30, 40, 4)


def test_oracle_is_deterministic_and_mt_mode_agrees_without_ties():
    s, cam = synth.mesh_scene(subdiv=12)
    a = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(ob.OracleBackend())).Render()
    b = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(ob.OracleBackend())).Render()
    c = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(ob.OracleBackend(threads=4))).Render()
    assert np.array_equal(a, b)
    assert (np.abs(a.astype(int) - c.astype(int)).max(axis=2) > 0).mean() < 0.01  # Workers>1 only differs at depth ties


def test_oracle_mt_mode_equals_the_sequential_oracle_at_depth_ties():
    """bug-list 11: with Workers > 1 the reference is non-deterministic where fragments tie in depth. The oracle's multithreaded
    mode (CPU-baseline timing, full-size parity test) resolves a tie in favour of the fragment drawn earlier, i.e. exactly like
    the sequential pass, whatever the thread timing: 6000 random coplanar triangles in two objects, 8 threads, several runs."""
    rng = np.random.default_rng(3)

    def tris(n):
        c = rng.uniform(-1, 1, size=(n, 1, 2)).astype(np.float32)
        d = rng.uniform(-0.15, 0.15, size=(n, 3, 2)).astype(np.float32)
        p = np.concatenate([c + d, np.zeros((n, 3, 1), np.float32)], axis=2)
        e1, e2 = p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]
        cw = (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]) < 0
        p[cw] = p[cw][:, [0, 2, 1]]
        return p

    mat = material.BlinnPhong(texture=synth.checker_texture(64, seed=3))
    s = scene.Scene(light.Point(intensity=3, position=(0, 0, 3)), light.Ambient(intensity=0.4),
                    scene.Geometry(tris(3000), None, None, None, np.zeros(3000, np.int32), [mat]),
                    scene.Geometry(tris(3000), None, None, None, np.zeros(3000, np.int32), [mat]))
    cam = camera.Perspective(position=(0, 0, 3), fov=45, aspect=1, near=0.1, far=10)

    def run(threads):
        be = ob.OracleBackend(threads)
        img = render.NewRenderer(render.Camera(cam), render.Size(256, 256), render.Scene(s), render._Backend(be)).Render(keep_gbuffer=True).copy()
        return img, be.read_gbuffer(256, 256)

    a, ga = run(1)
    covered = ga["ok"].astype(bool)
    assert covered.sum() > 40000
    for _ in range(4):
        b, gb = run(8)
        assert np.array_equal(ga["tri"], gb["tri"]) and np.array_equal(ga["sub"], gb["sub"]) and np.array_equal(a, b)


def test_view_frames_match_render_views_on_the_oracle():
    """ViewFrames() (PRC_FRAME_SHADOW_RESET, views submitted back to back) gives the frames RenderViews() gives
    (Options(Camera) + Render per view) — checked on the CPU oracle; the CUDA twin is in test_gpu_parity.py."""
    import math
    import oracle_binding as ob
    from polyred_b200 import render, synth
    cams = []
    for k in range(2):
        s, cam = synth.city_scene(n_objects=6, obj_stacks=8, obj_slices=8, ground_cells=12, tex_size=16, cam_angle=2 * math.pi * k / 2, cam_radius=2.6, cam_height=1.1)
        cams.append(cam)
    opts = [render.Camera(cams[0]), render.Size(96, 54), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    want = render.RenderViews(render.NewRenderer(*opts, render._Backend(ob.OracleBackend())), cams)
    rc = render.NewRenderer(*opts, render._Backend(ob.OracleBackend()))
    rc._ensure_uploaded()
    for fd, w in zip(render.ViewFrames(rc, cams), want):
        out = np.zeros((54, 96, 4), np.uint8)
        rc._backend.render(fd, out)
        assert np.array_equal(out, w)
    assert not np.array_equal(want[0], want[1])


def test_scene_cache_tracks_transforms_and_membership():
    """the flattened scene is uploaded once; moving a TransformContext only recomputes the per-object
    matrices (no upload), adding an object re-flattens and re-uploads — all without InvalidateScene()."""
    class Counting(ob.OracleBackend):
        uploads = 0

        def scene_upload(self, sd):
            Counting.uploads += 1
            return super().scene_upload(sd)

    s, cam = synth.mesh_scene(subdiv=10)
    be = Counting()
    r = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(be))
    a = r.Render().copy()
    b = r.Render().copy()
    assert Counting.uploads == 1 and np.array_equal(a, b)
    geo = s.geometries()[0][0]
    geo.Translate(0.3, 0.1, 0.0)
    c = r.Render().copy()
    assert Counting.uploads == 1 and not np.array_equal(a, c)          # matrices moved, soup stayed resident
    fresh = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(ob.OracleBackend())).Render()
    assert np.array_equal(c, fresh)                                     # same frame as a renderer that never cached
    s.root.Scale(0.5, 0.5, 0.5)                                         # a group transform above the leaf
    d = r.Render().copy()
    assert Counting.uploads == 1 and not np.array_equal(c, d)
    extra, _ = synth.mesh_scene(subdiv=6)
    g2 = extra.geometries()[0][0]
    g2.Translate(-0.8, 0.0, 0.0)
    s.Add(g2)
    e = r.Render().copy()
    assert Counting.uploads == 2 and not np.array_equal(d, e)           # membership changed: re-flattened and re-uploaded
    fresh = render.NewRenderer(render.Camera(cam), render.Size(96, 60), render.Scene(s), render._Backend(ob.OracleBackend())).Render()
    assert np.array_equal(e, fresh)


def test_scene_signature_is_cached_until_something_mutates():
    """Scene.signature() walks the graph only after a mutation (process-wide epoch): transforms, Add(), direct edits of a
    group's object list and replaced vertex arrays are all seen; an unchanged scene costs O(1) per frame."""
    from polyred_b200 import gomath as gm
    s, _ = synth.mesh_scene(subdiv=6)
    geo = s.geometries()[0][0]
    a = s.signature()
    e0 = gm.EPOCH[0]
    assert s.signature() == a and s._sig_cache[0] == e0 == gm.EPOCH[0]          # served from the cache
    geo.RotateY(0.25)
    b = s.signature()
    assert b[0] == a[0] and b[1] != a[1]                                        # same members, moved transform
    extra = synth.mesh_scene(subdiv=4)[0].geometries()[0][0]
    s.root.objects.append(extra)                                                # not through Add()
    c = s.signature()
    assert c[0] != b[0]
    s.root.objects.pop()
    assert s.signature()[0] == b[0]
    geo.pos = geo.pos.copy()                                                    # a new vertex array object
    assert s.signature()[0] != b[0]
    s.root.objects = [geo]                                                      # replaced list stays tracked
    d = s.signature()
    s.root.objects.append(extra)
    assert s.signature()[0] != d[0]


def test_pixel_format_bgra_swaps_red_and_blue():
    """render.PixelFormat(buffer.PixelFormatBGRA): the colour bytes are stored B,G,R,A (buffer/buffer.go:242-251)."""
    s, cam = synth.mesh_scene(subdiv=10)
    opts = [render.Camera(cam), render.Size(96, 60), render.Scene(s), render.GammaCorrection(True), render.MSAA(2)]
    rgba = render.NewRenderer(*opts, render._Backend(ob.OracleBackend())).Render()
    bgra = render.NewRenderer(*opts, render.PixelFormat(1), render._Backend(ob.OracleBackend())).Render()
    assert np.array_equal(bgra, rgba[..., [2, 1, 0, 3]]) and not np.array_equal(bgra, rgba)


def test_msaa_strip_halo_covers_the_resize_filter():
    """prc_render_peer shades `msaa` supersampled rows beyond a strip before it downsamples the strip's own output rows (MSAA
    frames cut into strips need no exchange). That halo must contain every input row imageutil.Resize reads with a non-zero
    coefficient (createWeights8, resize.go:143-164) for the strip's output rows — checked here for every MSAA factor and a range
    of heights, and end to end: a strip-wise downsample with the halo equals the whole-frame Resize."""
    from polyred_b200 import imageutil
    for m in range(2, 9):
        for out_h in (1, 2, 5, 17, 100, 135, 540):
            coeffs, start, fl = imageutil._create_weights8(out_h, 2, np.float32(out_h * m) / np.float32(out_h))
            rows = start[:, None] + np.arange(fl)[None, :]
            used = coeffs != 0
            y = np.arange(out_h)[:, None]
            assert (rows[used] >= (m * y - m + 0 * rows)[used]).all() and (rows[used] < (m * y + 2 * m + 0 * rows)[used]).all(), (m, out_h)
    rng = np.random.default_rng(9)
    m, out_h, out_w = 2, 30, 16
    big = rng.integers(0, 256, size=(out_h * m, out_w * m, 4), dtype=np.uint8)
    want = imageutil.resize(out_w, out_h, big)
    got = np.zeros_like(want)
    for o0, o1 in ((0, 7), (7, 19), (19, 30)):                      # three strips of output rows
        a, b = max(0, o0 * m - m), min(out_h * m, o1 * m + m)      # the rows a rank shades: its strip + the halo
        masked = np.zeros_like(big)
        masked[a:b] = big[a:b]                                     # everything else is garbage to this rank (zeros here)
        got[o0:o1] = imageutil.resize(out_w, out_h, masked)[o0:o1]
    assert np.array_equal(got, want)
