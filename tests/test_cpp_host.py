"""CPU: the C++ host side (host/polyred_host.hpp, driven through host/capi.cpp) against the Python mirror.

polyred is compiled code, so the host side above the C ABI also exists in C++ (the reference's option / camera / light /
material / scene interface). Both mirrors must hand the library the SAME BITS: the math restatements, the texture mip
chains, every uniform of prc_frame, and the frames rendered from them through the same backend (the CPU oracle here, which
exports the ABI of libpolyred_cuda.so under the "orc_" prefix) are compared byte for byte."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

import oracle_binding as ob
from polyred_b200 import _abi as A
from polyred_b200 import camera, gomath as gm, imageutil, light, material, render, scene, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "host")
vp = C.c_void_p


@pytest.fixture(scope="module")
def H():
    subprocess.check_call(["make", "-C", HOST, "-s", "libpolyred_host.so"])
    ob.load()  # builds oracle/libpr_oracle.so if needed
    L = C.CDLL(os.path.join(HOST, "libpolyred_host.so"))
    for name in ("pth_scene_new", "pth_group_new", "pth_texture_new", "pth_material_new", "pth_geometry_new", "pth_camera_perspective",
                 "pth_camera_orthographic", "pth_renderer_new"):
        getattr(L, name).restype = vp
    L.pth_last_frame.restype = C.POINTER(A.prc_frame)
    L.pth_mat4_det.restype = C.c_float
    L.pth_texture_new.argtypes = [C.c_int, C.c_int, vp, C.c_int]
    L.pth_texture_level.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), vp]
    L.pth_material_new.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_float, C.c_int, C.c_int, C.c_int]
    L.pth_geometry_new.argtypes = [C.c_uint64, vp, vp, vp, vp, vp, C.POINTER(vp), C.c_int]
    L.pth_group_add.argtypes = [vp, vp]
    L.pth_scene_add.argtypes = [vp, vp]
    L.pth_scene_add_light.argtypes = [vp, C.c_int, C.c_float, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_int]
    L.pth_xf.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.pth_scene_root_xf.argtypes = [vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    L.pth_model_matrix.argtypes = [vp, vp]
    L.pth_camera_perspective.argtypes = [vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float]
    L.pth_camera_orthographic.argtypes = [vp, vp, vp] + [C.c_float] * 6
    L.pth_camera_matrices.argtypes = [vp, vp, vp]
    L.pth_renderer_new.argtypes = [C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.pth_render.argtypes = [vp, vp, C.c_char_p, C.c_int]
    L.pth_render_view.argtypes = [vp, vp, C.c_char_p, C.c_int]
    L.pth_set_camera.argtypes = [vp, vp, C.c_char_p, C.c_int]
    L.pth_last_frame.argtypes = [vp]
    L.pth_renderer_free.argtypes = [vp]
    L.pth_shadow_map_image.argtypes = [vp, C.c_int, vp, C.c_char_p, C.c_int]
    L.pth_set_debug.argtypes = [vp, C.c_int, C.c_char_p, C.c_int]
    L.pth_set_blending.argtypes = [vp, C.c_char_p, C.c_int]
    L.pth_mat4_inv.argtypes = [vp, vp]
    L.pth_mat4_mulm.argtypes = [vp, vp, vp]
    L.pth_mat4_det.argtypes = [vp]
    L.pth_gamma_lut.argtypes = [vp]
    L.pth_resize.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
    return L


def _p(a):
    return a.ctypes.data_as(vp)


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_math_restatements_are_bit_identical(H):
    rng = np.random.default_rng(11)
    for _ in range(200):
        a = rng.normal(size=(4, 4)).astype(np.float32) * np.float32(10 ** rng.uniform(-2, 2))
        b = rng.normal(size=(4, 4)).astype(np.float32)
        out = np.zeros((4, 4), np.float32)
        H.pth_mat4_mulm(_p(a), _p(b), _p(out))
        assert np.array_equal(_bits(out), _bits(gm.mulm(a, b)))
        H.pth_mat4_inv(_p(a), _p(out))
        assert np.array_equal(_bits(out), _bits(gm.inv(a)))
        assert np.float32(H.pth_mat4_det(_p(a))).view(np.uint32) == np.float32(gm.det(a)).view(np.uint32)
    lut = np.zeros(256, np.uint8)
    H.pth_gamma_lut(_p(lut))
    assert np.array_equal(lut, imageutil.gamma_lut_u8())
    for (iw, ih, ow, oh) in ((64, 48, 32, 24), (100, 70, 37, 29), (33, 17, 11, 5)):
        img = rng.integers(0, 256, size=(ih, iw, 4), dtype=np.uint8)
        got = np.zeros((oh, ow, 4), np.uint8)
        H.pth_resize(_p(img), iw, ih, _p(got), ow, oh)
        assert np.array_equal(got, imageutil.resize(ow, oh, img))


def test_cameras_transform_contexts_and_mip_chains(H):
    rng = np.random.default_rng(12)
    f3 = lambda v: np.array(v, np.float32)
    view, proj = np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32)
    for _ in range(50):
        pos, tgt = rng.normal(size=3).astype(np.float32) * 3, rng.normal(size=3).astype(np.float32)
        fov, asp, n, f = np.float32(rng.uniform(20, 90)), np.float32(rng.uniform(0.5, 2.5)), np.float32(rng.uniform(0.01, 1)), np.float32(rng.uniform(2, 1000))
        c = H.pth_camera_perspective(_p(pos), _p(tgt), _p(f3([0, 1, 0])), fov, asp, n, f)
        H.pth_camera_matrices(c, _p(view), _p(proj))
        pc = camera.Perspective(position=pos, target=tgt, up=(0, 1, 0), fov=fov, aspect=asp, near=n, far=f)
        assert np.array_equal(_bits(view), _bits(pc.ViewMatrix())) and np.array_equal(_bits(proj), _bits(pc.ProjMatrix()))
        l, r, b, t = (np.float32(x) for x in (rng.uniform(-5, -1), rng.uniform(1, 5), rng.uniform(-4, -1), rng.uniform(1, 4)))
        c = H.pth_camera_orthographic(_p(pos), _p(tgt), _p(f3([0, 1, 0])), l, r, b, t, np.float32(1.5), np.float32(-3))
        H.pth_camera_matrices(c, _p(view), _p(proj))
        oc = camera.Orthographic(position=pos, target=tgt, up=(0, 1, 0), left=l, right=r, bottom=b, top=t, near=1.5, far=-3)
        assert np.array_equal(_bits(view), _bits(oc.ViewMatrix())) and np.array_equal(_bits(proj), _bits(oc.ProjMatrix()))
    # TransformContext: the same operations on both sides
    g = H.pth_group_new()
    pg = scene.Group()
    mm = np.zeros((4, 4), np.float32)
    for _ in range(40):
        op = int(rng.integers(0, 3))
        a, b, c_, d = (np.float32(x) for x in rng.uniform(-2, 2, size=4))
        if op == 0:
            pg.Scale(a, b, c_)
        elif op == 1:
            pg.Translate(a, b, c_)
        else:
            pg.Rotate(gm.v3(a, b, c_), d)
        H.pth_xf(g, op, a, b, c_, d)
        H.pth_model_matrix(g, _p(mm))
        assert np.array_equal(_bits(mm), _bits(pg.ModelMatrix()))
    # buffer.NewTexture mip chain
    img = rng.integers(0, 256, size=(80, 96, 4), dtype=np.uint8)  # non-square: 80 rows x 96 columns
    t = H.pth_texture_new(img.shape[1], img.shape[0], _p(img), 1)
    want = imageutil.build_mipmap(img)
    w, h = C.c_int(), C.c_int()
    assert H.pth_texture_level(t, 0, C.byref(w), C.byref(h), None) == len(want)
    for i, lv in enumerate(want):
        H.pth_texture_level(t, i, C.byref(w), C.byref(h), None)
        assert (h.value, w.value) == lv.shape[:2]
        got = np.zeros(lv.shape, np.uint8)
        H.pth_texture_level(t, i, C.byref(w), C.byref(h), _p(got))
        assert np.array_equal(got, lv)


def _build_both(H, msaa):
    """The same scene through both mirrors: two textured meshes (one inside a transformed group), a ground, a non-casting
    directional light, a shadow-casting point light, ambient light; every material receives shadows."""
    rng = np.random.default_rng(5)
    tex_img = rng.integers(0, 256, size=(64, 64, 4), dtype=np.uint8)
    tex_img[..., 3] = 255
    sp, sn, su = synth.sphere_mesh(10, 12, seed=2)
    gp, gn, gu = synth.ground_mesh(4, half=2.0, amp=0.05)
    keep = []

    # ---- Python mirror
    ptex = material.Texture(tex_img)
    pm1 = material.BlinnPhong(texture=ptex, diffuse=(200, 190, 180, 255), specular=(90, 90, 90, 255), shininess=16, receive_shadow=True)
    pm2 = material.BlinnPhong(texture=material.Texture.uniform((180, 200, 160, 255)), shininess=8, receive_shadow=True, flat_shading=True)
    ps = scene.Scene(light.Directional(intensity=0.9, direction=(-1, -1, -1)), light.Point(intensity=3, position=(4, 4, 2), cast_shadow=True),
                     light.Ambient(intensity=0.5))
    pg1 = scene.Geometry(sp, sn, su, None, np.zeros(sp.shape[0], np.int32), [pm1])
    pg1.Scale(0.5, 0.5, 0.5); pg1.Rotate(gm.v3(0.2, 1, 0.1), 0.7); pg1.Translate(0, 0.55, 0)
    ps.Add(pg1)
    ps.Add(scene.Geometry(gp, gn, gu, None, np.zeros(gp.shape[0], np.int32), [pm2]))
    mats = np.where(np.arange(sp.shape[0]) % 3 == 0, 1, 0).astype(np.int32)
    pg2 = scene.Geometry(sp, sn, su, None, mats, [pm2, pm1])
    pg2.Scale(0.3, 0.3, 0.3)
    grp = scene.Group(pg2)
    grp.Translate(0.9, 0.4, -0.3); grp.Rotate(gm.v3(0, 1, 0), 1.1)
    ps.Add(grp)
    pcam = camera.Perspective(position=(0, 1.0, 2.2), target=(0, 0.4, 0), up=(0, 1, 0), fov=45, aspect=1.6, near=0.1, far=10)
    pr = render.NewRenderer(render.Camera(pcam), render.Size(160, 100), render.Scene(ps), render.ShadowMap(True), render.GammaCorrection(True),
                            render.MSAA(msaa), render._Backend(ob.OracleBackend()))

    # ---- C++ mirror, the same calls
    tex = H.pth_texture_new(64, 64, _p(tex_img), 1)
    m1 = H.pth_material_new(tex, A.pack_rgba((200, 190, 180, 255)), A.pack_rgba((90, 90, 90, 255)), 16, 0, 0, 1)
    uni = np.array([180, 200, 160, 255], np.uint8)
    m2 = H.pth_material_new(H.pth_texture_new(1, 1, _p(uni), 1), A.pack_rgba(material.color_from_value(0.5, 0.5, 0.5, 1.0)),
                            A.pack_rgba(material.color_from_value(0.5, 0.5, 0.5, 1.0)), 8, 1, 0, 1)
    s = H.pth_scene_new()
    d = gm.v3(-1, -1, -1)
    H.pth_scene_add_light(s, 1, 0.9, A.pack_rgba((255, 255, 255, 255)), d[0], d[1], d[2], 0)
    H.pth_scene_add_light(s, 0, 3, A.pack_rgba((255, 255, 255, 255)), 4, 4, 2, 1)
    H.pth_scene_add_light(s, 2, 0.5, A.pack_rgba((255, 255, 255, 255)), 0, 0, 0, 0)
    z = np.zeros(sp.shape[0], np.int32)
    g1 = H.pth_geometry_new(sp.shape[0], _p(sp), _p(sn), _p(su), None, _p(z), (vp * 1)(m1), 1)
    H.pth_xf(g1, 0, 0.5, 0.5, 0.5, 0); H.pth_xf(g1, 2, 0.2, 1, 0.1, 0.7); H.pth_xf(g1, 1, 0, 0.55, 0, 0)
    H.pth_scene_add(s, g1)
    zg = np.zeros(gp.shape[0], np.int32)
    H.pth_scene_add(s, H.pth_geometry_new(gp.shape[0], _p(gp), _p(gn), _p(gu), None, _p(zg), (vp * 1)(m2), 1))
    g2 = H.pth_geometry_new(sp.shape[0], _p(sp), _p(sn), _p(su), None, _p(mats), (vp * 2)(m2, m1), 2)
    H.pth_xf(g2, 0, 0.3, 0.3, 0.3, 0)
    grp_c = H.pth_group_new()
    H.pth_group_add(grp_c, g2)
    H.pth_xf(grp_c, 1, 0.9, 0.4, -0.3, 0); H.pth_xf(grp_c, 2, 0, 1, 0, 1.1)
    H.pth_scene_add(s, grp_c)
    f3 = lambda v: np.array(v, np.float32)
    pos, tgt, up = f3([0, 1.0, 2.2]), f3([0, 0.4, 0]), f3([0, 1, 0])
    cam = H.pth_camera_perspective(_p(pos), _p(tgt), _p(up), 45, np.float32(1.6), np.float32(0.1), 10)
    err = C.create_string_buffer(512)
    so = os.path.join(ROOT, "oracle", "libpr_oracle.so").encode()
    r = H.pth_renderer_new(160, 100, cam, s, 1, 1, msaa, 0, 0, so, b"orc_", 0, err, 512)
    assert r, err.value
    keep += [tex_img, sp, sn, su, gp, gn, gu, z, zg, mats, uni, pos, tgt, up]
    return pr, r, keep


@pytest.mark.parametrize("msaa", [1, 2])
def test_uniforms_and_frames_match_the_python_mirror(H, msaa):
    pr, r, keep = _build_both(H, msaa)
    want = pr.Render()
    got = np.zeros_like(want)
    err = C.create_string_buffer(512)
    assert H.pth_render(r, _p(got), err, 512) == 0, err.value
    # every uniform of prc_frame, bit for bit
    fc, fp = H.pth_last_frame(r).contents, pr._last_frame.struct
    for name in ("flags", "width", "height", "n_objects", "n_lights", "n_ambient", "background_rgba", "row0", "row1", "msaa"):
        assert getattr(fc, name) == getattr(fp, name), name
    for name in ("viewport", "viewport_inv", "proj_inv", "view_inv", "viewport_to_world", "cam_pos"):
        assert np.array_equal(_bits(np.array(getattr(fc, name)[:])), _bits(np.array(getattr(fp, name)[:]))), name
    assert bytes(fc.gamma_lut) == bytes(fp.gamma_lut)
    n = fc.n_objects
    for o in range(n):
        for name in ("trans", "normal"):
            assert np.array_equal(_bits(np.array(getattr(fc.objects[o], name)[:])), _bits(np.array(getattr(fp.objects[o], name)[:]))), (o, name)
    for i in range(fc.n_lights):
        lc, lp = fc.lights[i], fp.lights[i]
        assert (lc.kind, lc.cast_shadow, lc.color_rgba) == (lp.kind, lp.cast_shadow, lp.color_rgba)
        assert np.array_equal(_bits(np.array(lc.pos[:] + [lc.intensity])), _bits(np.array(lp.pos[:] + [lp.intensity])))
        if lc.cast_shadow:
            assert np.array_equal(_bits(np.array(lc.view[:] + lc.proj[:])), _bits(np.array(lp.view[:] + lp.proj[:])))
            assert np.array_equal(_bits(np.ctypeslib.as_array(lc.shadow_trans, (n * 16,))), _bits(np.ctypeslib.as_array(lp.shadow_trans, (n * 16,))))
    assert np.array_equal(np.ctypeslib.as_array(fc.ambient_intensity, (fc.n_ambient,)), np.ctypeslib.as_array(fp.ambient_intensity, (fp.n_ambient,)))
    # and the frame rendered from them
    assert np.array_equal(got, want) and int((want[..., 3] > 0).sum()) > 3000
    # the zero-copy path (rgba_out = NULL + host_image) returns the same frame
    got2 = np.zeros_like(want)
    assert H.pth_render_view(r, _p(got2), err, 512) == 0, err.value
    assert np.array_equal(got2, want)
    H.pth_renderer_free(r)


def test_debug_option_and_shadow_map_dump_match_the_python_mirror(H, tmp_path, monkeypatch, capfd):
    """render.Debug(true) (options.go:101-107, shadow.go:98-118) in the C++ host: the same shadow-map picture as the Python
    mirror (which reproduces the reference's committed dump byte for byte), timings on stdout, shadow-<i>.ppm in the working
    directory; Workers / BatchSize are accepted; Blending is rejected, not ignored."""
    pr, r, keep = _build_both(H, 1)
    want = pr.Render()
    err = C.create_string_buffer(512)
    monkeypatch.chdir(tmp_path)
    assert H.pth_set_debug(r, 1, err, 512) == 0, err.value
    got = np.zeros_like(want)
    assert H.pth_render(r, _p(got), err, 512) == 0, err.value
    assert np.array_equal(got, want)
    out = capfd.readouterr().out
    assert "entire rendering:" in out and "saving (shadow map)... shadow-1.ppm" in out      # light 1 is the casting point light
    dump = pr.shadow_map_image(1)
    img = np.zeros_like(dump)
    assert H.pth_shadow_map_image(r, 1, _p(img), err, 512) == 0, err.value
    assert np.array_equal(img, dump) and int((dump[..., 0] > 0).sum()) > 300
    ppm = (tmp_path / "shadow-1.ppm").read_bytes()
    head = b"P6\n160 100\n255\n"
    assert ppm.startswith(head) and np.array_equal(np.frombuffer(ppm[len(head):], np.uint8).reshape(100, 160, 3), dump[..., :3])
    assert not (tmp_path / "shadow-0.ppm").exists()                                           # the directional light casts nothing
    assert H.pth_set_blending(r, err, 512) != 0 and b"Blending" in err.value
    H.pth_renderer_free(r)


def test_errors_surface(H):
    err = C.create_string_buffer(512)
    s = H.pth_scene_new()
    f3 = lambda v: np.array(v, np.float32)
    cam = H.pth_camera_perspective(_p(f3([0, 0, 2])), _p(f3([0, 0, 0])), _p(f3([0, 1, 0])), 45, 1.0, 0.1, 10)
    assert not H.pth_renderer_new(64, 64, cam, s, 0, 0, 1, 0, 0, b"/nonexistent/libpolyred_cuda.so", b"prc_", 0, err, 512)
    assert b"cannot load" in err.value and b"no CPU fallback" in err.value
    so = os.path.join(ROOT, "oracle", "libpr_oracle.so").encode()
    assert not H.pth_renderer_new(64, 64, cam, s, 0, 0, 9, 0, 0, so, b"orc_", 0, err, 512)
    assert b"MSAA" in err.value
    # a shadow-casting Directional light has no light camera in the reference (it panics): an error, not a guess
    H.pth_scene_add_light(s, 1, 1.0, 0xFFFFFFFF, 0, -1, 0, 1)
    assert not H.pth_renderer_new(64, 64, cam, s, 1, 0, 1, 0, 0, so, b"orc_", 0, err, 512)
    assert b"Directional" in err.value
