"""CPU: the oracle and the host-side math mirror replay the reference's own known-answer tests
(tests/golden/kat.json, extracted from the Go test sources by tests/golden/make_fixtures.py)."""
import ctypes as C
import json
import math
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle_binding import f32a, fptr
from polyred_b200 import _abi as A
from polyred_b200 import camera, gomath as gm, imageutil, material

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
ASSETS = os.path.join(HERE, "golden", "assets")


def mat_eq(a, b, eps=np.float32(1e-7)):
    """Mat4.Eq (math/mat4.go:162-179): every element ApproxEq within float32(1e-7)."""
    a, b = np.asarray(a, np.float32).reshape(-1), np.asarray(b, np.float32).reshape(-1)
    return bool(np.all(np.abs(a - b) <= eps))


def test_barycoord(oracle_lib):
    k = KAT["barycoord"]
    out = f32a(0, 0, 0)
    oracle_lib.orc_barycoord(fptr(f32a(*k["p"])), fptr(f32a(*k["t1"])), fptr(f32a(*k["t2"])), fptr(f32a(*k["t3"])), fptr(out))
    assert out.tolist() == k["want"]


def test_lerpc(oracle_lib):
    k = KAT["lerpc"]
    t = f32a(k["t"])
    got = oracle_lib.orc_lerpc(A.pack_rgba(k["from"]), A.pack_rgba(k["to"]), fptr(t))
    assert got == A.pack_rgba(k["want"])


def test_mat4_mulm_oracle_and_host(oracle_lib):
    k = KAT["mat4_mulm"]
    out = np.zeros(16, np.float32)
    oracle_lib.orc_mat4_mulm(fptr(f32a(*k["a"])), fptr(f32a(*k["b"])), fptr(out))
    assert out.tolist() == k["want"]
    assert gm.mulm(np.array(k["a"], np.float32).reshape(4, 4), np.array(k["b"], np.float32).reshape(4, 4)).reshape(-1).tolist() == k["want"]


def test_mat4_mulv(oracle_lib):
    k = KAT["mat4_mulv"]
    out = np.zeros(4, np.float32)
    oracle_lib.orc_mat4_mulv(fptr(f32a(*k["m"])), fptr(f32a(*k["v"])), fptr(out))
    assert out.tolist() == k["want"]
    assert gm.mulv(np.array(k["m"], np.float32).reshape(4, 4), f32a(*k["v"])).tolist() == k["want"]


def test_mat4_det_transpose_inv_host():
    k = KAT["mat4_det"]
    m = np.array(k["m"], np.float32).reshape(4, 4)
    assert float(gm.det(m)) == k["want"]
    assert np.array_equal(gm.transpose(m), m.T)
    k = KAT["mat4_inv"]
    m = np.array(k["m"], np.float32).reshape(4, 4)
    assert mat_eq(gm.inv(m), np.array(k["want"], np.float64).astype(np.float32))
    with pytest.raises(ZeroDivisionError):
        gm.inv(np.zeros((4, 4), np.float32))  # the reference panics ("zero determinant")


def test_inv_broadcasts_bitwise():
    rng = np.random.default_rng(0)
    ms = (rng.standard_normal((7, 4, 4)) * 3).astype(np.float32)
    batched = gm.inv(ms)
    for i in range(7):
        assert np.array_equal(batched[i].view(np.uint32), gm.inv(ms[i]).view(np.uint32))
    assert np.array_equal(gm.mulm(ms, ms[::-1])[2].view(np.uint32), gm.mulm(ms[2], ms[4]).view(np.uint32))


def test_view_and_projection_matrices():
    k = KAT["view_matrix"]
    assert mat_eq(camera.ViewMatrix(k["pos"], k["target"], k["up"]), k["want"])
    k = KAT["proj_perspective"]
    cam = camera.Perspective(fov=k["fov"], aspect=k["aspect"], near=k["near"], far=k["far"])
    assert mat_eq(cam.ProjMatrix(), k["want"])
    k = KAT["proj_orthographic"]
    cam = camera.Orthographic(left=k["left"], right=k["right"], top=k["top"], bottom=k["bottom"], near=k["near"], far=k["far"])
    assert mat_eq(cam.ProjMatrix(), k["want"])


def test_viewport_matrix():
    """math/math_test.go:40 — ViewportMatrix(w,h) = [w/2 0 0 w/2; 0 h/2 0 h/2; 0 0 1 0; 0 0 0 1]."""
    m = gm.viewport_matrix(800, 500)
    assert m.reshape(-1).tolist() == [400, 0, 0, 400, 0, 250, 0, 250, 0, 0, 1, 0, 0, 0, 0, 1]


def _query(oracle_lib, tex: material.Texture, lod, u, v):
    n = len(tex.mipmap)
    levels = [np.ascontiguousarray(l) for l in tex.mipmap]
    w = np.array([l.shape[1] for l in levels], np.uint32)
    h = np.array([l.shape[0] for l in levels], np.uint32)
    ptrs = (C.c_void_p * n)(*[l.ctypes.data for l in levels])
    got = oracle_lib.orc_texture_query(n, fptr(w), fptr(h), ptrs, 1 if tex.use_mipmap else 0, fptr(f32a(lod)), fptr(f32a(u)), fptr(f32a(v)))
    return [got & 255, (got >> 8) & 255, (got >> 16) & 255, (got >> 24) & 255]


@pytest.fixture(scope="module")
def textures():
    k = KAT["texture_query"]
    return {
        "default": material.Texture(),
        "2x2": material.Texture(np.array(k["data_2x2"], np.uint8).reshape(2, 2, 4), use_mipmap=True),
        # mustLoadTexture (buffer/texture_test.go:168-178): LoadImage without gamma correction
        "ground.png": material.Texture(imageutil.load_image(os.path.join(ASSETS, "ground.png")), use_mipmap=True),
        "pic.jpg": material.Texture(imageutil.load_image(os.path.join(ASSETS, "pic.jpg")), use_mipmap=True),
    }


@pytest.mark.parametrize("idx", range(len(KAT["texture_query"]["cases"])))
def test_texture_query(oracle_lib, textures, idx):
    """buffer/texture_test.go:33-166: Texture.Query incl. fractional LODs. This pins the oracle's
    Query/queryBilinear/queryTrilinear/LerpC AND the host-side imageutil.Resize mip chain."""
    c = KAT["texture_query"]["cases"][idx]
    got = _query(oracle_lib, textures[c["texture"]], c["lod"], c["u"], c["v"])
    if c["texture"] == "pic.jpg" and got != c["want"]:
        pytest.xfail(f"JPEG decoders differ (PIL vs Go image/jpeg IDCT): got {got}, reference {c['want']}")
    assert got == c["want"], (c, got)


def test_aabb(oracle_lib):
    k = KAT["aabb"]
    a1 = f32a(*k["aabb1"])
    for name, want in k["intersect"].items():
        assert bool(oracle_lib.orc_aabb_intersect(fptr(a1), fptr(f32a(*k[name])))) == want
    for p in k["contains_true"]:
        assert oracle_lib.orc_aabb_contains(fptr(a1), fptr(f32a(*p)))
    for p in k["contains_false"]:
        assert not oracle_lib.orc_aabb_contains(fptr(a1), fptr(f32a(*p)))
    # the Z quirk (box.go:38): maxZ compares against the RECEIVER's Max.Y, so a box far beyond +Z still "intersects"
    vp = f32a(0, 0, -1, 800, 500, 1)
    assert oracle_lib.orc_aabb_intersect(fptr(vp), fptr(f32a(10, 10, 5, 20, 20, 6)))
    assert not oracle_lib.orc_aabb_intersect(fptr(vp), fptr(f32a(10, 10, -6, 20, 20, -5)))


def test_triangle_is_valid(oracle_lib):
    for c in KAT["triangle_is_valid"]["cases"]:
        assert bool(oracle_lib.orc_triangle_is_valid(fptr(f32a(*c["p"])))) == c["valid"]
    assert not oracle_lib.orc_triangle_is_valid(fptr(f32a(1, 1, 1, 1, 1, 1, 2, 2, 2)))  # zero-length edge


def test_interp_world_pos(oracle_lib):
    k = KAT["interp_world_pos"]
    out = np.zeros(4, np.float32)
    third = np.float32(1.0) / np.float32(3)
    oracle_lib.orc_interp_world_pos(fptr(f32a(third, third, third)), fptr(f32a(*k["m1"])), fptr(f32a(*k["m2"])), fptr(f32a(*k["m3"])), fptr(out))
    assert 9.99 < float(out[0] + out[1] + out[2]) < 10.01
    oracle_lib.orc_interp_world_pos(fptr(f32a(1, 0, 0)), fptr(f32a(*k["m1"])), fptr(f32a(*k["m2"])), fptr(f32a(*k["m3"])), fptr(out))
    assert out[:3].tolist() == k["m1"][:3]


def test_ao_constants_are_go_constant_arithmetic(oracle_lib):
    """material/ao.go:28-32: Go evaluates untyped constant expressions exactly and rounds ONCE to float32."""
    out = np.zeros(5, np.float32)
    oracle_lib.orc_ao_constants(fptr(out))
    pi = Fraction("3.14159265358979323846264338327950288419716939937510582097494459")

    def f32_of(fr):  # correctly rounded float32 of an exact rational
        x = np.float32(float(fr))
        cands = [np.nextafter(x, np.float32(-np.inf)), x, np.nextafter(x, np.float32(np.inf))]
        return min(cands, key=lambda c: abs(Fraction(float(c)) - fr))

    assert out[0] == f32_of(pi) and out[1] == f32_of(pi / 4) and out[2] == f32_of(pi / 2) and out[3] == f32_of(pi * 4)
    assert out[4] == f32_of(pi * 2 - Fraction("0.0001"))


def test_srgb_lut_and_gamma_table():
    """color/color_test.go:36-66 + shader/gamma.go:13-18."""
    assert imageutil.linear_to_srgb(np.float32(0)) == 0 and imageutil.linear_to_srgb(np.float32(1)) == 1
    assert imageutil.srgb_to_linear(np.float32(0)) == 0 and imageutil.srgb_to_linear(np.float32(1)) == 1
    v = imageutil.linear_to_srgb(np.float32(0.5))
    assert abs(float(imageutil.srgb_to_linear(v)) - 0.5) <= 1e-7 + 1e-6  # LUT round trip (ApproxEq in the reference; LUT interpolation error)
    lut = imageutil.gamma_lut_u8()
    assert lut[0] == 0 and lut[255] == 255 and np.all(np.diff(lut.astype(int)) >= 0)
    # analytic sRGB of the mid grey, +-1 (kernels/srgb_test.go:14 tolerance 1e-6 on the float value)
    assert abs(int(lut[128]) - round(255 * (1.055 * (128 / 255) ** (1 / 2.4) - 0.055))) <= 1


def _kernels_shade_float64(k, n, p, base):
    """gpu/shader/gpumath/kernels/shade.go:15-56 restated in float64 (the reference's own second opinion)."""
    N, x, col = np.array(n + [0.0]), np.array(p + [1.0]), np.array(base, float)
    cam = np.array(k["cam"] + [1.0])
    acc = col * k["ambient"]
    for l in k["lights"]:
        lc = np.array(l["color"], float)
        if l["kind"] == "point":
            Ld = np.array(l["pos"] + [1.0]) - x
            L, I = Ld / np.linalg.norm(Ld), l["intensity"] / np.linalg.norm(Ld)
        else:
            d = np.array(l["dir"], float)
            d = d / np.linalg.norm(d)
            L, I = np.array([-d[0], -d[1], -d[2], 0.0]), l["intensity"]
        V = (cam - x) / np.linalg.norm(cam - x)
        H = (L + V) / np.linalg.norm(L + V)
        Ld_, Ls_ = min(max(N @ L, 0), 1), min(max(N @ H, 0), 1) ** k["shininess"]
        acc = acc + np.array(k["diffuse"], float) * (col * Ld_ * I) / 255.0 + np.array(k["specular"], float) * (lc * Ls_ * I) / 255.0
    return [int(min(max(math.floor(abs(v) + 0.5) * (1 if v >= 0 else -1), 0), 255)) for v in acc[:3]]


def test_fragment_shader_equivalence(oracle_lib):
    """render/shading_equiv_test.go:33-107: FragmentShader agrees with kernels.Shade within 1 LSB."""
    import oracle_binding as ob
    from polyred_b200 import light, render, scene
    k = KAT["shading_equivalence"]
    tex = material.Texture.uniform(k["texture_rgba"])
    mat = material.BlinnPhong(texture=tex, diffuse=k["diffuse"], specular=k["specular"], shininess=k["shininess"])
    l0, l1 = k["lights"]
    s = scene.Scene(light.Point(intensity=l0["intensity"], color=l0["color"], position=l0["pos"]),
                    light.Directional(intensity=l1["intensity"], color=l1["color"], direction=l1["dir"]), light.Ambient(intensity=k["ambient"]))
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    s.Add(scene.Geometry(tri, materials=[mat]))
    be = ob.OracleBackend()
    r = render.NewRenderer(render.Camera(camera.Perspective(position=k["cam"])), render.Size(8, 8), render.Scene(s), render._Backend(be))
    r._ensure_uploaded()
    fd = r.frame_desc()
    for n, p in zip(k["normals"], k["positions"]):
        got = oracle_lib.orc_fragment_shader(be.h, C.byref(fd.struct), 0, fptr(f32a(*n)), fptr(f32a(*n)), fptr(f32a(*p)), fptr(f32a(0.5, 0.5, 0, 0)), A.pack_rgba(k["texture_rgba"]))
        cpu = [got & 255, (got >> 8) & 255, (got >> 16) & 255]
        ref = _kernels_shade_float64(k, n, p, k["texture_rgba"])
        assert max(abs(a - b) for a, b in zip(cpu, ref)) <= 1, (n, p, cpu, ref)
        assert (got >> 24) == 255


# ---- more of the reference's own known-answer tests, transcribed with their expected values ----
def test_transform_context_model_matrices():
    """math/context_test.go:13-82 (TestTransformationContext): scale, translate, re-scale, quarter turns about Y / X / Z."""
    ctx = gm.TransformContext()
    ctx.Scale(1, 2, 3)
    ctx.Translate(1, 2, 3)
    assert mat_eq(ctx.ModelMatrix(), gm.mat4(1, 0, 0, 1, 0, 2, 0, 2, 0, 0, 3, 3, 0, 0, 0, 1))
    ctx.Scale(1, 2, 3)
    assert mat_eq(ctx.ModelMatrix(), gm.mat4(1, 0, 0, 1, 0, 4, 0, 4, 0, 0, 9, 9, 0, 0, 0, 1))
    half_pi = np.float32(math.pi / 2)
    for axis, want in (((0, 1, 0), (0, 0, 1, 0, 0, 1, 0, 0, -1, 0, 0, 0, 0, 0, 0, 1)),
                       ((1, 0, 0), (1, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1)),
                       ((0, 0, 1), (0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1))):
        ctx.ResetContext()
        ctx.Rotate(gm.v3(*axis), half_pi)
        assert mat_eq(ctx.ModelMatrix(), gm.mat4(*want)), axis


def test_quaternion_to_rotation_matrix():
    """math/quaternion_test.go:13-70 (TestQuaternionToRotationMatrix): pi/3 about X, Y, Z."""
    angle = np.float32(math.pi) / np.float32(3)
    c, s = np.float32(math.cos(float(angle * np.float32(0.5)))), np.float32(math.sin(float(angle * np.float32(0.5))))
    h = np.float32(0.8660254)
    for u, want in (((1, 0, 0), (1, 0, 0, 0, 0, 0.5, -h, 0, 0, h, 0.5, 0, 0, 0, 0, 1)),
                    ((0, 1, 0), (0.5, 0, h, 0, 0, 1, 0, 0, -h, 0, 0.5, 0, 0, 0, 0, 1)),
                    ((0, 0, 1), (0.5, -h, 0, 0, h, 0.5, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1))):
        q = gm.Quaternion(c, s * np.float32(u[0]), s * np.float32(u[1]), s * np.float32(u[2]))
        assert mat_eq(q.to_romat(), gm.mat4(*want)), u


def test_vec4_apply_dot_cross_unit_oracle_and_host(oracle_lib):
    """math/vec_test.go: TestVec_Apply (:383-396), TestVec_Dot (:180-189), TestVec_Cross (:409-418), TestVec_Len/Unit
    (:303-316, 337-351) for Vec4 - the operations every kernel is built from - on the oracle's restatement and the host mirror."""
    L = oracle_lib
    out4 = (C.c_float * 4)()
    m = f32a(1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16)
    L.orc_vec4_apply(fptr(f32a(1, 1, 1, 1)), fptr(m), out4)
    assert list(out4) == [10, 26, 42, 58]
    assert list(gm.v4_apply(np.array([1, 1, 1, 1], np.float32), gm.mat4(*range(1, 17)))) == [10, 26, 42, 58]
    d = C.c_float()
    L.orc_vec4_dot(fptr(f32a(1, 1, 2, 5)), fptr(f32a(2, 2, 2, 5)), C.byref(d))
    assert d.value == 33.0
    L.orc_vec4_cross(fptr(f32a(1, 0, 0, 0)), fptr(f32a(0, 1, 0, 0)), out4)
    assert list(out4) == [0, 0, 1, 0]
    assert list(gm.v3_cross(gm.v3(1, 0, 0), gm.v3(0, 1, 0))) == [0, 0, 1]
    L.orc_vec4_unit(fptr(f32a(1, 1, 1, 0)), out4)
    r3 = np.float32(1) / np.float32(math.sqrt(3.0))
    assert mat_eq(list(out4), [r3, r3, r3, 0])
    L.orc_vec4_unit(fptr(f32a(1, 1, 1, 1)), out4)
    assert mat_eq(list(out4), [0.5, 0.5, 0.5, 0.5])
    assert mat_eq(gm.v3_unit(gm.v3(1, 1, 1)), [r3, r3, r3])
