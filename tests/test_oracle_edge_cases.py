"""CPU: the oracle on the corners of the input domain (tests/edge_scenes.py). No reference fixture exists for these scenes;
what is asserted are properties the reference's code implies (cited per case) plus determinism — the GPU counterpart
(tests/test_gpu_edge_cases.py) then compares the CUDA path with the oracle bit for bit on the same list."""
import numpy as np
import pytest

import edge_scenes
import oracle_binding as ob
from polyred_b200 import render

SCENES = edge_scenes.edge_scenes()


def _render(name, keep=True):
    s, cam, w, h, o = SCENES[name]
    be = ob.OracleBackend()
    r = render.NewRenderer(*edge_scenes.options(render, s, cam, w, h, o), render._Backend(be))
    img = r.Render(keep_gbuffer=keep).copy()
    return img, be.read_gbuffer(w, h), be, (w, h, o)


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_renders_every_edge_scene_deterministically(name):
    a, ga, _, (w, h, o) = _render(name)
    b, gb, _, _ = _render(name)
    assert a.shape == (h, w, 4) and np.array_equal(a, b)
    for k in ga:
        assert np.array_equal(ga[k], gb[k], equal_nan=True) if ga[k].dtype.kind == "f" else np.array_equal(ga[k], gb[k]), k


def test_empty_and_invalid_scenes_are_all_background():
    """render/raster.go:326-328 (shade returns Background for !Ok) and triangle.go:63-80 (IsValid rejects zero-length edges and
    collinear triangles in model space)."""
    for name in ("empty", "all_invalid"):
        img, g, be, (w, h, o) = _render(name)
        assert not g["ok"].any()
        assert (img == np.array(o["background"], np.uint8)).all()
        assert int(be.timings().n_valid_tris) == 0


def test_nan_depth_fragments_of_a_degenerate_triangle():
    """bug-list 8: a triangle seen exactly edge-on has Sabc = 0; the pixel centres on its line get 0/0 = NaN barycentrics, which
    pass the `< -eps` test (raster.go:491), and their NaN depth wins over an empty pixel (buffer.go:279). The oracle reproduces
    and counts them; the CUDA path counts and drops them (the documented deviation)."""
    _, g, be, _ = _render("nan_depth_degenerate")
    ok = g["ok"].astype(bool)
    assert ok.sum() == 21 and np.isnan(g["depth"][ok]).all() and set(np.nonzero(ok)[1]) == {32}
    assert int(be.timings().n_valid_tris) == 1 and int(be.timings().n_nan_frags) == 21


def test_sliver_of_a_collinear_triangle_that_is_valid_in_float32():
    """triangle.go:63-80: cos = 1.0000001 in float32 is outside IsValid's 1e-7 window, so this collinear triangle is drawn; its
    screen area is a rounding residue, its depths are finite."""
    _, g, be, _ = _render("sliver_collinear")
    ok = g["ok"].astype(bool)
    assert int(be.timings().n_valid_tris) == 1 and ok.sum() == 59 and np.isfinite(g["depth"][ok]).all() and int(be.timings().n_nan_frags) == 0


def test_vertex_colours_pass_through():
    """raster.go:330-333: a negative MaterialID resolves to no material and the interpolated vertex colour is the pixel."""
    img, g, _, (w, h, _) = _render("vertex_colours")
    ok = g["ok"].astype(bool)[::-1]  # G-buffer is in screen coordinates (y up), the image has row 0 on top
    assert ok.sum() > 500
    packed = img[..., 0].astype(np.uint32) | (img[..., 1].astype(np.uint32) << 8) | (img[..., 2].astype(np.uint32) << 16) | (img[..., 3].astype(np.uint32) << 24)
    assert np.array_equal(packed[ok], g["col"][::-1][ok])
    assert (g["mat"][g["ok"].astype(bool)] < 0).all()


def test_flat_shading_uses_one_normal_per_triangle():
    """shader/blinn_cpu.go: FlatShading shades with Fragment.FaceNor, so all pixels of a triangle lit by the ambient + one
    distant-ish light vary only through position; the smooth variant of the same mesh must differ."""
    img, g, _, _ = _render("flat_shading")
    s, cam, w, h, o = SCENES["flat_shading"]
    geo = s.geometries()[0][0]
    geo.materials[0].flat_shading = False
    try:
        be = ob.OracleBackend()
        smooth = render.NewRenderer(*edge_scenes.options(render, s, cam, w, h, o), render._Backend(be)).Render().copy()
    finally:
        geo.materials[0].flat_shading = True
    assert not np.array_equal(img, smooth) and np.array_equal(img[..., 3], smooth[..., 3])


def test_depth_tie_keeps_the_first_triangle_drawn():
    """buffer.go:279 (`z > stored`): of two coincident triangles the one drawn first stays."""
    _, g, _, _ = _render("screen_filling_and_depth_tie")
    ok = g["ok"].astype(bool)
    assert ok.all()                                    # the huge triangles cover every pixel
    assert set(np.unique(g["tri"][ok])) == {0, 2}      # triangle 1 (the duplicate, drawn second) never wins; 2 is the small one in front
    assert (g["sub"][g["tri"] == 0] >= 1).all()        # drawn through clipTriangle's fan (clipping.go:73)


def test_pixel00_quirk_scene():
    img, g, _, (w, h, o) = _render("pixel00_quirk")
    assert g["ok"][0, 0] and not g["ok"][h - 1, w - 1]
    assert tuple(img[0, w - 1]) == (10, 200, 30, 255)  # uncovered, yet shaded from G(0,0) (bug-list 3)


def test_lights_cases():
    """shader/blinn_cpu.go:38-42: without light SOURCES the fragment shader returns the texture colour — ambient terms are not
    applied either, so "no lights" and "ambient only" give the same picture."""
    img, g, _, _ = _render("no_lights")
    ok = g["ok"].astype(bool)[::-1]
    a, _, _, _ = _render("ambient_only_two_terms")
    assert ok.sum() > 100 and np.array_equal(img, a)
    assert (img[ok][:, 3] == 255).all() and (img[ok][:, :3].min(axis=1) >= 80).all()       # the checker texture's range, unlit
    d, _, _, _ = _render("directional_only")
    assert (d[ok][:, :3].max(axis=1) > 0).any() and (d[ok][:, :3].max(axis=1) == 0).any()   # lit side and unlit side


def test_shadow_scenes_write_shadow_maps_and_darken():
    for name in ("two_casters_grouped", "orthographic_shadow"):
        s, cam, w, h, o = SCENES[name]
        img, g, be, _ = _render(name)
        assert (be.read_shadowmap(0, w, h) > 0).sum() > 50 and (be.read_shadowmap(1, w, h) > 0).sum() > 50
        o2 = dict(o, shadow=False)
        be2 = ob.OracleBackend()
        plain = render.NewRenderer(*edge_scenes.options(render, s, cam, w, h, o2), render._Backend(be2)).Render()
        ok = g["ok"].astype(bool)[::-1]
        assert (img[ok][:, :3].astype(int).sum(axis=1) <= plain[ok][:, :3].astype(int).sum(axis=1)).all()   # visibility only halves colours
        assert (img[ok][:, :3].astype(int).sum(axis=1) < plain[ok][:, :3].astype(int).sum(axis=1)).any()


def test_ambient_occlusion_only_darkens_and_shadow_needs_receive_flag():
    """The two property tests the reference has for these stages (gpu/shader/gpumath/kernels/ao_test.go:5-26: with the AO flag
    off the colour is unchanged, with it on every channel ends in [0, original]; kernels/shadow_test.go:5-12: a fragment whose
    material does not receive shadows keeps its colour), replayed on the oracle's CPU restatement of material/ao.go:20-73 and
    render/raster.go:339-353."""
    from polyred_b200 import synth

    def frame(ao, shadows, recv=True):
        s, cam = synth.mesh_scene(subdiv=24, with_ground=True, shadows=True, ao=ao)
        if not recv:
            for g, _ in s.geometries():
                for m in g.materials:
                    m.receive_shadow = False
        be = ob.OracleBackend()
        r = render.NewRenderer(render.Camera(cam), render.Size(160, 100), render.Scene(s), render.ShadowMap(shadows), render._Backend(be))
        return r.Render(keep_gbuffer=True).astype(int), be.read_gbuffer(160, 100)["ok"].astype(bool)[::-1]

    plain, ok = frame(ao=False, shadows=False)
    with_ao, _ = frame(ao=True, shadows=False)
    assert (with_ao[ok][:, :3] <= plain[ok][:, :3]).all() and (with_ao[ok][:, :3] < plain[ok][:, :3]).any()
    assert np.array_equal(with_ao[..., 3], plain[..., 3])                                  # alpha untouched (ao.go:36)
    shadowed, _ = frame(ao=False, shadows=True)
    assert (shadowed[ok][:, :3] <= plain[ok][:, :3]).all() and (shadowed[ok][:, :3] < plain[ok][:, :3]).any()
    not_receiving, _ = frame(ao=False, shadows=True, recv=False)
    assert np.array_equal(not_receiving, plain)                                            # ShadowMap on, ReceiveShadow off: unchanged
