"""-m gpu: CUDA path vs the CPU oracle on identical inputs, through the C ABI (libpolyred_cuda.so)."""
import math
import os

import numpy as np
import pytest

from polyred_b200 import camera, light, material, render, scene, synth
from parity_util import assert_bit_exact, assert_mixed, assert_north_star_gate, compare_frames, make_renderers

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _exact(monkeypatch):
    monkeypatch.setenv("PRC_FMA", "exact")


def _report(name, st):
    print(f"\n[{name}] " + " ".join(f"{k}={v}" for k, v in st.items()))


def test_c1_mesh_point_light():
    """BASELINE config 1 (reduced mesh): one closed mesh, point light + ambient, no shadows."""
    s, cam = synth.mesh_scene(subdiv=60, aspect=1.6)
    g, c = make_renderers(s, cam, 400, 250)
    st, ig, ic = compare_frames(g, c, 400, 250)
    _report("c1", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st


def test_c2_shadow_ao_gamma_clipped_ground():
    """BASELINE config 2 (reduced): mesh + ground quad that straddles the viewport (clip path),
    non-casting directional + casting point light, shadow map, AO, gamma."""
    s, cam = synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=True)
    g, c = make_renderers(s, cam, 320, 200, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 320, 200, n_lights_cast=(1,))
    _report("c2", st)
    assert_bit_exact(st)
    assert st["shadow1_written"] > 0
    # AO goes through atan/pow(.,10000) in float64 on both sides (libm vs CUDA): allow 1 LSB
    assert_north_star_gate(st)


def test_c3_city_eight_lights_four_casting():
    """BASELINE config 3 (reduced): ground heightfield + instanced meshes, 8 materials, 8 point
    lights of which 4 cast shadows; micro-triangles dominate."""
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    g, c = make_renderers(s, cam, 480, 270, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 480, 270, n_lights_cast=(0, 2, 4, 6))
    _report("c3", st)
    assert_bit_exact(st)
    assert_north_star_gate(st)
    assert st["rgba_px_diff"] == 0, st


def test_persistent_shadow_maps_and_second_frame():
    """Shadow maps are never cleared by Render() (render/shadow.go:221-228): a second frame of a
    static scene is identical, and prc_shadow_reset zeroes them."""
    s, cam = synth.city_scene(n_objects=9, obj_stacks=10, obj_slices=10, ground_cells=20, tex_size=32)
    g, c = make_renderers(s, cam, 240, 135, shadow=True)
    a = g.Render()
    b = g.Render()
    assert np.array_equal(a, b)
    assert np.array_equal(a, c.Render())


def test_default_mixed_mode_gbuffer_exact_rgba_within_gate(monkeypatch):
    """Default PRC_FMA=mixed: coverage, triangle ids, depth, UV and the shadow maps use the exact float64-FMA
    emulation (bit-exact asserted); shading-only attributes and the deferred shading use fmaf (a few ulp /
    north_star RGBA gate asserted, differences reported)."""
    monkeypatch.delenv("PRC_FMA", raising=False)
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    g, c = make_renderers(s, cam, 480, 270, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 480, 270, n_lights_cast=(0, 2, 4, 6))
    _report("c3-mixed", st)
    assert_mixed(st)
    assert_north_star_gate(st)
    s, cam = synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=True)
    g, c = make_renderers(s, cam, 320, 200, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 320, 200, n_lights_cast=(1,))
    _report("c2-mixed", st)
    assert_mixed(st)
    assert_north_star_gate(st)


def test_fast_fma_mode_within_gate(monkeypatch):
    """PRC_FMA=fast uses single-rounding fmaf where the reference uses a float64 FMA rounded to
    float32: the north_star gate (not bit-exactness) is asserted and the tie count reported."""
    monkeypatch.setenv("PRC_FMA", "fast")
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    g, c = make_renderers(s, cam, 480, 270, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 480, 270, n_lights_cast=(0, 2, 4, 6))
    _report("c3-fast", st)
    assert_north_star_gate(st, tie_budget=8)


def test_c4_no_shadow_single_light_large_frame():
    """BASELINE config 4 flavour (reduced): one point light + ambient, no shadows/AO, larger frame so that the
    ground triangles take the tile path while the instanced meshes stay on the in-thread path."""
    s, cam = synth.city_scene(n_objects=36, obj_stacks=24, obj_slices=24, ground_cells=12, n_lights=1, casting_every=0, tex_size=64, receive_shadow=False)
    g, c = make_renderers(s, cam, 1280, 720)
    st, ig, ic = compare_frames(g, c, 1280, 720)
    _report("c4", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st
    assert g._backend.timings().n_large_items > 0


def test_c5_multi_view_shadows_refit_per_view():
    """BASELINE config 5 flavour (reduced): the same scene from 3 orbit cameras; the light cameras are re-fitted
    and the shadow maps zeroed per view (bug-list 5), the scene stays resident."""
    import oracle_binding as ob
    from polyred_b200._lib import CudaBackend
    cams = []
    for k in range(3):
        s, cam = synth.city_scene(n_objects=16, obj_stacks=12, obj_slices=12, ground_cells=30, tex_size=32, cam_angle=2 * math.pi * k / 3, cam_radius=2.6, cam_height=1.1)
        cams.append(cam)
    opts = [render.Camera(cams[0]), render.Size(320, 180), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    rg = render.NewRenderer(*opts, render.CUDA(0))
    rc = render.NewRenderer(*opts, render._Backend(ob.OracleBackend()))
    vg, vc = render.RenderViews(rg, cams), render.RenderViews(rc, cams)
    for a, b in zip(vg, vc):
        assert np.array_equal(a, b)
    assert not np.array_equal(vg[0], vg[1])


def test_c5_view_frames_submitted_back_to_back():
    """ViewFrames(): the same views as RenderViews(), submitted without a host sync in between (PRC_FRAME_SHADOW_RESET
    zeroes the shadow maps in stream order); CUDA and oracle both honour the flag."""
    import oracle_binding as ob
    cams = []
    for k in range(3):
        s, cam = synth.city_scene(n_objects=16, obj_stacks=12, obj_slices=12, ground_cells=30, tex_size=32, cam_angle=2 * math.pi * k / 3, cam_radius=2.6, cam_height=1.1)
        cams.append(cam)
    opts = [render.Camera(cams[0]), render.Size(320, 180), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    rg = render.NewRenderer(*opts, render.CUDA(0))
    want = render.RenderViews(rg, cams)
    rg2 = render.NewRenderer(*opts, render.CUDA(0))
    rg2._ensure_uploaded()
    rc = render.NewRenderer(*opts, render._Backend(ob.OracleBackend()))
    rc._ensure_uploaded()
    for fd, fc, w in zip(render.ViewFrames(rg2, cams), render.ViewFrames(rc, cams), want):
        rg2._backend.render(fd, None)
        assert np.array_equal(rg2._backend.host_image(320, 180), w)
        out = np.zeros((180, 320, 4), np.uint8)
        rc._backend.render(fc, out)
        assert np.array_equal(out, w)


@pytest.mark.parametrize("mode", ["exact", "mixed", "fast"])
def test_fused_resolve_shade_equals_two_kernel_path(monkeypatch, mode):
    """Without PRC_FRAME_KEEP_GBUFFER (and without AO materials) resolve and shading run as ONE kernel that never writes
    the G-buffer (k_resolve_shade); the frame must be byte-identical to the two-kernel path, and in exact mode to the oracle."""
    import oracle_binding as ob
    monkeypatch.setenv("PRC_FMA", mode)
    s, cam = synth.city_scene(n_objects=36, obj_stacks=16, obj_slices=16, ground_cells=60, tex_size=64)
    opts = [render.Camera(cam), render.Size(640, 360), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    rg = render.NewRenderer(*opts, render.CUDA(0))
    fused = rg.Render().copy()
    two = rg.Render(keep_gbuffer=True).copy()
    assert np.array_equal(fused, two)
    if mode == "exact":
        want = render.NewRenderer(*opts, render._Backend(ob.OracleBackend())).Render()
        assert np.array_equal(fused, want)


@pytest.mark.parametrize("msaa", [2, 3])
def test_msaa_supersample_and_resize(monkeypatch, msaa):
    """render.MSAA(n) = frame buffer n times render.Size, cull / clip box n times larger AGAIN (the
    double-MSAA quirk: triangles up to n screens away are neither culled nor clipped), then imageutil.Resize (two-pass
    fixed-point bilinear, k_resize) down to render.Size. The supersampled G-buffer and shadow maps and the downsampled
    frame are bit-identical to the oracle's; the camera is close enough that part of the scene lies outside the screen."""
    monkeypatch.setenv("PRC_FMA", "exact")
    s, cam = synth.city_scene(n_objects=25, obj_stacks=14, obj_slices=14, ground_cells=40, tex_size=64, cam_radius=1.4, cam_height=0.6)
    w, h = 320, 180
    g, c = make_renderers(s, cam, w, h, shadow=True, gamma=True, msaa=msaa)
    st, ig, ic = compare_frames(g, c, w * msaa, h * msaa, n_lights_cast=(0, 2, 4, 6))
    _report(f"msaa{msaa}", st)
    assert ig.shape == (h, w, 4) and ic.shape == (h, w, 4)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st
    # the fused (no G-buffer) path gives the same downsampled frame
    assert np.array_equal(g.Render(), ig)


def test_pixel_format_bgra():
    """PRC_FRAME_BGRA on the CUDA path (fused and two-kernel shading, with and without MSAA) = the RGBA frame with red
    and blue swapped, as in the oracle."""
    s, cam = synth.city_scene(n_objects=9, obj_stacks=10, obj_slices=10, ground_cells=20, tex_size=32)
    for msaa in (1, 2):
        opts = [render.Camera(cam), render.Size(320, 180), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.MSAA(msaa)]
        rgba = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
        rb = render.NewRenderer(*opts, render.PixelFormat(1), render.CUDA(0))
        assert np.array_equal(rb.Render(), rgba[..., [2, 1, 0, 3]])
        assert np.array_equal(rb.Render(keep_gbuffer=True), rgba[..., [2, 1, 0, 3]])


def test_bunny_msaa2_against_the_reference_render(monkeypatch):
    """The CUDA path against the reference's own published MSAA(2) render (internal/examples/bunny_test.go ->
    examples/out/bunny.png, the fixture test_oracle_golden.py pins the oracle with): alpha identical in every pixel,
    fully covered pixels within 1 LSB — and the whole frame byte-identical to the oracle's."""
    import oracle_binding as ob
    from PIL import Image
    from polyred_b200 import model
    monkeypatch.setenv("PRC_FMA", "exact")
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    s = scene.Scene(light.Point(intensity=200, color=(255, 255, 255, 255), position=(-200, 250, 600)), light.Ambient(intensity=0.7))
    m = model.Load(os.path.join(G, "assets", "bunny_textured", "bunny.obj"))
    m.Scale(1500, 1500, 1500)
    m.Translate(-700, -5, 350)
    s.Add(m)
    cam = camera.Perspective(position=(-550, 194, 734), target=(-1000, 0, 0), up=(0, 1, 1), fov=45, aspect=np.float32(960) / np.float32(540), near=100, far=600)
    opts = [render.Camera(cam), render.Size(960, 540), render.Scene(s), render.MSAA(2), render.ShadowMap(True)]
    img = render.NewRenderer(*opts, render.CUDA(0)).Render()
    gold = np.array(Image.open(os.path.join(G, "ref_renders", "bunny_msaa2.png")))
    assert np.array_equal(img[..., 3], gold[..., 3])
    full = gold[..., 3] == 255
    d = np.abs(img[..., :3].astype(int) - gold[..., :3].astype(int)).max(axis=2)[full]
    assert int(d.max()) <= 1 and int((d > 0).sum()) <= 800
    assert np.array_equal(img, render.NewRenderer(*opts, render._Backend(ob.OracleBackend())).Render())


@pytest.mark.parametrize("mode", ["exact", "mixed"])
def test_shortcuts_equal_the_literal_sequences(monkeypatch, mode):
    """Every arithmetic shortcut of DESIGN.md 4 switched off (plain masks, standard-viewport collapse, affine light
    transforms, structured shadow lookup / unprojection, fused resolve+shade: all vertices then take geom_generic, the
    literal transcription) must give the same frame as the default build of the same FMA mode; in exact mode both equal
    the oracle bit for bit."""
    monkeypatch.setenv("PRC_FMA", mode)
    s, cam = synth.city_scene(n_objects=25, obj_stacks=14, obj_slices=14, ground_cells=40, tex_size=64)
    opts = [render.Camera(cam), render.Size(480, 270), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    fast_path = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
    for k in ("PRC_NO_PLAIN", "PRC_NO_VPSTD", "PRC_NO_AFFINE", "PRC_NO_PERSP_CAM", "PRC_NO_UNPROJ_STD", "PRC_NO_FUSED_SHADE"):
        monkeypatch.setenv(k, "1")
    literal = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
    assert np.array_equal(fast_path, literal)
    if mode == "exact":
        import oracle_binding as ob
        assert np.array_equal(literal, render.NewRenderer(*opts, render._Backend(ob.OracleBackend())).Render())


def test_candidate_queue_overflow_and_unshared_vertices(monkeypatch):
    """k_geom_raster corner cases, bit-exact against the oracle: (a) a ground of ~3-pixel cells seen from above gives every
    256-triangle chunk ~3000 candidate pixels, more than the CTA queue holds (PRC_QCAP = 2048), so part of each chunk takes
    the in-thread fallback; (b) the same mesh with every triangle's vertices jittered apart shares NO vertex (768 distinct
    vertices per chunk, the worst case of the chunk-local vertex table)."""
    monkeypatch.setenv("PRC_FMA", "exact")
    s, _ = synth.city_scene(n_objects=0, ground_cells=150, tex_size=32)
    cam = camera.Perspective(position=(0.0, 2.6, 0.05), target=(0, 0, 0), up=(0, 1, 0), fov=45, aspect=16 / 9, near=0.5, far=6.0)
    g, c = make_renderers(s, cam, 960, 540, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 960, 540, n_lights_cast=(0, 2, 4, 6))
    _report("queue-overflow", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0 and st["covered"] > 100000, st
    geo = [o for o in s.root.leaves() if hasattr(o, "pos")][0]
    rng = np.random.default_rng(7)
    geo.pos = (geo.pos + rng.uniform(-1e-4, 1e-4, size=geo.pos.shape)).astype(np.float32)
    g, c = make_renderers(s, cam, 960, 540, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 960, 540, n_lights_cast=(0, 2, 4, 6))
    _report("unshared-vertices", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st


def test_chunk_culling_is_exact(monkeypatch):
    """Chunk culling (k_chunk_cull) skips 256-triangle chunks that cannot touch a view's rows / the screen. It is
    enabled automatically for partial-row views (multi-GPU); forced here on full frames, including a camera that
    leaves most of the scene off screen and the clipped-ground scene."""
    monkeypatch.setenv("PRC_FORCE_CHUNK_CULL", "1")
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64, cam_radius=1.2, cam_height=0.5)
    g, c = make_renderers(s, cam, 480, 270, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 480, 270, n_lights_cast=(0, 2, 4, 6))
    _report("cull-close", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st
    s, cam = synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=False)
    g, c = make_renderers(s, cam, 320, 200, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 320, 200, n_lights_cast=(1,))
    _report("cull-clip", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st


def test_bin_overflow_rerenders_frame(monkeypatch):
    """The (tile, triangle) bin array is sized optimistically; on overflow the frame is re-rendered
    with a larger array (no host round trip in the common case). Force the path with a tiny array."""
    monkeypatch.setenv("PRC_BINS_INIT", "16")
    s, cam = synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=False)
    g, c = make_renderers(s, cam, 320, 200, shadow=True, gamma=True)
    st, ig, ic = compare_frames(g, c, 320, 200, n_lights_cast=(1,))
    _report("bins-retry", st)
    assert_bit_exact(st)
    assert st["rgba_px_diff"] == 0, st


def test_async_frames(monkeypatch):
    """PRC_FRAME_ASYNC: frames that stay on the device are submitted back to back; prc_sync finishes them (timings summed),
    the frames are the ones a synchronous render gives, and a queue overflow inside the batch surfaces as PRC_ERR_RETRY
    with the queue grown (the resubmitted batch then succeeds)."""
    from polyred_b200 import _abi as A
    from polyred_b200._lib import PolyredCudaError
    s, cam = synth.mesh_scene(subdiv=40, with_ground=True, shadows=True, ao=False)
    opts = [render.Camera(cam), render.Size(320, 200), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
    want = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*opts, render.CUDA(0))
    be = r._backend
    r._ensure_uploaded()
    r.Render()  # one synchronous frame sizes the queues
    per_frame = int(be.timings().gpu_launches)
    fd = r.frame_desc(no_readback=True)
    fd.struct.flags |= A.PRC_FRAME_ASYNC
    for _ in range(3):
        be.render(fd, None)
    be.sync()
    tm = be.timings()
    assert int(tm.gpu_launches) == 3 * per_frame and sum(tm.kernel_launches) > 0
    assert np.array_equal(r.Render(), want)  # a synchronous frame after the batch: same image (shadow maps only grow)
    # overflow inside an asynchronous batch
    monkeypatch.setenv("PRC_BINS_INIT", "16")
    r2 = render.NewRenderer(*opts, render.CUDA(0))
    r2._ensure_uploaded()
    fd2 = r2.frame_desc(no_readback=True)
    fd2.struct.flags |= A.PRC_FRAME_ASYNC
    for attempt in range(6):
        r2._backend.render(fd2, None)
        r2._backend.render(fd2, None)
        try:
            r2._backend.sync()
            break
        except PolyredCudaError as e:
            assert e.code == A.PRC_ERR_RETRY
    else:
        raise AssertionError("the queues never became large enough")
    assert attempt >= 1  # the tiny bin array did overflow at least once
    assert np.array_equal(r2.Render(), want)


def test_errors_surface_no_fallback():
    from polyred_b200._lib import PolyredCudaError, CudaBackend
    from polyred_b200 import _abi as A
    import ctypes as C
    be = CudaBackend(0)
    fr = A.prc_frame(abi_version=A.PRC_ABI_VERSION, width=16, height=16, row1=16)
    with pytest.raises(PolyredCudaError):
        be._check(be.L.prc_render(be.h, C.byref(fr), None))  # no scene uploaded
    be.close()
