"""-m gpu: device groups behind the C ABI (include/polyred_cuda.h prc_group_*, csrc/prc_group.cpp) and view batches
(prc_render_batch). The reference has no multi-device path (gpu/device.go:77-89 opens one device): the oracle of a group frame
is the frame of ONE context, bit for bit — and that frame is compared with the CPU oracle by the other -m gpu tests.

A group may list one device several times (several contexts of one GPU), so the whole protocol — export, connect, the library's
own submit threads, the strip partition and its re-balancing, the retry vote, the double-buffered host image — runs on a single
B200; with two or more devices visible the same checks also run across devices (NVLink peer access)."""
import ctypes as C
import os

import numpy as np
import pytest

from polyred_b200 import _abi as A
from polyred_b200 import render, synth

pytestmark = pytest.mark.gpu


def _n_devices():
    from polyred_b200 import _lib
    return int(_lib.lib().prc_device_count())


def _city(w=480, h=272):
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    return s, cam, w, h


def _opts(s, cam, w, h, msaa=1):
    return [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.MSAA(msaa)]


def _device_sets():
    sets = [(0,), (0, 0), (0, 0, 0)]
    n = _n_devices()
    if n >= 2:
        sets.append((0, 1))
    if n >= 4:
        sets.append((0, 1, 2, 3))
    if n >= 8:
        sets.append(tuple(range(8)))
    return sets


def test_group_rank_that_submits_late_does_not_stall_the_group(monkeypatch):
    """Kept FIRST in this file: it should be the first peer frame of the process. Rank 1 submits every frame 40 ms after rank 0
    (PRC_GROUP_STAGGER_MS), so rank 0's one-warp wait kernel is already spinning on the shared GPU when rank 1 launches each of
    its kernels for the first time. With CUDA's lazy module loading such a first launch may synchronise the whole context - i.e.
    wait for the spinning kernel, which waits for rank 1: the wait then gave up after 4 s and the frame failed with PRC_ERR_PEER
    (round 2, first seen on an 8-GPU box). prc_peer_connect now loads every kernel of the library before the first frame."""
    monkeypatch.setenv("PRC_GROUP_STAGGER_MS", "40")
    s, cam, w, h = _city()
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0))
    for k in range(4):
        assert np.array_equal(r.Render(), ref), f"group frame {k + 1} differs"
    r._backend.close()


@pytest.mark.parametrize("devices", _device_sets() if os.environ.get("PRC_TEST_GROUP", "1") != "0" else [])
def test_group_frame_equals_the_one_context_frame(devices):
    s, cam, w, h = _city()
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(*devices)) if len(devices) > 1 else None
    if r is None:
        from polyred_b200._lib import GroupBackend
        r = render.NewRenderer(*_opts(s, cam, w, h), render._Backend(GroupBackend(devices)))
    be = r._backend
    assert be.size() == len(devices)
    first = r.Render()
    assert np.array_equal(first, ref), f"first group frame differs in {(np.abs(first.astype(int) - ref.astype(int)).max(axis=2) > 0).sum()} px"
    keep = first  # read in place: must stay valid during the NEXT Render() (the reference's double buffer, raster.go:201-206)
    strips_seen = set()
    for k in range(12):  # the first frames after connecting re-balance the strips: every partition must give the same frame
        img = r.Render()
        strips_seen.add(tuple(be.strips()))
        assert np.array_equal(img, ref), f"group frame {k + 2} differs (strips {be.strips()})"
        if k == 0:
            assert np.array_equal(keep, ref), "the previous frame's in-place image was overwritten by the next Render()"
    rows = be.strips()
    assert rows[0][1] == h and rows[-1][0] == 0 and all(a[0] == b[1] for a, b in zip(rows, rows[1:])), rows  # rank 0 = top image rows
    # every rank holds the merged shadow maps
    sources, _ = s.Lights()
    cast = [i for i, l in enumerate(sources) if l.cast_shadow]
    one = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0))
    one.Render()
    for k in range(len(devices)):
        for li in cast[:2]:
            assert np.array_equal(be.rank(k).read_shadowmap(li, w, h), one._backend.read_shadowmap(li, w, h)), (k, li)
    be.close()


def test_group_frames_left_on_the_device_and_submitted_back_to_back():
    s, cam, w, h = _city()
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0))
    be = r._backend
    r._ensure_uploaded()
    fd = r.frame_desc(no_readback=True)
    from polyred_b200._lib import PolyredCudaError
    for attempt in range(4):
        fd.struct.flags |= A.PRC_FRAME_ASYNC
        for _ in range(5):
            be.render(fd, None)
        try:
            be.sync()
            break
        except PolyredCudaError as e:  # a queue grew: submit again (same contract as prc_sync)
            assert e.code == A.PRC_ERR_RETRY
    img = be.rank(0).read_image(w, h)  # the strips were gathered into context 0's device image
    assert np.array_equal(img, ref)
    # a synchronous device-resident frame
    fd.struct.flags &= ~A.PRC_FRAME_ASYNC
    be.render(fd, None)
    assert np.array_equal(be.rank(0).read_image(w, h), ref)
    be.close()


def test_group_with_ambient_occlusion_and_with_msaa():
    # AO: a rank needs the depths of 100 rows around its strip and of rows 0..99 for pixel (0,0) (bug-list 3)
    s, cam = synth.mesh_scene(subdiv=30, with_ground=True, shadows=True, ao=True)
    w, h = 320, 416
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0, 0))
    for _ in range(3):
        assert np.array_equal(r.Render(), ref)
    r._backend.close()
    # MSAA(2): every rank shades a halo of msaa rows and downsamples its own output rows
    s, cam, w, h = _city(240, 136)
    ref = render.NewRenderer(*_opts(s, cam, w, h, msaa=2), render.CUDA(0)).Render().copy()
    r = render.NewRenderer(*_opts(s, cam, w, h, msaa=2), render.CUDA(0, 0))
    for _ in range(3):
        assert np.array_equal(r.Render(), ref)
    r._backend.close()


def test_group_follows_a_moving_camera_and_a_new_frame_size():
    """Options() between frames: new camera = light cameras re-fitted and shadow maps zeroed on EVERY rank
    (prc_group_shadow_reset); a new size = the group exports and connects again by itself."""
    s, cam, w, h = _city()
    one = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0))
    grp = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0))
    assert np.array_equal(grp.Render(), one.Render())
    cam2 = synth.orbit_camera(1.0, aspect=w / h)
    for r in (one, grp):
        r.Options(render.Camera(cam2))
    assert np.array_equal(grp.Render(), one.Render())
    for r in (one, grp):
        r.Options(render.Size(320, 200))
    assert np.array_equal(grp.Render(), one.Render())
    grp._backend.close()


def test_view_batches_equal_one_view_at_a_time():
    """prc_render_batch / prc_group_render_views (BASELINE configs[4]) against Options(Camera) + Render() per view."""
    s, cam, w, h = _city(320, 180)
    cams = [synth.orbit_camera(2 * np.pi * k / 5, aspect=w / h) for k in range(5)]
    one = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0))
    want = []
    for c in cams:  # the literal loop a reference caller writes
        sd = one._scene_desc
        one.Options(render.Camera(c))
        one._scene_desc = sd
        want.append(one.Render().copy())
    got = render.RenderViews(render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)), cams)  # prc_render_batch
    assert len(got) == len(want)
    for v, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"view {v} of the batch differs"
    grp = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0))
    got = render.RenderViews(grp, cams)  # prc_group_render_views: view v on context v mod 2
    for v, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), f"view {v} of the group batch differs"
    # and a whole-group frame after a batch (the group reconnects by itself)
    grp.Options(render.Camera(cam))
    one.Options(render.Camera(cam))
    assert np.array_equal(grp.Render(), one.Render())
    grp._backend.close()


def test_group_rejects_what_it_cannot_do():
    from polyred_b200._lib import GroupBackend, PolyredCudaError
    s, cam, w, h = _city(64, 2)
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0, 0))
    with pytest.raises(PolyredCudaError) as e:  # 2 rows cannot be cut into 3 strips: an error, not a hang
        r.Render()
    assert e.value.code == A.PRC_ERR_INVALID
    r._backend.close()
    s, cam, w, h = _city(160, 96)
    r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0, 0))
    r._ensure_uploaded()
    fd = r.frame_desc(keep_gbuffer=True)
    with pytest.raises(PolyredCudaError) as e:
        r._backend.render(fd, None)
    assert e.value.code == A.PRC_ERR_UNSUPPORTED
    fd = r.frame_desc()
    fd.struct.row0 = 8
    with pytest.raises(PolyredCudaError) as e:
        r._backend.render(fd, None)
    assert e.value.code == A.PRC_ERR_INVALID
    r._backend.close()
    with pytest.raises(PolyredCudaError):
        GroupBackend([0, 4096])  # no such device
