"""-m gpu: frames over NVLink peer memory (prc_render_peer) must equal the 1-GPU frame bit for bit.

Run on one B200 at the end of round 1 as tools/peer_selfcheck.py (profiles/peer_selfcheck_r1.log: 1, 2 and 3 ranks on one
GPU, 0 differing pixels, 0 differing shadow texels); the multi-process form of the same check (CUDA IPC, one process per GPU) is
the peer leg of tests/multigpu_check.py. PRC_TEST_PEER=0 skips this file.

  world = 1 : the whole protocol with no peer (waits and signals are skipped), 1 GPU
  world = 2 : two contexts of ONE process, on one GPU and on two GPUs (the library uses the peers' pointers directly instead of
              IPC handles); frames are submitted to both contexts from this single host thread, which only works
              because nothing in prc_render_peer waits on the host
"""
import os

import numpy as np
import pytest

from polyred_b200 import _abi as A
from polyred_b200 import partition, render, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("PRC_TEST_PEER") == "0", reason="PRC_TEST_PEER=0")]


def _n_devices():
    from polyred_b200 import _lib
    return int(_lib.lib().prc_device_count())


def _scene(w=480, h=272):
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    return s, cam, w, h


def _opts(s, cam, w, h):
    return [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]


def _group(s, cam, w, h, devices):
    world = len(devices)
    _, rows = partition.strips(h, world)
    rs, fds, handles = [], [], []
    for k, dev in enumerate(devices):
        r = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(dev))
        r._ensure_uploaded()
        fd = r.frame_desc(no_readback=True)
        fd.struct.row0, fd.struct.row1 = rows[k]
        handles.append(r._backend.peer_export(fd))
        rs.append(r)
        fds.append(fd)
    for k, r in enumerate(rs):
        r._backend.peer_connect(k, world, handles)
    mine = [rows] * world  # every call carries the strips of all ranks
    return rs, fds, mine


def _run(devices, frames=3, size=(480, 272), scene=None, extra_flags=0):
    s, cam, w, h = _scene(*size) if scene is None else (*scene, *size)
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(devices[0])).Render().copy()
    from polyred_b200._lib import PolyredCudaError
    rs, fds, mine = _group(s, cam, w, h, devices)
    for fd in fds:
        fd.struct.flags |= extra_flags
    for attempt in range(4):
        for _ in range(frames):
            for k, r in enumerate(rs):
                r._backend.render_peer(fds[k], mine[k], 1)
        again = False
        for r in rs:
            try:
                r._backend.sync()
            except PolyredCudaError as e:  # a queue grew / the binned tile path or NaN mode was switched on: every rank submits again
                assert e.code == A.PRC_ERR_RETRY
                again = True
        if not again:
            break
        state = 0
        for r in rs:
            state |= r._backend.frame_state()
        for r in rs:
            r._backend.set_frame_state(state)  # the ranks must agree on NaN mode (distributed.PeerFrames.finish does the same)
    assert not again
    out = rs[0]._backend.read_image(w, h)
    for r in rs:
        r._backend.peer_disconnect()
    return ref, out


def test_peer_world1_equals_render(monkeypatch):
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0])
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_two_contexts_one_gpu(monkeypatch):
    """Two ranks on ONE GPU (two contexts, two streams): the whole protocol — sparse shadow push, epoch signals, strip copy —
    without needing a second device. The one-warp wait kernels leave the GPU free for the other context's kernels."""
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0, 0])
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_three_ranks_ragged_rows_one_gpu(monkeypatch):
    """Strips and shadow shards that do not divide evenly (272 rows over 3 ranks; 4 maps over 3 ranks = units that span two lights)."""
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0, 0, 0], frames=3)
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_image_complete_at_sync(monkeypatch):
    """PRC_FRAME_IMAGE_AT_SYNC: the consumer (rank 0) waits for its peers' strips beside its stream and starts its next frame at
    once; after prc_sync its image holds the last frame, all strips of it. Five frames back to back, three ranks."""
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0, 0, 0], frames=5, extra_flags=A.PRC_FRAME_IMAGE_AT_SYNC)
    assert int((ref != out).any(axis=2).sum()) == 0
    ref, out = _run([0, 0], frames=1, extra_flags=A.PRC_FRAME_IMAGE_AT_SYNC)
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_ao_strips_pixel00_halo(monkeypatch):
    """An AO material over strips much taller than the 100-row AO halo, with pixel (0,0) covered: the colour of every uncovered
    pixel comes from shading pixel (0,0) (bug-list 3), whose AO rays read depths of rows 0..99 — every rank must have rasterised
    and resolved those rows, not only its own strip."""
    monkeypatch.setenv("PRC_FMA", "exact")
    s, cam = synth.mesh_scene(subdiv=24, with_ground=True, shadows=True, ao=True, aspect=240 / 640)
    ref, out = _run([0, 0, 0], frames=2, size=(240, 640), scene=(s, cam))
    assert int((ref != out).any(axis=2).sum()) == 0
    # the case is only meaningful when pixel (0,0) is covered and some pixels are not
    g = render.NewRenderer(*_opts(s, cam, 240, 640), render.CUDA(0))
    g.Render(keep_gbuffer=True)
    ok = g._backend.read_gbuffer(240, 640)["ok"].reshape(640, 240)
    assert ok[0, 0] == 1 and (ok == 0).any(), "test scene no longer exercises the pixel-(0,0) AO quirk"


def test_peer_strips_move_back_and_forth_under_a_moving_camera():
    """Re-balancing moves the strip boundaries between frames while the camera moves: a row that a rank loses and regains two
    frames later must not still hold the visibility keys it received for the OLD camera (the merged key planes alternate with
    the frame parity and are cleared where the previous frame of that parity wrote, not where the current strip lies)."""
    from polyred_b200._lib import PolyredCudaError
    s, cam, w, h = _scene()
    cams = [synth.orbit_camera(0.35 * k, aspect=w / h) for k in range(7)]
    parts = [[(136, 272), (0, 136)], [(48, 272), (0, 48)], [(200, 272), (0, 200)]]
    one = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0))
    one._ensure_uploaded()
    rs, fds, _ = _group(s, cam, w, h, [0, 0])
    for k, c in enumerate(cams):
        rows = parts[k % 2] if k < 4 else parts[(k % 2) * 2]  # A B A B | A C A: every boundary is crossed in both directions
        one.cfg.Camera = c   # (no Options(): the light cameras stay, the persistent maps keep growing - on both sides)
        want = one.Render().copy()
        for attempt in range(4):
            for j, r in enumerate(rs):
                r.cfg.Camera = c
                fd = r.frame_desc(no_readback=True)
                fd.struct.row0, fd.struct.row1 = rows[j]
                r._backend.render_peer(fd, rows, 1)
            again, state = False, 0
            for r in rs:
                try:
                    r._backend.sync()
                except PolyredCudaError as e:
                    assert e.code == A.PRC_ERR_RETRY
                    again = True
                state |= r._backend.frame_state()
            if not again:
                break
            for r in rs:
                r._backend.set_frame_state(state)
        assert not again
        got = rs[0]._backend.read_image(w, h)
        assert int((got != want).any(axis=2).sum()) == 0, f"frame {k} (strips {rows}) differs from the one-context frame"
    for r in rs:
        r._backend.peer_disconnect()


def test_peer_two_contexts_one_process(monkeypatch):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0, 1])
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_rejects_unsupported_frames(monkeypatch):
    from polyred_b200._lib import PolyredCudaError
    s, cam, w, h = _scene()
    rs, fds, mine = _group(s, cam, w, h, [0])
    be = rs[0]._backend
    fds[0].struct.flags |= A.PRC_FRAME_SHADOW_RESET
    with pytest.raises(PolyredCudaError):
        be.render_peer(fds[0], mine[0], 1)
    fds[0].struct.flags &= ~A.PRC_FRAME_SHADOW_RESET
    be.render_peer(fds[0], mine[0], 1)
    be.render_peer(fds[0], mine[0], 0)  # the consumer set may change from frame to frame
    be.sync()
    be.peer_disconnect()
    with pytest.raises(PolyredCudaError):
        be.render_peer(fds[0], mine[0], 1)


def test_peer_odd_frame_size_takes_the_scalar_push(monkeypatch):
    """A frame width that is not a multiple of 4: the shards are not float4-aligned and k_shadow_push<1> runs."""
    monkeypatch.setenv("PRC_FMA", "exact")
    ref, out = _run([0, 0, 0], frames=2, size=(483, 271))
    assert int((ref != out).any(axis=2).sum()) == 0


def test_peer_strip_readback_into_one_host_image(monkeypatch):
    """Every rank DMAs its own strip into ONE host image (prc_set_host_image), no device-side gather (image_mask = 0)."""
    monkeypatch.setenv("PRC_FMA", "exact")
    s, cam, w, h = _scene()
    ref = render.NewRenderer(*_opts(s, cam, w, h), render.CUDA(0)).Render().copy()
    rs, fds, mine = _group(s, cam, w, h, [0, 0])
    host = np.zeros((h, w, 4), np.uint8)
    for k, r in enumerate(rs):
        r._backend.set_host_image(host.ctypes.data, host.nbytes)
        fds[k].struct.flags &= ~A.PRC_FRAME_NO_READBACK
    for _ in range(2):
        for k, r in enumerate(rs):
            r._backend.render_peer(fds[k], mine[k], 0)
    for r in rs:
        r._backend.sync()
    assert int((ref != host).any(axis=2).sum()) == 0
    for r in rs:
        r._backend.set_host_image(None)
        r._backend.peer_disconnect()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_msaa_strips_downsample_locally(monkeypatch, world):
    """render.MSAA(2) over strips: every rank shades `msaa` supersampled rows beyond its strip, runs imageutil.Resize on its own
    output rows and reads them back into the one host image; no exchange. Must equal the 1-GPU MSAA frame byte for byte."""
    monkeypatch.setenv("PRC_FMA", "exact")
    s, cam, w, h = _scene(240, 136)
    opts = _opts(s, cam, w, h) + [render.MSAA(2)]
    ref = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
    rows = [(a * 2, b * 2) for a, b in partition.strips_from_bounds(h, partition.equal_bounds(h, world))]
    host = np.zeros((h, w, 4), np.uint8)
    rs, fds = [], []
    handles = []
    for k in range(world):
        r = render.NewRenderer(*opts, render.CUDA(0))
        r._ensure_uploaded()
        fd = r.frame_desc(no_readback=False)
        fd.struct.row0, fd.struct.row1 = rows[k]
        handles.append(r._backend.peer_export(fd))
        rs.append(r)
        fds.append(fd)
    for k, r in enumerate(rs):
        r._backend.peer_connect(k, world, handles)
        r._backend.set_host_image(host.ctypes.data, host.nbytes)
    for _ in range(2):
        for k, r in enumerate(rs):
            r._backend.render_peer(fds[k], rows, 0)
    for r in rs:
        r._backend.sync()
    assert int((ref != host).any(axis=2).sum()) == 0
    for r in rs:
        r._backend.set_host_image(None)
        r._backend.peer_disconnect()



def test_multi_process_frames_equal_the_one_gpu_frame():
    """tests/multigpu_check.py under torchrun, one process per GPU (CUDA IPC + NVLink peer memory, and the NCCL path of round 1):
    every N-GPU frame — equal strips, re-balanced strips, strips read back into one shared host image, a moving camera with
    alternating partitions — must equal rank 0's own 1-GPU frame bit for bit. Needs two GPUs (the driver's test box has one;
    the logs of the 2- and 8-GPU runs are profiles/multigpu_check_{2,8}_r2.log)."""
    import subprocess
    import sys
    n = _n_devices()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29400 + os.getpid() % 500), os.path.join(root, "tests", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=root)
    if os.environ.get("PRC_MULTIGPU_LOG"):  # keep the run's output (tools/r2_call_j.sh -> profiles/)
        with open(os.environ["PRC_MULTIGPU_LOG"], "w") as f:
            f.write(p.stdout + p.stderr)
    lines = [l for l in p.stdout.splitlines() if l.startswith("[multigpu_check]") and "pixels differing" in l]
    assert p.returncode == 0 and lines and all(l.rstrip().endswith("= 0") for l in lines), p.stdout[-3000:] + p.stderr[-3000:]
