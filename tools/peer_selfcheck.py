#!/usr/bin/env python
"""Quick hardware check of prc_render_peer on ONE GPU, no torch import (starts in a few seconds):
world = 1, then two and three ranks as contexts of this one process on device 0 (tests/test_gpu_peer.py as a script).
    python tools/peer_selfcheck.py [max_world]
Prints one line per case; exit status 0 iff every frame equals the 1-GPU frame bit for bit."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PRC_FMA", "exact")

T0 = time.perf_counter()


def log(msg):
    print(f"[peer_selfcheck {time.perf_counter() - T0:6.2f}s] {msg}", flush=True)


def main():
    import numpy as np
    from polyred_b200 import _abi as A
    from polyred_b200 import partition, render, synth
    max_world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    s, cam = synth.city_scene(n_objects=25, obj_stacks=20, obj_slices=20, ground_cells=60, tex_size=64)
    log("scene built")
    sources, _ = s.Lights()
    cast = [i for i, l in enumerate(sources) if l.cast_shadow]
    bad = 0
    # every group size at a float4-aligned frame size, then the largest group at an odd size (scalar shadow push)
    cases = [(world, 480, 272) for world in range(1, max_world + 1)] + [(max_world, 483, 271)]
    refs = {}
    for world, w, h in cases:
        opts = [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True)]
        if (w, h) not in refs:
            refs[(w, h)] = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
            log(f"1-GPU reference frame {w}x{h} rendered ({int((refs[(w, h)][..., 3] > 0).sum())} px with alpha)")
        ref = refs[(w, h)]
        _, rows = partition.strips(h, world)
        rs, fds, handles = [], [], []
        for k in range(world):
            r = render.NewRenderer(*opts, render.CUDA(0))
            r._ensure_uploaded()
            fd = r.frame_desc(no_readback=True)
            fd.struct.row0, fd.struct.row1 = rows[k]
            handles.append(r._backend.peer_export(fd))
            rs.append(r)
            fds.append(fd)
        for k, r in enumerate(rs):
            r._backend.peer_connect(k, world, handles)
        mine = [rows] * world  # every call carries the strips of all ranks
        t = time.perf_counter()
        status = "ok"
        try:
            for _ in range(3):
                for k, r in enumerate(rs):
                    r._backend.render_peer(fds[k], mine[k], 1)
            for r in rs:
                r._backend.sync()
        except Exception as e:  # noqa: BLE001 - report and go on to the comparison
            status = f"ERROR {e}"
        dt = time.perf_counter() - t
        out = rs[0]._backend.read_image(w, h)
        nd = int((ref != out).any(axis=2).sum())
        # where do differing pixels lie? (which strip: tells a missing strip copy from a missing shadow push)
        per_strip = [int((ref[h - r1:h - r0] != out[h - r1:h - r0]).any(axis=2).sum()) for r0, r1 in rows]
        sm_diff = []
        for li in cast:
            m0 = rs[0]._backend.read_shadowmap(li, w, h)
            # every rank's merged map must be the same (and equal the 1-GPU map: checked through the frame)
            sm_diff.append(max(int((m0 != r._backend.read_shadowmap(li, w, h)).sum()) for r in rs))
        log(f"world={world} {w}x{h}: {status}; 3 frames in {dt * 1e3:.1f} ms; pixels differing from the 1-GPU frame = {nd} (per strip {per_strip}); "
            f"shadow texels differing between the ranks' merged maps = {sm_diff}")
        bad += nd + (status != "ok")
        # strips read back by every rank into ONE host image (prc_set_host_image), no device-side gather
        host = np.zeros((h, w, 4), np.uint8)
        status = "ok"
        try:
            for r in rs:
                r._backend.set_host_image(host.ctypes.data, host.nbytes)
            for k, r in enumerate(rs):
                fds[k].struct.flags &= ~A.PRC_FRAME_NO_READBACK
            for _ in range(2):
                for k, r in enumerate(rs):
                    r._backend.render_peer(fds[k], mine[k], 0)
            for r in rs:
                r._backend.sync()
        except Exception as e:  # noqa: BLE001
            status = f"ERROR {e}"
        nd = int((ref != host).any(axis=2).sum())
        log(f"world={world} strip readback into one host image: {status}; pixels differing = {nd}")
        bad += nd + (status != "ok")
        for r in rs:
            try:
                r._backend.set_host_image(None)
            except Exception as e:  # noqa: BLE001
                log(f"set_host_image(None): {e}")
        for r in rs:
            try:
                r._backend.peer_disconnect()
            except Exception as e:  # noqa: BLE001
                log(f"disconnect: {e}")
            r._backend.close()
    # MSAA(2) over strips: local downsample of each rank's own output rows, read back into one host image
    w, h, m = 240, 136, 2
    opts = [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.MSAA(m)]
    ref = render.NewRenderer(*opts, render.CUDA(0)).Render().copy()
    for world in range(2, max_world + 1):
        rows = [(a * m, b * m) for a, b in partition.strips_from_bounds(h, partition.equal_bounds(h, world))]
        host = np.zeros((h, w, 4), np.uint8)
        rs, fds, handles = [], [], []
        status = "ok"
        try:
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
        except Exception as e:  # noqa: BLE001
            status = f"ERROR {e}"
        nd = int((ref != host).any(axis=2).sum())
        log(f"world={world} MSAA({m}) strips {w}x{h}: {status}; pixels differing from the 1-GPU MSAA frame = {nd}")
        bad += nd + (status != "ok")
        for r in rs:
            try:
                r._backend.set_host_image(None)
                r._backend.peer_disconnect()
            except Exception as e:  # noqa: BLE001
                log(f"cleanup: {e}")
            r._backend.close()
    log("PASS" if bad == 0 else "FAIL")
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
