#!/usr/bin/env python
"""One process, one scene, several contexts opened under different tuning switches (read at prc_open): the end-to-end leg of
bench.py (host prc_frame in, RGBA8 frame out through the library's page-locked double buffer) and the device-resident leg,
per variant, with the frame's CRC so that a variant that changes a pixel is seen at once.
    python tools/e2e_variants.py [--workload C3] [--steps 30] "PRC_SHADE_BANDS=16" "PRC_SHADE_BANDS=16 PRC_STAGE_UNIFORMS=1" ...
"""
import argparse, json, os, sys, time, zlib
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import bench
from polyred_b200 import _abi as A
from polyred_b200 import render

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C3")
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("variants", nargs="*")
a = ap.parse_args()
wl, s, cam, _ = bench.build_scene(a.workload)
w, h = wl["w"], wl["h"]
opts = [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]), render.GammaCorrection(wl["gamma"])]
sd = None
for spec in ["default"] + a.variants:
    env = dict(kv.split("=", 1) for kv in spec.split()) if spec != "default" else {}
    for k, v in env.items():
        os.environ[k] = v
    r = render.NewRenderer(*opts, render.CUDA(0))
    if sd is not None:
        r._scene_desc, r._scene_desc_for = sd, r.cfg.Scene  # flatten the 10 M triangles once
    sd = r._ensure_uploaded()
    be = r._backend
    fd_dev, fd_e2e = r.frame_desc(no_readback=True), r.frame_desc(no_readback=False)

    def timed(fn, n):
        for _ in range(5):
            fn()
        be.sync(); t0 = time.perf_counter()
        for _ in range(n):
            fn()
        be.sync()
        return (time.perf_counter() - t0) / n * 1e3
    e2e = min(timed(lambda: be.render(fd_e2e, None), a.steps) for _ in range(3))
    crc = zlib.crc32(be.host_image(w, h).tobytes())
    dev = min(timed(lambda: be.render(fd_dev, None), a.steps) for _ in range(2))
    fd_dev.struct.flags |= A.PRC_FRAME_ASYNC | A.PRC_FRAME_UNIFORMS_RESIDENT
    be.render(fd_dev, None); be.sync()
    asy = min(timed(lambda: be.render(fd_dev, None), a.steps) for _ in range(2))
    t = be.timings()
    print(json.dumps({"variant": spec, "e2e_ms": round(e2e, 4), "device_sync_ms": round(dev, 4), "device_async_ms": round(asy, 4), "frame_crc": crc,
                      "kernel_ms_last_frame": {k: round(float(v), 4) for k, v in zip(A.KERNEL_CLASSES, t.kernel_ms) if v}}), flush=True)
    be.close()
    for k in env:
        del os.environ[k]
