#!/usr/bin/env python
"""One GPU: how the passes of a C3 frame scale with the rows a context owns (what a rank of an N-GPU group runs).
For each strip [row0, row1) of the screen and each shard of the stacked shadow rows: kernel-class times from prc_get_timings.
    gpurun -- 'python tools/strip_probe.py'"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from polyred_b200 import _abi as A, render  # noqa: E402

wl, s, cam, _ = bench.build_scene(os.environ.get("WORKLOAD", "C3"))
w, h = wl["w"], wl["h"]
r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]), render.GammaCorrection(wl["gamma"]), render.CUDA(0))
be = r._backend
r._ensure_uploaded()
names = A.KERNEL_CLASSES


def run(tag, fn, n=8):
    for _ in range(3):
        fn()
    k = np.zeros(8)
    for _ in range(n):
        fn()
        k += np.array(list(be.timings().kernel_ms))
    k /= n
    print(f"{tag:34s} " + " ".join(f"{names[i][:12]}={k[i]:.4f}" for i in range(8) if k[i] > 0), flush=True)


fd = r.frame_desc(no_readback=True)
run("full frame", lambda: be.render(fd, None))
for n in (2, 4, 8):
    for k in sorted({0, n // 2, n - 1}):
        fd.struct.row0, fd.struct.row1 = k * h // n, (k + 1) * h // n
        run(f"main rows [{fd.struct.row0},{fd.struct.row1})", lambda: be.render_main(fd, None))
fd.struct.row0, fd.struct.row1 = 0, h
sources, _ = s.Lights()
cast = [i for i, l in enumerate(sources) if l.cast_shadow]
run("shadows, 4 lights, all rows", lambda: be.render_shadow_units(fd, [(li, 0, h) for li in cast]))
run("shadows, 1 light, all rows", lambda: be.render_shadow_units(fd, [(cast[0], 0, h)]))
for a, b in ((0, h // 2), (h // 2, h), (h // 2, 5 * h // 8), (5 * h // 8, 6 * h // 8), (1080, 1215), (1215, 1350), (1350, 1485), (1485, 1620)):
    run(f"shadows, light {cast[0]}, rows [{a},{b})", lambda: be.render_shadow_units(fd, [(cast[0], a, b)]))
