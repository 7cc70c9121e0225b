import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
import bench
from polyred_b200 import render
wl, s, cam, _ = bench.build_scene('C3')
w, h = wl['w'], wl['h']
r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.CUDA(0))
be = r._backend
r._ensure_uploaded()
fd0 = r.frame_desc(no_readback=True); fd1 = r.frame_desc(no_readback=False)
out = np.zeros((h, w, 4), np.uint8)
def t(fn, n=10):
    for _ in range(3): fn()
    be.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    be.sync(); return (time.perf_counter() - t0) / n * 1e3
print('no readback      ms', t(lambda: be.render(fd0, None)))
print('readback pinned  ms', t(lambda: be.render(fd1, None)))
print('readback user    ms', t(lambda: be.render(fd1, out)))
print('frame_desc       ms', t(lambda: r.frame_desc(no_readback=False)))
print('Render()         ms', t(lambda: r.Render()))
import torch
a = torch.empty(w*h*4, dtype=torch.uint8, device='cuda'); b = torch.empty(w*h*4, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
def cp():
    b.copy_(a, non_blocking=True); torch.cuda.synchronize()
print('torch pinned D2H 33MB ms', t(cp))
