"""torchrun probe: per-phase wall time of the multi-GPU frame (serialised with syncs)."""
import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch, torch.distributed as dist
import bench
from polyred_b200 import render, partition, _abi as A
from polyred_b200.distributed import DistributedFrame
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl, s, cam, _ = bench.build_scene('C3')
w, h = wl['w'], wl['h']
r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.CUDA(local))
be = r._backend; r._ensure_uploaded()
df = DistributedFrame(r, rank, world, local)
fd = df.prepare(r.frame_desc(no_readback=True))
for _ in range(3): df.render(fd, False)
def sync():
    torch.cuda.synchronize(); be.sync()
T = {}
def timed(name, fn):
    sync(); dist.barrier(); sync(); t = time.perf_counter(); fn(); sync(); T[name] = T.get(name, 0) + (time.perf_counter() - t) * 1e3
N = 10
for _ in range(N):
    with torch.cuda.stream(df.stream):
        timed('shadows', lambda: be.render_shadow_units(fd, [(li, a, b) for li, a, b, o in df.units if o == rank]))
        ptr, nbytes, cap = be.device_shadow_all(); cb = df.chunk * w * 4
        full = df._view(ptr, cb * world)
        timed('allgather', lambda: dist.all_gather_into_tensor(full, full[rank * cb:(rank + 1) * cb]))
        fd.struct.flags |= A.PRC_FRAME_UNIFORMS_RESIDENT
        timed('forward', lambda: be.render_forward(fd))
        timed('deferred', lambda: be.render_deferred(fd, None))
        fd.struct.flags &= ~A.PRC_FRAME_UNIFORMS_RESIDENT
        ptr, nbytes, cap = be.device_image(); icb = df.img_chunk * w * 4
        img = df._view(ptr, icb * world)
        def gather():
            dist.all_gather_into_tensor(img, img[rank * icb:(rank + 1) * icb])
        timed('img_gather', gather)
    timed('whole_frame', lambda: df.render(fd, False))
    fd.struct.flags |= A.PRC_FRAME_UNIFORMS_RESIDENT
    timed('whole_frame_resident', lambda: df.render(fd, False))
    fd.struct.flags &= ~A.PRC_FRAME_UNIFORMS_RESIDENT
    def comm_only():
        with torch.cuda.stream(df.stream):
            dist.all_gather_into_tensor(full, full[rank * cb:(rank + 1) * cb])
            dist.all_gather_into_tensor(img, img[rank * icb:(rank + 1) * icb])
    timed('comm_only', comm_only)
    timed('barrier_only', lambda: None)
allT = [None] * world
dist.all_gather_object(allT, T)
if rank == 0:
    print("rank0", {k: round(v / N, 3) for k, v in T.items()})
    print("max  ", {k: round(max(t[k] for t in allT) / N, 3) for k in T})
    print("min  ", {k: round(min(t[k] for t in allT) / N, 3) for k in T})
dist.barrier(); dist.destroy_process_group()
