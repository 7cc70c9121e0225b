#!/usr/bin/env python
"""BASELINE.json configs[4] (C5): a batch of camera views at 1920x1080 of the C3 scene, the views partitioned across
the ranks (rank r renders views r, r+N, ...). Views are independent: no data-path collective, each rank uploads the
scene once and submits its views back to back (render.ViewFrames: per-view uniforms + PRC_FRAME_SHADOW_RESET).

  python tools/multiview_bench.py --views 256                           # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 ... tools/multiview_bench.py --views 256

Timed region (per rank, CUDA events on the library's stream, max over ranks): every view = host->device copy of
its uniforms (per-object matrices, light cameras) + all kernels + device->host copy of the RGBA8 frame into the
library's page-locked buffer. Prints one JSON line (rank 0)."""
import argparse, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=256)
    ap.add_argument("--repeat", type=int, default=2, help="timed passes over this rank's views (after one warm-up pass)")
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    args = ap.parse_args()
    import torch
    import bench
    from polyred_b200 import render, synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    real_stdout = os.dup(1); os.dup2(2, 1)  # NCCL banners etc. must not corrupt the JSON line
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = bench.WORKLOADS[args.workload]
    w, h = args.width, args.height
    s, cam0 = synth.city_scene(aspect=w / h, **wl["gen"])
    r = render.NewRenderer(render.Camera(cam0), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True), render.CUDA(local))
    be = r._backend
    r._ensure_uploaded()
    mine = list(range(rank, args.views, world))
    cams = [synth.orbit_camera(2 * math.pi * k / args.views, aspect=w / h) for k in mine]  # orbit, angle 2 pi k / views
    t0 = time.time()
    frames = render.ViewFrames(r, cams)
    t_uniforms = time.time() - t0
    stream = torch.cuda.ExternalStream(be.stream(), device=torch.device("cuda", local))

    def one_pass(check=None):
        for i, fd in enumerate(frames):
            be.render(fd, None)
            img = be.host_image(w, h)
            if check is not None:
                check.append(int(img[::8, ::8].astype(np.uint32).sum()))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); be.sync()

    sums = []
    one_pass(sums)  # warm-up (also grows queues)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    e0.record(stream)
    for _ in range(args.repeat):
        one_pass()
    e1.record(stream)
    barrier()
    wall = time.perf_counter() - tw
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall_ms = float(t[0]), float(t[1])
    n_done = args.views * args.repeat
    tm = be.timings()
    if rank == 0:
        line = {"metric": "multi-view frames/s (C5: views partitioned across GPUs, no collective)", "value": n_done / (wall_ms * 1e-3), "unit": "frames/s",
                "n_gpus": world, "views": args.views, "passes": args.repeat, "ms_per_view_per_gpu": ms / (len(mine) * args.repeat),
                "device_ms_total": ms, "wall_ms_total": wall_ms, "mtris_per_s": n_done * tm.n_valid_tris / (wall_ms * 1e-3) / 1e6,
                "mpixels_per_s": n_done * w * h / (wall_ms * 1e-3) / 1e6, "scaling": "strong", "data": "synthetic", "dtype": "f32",
                "config": {"workload": f"C5: {args.views} orbit views of the {args.workload} scene at {w}x{h}, shadows re-fitted and zeroed per view",
                           "timed": "per view: uniforms H2D + all kernels + RGBA8 D2H into page-locked memory; wall clock, max over ranks",
                           "views_per_rank": len(mine)},
                "host_uniform_prep_seconds_untimed": t_uniforms, "view_checksums_head": sums[:4]}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
