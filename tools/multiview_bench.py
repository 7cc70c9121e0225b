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
    ap.add_argument("--devices", default="", help="ONE process, a device group behind the C ABI (prc_group_render_views), e.g. 0,1,2,3; without torchrun")
    ap.add_argument("--per-view-calls", action="store_true", help="round 1's form: one synchronous prc_render per view instead of prc_render_batch")
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
    devices = [int(d) for d in args.devices.split(",") if d != ""] if world == 1 else []
    r = render.NewRenderer(render.Camera(cam0), render.Size(w, h), render.Scene(s), render.ShadowMap(True), render.GammaCorrection(True),
                           render.CUDA(*devices) if len(devices) > 1 else render.CUDA(local))
    be = r._backend
    r._ensure_uploaded()
    mine = list(range(rank, args.views, world))
    cams = [synth.orbit_camera(2 * math.pi * k / args.views, aspect=w / h) for k in mine]  # orbit, angle 2 pi k / views
    t0 = time.time()
    frames = render.ViewFrames(r, cams)
    t_uniforms = time.time() - t0
    grouped = len(devices) > 1
    stream = None if grouped else torch.cuda.ExternalStream(be.stream(), device=torch.device("cuda", local))
    outs = [np.zeros((h, w, 4), np.uint8) for _ in frames]  # the caller-owned images of the batch API

    def one_pass(check=None):
        if args.per_view_calls and not grouped:
            for i, fd in enumerate(frames):
                be.render(fd, None)
                img = be.host_image(w, h)
                if check is not None:
                    check.append(int(img[::8, ::8].astype(np.uint32).sum()))
            return
        be.render_batch(frames, outs)  # prc_render_batch / prc_group_render_views: one call for all views of this process
        if check is not None:
            check.extend(int(o[::8, ::8].astype(np.uint32).sum()) for o in outs)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(); be.sync()

    sums = []
    one_pass(sums)  # warm-up (also grows queues)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw = time.perf_counter()
    if stream is not None:
        e0.record(stream)
    for _ in range(args.repeat):
        one_pass()
    if stream is not None:
        e1.record(stream)
    barrier()
    wall = time.perf_counter() - tw
    ms = e0.elapsed_time(e1) if stream is not None else wall * 1e3
    t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall_ms = float(t[0]), float(t[1])
    n_done = args.views * args.repeat
    tm = be.timings()
    if rank == 0:
        line = {"metric": "multi-view frames/s (C5: views partitioned across GPUs, no collective)", "value": n_done / (wall_ms * 1e-3), "unit": "frames/s",
                "n_gpus": max(world, len(devices)), "api": ("prc_group_render_views (one process)" if grouped else "prc_render per view" if args.per_view_calls else "prc_render_batch"), "views": args.views, "passes": args.repeat, "ms_per_view_per_gpu": ms / (len(mine) * args.repeat),
                "device_ms_total": ms, "wall_ms_total": wall_ms, "mtris_per_s": n_done * tm.n_valid_tris / (wall_ms * 1e-3) / 1e6,
                "mpixels_per_s": n_done * w * h / (wall_ms * 1e-3) / 1e6, "scaling": "strong", "data": "synthetic", "dtype": "f32",
                "config": {"workload": f"C5: {args.views} orbit views of the {args.workload} scene at {w}x{h}, shadows re-fitted and zeroed per view",
                           "timed": "per view: uniforms H2D + all kernels + RGBA8 D2H into page-locked memory + host copy into the caller's image; wall clock, max over ranks",
                           "views_per_rank": len(mine)},
                "host_uniform_prep_seconds_untimed": t_uniforms, "view_checksums_head": sums[:4]}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
