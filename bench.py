#!/usr/bin/env python
"""bench.py — headline benchmark of the render pass (BASELINE.json metric: Mtris/s and Mpixels/s
shaded at 4K on the 10 M-triangle, 8-light, 4-caster scene = configs[2], "C3").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference CPU path (oracle port, host cores)

A step = one `Render()` of the frame. `value` = valid scene triangles x frames/s / 1e6 with the scene
resident in HBM and the image left on the device (CUDA events on the library's stream); `e2e` = the
same through the C-ABI call with HOST buffers: per step the frame uniforms go host->device and the
RGBA8 image comes back device->host inside the timed region.
N>1 (torchrun, one process per GPU), `scaling` = "strong" (the same frame on more GPUs). `--mgpu peer` (default): frames
through prc_render_peer — every rank rasterises ITS SHARE OF THE TRIANGLES (camera pass + all shadow lights) into private
buffers, one kernel merges them into the peers over NVLink peer memory (CUDA IPC; RED.MAX on visibility keys and shadow
depths), every rank shades a strip of rows balanced by measured time and copies it into rank 0's image; ordering by epoch words
in peer memory, frames submitted back to back with no host wait and no collective; e2e through one
shared host image that each GPU writes its strip into. `--mgpu nccl`: round 1's path (screen strips + shadow shards, shadow
maps and image strips all-gathered with NCCL, one frame at a time), kept for comparison. Every N>1 line carries
`matches_1gpu`: the frame's CRC against rank 0's own 1-GPU render of the same frame.
Roofline: `roofline` = the dominant kernel against the measured HBM peak with SURVEY 8(d)'s algorithmic
bytes (per rank at N>1: the bytes of the rows that rank owns); `shading_roofline` = the shading kernel
against the FP32 FMA peak measured by prc_measure_fp32_peak in this run (SURVEY 8(d) flop per covered pixel);
`frame_roofline` additionally states the covered pixels and a coverage-weighted fraction.
`e2e.python_renderer_render_ms_per_step` (N=1, informational): the same frame through the Python
mirror's Renderer.Render(), i.e. including the host-side uniforms.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (width, height, generator kwargs)
    "C3": dict(w=3840, h=2160, gen=dict(n_objects=1600, obj_stacks=50, obj_slices=50, ground_cells=1000, n_lights=8, casting_every=2,
                                        n_materials=8, tex_size=256), shadow=True, gamma=True,
               desc="3840x2160, 10.0M-triangle synthetic scene (2M-tri ground heightfield + 1600 instanced 5000-tri meshes), 8 materials, "
                    "8 point lights, 4 shadow-casting, gamma"),
    # the C3 scene from a camera inside the city: 85 % of the 4K frame is covered (the orbit camera of C3 leaves 75 % sky), so the
    # shading pass is actually loaded; a second data point (`--workload C3-close`), the bench line stays C3
    "C3-close": dict(w=3840, h=2160, gen=dict(n_objects=1600, obj_stacks=50, obj_slices=50, ground_cells=1000, n_lights=8, casting_every=2,
                                              n_materials=8, tex_size=256, cam_radius=1.0, cam_height=0.9), shadow=True, gamma=True,
                     desc="3840x2160, the C3 scene (10.0M triangles, 8 lights, 4 casting, gamma) seen from inside the city: ~85 % of the pixels covered"),
    # BASELINE configs[3]; not a bench line (the bench line is C3): `--workload C4 --no-cpu-baseline` records one data point
    "C4": dict(w=7680, h=4320, gen=dict(n_objects=18000, obj_stacks=50, obj_slices=50, ground_cells=2236, n_lights=1, casting_every=0,
                                        n_materials=8, tex_size=256), shadow=False, gamma=False,
               desc="7680x4320, 100.0M-triangle synthetic scene (10M-tri ground heightfield + 18000 instanced 5000-tri meshes), 8 materials, "
                    "1 point light + ambient, no shadows, no gamma"),
    # BASELINE configs[0] / configs[1] (parity-test cases; `--workload C1|C2` records a data point)
    "C1": dict(w=800, h=500, mesh=dict(subdiv=187), shadow=False, gamma=False,
               desc="800x500, 69 938-triangle closed mesh, textured Blinn-Phong, 1 point light + ambient, no shadows"),
    "C2": dict(w=1920, h=1080, mesh=dict(subdiv=187, with_ground=True, shadows=True, ao=True), shadow=True, gamma=True,
               desc="1920x1080, 69 938-triangle mesh + ground quad, directional + shadow-casting point light, shadow map, "
                    "ambient occlusion on every material, gamma"),
    "C3-small": dict(w=960, h=540, gen=dict(n_objects=100, obj_stacks=20, obj_slices=20, ground_cells=100, n_lights=8, casting_every=2,
                                            n_materials=8, tex_size=64), shadow=True, gamma=True, desc="reduced C3 for plumbing tests"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions: NVML polled every 2 ms from a thread (the nvidia-smi
    loop of the profiling recipe buffers its output, a 30 ms timed region would see one line); falls back to nvidia-smi."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0):
        self.rows, self.windows, self.gpu = [], [], gpu_index  # rows: (t, sm_mhz, max_mhz, reason bitmask)
        self.proc = self.thread = self.nvml = None
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def poll():
                while not self._stop:
                    try:
                        self.rows.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), mx, int(get_reasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nvml = pynvml
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        order = (0x8, 0x40, 0x20, 0x4)
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                mask = sum(bit for bit, v in zip(order, r[2:6]) if v.lower().startswith("active"))
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]), mask))
            except Exception:
                pass

    def window(self, t0, t1):
        """A timed region (perf_counter interval): the statistics are taken over the samples inside the windows."""
        self.windows.append((t0, t1))

    def stop(self):
        self._stop = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"], "samples": 0}
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in self.windows)]
        rows = inside or self.rows
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": max(r[2] for r in rows),
                "reasons": [n for n, bit in self.REASONS if mask & bit], "samples": len(rows),
                "source": ("nvml" if self.nvml else "nvidia-smi") + (", samples inside the timed regions" if inside else ", whole run (none fell inside the timed regions)")}


def build_scene(name):
    from polyred_b200 import synth
    wl = WORKLOADS[name]
    t = time.time()
    if "mesh" in wl:
        s, cam = synth.mesh_scene(aspect=wl["w"] / wl["h"], **wl["mesh"])
    else:
        s, cam = synth.city_scene(aspect=wl["w"] / wl["h"], **wl["gen"])
    return wl, s, cam, time.time() - t


def algorithmic_bytes(n_tris, px, n_cast):
    """bytes/frame = 112 N + Ls 36 N + px (148 + 12 Ls)."""
    return 112 * n_tris + n_cast * 36 * n_tris + px * (148 + 12 * n_cast)


def shading_flops(px_covered, n_lights, n_cast, ao=False):
    """flop per lit pixel = 60 (texture query) + 110 per light + 70 per casting light (+ 28 000 with AO)."""
    return px_covered * (60 + 110 * n_lights + 70 * n_cast + (28000 if ao else 0))


def workload_config(name, wl, n_tris, n_valid):
    """The `config` object — the same keys and values in both arms (the driver compares them)."""
    return {"workload": f"{name}: {wl['desc']}", "n_tris": int(n_tris), "n_valid_tris": int(n_valid), "width": wl["w"], "height": wl["h"],
            # timing rule: no explicit L2 flush between the timed frames — every frame streams the scene (112 B/triangle + shared vertices)
            # and rewrites the frame buffers, both far larger than the 126 MB L2 for C3 / C4 (C1 / C2 fit: they are not bench lines)
            "l2": "no flush between frames: scene and frame buffers are larger than L2" if n_tris >= 1_000_000 else "no flush between frames (the workload fits L2)"}


def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores: the oracle port
    (oracle/libpr_oracle.so, the C++ restatement of render/*.go) in its multithreaded mode — the Go
    binary cannot be built in this image (no Go toolchain). Rank 0 only. Each step renders the full frame
    of the workload (about 1 s on 16 cores for C3), so the default K/W finish within a minute."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle_binding as ob
    from polyred_b200 import render
    wl, s, cam, tgen = build_scene(args.workload)
    cores = os.cpu_count() or 1
    be = ob.OracleBackend(threads=cores)
    w, h = wl["w"], wl["h"]
    r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]),
                           render.GammaCorrection(wl["gamma"]), render._Backend(be))
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        r.Render()
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    tm = be.timings()
    sec = float(np.mean(times))
    val = tm.n_valid_tris / sec / 1e6
    line = {
        "impl": "reference", "metric": "Mtris/s", "value": val, "unit": "Mtris/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mpixels_per_s": w * h / sec / 1e6,
        "config": workload_config(args.workload, wl, r.scene_desc().n_tris, tm.n_valid_tris),
        "cpu_baseline": {"value": val, "unit": "Mtris/s", "cores": cores, "kind": "port",
                         "sample": "full frame of the same workload per step, C++ restatement of the reference CPU path (not the Go binary), one task per 256 triangles / 32 pixels, per-pixel spinlocks"},
        "e2e": {"value": val, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_cuda(args):
    import zlib

    import torch
    from polyred_b200 import _abi as A
    from polyred_b200 import render
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus must equal WORLD_SIZE")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl, s, cam, tgen = build_scene(args.workload)
    w, h = wl["w"], wl["h"]
    px = w * h
    r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]),
                           render.GammaCorrection(wl["gamma"]), render.CUDA(local))
    be = r._backend
    t = time.time()
    sd = r._ensure_uploaded()
    be.sync()
    t_upload = time.time() - t
    sources, _ = s.Lights()
    cast = [i for i, l in enumerate(sources) if l.cast_shadow] if wl["shadow"] else []
    any_ao = any(getattr(m, "ambient_occlusion", False) for m in sd.materials)
    stream = torch.cuda.ExternalStream(be.stream(), device=torch.device("cuda", local))
    fp32_peak = be.measure_fp32_peak()  # TFLOP/s, FMA micro-benchmark on this GPU, before anything is timed

    fd = r.frame_desc(no_readback=True)
    fd_e2e = r.frame_desc(no_readback=False)
    e2e_out = [None]

    # ---- multi-GPU partition: screen strips + shadow (light, row-range) shards (polyred_b200/distributed.py) ----
    df = pf = None
    units = []
    if world > 1:
        from polyred_b200.distributed import DistributedFrame, PeerFrames
        if args.mgpu == "peer":
            # frames over NVLink peer memory (prc_render_peer), submitted back to back, no collective
            pf = PeerFrames(r, rank, world, local, root=0)
            shared_host_image = pf.share_host_image()  # e2e leg: every GPU DMAs its own strip into one shared host image
        else:
            df = DistributedFrame(r, rank, world, local)
            df.prepare(fd)
            df.prepare(fd_e2e)
            units = df.units

    def step(fdesc, host_out):
        if world == 1:
            # e2e: rgba_out = NULL + prc_host_image = the frame DMA'd into the library's page-locked double buffer and
            # read in place by the caller (what the Go shim does); the device-resident leg passes PRC_FRAME_NO_READBACK
            be.render(fdesc, None)
            if host_out is not None:
                host_out[0] = be.host_image(w, h)
        elif pf is not None:
            if host_out is None:
                pf.submit(fdesc)  # device leg: strips gathered into rank 0's device image over NVLink
            else:
                # e2e: uniforms in (every rank), the frame out: each rank's strip over its own PCIe link into the shared host
                # image; the frame is complete when every rank has passed the barrier
                pf.submit(fdesc, gather=False)
                pf.finish(fast=True)
                host_out[0] = shared_host_image
        else:
            img = df.render(fdesc, host_out is not None)
            if host_out is not None:
                host_out[0] = img

    def barrier():
        if pf is not None:
            pf.finish()  # completes the submitted frames on every rank (and re-submits them if a queue had to grow)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        be.sync()

    # clocks / throttle reasons are sampled from the warm-up on (the timed region of a 2 ms frame is only tens of ms long)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- warm-up ----
    for _ in range(max(3, args.warmup)):
        step(fd, None)
    barrier()
    balance_log = None
    if pf is not None and args.balance:
        # peer-memory frames take any row ranges: move the strip / shadow-shard boundaries until every rank takes the same
        # time (per-rank kernel times of two frames per round); the frame itself does not depend on the partition
        balance_log = []
        for _ in range(args.balance_rounds):
            for _ in range(2):
                step(fd, None)
            barrier()
            costs = pf.rebalance()
            balance_log.append({"raster_ms": [round(c[0] / 2, 4) for c in costs], "shade_ms": [round(c[1] / 2, 4) for c in costs]})
        for _ in range(2):
            step(fd, None)
        barrier()
    if args.resident_uniforms:
        # device-resident leg: the per-frame uniforms (object matrices, lights) are inputs already in HBM; the warm-up
        # frames uploaded them, the timed frames reuse them (the e2e leg uploads them every frame)
        fd.struct.flags |= A.PRC_FRAME_UNIFORMS_RESIDENT
    n_valid = int(be.timings().n_valid_tris)

    # ---- timed region 1: device-resident (value) ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    back_to_back = (world == 1 and args.async_frames) or pf is not None

    def timed_frames(k_frames, kernel_timers):
        """K frames between two events on the library's stream; returns (device ms, wall s, timings of the K frames)."""
        ksum, klaunch, launches = np.zeros(8), np.zeros(8, np.int64), 0
        if not kernel_timers:
            fd.struct.flags |= A.PRC_FRAME_NO_KERNEL_TIMERS
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
        t0 = time.perf_counter()
        if back_to_back:
            # frames stay on the device: submit them back to back, one prc_sync inside the timed bracket; the per-class
            # kernel timings come back summed over the K frames
            if pf is None:
                fd.struct.flags |= A.PRC_FRAME_ASYNC
            for _ in range(k_frames):
                step(fd, None)
            with torch.cuda.stream(stream):
                e1.record(stream)
            barrier()  # be.sync() finishes the asynchronous frames (raises if a queue overflowed: warm-up sized them)
            fd.struct.flags &= ~A.PRC_FRAME_ASYNC
            tm = be.timings()
            ksum += np.array(list(tm.kernel_ms))
            klaunch += np.array(list(tm.kernel_launches))
            launches += int(tm.gpu_launches)
        else:
            for _ in range(k_frames):
                step(fd, None)
                tm = be.timings()
                ksum += np.array(list(tm.kernel_ms))
                klaunch += np.array(list(tm.kernel_launches))
                launches += int(tm.gpu_launches)
            with torch.cuda.stream(stream):
                e1.record(stream)
            barrier()
        wall = time.perf_counter() - t0
        fd.struct.flags &= ~A.PRC_FRAME_NO_KERNEL_TIMERS
        return e0.elapsed_time(e1), wall, t0, tm, ksum, klaunch, launches

    # One GPU: the per-class event brackets stay on inside the timed region (they cost ~10 us of a 1.1 ms frame). N > 1: a
    # frame is ~0.2 ms per rank, so the timed frames run without them and a second, untimed pass of K frames measures the classes.
    dev_ms, wall, t0, tm, ksum, klaunch, launches = timed_frames(args.steps, kernel_timers=(world == 1))
    sampler.window(t0, t0 + wall)
    peer_wait = be.peer_wait_ms() if pf is not None else None  # PRC_PEER_TRACE=1: where this rank idled for its peers
    kernel_pass = "inside the timed region"
    if world > 1:
        _, _, _, _, ksum, klaunch, _ = timed_frames(args.steps, kernel_timers=True)
        kernel_pass = "a second pass of K frames with the event brackets on (the timed frames run without them)"
    clocks = sampler.stop() if rank == 0 else None

    # ---- timed region 2: end to end through the C ABI with host buffers ----
    for _ in range(3):  # warm-up of the readback path (the page-locked host images are created on first use)
        step(fd_e2e, e2e_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(fd_e2e, e2e_out)
    barrier()
    wall_e2e = time.perf_counter() - t0

    # ---- checks outside the timed regions: frame CRC (N > 1: against this rank's own 1-GPU render), covered pixels ----
    frame_crc = matches = covered_px = None
    if rank == 0:
        img = np.ascontiguousarray(e2e_out[0]).copy()
        frame_crc = zlib.crc32(img.tobytes())
    if world > 1:
        barrier()
        if rank == 0:
            r1 = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]),
                                    render.GammaCorrection(wl["gamma"]), render.CUDA(local))
            ref = r1.Render()
            matches = bool(zlib.crc32(np.ascontiguousarray(ref).tobytes()) == frame_crc)
            covered_px = int(r1._backend.covered_pixels(w, h)) if hasattr(r1._backend, "covered_pixels") else None
            r1._backend.close()
    elif hasattr(be, "covered_pixels"):
        covered_px = int(be.covered_pixels(w, h))

    # the same frames through the host mirror's Renderer.Render() (scene-graph walk, uniforms, prc_render, image in place): what a
    # Python caller pays on top of the C-ABI call. Informational only; never allowed to break the bench line.
    mirror_ms = None
    if world == 1:
        try:
            for _ in range(2):
                r.Render()
            t1 = time.perf_counter()
            for _ in range(args.steps):
                r.Render()
            mirror_ms = (time.perf_counter() - t1) * 1e3 / args.steps
        except Exception as e:  # noqa: BLE001
            print(f"Renderer.Render() leg skipped: {e}", file=sys.stderr)

    if dist is not None:
        tt = torch.tensor([dev_ms, wall * 1e3, wall_e2e * 1e3], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, wall_e2e_ms = [float(x) for x in tt.tolist()]
    else:
        wall_ms, wall_e2e_ms = wall * 1e3, wall_e2e * 1e3
    if rank != 0:
        if pf is not None:
            pf.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    ms = dev_ms / args.steps
    fps = 1e3 / ms
    fps_e2e = 1e3 / (wall_e2e_ms / args.steps)
    pk, pk_kind = peaks()
    names = A.KERNEL_CLASSES
    n_tris = sd.n_tris
    # ---- roofline of the dominant kernel class (this rank's launches, this rank's share of the algorithmic bytes) ----
    dom = int(np.argmax(ksum))
    if pf is not None:
        # peer groups: the raster passes take 1/N of the triangles (every row), the shading its strip of rows
        f_img = (pf.rows[rank][1] - pf.rows[rank][0]) / h
        f_sh = f_cam = 1.0 / world
    elif df is not None:
        f_img = f_cam = (df.rows[rank][1] - df.rows[rank][0]) / h       # this rank's share of the screen rows
        f_sh = sum(b - a for _, a, b, owner in units if owner == rank) / max(1, len(cast) * h)  # ... of the stacked shadow-map rows
    else:
        f_img = f_sh = f_cam = 1.0
    sweeps = max(1, int(round(klaunch[0] / args.steps)))    # shadow sweeps per frame on this rank (up to 8 views share one)
    # 1 GPU / peer frames without an AO material: resolve and shading run as ONE kernel (k_resolve_shade, timed in the "shade"
    # class); it is charged the algorithmic bytes of both stages, G-buffer round trip included
    one_kernel_shade = (world == 1 or pf is not None) and not os.environ.get("PRC_NO_FUSED_SHADE") and not any_ao
    shade_bytes = f_img * ((64 + 4 + 4 * len(cast)) * px + ((8 + 64) * px if one_kernel_shade else 0))
    per_launch_bytes = {0: f_sh * len(cast) * (36 * n_tris + 8 * px) / sweeps, 1: f_cam * (112 * n_tris + 16 * px), 6: f_img * (8 + 64) * px, 7: shade_bytes}
    alg = per_launch_bytes.get(dom, 0.0)
    launches_per_step = float(klaunch[dom]) / args.steps
    avg_ms = ksum[dom] / max(1, klaunch[dom])
    achieved = alg / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
    frame_bytes = algorithmic_bytes(n_tris, px, len(cast))
    covered_bytes = (112 * n_tris + len(cast) * 36 * n_tris + (8 + 8) * px + len(cast) * 12 * px + (128 + 4) * (covered_px if covered_px is not None else px)
                     + 4 * (px - (covered_px if covered_px is not None else px)))
    # DRAM traffic of the dominant kernel: from the committed `ncu --set full` capture of this command (profiles/), labelled
    traffic = traffic_src = None
    import glob
    tpaths = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_r*.json")))  # the latest round's capture
    if world == 1 and args.workload == "C3" and tpaths:
        tj = json.load(open(tpaths[-1]))
        traffic = tj.get(names[dom], {}).get("dram_bytes_per_launch")
        traffic_src = f"{os.path.relpath(tpaths[-1], ROOT)} ({tj.get('_commit', 'commit not recorded')})"
    shade_ms = float(ksum[7]) / args.steps
    sflops = shading_flops((covered_px if covered_px is not None else px) * f_img, len(sources), len(cast), any_ao)
    h2d = fd.struct.n_objects * 128 + len(cast) * fd.struct.n_objects * 64 + len(sources) * 168 + 256 + 4 * fd.struct.n_ambient
    line = {
        "metric": "Mtris/s", "value": n_valid * fps / 1e6, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mpixels_per_s": px * fps / 1e6, "frames_per_s": fps, "wall_ms_per_step": wall_ms / args.steps,
        "config": workload_config(args.workload, wl, n_tris, n_valid),
        "run": {"fma": os.environ.get("PRC_FMA", "mixed"), "l2": "inputs (1.1 GB scene, 0.9 GB frame buffers) larger than L2; no explicit flush",
                "submit": ("K frames back to back (PRC_FRAME_ASYNC), one prc_sync inside the timed bracket" if (world == 1 and args.async_frames)
                           else "K frames back to back (prc_render_peer), one prc_sync inside the timed bracket" if pf is not None
                           else "one synchronous call per frame"),
                "partition": "1 GPU" if world == 1
                else f"raster passes (camera + all shadow lights): 1/{world} of the triangles per rank into private buffers, merged into the peers with atomicMax over NVLink peer memory (shadow texels into every rank, visibility keys into the rank shading the row); shading: {world} screen strips balanced by measured time; image strips copied to rank 0; no collective, no host wait inside a frame" if pf is not None
                else f"{world} screen strips + {len(units)} shadow shards (one in-place NCCL all-gather overlapped with the camera pass, one in-place all-gather of the image strips)",
                "kernel_timings": kernel_pass},
        "roofline": {"bound": "hbm", "kernel": ("resolve_shade (k_resolve_shade, one kernel)" if (dom == 7 and one_kernel_shade) else names[dom]), "achieved": achieved, "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s",
                     "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_ms,
                     "launches_per_step": launches_per_step,
                     "scope": "rank 0's launches against rank 0's share of the algorithmic bytes (rows it owns)" if world > 1 else "the whole frame on one GPU"},
        "shading_roofline": {"bound": "fp32", "kernel": "resolve_shade" if one_kernel_shade else "shade", "algorithmic_flop_per_launch_set": sflops,
                             "achieved": sflops / (shade_ms * 1e-3) / 1e12 if shade_ms > 0 else 0.0, "peak": fp32_peak, "unit": "TFLOP/s",
                             "peak_kind": "measured in this run (prc_measure_fp32_peak: pure FMA chains, 2 flop each)",
                             "frac": (sflops / (shade_ms * 1e-3) / 1e12 / fp32_peak) if (shade_ms > 0 and fp32_peak > 0) else 0.0,
                             "ms_per_step": shade_ms,
                             "note": "SURVEY 8(d) flop per COVERED pixel (60 + 110 L + 70 Ls); issue-slot and pipe utilisation of this kernel: profiles/ncu_shade_r2.txt"},
        "frame_roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": frame_bytes, "achieved": frame_bytes / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"] * world,
                           "frac": frame_bytes / (ms * 1e-3) / 1e9 / (pk["hbm_gbs"] * world), "unit": "GB/s", "covered_px": covered_px,
                           "coverage": (covered_px / px) if covered_px is not None else None,
                           "coverage_weighted_bytes_per_frame": covered_bytes,
                           "coverage_weighted_frac": covered_bytes / (ms * 1e-3) / 1e9 / (pk["hbm_gbs"] * world),
                           "note": "peak = N x the measured single-GPU HBM peak; coverage-weighted: the 128 B/px G-buffer round trip charged to covered pixels only"},
        "kernel_ms_per_step": {names[k]: float(ksum[k]) / args.steps for k in range(8)},
        "e2e": {"value": n_valid * fps_e2e / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(px * 4),
                "ms_per_step": wall_e2e_ms / args.steps, "mpixels_per_s": px * fps_e2e / 1e6,
                "note": ("prc_render_peer through the C ABI on every rank: host prc_frame (per-object matrices) in on every rank, host RGBA8 out as one shared-memory image "
                         "that each GPU writes its strip into over its own PCIe link; scene resident after one prc_scene_upload per rank" if pf is not None else
                         "prc_render through the C ABI: host prc_frame (per-object matrices) in, host RGBA8 out; scene resident after one prc_scene_upload"),
                "scene_upload_once": {"bytes": int(sd.upload_bytes()), "seconds": t_upload},
                "python_renderer_render_ms_per_step": mirror_ms},
        "frame_crc": frame_crc, "matches_1gpu": matches,
        "stats_last_frame": {"n_large_items": int(tm.n_large_items), "n_clipped": int(tm.n_clipped), "n_bin_entries": int(tm.n_bin_entries), "n_nan_frags": int(tm.n_nan_frags)},
        "gpu_launches": launches, "clocks": clocks, "scene_gen_seconds": tgen,
    }
    if pf is not None:
        line["run"]["strip_rows"] = [r1_ - r0_ for r0_, r1_ in pf.rows]
        line["balance_rounds"] = balance_log
        line["peer_wait_ms_per_step_rank0"] = {k: v / args.steps for k, v in peer_wait.items()} if peer_wait else None
        pf.close()
    if args.cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(args, wl, s, cam)
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(args, wl, s, cam):
    """The oracle port timed on this box's host cores on a bounded sample of the same workload."""
    import oracle_binding as ob
    from polyred_b200 import render
    cores = os.cpu_count() or 1
    be = ob.OracleBackend(threads=cores)
    w, h = wl["w"], wl["h"]
    r = render.NewRenderer(render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]),
                           render.GammaCorrection(wl["gamma"]), render._Backend(be))
    t = time.perf_counter()
    r.Render()
    sec = time.perf_counter() - t
    tm = be.timings()
    return {"value": tm.n_valid_tris / sec / 1e6, "unit": "Mtris/s", "cores": cores, "kind": "port", "seconds": sec,
            "mpixels_per_s": w * h / sec / 1e6,
            "sample": "one full frame of the same workload (no warm-up), C++ restatement of the reference CPU path in multithreaded mode "
                      "(tasks of 256 triangles / 32 pixels, per-pixel spinlocks); not the Go binary"}


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to fd 1 (e.g. "NCCL version ...") must not corrupt the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-resident-uniforms", dest="resident_uniforms", action="store_false")
    ap.add_argument("--no-async-frames", dest="async_frames", action="store_false",
                    help="device-resident leg: wait for every frame (prc_render) instead of submitting the K frames back to back")
    ap.add_argument("--no-balance", dest="balance", action="store_false", help="--mgpu peer: keep equal strips / shadow shards")
    ap.add_argument("--balance-rounds", type=int, default=10)
    ap.add_argument("--mgpu", default=os.environ.get("PRC_MGPU", "peer"), choices=["nccl", "peer"],
                    help="N > 1: 'peer' (default) = frames submitted back to back, exchange pushed over NVLink peer memory by the library "
                         "(prc_render_peer); 'nccl' = round 1's path: one frame at a time, shadow maps and image strips all-gathered with NCCL")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
