#!/usr/bin/env python
"""ONE process, N devices behind the C ABI (prc_group_*, csrc/prc_group.cpp): what the Go shim's render.CUDA(0, 1, ...) runs.
Not the bench line (bench.py --gpus N is one process per GPU under torchrun, as the driver launches it): a data point next to it.

    python tools/group_bench.py --devices 0,1,2,3,4,5,6,7 --steps 40

device leg: K frames with PRC_FRAME_NO_READBACK | PRC_FRAME_ASYNC back to back, one prc_group_sync inside the wall-clock bracket
            (strips gathered into device 0's image over NVLink);
e2e leg:    K synchronous prc_group_render calls, host uniforms in, the frame read in place from the group's page-locked host image
            (every GPU DMAs its own strip over its own PCIe link).
Prints one JSON line; `matches_1gpu` = CRC of the group frame against a single-context render of the same frame."""
import argparse, json, os, sys, time, zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--workload", default="C3")
    args = ap.parse_args()
    import bench
    from polyred_b200 import _abi as A, render
    from polyred_b200._lib import PolyredCudaError
    devices = [int(d) for d in args.devices.split(",")]
    wl, s, cam, _ = bench.build_scene(args.workload)
    w, h = wl["w"], wl["h"]
    opts = [render.Camera(cam), render.Size(w, h), render.Scene(s), render.ShadowMap(wl["shadow"]), render.GammaCorrection(wl["gamma"])]
    r = render.NewRenderer(*opts, render.CUDA(*devices)) if len(devices) > 1 else render.NewRenderer(*opts, render.CUDA(devices[0]))
    be = r._backend
    t = time.time()
    r._ensure_uploaded()
    be.sync()
    t_upload = time.time() - t
    fd_sync = r.frame_desc(no_readback=False)
    for _ in range(args.warmup):  # grows the queues, balances the strips (the first synchronous frames after connecting)
        be.render(fd_sync, None)
    strips = be.strips() if hasattr(be, "strips") else [(0, h)]
    fd = r.frame_desc(no_readback=True)
    fd.struct.flags |= A.PRC_FRAME_ASYNC | A.PRC_FRAME_UNIFORMS_RESIDENT  # device leg: the uniforms are already in HBM (uploaded by the frames above), like bench.py's

    def device_leg(k):
        for attempt in range(4):
            t0 = time.perf_counter()
            for _ in range(k):
                be.render(fd, None)
            try:
                be.sync()
                return time.perf_counter() - t0
            except PolyredCudaError as e:
                if e.code != A.PRC_ERR_RETRY:
                    raise
        raise SystemExit("frames kept asking for a retry")

    device_leg(8)
    dev_s = device_leg(args.steps)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        be.render(fd_sync, None)
        img = be.host_image(w, h)
    e2e_s = time.perf_counter() - t0
    crc = zlib.crc32(np.ascontiguousarray(img).tobytes())
    tm = be.timings()
    n_valid = int(tm.n_valid_tris)
    one = render.NewRenderer(*opts, render.CUDA(devices[0]))
    ref = one.Render()
    match = bool(zlib.crc32(np.ascontiguousarray(ref).tobytes()) == crc)
    ms, ms_e2e = dev_s * 1e3 / args.steps, e2e_s * 1e3 / args.steps
    print(json.dumps({
        "api": "prc_group_render (one process, one context + one submit thread per device)", "n_gpus": len(devices), "devices": devices,
        "workload": f"{args.workload}: {wl['desc']}", "steps": args.steps,
        "ms_per_frame_device_resident_wall": ms, "mtris_per_s": n_valid / ms / 1e3, "mpixels_per_s": w * h / ms / 1e3,
        "e2e_ms_per_frame": ms_e2e, "e2e_mtris_per_s": n_valid / ms_e2e / 1e3,
        "strip_rows": [b - a for a, b in strips], "matches_1gpu": match, "frame_crc": crc, "scene_upload_seconds_all_devices": t_upload,
        "timing": "host wall clock around K frames (device leg: back to back, one prc_group_sync; e2e: one synchronous call per frame, image read in place)"}))
    be.close()


if __name__ == "__main__":
    main()
