"""ctypes binding of libpolyred_cuda.so (include/polyred_cuda.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is
raised (north_star: no CPU fallback; the reference's per-pass fallback, render/raster.go:68-75,
is deliberately not mirrored)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi as A

_LIB = None
LIB_PATH = os.environ.get("PRC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libpolyred_cuda.so")  # PRC_LIB: a tuning build of the same library


class PolyredCudaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpolyred_cuda: error {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PolyredCudaError(A.PRC_ERR_INVALID, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.prc_abi_version.restype = C.c_uint32
        L.prc_device_count.restype = C.c_int32
        L.prc_open.argtypes = [C.c_int32, C.POINTER(vp)]
        L.prc_close.argtypes = [vp]
        L.prc_last_error.argtypes = [vp]
        L.prc_last_error.restype = C.c_char_p
        L.prc_scene_upload.argtypes = [vp, C.POINTER(A.prc_scene)]
        L.prc_shadow_reset.argtypes = [vp]
        L.prc_render.argtypes = [vp, C.POINTER(A.prc_frame), vp]
        L.prc_read_gbuffer.argtypes = [vp, C.POINTER(A.prc_gbuffer_host)]
        L.prc_read_shadowmap.argtypes = [vp, C.c_uint32, vp]
        L.prc_get_timings.argtypes = [vp, C.POINTER(A.prc_timings)]
        L.prc_read_image.argtypes = [vp, vp]
        L.prc_device_image.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.prc_device_shadowmap.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.prc_render_shadows.argtypes = [vp, C.POINTER(A.prc_frame), C.c_uint32, C.c_uint32, C.c_uint32]
        L.prc_render_main.argtypes = [vp, C.POINTER(A.prc_frame), vp]
        L.prc_render_shadow_units.argtypes = [vp, C.POINTER(A.prc_frame), C.c_uint32, vp, vp, vp]
        L.prc_render_forward.argtypes = [vp, C.POINTER(A.prc_frame)]
        L.prc_render_deferred.argtypes = [vp, C.POINTER(A.prc_frame), vp]
        L.prc_device_shadow_all.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.prc_host_image.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.prc_stream.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.prc_sync.argtypes = [vp]
        L.prc_peer_export.argtypes = [vp, C.POINTER(A.prc_frame), C.POINTER(A.prc_peer_handle)]
        L.prc_peer_connect.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(A.prc_peer_handle)]
        L.prc_peer_disconnect.argtypes = [vp]
        L.prc_render_peer.argtypes = [vp, C.POINTER(A.prc_frame), C.c_uint32, vp, vp, C.c_uint32]
        L.prc_frame_state.argtypes = [vp, C.POINTER(C.c_uint32)]
        L.prc_set_frame_state.argtypes = [vp, C.c_uint32]
        L.prc_set_exact_fma.argtypes = [vp, C.c_int32]
        L.prc_set_host_image.argtypes = [vp, vp, C.c_uint64]
        L.prc_peer_wait_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
        L.prc_measure_fp32_peak.argtypes = [vp, C.POINTER(C.c_double)]
        L.prc_count_covered.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.prc_render_batch.argtypes = [vp, C.c_uint32, C.POINTER(A.prc_frame), C.POINTER(vp)]
        L.prc_set_host_image_offset.argtypes = [vp, C.c_uint64]
        # device groups (one process, one context per device)
        L.prc_group_open.argtypes = [C.POINTER(C.c_int32), C.c_uint32, C.POINTER(vp)]
        L.prc_group_close.argtypes = [vp]
        L.prc_group_last_error.argtypes = [vp]
        L.prc_group_last_error.restype = C.c_char_p
        L.prc_group_size.argtypes = [vp]
        L.prc_group_size.restype = C.c_uint32
        L.prc_group_ctx.argtypes = [vp, C.c_uint32, C.POINTER(vp)]
        L.prc_group_scene_upload.argtypes = [vp, C.POINTER(A.prc_scene)]
        L.prc_group_shadow_reset.argtypes = [vp]
        L.prc_group_render.argtypes = [vp, C.POINTER(A.prc_frame), vp]
        L.prc_group_sync.argtypes = [vp]
        L.prc_group_host_image.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.prc_group_strips.argtypes = [vp, vp, vp]
        L.prc_group_render_views.argtypes = [vp, C.c_uint32, C.POINTER(A.prc_frame), C.POINTER(vp)]
        for name in ("prc_render_batch", "prc_set_host_image_offset", "prc_group_open", "prc_group_close", "prc_group_ctx", "prc_group_scene_upload",
                     "prc_group_shadow_reset", "prc_group_render", "prc_group_sync", "prc_group_host_image", "prc_group_strips", "prc_group_render_views"):
            getattr(L, name).restype = C.c_int32
        for name in ("prc_open", "prc_close", "prc_scene_upload", "prc_shadow_reset", "prc_render", "prc_read_gbuffer",
                     "prc_read_shadowmap", "prc_get_timings", "prc_device_image", "prc_device_shadowmap",
                     "prc_render_shadows", "prc_render_main", "prc_stream", "prc_sync", "prc_host_image", "prc_render_forward",
                     "prc_render_deferred", "prc_device_shadow_all", "prc_render_shadow_units", "prc_peer_export", "prc_peer_connect",
                     "prc_peer_disconnect", "prc_render_peer", "prc_set_exact_fma", "prc_read_image", "prc_set_host_image", "prc_peer_wait_ms",
                     "prc_measure_fp32_peak", "prc_count_covered", "prc_frame_state", "prc_set_frame_state"):
            getattr(L, name).restype = C.c_int32
        if L.prc_abi_version() != A.PRC_ABI_VERSION:
            raise PolyredCudaError(A.PRC_ERR_INVALID, "ABI version mismatch")
        _LIB = L
    return _LIB


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class Backend:
    """Common surface of a render backend: the CUDA library here, the CPU oracle in tests/."""

    prefix = "prc"

    def __init__(self, L, handle):
        self.L, self.h = L, handle

    def _f(self, name):
        return getattr(self.L, f"{self.prefix}_{name}")

    def _check(self, rc):
        if rc != 0:
            raise PolyredCudaError(rc, (self._f("last_error")(self.h) or b"").decode())

    def close(self):
        if self.h:
            self._f("close")(self.h)
            self.h = None

    def scene_upload(self, sd):
        self._check(self._f("scene_upload")(self.h, C.byref(sd.struct)))

    def shadow_reset(self):
        self._check(self._f("shadow_reset")(self.h))

    def render(self, fd, out: np.ndarray | None):
        self._check(self._f("render")(self.h, C.byref(fd.struct), out.ctypes.data if out is not None else None))

    def timings(self) -> A.prc_timings:
        t = A.prc_timings()
        self._check(self._f("get_timings")(self.h, C.byref(t)))
        return t

    def read_gbuffer(self, w, h):
        n = w * h
        g = {
            "ok": np.zeros(n, np.uint8), "tri": np.zeros(n, np.int32), "sub": np.zeros(n, np.int32),
            "depth": np.zeros(n, np.float32), "uv": np.zeros((n, 2), np.float32), "dudv": np.zeros((n, 2), np.float32),
            "nor": np.zeros((n, 3), np.float32), "facenor": np.zeros((n, 3), np.float32), "wpos": np.zeros((n, 3), np.float32),
            "col": np.zeros(n, np.uint32), "mat": np.zeros(n, np.int32),
        }
        s = A.prc_gbuffer_host(abi_version=A.PRC_ABI_VERSION)
        s.ok = _ptr(g["ok"], C.c_uint8); s.tri = _ptr(g["tri"], C.c_int32); s.sub = _ptr(g["sub"], C.c_int32)
        s.depth = _ptr(g["depth"], C.c_float); s.uv = _ptr(g["uv"], C.c_float); s.dudv = _ptr(g["dudv"], C.c_float)
        s.nor = _ptr(g["nor"], C.c_float); s.facenor = _ptr(g["facenor"], C.c_float); s.wpos = _ptr(g["wpos"], C.c_float)
        s.col = _ptr(g["col"], C.c_uint32); s.mat = _ptr(g["mat"], C.c_int32)
        self._check(self._f("read_gbuffer")(self.h, C.byref(s)))
        return {k: v.reshape((h, w) + v.shape[1:]) for k, v in g.items()}

    def read_shadowmap(self, light, w, h):
        out = np.zeros((h, w), np.float32)
        self._check(self._f("read_shadowmap")(self.h, light, out.ctypes.data))
        return out


class CudaBackend(Backend):
    def __init__(self, device: int = 0):
        L = lib()
        h = C.c_void_p()
        rc = L.prc_open(device, C.byref(h))
        if rc != 0:
            raise PolyredCudaError(rc, f"prc_open(device={device}) failed (is a CUDA device visible? there is no CPU fallback)")
        super().__init__(L, h)
        self.device = device

    def host_image(self, w, h) -> np.ndarray:
        """The last frame, in place in the library's page-locked double buffer (no copy)."""
        p, n = C.c_uint64(), C.c_uint64()
        self._check(self.L.prc_host_image(self.h, C.byref(p), C.byref(n)))
        key = (p.value, n.value, w, h)
        views = self.__dict__.setdefault("_host_views", {})
        if key not in views:
            buf = (C.c_uint8 * n.value).from_address(p.value)
            views[key] = np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 4)
        return views[key]

    def read_image(self, w, h) -> np.ndarray:
        """The device-resident frame (PRC_FRAME_NO_READBACK) copied to a fresh host array."""
        out = np.zeros((h, w, 4), np.uint8)
        self._check(self.L.prc_read_image(self.h, out.ctypes.data))
        return out

    def device_image(self):
        p, n, cap = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.L.prc_device_image(self.h, C.byref(p), C.byref(n), C.byref(cap)))
        return p.value, n.value, cap.value

    def device_shadowmap(self, light):
        p, n = C.c_uint64(), C.c_uint64()
        self._check(self.L.prc_device_shadowmap(self.h, light, C.byref(p), C.byref(n)))
        return p.value, n.value

    def render_shadows(self, fd, light_mask, row0, row1):
        self._check(self.L.prc_render_shadows(self.h, C.byref(fd.struct), light_mask, row0, row1))

    def render_shadow_units(self, fd, units):
        """units: [(light, row0, row1)] — one fused sweep per 8 units."""
        n = len(units)
        if not n:
            return
        a = np.array(units, dtype=np.uint32).reshape(n, 3)
        li, r0, r1 = (np.ascontiguousarray(a[:, k]) for k in range(3))
        self._check(self.L.prc_render_shadow_units(self.h, C.byref(fd.struct), n, li.ctypes.data, r0.ctypes.data, r1.ctypes.data))

    def render_main(self, fd, out):
        self._check(self.L.prc_render_main(self.h, C.byref(fd.struct), out.ctypes.data if out is not None else None))

    def render_forward(self, fd):
        self._check(self.L.prc_render_forward(self.h, C.byref(fd.struct)))

    def render_deferred(self, fd, out):
        self._check(self.L.prc_render_deferred(self.h, C.byref(fd.struct), out.ctypes.data if out is not None else None))

    def device_shadow_all(self):
        p, n, cap = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self.L.prc_device_shadow_all(self.h, C.byref(p), C.byref(n), C.byref(cap)))
        return p.value, n.value, cap.value

    def stream(self):
        s = C.c_uint64()
        self._check(self.L.prc_stream(self.h, C.byref(s)))
        return s.value

    def sync(self):
        self._check(self.L.prc_sync(self.h))

    # ---- frames over NVLink peer memory (include/polyred_cuda.h: prc_render_peer) ----
    def peer_export(self, fd) -> bytes:
        """This context's exchange handle as bytes (to be gathered from every rank by the host)."""
        h = A.prc_peer_handle()
        self._check(self.L.prc_peer_export(self.h, C.byref(fd.struct), C.byref(h)))
        return bytes(h)

    def peer_connect(self, rank: int, world: int, handles: list[bytes]):
        arr = (A.prc_peer_handle * world)()
        for k, b in enumerate(handles):
            if len(b) != C.sizeof(A.prc_peer_handle):
                raise PolyredCudaError(A.PRC_ERR_INVALID, "peer handle of the wrong size")
            C.memmove(C.byref(arr[k]), b, len(b))
        self._check(self.L.prc_peer_connect(self.h, rank, world, arr))

    def peer_disconnect(self):
        self._check(self.L.prc_peer_disconnect(self.h))

    def set_host_image(self, address: int | None, nbytes: int = 0):
        """Readback destination of prc_render_peer frames: a caller-owned host image (e.g. shared memory mapped by every rank)."""
        self._check(self.L.prc_set_host_image(self.h, address, nbytes))

    @staticmethod
    def row_arrays(rows):
        """[(row0, row1)] of every rank -> the two contiguous uint32 arrays prc_render_peer takes (build once, submit many frames)."""
        a = np.array(rows, dtype=np.uint32).reshape(len(rows), 2)
        return len(rows), (np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1]))

    def render_peer_arrays(self, fd, n, arrays, image_mask: int = 1):
        r0, r1 = arrays
        self._check(self.L.prc_render_peer(self.h, C.byref(fd.struct), n, r0.ctypes.data, r1.ctypes.data, image_mask))

    def render_peer(self, fd, rows, image_mask: int = 1):
        """Submit one frame of the group without waiting. rows: the strip (row0, row1) of EVERY rank, in rank order (this rank's
        entry must be fd's row0/row1)."""
        n, arrays = self.row_arrays(rows)
        self.render_peer_arrays(fd, n, arrays, image_mask)

    def frame_state(self) -> int:
        out = C.c_uint32(0)
        self._check(self.L.prc_frame_state(self.h, C.byref(out)))
        return int(out.value)

    def set_frame_state(self, state: int):
        self._check(self.L.prc_set_frame_state(self.h, state))

    def render_batch(self, fds, outs=None):
        """prc_render_batch: the views `fds` (FrameDescs) submitted back to back; outs[v] = (H, W, 4) u8 array or None."""
        _render_views(self.L.prc_render_batch, self, fds, outs)

    def covered_pixels(self, w=None, h=None) -> int:
        """Covered pixels (visibility key set) of the last frame rendered by this context."""
        out = C.c_uint64(0)
        self._check(self.L.prc_count_covered(self.h, C.byref(out)))
        return int(out.value)

    def measure_fp32_peak(self) -> float:
        """Measured FP32 FMA throughput of the device, TFLOP/s (the shading kernels' roofline denominator)."""
        out = C.c_double(0.0)
        self._check(self.L.prc_measure_fp32_peak(self.h, C.byref(out)))
        return float(out.value)

    def peer_wait_ms(self) -> dict:
        """PRC_PEER_TRACE=1: where this rank's stream idled for its peers during the frames finished by the last sync()."""
        out = (C.c_float * 4)()
        self._check(self.L.prc_peer_wait_ms(self.h, C.byref(out)))
        return dict(zip(("shadow_rows", "peers_shaded", "image_strips", "image_free"), (float(x) for x in out)))


def _render_views(fn, backend, fds, outs):
    n = len(fds)
    frames = (A.prc_frame * max(1, n))()
    for v, fd in enumerate(fds):
        C.memmove(C.byref(frames[v]), C.byref(fd.struct), C.sizeof(A.prc_frame))
    ptrs = (C.c_void_p * max(1, n))()
    for v in range(n):
        o = outs[v] if outs is not None else None
        ptrs[v] = o.ctypes.data if o is not None else None
    backend._check(fn(backend.h, n, frames, ptrs))


class GroupBackend(Backend):
    """A device group behind the C ABI (include/polyred_cuda.h prc_group_*): one process, one context per device, one submit
    thread per device inside the library. Same surface as CudaBackend; render() produces the 1-GPU frame bit for bit."""

    prefix = "prc_group"

    def __init__(self, devices):
        L = lib()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = L.prc_group_open(devs, len(devices), C.byref(h))
        if rc != 0:
            raise PolyredCudaError(rc, f"prc_group_open(devices={list(devices)}) failed (are the CUDA devices visible? there is no CPU fallback)")
        super().__init__(L, h)
        self.devices = [int(d) for d in devices]

    def size(self) -> int:
        return int(self.L.prc_group_size(self.h))

    def rank(self, k: int) -> "CudaBackend":
        """Rank k's context as a (non-owning) CudaBackend, for the read-only / debug calls."""
        c = C.c_void_p()
        self._check(self.L.prc_group_ctx(self.h, k, C.byref(c)))
        b = CudaBackend.__new__(CudaBackend)
        Backend.__init__(b, self.L, c)
        b.device = self.devices[k]
        b.close = lambda: None  # owned by the group
        return b

    def host_image(self, w, h) -> np.ndarray:
        """The last frame, in place in the group's page-locked double buffer (valid during the next render(), like the reference's)."""
        p, n = C.c_uint64(), C.c_uint64()
        self._check(self.L.prc_group_host_image(self.h, C.byref(p), C.byref(n)))
        key = (p.value, n.value, w, h)
        views = self.__dict__.setdefault("_host_views", {})
        if key not in views:
            buf = (C.c_uint8 * n.value).from_address(p.value)
            views[key] = np.frombuffer(buf, dtype=np.uint8).reshape(h, w, 4)
        return views[key]

    def sync(self):
        self._check(self.L.prc_group_sync(self.h))

    def strips(self):
        n = self.size()
        r0, r1 = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        self._check(self.L.prc_group_strips(self.h, r0.ctypes.data, r1.ctypes.data))
        return list(zip(r0.tolist(), r1.tolist()))

    def timings(self):
        """Rank 0's timings (per-rank: rank(k).timings())."""
        return self.rank(0).timings()

    def read_shadowmap(self, light, w, h):
        return self.rank(0).read_shadowmap(light, w, h)  # every rank holds the merged maps

    def read_gbuffer(self, w, h):
        raise PolyredCudaError(A.PRC_ERR_UNSUPPORTED, "a group frame keeps no G-buffer (PRC_FRAME_KEEP_GBUFFER is a single-context debug option)")

    def render_batch(self, fds, outs=None):
        """prc_group_render_views: view v on device v mod n, every device running prc_render_batch on its share."""
        _render_views(self.L.prc_group_render_views, self, fds, outs)
