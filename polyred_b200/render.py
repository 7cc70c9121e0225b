"""`render.NewRenderer(...).Render()` with a CUDA backend option.

Host-side mirror of polyred's renderer front end (reference: render/raster.go:84-199,
render/options.go:16-141, render/shadow.go:33-90). The options keep the reference's names and
meaning; the one addition is `CUDA(device)`, which mirrors `render.GPU(dev)`
(render/options.go:103-110). Everything per pixel / per triangle happens inside
libpolyred_cuda.so; this module only
  * flattens the scene graph into the SoA triangle soup + flat material table
    (what cpuForwardPass walks every frame, render/raster.go:241-270),
  * computes the per-frame uniforms with the reference's float32 semantics
    (render/raster.go:232-246, 281-295; render/shadow.go:41-86, 121-135),
  * and returns the image (`*image.RGBA` → uint8 [H, W, 4], row 0 = top).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A
from . import gomath as gm
from . import imageutil
from .camera import Orthographic, Perspective
from .light import Directional, Point
from .scene import Scene as _Scene

f32 = np.float32


@dataclass
class _Option:  # render/options.go:16-33
    Width: int = 800
    Height: int = 600
    MSAA: int = 1
    ShadowMap: bool = False
    GammaCorrect: bool = False
    Debug: bool = False
    Scene: _Scene | None = None
    Camera: object = None
    Perspect: bool = False
    Background: tuple = (0, 0, 0, 0)  # transparent black (options.go:21)
    Workers: int = 0
    BatchSize: int = 32
    BlendFunc: object = None
    Format: int = 0
    CUDADevice: int | None = None
    backend: object = None


def Size(w, h):
    return lambda o: (setattr(o, "Width", int(w)), setattr(o, "Height", int(h)))


def Camera(cam):
    def f(o):
        o.Camera = cam
        o.Perspect = isinstance(cam, Perspective)  # options.go:49-56
    return f


def Scene(s):
    return lambda o: setattr(o, "Scene", s)


def ShadowMap(enable):
    return lambda o: setattr(o, "ShadowMap", bool(enable))


def GammaCorrection(enable):
    return lambda o: setattr(o, "GammaCorrect", bool(enable))


def Background(rgba):
    return lambda o: setattr(o, "Background", tuple(rgba))


def MSAA(n):
    return lambda o: setattr(o, "MSAA", int(n))


def Debug(enable):
    return lambda o: setattr(o, "Debug", bool(enable))


def Workers(n):
    return lambda o: setattr(o, "Workers", int(n))


def BatchSize(n):
    return lambda o: setattr(o, "BatchSize", int(n))


def Blending(fn):
    return lambda o: setattr(o, "BlendFunc", fn)


def PixelFormat(fmt):
    return lambda o: setattr(o, "Format", int(fmt))


def CUDA(*devices: int):
    """Backend selection, the analogue of render.GPU(dev) (render/options.go:103-110). Several devices = one frame over all of
    them (a device group inside the library, include/polyred_cuda.h prc_group_*): the same image, bit for bit."""
    devs = [int(d) for d in devices] or [0]
    return lambda o: setattr(o, "CUDADevice", devs[0] if len(devs) == 1 else tuple(devs))


def _Backend(b):
    """Test hook: inject a backend object (tests drive the CPU oracle through the same host code)."""
    return lambda o: setattr(o, "backend", b)


class SceneDesc:
    """Flat upload form of a Scene (prc_scene). Keeps the numpy arrays alive."""

    def __init__(self, scene: _Scene):
        geos = scene.geometries()
        self.geos = geos
        self.n_objects = len(geos)
        counts = [g.pos.shape[0] for g, _ in geos]
        self.obj_start = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
        n = int(self.obj_start[-1])
        cat = lambda xs, shape, dt: (np.ascontiguousarray(np.concatenate(xs), dtype=dt) if xs else np.zeros(shape, dt))
        self.pos = cat([g.pos for g, _ in geos], (0, 3, 3), np.float32)
        self.nor = cat([g.nor for g, _ in geos], (0, 3, 3), np.float32)
        self.uv = cat([g.uv for g, _ in geos], (0, 3, 2), np.float32)
        self.col = cat([g.col for g, _ in geos], (0, 3), np.uint32)
        # flat material table (raster.go:252-262)
        mats, flat = [], []
        for g, _ in geos:
            base = len(mats)
            mats.extend(g.materials)
            flat.append(np.where(g.mat >= 0, g.mat + base, g.mat).astype(np.int32))
        self.mat = cat(flat, (0,), np.int32)
        self.materials = mats
        # textures, deduplicated by identity
        tex_index, textures = {}, []
        for m in mats:
            t = getattr(m, "texture", None)
            if t is not None and id(t) not in tex_index:
                tex_index[id(t)] = len(textures)
                textures.append(t)
        self.mat_arr = (A.prc_material * max(1, len(mats)))()
        for i, m in enumerate(mats):
            pm = self.mat_arr[i]
            if m is None or getattr(m, "texture", None) is None:
                pm.flags = A.PRC_MAT_NIL
                pm.texture = -1
                continue
            pm.diffuse_rgba = A.pack_rgba(m.diffuse)
            pm.specular_rgba = A.pack_rgba(m.specular)
            pm.shininess = float(m.shininess)
            pm.texture = tex_index[id(m.texture)]
            pm.flags = ((A.PRC_MAT_FLAT_SHADING if m.flat_shading else 0) | (A.PRC_MAT_AMBIENT_OCCLUSION if m.ambient_occlusion else 0)
                        | (A.PRC_MAT_RECEIVE_SHADOW if m.receive_shadow else 0) | (0 if m.texture.use_mipmap else A.PRC_MAT_NO_MIPMAP))
        first, lw, lh, off, blobs, cur = [0], [], [], [], [], 0
        for t in textures:
            for lv in t.mipmap:
                lv = np.ascontiguousarray(lv, dtype=np.uint8)
                lh.append(lv.shape[0]); lw.append(lv.shape[1]); off.append(cur)
                blobs.append(lv.reshape(-1)); cur += lv.size
            first.append(len(lw))
        self.tex_first = np.array(first, np.uint32)
        self.level_w = np.array(lw if lw else [0], np.uint32)
        self.level_h = np.array(lh if lh else [0], np.uint32)
        self.level_off = np.array(off if off else [0], np.uint64)
        self.tex_data = np.concatenate(blobs) if blobs else np.zeros(4, np.uint8)
        s = A.prc_scene(abi_version=A.PRC_ABI_VERSION, n_tris=n)
        p = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
        s.pos, s.nor, s.uv = p(self.pos, C.c_float), p(self.nor, C.c_float), p(self.uv, C.c_float)
        s.col, s.mat = p(self.col, C.c_uint32), p(self.mat, C.c_int32)
        s.n_objects, s.n_materials = self.n_objects, len(mats)
        s.obj_tri_start = p(self.obj_start, C.c_uint64)
        s.materials = C.cast(self.mat_arr, C.POINTER(A.prc_material))
        s.n_textures, s.n_tex_levels = len(textures), len(lw)
        s.tex_first_level, s.level_w, s.level_h = p(self.tex_first, C.c_uint32), p(self.level_w, C.c_uint32), p(self.level_h, C.c_uint32)
        s.level_offset, s.tex_data, s.tex_bytes = p(self.level_off, C.c_uint64), p(self.tex_data, C.c_uint8), cur
        self.struct = s
        self.n_tris = n

    def upload_bytes(self):
        return self.pos.nbytes + self.nor.nbytes + self.uv.nbytes + self.col.nbytes + self.mat.nbytes + self.tex_data.nbytes


class FrameDesc:
    """Per-frame uniforms (prc_frame). Keeps the arrays alive."""

    def __init__(self):
        self.struct = A.prc_frame(abi_version=A.PRC_ABI_VERSION)
        self.keep = []


def _m16(m):
    return (C.c_float * 16)(*np.asarray(m, np.float32).reshape(16).tolist())


class Renderer:
    """render.Renderer (render/raster.go:31-61)."""

    def __init__(self, *opts):
        self.cfg = _Option()
        for o in opts:
            o(self.cfg)
        self._scene_desc = None
        self._scene_uploaded_to = None
        self._light_cams = []
        self._backend = self.cfg.backend
        if self._backend is None:
            if self.cfg.CUDADevice is None:
                raise ValueError("render: no backend selected — pass render.CUDA(device); this package has no CPU renderer")
            from ._lib import CudaBackend, GroupBackend
            dev = self.cfg.CUDADevice
            self._backend = GroupBackend(dev) if isinstance(dev, tuple) else CudaBackend(dev)
        self._validate()
        if self.cfg.Scene is not None and self.cfg.ShadowMap:
            self.initShadowMaps()

    def _validate(self):
        c = self.cfg
        if not (1 <= c.MSAA <= 8):
            raise ValueError("render: MSAA must be in 1..8")
        if c.BlendFunc is not None:
            raise NotImplementedError("render: Blending is not on the CUDA path (SURVEY 8f-4)")
        if c.Format not in (0, 1):
            raise ValueError("render: PixelFormat must be PixelFormatRGBA (0) or PixelFormatBGRA (1)")

    # -- render/options.go:125-141
    def Options(self, *opts):
        for o in opts:
            o(self.cfg)
        self._validate()
        self._scene_desc = None
        if self.cfg.Scene is not None and self.cfg.ShadowMap:
            self.initShadowMaps()

    # -- render/shadow.go:33-90
    def initShadowMaps(self):
        c = self.cfg
        sources, _ = c.Scene.Lights()
        self._light_cams = [None] * len(sources)
        center = c.Scene.Center()
        for i, l in enumerate(sources):
            if not l.cast_shadow:
                continue
            from .camera import ViewMatrix
            tm = gm.mulm(gm.mulm(ViewMatrix(l.Position(), center, gm.v3(0, 1, 0)), gm.inv(c.Camera.ViewMatrix())), gm.inv(c.Camera.ProjMatrix()))
            corners = [(1, 1, 1), (1, 1, -1), (1, -1, 1), (-1, 1, 1), (-1, -1, 1), (1, -1, -1), (-1, 1, -1), (-1, -1, -1)]
            vs = np.stack([gm.v4_pos(gm.v4_apply(np.array([x, y, z, 1], np.float32), tm))[:3] for x, y, z in corners])
            mn, mx = vs.min(axis=0), vs.max(axis=0)
            le, ri, bo, to, ne, fa = mn[0], mx[0], mn[1], mx[1], mx[2], mn[2] - f32(2)
            if isinstance(l, Point):
                self._light_cams[i] = Orthographic(position=l.Position(), target=center, up=(0, 1, 0), left=le, right=ri, bottom=bo, top=to, near=ne, far=fa)
            else:
                # shadow.go:67-86: only *light.Point gets a camera; a casting Directional leaves it
                # nil and passShadows panics (shadow.go:123). No defined reference result => error.
                raise NotImplementedError("render: a shadow-casting Directional light has no light camera in the reference (it panics)")
        self._backend_shadow_reset = True

    def InvalidateScene(self):
        """Forget the flattened scene (geometry, materials, model matrices, lights): the next Render() re-flattens
        and re-uploads it. Needed only after editing a vertex / material array IN PLACE: adding or removing objects and
        moving TransformContexts are detected automatically (Scene.signature)."""
        self._scene_desc = None

    # -- flatten + uniforms
    def scene_desc(self) -> SceneDesc:
        """The flattened scene, cached: rebuilt (and re-uploaded) when the Scene object or its membership
        changes; when only a TransformContext moved, just the per-object matrices are recomputed — the reference walks the
        scene graph and rebuilds everything every frame (raster.go:241-270)."""
        sig = self.cfg.Scene.signature()
        sd = self._scene_desc
        if sd is None or self._scene_desc_for is not self.cfg.Scene or sd._membership != sig[0]:
            sd = self._scene_desc = SceneDesc(self.cfg.Scene)
            sd._membership = sig[0]
            self._scene_desc_for = self.cfg.Scene
            self._scene_uploaded_to = None
        if getattr(sd, "_transforms", None) != sig[1]:
            sd._static_xf = None  # recomputed by frame_desc()
            sd._transforms = sig[1]
        return sd

    def frame_desc(self, keep_gbuffer=False, no_readback=False) -> FrameDesc:
        c = self.cfg
        sd = self.scene_desc()
        fd = FrameDesc()
        s = fd.struct
        W, H = c.Width * c.MSAA, c.Height * c.MSAA
        view, proj = c.Camera.ViewMatrix(), c.Camera.ProjMatrix()
        vp = gm.viewport_matrix(W, H)
        view_inv, proj_inv, vp_inv = gm.inv(view), gm.inv(proj), gm.inv(vp)
        nobj = sd.n_objects
        if nobj:
            # Model / Normal depend on the scene graph only: computed once per flattened scene (the reference
            # recomputes them every frame, raster.go:242-243; call InvalidateScene() after moving an object)
            if getattr(sd, "_static_xf", None) is None:
                geos = c.Scene.geometries()  # fresh walk: group transforms above the leaves may have moved (same leaves, Scene.signature)
                chain = np.stack([m for _, m in geos])
                own = np.stack([g.ModelMatrix() for g, _ in geos])
                m_ = gm.mulm(chain, own)                   # raster.go:242
                sd._static_xf = (m_, gm.transpose(gm.inv(m_)))  # raster.go:243
                sd._lights = c.Scene.Lights()
            model, normal = sd._static_xf
            # raster.go:382: Proj.MulM(View).MulM(Model) per object; unchanged while neither the camera nor a model matrix moves
            key = (view.tobytes(), proj.tobytes())
            hit = getattr(sd, "_object_xf", None)
            if hit is not None and hit[0] is model and hit[1] == key:
                flat = hit[2]
            else:
                trans = gm.mulm(gm.mulm(proj, view)[None], model)
                flat = np.ascontiguousarray(np.concatenate([trans.reshape(nobj, 16), normal.reshape(nobj, 16)], axis=1), dtype=np.float32)
                sd._object_xf = (model, key, flat)
        else:
            model = normal = np.zeros((0, 4, 4), np.float32)
            flat = np.zeros((0, 32), np.float32)
            sd._lights = c.Scene.Lights()
        objs = (A.prc_object_xf * max(1, nobj))()
        C.memmove(objs, flat.ctypes.data, flat.nbytes)
        fd.keep.append(objs)
        sources, envs = sd._lights
        lights = (A.prc_light * max(1, len(sources)))()
        for i, l in enumerate(sources):
            pl = lights[i]
            pl.kind = A.PRC_LIGHT_POINT if isinstance(l, Point) else A.PRC_LIGHT_DIRECTIONAL
            v = l.Position() if isinstance(l, Point) else l.direction
            pl.pos = (C.c_float * 3)(*[float(x) for x in v])
            pl.intensity = float(l.intensity)
            pl.color_rgba = A.pack_rgba(l.color)
            pl.cast_shadow = 1 if (l.cast_shadow and c.ShadowMap) else 0
            if c.ShadowMap and l.cast_shadow:
                cam = self._light_cams[i]
                mats = getattr(cam, "_matrices", None)  # a light camera is immutable once initShadowMaps has fitted it
                if mats is None:
                    mats = cam._matrices = (cam.ViewMatrix(), cam.ProjMatrix())
                lv, lp = mats
                pl.view, pl.proj = _m16(lv), _m16(lp)
                # shadow.go:155: lightProj.MulM(lightView).MulM(Model) per object. It depends on the light camera (re-fitted only by
                # NewRenderer / Options) and the model matrices, not on the frame: cached per light until either changes
                cache = sd.__dict__.setdefault("_shadow_trans", {})
                key = (lv.tobytes(), lp.tobytes())
                hit = cache.get(i)
                if hit is not None and hit[0] is model and hit[1] == key:
                    st = hit[2]
                else:
                    st = np.ascontiguousarray(gm.mulm(gm.mulm(lp, lv)[None], model).reshape(nobj, 16), dtype=np.float32)
                    cache[i] = (model, key, st)
                fd.keep.append(st)
                pl.shadow_trans = st.ctypes.data_as(C.POINTER(C.c_float))
        fd.keep.append(lights)
        amb = np.array([e.intensity for e in envs] or [0], np.float32)
        fd.keep.append(amb)
        s.flags = ((A.PRC_FRAME_PERSPECT if c.Perspect else 0) | (A.PRC_FRAME_SHADOWMAP if c.ShadowMap else 0)
                   | (A.PRC_FRAME_GAMMA if c.GammaCorrect else 0) | (A.PRC_FRAME_KEEP_GBUFFER if keep_gbuffer else 0)
                   | (A.PRC_FRAME_NO_READBACK if no_readback else 0) | (A.PRC_FRAME_BGRA if c.Format == 1 else 0))
        s.width, s.height = W, H
        s.n_objects, s.n_lights, s.n_ambient = nobj, len(sources), len(envs)
        s.background_rgba = A.pack_rgba(c.Background)
        s.objects = C.cast(objs, C.POINTER(A.prc_object_xf))
        s.lights = C.cast(lights, C.POINTER(A.prc_light))
        s.ambient_intensity = amb.ctypes.data_as(C.POINTER(C.c_float))
        s.viewport, s.viewport_inv, s.proj_inv, s.view_inv = _m16(vp), _m16(vp_inv), _m16(proj_inv), _m16(view_inv)
        s.viewport_to_world = _m16(gm.mulm(gm.mulm(view_inv, proj_inv), vp_inv))  # raster.go:287
        s.cam_pos = (C.c_float * 3)(*[float(x) for x in c.Camera.Position()])
        lut = imageutil.gamma_lut_u8()
        C.memmove(s.gamma_lut, lut.ctypes.data, 256)
        s.row0, s.row1 = 0, H
        s.msaa = c.MSAA  # W, H above are the supersampled buffer (raster.go:149); the backend resizes to Width x Height (raster.go:377)
        return fd

    def _ensure_uploaded(self):
        sd = self.scene_desc()
        if self._scene_uploaded_to is not self._backend:
            self._backend.scene_upload(sd)
            self._scene_uploaded_to = self._backend
        if getattr(self, "_backend_shadow_reset", False):
            self._backend.shadow_reset()
            self._backend_shadow_reset = False
        return sd

    # -- render/raster.go:155-199
    def Render(self, keep_gbuffer=False) -> np.ndarray:
        c = self.cfg
        if c.Scene is None or c.Camera is None:
            raise ValueError("render: Scene and Camera are required")
        self._ensure_uploaded()
        fd = self.frame_desc(keep_gbuffer=keep_gbuffer)
        if hasattr(self._backend, "host_image"):
            # zero copy: the frame is read in place from the library's page-locked double buffer, which — like the
            # reference's (raster.go:86,201-206) — stays valid until two frames later
            self._backend.render(fd, None)
            out = self._backend.host_image(c.Width, c.Height)
        else:
            out = np.zeros((c.Height, c.Width, 4), np.uint8)
            self._backend.render(fd, out)
        self._last_frame = fd
        if c.Debug:
            self._debug_report()
        return out

    def shadow_map_image(self, index: int) -> np.ndarray:
        """The picture passShadows saves as shadow-<index>.png under render.Debug(true) (render/shadow.go:98-118):
        pixel (i, j) = uint8(depths[i + (H-j-1)*W] * 255) in R, G and B, alpha 255."""
        c = self.cfg
        W, H = c.Width * c.MSAA, c.Height * c.MSAA
        z = self._backend.read_shadowmap(index, W, H)[::-1]  # image row j = map row H-1-j
        g = (z.astype(np.float32) * f32(255)).astype(np.int64).astype(np.uint8)  # Go's uint8(float32): truncation; depths lie in [0, 1]
        return np.stack([g, g, g, np.full_like(g, 255)], axis=2)

    def _debug_report(self, directory: str = "."):
        """render.Debug(true): per-pass timings (profiling.Timed, raster.go:156-161) and one shadow-<i>.png per casting
        light (shadow.go:98-118), written like the reference does into the working directory."""
        import os
        c = self.cfg
        if hasattr(self._backend, "timings"):
            t = self._backend.timings()
            print(f"forward pass (shadow): {t.shadow_ms:.3f} ms\nforward pass (world): {t.forward_ms:.3f} ms\n"
                  f"deferred pass (shading): {t.shade_ms:.3f} ms\nentire rendering: {t.total_ms:.3f} ms")
        if c.ShadowMap:
            from PIL import Image
            sources, _ = c.Scene.Lights()
            for i, l in enumerate(sources):
                if l.cast_shadow:
                    file = os.path.join(directory, f"shadow-{i}.png")
                    print(f"saving (shadow map)... {file}")
                    Image.fromarray(self.shadow_map_image(i), "RGBA").save(file)


def ViewFrames(r: Renderer, cameras) -> list:
    """The per-view uniforms of RenderViews() as FrameDescs, for callers that submit the views back to back
    (tools/multiview_bench.py): each carries PRC_FRAME_SHADOW_RESET, the stream-ordered form of the shadow-map
    zeroing Options() does between views, so no host synchronisation separates two views."""
    out = []
    for cam in cameras:
        sd = r._scene_desc
        r.Options(Camera(cam))
        r._scene_desc = sd
        r._backend_shadow_reset = False
        fd = r.frame_desc()
        fd.struct.flags |= A.PRC_FRAME_SHADOW_RESET
        out.append(fd)
    return out


def RenderViews(r: Renderer, cameras) -> list:
    """BASELINE config 5 (multi-view): the same scene from several cameras. Each view goes through
    Options(Camera(c)) exactly as a reference caller would, which re-fits the light cameras to the new view
    frustum and zeroes the shadow maps (render/options.go:125-141, bug-list 5), then Render(). The flattened
    scene stays resident in HBM across views."""
    if hasattr(r._backend, "render_batch") and not r.cfg.Debug:
        # one call for the whole batch (prc_render_batch / prc_group_render_views): the views are submitted back to back, the
        # zeroing of the shadow maps that Options() does between views travels as PRC_FRAME_SHADOW_RESET
        r._ensure_uploaded()
        fds = ViewFrames(r, cameras)
        outs = [np.zeros((r.cfg.Height, r.cfg.Width, 4), np.uint8) for _ in fds]
        r._backend.render_batch(fds, outs)
        if fds:
            r._last_frame = fds[-1]
        return outs
    out = []
    for cam in cameras:
        sd = r._scene_desc
        r.Options(Camera(cam))
        r._scene_desc = sd          # Options() drops the flattened scene; the geometry did not change
        out.append(r.Render().copy())
    return out


def NewRenderer(*opts) -> Renderer:
    """render.NewRenderer (render/raster.go:84-143)."""
    return Renderer(*opts)
