"""ctypes mirror of include/polyred_cuda.h (structs and constants only, no logic)."""
from __future__ import annotations

import ctypes as C

PRC_ABI_VERSION = 1

PRC_OK = 0
PRC_ERR_INVALID = -1
PRC_ERR_CUDA = -2
PRC_ERR_UNSUPPORTED = -3
PRC_ERR_NO_SCENE = -4
PRC_ERR_NCCL = -5
PRC_ERR_RETRY = -6
PRC_ERR_PEER = -7

PRC_MAT_FLAT_SHADING = 1
PRC_MAT_AMBIENT_OCCLUSION = 2
PRC_MAT_RECEIVE_SHADOW = 4
PRC_MAT_NIL = 8
PRC_MAT_NO_MIPMAP = 16

PRC_LIGHT_POINT = 0
PRC_LIGHT_DIRECTIONAL = 1

PRC_FRAME_PERSPECT = 1
PRC_FRAME_SHADOWMAP = 2
PRC_FRAME_GAMMA = 4
PRC_FRAME_KEEP_GBUFFER = 8
PRC_FRAME_NO_READBACK = 16
PRC_FRAME_UNIFORMS_RESIDENT = 32
PRC_FRAME_SHADOW_RESET = 64
PRC_FRAME_BGRA = 128
PRC_FRAME_ASYNC = 256
PRC_FRAME_NO_KERNEL_TIMERS = 512
PRC_FRAME_IMAGE_AT_SYNC = 1024

F16 = C.c_float * 16
F3 = C.c_float * 3
P = C.POINTER


class prc_material(C.Structure):
    _fields_ = [
        ("diffuse_rgba", C.c_uint32),
        ("specular_rgba", C.c_uint32),
        ("shininess", C.c_float),
        ("texture", C.c_int32),
        ("flags", C.c_uint32),
        ("_pad", C.c_uint32),
    ]


class prc_scene(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("flags", C.c_uint32),
        ("n_tris", C.c_uint64),
        ("pos", P(C.c_float)),
        ("nor", P(C.c_float)),
        ("uv", P(C.c_float)),
        ("col", P(C.c_uint32)),
        ("mat", P(C.c_int32)),
        ("n_objects", C.c_uint32),
        ("n_materials", C.c_uint32),
        ("obj_tri_start", P(C.c_uint64)),
        ("materials", P(prc_material)),
        ("n_textures", C.c_uint32),
        ("n_tex_levels", C.c_uint32),
        ("tex_first_level", P(C.c_uint32)),
        ("level_w", P(C.c_uint32)),
        ("level_h", P(C.c_uint32)),
        ("level_offset", P(C.c_uint64)),
        ("tex_data", P(C.c_uint8)),
        ("tex_bytes", C.c_uint64),
    ]


class prc_object_xf(C.Structure):
    _fields_ = [("trans", F16), ("normal", F16)]


class prc_light(C.Structure):
    _fields_ = [
        ("kind", C.c_uint32),
        ("cast_shadow", C.c_uint32),
        ("pos", F3),
        ("intensity", C.c_float),
        ("color_rgba", C.c_uint32),
        ("_pad", C.c_uint32),
        ("view", F16),
        ("proj", F16),
        ("shadow_trans", P(C.c_float)),
    ]


class prc_frame(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("flags", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("n_objects", C.c_uint32),
        ("n_lights", C.c_uint32),
        ("n_ambient", C.c_uint32),
        ("background_rgba", C.c_uint32),
        ("objects", P(prc_object_xf)),
        ("lights", P(prc_light)),
        ("ambient_intensity", P(C.c_float)),
        ("viewport", F16),
        ("viewport_inv", F16),
        ("proj_inv", F16),
        ("view_inv", F16),
        ("viewport_to_world", F16),
        ("cam_pos", F3),
        ("msaa", C.c_uint32),
        ("gamma_lut", C.c_uint8 * 256),
        ("row0", C.c_uint32),
        ("row1", C.c_uint32),
    ]


class prc_gbuffer_host(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("_pad", C.c_uint32),
        ("ok", P(C.c_uint8)),
        ("tri", P(C.c_int32)),
        ("sub", P(C.c_int32)),
        ("depth", P(C.c_float)),
        ("uv", P(C.c_float)),
        ("dudv", P(C.c_float)),
        ("nor", P(C.c_float)),
        ("facenor", P(C.c_float)),
        ("wpos", P(C.c_float)),
        ("col", P(C.c_uint32)),
        ("mat", P(C.c_int32)),
    ]


class prc_timings(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("_pad", C.c_uint32),
        ("shadow_ms", C.c_float),
        ("forward_ms", C.c_float),
        ("shade_ms", C.c_float),
        ("total_ms", C.c_float),
        ("n_valid_tris", C.c_uint64),
        ("n_nan_frags", C.c_uint64),
        ("gpu_launches", C.c_uint64),
        ("kernel_ms", C.c_float * 8),
        ("kernel_launches", C.c_uint32 * 8),
        ("n_large_items", C.c_uint64),
        ("n_clipped", C.c_uint64),
        ("n_bin_entries", C.c_uint64),
    ]


class prc_peer_handle(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("device", C.c_uint32),
        ("pid", C.c_uint64),
        ("shadow_ptr", C.c_uint64),
        ("image_ptr", C.c_uint64),
        ("signals_ptr", C.c_uint64),
        ("shadow_off", C.c_uint64),
        ("image_off", C.c_uint64),
        ("signals_off", C.c_uint64),
        ("shadow_bytes", C.c_uint64),
        ("image_bytes", C.c_uint64),
        ("shadow_ipc", C.c_uint8 * 64),
        ("image_ipc", C.c_uint8 * 64),
        ("signals_ipc", C.c_uint8 * 64),
        ("mkeys_ptr", C.c_uint64),
        ("mkeys_off", C.c_uint64),
        ("mkeys_bytes", C.c_uint64),
        ("mkeys_ipc", C.c_uint8 * 64),
    ]


KERNEL_CLASSES = ("geom_raster_shadow", "geom_raster_camera", "peer_push", "binning", "medium_raster", "tile_raster", "resolve", "shade")


def pack_rgba(c) -> int:
    r, g, b, a = (int(x) & 0xFF for x in c)
    return r | (g << 8) | (b << 16) | (a << 24)
