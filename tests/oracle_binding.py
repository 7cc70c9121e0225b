"""ctypes binding of oracle/libpr_oracle.so — TEST INFRASTRUCTURE ONLY.

The oracle consumes the same prc_scene/prc_frame structs as the CUDA library, so the tests
drive both through the same host code (polyred_b200.render) and compare outputs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from polyred_b200 import _abi as A
from polyred_b200._lib import Backend

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None


def load():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "libpr_oracle.so")
        srcs = [os.path.join(ORACLE_DIR, f) for f in ("pr_oracle.cpp", "pr_math.h")] + [os.path.join(ROOT, "include", "polyred_cuda.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        L = C.CDLL(so)
        vp = C.c_void_p
        L.orc_open.argtypes = [C.POINTER(vp)]
        L.orc_close.argtypes = [vp]
        L.orc_last_error.argtypes = [vp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_threads.argtypes = [vp, C.c_int32]
        L.orc_scene_upload.argtypes = [vp, C.POINTER(A.prc_scene)]
        L.orc_shadow_reset.argtypes = [vp]
        L.orc_render.argtypes = [vp, C.POINTER(A.prc_frame), vp]
        L.orc_read_gbuffer.argtypes = [vp, C.POINTER(A.prc_gbuffer_host)]
        L.orc_read_shadowmap.argtypes = [vp, C.c_uint32, vp]
        L.orc_get_timings.argtypes = [vp, C.POINTER(A.prc_timings)]
        L.orc_lerpc.restype = C.c_uint32
        L.orc_texture_query.restype = C.c_uint32
        L.orc_fragment_shader.restype = C.c_uint32
        L.orc_fragment_shader.argtypes = [vp, C.POINTER(A.prc_frame), C.c_int32, vp, vp, vp, vp, C.c_uint32]
        _LIB = L
    return _LIB


class OracleBackend(Backend):
    prefix = "orc"

    def __init__(self, threads: int = 1):
        L = load()
        h = C.c_void_p()
        assert L.orc_open(C.byref(h)) == 0
        super().__init__(L, h)
        L.orc_set_threads(h, threads)


def fptr(a):
    return a.ctypes.data_as(C.c_void_p)


def f32a(*v):
    return np.array(v, dtype=np.float32)
