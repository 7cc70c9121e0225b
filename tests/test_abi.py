"""CPU: libpolyred_cuda.so loads without a GPU, exports every symbol include/polyred_cuda.h declares,
and the ctypes mirror has the same struct layout as the C header (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from polyred_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "polyred_cuda.h")


@pytest.fixture(scope="module")
def lib():
    from polyred_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.dirname(_lib.LIB_PATH), "-s"])
    return _lib.lib()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prc_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 27, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/polyred_cuda.h but not exported"


def test_abi_version(lib):
    assert lib.prc_abi_version() == A.PRC_ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    structs = ["prc_material", "prc_scene", "prc_object_xf", "prc_light", "prc_frame", "prc_gbuffer_host", "prc_timings", "prc_peer_handle"]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "polyred_cuda.h"\nint main(){\n'
    for s in structs:
        prog += f'printf("{s} %zu\\n", sizeof({s}));\n'
        for name, _ in getattr(A, s)._fields_:
            prog += f'printf("{s}.{name} %zu\\n", offsetof({s}, {name}));\n'
    prog += "return 0;}\n"
    c = tmp_path / "layout.c"
    c.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(c)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for s in structs:
        ct = getattr(A, s)
        assert int(out[s]) == C.sizeof(ct), s
        for name, _ in ct._fields_:
            assert int(out[f"{s}.{name}"]) == getattr(ct, name).offset, (s, name)


def test_no_gpu_means_error_not_fallback(lib):
    """north_star: no CPU fallback. Without a CUDA device prc_open fails and the Python option raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from polyred_b200 import render
    from polyred_b200._lib import PolyredCudaError
    assert lib.prc_device_count() == 0
    h = C.c_void_p()
    assert lib.prc_open(0, C.byref(h)) < 0
    with pytest.raises(PolyredCudaError):
        render.NewRenderer(render.CUDA(0))
    with pytest.raises(ValueError):
        render.NewRenderer()  # no backend selected: this package has no CPU renderer


def test_product_code_never_references_the_oracle():
    """The oracle is test infrastructure: nothing under polyred_b200/ may import, load or link it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "polyred_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libpr_oracle" not in txt and "oracle_binding" not in txt and "orc_" not in txt, os.path.join(dirpath, f)
