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


# ---- the Go shim (go/cuda.go) cannot be compiled here (no Go toolchain): its fixed-layout mirrors are checked by computing the
# ---- layout Go gives them (gc, amd64: natural alignment, fields in declaration order) and comparing with the C compiler's ----
GO_SIZES = {"uint32": (4, 4), "int32": (4, 4), "float32": (4, 4), "uint64": (8, 8), "int64": (8, 8), "uintptr": (8, 8), "unsafe.Pointer": (8, 8),
            "uint8": (1, 1), "byte": (1, 1), "uint16": (2, 2), "int16": (2, 2)}


def _go_structs():
    src = open(os.path.join(ROOT, "go", "cuda.go")).read()
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for name, body in re.findall(r"type\s+(prc\w+)\s+struct\s*\{(.*?)\}", src, flags=re.S):
        fields = []
        for decl in re.split(r"[\n;]", body):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"^([\w\s,]+?)\s+((?:\[\d+\])?[\w.]+)$", decl)
            assert m, (name, decl)
            names, ty = [n.strip() for n in m.group(1).split(",")], m.group(2)
            a = re.match(r"\[(\d+)\](.+)", ty)
            size, align = GO_SIZES[a.group(2) if a else ty]
            if a:
                size *= int(a.group(1))
            for n in names:
                fields.append((n, size, align))
        out[name] = fields
    return out


def _go_layout(fields):
    off, maxal, res = 0, 1, []
    for n, size, align in fields:
        off = (off + align - 1) // align * align
        res.append((n, off, size))
        off += size
        maxal = max(maxal, align)
    return res, (off + maxal - 1) // maxal * maxal


def test_go_shim_struct_layouts_match_header(tmp_path):
    """go/cuda.go had only ever been checked by eye. Every mirrored struct must have the C struct's size and the
    same sequence of field offsets (names differ; blank Go fields are the C padding members), and the frame-flag constants
    (1 << iota) must equal the header's values."""
    pairs = {"prcMaterial": "prc_material", "prcScene": "prc_scene", "prcObjectXf": "prc_object_xf", "prcLight": "prc_light", "prcFrame": "prc_frame"}
    go = _go_structs()
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "polyred_cuda.h"\nint main(){\n'
    for g, c in pairs.items():
        assert g in go, f"go/cuda.go no longer declares {g}"
        prog += f'printf("{c} %zu\\n", sizeof({c}));\n'
        for name, _ in getattr(A, c)._fields_:
            prog += f'printf("{c}.{name} %zu\\n", offsetof({c}, {name}));\n'
    prog += "return 0;}\n"
    src = tmp_path / "golayout.c"
    src.write_text(prog)
    exe = tmp_path / "golayout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for g, c in pairs.items():
        layout, size = _go_layout(go[g])
        assert size == int(out[c]), (g, size, out[c])
        c_offsets = [int(out[f"{c}.{name}"]) for name, _ in getattr(A, c)._fields_]
        go_named = [off for n, off, _ in layout if n != "_"]
        # every C member starts where a Go field starts; Go may split nothing and may only add blank padding fields
        c_named = [o for (name, _), o in zip(getattr(A, c)._fields_, c_offsets) if not name.startswith("_")]
        assert go_named == c_named, (g, go_named, c_named)
    # frame flags: const ( prcFramePerspect = 1 << iota ... )
    gosrc = open(os.path.join(ROOT, "go", "cuda.go")).read()
    block = re.search(r"const \(\s*prcFramePerspect = 1 << iota(.*?)\)", re.sub(r"//[^\n]*", "", gosrc), flags=re.S).group(1)
    names = ["prcFramePerspect"] + re.findall(r"^\s*(prcFrame\w+)", block, flags=re.M)
    want = {"prcFramePerspect": A.PRC_FRAME_PERSPECT, "prcFrameShadowMap": A.PRC_FRAME_SHADOWMAP, "prcFrameGamma": A.PRC_FRAME_GAMMA,
            "prcFrameKeepGBuffer": A.PRC_FRAME_KEEP_GBUFFER, "prcFrameNoReadback": A.PRC_FRAME_NO_READBACK,
            "prcFrameUniformsResident": A.PRC_FRAME_UNIFORMS_RESIDENT, "prcFrameShadowReset": A.PRC_FRAME_SHADOW_RESET,
            "prcFrameBGRA": A.PRC_FRAME_BGRA, "prcFrameAsync": A.PRC_FRAME_ASYNC, "prcFrameNoKernelTimers": A.PRC_FRAME_NO_KERNEL_TIMERS,
            "prcFrameImageAtSync": A.PRC_FRAME_IMAGE_AT_SYNC}
    assert len(names) == len(want), names
    for i, n in enumerate(names):
        assert want[n] == 1 << i, (n, i)


def test_python_constants_match_header():
    """Every PRC_* #define with an integer value in include/polyred_cuda.h equals the constant of the same name in _abi.py."""
    src = open(HEADER).read()
    n = 0
    for name, val in re.findall(r"^#define\s+(PRC_[A-Z0-9_]+)\s+(-?\d+)u?\b", src, flags=re.M):
        if hasattr(A, name):
            assert getattr(A, name) == int(val), name
            n += 1
    assert n >= 20
    assert A.PRC_FRAME_IMAGE_AT_SYNC == 1024


def test_go_shim_binds_only_exported_symbols_with_the_declared_arity(lib):
    """Every symbol name go/cuda.go passes to purego.Dlsym is exported by the library, and every purego.SyscallN call on it passes
    as many arguments as the C declaration has parameters (a typo or a stale call shape would only show at run time in Go)."""
    gosrc = open(os.path.join(ROOT, "go", "cuda.go")).read()
    names = sorted(set(re.findall(r'"(prc_[a-z0-9_]+)"', gosrc)))
    assert len(names) >= 9, names
    for n in names:
        assert hasattr(lib, n), f"go/cuda.go binds {n}, which libpolyred_cuda.so does not export"
    # arity: the C prototypes
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    arity = {}
    for name, params in re.findall(r"\b(prc_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr):
        params = params.strip()
        arity[name] = 0 if params in ("", "void") else params.count(",") + 1
    # which struct field holds which symbol: fnX: sym("prc_...")
    field = dict(re.findall(r"(fn[A-Za-z]+):\s*sym\(\"(prc_[a-z0-9_]+)\"\)", gosrc))
    assert len(field) >= 8, field
    checked = 0
    for m in re.finditer(r"purego\.SyscallN\(\s*(?:b|r\.cuda)\.(fn[A-Za-z]+)", gosrc):
        fn = m.group(1)
        if fn not in field:
            continue
        depth, n_args, i = 1, 0, m.end()  # walk to the parenthesis that closes SyscallN( , counting top-level commas
        while depth > 0:
            ch = gosrc[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
            elif ch == "," and depth == 1:
                n_args += 1
            i += 1
        assert n_args == arity[field[fn]], f"go/cuda.go calls {field[fn]} with {n_args} arguments, the header declares {arity[field[fn]]}"
        checked += 1
    calls = checked
    assert checked >= 6, (checked, calls)
