"""Shared helpers for the parity tests: run the same host-side Renderer against the CUDA
library and against the CPU oracle and compare every output (north_star parity gate)."""
from __future__ import annotations

import numpy as np

from polyred_b200 import render


def make_renderers(scene, cam, w, h, shadow=False, gamma=False, cuda_device=0, background=(0, 0, 0, 0), msaa=1):
    import oracle_binding as ob
    from polyred_b200._lib import CudaBackend

    opts = [render.Camera(cam), render.Size(w, h), render.Scene(scene), render.ShadowMap(shadow), render.GammaCorrection(gamma),
            render.Background(background), render.MSAA(msaa)]
    r_gpu = render.NewRenderer(*opts, render.CUDA(cuda_device))
    r_cpu = render.NewRenderer(*opts, render._Backend(ob.OracleBackend()))
    return r_gpu, r_cpu


def f32_ulp_diff(a, b):
    """ULP distance between float32 arrays (NaN == NaN, +0 == -0)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ia - ib)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)


def compare_frames(r_gpu, r_cpu, w, h, n_lights_cast=(), exact=True):
    """Renders one frame on both sides; returns a dict of mismatch statistics."""
    img_g = r_gpu.Render(keep_gbuffer=True)
    img_c = r_cpu.Render(keep_gbuffer=True)
    gg = r_gpu._backend.read_gbuffer(w, h)
    gc = r_cpu._backend.read_gbuffer(w, h)
    st = {}
    st["n_px"] = w * h
    st["covered"] = int(gc["ok"].sum())
    st["coverage_xor"] = int((gg["ok"] != gc["ok"]).sum())
    both = (gg["ok"] == 1) & (gc["ok"] == 1)
    st["tri_mismatch"] = int(((gg["tri"] != gc["tri"]) & both).sum())
    st["sub_mismatch"] = int(((gg["sub"] != gc["sub"]) & both).sum())
    same = both & (gg["tri"] == gc["tri"]) & (gg["sub"] == gc["sub"])
    st["depth_max_ulp"] = int(f32_ulp_diff(gg["depth"], gc["depth"])[same].max()) if same.any() else 0
    for k in ("uv", "dudv", "nor", "facenor", "wpos"):
        d = f32_ulp_diff(gg[k], gc[k])[same]
        st[f"{k}_max_ulp"] = int(d.max()) if d.size else 0
        st[f"{k}_n_diff"] = int((d > 0).sum())
    st["mat_mismatch"] = int((gg["mat"] != gc["mat"])[same].sum())
    for li in n_lights_cast:
        sg = r_gpu._backend.read_shadowmap(li, w, h)
        sc = r_cpu._backend.read_shadowmap(li, w, h)
        st[f"shadow{li}_max_ulp"] = int(f32_ulp_diff(sg, sc).max())
        st[f"shadow{li}_written"] = int((sc > 0).sum())
    d = np.abs(img_g.astype(np.int32) - img_c.astype(np.int32)).max(axis=2)
    st["rgba_px_diff"] = int((d > 0).sum())
    st["rgba_px_diff_gt1"] = int((d > 1).sum())
    st["rgba_max_diff"] = int(d.max())
    st["nan_gpu"] = int(r_gpu._backend.timings().n_nan_frags)
    st["nan_cpu"] = int(r_cpu._backend.timings().n_nan_frags)
    st["clipped_gpu"] = int(r_gpu._backend.timings().n_clipped)   # triangles that went through clipTriangle (informational)
    st["clipped_cpu"] = int(r_cpu._backend.timings().n_clipped)
    st["valid_gpu"] = int(r_gpu._backend.timings().n_valid_tris)
    st["valid_cpu"] = int(r_cpu._backend.timings().n_valid_tris)
    return st, img_g, img_c


def assert_bit_exact(st):
    assert st["valid_gpu"] == st["valid_cpu"], st
    assert st["coverage_xor"] == 0, st
    assert st["tri_mismatch"] == 0 and st["sub_mismatch"] == 0, st
    assert st["depth_max_ulp"] == 0, st
    for k in ("uv", "dudv", "nor", "facenor", "wpos"):
        assert st[f"{k}_max_ulp"] == 0, (k, st)
    assert st["mat_mismatch"] == 0, st
    for k, v in st.items():
        if k.startswith("shadow") and k.endswith("max_ulp"):
            assert v == 0, (k, st)


def assert_mixed(st, attr_ulp=8):
    """PRC_FMA=mixed (default): everything that decides coverage, triangle id, depth, perspective-correct UV and the
    shadow maps is bit-exact; the shading-only attributes (normal, face normal, world position) may differ from
    the float64-FMA reference by double-rounding cases of fmaf (a few ulp after normalisation)."""
    assert st["valid_gpu"] == st["valid_cpu"], st
    assert st["coverage_xor"] == 0 and st["tri_mismatch"] == 0 and st["sub_mismatch"] == 0, st
    assert st["depth_max_ulp"] == 0 and st["uv_max_ulp"] == 0 and st["mat_mismatch"] == 0, st
    for k in ("nor", "facenor", "wpos"):
        assert st[f"{k}_max_ulp"] <= attr_ulp, (k, st)
    for k, v in st.items():
        if k.startswith("shadow") and k.endswith("max_ulp"):
            assert v == 0, (k, st)


def assert_north_star_gate(st, tie_budget=0):
    """north_star: coverage + tri-ID bit-exact apart from a stated count of edge-tie pixels, depth
    within 1 ulp, RGBA8 within 1/255 on >= 99.9 % of pixels."""
    assert st["coverage_xor"] <= tie_budget, st
    assert st["tri_mismatch"] <= tie_budget, st
    assert st["depth_max_ulp"] <= 1, st
    assert st["rgba_px_diff_gt1"] <= 0.001 * st["n_px"], st
