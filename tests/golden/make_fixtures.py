#!/usr/bin/env python
"""Builds tests/golden/ from the mounted reference (/root/reference). Run in the build container only;
the committed outputs are what the tests read (the GPU box has no /root/reference).

  kat.json            known answers extracted from the reference's own Go tests (regex over the sources):
                      buffer/texture_test.go (24 Texture.Query colours), math/interpolate_test.go,
                      math/mat_test.go, camera/camera_test.go, geometry/primitive/{box,triangle}_test.go
  assets/             the few small input assets those tests and the soft goldens need
  ref_renders/        renders committed in the reference (soft goldens): MSAA(1) ground / perspect / gopher
                      (internal/examples/{ground,perspect,gopher}_test.go), MSAA(2) bunny / shadow / dragon
                      (internal/examples/{bunny,shadow,dragon}_test.go), the benchmark's shadow-map dump and coverage
                      (internal/examples/benchmark/)
  scene_gopher.npz    gopher.obj as flattened by polyred_b200.model.Load (the .obj is 2.6 MB of text)
"""
import json
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference"


def nums(s):
    return [float(x) for x in re.findall(r"(?<![\w.])-?\d+\.?\d*(?:e-?\d+)?(?![\w])", s)]


def extract_kats():
    k = {}
    # --- buffer/texture_test.go:33-150
    src = open(f"{REF}/buffer/texture_test.go").read()
    table = src[src.index("tests = []struct"):src.index("func TestQuery")]
    cases = []
    for m in re.finditer(r'\{\s*"([^"]+)",\s*(buffer\.NewTexture\(\),|buffer\.NewTexture\([^{}]*?\),|mustLoadTexture\("([^"]+)"\),)\s*([-\d.]+),\s*([-\d.]+),\s*([-\d.]+),\s*color\.RGBA\{R: (\d+), G: (\d+), B: (\d+), A: (\d+)\}', table, re.S):
        name, ctor, path, u, v, lod, r, g, b, a = m.groups()
        cases.append({"name": name, "texture": ("default" if ctor == "buffer.NewTexture()," else "2x2" if "TextureImage(data)" in ctor else os.path.basename(path)),
                      "u": float(u), "v": float(v), "lod": float(lod), "want": [int(r), int(g), int(b), int(a)]})
    assert len(cases) == 22, len(cases)
    k["texture_query"] = {"source": "buffer/texture_test.go:33-150", "data_2x2": [255, 255, 255, 255, 0, 0, 0, 0, 0, 0, 0, 0, 255, 255, 255, 255], "cases": cases}
    # --- math/interpolate_test.go:83-115
    k["lerpc"] = {"source": "math/interpolate_test.go:83-91", "from": [0, 0, 0, 255], "to": [255, 255, 255, 255], "t": 0.5, "want": [127, 127, 127, 255]}
    k["barycoord"] = {"source": "math/interpolate_test.go:104-115", "p": [5, 5], "t1": [0, 0], "t2": [20, 0], "t3": [0, 20], "want": [0.5, 0.25, 0.25]}
    # --- math/mat_test.go
    src = open(f"{REF}/math/mat_test.go").read()
    def block(start, end):
        return src[src.index(start):src.index(end, src.index(start))]
    mm = block("func TestMat_MulM", "func TestMat_MulV")
    m4 = mm[mm.index('t.Run("Mat4"'):]
    vals = nums(m4)
    k["mat4_mulm"] = {"source": "math/mat_test.go:528-560", "a": vals[0:16], "b": vals[16:32], "want": vals[32:48]}
    mv = block("func TestMat_MulV", "func TestMat_Det")
    m4 = mv[mv.index('t.Run("Mat4"'):]
    vals = nums(m4)
    k["mat4_mulv"] = {"source": "math/mat_test.go:590-603", "v": vals[0:4], "m": vals[4:20], "want": vals[20:24]}
    md = block("func TestMat_Det", "func TestMat_T")
    m4 = md[md.index('t.Run("Mat4"'):]
    vals = nums(m4)
    k["mat4_det"] = {"source": "math/mat_test.go:638-651", "m": vals[0:16], "want": vals[16]}
    mi = block("func TestMat_Inv", 't.Run("Mat4_Invalid"')
    vals = nums(mi)
    m = vals[0:16]
    rest = vals[16:]
    want = [rest[2 * i] / rest[2 * i + 1] for i in range(16)]
    k["mat4_inv"] = {"source": "math/mat_test.go:702-723 (compared with Mat4.Eq, eps 1e-7)", "m": m, "want": want}
    # --- camera/camera_test.go:106-194
    src = open(f"{REF}/camera/camera_test.go").read()
    tv = src[src.index("func TestViewMatrix"):src.index("func TestProjMatrix")]
    vals = nums(tv[tv.index("want := math.NewMat4"):tv.index("vm := camera.ViewMatrix")])
    k["view_matrix"] = {"source": "camera/camera_test.go:106-125 (Mat4.Eq, eps 1e-7)", "pos": [-550, 194, 734], "target": [-1000, 0, 0], "up": [0, 1, 1], "want": vals[0:16]}
    tp = src[src.index("func TestProjMatrix"):src.index("func BenchmarkCamera")]
    w1 = nums(tp[tp.index("want := math.NewMat4"):tp.index("cp := camera.NewPerspective")])
    w2 = nums(tp[tp.index("want = math.NewMat4"):tp.index("op := camera.NewOrthographic")])
    k["proj_perspective"] = {"source": "camera/camera_test.go:147-170", "fov": 45.0, "aspect": 1.6, "near": -100.0, "far": -600.0, "want": w1[0:16]}
    k["proj_orthographic"] = {"source": "camera/camera_test.go:172-193", "left": -0.5, "right": 0.5, "top": 1.0, "bottom": -0.5, "near": 0.0, "far": -3.0, "want": w2[0:16]}
    # --- geometry/primitive/box_test.go:30-110, triangle_test.go:29-49
    k["aabb"] = {"source": "geometry/primitive/box_test.go:30-64,97-110",
                 "aabb1": [0, 0, 0, 1, 1, 1], "aabb2": [-1, -1, -1, -0.5, -0.5, -0.5], "aabb3": [0, 0, 0, 0.5, 0.5, 0.5], "aabb4": [-1, -1, -1, 0, 0, 0],
                 "intersect": {"aabb2": False, "aabb3": True, "aabb4": True},
                 "contains_true": [[1, 0, 0], [0, 1, 0], [0, 0, 1]], "contains_false": [[-1, -1, -1]]}
    k["triangle_is_valid"] = {"source": "geometry/primitive/triangle_test.go:29-49",
                              "cases": [{"p": [1, 0, 0, 2, 0, 0, 3, 0, 0], "valid": False}, {"p": [1, 0, 0, 2, 0, 0, 0, 1, 0], "valid": True}]}
    k["interp_world_pos"] = {"source": "render/worldpos_test.go:17-32", "m1": [10, 0, 0, 0], "m2": [0, 10, 0, 0], "m3": [0, 0, 10, 0]}
    k["shading_equivalence"] = {
        "source": "render/shading_equiv_test.go:33-107 (FragmentShader vs kernels.Shade within 1 LSB)",
        "texture_rgba": [200, 150, 100, 255], "diffuse": [220, 180, 160, 255], "specular": [255, 255, 255, 255], "shininess": 32,
        "cam": [0, 1.5, 3], "ambient": 0.4,
        "lights": [{"kind": "point", "intensity": 3, "color": [255, 240, 220, 255], "pos": [-2, 3, 4]},
                   {"kind": "directional", "intensity": 1, "color": [180, 200, 255, 255], "dir": [0, -1, -1]}],
        "normals": [[0, 1, 0], [0, 0, 1], [1, 0, 0], [0.577, 0.577, 0.577], [-0.4, 0.8, 0.45], [0.3, -0.2, 0.93]],
        "positions": [[0, 0, 0], [1, 0.5, -1], [-1, 1, 0.5], [0.2, -0.3, 1], [-0.6, 0.1, -0.4], [0.9, 0.9, 0.2]]}
    json.dump(k, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    return k


def copy_assets():
    os.makedirs(os.path.join(HERE, "assets"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "ref_renders"), exist_ok=True)
    for f in ("ground.obj", "ground.mtl", "ground.png", "perspect.obj", "perspect.mtl", "uvgrid2.png", "bunny.obj", "pic.jpg", "dragon.obj"):
        shutil.copy(f"{REF}/internal/testdata/{f}", os.path.join(HERE, "assets", f))
    # bunny.mtl references a 0.7 MB texture the fixtures do not need: keep the material, drop the map
    mtl = open(f"{REF}/internal/testdata/bunny.mtl").read()
    open(os.path.join(HERE, "assets", "bunny.mtl"), "w").write("\n".join(l for l in mtl.splitlines() if not l.startswith("map_Kd")) + "\n")
    # the textured bunny of internal/examples/bunny_test.go (MSAA(2) golden): original .mtl + its 1024x1024 texture
    os.makedirs(os.path.join(HERE, "assets", "bunny_textured"), exist_ok=True)
    for f in ("bunny.obj", "bunny.mtl", "bunny.png"):
        shutil.copy(f"{REF}/internal/testdata/{f}", os.path.join(HERE, "assets", "bunny_textured", f))
    for src, dst in (("examples/out/ground.png", "ground.png"), ("examples/out/perspect.png", "perspect.png"), ("examples/out/gopher.png", "gopher.png"), ("examples/out/bunny.png", "bunny_msaa2.png"), ("examples/out/shadow.png", "shadow_msaa2.png"), ("examples/out/dragon.png", "dragon_msaa2.png"),
                     ("examples/benchmark/shadow-0.png", "benchmark_shadow-0.png"), ("testdata/render.png", "testrender_msaa2.png"),
                     ("examples/out/plane.png", "plane_msaa2.png"), ("examples/out/normalize.png", "normalize_msaa2.png")):
        shutil.copy(f"{REF}/internal/{src}", os.path.join(HERE, "ref_renders", dst))
    # benchmark.png: only its coverage (alpha) is a usable golden (rendered by older shading code)
    from PIL import Image
    a = np.asarray(Image.open(f"{REF}/internal/examples/benchmark/benchmark.png").convert("RGBA"))[..., 3] > 0
    np.savez_compressed(os.path.join(HERE, "ref_renders", "benchmark_coverage.npz"), covered=np.packbits(a), shape=np.array(a.shape))


def gopher_scene():
    from polyred_b200 import model
    g = model.Load(f"{REF}/internal/testdata/gopher.obj")
    out = {"n_geometries": np.array(len(g.objects))}
    for i, geo in enumerate(g.objects):
        out[f"pos{i}"], out[f"nor{i}"], out[f"uv{i}"], out[f"mat{i}"] = geo.pos, geo.nor, geo.uv, geo.mat
        out[f"materials{i}"] = np.array([[*m.diffuse, *m.specular, float(m.shininess), *m.texture.image.reshape(-1)[:4]] for m in geo.materials], np.float32).reshape(-1, 13)
    np.savez_compressed(os.path.join(HERE, "scene_gopher.npz"), **out)


if __name__ == "__main__":
    extract_kats()
    copy_assets()
    gopher_scene()
    print("fixtures written to", HERE)
