// oracle/pr_oracle.cpp — TEST INFRASTRUCTURE ONLY. Not part of the product.
//
// Sequential, bug-for-bug CPU restatement of polyred's CPU render pass
// (render.NewRenderer(render.CPU(), render.Workers(1), render.BatchSize(1)).Render()),
// consuming the same prc_scene / prc_frame descriptors as libpolyred_cuda.so so that the
// parity tests feed both sides identical bytes. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
//
// The Go reference cannot be compiled in this image (no Go toolchain, purego module absent),
// so this oracle is pinned (a) bottom-up by the reference's own known-answer tests
// (tests/golden/kat_*.json extracted from math/*_test.go, buffer/texture_test.go, ...) and
// (b) top-down by the reference's committed renders: MSAA(1) internal/examples/out/ground.png, perspect.png,
// gopher.png, the MSAA(2) bunny.png, shadow.png and dragon.png (supersampling, double-MSAA cull box, imageutil.Resize),
// internal/testdata/render.png (the output of the reference's own TestRender: newscene() at 1920x1080, MSAA(2)) and the
// benchmark's shadow-map dump — see tests/test_oracle_golden.py.
// NOT pinned by any reference fixture: the AO transcendental chain (Atan/Cos/Sin/Pow(.,10000))
// and Log2 in the LOD formula use libm where Go uses its own routines ("parity unpinned" for
// those two).
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (oracle/Makefile).
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../include/polyred_cuda.h"
#include "pr_math.h"

using namespace orc;

namespace {

// buffer.Fragment (buffer/buffer.go:54-58) + primitive.Fragment (geometry/primitive/fragment.go:14-24)
struct Fragment {
  bool ok;
  int64_t X, Y;
  f32 depth, u, v, du, dv;
  Vec4 nor;
  RGBA col;
  int64_t mat;
  Vec4 facenor, wpos;
  int32_t tri, sub;  // bookkeeping for the parity maps (not in the reference struct); tri = index + 1
};

struct Vertex {  // primitive.Vertex (geometry/primitive/vertex.go:15-21)
  Vec4 pos, nor;
  RGBA col;
  Vec2 uv;
};

struct TexLevel { uint32_t w, h; const uint8_t* pix; };

struct Ctx {
  // scene (deep copy so the caller's arrays are borrowed only for the call)
  uint64_t n_tris = 0;
  std::vector<f32> pos, nor, uv;
  std::vector<uint32_t> col;
  std::vector<int32_t> mat;
  std::vector<uint64_t> obj_start;
  std::vector<prc_material> materials;
  std::vector<uint32_t> tex_first;
  std::vector<TexLevel> levels;
  std::vector<uint8_t> tex_data;
  bool has_scene = false;

  // frame state
  int W = 0, H = 0;
  std::vector<uint8_t> host_img[2];      // rgba_out == NULL: the frame is kept here, alternately (prc_host_image semantics)
  int host_cur = 0;
  int msaa = 1;                          // render.MSAA(n): W, H above are already n times render.Size (raster.go:149)
  std::vector<Fragment> frags;           // FragmentBuffer.fragments, stored in SCREEN coords [y*W+x]
  std::vector<std::vector<f32>> shadow;  // shadowInfo.depths per light (render/shadow.go:26-31)
  std::vector<std::atomic<uint32_t>> locks;  // MT baseline only (spinlock per pixel, buffer.go:96)
  int threads = 1;
  std::string err;
  prc_timings tm{};
};

inline bool in_rect(const Ctx& c, int64_t x, int64_t y) { return x >= 0 && x < c.W && y >= 0 && y < c.H; }

Mat4 M(const float* p) { Mat4 m; std::memcpy(m.m, p, 64); return m; }

// ---------- Triangle.IsValid (geometry/primitive/triangle.go:63-80) ----------
bool tri_is_valid(Vec4 p1, Vec4 p2, Vec4 p3) {
  Vec4 p1p2 = sub(p2, p1);
  Vec4 p1p3 = sub(p3, p1);
  if (is_zero(p1p2)) return false;
  if (is_zero(p1p3)) return false;
  f32 d = dot(p1p2, p1p3) / (len(p1p2) * len(p1p3));
  return !ApproxEq(d, 1, Epsilon) && !ApproxEq(d, -1, Epsilon);
}

// ---------- AABB (geometry/primitive/box.go) ----------
struct AABB { Vec3 mn, mx; };
AABB aabb3(Vec4 a, Vec4 b, Vec4 c) {  // NewAABB :15-27
  const f32 F = std::numeric_limits<f32>::max();
  AABB r{Vec3{F, F, F}, Vec3{-F, -F, -F}};
  const Vec4 vs[3] = {a, b, c};
  for (int i = 0; i < 3; i++) {
    r.mn.x = Min2(r.mn.x, vs[i].x); r.mn.y = Min2(r.mn.y, vs[i].y); r.mn.z = Min2(r.mn.z, vs[i].z);
    r.mx.x = Max2(r.mx.x, vs[i].x); r.mx.y = Max2(r.mx.y, vs[i].y); r.mx.z = Max2(r.mx.z, vs[i].z);
  }
  return r;
}
bool aabb_intersect(const AABB& a, const AABB& b) {  // :32-41 (note maxZ uses a.Max.Y — the Z quirk)
  f32 minX = Max2(a.mn.x, b.mn.x);
  f32 minY = Max2(a.mn.y, b.mn.y);
  f32 minZ = Max2(a.mn.z, b.mn.z);
  f32 maxX = Min2(a.mx.x, b.mx.x);
  f32 maxY = Min2(a.mx.y, b.mx.y);
  f32 maxZ = Min2(a.mx.y, b.mx.z);
  return minX <= maxX && minY <= maxY && minZ <= maxZ;
}
bool less_eq(f32 v1, f32 v2) { return ApproxEq(v1, v2, Epsilon) || ApproxLess(v1, v2, Epsilon); }  // :62-64
bool aabb_contains(const AABB& a, Vec4 p) {  // :61-75
  return less_eq(a.mn.x, p.x) && less_eq(a.mn.y, p.y) && less_eq(a.mn.z, p.z) && less_eq(p.x, a.mx.x) &&
         less_eq(p.y, a.mx.y) && less_eq(p.z, a.mx.z);
}

// ---------- clipping (render/clipping.go) ----------
struct Plane { Vec4 pos, nor; };
bool point_in_front(const Plane& p, Vec4 v) { return dot(sub(v, p.pos), p.nor) > 0; }  // :18-20
Vec4 intersect_segment(const Plane& p, Vec4 v0, Vec4 v1) {                               // :22-29
  Vec4 u = sub(v1, v0);
  Vec4 w = sub(v0, p.pos);
  f32 d = dot(p.nor, u);
  f32 n = -dot(p.nor, w);
  f32 s = n / d;
  return add(v0, scale(u, s, s, s, s));
}
int sutherland_hodgman(const Vec4* pts, int n, f32 w, f32 h, Vec4* outp) {  // :31-65
  Plane planes[6] = {
      {Vec4{w, 0, 0, 1}, Vec4{-1, 0, 0, 1}}, {Vec4{0, 0, 0, 1}, Vec4{1, 0, 0, 1}},
      {Vec4{0, h, 0, 1}, Vec4{0, -1, 0, 1}}, {Vec4{0, 0, 0, 1}, Vec4{0, 1, 0, 1}},
      {Vec4{0, 0, 1, 1}, Vec4{0, 0, -1, 1}}, {Vec4{0, 0, -1, 1}, Vec4{0, 0, 1, 1}},
  };
  Vec4 a[16], b[16];
  int na = n;
  for (int i = 0; i < n; i++) a[i] = pts[i];
  for (int k = 0; k < 6; k++) {
    if (na == 0) return 0;
    int nb = 0;
    Vec4 s = a[na - 1];
    for (int i = 0; i < na; i++) {
      Vec4 e = a[i];
      if (point_in_front(planes[k], e)) {
        if (!point_in_front(planes[k], s)) b[nb++] = intersect_segment(planes[k], s, e);
        b[nb++] = e;
      } else if (point_in_front(planes[k], s)) {
        b[nb++] = intersect_segment(planes[k], s, e);
      }
      s = e;
    }
    na = nb;
    for (int i = 0; i < nb; i++) a[i] = b[i];
  }
  for (int i = 0; i < na; i++) outp[i] = a[i];
  return na;
}
// one vertex of clipTriangle's fan (clipping.go:73-155): attributes by screen-space
// barycentrics of the ORIGINAL triangle, Pos.W = 1, colour truncated to u8.
Vertex clip_vertex(Vec4 c, const Vertex& v1, const Vertex& v2, const Vertex& v3) {
  f32 b[3];
  barycoord(Vec2{c.x, c.y}, Vec2{v1.pos.x, v1.pos.y}, Vec2{v2.pos.x, v2.pos.y}, Vec2{v3.pos.x, v3.pos.y}, b);
  Vertex t;
  t.pos = Vec4{b[0] * v1.pos.x + b[1] * v2.pos.x + b[2] * v3.pos.x, b[0] * v1.pos.y + b[1] * v2.pos.y + b[2] * v3.pos.y,
               b[0] * v1.pos.z + b[1] * v2.pos.z + b[2] * v3.pos.z, 1};
  t.uv = Vec2{b[0] * v1.uv.x + b[1] * v2.uv.x + b[2] * v3.uv.x, b[0] * v1.uv.y + b[1] * v2.uv.y + b[2] * v3.uv.y};
  t.nor = Vec4{b[0] * v1.nor.x + b[1] * v2.nor.x + b[2] * v3.nor.x, b[0] * v1.nor.y + b[1] * v2.nor.y + b[2] * v3.nor.y,
               b[0] * v1.nor.z + b[1] * v2.nor.z + b[2] * v3.nor.z, 0};
  t.col = RGBA{go_u8(Clamp(b[0] * (f32)v1.col.r + b[1] * (f32)v2.col.r + b[2] * (f32)v3.col.r, 0, 0xff)),
               go_u8(Clamp(b[0] * (f32)v1.col.g + b[1] * (f32)v2.col.g + b[2] * (f32)v3.col.g, 0, 0xff)),
               go_u8(Clamp(b[0] * (f32)v1.col.b + b[1] * (f32)v2.col.b + b[2] * (f32)v3.col.b, 0, 0xff)),
               go_u8(Clamp(b[0] * (f32)v1.col.a + b[1] * (f32)v2.col.a + b[2] * (f32)v3.col.a, 0, 0xff))};
  return t;
}

struct FrameU {  // uniforms of the forward pass (shader.MVP, shader/mvp.go:7-26)
  Mat4 viewport, viewport_inv, proj_inv, view_inv;
  bool perspect;
};

// ---------- FragmentBuffer.DepthTest + Set (buffer/buffer.go:221-280) ----------
// Sequential: both under the pixel lock, so test-then-set is atomic per fragment.
// MT baseline: the same two steps under a per-pixel spinlock (arrival-order ties, as the
// reference with Workers>1).
struct PixelLock {
  std::atomic<uint32_t>* l;
  explicit PixelLock(std::atomic<uint32_t>* p) : l(p) {
    if (l) {
      uint32_t e = 0;
      while (!l->compare_exchange_weak(e, 1, std::memory_order_acquire)) e = 0;
    }
  }
  ~PixelLock() { if (l) l->store(0, std::memory_order_release); }
};

// ---------- drawClipped (render/raster.go:462-572) ----------
void draw_clipped(Ctx& c, const FrameU& u, const Vertex& t1, const Vertex& t2, const Vertex& t3, const f32 recipw[3],
                  int64_t material_id, int32_t tri, int32_t subidx) {
  Vec4 m1 = apply(apply(apply(t1.pos, u.viewport_inv), u.proj_inv), u.view_inv);  // :467-469, no divide by W
  Vec4 m2 = apply(apply(apply(t2.pos, u.viewport_inv), u.proj_inv), u.view_inv);
  Vec4 m3 = apply(apply(apply(t3.pos, u.viewport_inv), u.proj_inv), u.view_inv);

  AABB bb = aabb3(t1.pos, t2.pos, t3.pos);
  int64_t xmin = go_int(Round(bb.mn.x) - 1);
  int64_t xmax = go_int(Round(bb.mx.x) + 1);
  int64_t ymin = go_int(Round(bb.mn.y) - 1);
  int64_t ymax = go_int(Round(bb.mx.y) + 1);

  Vec4 fN = unit(cross(sub(m2, m1), sub(m3, m1)));  // :479

  // the reference loops the whole AABB and skips !buf.In(x,y) (:483); clamping is equivalent
  if (xmin < 0) xmin = 0;
  if (ymin < 0) ymin = 0;
  if (xmax > c.W - 1) xmax = c.W - 1;
  if (ymax > c.H - 1) ymax = c.H - 1;
  Vec2 a{t1.pos.x, t1.pos.y}, b{t2.pos.x, t2.pos.y}, cc{t3.pos.x, t3.pos.y};
  const bool mt = c.threads > 1;
  for (int64_t x = xmin; x <= xmax; x++) {
    for (int64_t y = ymin; y <= ymax; y++) {
      Vec2 p{(f32)x + 0.5f, (f32)y + 0.5f};
      f32 bc[3];
      barycoord(p, a, b, cc, bc);
      if (bc[0] < -Epsilon || bc[1] < -Epsilon || bc[2] < -Epsilon) continue;  // :491

      f32 z = bc[0] * t1.pos.z + bc[1] * t2.pos.z + bc[2] * t3.pos.z;  // :496
      Fragment& dst = c.frags[(size_t)y * c.W + x];
      {
        PixelLock lk(mt ? &c.locks[(size_t)y * c.W + x] : nullptr);
        // DepthTest buffer.go:279. Multithreaded mode only: at EQUAL depth the fragment drawn earlier wins whatever the arrival
        // order, which is what the sequential pass (the parity oracle) produces — the reference itself is non-deterministic
        // there with Workers > 1 (bug-list 11); the full-size parity test runs this mode and must not depend on thread timing.
        const bool earlier = mt && dst.ok && z == dst.depth && ((int64_t)(tri + 1) * 8 + subidx) < ((int64_t)dst.tri * 8 + dst.sub);
        if (!((!dst.ok) || z > dst.depth || earlier)) continue;
      }
      if (std::isnan(z)) __atomic_fetch_add(&c.tm.n_nan_frags, 1, __ATOMIC_RELAXED);

      f32 wc1 = recipw[0] * bc[0], wc2 = recipw[1] * bc[1], wc3 = recipw[2] * bc[2];
      f32 norm = 1.0f;
      if (u.perspect) norm = 1 / (wc1 + wc2 + wc3);
      f32 uvX = (wc1 * t1.uv.x + wc2 * t2.uv.x + wc3 * t3.uv.x) * norm;
      f32 uvY = (wc1 * t1.uv.y + wc2 * t2.uv.y + wc3 * t3.uv.y) * norm;

      f32 du = 0, dv = 0;
      if (material_id >= 0) {  // :517-535
        Vec2 p1{p.x + 1, p.y};
        Vec2 p2{p.x, p.y + 1};
        f32 bcx[3], bcy[3];
        barycoord(p1, a, b, cc, bcx);
        f32 wc1x = recipw[0] * bcx[0], wc2x = recipw[1] * bcx[1], wc3x = recipw[2] * bcx[2];
        f32 normx = 1 / (wc1x + wc2x + wc3x);
        barycoord(p2, a, b, cc, bcy);
        f32 wc1y = recipw[0] * bcy[0], wc2y = recipw[1] * bcy[1], wc3y = recipw[2] * bcy[2];
        f32 normy = 1 / (wc1y + wc2y + wc3y);
        f32 uvdU = (wc1x * t1.uv.x + wc2x * t2.uv.x + wc3x * t3.uv.x) * normx;
        f32 uvdX = (wc1x * t1.uv.y + wc2x * t2.uv.y + wc3x * t3.uv.y) * normx;
        f32 uvdV = (wc1y * t1.uv.x + wc2y * t2.uv.x + wc3y * t3.uv.x) * normy;
        f32 uvdY = (wc1y * t1.uv.y + wc2y * t2.uv.y + wc3y * t3.uv.y) * normy;
        du = (uvdU - uvX) * (uvdU - uvX) + (uvdX - uvY) * (uvdX - uvY);
        dv = (uvdV - uvX) * (uvdV - uvX) + (uvdY - uvY) * (uvdY - uvY);
      }

      Vec4 n = unit(Vec4{(bc[0] * t1.nor.x + bc[1] * t2.nor.x + bc[2] * t3.nor.x),
                         (bc[0] * t1.nor.y + bc[1] * t2.nor.y + bc[2] * t3.nor.y),
                         (bc[0] * t1.nor.z + bc[1] * t2.nor.z + bc[2] * t3.nor.z), 0});
      Vec4 wp{bc[0] * m1.x + bc[1] * m2.x + bc[2] * m3.x, bc[0] * m1.y + bc[1] * m2.y + bc[2] * m3.y,
              bc[0] * m1.z + bc[1] * m2.z + bc[2] * m3.z, 1};  // interpWorldPos :453-460
      RGBA col{
          go_u8(Clamp((wc1 * (f32)t1.col.r + wc2 * (f32)t2.col.r + wc3 * (f32)t3.col.r) * norm, 0, 0xff)),
          go_u8(Clamp((wc1 * (f32)t1.col.g + wc2 * (f32)t2.col.g + wc3 * (f32)t3.col.g) * norm, 0, 0xff)),
          go_u8(Clamp((wc1 * (f32)t1.col.b + wc2 * (f32)t2.col.b + wc3 * (f32)t3.col.b) * norm, 0, 0xff)),
          go_u8(Clamp((wc1 * (f32)t1.col.a + wc2 * (f32)t2.col.a + wc3 * (f32)t3.col.a) * norm, 0, 0xff)),
      };

      PixelLock lk(mt ? &c.locks[(size_t)y * c.W + x] : nullptr);
      if (dst.ok && z <= dst.depth &&
          !(mt && z == dst.depth && ((int64_t)(tri + 1) * 8 + subidx) < ((int64_t)dst.tri * 8 + dst.sub)))
        continue;  // Set buffer.go:230 (+ the deterministic tie rule of the multithreaded mode, see above)
      dst.ok = true;
      dst.X = x; dst.Y = y;
      dst.depth = z; dst.u = uvX; dst.v = uvY; dst.du = du; dst.dv = dv;
      dst.nor = n; dst.facenor = fN; dst.wpos = wp; dst.col = col; dst.mat = material_id;
      dst.tri = tri + 1; dst.sub = subidx;
    }
  }
}

struct TriIn { Vertex v[3]; int64_t mat; };
TriIn load_tri(const Ctx& c, uint64_t i) {
  TriIn t;
  for (int k = 0; k < 3; k++) {
    const f32* p = &c.pos[i * 9 + k * 3];
    const f32* n = &c.nor[i * 9 + k * 3];
    t.v[k].pos = Vec4{p[0], p[1], p[2], 1};
    t.v[k].nor = Vec4{n[0], n[1], n[2], 0};
    t.v[k].uv = Vec2{c.uv[i * 6 + k * 2], c.uv[i * 6 + k * 2 + 1]};
    t.v[k].col = unpack(c.col[i * 3 + k]);
  }
  t.mat = c.mat[i];
  return t;
}

// ---------- (*Renderer).draw (render/raster.go:380-444) ----------
void draw(Ctx& c, const FrameU& u, const Mat4& trans, const Mat4& normal, const TriIn& t, int32_t tri) {
  Vertex t1, t2, t3;
  t1.pos = mulv(trans, t.v[0].pos); t1.col = t.v[0].col; t1.uv = t.v[0].uv; t1.nor = apply(t.v[0].nor, normal);
  t2.pos = mulv(trans, t.v[1].pos); t2.col = t.v[1].col; t2.uv = t.v[1].uv; t2.nor = apply(t.v[1].nor, normal);
  t3.pos = mulv(trans, t.v[2].pos); t3.col = t.v[2].col; t3.uv = t.v[2].uv; t3.nor = apply(t.v[2].nor, normal);

  f32 recipw[3] = {1, 1, 1};
  if (u.perspect) { recipw[0] = -1 / t1.pos.w; recipw[1] = -1 / t2.pos.w; recipw[2] = -1 / t3.pos.w; }

  t1.pos = pos(apply(t1.pos, u.viewport));
  t2.pos = pos(apply(t2.pos, u.viewport));
  t3.pos = pos(apply(t3.pos, u.viewport));
  if (cross(sub(t2.pos, t1.pos), sub(t3.pos, t1.pos)).z < 0) return;  // cullBackFace cull.go:26-28

  // :416-423: Max = (MSAA*buf.Dx, MSAA*buf.Dy, 1) where buf is already MSAA times the output size (the double-MSAA quirk)
  AABB vp{Vec3{0, 0, -1}, Vec3{(f32)(c.msaa * c.W), (f32)(c.msaa * c.H), 1}};
  AABB tb = aabb3(t1.pos, t2.pos, t3.pos);
  if (!aabb_intersect(vp, tb)) return;
  if (aabb_contains(vp, t1.pos) && aabb_contains(vp, t2.pos) && aabb_contains(vp, t3.pos)) {
    draw_clipped(c, u, t1, t2, t3, recipw, t.mat, tri, 0);
    return;
  }
  Vec4 pts[3] = {t1.pos, t2.pos, t3.pos};
  Vec4 clips[16];
  int nc = sutherland_hodgman(pts, 3, (f32)(c.msaa * c.W), (f32)(c.msaa * c.H), clips);  // raster.go:438-439
  __atomic_fetch_add(&c.tm.n_clipped, 1, __ATOMIC_RELAXED);  // statistic only (compared with the CUDA path's count in the edge-case tests)
  for (int i = 2; i < nc; i++) {  // clipTriangle fan, clipping.go:73; parent's recipw reused (raster.go:440-443)
    Vertex a = clip_vertex(clips[0], t1, t2, t3);
    Vertex b = clip_vertex(clips[i - 1], t1, t2, t3);
    Vertex d = clip_vertex(clips[i], t1, t2, t3);
    draw_clipped(c, u, a, b, d, recipw, t.mat, tri, i - 1);
  }
}

// ---------- drawDepth (render/shadow.go:152-219) ----------
void draw_depth(Ctx& c, std::vector<f32>& depths, const Mat4& strans, const Mat4& viewport, const TriIn& t) {
  Vec4 p1 = pos(apply(mulv(strans, t.v[0].pos), viewport));
  Vec4 p2 = pos(apply(mulv(strans, t.v[1].pos), viewport));
  Vec4 p3 = pos(apply(mulv(strans, t.v[2].pos), viewport));
  if (cross(sub(p2, p1), sub(p3, p1)).z < 0) return;
  // cullViewFrustum cull.go:15-24: NewAABB((MSAA*W,MSAA*H,1),(0,0,0),(0,0,-1))
  AABB vp{Vec3{0, 0, -1}, Vec3{(f32)(c.msaa * c.W), (f32)(c.msaa * c.H), 1}};
  AABB tb = aabb3(p1, p2, p3);
  if (!aabb_intersect(vp, tb)) return;
  int64_t xmin = go_int(Round(tb.mn.x) - 1);
  int64_t xmax = go_int(Round(tb.mx.x) + 1);
  int64_t ymin = go_int(Round(tb.mn.y) - 1);
  int64_t ymax = go_int(Round(tb.mx.y) + 1);
  if (xmin < 0) xmin = 0;
  if (ymin < 0) ymin = 0;
  if (xmax > c.W - 1) xmax = c.W - 1;
  if (ymax > c.H - 1) ymax = c.H - 1;
  Vec2 a{p1.x, p1.y}, b{p2.x, p2.y}, cc{p3.x, p3.y};
  const bool mt = c.threads > 1;
  for (int64_t x = xmin; x <= xmax; x++) {
    for (int64_t y = ymin; y <= ymax; y++) {
      f32 bc[3];
      barycoord(Vec2{(f32)x + 0.5f, (f32)y + 0.5f}, a, b, cc, bc);
      if (bc[0] < -Epsilon || bc[1] < -Epsilon || bc[2] < -Epsilon) continue;
      f32 z = bc[0] * p1.z + bc[1] * p2.z + bc[2] * p3.z;
      size_t idx = (size_t)x + (size_t)y * c.W;  // no y flip (:210)
      PixelLock lk(mt ? &c.locks[idx] : nullptr);
      if (!(z <= depths[idx])) depths[idx] = z;  // shadowDepthTest :221-228 then store :211-215
    }
  }
}

// ---------- Texture.Query (buffer/texture.go:83-187) ----------
RGBA rgba_at(const TexLevel& t, int64_t x, int64_t y) {  // image.RGBA.RGBAAt: zero outside Rect
  if (x < 0 || y < 0 || x >= (int64_t)t.w || y >= (int64_t)t.h) return RGBA{0, 0, 0, 0};
  const uint8_t* p = t.pix + ((size_t)y * t.w + (size_t)x) * 4;
  return RGBA{p[0], p[1], p[2], p[3]};
}
RGBA query_bilinear(const TexLevel& buf, f32 u, f32 v) {  // :150-187
  int64_t dx = buf.w, dy = buf.h;
  if (dx == 1 && dy == 1) return rgba_at(buf, 0, 0);
  f32 x = u * ((f32)dx - 1);
  f32 y = v * ((f32)dy - 1);
  f32 x0 = Floor(x), y0 = Floor(y);
  int64_t i = go_int(x0), j = go_int(y0);
  RGBA p1 = rgba_at(buf, i, j), p2, p3, p4;
  p2 = (i < dx - 1) ? rgba_at(buf, i + 1, j) : rgba_at(buf, i, j);
  RGBA interpo1 = lerpc(p1, p2, x - x0);
  p3 = (j < dy - 1) ? rgba_at(buf, i, j + 1) : rgba_at(buf, i, j);
  p4 = (i < dx - 1 && j < dy - 1) ? rgba_at(buf, i + 1, j + 1) : rgba_at(buf, i, j);
  RGBA interpo2 = lerpc(p3, p4, x - x0);
  return lerpc(interpo1, interpo2, y - y0);
}
void go_modf(f32 f, f32* ip, f32* fp) {  // math.Modf (math/math.go:113-116)
  double i;
  double fr = std::modf((double)f, &i);
  *ip = (f32)i;
  *fp = (f32)fr;
}
RGBA tex_query(const TexLevel* mip, int nlev, bool use_mipmap, f32 lod, f32 u, f32 v) {  // :83-132
  f32 iu, iv;
  go_modf(u, &iu, &u);
  if (iu != 0 && u == 0) u = 1;
  if (u < 0) u = 1 - u;
  go_modf(v, &iv, &v);
  if (iv != 0 && v == 0) v = 1;
  if (v < 0) v = 1 - v;
  if (!use_mipmap) {  // queryL0 :134-143
    const TexLevel& t = mip[0];
    f32 dx = (f32)t.w, dy = (f32)t.h;
    if (dx == 1 && dy == 1) return rgba_at(t, 0, 0);
    return rgba_at(t, go_int(Floor(u * (dx - 1))), go_int(Floor(v * (dy - 1))));
  }
  if (lod < 0) lod = 0;
  else if (lod >= (f32)nlev) lod = (f32)(nlev - 1);
  if (lod <= 1) return query_bilinear(mip[0], u, v);
  lod -= 1;
  int64_t h = go_int(Floor(lod));
  int64_t l = h + 1;
  if (l >= nlev) return query_bilinear(mip[h], u, v);
  f32 p = lod - (f32)h;
  if (ApproxEq(p, 0, Epsilon)) return query_bilinear(mip[h], u, v);
  RGBA L1 = query_bilinear(mip[h], u, v);
  RGBA L2 = query_bilinear(mip[l], u, v);
  return lerpc(L1, L2, p);
}

// ---------- shader.FragmentShader (shader/blinn_cpu.go:24-106) ----------
RGBA fragment_shader(const Ctx& c, const prc_material& m, const Fragment& info, Vec3 cam, const prc_frame& fr) {
  const TexLevel* mip = &c.levels[c.tex_first[m.texture]];
  int nlev = (int)(c.tex_first[m.texture + 1] - c.tex_first[m.texture]);
  bool use_mip = !(m.flags & PRC_MAT_NO_MIPMAP);
  f32 lod = 0;
  if (use_mip) {
    f32 siz = (f32)mip[0].w * Sqrt(Max2(info.du, info.dv));
    if (siz < 1) siz = 1;
    lod = Log2(siz);
  }
  RGBA col = tex_query(mip, nlev, use_mip, lod, info.u, 1 - info.v);
  if (fr.n_lights == 0) return col;

  f32 LaR = 0, LaG = 0, LaB = 0;
  for (uint32_t e = 0; e < fr.n_ambient; e++) {
    LaR += fr.ambient_intensity[e] * (f32)col.r;
    LaG += fr.ambient_intensity[e] * (f32)col.g;
    LaB += fr.ambient_intensity[e] * (f32)col.b;
  }
  f32 LdR = 0, LdG = 0, LdB = 0, LsR = 0, LsG = 0, LsB = 0;
  Vec4 n = info.nor;
  if (m.flags & PRC_MAT_FLAT_SHADING) n = info.facenor;
  Vec4 x = info.wpos;
  for (uint32_t li = 0; li < fr.n_lights; li++) {
    const prc_light& l = fr.lights[li];
    Vec4 L{0, 0, 0, 0};
    f32 I = 0;
    if (l.kind == PRC_LIGHT_POINT) {
      Vec4 Ldir = sub(Vec4{l.pos[0], l.pos[1], l.pos[2], 1}, x);
      L = unit(Ldir);
      I = l.intensity / len(Ldir);
    } else if (l.kind == PRC_LIGHT_DIRECTIONAL) {
      L = scale(Vec4{l.pos[0], l.pos[1], l.pos[2], 0}, -1, -1, -1, 1);
      I = l.intensity;
    }
    Vec4 V = unit(sub(Vec4{cam.x, cam.y, cam.z, 1}, x));
    Vec4 Hh = unit(add(L, V));
    f32 Ld = Clamp(dot(n, L), 0, 1);
    f32 Ls = Pow(Clamp(dot(n, Hh), 0, 1), m.shininess);
    LdR += Ld * (f32)col.r * I;
    LdG += Ld * (f32)col.g * I;
    LdB += Ld * (f32)col.b * I;
    RGBA lc = unpack(l.color_rgba);
    LsR += Ls * (f32)lc.r * I;
    LsG += Ls * (f32)lc.g * I;
    LsB += Ls * (f32)lc.b * I;
  }
  RGBA D = unpack(m.diffuse_rgba), S = unpack(m.specular_rgba);
  f32 r = Round(LaR + ((f32)D.r * LdR / 255.0f) + ((f32)S.r * LsR / 255.0f));
  f32 g = Round(LaG + ((f32)D.g * LdG / 255.0f) + ((f32)S.g * LsG / 255.0f));
  f32 b = Round(LaB + ((f32)D.b * LdB / 255.0f) + ((f32)S.b * LsB / 255.0f));
  return RGBA{go_u8(Clamp(r, 0, 0xff)), go_u8(Clamp(g, 0, 0xff)), go_u8(Clamp(b, 0, 0xff)), go_u8(Clamp((f32)col.a, 0, 0xff))};
}

// ---------- shadingVisibility (render/shadow.go:230-283) ----------
bool shading_visibility(const Ctx& c, const prc_frame& fr, uint32_t li, const Fragment& info) {
  const prc_light& l = fr.lights[li];
  if (!l.cast_shadow) return true;
  Vec4 sc = pos(apply(apply(apply(apply(Vec4{(f32)info.X, (f32)info.Y, info.depth, 1}, M(fr.viewport_to_world)), M(l.view)),
                            M(l.proj)), M(fr.viewport)));
  int64_t lightX = go_int(sc.x), lightY = go_int(sc.y);
  int64_t bufIdx = lightX + lightY * (int64_t)c.W;  // wraps like Go int arithmetic (two's complement)
  f32 shadow = 0;
  const std::vector<f32>& d = c.shadow[li];
  if (bufIdx > 0 && bufIdx < (int64_t)d.size()) {
    f32 shadowZ = d[bufIdx];
    const f32 bias = 0.03f;
    if (sc.z < shadowZ - bias) shadow++;
  }
  return shadow > 0;
}

// ---------- material.AmbientOcclusionShade (material/ao.go:20-73) ----------
const Fragment kEmpty{};
const Fragment& buf_get(const Ctx& c, int64_t x, int64_t y) {  // FragmentBuffer.Get buffer.go:209-219
  if (!in_rect(c, x, y)) return kEmpty;
  return c.frags[(size_t)y * c.W + x];
}
f32 max_elevation_angle(const Ctx& c, const Fragment& info, f32 dirX, f32 dirY) {
  Vec4 p{(f32)info.X, (f32)info.Y, 0, 1};
  Vec4 dir{dirX, dirY, 0, 0};
  f32 maxangle = 0;
  for (f32 t = 0; t < 100; t += 1) {
    Vec4 cur = add(p, scale(dir, t, t, 1, 1));
    if (!in_rect(c, go_int(cur.x), go_int(cur.y))) return maxangle;
    f32 distance = len(sub(p, cur));
    if (distance < 1) continue;
    const Fragment& shadeInfo = buf_get(c, go_int(cur.x), go_int(cur.y));
    const Fragment& traceInfo = buf_get(c, go_int(p.x), go_int(p.y));
    f32 shadeDepth = shadeInfo.depth, traceDepth = traceInfo.depth;
    if (!shadeInfo.ok) shadeDepth = -1;
    if (!traceInfo.ok) traceDepth = -1;
    f32 elevation = shadeDepth - traceDepth;
    maxangle = Max2(maxangle, Atan(elevation / distance));
  }
  return maxangle;
}
// Go untyped-constant arithmetic is exact and rounds once to float32:
//   math.Pi/4, math.HalfPi, (math.Pi/2)*8 are power-of-two scalings of float32(Pi);
//   math.TwoPi-1e-4 rounded to float32 is 6.2830853 (0x40c90fce), verified in tests/test_oracle_kat.py.
const f32 kPi = 3.14159265358979323846f;
const f32 kQuarterPi = kPi / 4;   // exact scaling
const f32 kHalfPi = kPi / 2;
const f32 kFourPi = kPi * 4;
const f32 kTwoPiMinus = 6.2830853071795864769f;  // float32(2*Pi - 1e-4)
RGBA ambient_occlusion_shade(const Ctx& c, const Fragment& info, const prc_material* mat) {
  if (mat == nullptr || !(mat->flags & PRC_MAT_AMBIENT_OCCLUSION)) return info.col;
  f32 total = 0;
  for (f32 a = 0; a < kTwoPiMinus; a += kQuarterPi) total += kHalfPi - max_elevation_angle(c, info, Cos(a), Sin(a));
  total /= kFourPi;
  total = Pow(total, 10000);
  return RGBA{go_u8(total * (f32)info.col.r), go_u8(total * (f32)info.col.g), go_u8(total * (f32)info.col.b), info.col.a};
}

const prc_material* mat_at(const Ctx& c, int64_t id) {  // matAt render/gpudeferred.go:77-82
  if (id < 0 || id >= (int64_t)c.materials.size()) return nullptr;
  const prc_material* m = &c.materials[id];
  if (m->flags & PRC_MAT_NIL) return nullptr;
  return m;
}

// ---------- (*Renderer).shade (render/raster.go:324-359) ----------
RGBA shade(Ctx& c, const prc_frame& fr, Fragment& frag /* the pixel's own fragment, mutated: frag.Col */) {
  const Fragment info = buf_get(c, frag.X, frag.Y);  // UnsafeGet(frag.X, frag.Y): pixel (0,0) for uncovered pixels
  if (!info.ok) return unpack(fr.background_rgba);
  RGBA col = info.col;
  const prc_material* mat = mat_at(c, frag.mat);
  if (mat != nullptr) {
    col = fragment_shader(c, *mat, info, Vec3{fr.cam_pos[0], fr.cam_pos[1], fr.cam_pos[2]}, fr);
    if ((fr.flags & PRC_FRAME_SHADOWMAP) && (mat->flags & PRC_MAT_RECEIVE_SHADOW)) {
      f32 visibles = 0;
      for (uint32_t i = 0; i < fr.n_lights; i++)
        if (shading_visibility(c, fr, i, info)) visibles++;
      f32 w = Pow(0.5f, visibles);
      col = RGBA{go_u8((f32)col.r * w), go_u8((f32)col.g * w), go_u8((f32)col.b * w), col.a};
    }
  }
  frag.col = col;
  return ambient_occlusion_shade(c, frag, mat_at(c, frag.mat));
}

uint32_t object_of(const Ctx& c, uint64_t tri) {  // index of the Geometry that owns triangle `tri`
  uint32_t lo = 0, hi = (uint32_t)c.obj_start.size() - 1;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (c.obj_start[mid] <= tri) lo = mid; else hi = mid;
  }
  return lo;
}

void run_parallel(int threads, uint64_t n, uint64_t chunk, const std::function<void(uint64_t, uint64_t)>& fn) {
  if (threads <= 1) { fn(0, n); return; }
  std::atomic<uint64_t> next{0};
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&]() {
      for (;;) {
        uint64_t s = next.fetch_add(chunk);
        if (s >= n) break;
        fn(s, std::min(n, s + chunk));
      }
    });
  for (auto& t : th) t.join();
}
}  // namespace

extern "C" {

struct orc_ctx { Ctx c; };

int32_t orc_open(orc_ctx** out) { *out = new orc_ctx(); return PRC_OK; }
int32_t orc_close(orc_ctx* x) { delete x; return PRC_OK; }
const char* orc_last_error(orc_ctx* x) { return x->c.err.c_str(); }
// threads<=1: sequential (THE parity oracle, Workers(1) order). threads>1: CPU-baseline timing mode.
int32_t orc_set_threads(orc_ctx* x, int32_t threads) { x->c.threads = threads < 1 ? 1 : threads; return PRC_OK; }

int32_t orc_scene_upload(orc_ctx* x, const prc_scene* s) {
  Ctx& c = x->c;
  if (!s || s->abi_version != PRC_ABI_VERSION) { c.err = "bad abi_version"; return PRC_ERR_INVALID; }
  c.n_tris = s->n_tris;
  c.pos.assign(s->pos, s->pos + s->n_tris * 9);
  c.nor.assign(s->nor, s->nor + s->n_tris * 9);
  c.uv.assign(s->uv, s->uv + s->n_tris * 6);
  c.col.assign(s->col, s->col + s->n_tris * 3);
  c.mat.assign(s->mat, s->mat + s->n_tris);
  c.obj_start.assign(s->obj_tri_start, s->obj_tri_start + s->n_objects + 1);
  c.materials.assign(s->materials, s->materials + s->n_materials);
  c.tex_first.assign(s->tex_first_level, s->tex_first_level + s->n_textures + 1);
  c.tex_data.assign(s->tex_data, s->tex_data + s->tex_bytes);
  c.levels.resize(s->n_tex_levels);
  for (uint32_t i = 0; i < s->n_tex_levels; i++) c.levels[i] = TexLevel{s->level_w[i], s->level_h[i], c.tex_data.data() + s->level_offset[i]};
  c.has_scene = true;
  return PRC_OK;
}

int32_t orc_shadow_reset(orc_ctx* x) {
  for (auto& d : x->c.shadow) std::fill(d.begin(), d.end(), 0.0f);
  return PRC_OK;
}

// ---------- imageutil.Resize (internal/imageutil/resize.go:16-164), width and height both given ----------
// createWeights8 (:146-164): int16 coefficients in [-256, 256] of the `linear` kernel (:64-70), float32 arithmetic
static void create_weights8(int dy, int filter_length, f32 scale, std::vector<int16_t>& coeffs, std::vector<int>& start, int& flen) {
  flen = filter_length * (int)go_int(Max2((f32)std::ceil((double)scale), 1.0f));  // math.Ceil via float64 (math/math.go)
  const f32 filter_factor = Min2(1.0f / scale, 1.0f);
  coeffs.assign((size_t)dy * flen, 0);
  start.assign(dy, 0);
  for (int y = 0; y < dy; y++) {
    f32 interp = scale * ((f32)y + 0.5f) - 0.5f;
    start[y] = (int)go_int(interp) - flen / 2 + 1;
    interp -= (f32)start[y];
    for (int i = 0; i < flen; i++) {
      f32 in = (interp - (f32)i) * filter_factor;
      in = Abs(in);
      const f32 k = in <= 1 ? 1 - in : 0;          // linear
      coeffs[(size_t)y * flen + i] = (int16_t)go_int(k * 256);
    }
  }
}
// resizeRGBA (:72-111): `in` has in_w x in_h pixels; output pixel (x, y) of the TRANSPOSED result filters input row x
// around column start[y]; out has out_w = in_h columns and out_h = dy rows
static void resize_pass(const uint8_t* in, int in_w, int in_h, uint8_t* out, int dy, const std::vector<int16_t>& coeffs, const std::vector<int>& start, int flen) {
  const int maxX = in_w - 1;
  for (int x = 0; x < in_h; x++) {
    const uint8_t* row = in + (size_t)x * in_w * 4;
    for (int y = 0; y < dy; y++) {
      int32_t acc[4] = {0, 0, 0, 0}, sum = 0;
      for (int i = 0; i < flen; i++) {
        const int32_t coeff = coeffs[(size_t)y * flen + i];
        if (coeff == 0) continue;
        int xi = start[y] + i;
        if ((unsigned)xi < (unsigned)maxX) xi *= 4;
        else if (xi >= maxX) xi = 4 * maxX;
        else xi = 0;
        for (int ch = 0; ch < 4; ch++) acc[ch] += coeff * (int32_t)row[xi + ch];
        sum += coeff;
      }
      uint8_t* o = out + ((size_t)y * in_h + x) * 4;
      for (int ch = 0; ch < 4; ch++) {
        int v = acc[ch] / sum;  // Go and C both truncate toward zero
        o[ch] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
    }
  }
}
static void resize_rgba(const uint8_t* img, int iw, int ih, uint8_t* out, int ow, int oh) {
  if (ow == iw && oh == ih) { std::memcpy(out, img, (size_t)iw * ih * 4); return; }  // :26-28
  const f32 scaleX = (f32)iw / (f32)ow, scaleY = (f32)ih / (f32)oh;                    // calcFactors :114-131
  std::vector<uint8_t> temp((size_t)ih * ow * 4);                                       // transposed: ih columns, ow rows
  std::vector<int16_t> coeffs; std::vector<int> start; int flen;
  create_weights8(ow, 2, scaleX, coeffs, start, flen);
  resize_pass(img, iw, ih, temp.data(), ow, coeffs, start, flen);
  create_weights8(oh, 2, scaleY, coeffs, start, flen);
  resize_pass(temp.data(), ih, ow, out, oh, coeffs, start, flen);
}
extern "C" int32_t orc_resize(const uint8_t* img, int32_t iw, int32_t ih, uint8_t* out, int32_t ow, int32_t oh) {
  if (!img || !out || iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0) return PRC_ERR_INVALID;
  resize_rgba(img, iw, ih, out, ow, oh);
  return PRC_OK;
}

// (*Renderer).Render (render/raster.go:155-199)
int32_t orc_render(orc_ctx* x, const prc_frame* fr, uint8_t* rgba_out) {
  Ctx& c = x->c;
  if (!fr || fr->abi_version != PRC_ABI_VERSION) { c.err = "bad abi_version"; return PRC_ERR_INVALID; }
  if (!c.has_scene) { c.err = "no scene"; return PRC_ERR_NO_SCENE; }
  if (fr->n_objects + 1 != c.obj_start.size()) { c.err = "n_objects mismatch"; return PRC_ERR_INVALID; }
  auto T0 = std::chrono::steady_clock::now();
  const int W = fr->width, H = fr->height;
  const size_t npx = (size_t)W * H;
  c.msaa = fr->msaa > 1 ? (int)fr->msaa : 1;
  if (W % c.msaa || H % c.msaa) { c.err = "frame size is not a multiple of msaa"; return PRC_ERR_INVALID; }
  if (c.W != W || c.H != H || c.shadow.size() != fr->n_lights) {  // resetBufs + initShadowMaps
    c.W = W; c.H = H;
    c.shadow.assign(fr->n_lights, std::vector<f32>());
  }
  for (uint32_t i = 0; i < fr->n_lights; i++)
    if (fr->lights[i].cast_shadow && c.shadow[i].size() != npx) c.shadow[i].assign(npx, 0.0f);
  // NextBuffer().Clear() (raster.go:202-206); Fragment.tri holds index+1 so the zero value means "none"
  if (c.frags.size() != npx) c.frags.resize(npx);
  run_parallel(c.threads, npx, 1 << 16, [&](uint64_t a, uint64_t b) { std::memset((void*)&c.frags[a], 0, (b - a) * sizeof(Fragment)); });
  if (c.threads > 1) {
    if (c.locks.size() != npx) { std::vector<std::atomic<uint32_t>> l(npx); c.locks.swap(l); }
    run_parallel(c.threads, npx, 1 << 16, [&](uint64_t a, uint64_t b) { for (uint64_t i = a; i < b; i++) c.locks[i].store(0, std::memory_order_relaxed); });
  }
  c.tm = prc_timings{};
  c.tm.abi_version = PRC_ABI_VERSION;

  const Mat4 viewport = M(fr->viewport);
  // --- passShadows (render/shadow.go:92-150), per light in order ---
  if (fr->flags & PRC_FRAME_SHADOWMAP) {
    if (fr->flags & PRC_FRAME_SHADOW_RESET)  // Options() between two views zeroes the maps (render/options.go:125-141)
      for (auto& d : c.shadow) std::fill(d.begin(), d.end(), 0.0f);
    for (uint32_t li = 0; li < fr->n_lights; li++) {
      const prc_light& l = fr->lights[li];
      if (!l.cast_shadow) continue;
      run_parallel(c.threads, c.n_tris, 1024, [&](uint64_t a, uint64_t b) {
        uint32_t o = object_of(c, a);
        Mat4 st = M(l.shadow_trans + (size_t)o * 16);
        for (uint64_t i = a; i < b; i++) {
          while (i >= c.obj_start[o + 1]) { o++; st = M(l.shadow_trans + (size_t)o * 16); }
          TriIn t = load_tri(c, i);
          if (!tri_is_valid(t.v[0].pos, t.v[1].pos, t.v[2].pos)) continue;
          draw_depth(c, c.shadow[li], st, viewport, t);
        }
      });
    }
  }
  auto T1 = std::chrono::steady_clock::now();
  // --- cpuForwardPass (render/raster.go:226-273) ---
  FrameU u{viewport, M(fr->viewport_inv), M(fr->proj_inv), M(fr->view_inv), (fr->flags & PRC_FRAME_PERSPECT) != 0};
  std::atomic<uint64_t> nvalid{0};
  run_parallel(c.threads, c.n_tris, 1024, [&](uint64_t a, uint64_t b) {
    uint32_t o = object_of(c, a);
    Mat4 trans = M(fr->objects[o].trans), normal = M(fr->objects[o].normal);
    uint64_t nv = 0;
    for (uint64_t i = a; i < b; i++) {
      while (i >= c.obj_start[o + 1]) { o++; trans = M(fr->objects[o].trans); normal = M(fr->objects[o].normal); }
      TriIn t = load_tri(c, i);
      if (!tri_is_valid(t.v[0].pos, t.v[1].pos, t.v[2].pos)) continue;
      nv++;
      draw(c, u, trans, normal, t, (int32_t)i);
    }
    nvalid += nv;
  });
  c.tm.n_valid_tris = nvalid;
  auto T2 = std::chrono::steady_clock::now();
  // --- passDeferred (raster.go:275-315) via DrawFragments/DrawFragment (raster_screen.go:17-87) ---
  // Sequential order j = 0..n-1, x = j%w, y = j/w; the shader result replaces fragments[i].Col
  // in place (UnsafeSet), which later pixels observe through pixel (0,0) (bug-list 3).
  std::vector<uint32_t> color(npx);  // FragmentBuffer.color in SCREEN coords
  if (c.threads <= 1) {
    for (size_t j = 0; j < npx; j++) {
      Fragment& f = c.frags[j];
      f.col = shade(c, *fr, f);
      color[j] = pack(f.col);
    }
  } else {
    // MT baseline: pixel (0,0) first (what a sequential run observes), then 32-pixel tasks.
    if (npx) { Fragment& f = c.frags[0]; f.col = shade(c, *fr, f); color[0] = pack(f.col); }
    run_parallel(c.threads, npx, 32, [&](uint64_t a, uint64_t b) {
      for (uint64_t j = (a == 0 ? 1 : a); j < b; j++) {
        Fragment& f = c.frags[j];
        f.col = shade(c, *fr, f);
        color[j] = pack(f.col);
      }
    });
  }
  // --- passAntialiasing (raster.go:361-378): gamma via the host-computed u8 LUT (shader/gamma.go:13-18) ---
  if (fr->flags & PRC_FRAME_GAMMA) {
    run_parallel(c.threads, npx, 4096, [&](uint64_t a, uint64_t b) {
      for (uint64_t j = a; j < b; j++) {
        RGBA p = c.frags[j].col;
        p.r = fr->gamma_lut[p.r]; p.g = fr->gamma_lut[p.g]; p.b = fr->gamma_lut[p.b];
        c.frags[j].col = p;
        color[j] = pack(p);
      }
    });
  }
  auto T3 = std::chrono::steady_clock::now();
  if (fr->flags & PRC_FRAME_BGRA)  // PixelFormatBGRA: Set() stores the colour bytes as B,G,R,A (buffer.go:242-251)
    for (size_t j = 0; j < npx; j++) color[j] = (color[j] & 0xff00ff00u) | ((color[j] & 0xffu) << 16) | ((color[j] >> 16) & 0xffu);
  // buf.Image(): image row r = screen y = H-1-r (buffer.go:160-166, 225)
  if (!rgba_out && !(fr->flags & PRC_FRAME_NO_READBACK)) {  // same contract as prc_render: read the frame through orc_host_image
    c.host_cur ^= 1;
    c.host_img[c.host_cur].resize((size_t)(W / c.msaa) * (H / c.msaa) * 4);
    rgba_out = c.host_img[c.host_cur].data();
  }
  if (rgba_out && c.msaa == 1)
    for (int r = 0; r < H; r++) std::memcpy(rgba_out + (size_t)r * W * 4, &color[(size_t)(H - 1 - r) * W], (size_t)W * 4);
  if (rgba_out && c.msaa > 1) {
    // r.outBuf = imageutil.Resize(cfg.Width, cfg.Height, CurrBuffer().Image()) (raster.go:377)
    std::vector<uint8_t> big((size_t)W * H * 4);
    for (int r = 0; r < H; r++) std::memcpy(big.data() + (size_t)r * W * 4, &color[(size_t)(H - 1 - r) * W], (size_t)W * 4);
    resize_rgba(big.data(), W, H, rgba_out, W / c.msaa, H / c.msaa);
  }
  auto ms = [](auto a, auto b) { return std::chrono::duration<float, std::milli>(b - a).count(); };
  c.tm.shadow_ms = ms(T0, T1); c.tm.forward_ms = ms(T1, T2); c.tm.shade_ms = ms(T2, T3); c.tm.total_ms = ms(T0, T3);
  return PRC_OK;
}

int32_t orc_host_image(orc_ctx* x, uint64_t* host_ptr, uint64_t* bytes) {
  Ctx& c = x->c;
  if (c.host_img[c.host_cur].empty()) return PRC_ERR_INVALID;
  *host_ptr = (uint64_t)(uintptr_t)c.host_img[c.host_cur].data();
  *bytes = c.host_img[c.host_cur].size();
  return PRC_OK;
}

int32_t orc_read_gbuffer(orc_ctx* x, prc_gbuffer_host* g) {
  Ctx& c = x->c;
  size_t n = c.frags.size();
  for (size_t i = 0; i < n; i++) {
    const Fragment& f = c.frags[i];
    if (g->ok) g->ok[i] = f.ok;
    if (g->tri) g->tri[i] = f.ok ? f.tri - 1 : -1;
    if (g->sub) g->sub[i] = f.ok ? f.sub : 0;
    if (g->depth) g->depth[i] = f.depth;
    if (g->uv) { g->uv[2 * i] = f.u; g->uv[2 * i + 1] = f.v; }
    if (g->dudv) { g->dudv[2 * i] = f.du; g->dudv[2 * i + 1] = f.dv; }
    if (g->nor) { g->nor[3 * i] = f.nor.x; g->nor[3 * i + 1] = f.nor.y; g->nor[3 * i + 2] = f.nor.z; }
    if (g->facenor) { g->facenor[3 * i] = f.facenor.x; g->facenor[3 * i + 1] = f.facenor.y; g->facenor[3 * i + 2] = f.facenor.z; }
    if (g->wpos) { g->wpos[3 * i] = f.wpos.x; g->wpos[3 * i + 1] = f.wpos.y; g->wpos[3 * i + 2] = f.wpos.z; }
    // NOTE: after the deferred pass fragments[i].Col holds the SHADED colour (UnsafeSet, raster_screen.go:86)
    if (g->col) g->col[i] = pack(f.col);
    if (g->mat) g->mat[i] = (int32_t)f.mat;
  }
  return PRC_OK;
}

int32_t orc_read_shadowmap(orc_ctx* x, uint32_t light, float* out) {
  Ctx& c = x->c;
  if (light >= c.shadow.size() || c.shadow[light].empty()) { c.err = "no such shadow map"; return PRC_ERR_INVALID; }
  std::memcpy(out, c.shadow[light].data(), c.shadow[light].size() * 4);
  return PRC_OK;
}

int32_t orc_get_timings(orc_ctx* x, prc_timings* t) { *t = x->c.tm; return PRC_OK; }

// ------------------------------------------------------------------------------------------
// Unit entry points, so the reference's own known-answer tests can be replayed on the oracle.
// ------------------------------------------------------------------------------------------
void orc_barycoord(const float p[2], const float t1[2], const float t2[2], const float t3[2], float out[3]) {
  barycoord(Vec2{p[0], p[1]}, Vec2{t1[0], t1[1]}, Vec2{t2[0], t2[1]}, Vec2{t3[0], t3[1]}, out);
}
uint32_t orc_lerpc(uint32_t from, uint32_t to, const float* t) { return pack(lerpc(unpack(from), unpack(to), *t)); }
void orc_mat4_mulm(const float a[16], const float b[16], float out[16]) { Mat4 r = mulm(M(a), M(b)); std::memcpy(out, r.m, 64); }
void orc_mat4_mulv(const float a[16], const float v[4], float out[4]) { Vec4 r = mulv(M(a), Vec4{v[0], v[1], v[2], v[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void orc_vec4_apply(const float v[4], const float a[16], float out[4]) { Vec4 r = apply(Vec4{v[0], v[1], v[2], v[3]}, M(a)); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void orc_vec4_dot(const float v[4], const float u[4], float* out) { *out = dot(Vec4{v[0], v[1], v[2], v[3]}, Vec4{u[0], u[1], u[2], u[3]}); }
void orc_vec4_cross(const float v[4], const float u[4], float out[4]) { Vec4 r = cross(Vec4{v[0], v[1], v[2], v[3]}, Vec4{u[0], u[1], u[2], u[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void orc_vec4_unit(const float v[4], float out[4]) { Vec4 r = unit(Vec4{v[0], v[1], v[2], v[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
void orc_vec4_pos(const float v[4], float out[4]) { Vec4 r = pos(Vec4{v[0], v[1], v[2], v[3]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w; }
int32_t orc_triangle_is_valid(const float p[9]) {
  return tri_is_valid(Vec4{p[0], p[1], p[2], 1}, Vec4{p[3], p[4], p[5], 1}, Vec4{p[6], p[7], p[8], 1});
}
int32_t orc_aabb_intersect(const float a[6], const float b[6]) {
  AABB A{Vec3{a[0], a[1], a[2]}, Vec3{a[3], a[4], a[5]}}, B{Vec3{b[0], b[1], b[2]}, Vec3{b[3], b[4], b[5]}};
  return aabb_intersect(A, B);
}
int32_t orc_aabb_contains(const float a[6], const float p[3]) {
  AABB A{Vec3{a[0], a[1], a[2]}, Vec3{a[3], a[4], a[5]}};
  return aabb_contains(A, Vec4{p[0], p[1], p[2], 1});
}
// Texture.Query on a caller-supplied mip chain
uint32_t orc_texture_query(uint32_t nlev, const uint32_t* w, const uint32_t* h, const uint8_t* const* pix, int32_t use_mipmap,
                           const float* lod, const float* u, const float* v) {
  std::vector<TexLevel> lv(nlev);
  for (uint32_t i = 0; i < nlev; i++) lv[i] = TexLevel{w[i], h[i], pix[i]};
  return pack(tex_query(lv.data(), (int)nlev, use_mipmap != 0, *lod, *u, *v));
}
// FragmentShader on one fragment of the uploaded scene's material `mat`
uint32_t orc_fragment_shader(orc_ctx* x, const prc_frame* fr, int32_t mat, const float nor[3], const float facenor[3],
                             const float wpos[3], const float uvdudv[4], uint32_t col) {
  Fragment f{};
  f.ok = true;
  f.nor = Vec4{nor[0], nor[1], nor[2], 0};
  f.facenor = Vec4{facenor[0], facenor[1], facenor[2], 0};
  f.wpos = Vec4{wpos[0], wpos[1], wpos[2], 1};
  f.u = uvdudv[0]; f.v = uvdudv[1]; f.du = uvdudv[2]; f.dv = uvdudv[3];
  f.col = unpack(col);
  return pack(fragment_shader(x->c, x->c.materials[mat], f, Vec3{fr->cam_pos[0], fr->cam_pos[1], fr->cam_pos[2]}, *fr));
}
// interpWorldPos (render/raster.go:453-460)
void orc_interp_world_pos(const float bc[3], const float m1[4], const float m2[4], const float m3[4], float out[4]) {
  out[0] = bc[0] * m1[0] + bc[1] * m2[0] + bc[2] * m3[0];
  out[1] = bc[0] * m1[1] + bc[1] * m2[1] + bc[2] * m3[1];
  out[2] = bc[0] * m1[2] + bc[1] * m2[2] + bc[2] * m3[2];
  out[3] = 1;
}
// sutherlandHodgman (render/clipping.go:31-65) on a triangle; returns vertex count
int32_t orc_clip_polygon(const float tri[12], const float* w, const float* h, float out[64]) {
  Vec4 pts[3], o[16];
  for (int i = 0; i < 3; i++) pts[i] = Vec4{tri[4 * i], tri[4 * i + 1], tri[4 * i + 2], tri[4 * i + 3]};
  int n = sutherland_hodgman(pts, 3, *w, *h, o);
  for (int i = 0; i < n; i++) { out[4 * i] = o[i].x; out[4 * i + 1] = o[i].y; out[4 * i + 2] = o[i].z; out[4 * i + 3] = o[i].w; }
  return n;
}
void orc_ao_constants(float out[5]) { out[0] = kPi; out[1] = kQuarterPi; out[2] = kHalfPi; out[3] = kFourPi; out[4] = kTwoPiMinus; }

}  // extern "C"
