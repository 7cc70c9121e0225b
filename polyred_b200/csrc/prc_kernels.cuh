// prc_kernels.cuh — sm_100a kernels of the render pass. See DESIGN.md for the pipeline.
//
// Pipeline of one frame (all on one stream, no host round trip):
//   [k_chunk_cull]           per partial-row view (multi-GPU): which 256-triangle chunks can touch the view's rows
//   k_geom_raster<SHADOW>    all casting lights in one sweep: shared vertices -> screen, cull/classify, candidate-pixel
//                            queue -> atomicMax on the shadow maps; large boxes queued as setup records with a target
//   k_geom_raster<CAMERA>    same for the camera -> visibility keys (depth, ~draw sequence) by 64-bit atomicMax; clip queue
//   k_clip_raster            Sutherland-Hodgman fan for triangles straddling the viewport
//   k_bin_* / k_tile_raster  once per frame for every queued record (camera keys and shadow maps)
//   k_shade_special          pixel (0,0) (the uncovered-pixel quirk)
//   k_resolve_shade          key -> attributes -> Blinn-Phong + shadow lookup + gamma -> RGBA8 (image order), or
//   k_resolve + k_shade      the same through the G-buffer (KEEP_GBUFFER, AO materials, split-phase multi-GPU frames)
#pragma once
#include "../../include/polyred_cuda.h"
#include "prc_math.cuh"
#include "prc_prune.h"
#include "prc_pow.h"

namespace prc {

#define PRC_TILE 16
#define PRC_SMALL_MAX_PIXELS 16  // pixel-box area up to which a triangle goes through the CTA candidate queue (larger: tile path)

struct DevScene {
  const float* pos;
  const float* nor;
  const float* uv;
  const uint32_t* col;
  const int32_t* mat;
  const uint32_t* meta;  // bit31 = !IsValid, bits 0..23 = object index
  // chunk-local shared vertices (k_chunk_dedupe): chunk c = triangles [256c, 256c+256) owns the distinct (position, object)
  // pairs cverts[cvoff[c] .. cvoff[c+1]) = (x, y, z, object bits); lidx[tri] = three 10-bit indices into them
  // (0xFFFFFFFF for a triangle failing IsValid)
  const float4* cverts;
  const uint32_t* cvoff;
  const uint32_t* lidx;
  uint64_t n_tris;
  const prc_material* mats;
  uint32_t n_mats;
  const uint32_t* tex_first;
  const uint32_t* level_w;
  const uint32_t* level_h;
  const uint64_t* level_off;
  const uint8_t* tex_data;
};

struct DevLight {
  uint32_t kind, cast_shadow;
  float pos[3];
  float intensity;
  uint32_t color;
  float colf[3];  // the colour channels as float32 (converted once on the host instead of per light per pixel)
  float view[16], proj[16];
  uint32_t pm_view, pm_proj;  // plain masks (see apply4m)
  uint32_t persp_cam;         // rigid view with 1 = perspective, 2 = orthographic projection, all entries finite (0 = anything else)
  float* shadow_map;  // W*H floats (persistent)
};

struct DevFrame {
  int W, H;        // frame buffer = render.Size x MSAA
  float cullW, cullH;  // MSAA*W, MSAA*H: the box of the viewport cull / clip tests (the double-MSAA quirk, raster.go:416-423,438-439, cull.go:17)
  int row0, row1;  // rows shaded by this context
  int rr0, rr1;    // rows rasterised (row0/row1 widened by the AO halo)
  uint32_t flags;
  uint32_t n_lights, n_ambient;
  uint32_t background;
  float viewport[16], viewport_inv[16], proj_inv[16], view_inv[16], vtw[16];
  uint32_t pm_viewport, pm_viewport_inv, pm_proj_inv, pm_view_inv, pm_vtw;  // plain masks (see apply4m)
  uint32_t unproj_std;  // viewport_inv = [a 0 0 b; 0 c 0 d; 0 0 1 0; 0 0 0 1], proj_inv = [e 0 0 0; 0 f 0 0; 0 0 0 g; 0 0 h i], view_inv = [R t; 0 0 0 1], all finite
  uint32_t vp_std;  // viewport == [a 0 0 b; 0 c 0 d; 0 0 1 0; 0 0 0 1] (math.ViewportMatrix, math/math.go:270-277)
  float cam[3];
  const prc_object_xf* xf;   // per object trans / normal
  const DevLight* lights;
  const float* ambient;
  const uint8_t* gamma;
};

struct LargeRec {  // setup record of a triangle handed to the tile path (52 B)
  float x1, y1, z1, x2, y2, z2, x3, y3, z3;
  uint32_t seq;
  short bx0, by0, bx1, by1;  // clamped pixel bbox, inclusive
  uint32_t target;           // 0 = camera visibility keys, 1 + k = shadow map of the k-th casting light
};

struct Counters {
  // per raster pass (reset by the host with a stream-ordered memset of the first 16 bytes)
  unsigned int n_large;
  unsigned int n_clip;
  unsigned int n_bin_total;
  unsigned int n_huge;          // queued records too large for k_medium_raster (they need the binned tile path)
  // per frame, also reset between frames submitted back to back: length of the compacted chunk list of the frame's k-th raster pass
  unsigned int n_list[16];
  // per frame
  unsigned int large_overflow;  // a queue was too small: the frame is invalid and is re-rendered after growing
  unsigned int max_bins;        // largest n_bin_total of the frame (to size the bin array)
  unsigned int need_bins;       // a huge record was queued while the binned tile path was switched off: re-render with it
  unsigned int _pad;
  unsigned long long n_nan;         // NaN-depth fragments of the camera pass (bug-list 8; resolved by the first-fragment rule in NaN mode)
  unsigned long long n_nan_shadow;  // NaN-depth fragments of the shadow passes (counted and dropped: the one stated deviation)
  unsigned long long stat_large, stat_clip, stat_bins;
  // per scene
  unsigned long long n_valid;
  // Peer groups: row flags [1 + casting lights][H] (target 0 = camera keys) set by whatever writes a fragment into this rank's
  // private buffers, so that k_peer_push only scans (and clears) rows that hold something. nullptr outside a peer group.
  unsigned char* dirty;
  int dirty_h;
};
// Flags of the private buffers of a peer-group rank: one per 256-pixel SEGMENT of a row, and one flag per 32-byte sector — with
// the flags packed (H bytes = 17 cache lines for a 4K frame) every fragment of a pass stored into the same few L2 lines and those
// stores serialised (measured at 2 GPUs: the camera pass over half of the triangles took 0.140 ms against 0.174 ms for all of
// them on one GPU; 0.099 ms with one flag per sector).
#define PRC_DIRTY_STRIDE 32
#define PRC_DIRTY_SEG 256
__device__ __forceinline__ int dirty_nseg(int W) { return (W + PRC_DIRTY_SEG - 1) / PRC_DIRTY_SEG; }
__device__ __forceinline__ void mark_dirty(const Counters* cnt, uint32_t target, int x, int y, int W) {  // not for the hot path (a dependent global load)
  unsigned char* d = cnt->dirty;
  if (d) d[(((size_t)target * cnt->dirty_h + y) * dirty_nseg(W) + (x / PRC_DIRTY_SEG)) * PRC_DIRTY_STRIDE] = 1;
}

// warp-aggregated slot reservation: one atomicAdd per warp for all lanes that reach this point together
__device__ __forceinline__ unsigned int warp_push(unsigned int* counter) {
  const unsigned int m = __activemask();
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------------------
// Triangle setup shared by every stage (render/raster.go:380-444, render/shadow.go:152-178)
// ---------------------------------------------------------------------------------------------
struct ScreenTri {
  V4 p1, p2, p3;  // after Apply(Viewport).Pos()
  float cw1, cw2, cw3;  // clip-space W of the three vertices (for recipw)
};

enum { TRI_CULLED = 0, TRI_DIRECT = 1, TRI_CLIP = 2 };

// transform + viewport + back-face + AABB tests. `clip_allowed`=false is the shadow pass (no clipping,
// cullViewFrustum only — render/cull.go:15-24).
template <bool E>
__device__ __forceinline__ int tri_setup(const float* __restrict__ trans, const float* __restrict__ viewport, uint32_t pm_vp, float Wf, float Hf,
                                         const float* p, bool clip_allowed, ScreenTri& s) {
  V4 a = mulv(trans, V4{p[0], p[1], p[2], 1.0f});
  V4 b = mulv(trans, V4{p[3], p[4], p[5], 1.0f});
  V4 c = mulv(trans, V4{p[6], p[7], p[8], 1.0f});
  s.cw1 = a.w; s.cw2 = b.w; s.cw3 = c.w;
  s.p1 = pos4(apply4m<E>(a, viewport, pm_vp));
  s.p2 = pos4(apply4m<E>(b, viewport, pm_vp));
  s.p3 = pos4(apply4m<E>(c, viewport, pm_vp));
  // cullBackFace (render/cull.go:26-28)
  V4 e1 = sub4(s.p2, s.p1), e2 = sub4(s.p3, s.p1);
  if (fma32<E>(e1.x, e2.y, -(e1.y * e2.x)) < 0.0f) return TRI_CULLED;
  // AABB.Intersect with the viewport box (geometry/primitive/box.go:32-41; maxZ uses Max.Y of the receiver)
  float mnx = go_min3(s.p1.x, s.p2.x, s.p3.x), mxx = go_max3(s.p1.x, s.p2.x, s.p3.x);
  float mny = go_min3(s.p1.y, s.p2.y, s.p3.y), mxy = go_max3(s.p1.y, s.p2.y, s.p3.y);
  float mnz = go_min3(s.p1.z, s.p2.z, s.p3.z), mxz = go_max3(s.p1.z, s.p2.z, s.p3.z);
  float minX = go_max(0.0f, mnx), minY = go_max(0.0f, mny), minZ = go_max(-1.0f, mnz);
  float maxX = go_min(Wf, mxx), maxY = go_min(Hf, mxy), maxZ = go_min(Hf, mxz);
  if (!(minX <= maxX && minY <= maxY && minZ <= maxZ)) return TRI_CULLED;
  if (!clip_allowed) return TRI_DIRECT;
  // AABB.Contains for the three vertices (box.go:61-75)
  bool in = true;
  const V4* vs[3] = {&s.p1, &s.p2, &s.p3};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const V4& v = *vs[i];
    in = in && less_eq(0.0f, v.x) && less_eq(0.0f, v.y) && less_eq(-1.0f, v.z) && less_eq(v.x, Wf) && less_eq(v.y, Hf) && less_eq(v.z, 1.0f);
  }
  return in ? TRI_DIRECT : TRI_CLIP;
}

// pixel bbox of drawClipped/drawDepth (raster.go:473-477): int(Round(min)-1) .. int(Round(max)+1), clamped
__device__ __forceinline__ bool pixel_bbox(const V4& p1, const V4& p2, const V4& p3, int W, int r0, int r1, int& x0, int& y0, int& x1, int& y1,
                                           float& mnx, float& mny, float& mxx, float& mxy) {
  mnx = go_min3(p1.x, p2.x, p3.x); mxx = go_max3(p1.x, p2.x, p3.x);
  mny = go_min3(p1.y, p2.y, p3.y); mxy = go_max3(p1.y, p2.y, p3.y);
  long long xa = go_int(roundf(mnx) - 1.0f), xb = go_int(roundf(mxx) + 1.0f);
  long long ya = go_int(roundf(mny) - 1.0f), yb = go_int(roundf(mxy) + 1.0f);
  if (xa < 0) xa = 0;
  if (ya < r0) ya = r0;
  if (xb > W - 1) xb = W - 1;
  if (yb > r1 - 1) yb = r1 - 1;
  x0 = (int)xa; y0 = (int)ya; x1 = (int)xb; y1 = (int)yb;
  return xa <= xb && ya <= yb;
}

// ---- Sutherland-Hodgman in screen space (render/clipping.go:18-65) ----
template <bool E>
__device__ __forceinline__ bool in_front(const V4& ppos, const V4& pnor, const V4& v) { return dot4<E>(sub4(v, ppos), pnor) > 0.0f; }
template <bool E>
__device__ __forceinline__ V4 isect(const V4& ppos, const V4& pnor, const V4& v0, const V4& v1) {
  V4 u = sub4(v1, v0);
  V4 w = sub4(v0, ppos);
  float d = dot4<E>(pnor, u);
  float n = -dot4<E>(pnor, w);
  float s = __fdiv_rn(n, d);
  return add4(v0, V4{u.x * s, u.y * s, u.z * s, u.w * s});
}
template <bool E>
__device__ int clip_polygon(const ScreenTri& t, float Wf, float Hf, V4* out /*[12]*/) {
  V4 a[12], b[12];
  int na = 3;
  a[0] = t.p1; a[1] = t.p2; a[2] = t.p3;
  for (int k = 0; k < 6; k++) {
    V4 ppos, pnor;
    switch (k) {
      case 0: ppos = V4{Wf, 0, 0, 1}; pnor = V4{-1, 0, 0, 1}; break;
      case 1: ppos = V4{0, 0, 0, 1}; pnor = V4{1, 0, 0, 1}; break;
      case 2: ppos = V4{0, Hf, 0, 1}; pnor = V4{0, -1, 0, 1}; break;
      case 3: ppos = V4{0, 0, 0, 1}; pnor = V4{0, 1, 0, 1}; break;
      case 4: ppos = V4{0, 0, 1, 1}; pnor = V4{0, 0, -1, 1}; break;
      default: ppos = V4{0, 0, -1, 1}; pnor = V4{0, 0, 1, 1}; break;
    }
    if (na == 0) return 0;
    int nb = 0;
    V4 s = a[na - 1];
    for (int i = 0; i < na; i++) {
      V4 e = a[i];
      bool ein = in_front<E>(ppos, pnor, e), sin_ = in_front<E>(ppos, pnor, s);
      if (ein) {
        if (!sin_ && nb < 12) b[nb++] = isect<E>(ppos, pnor, s, e);
        if (nb < 12) b[nb++] = e;
      } else if (sin_) {
        if (nb < 12) b[nb++] = isect<E>(ppos, pnor, s, e);
      }
      s = e;
    }
    na = nb;
    for (int i = 0; i < nb; i++) a[i] = b[i];
  }
  for (int i = 0; i < na; i++) out[i] = a[i];
  return na;
}
// position of one fan vertex (clipping.go:80-85): barycentric of the clip point w.r.t. the original triangle
template <bool E>
__device__ __forceinline__ void clip_bary(const ScreenTri& t, const V4& c, float b[3]) {
  BarySetup bs = bary_setup<E>(t.p1.x, t.p1.y, t.p2.x, t.p2.y, t.p3.x, t.p3.y);
  bary_eval<E>(bs, c.x, c.y, b[0], b[1], b[2]);
}
__device__ __forceinline__ V4 clip_pos(const ScreenTri& t, const float b[3]) {
  return V4{b[0] * t.p1.x + b[1] * t.p2.x + b[2] * t.p3.x, b[0] * t.p1.y + b[1] * t.p2.y + b[2] * t.p3.y,
            b[0] * t.p1.z + b[1] * t.p2.z + b[2] * t.p3.z, 1.0f};
}

// ---------------------------------------------------------------------------------------------
// rasterisation of one screen triangle's pixel box by the calling thread (clipped fans, pixel (0,0), queue-overflow fallback)
// ---------------------------------------------------------------------------------------------
// NaN mode (NM, camera pass only; bug-list 8, buffer/buffer.go:230,279): DepthTest passes for ANY depth when the pixel is
// empty and `depth > NaN` is false for every later fragment, so a NaN-depth fragment that is the FIRST fragment of its pixel
// in draw order stays for good, and a NaN-depth fragment that arrives later is dropped. The order-independent form: every
// fragment also does atomicMin(first[pixel], seq << 1 | isnan(depth)); a pixel whose minimum has the NaN bit set shows that
// fragment (k_nan_fix rewrites its key before the resolve). A context enters NaN mode — and re-renders the frame — when a
// frame reported a NaN-depth fragment; frames without one (every benchmark and golden scene) never pay for it.
__device__ __forceinline__ void nan_first(unsigned long long* first, size_t idx, uint32_t seq, float z) {
  atomicMin(&first[idx], ((unsigned long long)seq << 1) | (isnan(z) ? 1ull : 0ull));
}

template <bool E, bool SHADOW, bool NM = false>
__device__ __forceinline__ void raster_one(const BarySetup& bs, const V4& p1, const V4& p2, const V4& p3, int x0, int y0, int x1, int y1, uint32_t seq, int W,
                                           unsigned long long* __restrict__ keys, float* __restrict__ smap, Counters* cnt, unsigned long long* first = nullptr,
                                           uint32_t target = 0) {
  for (int y = y0; y <= y1; y++) {
    float py = (float)y + 0.5f;
    for (int x = x0; x <= x1; x++) {
      float px = (float)x + 0.5f;
      float w1, w2, w3;
      bary_eval<E>(bs, px, py, w1, w2, w3);
      if (w1 < -PRC_EPS || w2 < -PRC_EPS || w3 < -PRC_EPS) continue;
      float z = w1 * p1.z + w2 * p2.z + w3 * p3.z;
      size_t idx = (size_t)y * W + x;
      if (NM && !SHADOW) nan_first(first, idx, seq, z);
      if (isnan(z)) { atomicAdd(SHADOW ? &cnt->n_nan_shadow : &cnt->n_nan, 1ULL); if (NM && !SHADOW) mark_dirty(cnt, 0, x, y, W); continue; }
      mark_dirty(cnt, target, x, y, W);
      if (SHADOW) {
        // shadowDepthTest (shadow.go:221-228): store iff !(z <= stored); stored starts at 0 and only grows,
        // so only z > 0 can ever be stored and positive floats order like their int bits.
        if (z > 0.0f && z > smap[idx]) atomicMax((int*)&smap[idx], __float_as_int(z));
      } else {
        unsigned long long key = ((unsigned long long)depth_key(z) << 32) | (unsigned long long)(0xFFFFFFFFu - seq);
        if (key > keys[idx]) atomicMax(&keys[idx], key);
      }
    }
  }
}

// pixel (0,0) is needed on every rank (uncovered pixels shade G(0,0), raster.go:326): when the raster rows do
// not include row 0 it is evaluated separately.
template <bool E, bool NM = false>
__device__ __forceinline__ void raster_pixel00(const V4& p1, const V4& p2, const V4& p3, uint32_t seq, unsigned long long* keys, Counters* cnt, unsigned long long* first = nullptr) {
  int x0, y0, x1, y1;
  float a, b, c, d;
  if (!pixel_bbox(p1, p2, p3, 1, 0, 1, x0, y0, x1, y1, a, b, c, d)) return;
  BarySetup bs = bary_setup<E>(p1.x, p1.y, p2.x, p2.y, p3.x, p3.y);
  raster_one<E, false, NM>(bs, p1, p2, p3, 0, 0, 0, 0, seq, 1 << 30, keys, nullptr, cnt, first);
}

template <bool E, bool SHADOW, bool NM = false>
__device__ __forceinline__ void emit_tri(const V4& p1, const V4& p2, const V4& p3, uint32_t seq, const DevFrame& F, int r0, int r1,
                                         unsigned long long* keys, float* smap, LargeRec* large, unsigned int large_cap, uint32_t target, Counters* cnt,
                                         unsigned long long* first = nullptr) {
  int x0, y0, x1, y1;
  float mnx, mny, mxx, mxy;
  if (!SHADOW && r0 > 0) raster_pixel00<E, NM>(p1, p2, p3, seq, keys, cnt, first);
  if (!pixel_bbox(p1, p2, p3, F.W, r0, r1, x0, y0, x1, y1, mnx, mny, mxx, mxy)) return;
  const BarySetup bs = bary_setup<E>(p1.x, p1.y, p2.x, p2.y, p3.x, p3.y);
  // exact-safe shrink of the reference's AABB+-1 loop (prc_prune.h)
  if (prune_ok(mnx, mny, mxx, mxy, bs.Sabc)) {
    x0 = max(x0, prune_first(mnx)); x1 = min(x1, prune_last(mxx));
    y0 = max(y0, prune_first(mny)); y1 = min(y1, prune_last(mxy));
    if (x0 > x1 || y0 > y1) return;
  }
  int area = (x1 - x0 + 1) * (y1 - y0 + 1);
  if (area <= PRC_SMALL_MAX_PIXELS) {
    raster_one<E, SHADOW, NM>(bs, p1, p2, p3, x0, y0, x1, y1, seq, F.W, keys, smap, cnt, first, target);
  } else {
    unsigned int slot = warp_push(&cnt->n_large);
    if (slot >= large_cap) { atomicExch(&cnt->large_overflow, 1u); return; }
    LargeRec r;
    r.x1 = p1.x; r.y1 = p1.y; r.z1 = p1.z; r.x2 = p2.x; r.y2 = p2.y; r.z2 = p2.z; r.x3 = p3.x; r.y3 = p3.y; r.z3 = p3.z;
    r.seq = seq; r.bx0 = (short)x0; r.by0 = (short)y0; r.bx1 = (short)x1; r.by1 = (short)y1; r.target = target;
    large[slot] = r;
  }
}

#define PRC_GEOM_THREADS 256
#ifndef PRC_GEOM_MIN_BLOCKS
#define PRC_GEOM_MIN_BLOCKS 8  // 32 registers, full occupancy: measured 1.29 ms/frame vs 1.34 (6 CTAs, 40 regs) and 1.40 (5 CTAs, 48 regs); more resident CTAs hide the phase barriers
#endif
template <bool E, bool SHADOW, bool NM = false>
// NOTE: takes the frame through a pointer to a DEVICE-RESIDENT copy. Passing the kernel-parameter struct by
// reference to a non-inlined function makes every thread copy it to local memory at kernel entry (measured:
// 1.76 GB of DRAM writes per launch).
__device__ __noinline__ void geom_generic(const DevFrame* __restrict__ Fg, const float* trans, const float* __restrict__ pos, uint32_t tri, unsigned long long* keys,
                                          float* smap, LargeRec* large, unsigned int large_cap, uint32_t target, unsigned int* clipq, unsigned int clip_cap, Counters* cnt,
                                          int vr0, int vr1) {
  float p[9];  // re-read from global memory: passing the caller's register array by pointer would force it into local memory
#pragma unroll
  for (int i = 0; i < 9; i++) p[i] = pos[(size_t)tri * 9 + i];
  const DevFrame& F = *Fg;
  ScreenTri st;
  int cls = tri_setup<E>(trans, F.viewport, F.pm_viewport, F.cullW, F.cullH, p, !SHADOW, st);
  if (cls == TRI_CULLED) return;
  if (cls == TRI_CLIP) {
    unsigned int slot = warp_push(&cnt->n_clip);
    if (slot < clip_cap) clipq[slot] = tri;
    else atomicExch(&cnt->large_overflow, 1u);
    return;
  }
  const int r0 = vr0, r1 = vr1;  // (the device-resident frame copy *Fg holds the strip's rows; a launch may cover another range)
  emit_tri<E, SHADOW, NM>(st.p1, st.p2, st.p3, tri * 8u, F, r0, r1, keys, smap, large, large_cap, target, cnt, NM ? keys + (size_t)F.W * F.H : nullptr);
}

// Apply(Viewport).Pos() for the standard viewport matrix [a 0 0 b; 0 c 0 d; 0 0 1 0; 0 0 0 1] and a vertex whose
// clip coordinates are finite with z != 0, w != 0: the zero entries contribute exact zeros, so
//   x' = FMA(a, x, b*w)   y' = FMA(c, y, d*w)   z' = z   w' = w      (math/vec4.go:108-116)
// bit for bit (adding +-0 to a non-zero value is the identity; an exactly cancelling FMA gives +0 either way).
template <bool E>
__device__ __forceinline__ bool viewport_pos_std(const float* __restrict__ vp, const V4& c, V4& out) {
  // (a non-finite clip coordinate always produces a non-finite screen coordinate below, which the caller's
  // finiteness test sends to the literal path; only the signed-zero cases need to be excluded here)
  if (c.z == 0.0f || c.w == 0.0f) return false;
  const float x = fma32<E>(vp[0], c.x, vp[3] * c.w), y = fma32<E>(vp[5], c.y, vp[7] * c.w);
  if (c.w == 1.0f) { out = V4{x, y, c.z, 1.0f}; return true; }
  const float invW = __fdiv_rn(1.0f, c.w);
  out = V4{x * invW, y * invW, c.z * invW, 1.0f};
  return true;
}

// Shadow passes of several lights share one sweep over the triangles (the chunk's vertices are fetched once per CTA).
struct GeomViews {
  int n;
  uint32_t affine;  // bit v: every object's trans of view v has last row (0,0,0,1) (checked on the host)
  const float* trans[8];  // per view: [n_obj][16] light trans (shadow) — unused for the camera (F.xf)
  float* smap[8];
  uint32_t target[8];
  int r0[8], r1[8];  // rows of the shadow map this view rasterises
  const unsigned char* vis[8];  // per view: [n_chunks] 0 = no triangle of this 256-triangle chunk can touch the view's rows / screen
  int any_vis;                  // some vis[v] is set
  // LIST launches (partial-row views: multi-GPU strips / shadow shards): chunks visible in at least one view, compacted by
  // k_chunk_cull_views (any order), and their number (device memory: no host round trip)
  const unsigned int* list;
  const unsigned int* n_list;
  unsigned int* n_list_hint;    // page-locked host word: the host sizes the NEXT frame's grid from it (a hint, never needed for correctness)
  unsigned int list_from;       // LIST == 2 (tail launch): first list position not covered by the direct launch
  // LIST == 3 (a rank of a multi-GPU group takes its share of the triangles): blocks of PRC_PART_BLOCK chunks dealt round-robin,
  // CTA b handles chunk part_first + (b / PRC_PART_BLOCK) * part_stride + b % PRC_PART_BLOCK
  unsigned int part_first, part_stride, n_chunks;
  unsigned char* dirty;         // LIST == 3: Counters.dirty (row flags of the private targets)
};
#define PRC_PART_BLOCK 16

// ---------------------------------------------------------------------------------------------
// K1: one CTA per 256-triangle chunk, three phases per view, all operands in shared memory.
//   phase 1  the chunk's DISTINCT vertices (k_chunk_dedupe: ~0.6-0.7 per triangle instead of 3) are transformed
//            to screen space once: Mat4.MulV + Apply(Viewport).Pos() depend only on (position, object matrix),
//            so sharing the result between the triangles of a chunk is bit-exact;
//   phase 2  one thread per triangle gathers its three screen vertices, runs the cull / classify / exact-prune
//            sequence; survivors with a small pixel box append one 16-bit entry per candidate pixel to a CTA queue
//            (large boxes -> tile queue, clipping -> clip queue, NaN/Inf/zero w -> geom_generic, as before);
//   phase 3  the queue is consumed by ALL threads, one candidate pixel each: the in-thread pixel loops ran with
//            ~3.5 of 32 lanes active (ncu), the queue runs full warps.
// Several shadow views share the vertex fetch; the camera pass is the same kernel with one view.
// ---------------------------------------------------------------------------------------------
#define PRC_QCAP 2048
template <bool E, bool SHADOW, bool NM = false, bool PEER = false>
__device__ __forceinline__ void small_pixel(const float p1x, const float p1y, const float p1z, const float p2x, const float p2y, const float p2z,
                                            const float p3x, const float p3y, const float p3z, const int x, const int y, const uint32_t seq, const int W,
                                            unsigned long long* keys, float* smap, Counters* cnt, unsigned long long* first = nullptr,
                                            unsigned char* dirty_rows = nullptr) {
  const BarySetup bs = bary_setup<E>(p1x, p1y, p2x, p2y, p3x, p3y);
  const float thr = 2e-7f * fabsf(bs.Sabc);
  const uint32_t sg = __float_as_uint(bs.Sabc);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  const float apx = px - bs.t1x, bpx = px - bs.t2x, apy = py - bs.t1y, bpy = py - bs.t2y;
  if (E) {
    // cheap certain rejection first: the single-rounding fmaf value is within 1 ulp of the reference's
    // double-rounded one, so |S_fmaf| > 4e-7 |Sabc| with the wrong sign implies |S| > 2e-7 |Sabc| below
    const float thr2 = thr + thr;
    const float q0 = cross2z<false>(bs.abx, bs.aby, apx, apy), q1 = cross2z<false>(apx, apy, bs.acx, bs.acy), q2 = cross2z<false>(bs.bcx, bs.bcy, bpx, bpy);
    if ((((__float_as_uint(q0) ^ sg) >> 31) && fabsf(q0) > thr2) || (((__float_as_uint(q1) ^ sg) >> 31) && fabsf(q1) > thr2) ||
        (((__float_as_uint(q2) ^ sg) >> 31) && fabsf(q2) > thr2))
      return;
  }
  const float Sabp = cross2z<E>(bs.abx, bs.aby, apx, apy);
  const float Sapc = cross2z<E>(apx, apy, bs.acx, bs.acy);
  const float Sbcp = cross2z<E>(bs.bcx, bs.bcy, bpx, bpy);
  // certain rejection without dividing: sign(S) != sign(Sabc) and |S| > 2e-7 |Sabc|  =>  RN(S/Sabc) < -1e-7
  if ((((__float_as_uint(Sabp) ^ sg) >> 31) && fabsf(Sabp) > thr) || (((__float_as_uint(Sapc) ^ sg) >> 31) && fabsf(Sapc) > thr) ||
      (((__float_as_uint(Sbcp) ^ sg) >> 31) && fabsf(Sbcp) > thr))
    return;
  const float w1 = __fdiv_rn(Sbcp, bs.Sabc), w2 = __fdiv_rn(Sapc, bs.Sabc), w3 = __fdiv_rn(Sabp, bs.Sabc);
  if (w1 < -PRC_EPS || w2 < -PRC_EPS || w3 < -PRC_EPS) return;
  const float z = w1 * p1z + w2 * p2z + w3 * p3z;
  const size_t idx = (size_t)y * W + x;
  if (NM && !SHADOW) nan_first(first, idx, seq, z);
  if (PEER && (!SHADOW || z > 0.0f || (NM && isnan(z)))) dirty_rows[((size_t)y * dirty_nseg(W) + (x / PRC_DIRTY_SEG)) * PRC_DIRTY_STRIDE] = 1;  // peer groups: this row of the private buffer holds something
  if (isnan(z)) { atomicAdd(SHADOW ? &cnt->n_nan_shadow : &cnt->n_nan, 1ULL); return; }
  // fire-and-forget reductions (RED.MAX): no pre-test load, so nothing waits on memory
  if (SHADOW) {
    if (z > 0.0f) atomicMax((int*)&smap[idx], __float_as_int(z));
  } else {
    unsigned long long key = ((unsigned long long)depth_key(z) << 32) | (unsigned long long)(0xFFFFFFFFu - seq);
    atomicMax(&keys[idx], key);
  }
}

struct GeomSmem {
  float x[2][768], y[2][768], z[2][768];  // screen-space distinct vertices of the chunk, double-buffered over views (NaN x = "take the literal path")
  uint32_t idx[PRC_GEOM_THREADS];  // per triangle with queued candidates: its three local vertex indices
  uint32_t box[PRC_GEOM_THREADS];  // x0 | y0 << 14 | (box width - 1) << 28 of its pixel box
  unsigned short q[PRC_QCAP];      // candidate = triangle slot | pixel number within the box << 8
  unsigned int qn[2], qv[2];       // reserved / valid entries, alternating between consecutive views
  unsigned int it;                 // LIST launches: the CTA's position in the chunk list (kept out of the 32 registers)
};
__constant__ unsigned short c_recip256[17] = {0, 256, 128, 86, 64, 52, 43, 37, 32, 29, 26, 24, 22, 20, 19, 18, 16};  // ceil(256 / bw): (j * r) >> 8 == j / bw for j < 16

// phase 1: the chunk's distinct vertices -> screen space of one view
template <bool E, bool SHADOW>
__device__ __forceinline__ void geom_vertices(const DevScene& S, const DevFrame& F, float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sz,
                                              const uint32_t voff, const uint32_t nv, const float* __restrict__ trans_base, const int trans_stride, const bool affine) {
  for (uint32_t k = threadIdx.x; k < nv; k += PRC_GEOM_THREADS) {
    const float4 q = __ldg(S.cverts + voff + k);
    const float* trans = trans_base + (size_t)__float_as_uint(q.w) * trans_stride;
    V4 c, p{0.0f, 0.0f, 0.0f, 1.0f};
    // affine: last row of trans is exactly (0,0,0,1) (orthographic light camera x affine model): w = 0*x+0*y+0*z+1*1 = 1
    // for finite x,y,z; non-finite coordinates make x/y/z non-finite too and are caught by the magnitude test below
    if (affine) c = mulv3(trans, q.x, q.y, q.z);
    else c = mulv(trans, V4{q.x, q.y, q.z, 1.0f});
    // literal path (geom_generic, per triangle) for: a non-standard viewport matrix, zero z or w, NaN / Inf, or
    // coordinates so large that differences of two of them could overflow
    if (!(F.vp_std && viewport_pos_std<E>(F.viewport, c, p)) || !(fabsf(p.x) + fabsf(p.y) + fabsf(p.z) < 3e29f)) p.x = __int_as_float(0x7fc00000);
    sx[k] = p.x; sy[k] = p.y; sz[k] = p.z;
  }
}

// phase 2 for one triangle
template <bool E, bool SHADOW, bool NM = false, bool PEER = false>
__device__ __forceinline__ void geom_classify(const DevScene& S, const DevFrame& F, GeomSmem& sm, const int buf, const int qsel, const uint32_t li, const unsigned int tri,
                                              const float* __restrict__ trans_base, const int trans_stride,
                                              unsigned long long* keys, float* smap, LargeRec* large, unsigned int large_cap, unsigned int* clipq,
                                              unsigned int clip_cap, Counters* cnt, const DevFrame* Fg, uint32_t target, const int vr0, const int vr1,
                                              unsigned char* dirty_rows = nullptr) {
  const uint32_t i0 = li & 1023u, i1 = (li >> 10) & 1023u, i2 = (li >> 20) & 1023u;
  const float* sx = sm.x[buf]; const float* sy = sm.y[buf]; const float* sz = sm.z[buf];
  const float p1x = sx[i0], p2x = sx[i1], p3x = sx[i2];
  const float fin = p1x + p2x + p3x;  // NaN iff a vertex was flagged in phase 1 (finite coordinates are < 3e29: no overflow)
  if (fin != fin) {
    const uint32_t obj = __ldg(S.meta + tri) & 0x00FFFFFFu;
    geom_generic<E, SHADOW, NM>(Fg, trans_base + (size_t)obj * trans_stride, S.pos, tri, keys, smap, large, large_cap, target, clipq, clip_cap, cnt, vr0, vr1);
    return;
  }
  const float p1y = sy[i0], p2y = sy[i1], p3y = sy[i2];
  // cullBackFace (render/cull.go:26-28): FMA(e1.x, e2.y, -(e1.y*e2.x)) < 0, the same expression as Barycoord's Sabc.
  // Only its SIGN is needed here, and the single-rounding FMA has the sign of the exactly computed value just like the
  // reference's double-rounded one (both round the same exact number, neither can round a non-zero value of this
  // magnitude to zero); a zero or subnormal result takes the exact route. The pruning test below tolerates the 1-ulp
  // difference (its threshold has orders of magnitude of slack, prc_prune.h); phase 3 recomputes Sabc exactly.
  float Sabc = cross2z<false>(p2x - p1x, p2y - p1y, p3x - p1x, p3y - p1y);
  if (E && !(fabsf(Sabc) >= 1.17549435e-38f)) Sabc = cross2z<true>(p2x - p1x, p2y - p1y, p3x - p1x, p3y - p1y);
  if (Sabc < 0.0f) return;
  const float p1z = sz[i0], p2z = sz[i1], p3z = sz[i2];
  // finite coordinates: Go's NaN-propagating Min/Max reduce to plain min/max (the sign of a zero is irrelevant below)
  const float mnx = fminf(fminf(p1x, p2x), p3x), mxx = fmaxf(fmaxf(p1x, p2x), p3x);
  const float mny = fminf(fminf(p1y, p2y), p3y), mxy = fmaxf(fmaxf(p1y, p2y), p3y);
  const float mnz = fminf(fminf(p1z, p2z), p3z), mxz = fmaxf(fmaxf(p1z, p2z), p3z);
  const float Wf = F.cullW, Hf = F.cullH;
  // AABB.Intersect (box.go:32-41): max(lo) <= min(hi) per axis; the Z test compares against Max.Y = H (the Z quirk)
  if (!(mxx >= 0.0f && mnx <= Wf && mxy >= 0.0f && mny <= Hf && mxz >= -1.0f && mnz <= Hf)) return;
  if (!SHADOW) {
    // AABB.Contains (box.go:61-75) is monotone per coordinate, so testing the extremes tests all three vertices
    const bool in = less_eq(0.0f, mnx) && less_eq(0.0f, mny) && less_eq(-1.0f, mnz) && less_eq(mxx, Wf) && less_eq(mxy, Hf) && less_eq(mxz, 1.0f);
    if (!in) {
      unsigned int slot = warp_push(&cnt->n_clip);
      if (slot < clip_cap) clipq[slot] = tri;
      else atomicExch(&cnt->large_overflow, 1u);
      return;
    }
  }
  const int r0 = SHADOW ? vr0 : F.rr0, r1 = SHADOW ? vr1 : F.rr1;
  const uint32_t seq = tri * 8u;
  if (!SHADOW && r0 > 0) raster_pixel00<E, NM>(V4{p1x, p1y, p1z, 1.0f}, V4{p2x, p2y, p2z, 1.0f}, V4{p3x, p3y, p3z, 1.0f}, seq, keys, cnt, NM ? keys + (size_t)F.W * F.H : nullptr);
  int x0, x1, y0, y1;
  if (prune_ok(mnx, mny, mxx, mxy, Sabc)) {
    // exact-safe shrink of the AABB+-1 loop (prc_prune.h). The pruned box [ceil(min-.5-M), floor(max-.5+M)] always lies
    // inside the reference's int(Round(min)-1) .. int(Round(max)+1), so the latter need not be computed here.
    x0 = max(0, prune_first(mnx)); x1 = min(F.W - 1, prune_last(mxx));
    y0 = max(r0, prune_first(mny)); y1 = min(r1 - 1, prune_last(mxy));
  } else {
    // pixel box int(Round(min)-1) .. int(Round(max)+1) clamped to the buffer (raster.go:473-485); clamping in float first
    x0 = (int)fmaxf(roundf(mnx) - 1.0f, 0.0f); x1 = (int)fminf(roundf(mxx) + 1.0f, (float)(F.W - 1));
    y0 = (int)fmaxf(roundf(mny) - 1.0f, (float)r0); y1 = (int)fminf(roundf(mxy) + 1.0f, (float)(r1 - 1));
  }
  if (x0 > x1 || y0 > y1) return;
  const int bw = x1 - x0 + 1, area = bw * (y1 - y0 + 1);
  if (area > PRC_SMALL_MAX_PIXELS) {
    unsigned int slot = warp_push(&cnt->n_large);
    if (slot >= large_cap) { atomicExch(&cnt->large_overflow, 1u); return; }
    LargeRec lr;
    lr.x1 = p1x; lr.y1 = p1y; lr.z1 = p1z; lr.x2 = p2x; lr.y2 = p2y; lr.z2 = p2z; lr.x3 = p3x; lr.y3 = p3y; lr.z3 = p3z;
    lr.seq = seq; lr.bx0 = (short)x0; lr.by0 = (short)y0; lr.bx1 = (short)x1; lr.by1 = (short)y1; lr.target = target;
    large[slot] = lr;
    return;
  }
  // candidate pixels of the box (render/raster.go:481-499 / render/shadow.go:191-215) -> CTA queue
  // (one shared-memory atomic per lane: a warp-aggregated reservation — five ballots for the prefix — measured 8 % slower, round 2)
  const unsigned int base = atomicAdd(&sm.qn[qsel], (unsigned int)area);
  if (base + area <= PRC_QCAP) {
    atomicMax(&sm.qv[qsel], base + area);
    sm.idx[threadIdx.x] = li;
    sm.box[threadIdx.x] = (uint32_t)x0 | ((uint32_t)y0 << 14) | ((uint32_t)(bw - 1) << 28);
    unsigned short e = (unsigned short)threadIdx.x;
    unsigned short* q = sm.q + base;
    q[0] = e;  // most survivors cover one or two candidate pixels
    if (area > 1) {
      q[1] = (unsigned short)(e + 256);
      e += 512;
      for (int j = 2; j < area; j++, e += 256) q[j] = e;
    }
  } else {
    // queue full (many multi-pixel triangles in one chunk): this triangle's pixels in-thread
    for (int y = y0; y <= y1; y++)
      for (int x = x0; x <= x1; x++)
        small_pixel<E, SHADOW, NM, PEER>(p1x, p1y, p1z, p2x, p2y, p2z, p3x, p3y, p3z, x, y, seq, F.W, keys, smap, cnt, NM ? keys + (size_t)F.W * F.H : nullptr, dirty_rows);
  }
}

// __grid_constant__: the per-view arrays are indexed with a run-time view number; without it the whole parameter
// struct is copied to local memory by every thread (ncu: 11 % of the kernel's instructions, STL at entry).
template <bool E, bool SHADOW, bool NM = false, int LIST = 0>
__global__ void __launch_bounds__(PRC_GEOM_THREADS, PRC_GEOM_MIN_BLOCKS) k_geom_raster(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F,
                                                                     const __grid_constant__ GeomViews V,
                                                                     unsigned long long* keys, LargeRec* large, unsigned int large_cap,
                                                                     unsigned int* clipq, unsigned int clip_cap, Counters* cnt, const DevFrame* Fg) {
  __shared__ GeomSmem sm;
  // LIST launches take their chunks from the compacted list of chunks that can touch some view (k_chunk_cull_views) instead of
  // launching one CTA per chunk that mostly reads its visibility byte and exits — on C3 a strip or shadow shard of an 8-GPU
  // frame touches a few thousand of the 39 063 chunks, and the empty CTAs alone cost 0.04 ms per pass (measured, round 2).
  //   LIST == 1  CTA b handles list[b]; the host sizes the grid from the previous frame's list length (n_list_hint) plus a margin
  //   LIST == 2  a small tail launch (stride loop) for list positions >= list_from, i.e. when the list outgrew that grid —
  //              normally it finds nothing. (A stride loop in the main launch cost 25 % at 32 registers: measured, round 2.)
  //   LIST == 3  no list: this rank's share of the chunks (see GeomViews.part_first)
  if (LIST == 1 && blockIdx.x == 0 && threadIdx.x == 0 && V.n_list_hint) *V.n_list_hint = *V.n_list;
  if (LIST == 1 && blockIdx.x >= *V.n_list) return;
  if (LIST == 2 && threadIdx.x == 0) sm.it = V.list_from + blockIdx.x;
  for (;;) {
  if (LIST == 2) {
    __syncthreads();  // sm.it is set; the previous chunk's shared vertices and queue are no longer read
    if (sm.it >= *V.n_list) return;
  }
  const unsigned int chunk = LIST == 1 ? V.list[blockIdx.x] : LIST == 2 ? V.list[sm.it]
                             : LIST == 3 ? V.part_first + (blockIdx.x / PRC_PART_BLOCK) * V.part_stride + (blockIdx.x % PRC_PART_BLOCK) : blockIdx.x;
  if (LIST == 3 && chunk >= V.n_chunks) return;
  const unsigned long long tri64 = (unsigned long long)chunk * PRC_GEOM_THREADS + threadIdx.x;
  const unsigned int tri = (unsigned int)tri64;
  const uint32_t li = tri64 < S.n_tris ? __ldg(S.lidx + tri) : 0xFFFFFFFFu;
  const uint32_t voff = __ldg(S.cvoff + chunk), nv = __ldg(S.cvoff + chunk + 1) - voff;
  const int n_views = SHADOW ? V.n : 1;
  const int trans_stride = SHADOW ? 16 : (int)(sizeof(prc_object_xf) / sizeof(float));
  // views this chunk can touch (k_chunk_cull; uniform over the CTA)
  uint32_t todo = (1u << n_views) - 1u;
  if (V.any_vis) {
    todo = 0;
    for (int v = 0; v < n_views; v++)
      if (V.vis[v] == nullptr || V.vis[v][chunk] != 0) todo |= 1u << v;
    if (!todo) {
      if (LIST != 2) return;
      __syncthreads();  // every thread has read sm.it
      if (threadIdx.x == 0) sm.it += gridDim.x;
      continue;
    }
  }
  if (threadIdx.x == 0) { sm.qn[0] = sm.qn[1] = 0; sm.qv[0] = sm.qv[1] = 0; }
  int v = __ffs(todo) - 1, buf = 0;
  todo &= todo - 1;
  geom_vertices<E, SHADOW>(S, F, sm.x[0], sm.y[0], sm.z[0], voff, nv, SHADOW ? V.trans[v] : reinterpret_cast<const float*>(F.xf), trans_stride,
                           SHADOW && ((V.affine >> v) & 1));
  __syncthreads();
#pragma unroll 1
  for (;;) {
    const float* trans_base = SHADOW ? V.trans[v] : reinterpret_cast<const float*>(F.xf);
    float* smap = SHADOW ? V.smap[v] : nullptr;
    // peer groups: the row flags of this view's private target (see Counters.dirty); V.dirty is uniform, no dependent load per fragment
    unsigned char* dirty_rows = LIST == 3 ? V.dirty + (size_t)(SHADOW ? V.target[v] : 0u) * F.H * dirty_nseg(F.W) * PRC_DIRTY_STRIDE : nullptr;
    // ---- phase 2: one thread per triangle
    if (li != 0xFFFFFFFFu)
      geom_classify<E, SHADOW, NM, LIST == 3>(S, F, sm, buf, buf, li, tri, trans_base, trans_stride, keys, smap, large, large_cap, clipq, clip_cap, cnt, Fg,
                                              SHADOW ? V.target[v] : 0u, SHADOW ? V.r0[v] : F.rr0, SHADOW ? V.r1[v] : F.rr1, dirty_rows);
    __syncthreads();
    // ---- phase 3 of this view (one thread per candidate pixel) overlapped with phase 1 of the next view (other buffer)
    const unsigned int nq = sm.qv[buf];
    const int vn = todo ? __ffs(todo) - 1 : -1;
    todo &= todo - 1;
    if (threadIdx.x == 0) { sm.qn[buf ^ 1] = 0; sm.qv[buf ^ 1] = 0; }  // last read before the barrier above
    if (vn >= 0)
      geom_vertices<E, SHADOW>(S, F, sm.x[buf ^ 1], sm.y[buf ^ 1], sm.z[buf ^ 1], voff, nv, V.trans[vn], trans_stride, (V.affine >> vn) & 1);
    {
      const float* sx = sm.x[buf]; const float* sy = sm.y[buf]; const float* sz = sm.z[buf];
      // dealt from the LAST warp down: the first warps are the ones busy with the next view's vertices
      for (unsigned int c = PRC_GEOM_THREADS - 1 - threadIdx.x; c < nq; c += PRC_GEOM_THREADS) {
        const uint32_t e = sm.q[c], t = e & 255u, j = e >> 8;
        const uint32_t ti = sm.idx[t], box = sm.box[t];
        const uint32_t i0 = ti & 1023u, i1 = (ti >> 10) & 1023u, i2 = (ti >> 20) & 1023u;
        const uint32_t bw = (box >> 28) + 1u, dy = (j * c_recip256[bw]) >> 8, dx = j - dy * bw;
        small_pixel<E, SHADOW, NM, LIST == 3>(sx[i0], sy[i0], sz[i0], sx[i1], sy[i1], sz[i1], sx[i2], sy[i2], sz[i2],
                                              (int)((box & 0x3FFFu) + dx), (int)(((box >> 14) & 0x3FFFu) + dy),
                                              (chunk * PRC_GEOM_THREADS + t) * 8u, F.W, keys, smap, cnt, NM ? keys + (size_t)F.W * F.H : nullptr, dirty_rows);
      }
    }
    if (vn < 0) break;
    __syncthreads();
    v = vn; buf ^= 1;
  }
    if (LIST != 2) break;
    __syncthreads();  // every thread is done with this chunk (and has read sm.it)
    if (threadIdx.x == 0) sm.it += gridDim.x;
  }
}

// K2: triangles straddling the viewport: clip, fan, emit (raster.go:438-443)
template <bool E, bool NM = false>
__global__ void k_clip_raster(DevScene S, DevFrame F, const unsigned int* __restrict__ clipq, unsigned long long* keys, LargeRec* large,
                              unsigned int large_cap, Counters* cnt) {
  const unsigned int n = cnt->n_clip;  // final: written by the preceding kernel on the same stream
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
  const unsigned int tri = clipq[i];
  const uint32_t obj = S.meta[tri] & 0x00FFFFFFu;
  float p[9];
#pragma unroll
  for (int k = 0; k < 9; k++) p[k] = S.pos[(size_t)tri * 9 + k];
  ScreenTri st;
  tri_setup<E>(F.xf[obj].trans, F.viewport, F.pm_viewport, F.cullW, F.cullH, p, true, st);
  V4 poly[12];
  int nc = clip_polygon<E>(st, F.cullW, F.cullH, poly);
  if (nc < 3) continue;
  float b0[3];
  clip_bary<E>(st, poly[0], b0);
  V4 q0 = clip_pos(st, b0);
  for (int k = 2; k < nc && k <= 8; k++) {
    float b1[3], b2[3];
    clip_bary<E>(st, poly[k - 1], b1);
    clip_bary<E>(st, poly[k], b2);
    V4 q1 = clip_pos(st, b1), q2 = clip_pos(st, b2);
    emit_tri<E, false, NM>(q0, q1, q2, tri * 8u + (uint32_t)(k - 1), F, F.rr0, F.rr1, keys, nullptr, large, large_cap, 0u, cnt, NM ? keys + (size_t)F.W * F.H : nullptr);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// tile path for large triangles, ONCE per frame for all raster passes: every queued record carries its
// target (camera keys or a shadow map); bins are indexed by virtual tile = target * n_tiles + tile.
//   k_bin_count -> k_scan_sums -> k_scan_apply (+ list of non-empty virtual tiles) -> k_bin_fill -> k_tile_raster
// ---------------------------------------------------------------------------------------------
#define PRC_MEDIUM_MAX_PIXELS 4096  // pixel-box area up to which a queued record is rasterised by one warp (k_medium_raster); larger: binned tile path
__device__ __forceinline__ bool rec_is_huge(const LargeRec& r) { return ((int)r.bx1 - r.bx0 + 1) * ((int)r.by1 - r.by0 + 1) > PRC_MEDIUM_MAX_PIXELS; }

__global__ void k_bin_count(const LargeRec* __restrict__ large, const Counters* cnt, unsigned int cap, int tiles_x, int n_tiles, unsigned int* tile_count) {
  const unsigned int n = min(cnt->n_large, cap);
  const unsigned int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < n; warp += nwarps) {
    const LargeRec r = large[warp];
    if (!rec_is_huge(r)) continue;  // rasterised by k_medium_raster
    unsigned int* tc = tile_count + (size_t)r.target * n_tiles;
    const int tx0 = r.bx0 / PRC_TILE, tx1 = r.bx1 / PRC_TILE, ty0 = r.by0 / PRC_TILE, ty1 = r.by1 / PRC_TILE;
    const int nx = tx1 - tx0 + 1, nt = nx * (ty1 - ty0 + 1);
    for (int t = lane; t < nt; t += 32) atomicAdd(&tc[(ty0 + t / nx) * tiles_x + tx0 + t % nx], 1u);
  }
}
// exclusive scan over the virtual tiles in chunks of 4096 (4 per thread, 128-bit loads; arrays are padded to a
// multiple of 4096 and zero-filled): pass 1 = chunk sums, pass 2 = scan within the chunk + chunk offset.
__device__ __forceinline__ unsigned int block_scan_1024(unsigned int local, unsigned int* warp_sum, unsigned int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned int w = warp_sum[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_sum[lane] = wi - w;
    if (lane == 31) warp_sum[32] = wi;
  }
  __syncthreads();
  total = warp_sum[32];
  return warp_sum[warp] + incl - local;  // exclusive prefix of this thread within the block
}
__global__ void __launch_bounds__(1024) k_scan_sums(const unsigned int* __restrict__ in, unsigned int* chunk_sum, const Counters* cnt) {
  __shared__ unsigned int warp_sum[33];
  if (cnt->n_huge == 0) return;
  const uint4 v = *reinterpret_cast<const uint4*>(in + (size_t)blockIdx.x * 4096 + threadIdx.x * 4);
  unsigned int total;
  block_scan_1024(v.x + v.y + v.z + v.w, warp_sum, total);
  if (threadIdx.x == 0) chunk_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan_apply(const unsigned int* __restrict__ in, unsigned int* out, unsigned int* cursor,
                                                      const unsigned int* __restrict__ chunk_sum, Counters* cnt, unsigned int bins_cap, unsigned int large_cap,
                                                      unsigned int* active, unsigned int* n_active) {
  __shared__ unsigned int warp_sum[33];
  __shared__ unsigned int base_s, grand_s;
  if (cnt->n_huge == 0) {  // nothing for the tile path this frame: n_active stays 0 (host memset)
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->n_bin_total = 0;
    return;
  }
  // offset of this chunk = sum of the previous chunk sums (gridDim.x <= 1024 chunks)
  {
    unsigned int c = threadIdx.x < gridDim.x ? chunk_sum[threadIdx.x] : 0u;
    unsigned int tot;
    const unsigned int ex = block_scan_1024(c, warp_sum, tot);
    if (threadIdx.x == blockIdx.x) base_s = ex;
    if (threadIdx.x == 0) grand_s = tot;
    __syncthreads();
  }
  const size_t i = (size_t)blockIdx.x * 4096 + threadIdx.x * 4;
  const uint4 v = *reinterpret_cast<const uint4*>(in + i);
  unsigned int tot;
  const unsigned int ex = base_s + block_scan_1024(v.x + v.y + v.z + v.w, warp_sum, tot);
  const uint4 o4 = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
  *reinterpret_cast<uint4*>(out + i) = o4;
  *reinterpret_cast<uint4*>(cursor + i) = o4;
  const unsigned int c[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (c[k]) active[atomicAdd(n_active, 1u)] = (unsigned int)(i + k);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned int total = grand_s;
    cnt->n_bin_total = total;
    if (total > cnt->max_bins) cnt->max_bins = total;
    if (total > bins_cap) cnt->large_overflow = 1u;
    cnt->stat_bins += total;  // (stat_large / stat_clip and the queue-overflow test: k_medium_raster)
  }
}
__global__ void k_bin_fill(const LargeRec* __restrict__ large, const Counters* cnt, unsigned int cap, int tiles_x, int n_tiles, unsigned int* cursor,
                           unsigned int* bins, unsigned int bins_cap) {
  const unsigned int n = min(cnt->n_large, cap);
  if (cnt->n_bin_total > bins_cap) return;  // frame flagged for a re-render with a larger bin array
  const unsigned int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; warp < n; warp += nwarps) {
    const LargeRec r = large[warp];
    if (!rec_is_huge(r)) continue;
    unsigned int* cur = cursor + (size_t)r.target * n_tiles;
    const int tx0 = r.bx0 / PRC_TILE, tx1 = r.bx1 / PRC_TILE, ty0 = r.by0 / PRC_TILE, ty1 = r.by1 / PRC_TILE;
    const int nx = tx1 - tx0 + 1, nt = nx * (ty1 - ty0 + 1);
    for (int t = lane; t < nt; t += 32) {
      unsigned int slot = atomicAdd(&cur[(ty0 + t / nx) * tiles_x + tx0 + t % nx], 1u);
      if (slot < bins_cap) bins[slot] = warp;
    }
  }
}

struct TileTargets { float* smap[33]; };  // [1 + k] = shadow map of the k-th casting light

// Queued records (pixel box > 16 px) up to PRC_MEDIUM_MAX_PIXELS: one warp per record walks the clamped pixel box, 32 pixels at
// a time, with the tile raster's arithmetic (bary_setup / bary_eval) and reduces straight into the target. On C3 that is every
// queued record (47 520 per frame, 1.45 tiles each): the binned tile path — five dependent launches whose fixed cost
// (0.034 ms of a 1.14 ms frame, 0.055 ms of an 8-GPU rank's 0.27 ms) exceeded their work — then never runs. Records above the limit
// (a ground quad filling the screen) are counted in n_huge and left to the bins; if those are switched off (`bins_on` = 0:
// no huge record seen so far) the frame is flagged and re-rendered with them, like a queue overflow.
template <bool E, bool NM = false>
__global__ void __launch_bounds__(256) k_medium_raster(const LargeRec* __restrict__ large, Counters* cnt, unsigned int cap, int W, int H,
                                                       unsigned long long* keys, const TileTargets* __restrict__ targets, int bins_on) {
  const unsigned int n_all = cnt->n_large, n = min(n_all, cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cnt->stat_large += n_all; cnt->stat_clip += cnt->n_clip;
    if (n_all > cap) cnt->large_overflow = 1u;
  }
  const unsigned int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long nan_cam = 0, nan_sh = 0;
  for (unsigned int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += nwarps) {
    const LargeRec r = large[w];
    if (rec_is_huge(r)) {
      if (lane == 0) {
        atomicAdd(&cnt->n_huge, 1u);
        if (!bins_on) cnt->need_bins = 1u;
      }
      continue;
    }
    const BarySetup bs = bary_setup<E>(r.x1, r.y1, r.x2, r.y2, r.x3, r.y3);
    const int bw = r.bx1 - r.bx0 + 1, area = bw * (r.by1 - r.by0 + 1);
    const bool shadow = r.target != 0;
    float* smap = shadow ? targets->smap[r.target] : nullptr;
    for (int p = (int)lane; p < area; p += 32) {
      const int dy = p / bw, x = r.bx0 + (p - dy * bw), y = r.by0 + dy;
      float w1, w2, w3;
      bary_eval<E>(bs, (float)x + 0.5f, (float)y + 0.5f, w1, w2, w3);
      if (w1 < -PRC_EPS || w2 < -PRC_EPS || w3 < -PRC_EPS) continue;
      const float z = w1 * r.z1 + w2 * r.z2 + w3 * r.z3;
      const size_t idx = (size_t)y * W + x;
      if (NM && !shadow) nan_first(keys + (size_t)W * H, idx, r.seq, z);
      if (isnan(z)) { if (shadow) nan_sh++; else { nan_cam++; if (NM) mark_dirty(cnt, 0, x, y, W); } continue; }
      mark_dirty(cnt, r.target, x, y, W);
      if (shadow) {
        if (z > 0.0f) atomicMax((int*)&smap[idx], __float_as_int(z));
      } else {
        atomicMax(&keys[idx], ((unsigned long long)depth_key(z) << 32) | (unsigned long long)(0xFFFFFFFFu - r.seq));
      }
    }
  }
  if (nan_cam) atomicAdd(&cnt->n_nan, nan_cam);
  if (nan_sh) atomicAdd(&cnt->n_nan_shadow, nan_sh);
}

struct TileRec {
  BarySetup bs;
  float z1, z2, z3;
  uint32_t seq;
  short bx0, by0, bx1, by1;
};
template <bool E, bool NM = false>
__global__ void __launch_bounds__(PRC_TILE * PRC_TILE) k_tile_raster(const LargeRec* __restrict__ large, const unsigned int* __restrict__ tile_start,
                                                                      const unsigned int* __restrict__ bins, int tiles_x, int n_tiles, int W, int H,
                                                                      unsigned long long* keys, const TileTargets* __restrict__ targets, Counters* cnt,
                                                                      const unsigned int* __restrict__ active, const unsigned int* __restrict__ n_active) {
  __shared__ TileRec recs[128];
  if (cnt->large_overflow) return;
  const unsigned int na = *n_active;
  for (unsigned int ai = blockIdx.x; ai < na; ai += gridDim.x) {
    const int vtile = (int)active[ai];
    const int target = vtile / n_tiles, tile = vtile - target * n_tiles;
    const unsigned int b = tile_start[vtile], e = tile_start[vtile + 1];
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const int x = tx * PRC_TILE + (threadIdx.x & (PRC_TILE - 1)), y = ty * PRC_TILE + (threadIdx.x / PRC_TILE);
    const float px = (float)x + 0.5f, py = (float)y + 0.5f;
    const bool live = x < W && y < H;
    const bool shadow = target != 0;
    unsigned long long best = 0, firstv = ~0ull;
    float bestz = 0.0f;
    unsigned long long nan_local = 0;
    for (unsigned int base = b; base < e; base += 128) {
      const int nb = min(128u, e - base);
      __syncthreads();
      if ((int)threadIdx.x < nb) {
        const LargeRec r = large[bins[base + threadIdx.x]];
        TileRec t;
        t.bs = bary_setup<E>(r.x1, r.y1, r.x2, r.y2, r.x3, r.y3);
        t.z1 = r.z1; t.z2 = r.z2; t.z3 = r.z3; t.seq = r.seq;
        t.bx0 = r.bx0; t.by0 = r.by0; t.bx1 = r.bx1; t.by1 = r.by1;
        recs[threadIdx.x] = t;
      }
      __syncthreads();
      if (!live) continue;
      for (int j = 0; j < nb; j++) {
        const TileRec& t = recs[j];
        if (x < t.bx0 || x > t.bx1 || y < t.by0 || y > t.by1) continue;  // the record's box is already clamped to the pass's rows
        float w1, w2, w3;
        bary_eval<E>(t.bs, px, py, w1, w2, w3);
        if (w1 < -PRC_EPS || w2 < -PRC_EPS || w3 < -PRC_EPS) continue;
        float z = w1 * t.z1 + w2 * t.z2 + w3 * t.z3;
        if (NM && !shadow) firstv = min(firstv, ((unsigned long long)t.seq << 1) | (isnan(z) ? 1ull : 0ull));
        if (isnan(z)) { nan_local++; continue; }
        if (shadow) {
          if (z > bestz) bestz = z;
        } else {
          unsigned long long key = ((unsigned long long)depth_key(z) << 32) | (unsigned long long)(0xFFFFFFFFu - t.seq);
          if (key > best) best = key;
        }
      }
    }
    if (nan_local) atomicAdd(shadow ? &cnt->n_nan_shadow : &cnt->n_nan, nan_local);
    if (!live) continue;
    const size_t idx = (size_t)y * W + x;
    if ((shadow ? bestz > 0.0f : best != 0) || (NM && firstv != ~0ull)) mark_dirty(cnt, (uint32_t)target, x, y, W);
    if (NM && firstv != ~0ull) atomicMin(&keys[(size_t)W * H + idx], firstv);
    if (shadow) {
      if (bestz > 0.0f) atomicMax((int*)&targets->smap[target][idx], __float_as_int(bestz));
    } else {
      if (best) atomicMax(&keys[idx], best);
    }
  }
}

// NaN mode: a pixel whose first fragment in draw order has a NaN depth shows that fragment (see nan_first). The key's depth
// half is not read by the resolve (it recomputes the fragment from the sequence number), it only has to be non-zero.
__global__ void k_nan_fix(unsigned long long* keys, const unsigned long long* __restrict__ first, size_t i0, size_t i1, int with00) {
  const size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (with00 && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long v = first[0];
    if (v != ~0ull && (v & 1ull)) keys[0] = (1ull << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(v >> 1));
  }
  if (i >= i1) return;
  const unsigned long long v = first[i];
  if (v != ~0ull && (v & 1ull)) keys[i] = (1ull << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)(v >> 1));
}

// ---------------------------------------------------------------------------------------------
// K3: resolve — visibility key -> fragment attributes (drawClipped, raster.go:462-572)
// G-buffer layout: 4 x float4 per pixel
//   ga = (depth, u, v, du)  gb = (dv, nx, ny, nz)  gc = (fx, fy, fz, wx)  gd = (wy, wz, col bits, mat bits)
// ---------------------------------------------------------------------------------------------
struct Frag {
  bool ok;
  int X, Y;
  float depth, u, v, du, dv;
  V4 nor, facenor, wpos;
  uint32_t col;
  int32_t mat;
};

struct VtxAttr { V4 pos, nor; float u, v; uint32_t col; };

// E: arithmetic of everything that decides position / depth; EA: arithmetic of the shading-only attributes
// (normals, world position, face normal, du/dv). PRC_FMA=mixed runs <true, false>.
// (x, y, z, 1).Apply(ViewportInv).Apply(ProjInv).Apply(ViewInv) for the matrix shapes flagged by DevFrame.unproj_std
__device__ __forceinline__ V4 unproject_std(const DevFrame& F, const V4& p) {
  const float *a = F.viewport_inv, *q = F.proj_inv, *r = F.view_inv;
  const float x1 = __fmaf_rn(a[0], p.x, a[3]), y1 = __fmaf_rn(a[5], p.y, a[7]), z1 = p.z;
  const float x2 = q[0] * x1, y2 = q[5] * y1, z2 = q[11], w2 = __fmaf_rn(q[14], z1, q[15]);
  V4 o;
  o.x = __fmaf_rn(r[0], x2, __fmaf_rn(r[1], y2, __fmaf_rn(r[2], z2, r[3] * w2)));
  o.y = __fmaf_rn(r[4], x2, __fmaf_rn(r[5], y2, __fmaf_rn(r[6], z2, r[7] * w2)));
  o.z = __fmaf_rn(r[8], x2, __fmaf_rn(r[9], y2, __fmaf_rn(r[10], z2, r[11] * w2)));
  o.w = w2;
  return o;
}

// Out-of-line literal sequences of the resolve stage (rare inputs; the hot kernels are instruction-cache bound, see query_bilinear_wide).
// The kernels take their parameter structs as __grid_constant__, so passing their addresses does not copy them to local memory.
template <bool E>
__device__ __noinline__ ScreenTri tri_setup_literal(const DevFrame* F, const float* __restrict__ trans, float p0, float p1, float p2, float p3, float p4, float p5,
                                                    float p6, float p7, float p8) {
  const float p[9] = {p0, p1, p2, p3, p4, p5, p6, p7, p8};
  ScreenTri st;
  tri_setup<E>(trans, F->viewport, F->pm_viewport, F->cullW, F->cullH, p, true, st);
  return st;
}
template <bool EA>
__device__ __noinline__ V4 unproject_literal(const DevFrame* F, V4 p) {  // raster.go:467-469
  return apply4m<EA>(apply4m<EA>(apply4m<EA>(p, F->viewport_inv, F->pm_viewport_inv), F->proj_inv, F->pm_proj_inv), F->view_inv, F->pm_view_inv);
}
// The screen triangle of the resolve stage. With the standard viewport matrix and finite, non-zero clip z / w the transform is the
// one phase 1 of k_geom_raster runs (Mat4.MulV, viewport_pos_std — bit for bit the reference's Apply(Viewport).Pos(), proof there);
// anything else takes the literal tri_setup, as the triangle did in the geometry pass (geom_generic).
template <bool E>
__device__ __forceinline__ void resolve_screen_tri(const DevFrame& F, const float* __restrict__ trans, const float* p, ScreenTri& st) {
  if (F.vp_std) {
    const V4 a = mulv(trans, V4{p[0], p[1], p[2], 1.0f}), b = mulv(trans, V4{p[3], p[4], p[5], 1.0f}), c = mulv(trans, V4{p[6], p[7], p[8], 1.0f});
    if (viewport_pos_std<E>(F.viewport, a, st.p1) && viewport_pos_std<E>(F.viewport, b, st.p2) && viewport_pos_std<E>(F.viewport, c, st.p3) &&
        fabsf(st.p1.x) + fabsf(st.p1.y) + fabsf(st.p1.z) + fabsf(st.p2.x) + fabsf(st.p2.y) + fabsf(st.p2.z) + fabsf(st.p3.x) + fabsf(st.p3.y) + fabsf(st.p3.z) < 3e29f) {
      st.cw1 = a.w; st.cw2 = b.w; st.cw3 = c.w;
      return;
    }
  }
  st = tri_setup_literal<E>(&F, trans, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
}

// FAN = false: the caller guarantees seq & 7 == 0 (an unclipped triangle) and the clipper is compiled out.
template <bool E, bool EA, bool FAN>
__device__ __forceinline__ void resolve_fragment_impl(const DevScene& S, const DevFrame& F, uint32_t seq, int x, int y, Frag& f) {
  const uint32_t tri = seq >> 3, sub = FAN ? (seq & 7u) : 0u;
  const uint32_t obj = S.meta[tri] & 0x00FFFFFFu;
  const float* trans = F.xf[obj].trans;
  const float* nrm = F.xf[obj].normal;
  float p[9];
#pragma unroll
  for (int k = 0; k < 9; k++) p[k] = __ldg(S.pos + (size_t)tri * 9 + k);
  ScreenTri st;
  resolve_screen_tri<E>(F, trans, p, st);
  const bool persp = (F.flags & PRC_FRAME_PERSPECT) != 0;
  float rw1 = 1.0f, rw2 = 1.0f, rw3 = 1.0f;
  if (persp) { rw1 = __fdiv_rn(-1.0f, st.cw1); rw2 = __fdiv_rn(-1.0f, st.cw2); rw3 = __fdiv_rn(-1.0f, st.cw3); }
  VtxAttr v[3];
  v[0].pos = st.p1; v[1].pos = st.p2; v[2].pos = st.p3;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float* n = S.nor + (size_t)tri * 9 + k * 3;
    v[k].nor = apply4<EA>(V4{__ldg(n), __ldg(n + 1), __ldg(n + 2), 0.0f}, nrm);
    v[k].u = __ldg(S.uv + (size_t)tri * 6 + k * 2);
    v[k].v = __ldg(S.uv + (size_t)tri * 6 + k * 2 + 1);
    v[k].col = __ldg(S.col + (size_t)tri * 3 + k);
  }
  const int32_t material_id = __ldg(S.mat + tri);
  if (FAN && sub != 0) {
    // clipTriangle (clipping.go:67-157): fan vertex attributes by screen-space barycentrics of the parent
    V4 poly[12];
    clip_polygon<E>(st, F.cullW, F.cullH, poly);
    const int which[3] = {0, (int)sub, (int)sub + 1};
    VtxAttr c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float b[3];
      clip_bary<E>(st, poly[which[k]], b);
      c[k].pos = clip_pos(st, b);
      c[k].u = b[0] * v[0].u + b[1] * v[1].u + b[2] * v[2].u;
      c[k].v = b[0] * v[0].v + b[1] * v[1].v + b[2] * v[2].v;
      c[k].nor = V4{b[0] * v[0].nor.x + b[1] * v[1].nor.x + b[2] * v[2].nor.x, b[0] * v[0].nor.y + b[1] * v[1].nor.y + b[2] * v[2].nor.y,
                    b[0] * v[0].nor.z + b[1] * v[1].nor.z + b[2] * v[2].nor.z, 0.0f};
      uint32_t cc = 0;
#pragma unroll
      for (int ch = 0; ch < 4; ch++)
        cc |= go_u8(clampf(b[0] * (float)chan(v[0].col, ch) + b[1] * (float)chan(v[1].col, ch) + b[2] * (float)chan(v[2].col, ch), 0.0f, 255.0f)) << (8 * ch);
      c[k].col = cc;
    }
    v[0] = c[0]; v[1] = c[1]; v[2] = c[2];
  }
  // un-project without dividing by W (raster.go:467-469)
  V4 m1, m2, m3;
  if (!EA && F.unproj_std && v[0].pos.w == 1.0f && v[1].pos.w == 1.0f && v[2].pos.w == 1.0f &&
      fabsf(v[0].pos.x) + fabsf(v[0].pos.y) + fabsf(v[0].pos.z) + fabsf(v[1].pos.x) + fabsf(v[1].pos.y) + fabsf(v[1].pos.z) + fabsf(v[2].pos.x) +
              fabsf(v[2].pos.y) + fabsf(v[2].pos.z) < 1e30f) {
    // single-rounding FMA mode: the zero / one entries of the three inverse matrices make 31 of the 48 FMAs per vertex
    // identities (FMA(0, v, c) = c, FMA(m, v, +-0) = m*v for finite v); same values, only the sign of an exact zero can differ
    m1 = unproject_std(F, v[0].pos); m2 = unproject_std(F, v[1].pos); m3 = unproject_std(F, v[2].pos);
  } else {
    if (EA) {  // exact mode: this IS the hot path, keep it inline
      m1 = apply4m<EA>(apply4m<EA>(apply4m<EA>(v[0].pos, F.viewport_inv, F.pm_viewport_inv), F.proj_inv, F.pm_proj_inv), F.view_inv, F.pm_view_inv);
      m2 = apply4m<EA>(apply4m<EA>(apply4m<EA>(v[1].pos, F.viewport_inv, F.pm_viewport_inv), F.proj_inv, F.pm_proj_inv), F.view_inv, F.pm_view_inv);
      m3 = apply4m<EA>(apply4m<EA>(apply4m<EA>(v[2].pos, F.viewport_inv, F.pm_viewport_inv), F.proj_inv, F.pm_proj_inv), F.view_inv, F.pm_view_inv);
    } else {
      m1 = unproject_literal<EA>(&F, v[0].pos); m2 = unproject_literal<EA>(&F, v[1].pos); m3 = unproject_literal<EA>(&F, v[2].pos);
    }
  }
  f.facenor = unit4<EA>(cross4<EA>(sub4(m2, m1), sub4(m3, m1)));
  BarySetup bs = bary_setup<E>(v[0].pos.x, v[0].pos.y, v[1].pos.x, v[1].pos.y, v[2].pos.x, v[2].pos.y);
  const float px = (float)x + 0.5f, py = (float)y + 0.5f;
  float b0, b1, b2;
  bary_eval<E>(bs, px, py, b0, b1, b2);
  f.ok = true; f.X = x; f.Y = y;
  f.depth = b0 * v[0].pos.z + b1 * v[1].pos.z + b2 * v[2].pos.z;
  float wc1 = rw1 * b0, wc2 = rw2 * b1, wc3 = rw3 * b2;
  float norm = 1.0f;
  if (persp) norm = __fdiv_rn(1.0f, (wc1 + wc2 + wc3));
  f.u = (wc1 * v[0].u + wc2 * v[1].u + wc3 * v[2].u) * norm;
  f.v = (wc1 * v[0].v + wc2 * v[1].v + wc3 * v[2].v) * norm;
  f.du = 0.0f; f.dv = 0.0f;
  if (material_id >= 0) {  // raster.go:517-535
    float x0, x1, x2, y0, y1, y2;
    bary_eval<EA>(bs, px + 1.0f, py, x0, x1, x2);
    float wc1x = rw1 * x0, wc2x = rw2 * x1, wc3x = rw3 * x2;
    float normx = __fdiv_rn(1.0f, (wc1x + wc2x + wc3x));
    bary_eval<EA>(bs, px, py + 1.0f, y0, y1, y2);
    float wc1y = rw1 * y0, wc2y = rw2 * y1, wc3y = rw3 * y2;
    float normy = __fdiv_rn(1.0f, (wc1y + wc2y + wc3y));
    float uvdU = (wc1x * v[0].u + wc2x * v[1].u + wc3x * v[2].u) * normx;
    float uvdX = (wc1x * v[0].v + wc2x * v[1].v + wc3x * v[2].v) * normx;
    float uvdV = (wc1y * v[0].u + wc2y * v[1].u + wc3y * v[2].u) * normy;
    float uvdY = (wc1y * v[0].v + wc2y * v[1].v + wc3y * v[2].v) * normy;
    f.du = (uvdU - f.u) * (uvdU - f.u) + (uvdX - f.v) * (uvdX - f.v);
    f.dv = (uvdV - f.u) * (uvdV - f.u) + (uvdY - f.v) * (uvdY - f.v);
  }
  f.nor = unit4<EA>(V4{b0 * v[0].nor.x + b1 * v[1].nor.x + b2 * v[2].nor.x, b0 * v[0].nor.y + b1 * v[1].nor.y + b2 * v[2].nor.y,
                      b0 * v[0].nor.z + b1 * v[1].nor.z + b2 * v[2].nor.z, 0.0f});
  f.wpos = V4{b0 * m1.x + b1 * m2.x + b2 * m3.x, b0 * m1.y + b1 * m2.y + b2 * m3.y, b0 * m1.z + b1 * m2.z + b2 * m3.z, 1.0f};
  uint32_t col = 0;
#pragma unroll
  for (int ch = 0; ch < 4; ch++)
    col |= go_u8(clampf((wc1 * (float)chan(v[0].col, ch) + wc2 * (float)chan(v[1].col, ch) + wc3 * (float)chan(v[2].col, ch)) * norm, 0.0f, 255.0f)) << (8 * ch);
  f.col = col;
  f.mat = material_id;
}

// Fragments of clipped (fan) triangles are rare; their path (Sutherland-Hodgman + re-derived attributes, ~2000
// instructions and 400 B of local arrays) lives in one non-inlined function so that the hot kernels stay small enough for
// the instruction cache (ncu: 22 % of k_resolve_shade's stall samples were instruction fetches). The kernels take their
// parameter structs as __grid_constant__, so passing their addresses here does not copy them to local memory.
template <bool E, bool EA>
__device__ __noinline__ Frag resolve_fragment_fan(const DevScene* S, const DevFrame* F, uint32_t seq, int x, int y) {
  Frag f;
  resolve_fragment_impl<E, EA, true>(*S, *F, seq, x, y, f);
  return f;
}
template <bool E, bool EA>
__device__ __forceinline__ void resolve_fragment(const DevScene& S, const DevFrame& F, uint32_t seq, int x, int y, Frag& f) {
  if (seq & 7u) f = resolve_fragment_fan<E, EA>(&S, &F, seq, x, y);
  else resolve_fragment_impl<E, EA, false>(S, F, seq, x, y, f);
}

struct GBuf {
  float4 *ga, *gb, *gc, *gd;
  float* ao_depth;  // ok ? depth : -1 (material/ao.go:56-63)
};

__device__ __forceinline__ void gbuf_store(const GBuf& G, size_t idx, const Frag& f) {
  G.ga[idx] = make_float4(f.depth, f.u, f.v, f.du);
  G.gb[idx] = make_float4(f.dv, f.nor.x, f.nor.y, f.nor.z);
  G.gc[idx] = make_float4(f.facenor.x, f.facenor.y, f.facenor.z, f.wpos.x);
  G.gd[idx] = make_float4(f.wpos.y, f.wpos.z, __uint_as_float(f.col), __int_as_float(f.mat));
}
__device__ __forceinline__ void gbuf_load(const GBuf& G, size_t idx, Frag& f) {
  float4 a = G.ga[idx], b = G.gb[idx], c = G.gc[idx], d = G.gd[idx];
  f.depth = a.x; f.u = a.y; f.v = a.z; f.du = a.w; f.dv = b.x;
  f.nor = V4{b.y, b.z, b.w, 0.0f};
  f.facenor = V4{c.x, c.y, c.z, 0.0f};
  f.wpos = V4{c.w, d.x, d.y, 1.0f};
  f.col = __float_as_uint(d.z);
  f.mat = __float_as_int(d.w);
}

#ifndef PRC_RESOLVE_MIN_BLOCKS
#define PRC_RESOLVE_MIN_BLOCKS 8  // 64 registers: latency-bound gathers, measured 0.168 ms vs 0.198 at 88 registers
#endif
template <bool E, bool EA>
__global__ void __launch_bounds__(128, PRC_RESOLVE_MIN_BLOCKS) k_resolve(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F, const unsigned long long* __restrict__ keys, GBuf G) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = F.rr0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  if (x >= F.W || y >= F.rr1) return;
  const size_t idx = (size_t)y * F.W + x;
  const unsigned long long key = keys[idx];
  if (key == 0) {
    if (G.ao_depth) G.ao_depth[idx] = -1.0f;
    return;
  }
  Frag f;
  resolve_fragment<E, EA>(S, F, 0xFFFFFFFFu - (uint32_t)key, x, y, f);
  gbuf_store(G, idx, f);
  if (G.ao_depth) G.ao_depth[idx] = f.depth;
}
// pixel (0,0) when it lies outside the rasterised rows (multi-GPU strips)
template <bool E, bool EA>
__global__ void k_resolve00(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F, const unsigned long long* __restrict__ keys, GBuf G) {
  const unsigned long long key = keys[0];
  if (key == 0) return;
  Frag f;
  resolve_fragment<E, EA>(S, F, 0xFFFFFFFFu - (uint32_t)key, 0, 0, f);
  gbuf_store(G, 0, f);
}

// ---------------------------------------------------------------------------------------------
// K4: deferred shading (render/raster.go:324-359, shader/blinn_cpu.go:24-106, buffer/texture.go:83-187,
//     render/shadow.go:230-283, material/ao.go:20-73, shader/gamma.go:13-18)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rgba_at(const DevScene& S, uint32_t lvl, long long x, long long y) {
  const long long w = S.level_w[lvl], h = S.level_h[lvl];
  if (x < 0 || y < 0 || x >= w || y >= h) return 0u;
  return __ldg(reinterpret_cast<const uint32_t*>(S.tex_data + S.level_off[lvl]) + (size_t)y * w + x);
}
// the literal 64-bit sequence (negative, huge or NaN texel coordinates: wild UVs), out of line — the kernels that shade are
// instruction-cache bound (ncu: 22 % of k_resolve_shade's stall samples were instruction fetches at 11 152 SASS instructions)
__device__ __noinline__ uint32_t query_bilinear_wide(const DevScene* Sp, uint32_t lvl, float x, float y, float x0, float y0) {
  const DevScene& S = *Sp;
  const long long dx = S.level_w[lvl], dy = S.level_h[lvl];
  long long i = go_int(x0), j = go_int(y0);
  uint32_t p1 = rgba_at(S, lvl, i, j);
  uint32_t p2 = (i < dx - 1) ? rgba_at(S, lvl, i + 1, j) : p1;
  uint32_t i1 = lerpc(p1, p2, x - x0);
  uint32_t p3 = (j < dy - 1) ? rgba_at(S, lvl, i, j + 1) : p1;
  uint32_t p4 = (i < dx - 1 && j < dy - 1) ? rgba_at(S, lvl, i + 1, j + 1) : p1;
  uint32_t i2 = lerpc(p3, p4, x - x0);
  return lerpc(i1, i2, y - y0);
}
__device__ __forceinline__ uint32_t query_bilinear(const DevScene& S, uint32_t lvl, float u, float v) {
  const uint32_t wx = S.level_w[lvl], wy = S.level_h[lvl];
  if (wx == 1 && wy == 1) return rgba_at(S, lvl, 0, 0);
  float x = u * ((float)wx - 1.0f), y = v * ((float)wy - 1.0f);
  float x0 = floorf(x), y0 = floorf(y);
  if (x0 >= 0.0f && y0 >= 0.0f && x0 < 1.0e9f && y0 < 1.0e9f && wx < 0x40000000u && wy < 0x40000000u) {
    // the usual case in 32-bit integers: same texel indices, same bounds tests as the 64-bit sequence below
    const uint32_t i = (uint32_t)__float2int_rz(x0), j = (uint32_t)__float2int_rz(y0);
    const uint32_t* tex = reinterpret_cast<const uint32_t*>(S.tex_data + S.level_off[lvl]);
    const bool in_i = i < wx, in_j = j < wy, in_i1 = i + 1 < wx, in_j1 = j + 1 < wy;  // i, j >= 0 here
    const uint32_t p1 = (in_i && in_j) ? __ldg(tex + (size_t)j * wx + i) : 0u;
    const uint32_t p2 = in_i1 ? (in_j ? __ldg(tex + (size_t)j * wx + i + 1) : 0u) : p1;
    const uint32_t p3 = in_j1 ? (in_i ? __ldg(tex + (size_t)(j + 1) * wx + i) : 0u) : p1;
    const uint32_t p4 = (in_i1 && in_j1) ? __ldg(tex + (size_t)(j + 1) * wx + i + 1) : p1;
    const float tx = x - x0;
    return lerpc(lerpc(p1, p2, tx), lerpc(p3, p4, tx), y - y0);
  }
  return query_bilinear_wide(&S, lvl, x, y, x0, y0);
}
__device__ __forceinline__ void go_modf(float f, float& ip, float& fp) {
  ip = truncf(f);
  fp = isinf(f) ? NAN : f - ip;  // exact for finite f; Modf(+-Inf) = +-Inf, NaN
  if (f == 0.0f) fp = f;
}
__device__ uint32_t tex_query(const DevScene& S, const prc_material& m, float lod, float u, float v) {
  const uint32_t first = S.tex_first[m.texture];
  const int nlev = (int)(S.tex_first[m.texture + 1] - first);
  float iu, iv;
  go_modf(u, iu, u);
  if (iu != 0.0f && u == 0.0f) u = 1.0f;
  if (u < 0.0f) u = 1.0f - u;
  go_modf(v, iv, v);
  if (iv != 0.0f && v == 0.0f) v = 1.0f;
  if (v < 0.0f) v = 1.0f - v;
  if (m.flags & PRC_MAT_NO_MIPMAP) {
    float dx = (float)S.level_w[first], dy = (float)S.level_h[first];
    if (dx == 1.0f && dy == 1.0f) return rgba_at(S, first, 0, 0);
    return rgba_at(S, first, go_int(floorf(u * (dx - 1.0f))), go_int(floorf(v * (dy - 1.0f))));
  }
  if (lod < 0.0f) lod = 0.0f;
  else if (lod >= (float)nlev) lod = (float)(nlev - 1);
  // texture.go:118-137: one level (bilinear) or two (trilinear); ONE inlined copy of the bilinear fetch serves both
  uint32_t lv = first;
  float p = 0.0f;
  bool two = false;
  if (lod > 1.0f) {
    lod -= 1.0f;
    const long long h = go_int(floorf(lod));
    lv = first + (uint32_t)h;
    if (h + 1 < nlev) {
      p = lod - (float)h;
      two = !approx_eq(p, 0.0f);
    }
  }
  uint32_t L[2] = {0u, 0u};
#pragma unroll 1
  for (int k = 0; k < (two ? 2 : 1); k++) L[k] = query_bilinear(S, lv + (uint32_t)k, u, v);
  return two ? lerpc(L[0], L[1], p) : L[0];
}
// math.Log2 via float64 (math/math.go:122-124); Go: Frexp, exact for powers of two, else Log(frac)*(1/Ln2)+exp
__device__ __noinline__ float go_log2(float x) {
  int e;
  double fr = frexp((double)x, &e);
  if (fr == 0.5) return (float)(e - 1);
  return (float)(log(fr) * (1.0 / 0.693147180559945309417232121458176568) + (double)e);
}

// math.Pow of the Go standard library (src/math/pow.go) for the arguments the path produces (finite x >= 0,
// finite y > 0): x**y = x**yf * x**yi, integer part by repeated squaring of the Frexp mantissa (only IEEE
// multiplications => bit-identical to Go for integer y such as every shininess in the fixtures and AO's 10000),
// fractional part by exp(yf*log(x)). Anything else goes to the CUDA library pow (special values agree).
__device__ __noinline__ double go_pow64(double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (y == 1.0) return x;
  if (!(x > 0.0) || !(y > 0.0) || isinf(x) || isinf(y)) {
    if (x == 0.0 && y > 0.0 && !isinf(y) && !signbit(x)) return 0.0;
    return pow(x, y);
  }
  if (y == 0.5) return sqrt(x);
  double yi;
  double yf = modf(y, &yi);
  if (yi >= 9223372036854775808.0) return pow(x, y);
  double a1 = 1.0;
  long long ae = 0;
  if (yf != 0.0) {
    if (yf > 0.5) { yf -= 1.0; yi += 1.0; }
    a1 = exp(yf * log(x));
  }
  int xe_i;
  double x1;
  {
    // Frexp by bit manipulation for normal x (x is finite and > 0 here); the library call handles subnormals
    const int hi = __double2hiint(x), e = (hi >> 20) & 0x7FF;
    if (e != 0) {
      xe_i = e - 1022;
      x1 = __hiloint2double((hi & (int)0x800FFFFF) | (1022 << 20), __double2loint(x));
    } else {
      x1 = frexp(x, &xe_i);
    }
  }
  long long xe = xe_i;
  if (yi < 2147483648.0) {  // the common case in 32-bit integer arithmetic (same sequence of operations)
    int xe32 = xe_i, ae32 = 0;
    for (unsigned int i = (unsigned int)yi; i != 0; i >>= 1) {
      if (xe32 < -(1 << 12) || (1 << 12) < xe32) { ae32 += xe32; break; }
      if (i & 1) { a1 = __dmul_rn(a1, x1); ae32 += xe32; }
      x1 = __dmul_rn(x1, x1);
      xe32 <<= 1;
      if (x1 < .5) { x1 = __dadd_rn(x1, x1); xe32--; }
    }
    // Ldexp: a1 >= 2^-32 (at most 31 mantissa products), so for -980 < ae <= 0 scaling by 2^ae stays normal and is exact
    if (ae32 > -980 && ae32 <= 0 && a1 < 1e100) return __dmul_rn(a1, __hiloint2double((ae32 + 1023) << 20, 0));
    ae = ae32;
  } else
  for (long long i = (long long)yi; i != 0; i >>= 1) {
    if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
    if (i & 1) { a1 = __dmul_rn(a1, x1); ae += xe; }
    x1 = __dmul_rn(x1, x1);
    xe <<= 1;
    if (x1 < .5) { x1 = __dadd_rn(x1, x1); xe--; }
  }
  if (ae > 100000) ae = 100000;
  if (ae < -100000) ae = -100000;
  return ldexp(a1, (int)ae);
}
__device__ __forceinline__ float go_pow(float x, float y) {
  if (pow_int_unit_ok(x, y)) return pow_int_unit(x, y);  // integer shininess, base in (0,1]: bit-identical shortcut (prc_pow.h)
  return (float)go_pow64((double)x, (double)y);
}

template <bool E>
__device__ uint32_t fragment_shader(const DevScene& S, const DevFrame& F, const prc_material& m, const Frag& info) {
  float lod = 0.0f;
  if (!(m.flags & PRC_MAT_NO_MIPMAP)) {
    float siz = (float)S.level_w[S.tex_first[m.texture]] * __fsqrt_rn(go_max(info.du, info.dv));
    if (siz < 1.0f) siz = 1.0f;
    lod = go_log2(siz);
  }
  const uint32_t col = tex_query(S, m, lod, info.u, 1.0f - info.v);
  if (F.n_lights == 0) return col;
  const float cr = (float)chan(col, 0), cg = (float)chan(col, 1), cb = (float)chan(col, 2);
  float LaR = 0, LaG = 0, LaB = 0;
  for (uint32_t e = 0; e < F.n_ambient; e++) {
    const float I = F.ambient[e];
    LaR += I * cr; LaG += I * cg; LaB += I * cb;
  }
  float LdR = 0, LdG = 0, LdB = 0, LsR = 0, LsG = 0, LsB = 0;
  const V4 n = (m.flags & PRC_MAT_FLAT_SHADING) ? info.facenor : info.nor;
  const V4 x = info.wpos;
  const V4 Vv = unit4<E>(sub4(V4{F.cam[0], F.cam[1], F.cam[2], 1.0f}, x));
  for (uint32_t li = 0; li < F.n_lights; li++) {
    const DevLight& l = F.lights[li];
    V4 L = V4{0, 0, 0, 0};
    float I = 0.0f;
    if (l.kind == PRC_LIGHT_POINT) {
      V4 Ldir = sub4(V4{l.pos[0], l.pos[1], l.pos[2], 1.0f}, x);
      float ln = len4<E>(Ldir);
      float inv = __fdiv_rn(1.0f, ln);
      L = V4{Ldir.x * inv, Ldir.y * inv, Ldir.z * inv, Ldir.w * inv};
      I = __fdiv_rn(l.intensity, ln);
    } else if (l.kind == PRC_LIGHT_DIRECTIONAL) {
      L = V4{l.pos[0] * -1.0f, l.pos[1] * -1.0f, l.pos[2] * -1.0f, 0.0f};
      I = l.intensity;
    }
    V4 Hh = unit4<E>(add4(L, Vv));
    float Ld = clampf(dot4<E>(n, L), 0.0f, 1.0f);
    float Ls = go_pow(clampf(dot4<E>(n, Hh), 0.0f, 1.0f), m.shininess);
    LdR += Ld * cr * I; LdG += Ld * cg * I; LdB += Ld * cb * I;
    LsR += Ls * l.colf[0] * I; LsG += Ls * l.colf[1] * I; LsB += Ls * l.colf[2] * I;  // colf[k] = float32(chan(color, k))
  }
  float r = roundf(LaR + __fdiv_rn((float)chan(m.diffuse_rgba, 0) * LdR, 255.0f) + __fdiv_rn((float)chan(m.specular_rgba, 0) * LsR, 255.0f));
  float g = roundf(LaG + __fdiv_rn((float)chan(m.diffuse_rgba, 1) * LdG, 255.0f) + __fdiv_rn((float)chan(m.specular_rgba, 1) * LsG, 255.0f));
  float b = roundf(LaB + __fdiv_rn((float)chan(m.diffuse_rgba, 2) * LdB, 255.0f) + __fdiv_rn((float)chan(m.specular_rgba, 2) * LsB, 255.0f));
  return go_u8(clampf(r, 0.0f, 255.0f)) | (go_u8(clampf(g, 0.0f, 255.0f)) << 8) | (go_u8(clampf(b, 0.0f, 255.0f)) << 16) |
         (go_u8(clampf((float)chan(col, 3), 0.0f, 255.0f)) << 24);
}

// the literal three Apply calls of shadingVisibility (shadow.go:238-243), out of line (see query_bilinear_wide)
template <bool E>
__device__ __noinline__ V4 light_screen_generic(const DevFrame* F, const DevLight* l, V4 world) {
  return pos4(apply4m<E>(apply4m<E>(apply4m<E>(world, l->view, l->pm_view), l->proj, l->pm_proj), F->viewport, F->pm_viewport));
}
template <bool E>
__device__ bool shading_visibility(const DevFrame& F, const DevLight& l, const V4& world) {
  if (!l.cast_shadow) return true;
  // `world` = (X, Y, Depth, 1).Apply(ViewportToWorld) is the same for every light (hoisted by the caller)
  V4 sc;
  if (!E && l.persp_cam && F.vp_std && world.w == 1.0f && fabsf(world.x) + fabsf(world.y) + fabsf(world.z) < 1e30f) {
    // single-rounding FMA mode only: with the zero / one entries of a rigid view, a perspective / orthographic projection and the
    // standard viewport, FMA(0, v, c) = c and FMA(m, v, +-0) = m*v for finite v, so the 48 FMAs of the three Apply calls
    // collapse to the 14-21 operations below with the same values (only the sign of an exact zero can differ)
    const float* a = l.view;
    const float vx = fma32<false>(a[0], world.x, fma32<false>(a[1], world.y, fma32<false>(a[2], world.z, a[3])));
    const float vy = fma32<false>(a[4], world.x, fma32<false>(a[5], world.y, fma32<false>(a[6], world.z, a[7])));
    const float vz = fma32<false>(a[8], world.x, fma32<false>(a[9], world.y, fma32<false>(a[10], world.z, a[11])));
    const float* q = l.proj;
    const float* vp = F.viewport;
    if (l.persp_cam == 1u) {
      const float px = q[0] * vx, py = q[5] * vy, pz = fma32<false>(q[10], vz, q[11]), pw = -vz;
      sc = pos4(V4{fma32<false>(vp[0], px, vp[3] * pw), fma32<false>(vp[5], py, vp[7] * pw), pz, pw});
    } else {  // orthographic: w stays 1, Pos() does not divide
      const float px = fma32<false>(q[0], vx, q[3]), py = fma32<false>(q[5], vy, q[7]), pz = fma32<false>(q[10], vz, q[11]);
      sc = V4{fma32<false>(vp[0], px, vp[3]), fma32<false>(vp[5], py, vp[7]), pz, 1.0f};
    }
  } else {
    sc = light_screen_generic<E>(&F, &l, world);
  }
  if (fabsf(sc.x) < 5e8f && fabsf(sc.y) * (float)F.W < 1.5e9f) {  // 32-bit fast path: int(x) + int(y)*W cannot overflow
    const int idx = __float2int_rz(sc.x) + __float2int_rz(sc.y) * F.W;
    if (idx > 0 && idx < F.W * F.H) {
      float shadowZ = l.shadow_map[idx];
      if (sc.z < shadowZ - 0.03f) return true;
    }
    return false;
  }
  long long lx = go_int(sc.x), ly = go_int(sc.y);
  long long idx = (long long)((unsigned long long)lx + (unsigned long long)ly * (unsigned long long)F.W);
  if (idx > 0 && idx < (long long)F.W * F.H) {
    float shadowZ = l.shadow_map[idx];
    if (sc.z < shadowZ - 0.03f) return true;
  }
  return false;
}

struct AoConsts { float cosv[8], sinv[8]; float half_pi, four_pi; };

// the literal arithmetic of one AO sample (ao.go:47-70) once it is known to lie inside the buffer
__device__ __forceinline__ float ao_sample(float m, float pxf, float pyf, float cx, float cy, float elevation) {
  const float dxv = pxf - cx, dyv = pyf - cy;
  // Vec4.Len of (dx, dy, 0, 0): FMA(dx,dx, FMA(dy,dy, FMA(0,0, 0*0))); the inner FMA(dy,dy,0) rounds the exact product
  // once to float64 (exact, 48 bits) and once to float32, i.e. it is the plain float32 product
  const float dist = __fsqrt_rn(fma32<true>(dxv, dxv, dyv * dyv));
  if (dist < 1.0f) return m;
  return go_max(m, __fdiv_rn(elevation, dist));
}
// trunc(c) for 0 <= c < 2^22 without a conversion-pipe instruction: the mantissa of c + 2^23 added with round-toward-zero
__device__ __forceinline__ int trunc_pos(float c) { return __float_as_int(__fadd_rz(c, 8388608.0f)) & 0x7FFFFF; }

__device__ __forceinline__ float max_elevation(int W, int H, const float* __restrict__ ao_depth, int X, int Y, float dirX, float dirY) {
  // max over t of Atan(e_t/d_t) == Atan(max over t of e_t/d_t): Atan and the float32 rounding are monotone and
  // Go's Max is NaN-propagating on both sides, so one atan per direction reproduces ao.go:40-73 exactly.
  // The 99 samples of a ray are filtered before the literal arithmetic (exact FMA, sqrt, IEEE divide, Go Max):
  //   * t = 0 is the fragment itself: inside the buffer, distance 0 < 1 -> `continue` (ao.go:52-54);
  //   * a sample that is not higher than the fragment (elevation <= 0) cannot raise the running maximum m >= 0;
  //   * the sample distance is |dir| t up to the rounding of cur = p + dir*t on the pixel grid: |distance - t| < 7.1e-4 for
  //     coordinates below 16384 (half an ulp, 2^-11, per axis), so for t >= 2 elevation/distance < (elevation/t) * 1.002
  //     and a sample whose approximate quotient times 1.002 is below m cannot raise it either (RN is monotone).
  // Only the few remaining candidates (typically the silhouette of the nearest occluder) run the literal sequence.
  // The first `tsafe` samples cannot leave the buffer (|dir| <= 1: sample t stays within t + 1 pixels of the fragment), so
  // they run without bounds tests and the loop can be unrolled (independent depth loads in flight).
  const float pxf = (float)X, pyf = (float)Y;
  float m = 0.0f;
  const float traceDepth = ao_depth[Y * W + X];  // W*H <= 2^28
  int tsafe = 99;
  if (dirX > 0.0f) tsafe = min(tsafe, W - 2 - X); else if (dirX < 0.0f) tsafe = min(tsafe, X - 1);
  if (dirY > 0.0f) tsafe = min(tsafe, H - 2 - Y); else if (dirY < 0.0f) tsafe = min(tsafe, Y - 1);
  if (!(fabsf(dirX) <= 1.0001f && fabsf(dirY) <= 1.0001f)) tsafe = 0;  // (never for Cos/Sin; keeps the argument honest)
  float t = 0.0f;
  int ti = 1;
  if (tsafe >= 1) {  // t = 1: no filter (the distance can round to just below 1)
    t = 1.0f;
    const float cx = pxf + dirX, cy = pyf + dirY;
    m = ao_sample(m, pxf, pyf, cx, cy, ao_depth[trunc_pos(cy) * W + trunc_pos(cx)] - traceDepth);
    ti = 2;
  }
#pragma unroll 4
  for (; ti <= tsafe; ti++) {
    t += 1.0f;  // exact: the reference's float32 loop counter (ao.go:46)
    const float cx = pxf + dirX * t, cy = pyf + dirY * t;
    const float elevation = ao_depth[trunc_pos(cy) * W + trunc_pos(cx)] - traceDepth;
    if (elevation == elevation) {  // (a NaN elevation takes the literal path: Max propagates it)
      if (!(elevation > 0.0f)) continue;
      if (__fdividef(elevation, t) * 1.002f < m) continue;
    }
    m = ao_sample(m, pxf, pyf, cx, cy, elevation);
  }
  for (; ti < 100; ti++) {
    t += 1.0f;
    const float cx = pxf + dirX * t, cy = pyf + dirY * t;
    int ix, iy;
    if (fabsf(cx) < 4.0e6f && fabsf(cy) < 4.0e6f) {
      // int(cur.X), int(cur.Y) truncate toward zero: anything <= -1 is outside the buffer, (-1, 0) maps to pixel 0
      if (cx <= -1.0f || cy <= -1.0f) break;
      ix = trunc_pos(fabsf(cx)); iy = trunc_pos(fabsf(cy));
      if (ix >= W || iy >= H) break;
    } else {
      const long long lx = go_int(cx), ly = go_int(cy);
      if (lx < 0 || ly < 0 || lx >= W || ly >= H) break;
      ix = (int)lx; iy = (int)ly;
    }
    const float elevation = ao_depth[iy * W + ix] - traceDepth;
    if (ti >= 2 && elevation == elevation) {
      if (!(elevation > 0.0f)) continue;
      if (__fdividef(elevation, t) * 1.002f < m) continue;
    }
    m = ao_sample(m, pxf, pyf, cx, cy, elevation);
  }
  return (float)atan((double)m);
}
__device__ __noinline__ uint32_t ao_shade(int W, int H, const AoConsts* __restrict__ A, const float* __restrict__ ao_depth, int X, int Y, uint32_t col) {
  float total = 0.0f;
#pragma unroll 1
  for (int k = 0; k < 8; k++) total += A->half_pi - max_elevation(W, H, ao_depth, X, Y, A->cosv[k], A->sinv[k]);
  total = __fdiv_rn(total, A->four_pi);
  total = go_pow(total, 10000.0f);
  return go_u8(total * (float)chan(col, 0)) | (go_u8(total * (float)chan(col, 1)) << 8) | (go_u8(total * (float)chan(col, 2)) << 16) | (col & 0xff000000u);
}

__device__ __forceinline__ const prc_material* mat_at(const DevScene& S, int32_t id) {  // gpudeferred.go:77-82
  if (id < 0 || (uint32_t)id >= S.n_mats) return nullptr;
  const prc_material* m = S.mats + id;
  return (m->flags & PRC_MAT_NIL) ? nullptr : m;
}

// (*Renderer).shade for a fragment `frag` (own X,Y,mat) whose G-buffer record is `info`
template <bool E>
__device__ uint32_t shade_pixel(const DevScene& S, const DevFrame& F, const AoConsts* __restrict__ A, const float* ao_depth, const Frag& info, int fragX, int fragY,
                                int32_t frag_mat) {
  uint32_t col = info.col;
  const prc_material* mat = mat_at(S, frag_mat);
  if (mat != nullptr) {
    col = fragment_shader<E>(S, F, *mat, info);
    if ((F.flags & PRC_FRAME_SHADOWMAP) && (mat->flags & PRC_MAT_RECEIVE_SHADOW)) {
      float visibles = 0.0f;
      const V4 world = apply4m<E>(V4{(float)info.X, (float)info.Y, info.depth, 1.0f}, F.vtw, F.pm_vtw);
      for (uint32_t i = 0; i < F.n_lights; i++)
        if (shading_visibility<E>(F, F.lights[i], world)) visibles += 1.0f;
      float w = go_pow(0.5f, visibles);
      col = go_u8((float)chan(col, 0) * w) | (go_u8((float)chan(col, 1) * w) << 8) | (go_u8((float)chan(col, 2) * w) << 16) | (col & 0xff000000u);
    }
    if (mat->flags & PRC_MAT_AMBIENT_OCCLUSION) col = ao_shade(F.W, F.H, A, ao_depth, fragX, fragY, col);
  }
  return col;
}

// special[0] = colour of pixel (0,0) (pre-gamma), special[1] = colour of every uncovered pixel (bug-list 3:
// an uncovered pixel carries a zero Fragment, so shade() reads G(0,0) with MaterialID 0 — raster.go:326-332)
template <bool E, bool ES>
__global__ void k_shade_special(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F, const AoConsts* __restrict__ A, const unsigned long long* __restrict__ keys, GBuf G, uint32_t* special) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long key = keys[0];
  if (key == 0) {
    special[0] = F.background;
    special[1] = F.background;
    return;
  }
  // pixel (0,0) is resolved here (not read from the G-buffer): the same values, and no dependency on a resolve kernel
  Frag info;
  resolve_fragment<E, ES>(S, F, 0xFFFFFFFFu - (uint32_t)key, 0, 0, info);
  info.nor.w = 0.0f; info.facenor.w = 0.0f; info.wpos.w = 1.0f;  // what a G-buffer round trip keeps (gbuf_load)
  info.ok = true; info.X = 0; info.Y = 0;
  uint32_t c00 = shade_pixel<ES>(S, F, A, G.ao_depth, info, 0, 0, info.mat);
  special[0] = c00;
  info.col = c00;  // UnsafeSet wrote the shaded colour back into fragments[(0,0)] (raster_screen.go:86)
  special[1] = shade_pixel<ES>(S, F, A, G.ao_depth, info, 0, 0, 0);
}

#ifndef PRC_SHADE_MIN_BLOCKS
#define PRC_SHADE_MIN_BLOCKS 8  // 64 registers: measured 0.313 ms vs 0.355 at 76 registers
#endif
template <bool E>
__global__ void __launch_bounds__(128, PRC_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F, const AoConsts* __restrict__ A, const unsigned long long* __restrict__ keys, GBuf G,
                                               const uint32_t* __restrict__ special, uint32_t* __restrict__ image, uint32_t* __restrict__ host_image = nullptr) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = F.row0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  if (x >= F.W || y >= F.row1) return;
  const size_t idx = (size_t)y * F.W + x;
  uint32_t col;
  if (x == 0 && y == 0) col = special[0];
  else if (keys[idx] == 0) col = special[1];
  else {
    Frag info;
    gbuf_load(G, idx, info);
    info.ok = true; info.X = x; info.Y = y;
    col = shade_pixel<E>(S, F, A, G.ao_depth, info, x, y, info.mat);
  }
  if (F.flags & PRC_FRAME_GAMMA)
    col = (uint32_t)F.gamma[chan(col, 0)] | ((uint32_t)F.gamma[chan(col, 1)] << 8) | ((uint32_t)F.gamma[chan(col, 2)] << 16) | (col & 0xff000000u);
  if (F.flags & PRC_FRAME_BGRA) col = __byte_perm(col, 0, 0x3012);  // PixelFormatBGRA: bytes B,G,R,A (buffer.go:242-251)
  image[(size_t)(F.H - 1 - y) * F.W + x] = col;  // image row r = screen y = H-1-r (buffer.go:225)
  // zero-copy readback: the same pixel straight into the caller-visible page-locked image (a warp writes one 128-byte line),
  // so the frame crosses PCIe while it is being shaded instead of in copies that start when a band is done
  if (host_image) __stcs(host_image + (size_t)(F.H - 1 - y) * F.W + x, col);
}

// K3+K4 fused: visibility key -> attributes -> colour in one kernel, without the 64 B/pixel G-buffer round trip. The
// attribute gathers of k_resolve are latency bound and the shading arithmetic of k_shade is issue bound; in one kernel
// warps in either stage cover for each other. Used when nothing needs the G-buffer afterwards: no PRC_FRAME_KEEP_GBUFFER
// and no ambient-occlusion material (AO reads its neighbours' depths, i.e. needs every pixel resolved first). Pixel (0,0)
// is resolved by k_shade_special itself. Same device functions, same values as the two-kernel path.
#ifndef PRC_FUSED_MIN_BLOCKS
#define PRC_FUSED_MIN_BLOCKS 8
#endif
template <bool E, bool ES>
__global__ void __launch_bounds__(128, PRC_FUSED_MIN_BLOCKS) k_resolve_shade(const __grid_constant__ DevScene S, const __grid_constant__ DevFrame F, const AoConsts* __restrict__ A, const unsigned long long* __restrict__ keys,
                                                       const uint32_t* __restrict__ special, uint32_t* __restrict__ image, uint32_t* __restrict__ host_image = nullptr) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = F.row0 + blockIdx.y * 4 + (threadIdx.x >> 5);
  if (x >= F.W || y >= F.row1) return;
  const size_t idx = (size_t)y * F.W + x;
  const unsigned long long key = keys[idx];
  uint32_t col;
  if (x == 0 && y == 0) col = special[0];
  else if (key == 0) col = special[1];
  else {
    Frag info;
    resolve_fragment<E, ES>(S, F, 0xFFFFFFFFu - (uint32_t)key, x, y, info);
    info.nor.w = 0.0f; info.facenor.w = 0.0f; info.wpos.w = 1.0f;  // what a G-buffer round trip keeps (gbuf_load)
    info.ok = true; info.X = x; info.Y = y;
    col = shade_pixel<ES>(S, F, A, nullptr, info, x, y, info.mat);
  }
  if (F.flags & PRC_FRAME_GAMMA)
    col = (uint32_t)F.gamma[chan(col, 0)] | ((uint32_t)F.gamma[chan(col, 1)] << 8) | ((uint32_t)F.gamma[chan(col, 2)] << 16) | (col & 0xff000000u);
  if (F.flags & PRC_FRAME_BGRA) col = __byte_perm(col, 0, 0x3012);  // PixelFormatBGRA: bytes B,G,R,A (buffer.go:242-251)
  image[(size_t)(F.H - 1 - y) * F.W + x] = col;  // image row r = screen y = H-1-r (buffer.go:225)
  // zero-copy readback: the same pixel straight into the caller-visible page-locked image (a warp writes one 128-byte line),
  // so the frame crosses PCIe while it is being shaded instead of in copies that start when a band is done
  if (host_image) __stcs(host_image + (size_t)(F.H - 1 - y) * F.W + x, col);
}

// ---------------------------------------------------------------------------------------------
// passAntialiasing's downsample (render/raster.go:377): imageutil.Resize (internal/imageutil/resize.go:16-117), the
// two-pass fixed-point bilinear filter, both passes in one kernel. The reference filters rows into an 8-bit TRANSPOSED
// temporary and then filters that; the value of temporary pixel (row r, out column c) depends only on input row r, so each
// output pixel recomputes the `fl_y` temporaries it needs (8-bit rounding and clamping included) instead of storing them.
// Coefficients / start offsets come from createWeights8 (:146-164), computed on the host in float32 (make_weights8).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize(const uint32_t* __restrict__ in, int iw, int ih, uint32_t* __restrict__ out, int ow, int oh,
                                                const short* __restrict__ cx, const int* __restrict__ sx, int flx,
                                                const short* __restrict__ cy, const int* __restrict__ sy, int fly) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= ow || y >= oh) return;
  const int maxX = iw - 1, maxY = ih - 1;
  int acc[4] = {0, 0, 0, 0}, sum = 0;
  for (int j = 0; j < fly; j++) {
    const int cj = cy[y * fly + j];
    if (cj == 0) continue;
    int r = sy[y] + j;  // row of the temporary = input row (resizeRGBA's index clamp, :88-95)
    r = ((unsigned)r < (unsigned)maxY) ? r : (r >= maxY ? maxY : 0);
    const uint32_t* row = in + (size_t)r * iw;
    int t[4] = {0, 0, 0, 0}, ts = 0;
    for (int i = 0; i < flx; i++) {
      const int ci = cx[x * flx + i];
      if (ci == 0) continue;
      int c = sx[x] + i;
      c = ((unsigned)c < (unsigned)maxX) ? c : (c >= maxX ? maxX : 0);
      const uint32_t p = __ldg(row + c);
#pragma unroll
      for (int ch = 0; ch < 4; ch++) t[ch] += ci * (int)((p >> (8 * ch)) & 0xffu);
      ts += ci;
    }
#pragma unroll
    for (int ch = 0; ch < 4; ch++) acc[ch] += cj * min(max(t[ch] / ts, 0), 255);  // the temporary is uint8(Clamp(sum/coeffsum, 0, 255))
    sum += cj;
  }
  uint32_t o = 0;
#pragma unroll
  for (int ch = 0; ch < 4; ch++) o |= (uint32_t)min(max(acc[ch] / sum, 0), 255) << (8 * ch);
  out[(size_t)y * ow + x] = o;
}

// ---------------------------------------------------------------------------------------------
// chunk culling: a chunk = the 256 consecutive triangles one CTA of k_geom_raster handles.
//   upload:    k_chunk_aabb  -> model-space AABB of the chunk + its object (or "mixed" when it spans objects)
//   per view:  k_chunk_cull  -> 0 when no triangle of the chunk can produce a fragment in the view's pixel rows / on
//              screen. Exactness: the screen position of every vertex inside the box lies in the convex hull of the
//              eight projected corners when all corners have clip w of one sign (projective maps preserve convexity);
//              the test uses a margin of 2 px + 1e-3 relative for the float rounding of the reference sequence, and the
//              reference's own pixel loop reaches at most 1.5 px beyond a triangle's bbox. Anything else stays visible.
// ---------------------------------------------------------------------------------------------
struct ChunkBox { float mn[3], mx[3]; uint32_t obj; uint32_t mixed; };

__global__ void __launch_bounds__(256) k_chunk_aabb(const float* __restrict__ pos, const uint32_t* __restrict__ meta, uint64_t n, ChunkBox* out) {
  __shared__ float smn[8][3], smx[8][3];
  __shared__ uint32_t sobj[2];
  const uint64_t tri = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  if (tri < n) {
    for (int k = 0; k < 3; k++)
      for (int c = 0; c < 3; c++) {
        const float v = pos[tri * 9 + k * 3 + c];
        mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
        if (!(fabsf(v) < 3e38f)) { mn[c] = -INFINITY; mx[c] = INFINITY; }  // NaN / Inf vertex: unbounded box
      }
  }
  for (int c = 0; c < 3; c++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  if ((threadIdx.x & 31) == 0)
    for (int c = 0; c < 3; c++) { smn[threadIdx.x >> 5][c] = mn[c]; smx[threadIdx.x >> 5][c] = mx[c]; }
  const uint64_t last = min(n - 1, (uint64_t)blockIdx.x * 256 + 255);
  if (threadIdx.x == 0) { sobj[0] = meta[(uint64_t)blockIdx.x * 256] & 0x00FFFFFFu; sobj[1] = meta[last] & 0x00FFFFFFu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    ChunkBox b;
    for (int c = 0; c < 3; c++) {
      b.mn[c] = smn[0][c]; b.mx[c] = smx[0][c];
      for (int w = 1; w < 8; w++) { b.mn[c] = fminf(b.mn[c], smn[w][c]); b.mx[c] = fmaxf(b.mx[c], smx[w][c]); }
    }
    b.obj = sobj[0];
    b.mixed = sobj[0] != sobj[1];  // triangles are grouped by object, so equal ends mean one object
    out[blockIdx.x] = b;
  }
}

// m: the object's matrix of this view (F.xf[obj].trans or shadow_trans[obj])
__device__ __forceinline__ bool chunk_visible(const ChunkBox& b, const float* __restrict__ m, const float* __restrict__ viewport, int W, int r0, int r1, int need_pixel00) {
  float xmin = 3.4e38f, xmax = -3.4e38f, ymin = 3.4e38f, ymax = -3.4e38f, wmin = 3.4e38f, wmax = -3.4e38f;
  bool finite = true;
  for (int k = 0; k < 8; k++) {
    const float x = (k & 1) ? b.mx[0] : b.mn[0], y = (k & 2) ? b.mx[1] : b.mn[1], z = (k & 4) ? b.mx[2] : b.mn[2];
    const float cx = m[0] * x + m[1] * y + m[2] * z + m[3], cy = m[4] * x + m[5] * y + m[6] * z + m[7];
    const float cw = m[12] * x + m[13] * y + m[14] * z + m[15];
    // Apply(Viewport).Pos(): sx = (vp00*cx + vp03*cw)/cw
    const float sx = (viewport[0] * cx + viewport[3] * cw) / cw, sy = (viewport[5] * cy + viewport[7] * cw) / cw;
    finite = finite && fabsf(sx) < 1e30f && fabsf(sy) < 1e30f;
    xmin = fminf(xmin, sx); xmax = fmaxf(xmax, sx); ymin = fminf(ymin, sy); ymax = fmaxf(ymax, sy);
    wmin = fminf(wmin, cw); wmax = fmaxf(wmax, cw);
  }
  // all corners strictly on one side of w = 0, with head-room against rounding
  const float wabs = fmaxf(fabsf(wmin), fabsf(wmax));
  const bool one_side = (wmin > 1e-4f * wabs && wmin > 0.0f) || (wmax < -1e-4f * wabs && wmax < 0.0f);
  if (!(finite && one_side)) return true;
  const float mx_ = 4.0f + 1e-3f * fmaxf(fmaxf(fabsf(xmin), fabsf(xmax)), fmaxf(fabsf(ymin), fabsf(ymax)));
  const bool hit_rows = ymax + mx_ >= (float)r0 && ymin - mx_ <= (float)r1;
  const bool hit_cols = xmax + mx_ >= 0.0f && xmin - mx_ <= (float)W;
  const bool hit00 = need_pixel00 && xmin - mx_ <= 1.0f && ymin - mx_ <= 1.0f && xmax + mx_ >= 0.0f && ymax + mx_ >= 0.0f;
  return (hit_rows && hit_cols) || hit00;
}

// All views of one raster pass in one launch: per view the visibility byte of every chunk, and the compacted list of the chunks
// some view can touch (warp-aggregated append; any order: depth resolution by atomicMax is order-independent).
struct CullViews {
  int n;
  const float* trans[8];   // per view: [n_obj] matrices, `stride` floats apart (F.xf -> 32, shadow_trans -> 16)
  int stride;
  int r0[8], r1[8];
  int test[8];             // 0: the view takes every chunk (all rows, culling not forced)
  int need_pixel00;
  unsigned char* vis[8];
};
__global__ void __launch_bounds__(256) k_chunk_cull_views(const ChunkBox* __restrict__ boxes, uint32_t n_chunks, const __grid_constant__ CullViews C,
                                                          const float* __restrict__ viewport, int W, unsigned int* list, unsigned int* n_list) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool any = false;
  if (i < n_chunks) {
    const ChunkBox b = boxes[i];
    const bool testable = !b.mixed && b.mn[0] > -3e38f && b.mx[0] < 3e38f && b.mn[1] > -3e38f && b.mx[1] < 3e38f && b.mn[2] > -3e38f && b.mx[2] < 3e38f;
    for (int v = 0; v < C.n; v++) {
      bool vis = true;
      if (C.test[v] && testable) vis = chunk_visible(b, C.trans[v] + (size_t)b.obj * C.stride, viewport, W, C.r0[v], C.r1[v], C.need_pixel00);
      C.vis[v][i] = vis ? 1 : 0;
      any = any || vis;
    }
  }
  if (any) list[warp_push(n_list)] = i;
}

// ---------------------------------------------------------------------------------------------
// upload-time kernels
// ---------------------------------------------------------------------------------------------
// Chunk-local vertex sharing: the distinct (position bits, object) pairs among the 768 vertices of a 256-triangle
// chunk, numbered in order of first appearance (deterministic: a group's representative is its smallest vertex
// number). Pass 1 (cverts == nullptr) only counts; the host scans the counts; pass 2 writes vertices + indices.
__global__ void __launch_bounds__(256) k_chunk_dedupe(const float* __restrict__ pos, const uint32_t* __restrict__ meta, uint64_t n,
                                                      const uint32_t* __restrict__ cvoff, uint32_t* nv_out, float4* cverts, uint32_t* lidx) {
  __shared__ uint32_t kx[768], ky[768], kz[768], ko[768];
  __shared__ int slot[1024];
  __shared__ unsigned short num[768];
  __shared__ uint32_t wsum[8];
  const uint64_t tri = (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const bool have = tri < n;
  const uint32_t m = have ? meta[tri] : 0x80000000u;
  const bool valid = !(m & 0x80000000u);
  for (int i = threadIdx.x; i < 1024; i += 256) slot[i] = 0x7fffffff;
  for (int k = 0; k < 3; k++) {
    const int i = threadIdx.x * 3 + k;
    kx[i] = have ? __float_as_uint(pos[tri * 9 + k * 3 + 0]) : 0u;
    ky[i] = have ? __float_as_uint(pos[tri * 9 + k * 3 + 1]) : 0u;
    kz[i] = have ? __float_as_uint(pos[tri * 9 + k * 3 + 2]) : 0u;
    ko[i] = m & 0x00FFFFFFu;
  }
  __syncthreads();
  int hs[3];
  for (int k = 0; k < 3; k++) {
    const int i = threadIdx.x * 3 + k;
    hs[k] = -1;
    if (!valid) continue;
    uint32_t h = (kx[i] * 0x9E3779B1u) ^ (ky[i] * 0x85EBCA77u) ^ (kz[i] * 0xC2B2AE3Du) ^ (ko[i] * 0x27D4EB2Fu);
    h = (h ^ (h >> 15)) * 0x2C1B3C6Du;
    h = (h >> 22) & 1023u;
    for (;;) {
      const int old = atomicCAS(&slot[h], 0x7fffffff, i);
      if (old == 0x7fffffff) break;  // claimed an empty slot for this key
      if (kx[old] == kx[i] && ky[old] == ky[i] && kz[old] == kz[i] && ko[old] == ko[i]) { atomicMin(&slot[h], i); break; }
      h = (h + 1) & 1023u;  // 768 keys at most in 1024 slots: the probe always terminates
    }
    hs[k] = (int)h;
  }
  __syncthreads();
  int rep[3], cnt = 0;
  for (int k = 0; k < 3; k++) {
    rep[k] = hs[k] >= 0 ? slot[hs[k]] : -1;
    if (rep[k] == (int)threadIdx.x * 3 + k) cnt++;
  }
  // exclusive scan of cnt over the CTA
  uint32_t inc = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if ((threadIdx.x & 31) >= o) inc += t;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
  __syncthreads();
  uint32_t base = inc - cnt, total = 0;
  for (int w = 0; w < 8; w++) { if (w < (int)(threadIdx.x >> 5)) base += wsum[w]; total += wsum[w]; }
  for (int k = 0; k < 3; k++)
    if (rep[k] == (int)threadIdx.x * 3 + k) num[rep[k]] = (unsigned short)base++;
  __syncthreads();
  if (cverts == nullptr) {
    if (threadIdx.x == 0) nv_out[blockIdx.x] = total;
    return;
  }
  const uint32_t off = cvoff[blockIdx.x];
  for (int k = 0; k < 3; k++) {
    const int i = threadIdx.x * 3 + k;
    if (rep[k] == i) cverts[off + num[i]] = make_float4(__uint_as_float(kx[i]), __uint_as_float(ky[i]), __uint_as_float(kz[i]), __uint_as_float(ko[i]));
  }
  if (have) lidx[tri] = valid ? ((uint32_t)num[rep[0]] | ((uint32_t)num[rep[1]] << 10) | ((uint32_t)num[rep[2]] << 20)) : 0xFFFFFFFFu;
}

// Triangle.IsValid (geometry/primitive/triangle.go:63-80); always evaluated with the exact FMA.
__global__ void k_validate(const float* __restrict__ pos, const uint64_t* __restrict__ obj_start, uint32_t n_obj, uint64_t n, uint32_t* meta,
                           unsigned long long* n_valid) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // object of triangle i: binary search over obj_start
  uint32_t lo = 0, hi = n_obj;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (obj_start[mid] <= i) lo = mid; else hi = mid;
  }
  const float* p = pos + i * 9;
  V4 p1{p[0], p[1], p[2], 1.0f}, p2{p[3], p[4], p[5], 1.0f}, p3{p[6], p[7], p[8], 1.0f};
  V4 a = sub4(p2, p1), b = sub4(p3, p1);
  bool valid = true;
  if (approx_eq(a.x, 0.0f) && approx_eq(a.y, 0.0f) && approx_eq(a.z, 0.0f)) valid = false;
  if (approx_eq(b.x, 0.0f) && approx_eq(b.y, 0.0f) && approx_eq(b.z, 0.0f)) valid = false;
  if (valid) {
    float d = __fdiv_rn(dot4<true>(a, b), (len4<true>(a) * len4<true>(b)));
    valid = !approx_eq(d, 1.0f) && !approx_eq(d, -1.0f);
  }
  meta[i] = lo | (valid ? 0u : 0x80000000u);
  if (valid) atomicAdd(n_valid, 1ULL);
}

// covered pixels of the last frame (visibility key != 0), for the coverage-weighted roofline of the measurement harness
__global__ void __launch_bounds__(256) k_count_covered(const unsigned long long* __restrict__ keys, size_t n, unsigned long long* out) {
  unsigned int c = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) c += keys[i] != 0 ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// FP32 FMA micro-benchmark (the roofline of the shading kernels is the FP32 pipe, measured on the box because
// MEASURED_PEAKS.json has no FP32 entry): 16 independent chains per thread, `iters` x 16 FMAs, nothing else in the loop.
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
  float x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = (float)(threadIdx.x + k) * 1e-3f;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = __fmaf_rn(x[k], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 16; k++) s += x[k];
  if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace prc
