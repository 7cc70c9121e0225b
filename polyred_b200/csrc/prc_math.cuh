// prc_math.cuh — device-side float32 arithmetic with polyred's rounding contract.
//
// The translation unit is compiled with -fmad=false, so a plain `a*b + c` in float is two
// correctly-rounded operations (Go on amd64 never fuses; math/mat4.go:224-230). The reference's
// math.FMA[float32] (math/math.go:265-267) is a float64 fma rounded back to float32; fma32<true>
// reproduces it bit for bit, fma32<false> is the single-rounding fmaf that differs from it only
// in double-rounding cases (~2^-30 per operation; see DESIGN.md "Arithmetic contract").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace prc {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
struct M4 { float m[16]; };

template <bool EXACT>
__device__ __forceinline__ float fma32(float x, float y, float z) {
  if (EXACT) return __double2float_rn(__fma_rn((double)x, (double)y, (double)z));
  return __fmaf_rn(x, y, z);
}

// ---- Go conversions ----
__device__ __forceinline__ long long go_int(float v) {  // amd64 CVTTSS2SQ
  if (!(v >= -9223372036854775808.0f && v < 9223372036854775808.0f)) return (long long)0x8000000000000000ULL;
  return (long long)v;
}
__device__ __forceinline__ uint32_t go_u8(float v) {  // CVTTSS2SL, low byte
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return 0;
  return (uint32_t)((int)v) & 0xffu;
}
__device__ __forceinline__ float clampf(float n, float lo, float hi) {  // math/clamp.go:8-16 (NaN passes)
  if (n < lo) return lo;
  if (n > hi) return hi;
  return n;
}
// Go math.Max / math.Min on two values folded from -+MaxFloat64 (math/math.go:246-261): NaN-propagating.
__device__ __forceinline__ float go_max(float a, float b) {
  if (isnan(a) || isnan(b)) {
    if (a == INFINITY || b == INFINITY) return INFINITY;
    return NAN;
  }
  if (a == 0.0f && b == 0.0f) return signbit(a) ? b : a;
  return a > b ? a : b;
}
__device__ __forceinline__ float go_min(float a, float b) {
  if (isnan(a) || isnan(b)) {
    if (a == -INFINITY || b == -INFINITY) return -INFINITY;
    return NAN;
  }
  if (a == 0.0f && b == 0.0f) return signbit(a) ? a : b;
  return a < b ? a : b;
}
__device__ __forceinline__ float go_max3(float a, float b, float c) { return go_max(go_max(a, b), c); }
__device__ __forceinline__ float go_min3(float a, float b, float c) { return go_min(go_min(a, b), c); }

#define PRC_EPS 1e-7f
__device__ __forceinline__ bool approx_eq(float a, float b) { return fabsf(a - b) <= PRC_EPS; }
__device__ __forceinline__ bool approx_less(float a, float b) { return a < b && fabsf(a - b) > PRC_EPS; }
__device__ __forceinline__ bool less_eq(float a, float b) { return approx_eq(a, b) || approx_less(a, b); }  // box.go:62-64

// ---- Vec4 (math/vec4.go) ----
__device__ __forceinline__ V4 sub4(V4 v, V4 u) { return V4{v.x - u.x, v.y - u.y, v.z - u.z, v.w - u.w}; }
__device__ __forceinline__ V4 add4(V4 v, V4 u) { return V4{v.x + u.x, v.y + u.y, v.z + u.z, v.w + u.w}; }
template <bool E>
__device__ __forceinline__ float dot4(V4 v, V4 u) {  // :89-93
  return fma32<E>(v.x, u.x, fma32<E>(v.y, u.y, fma32<E>(v.z, u.z, v.w * u.w)));
}
template <bool E>
__device__ __forceinline__ float len4(V4 v) { return __fsqrt_rn(dot4<E>(v, v)); }  // :96-98 (sqrt via f64 == sqrtf)
template <bool E>
__device__ __forceinline__ V4 unit4(V4 v) {  // :101-104
  float n = __fdiv_rn(1.0f, len4<E>(v));
  return V4{v.x * n, v.y * n, v.z * n, v.w * n};
}
template <bool E>
__device__ __forceinline__ V4 apply4(V4 v, const float* __restrict__ a) {  // :108-116
  V4 r;
  r.x = fma32<E>(a[0], v.x, fma32<E>(a[1], v.y, fma32<E>(a[2], v.z, a[3] * v.w)));
  r.y = fma32<E>(a[4], v.x, fma32<E>(a[5], v.y, fma32<E>(a[6], v.z, a[7] * v.w)));
  r.z = fma32<E>(a[8], v.x, fma32<E>(a[9], v.y, fma32<E>(a[10], v.z, a[11] * v.w)));
  r.w = fma32<E>(a[12], v.x, fma32<E>(a[13], v.y, fma32<E>(a[14], v.z, a[15] * v.w)));
  return r;
}
// Apply with a per-matrix "plain" mask: bit i set <=> entry i is 0 or +-2^k, for which
// FMA(m, v, c) == RN(m*v + c) with m*v exact, i.e. a plain float multiply-add is bit-identical to the
// reference's float64 FMA (no double rounding can occur when one addend is a float32 and the other an
// exactly representable float32 product). The mask is uniform (kernel parameter), so no divergence.
template <bool E>
__device__ __forceinline__ float fma32m(bool plain, float m, float v, float c) {
  // without the exact FMA the fused single-rounding result already equals the plain one for such entries
  if (E && plain) return m * v + c;
  return fma32<E>(m, v, c);
}
template <bool E>
__device__ __forceinline__ V4 apply4m(V4 v, const float* __restrict__ a, uint32_t pm) {
  V4 r;
  r.x = fma32m<E>(pm & 1u, a[0], v.x, fma32m<E>(pm & 2u, a[1], v.y, fma32m<E>(pm & 4u, a[2], v.z, a[3] * v.w)));
  r.y = fma32m<E>(pm & 16u, a[4], v.x, fma32m<E>(pm & 32u, a[5], v.y, fma32m<E>(pm & 64u, a[6], v.z, a[7] * v.w)));
  r.z = fma32m<E>(pm & 256u, a[8], v.x, fma32m<E>(pm & 512u, a[9], v.y, fma32m<E>(pm & 1024u, a[10], v.z, a[11] * v.w)));
  r.w = fma32m<E>(pm & 4096u, a[12], v.x, fma32m<E>(pm & 8192u, a[13], v.y, fma32m<E>(pm & 16384u, a[14], v.z, a[15] * v.w)));
  return r;
}
template <bool E>
__device__ __forceinline__ V4 cross4(V4 v, V4 u) {  // :130-137
  V4 r;
  r.x = fma32<E>(v.y, u.z, -(v.z * u.y));
  r.y = fma32<E>(v.z, u.x, -(v.x * u.z));
  r.z = fma32<E>(v.x, u.y, -(v.y * u.x));
  r.w = 0;
  return r;
}
__device__ __forceinline__ V4 pos4(V4 v) {  // :140-146
  if (v.w == 1.0f || v.w == 0.0f) return V4{v.x, v.y, v.z, 1.0f};
  float invW = __fdiv_rn(1.0f, v.w);
  return V4{v.x * invW, v.y * invW, v.z * invW, 1.0f};
}
// Mat4.MulV (math/mat4.go:224-230): plain float32, left to right (no contraction: -fmad=false)
__device__ __forceinline__ V4 mulv(const float* __restrict__ m, V4 v) {
  V4 r;
  r.x = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
  r.y = m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w;
  r.z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w;
  r.w = m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w;
  return r;
}
// Mat4.MulV for a point (w = 1) when the caller knows the result's w is exactly 1 (last row (0,0,0,1))
__device__ __forceinline__ V4 mulv3(const float* __restrict__ m, float x, float y, float z) {
  V4 r;
  r.x = m[0] * x + m[1] * y + m[2] * z + m[3];
  r.y = m[4] * x + m[5] * y + m[6] * z + m[7];
  r.z = m[8] * x + m[9] * y + m[10] * z + m[11];
  r.w = 1.0f;
  return r;
}
template <bool E>
__device__ __forceinline__ float cross2z(float vx, float vy, float ux, float uy) {  // Vec3.Cross .Z (math/vec3.go:113-120)
  return fma32<E>(vx, uy, -(vy * ux));
}

// math.Barycoord (math/interpolate.go:66-79), split into the per-triangle part and the per-point part.
struct BarySetup {
  float t1x, t1y, t2x, t2y;
  float abx, aby, acx, acy, bcx, bcy;
  float Sabc;
};
template <bool E>
__device__ __forceinline__ BarySetup bary_setup(float t1x, float t1y, float t2x, float t2y, float t3x, float t3y) {
  BarySetup s;
  s.t1x = t1x; s.t1y = t1y; s.t2x = t2x; s.t2y = t2y;
  s.abx = t2x - t1x; s.aby = t2y - t1y;
  s.acx = t3x - t1x; s.acy = t3y - t1y;
  s.bcx = t3x - t2x; s.bcy = t3y - t2y;
  s.Sabc = cross2z<E>(s.abx, s.aby, s.acx, s.acy);
  return s;
}
template <bool E>
__device__ __forceinline__ void bary_eval(const BarySetup& s, float px, float py, float& w1, float& w2, float& w3) {
  float apx = px - s.t1x, apy = py - s.t1y;
  float bpx = px - s.t2x, bpy = py - s.t2y;
  float Sabp = cross2z<E>(s.abx, s.aby, apx, apy);
  float Sapc = cross2z<E>(apx, apy, s.acx, s.acy);
  float Sbcp = cross2z<E>(s.bcx, s.bcy, bpx, bpy);
  w1 = __fdiv_rn(Sbcp, s.Sabc);
  w2 = __fdiv_rn(Sapc, s.Sabc);
  w3 = __fdiv_rn(Sabp, s.Sabc);
}

// colours
__device__ __forceinline__ uint32_t chan(uint32_t c, int i) { return (c >> (8 * i)) & 0xffu; }
// math.LerpC (math/interpolate.go:55-62)
__device__ __forceinline__ uint32_t lerpc(uint32_t from, uint32_t to, float t) {
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    // from,to in [0,255] and t in [0,1) keep the value inside [0,255], where Go's uint8(float32) is a plain
    // truncation; a NaN t gives NaN -> 0 in both Go (CVTTSS2SL low byte) and CUDA (cvt.rzi of NaN is 0)
    float f = (float)chan(from, i), g = (float)chan(to, i);
    r |= ((uint32_t)__float2int_rz(f + t * (g - f)) & 0xffu) << (8 * i);
  }
  return r;
}

// depth -> order-preserving uint (z > z' <=> key > key'); -0 and +0 compare equal in the reference
// (buffer.go:279 uses `>`), so zero is canonicalised. NaN never reaches here.
__device__ __forceinline__ uint32_t depth_key(float z) {
  uint32_t u = __float_as_uint(z);
  if ((u << 1) == 0u) u = 0u;  // -0 -> +0
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

}  // namespace prc
