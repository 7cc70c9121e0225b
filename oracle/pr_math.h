// oracle/pr_math.h — TEST INFRASTRUCTURE ONLY (parity oracle). Not part of the product.
//
// CPU restatement of polyred's float32 arithmetic contract (reference: poly.red/math).
// Every function cites the reference file:line it follows. Compile with
//   g++ -O2 -ffp-contract=off -fno-fast-math
// so that `a*b + c` in float is never fused (Go on amd64/GOAMD64=v1 never fuses) and
// math.FMA[float32] is reproduced as a float64 fma rounded back to float32.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

typedef float f32;

// math.FMA[T] (math/math.go:265-267): T(math.FMA(float64(x), float64(y), float64(z)))
static inline f32 FMA(f32 x, f32 y, f32 z) { return (f32)std::fma((double)x, (double)y, (double)z); }
// math.Sqrt (math/math.go:231-233)
static inline f32 Sqrt(f32 x) { return (f32)std::sqrt((double)x); }
// math.Round (math/math.go:58-60), half away from zero
static inline f32 Round(f32 x) { return (f32)std::round((double)x); }
// math.Floor (math/math.go:102-104)
static inline f32 Floor(f32 x) { return (f32)std::floor((double)x); }

// math.Pow of the Go standard library (go1.26 src/math/pow.go), restated: special cases, then x**y =
// x**yf * x**yi with the integer part by repeated squaring of the Frexp mantissa and the fractional part by
// Exp(yf*Log(x)). For integer y (every shininess in the fixtures and AO's 10000) the result involves only
// IEEE multiplications, so this restatement is bit-identical to Go's; for fractional y it inherits libm's
// exp/log instead of Go's (<= 1 ulp of float64 apart).
static inline bool go_is_odd_int(double x) {
  if (std::fabs(x) >= 9007199254740992.0) return false;  // 1<<53: all such floats are even integers
  double xi;
  double xf = std::modf(x, &xi);
  return xf == 0 && ((long long)xi & 1) == 1;
}
static inline double go_pow64(double x, double y) {
  if (y == 0 || x == 1) return 1;
  if (y == 1) return x;
  if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
  if (x == 0) {
    if (y < 0) {
      if (std::signbit(x) && go_is_odd_int(y)) return std::copysign(HUGE_VAL, x);
      return HUGE_VAL;
    }
    if (std::signbit(x) && go_is_odd_int(y)) return x;
    return 0;
  }
  if (std::isinf(y)) {
    if (x == -1) return 1;
    if ((std::fabs(x) < 1) == (y > 0)) return 0;
    return HUGE_VAL;
  }
  if (std::isinf(x)) {
    if (x < 0) return go_pow64(1 / x, -y);
    return y < 0 ? 0 : HUGE_VAL;
  }
  if (y == 0.5) return std::sqrt(x);
  if (y == -0.5) return 1 / std::sqrt(x);
  double yi;
  double yf = std::modf(std::fabs(y), &yi);
  if (yf != 0 && x < 0) return std::numeric_limits<double>::quiet_NaN();
  if (yi >= 9223372036854775808.0) {
    if (x == -1) return 1;
    if ((std::fabs(x) < 1) == (y > 0)) return 0;
    return HUGE_VAL;
  }
  double a1 = 1.0;
  long long ae = 0;
  if (yf != 0) {
    if (yf > 0.5) { yf--; yi++; }
    a1 = std::exp(yf * std::log(x));
  }
  int xe_i;
  double x1 = std::frexp(x, &xe_i);
  long long xe = xe_i;
  for (long long i = (long long)yi; i != 0; i >>= 1) {
    if (xe < -(1 << 12) || (1 << 12) < xe) {  // overflow / underflow is certain; let Ldexp produce it
      ae += xe;
      break;
    }
    if (i & 1) { a1 *= x1; ae += xe; }
    x1 *= x1;
    xe <<= 1;
    if (x1 < .5) { x1 += x1; xe--; }
  }
  if (y < 0) { a1 = 1 / a1; ae = -ae; }
  if (ae > 100000) ae = 100000;
  if (ae < -100000) ae = -100000;
  return std::ldexp(a1, (int)ae);
}
// math.Pow (math/math.go:91-93): float64 Pow of the Go standard library, rounded to float32
static inline f32 Pow(f32 x, f32 y) { return (f32)go_pow64((double)x, (double)y); }
// math.Log2 (math/math.go:122-124); Go: Frexp, exact for powers of two, else Log(frac)*(1/Ln2)+exp
static inline f32 Log2(f32 x) {
  double d = (double)x;
  int e;
  double fr = std::frexp(d, &e);
  if (fr == 0.5) return (f32)(double)(e - 1);
  const double Ln2 = 0.693147180559945309417232121458176568;
  return (f32)(std::log(fr) * (1.0 / Ln2) + (double)e);
}
static inline f32 Atan(f32 x) { return (f32)std::atan((double)x); }
static inline f32 Cos(f32 x) { return (f32)std::cos((double)x); }
static inline f32 Sin(f32 x) { return (f32)std::sin((double)x); }
static inline f32 Abs(f32 x) { return (f32)std::fabs((double)x); }

// Go math.Max/Min on float64 (NaN-propagating, Max(+0,-0)=+0), folded from ∓MaxFloat64
// as poly.red math.Max/Min do (math/math.go:246-261).
static inline double go_max64(double x, double y) {
  if (std::isinf(x) && x > 0) return x;
  if (std::isinf(y) && y > 0) return y;
  if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
  if (x == 0 && x == y) return std::signbit(x) ? y : x;
  return x > y ? x : y;
}
static inline double go_min64(double x, double y) {
  if (std::isinf(x) && x < 0) return x;
  if (std::isinf(y) && y < 0) return y;
  if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
  if (x == 0 && x == y) return std::signbit(x) ? x : y;
  return x < y ? x : y;
}
static inline f32 Max2(f32 a, f32 b) {
  double m = -std::numeric_limits<double>::max();
  m = go_max64(m, a);
  m = go_max64(m, b);
  return (f32)m;
}
static inline f32 Max3(f32 a, f32 b, f32 c) {
  double m = -std::numeric_limits<double>::max();
  m = go_max64(m, a); m = go_max64(m, b); m = go_max64(m, c);
  return (f32)m;
}
static inline f32 Min2(f32 a, f32 b) {
  double m = std::numeric_limits<double>::max();
  m = go_min64(m, a);
  m = go_min64(m, b);
  return (f32)m;
}
static inline f32 Min3(f32 a, f32 b, f32 c) {
  double m = std::numeric_limits<double>::max();
  m = go_min64(m, a); m = go_min64(m, b); m = go_min64(m, c);
  return (f32)m;
}
// math.Clamp (math/clamp.go:8-16): NaN passes through.
static inline f32 Clamp(f32 n, f32 lo, f32 hi) {
  if (n < lo) return lo;
  if (n > hi) return hi;
  return n;
}
// math.ApproxEq / ApproxLess (math/math.go:41-48) with T=float32, epsilon = float32(1e-7)
static const f32 Epsilon = 1e-7f;
static inline bool ApproxEq(f32 a, f32 b, f32 eps) { return Abs(a - b) <= eps; }
static inline bool ApproxLess(f32 a, f32 b, f32 eps) { return a < b && Abs(a - b) > eps; }

// Go float32 -> int (amd64 CVTTSS2SQ): NaN / out of range -> MinInt64.
static inline int64_t go_int(f32 v) {
  if (!(v >= -9223372036854775808.0f && v < 9223372036854775808.0f)) return INT64_MIN;
  return (int64_t)v;
}
// Go float32 -> uint8 (amd64: CVTTSS2SL then low byte): NaN / out of int32 range -> 0.
static inline uint8_t go_u8(f32 v) {
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return 0;
  return (uint8_t)(int32_t)v;
}

struct Vec2 { f32 x, y; };
struct Vec3 { f32 x, y, z; };
struct Vec4 { f32 x, y, z, w; };
struct Mat4 { f32 m[16]; };  // row-major X00..X33 (math/mat4.go:36-43)
struct RGBA { uint8_t r, g, b, a; };

static inline RGBA unpack(uint32_t c) { return RGBA{(uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16), (uint8_t)(c >> 24)}; }
static inline uint32_t pack(RGBA c) { return (uint32_t)c.r | ((uint32_t)c.g << 8) | ((uint32_t)c.b << 16) | ((uint32_t)c.a << 24); }

// Vec4 (math/vec4.go)
static inline Vec4 sub(Vec4 v, Vec4 u) { return Vec4{v.x - u.x, v.y - u.y, v.z - u.z, v.w - u.w}; }    // :61-63
static inline Vec4 add(Vec4 v, Vec4 u) { return Vec4{v.x + u.x, v.y + u.y, v.z + u.z, v.w + u.w}; }    // :56-58
static inline Vec4 scale(Vec4 v, f32 x, f32 y, f32 z, f32 w) { return Vec4{v.x * x, v.y * y, v.z * z, v.w * w}; }  // :74-76
static inline f32 dot(Vec4 v, Vec4 u) {  // :89-93
  return FMA(v.x, u.x, FMA(v.y, u.y, FMA(v.z, u.z, v.w * u.w)));
}
static inline f32 len(Vec4 v) { return Sqrt(dot(v, v)); }  // :96-98
static inline Vec4 unit(Vec4 v) {                           // :101-104
  f32 n = 1.0f / len(v);
  return Vec4{v.x * n, v.y * n, v.z * n, v.w * n};
}
static inline Vec4 apply(Vec4 v, const Mat4& m) {  // :108-116
  const f32* a = m.m;
  f32 x = FMA(a[0], v.x, FMA(a[1], v.y, FMA(a[2], v.z, a[3] * v.w)));
  f32 y = FMA(a[4], v.x, FMA(a[5], v.y, FMA(a[6], v.z, a[7] * v.w)));
  f32 z = FMA(a[8], v.x, FMA(a[9], v.y, FMA(a[10], v.z, a[11] * v.w)));
  f32 w = FMA(a[12], v.x, FMA(a[13], v.y, FMA(a[14], v.z, a[15] * v.w)));
  return Vec4{x, y, z, w};
}
static inline Vec4 cross(Vec4 v, Vec4 u) {  // :130-137
  f32 x = FMA(v.y, u.z, -v.z * u.y);
  f32 y = FMA(v.z, u.x, -v.x * u.z);
  f32 z = FMA(v.x, u.y, -v.y * u.x);
  return Vec4{x, y, z, 0};
}
static inline Vec4 pos(Vec4 v) {  // :140-146
  if (v.w == 1 || v.w == 0) return Vec4{v.x, v.y, v.z, 1};
  f32 invW = 1.0f / v.w;
  return Vec4{v.x * invW, v.y * invW, v.z * invW, 1};
}
static inline bool is_zero(Vec4 v) {  // :67-71
  return ApproxEq(v.x, 0, Epsilon) && ApproxEq(v.y, 0, Epsilon) && ApproxEq(v.z, 0, Epsilon);
}
// Vec3.Cross (math/vec3.go:113-120), only Z is ever consumed by Barycoord
static inline f32 cross2z(f32 vx, f32 vy, f32 ux, f32 uy) { return FMA(vx, uy, -vy * ux); }

// Mat4 (math/mat4.go)
static inline Vec4 mulv(const Mat4& mm, Vec4 v) {  // :224-230, plain float32, left to right
  const f32* m = mm.m;
  f32 x = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3] * v.w;
  f32 y = m[4] * v.x + m[5] * v.y + m[6] * v.z + m[7] * v.w;
  f32 z = m[8] * v.x + m[9] * v.y + m[10] * v.z + m[11] * v.w;
  f32 w = m[12] * v.x + m[13] * v.y + m[14] * v.z + m[15] * v.w;
  return Vec4{x, y, z, w};
}
static inline Mat4 mulm(const Mat4& A, const Mat4& B) {  // :201-220
  Mat4 r;
  const f32 *m = A.m, *n = B.m;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++)
      r.m[i * 4 + j] = m[i * 4 + 0] * n[0 * 4 + j] + m[i * 4 + 1] * n[1 * 4 + j] + m[i * 4 + 2] * n[2 * 4 + j] + m[i * 4 + 3] * n[3 * 4 + j];
  return r;
}
static inline Mat4 transpose(const Mat4& A) {  // :250-257
  Mat4 r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r.m[i * 4 + j] = A.m[j * 4 + i];
  return r;
}
// Det/Inv (math/mat4.go:233-283) are host-side only (uniform setup, SURVEY a-1): they live in
// the host mirror (polyred_b200/gomath.py) and are pinned there against math/mat_test.go.

// math.Barycoord (math/interpolate.go:66-79)
static inline void barycoord(Vec2 p, Vec2 t1, Vec2 t2, Vec2 t3, f32 w[3]) {
  f32 apx = p.x - t1.x, apy = p.y - t1.y;
  f32 abx = t2.x - t1.x, aby = t2.y - t1.y;
  f32 acx = t3.x - t1.x, acy = t3.y - t1.y;
  f32 bcx = t3.x - t2.x, bcy = t3.y - t2.y;
  f32 bpx = p.x - t2.x, bpy = p.y - t2.y;
  f32 Sabc = cross2z(abx, aby, acx, acy);
  f32 Sabp = cross2z(abx, aby, apx, apy);
  f32 Sapc = cross2z(apx, apy, acx, acy);
  f32 Sbcp = cross2z(bcx, bcy, bpx, bpy);
  w[0] = Sbcp / Sabc;
  w[1] = Sapc / Sabc;
  w[2] = Sabp / Sabc;
}
// math.LerpC (math/interpolate.go:55-62): uint8(from + t*(to-from)), truncation
static inline RGBA lerpc(RGBA from, RGBA to, f32 t) {
  return RGBA{go_u8((f32)from.r + t * ((f32)to.r - (f32)from.r)), go_u8((f32)from.g + t * ((f32)to.g - (f32)from.g)),
              go_u8((f32)from.b + t * ((f32)to.b - (f32)from.b)), go_u8((f32)from.a + t * ((f32)to.a - (f32)from.a))};
}

}  // namespace orc
