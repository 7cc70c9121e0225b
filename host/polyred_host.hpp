// polyred_host.hpp — C++ host side above the C ABI of include/polyred_cuda.h.
//
// polyred is compiled code (Go) and no Go toolchain exists in the build image, so next to the purego shim a maintainer
// would add (go/cuda.go) and the Python mirror the parity tests drive (polyred_b200/), this header is the host side in
// C++: the reference's option / camera / light / material / scene interface for the render pass — same names, argument
// meaning and error behaviour — producing the prc_scene / prc_frame descriptors and calling the library.
//
//   namespace polyred::math      float32 arithmetic in the reference's expression order (math/mat4.go, vec3.go, context.go)
//   namespace polyred::camera    Perspective / Orthographic (camera/*.go)
//   namespace polyred::light     Point / Directional / Ambient (light/*.go)
//   namespace polyred::material  Texture (mip chain via imageutil.Resize) / BlinnPhong (material/material.go, buffer/texture.go)
//   namespace polyred::scene     Geometry / Group / Scene: traversal order = draw order (scene/core.go)
//   namespace polyred::render    NewRenderer(Size, Camera, Scene, ShadowMap, GammaCorrection, Background, MSAA, PixelFormat,
//                                CUDA(device)) . Render()   (render/options.go, render/raster.go:84-199, render/shadow.go:33-90)
//
// Everything here is HOST work (uniforms, flattening); all pixels come from the library. Compile with
// -ffp-contract=off: Go on amd64 never fuses a*b+c, and the uniforms must be the reference's bits.
// The library is loaded with dlopen; tests/test_cpp_host.py drives this header against the CPU oracle (same ABI, "orc_"
// prefix) and checks uniforms and frames byte for byte against the Python mirror.
#pragma once
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/polyred_cuda.h"

namespace polyred {
using f32 = float;

// ------------------------------------------------------------------------------------------------ math
namespace math {
constexpr f32 Pi = 3.14159265358979323846f;  // float32(math.Pi) (math/math.go:16)

inline f32 FMA(f32 x, f32 y, f32 z) { return (f32)std::fma((double)x, (double)y, (double)z); }  // math/math.go:265-267
inline f32 Sqrt(f32 x) { return (f32)std::sqrt((double)x); }
inline f32 Tan(f32 x) { return (f32)std::tan((double)x); }
inline f32 Cos(f32 x) { return (f32)std::cos((double)x); }
inline f32 Sin(f32 x) { return (f32)std::sin((double)x); }

struct Vec3 {
  f32 x = 0, y = 0, z = 0;
  Vec3 operator-(const Vec3& u) const { return {x - u.x, y - u.y, z - u.z}; }
  Vec3 operator+(const Vec3& u) const { return {x + u.x, y + u.y, z + u.z}; }
  Vec3 operator*(f32 s) const { return {x * s, y * s, z * s}; }
  f32 Dot(const Vec3& u) const { return FMA(x, u.x, FMA(y, u.y, z * u.z)); }  // math/vec3.go:78-82
  f32 Len() const { return Sqrt(Dot(*this)); }
  Vec3 Unit() const { const f32 n = 1.0f / Len(); return {x * n, y * n, z * n}; }  // :90-93
  Vec3 Cross(const Vec3& u) const {  // :113-120
    return {FMA(y, u.z, -z * u.y), FMA(z, u.x, -x * u.z), FMA(x, u.y, -y * u.x)};
  }
};
struct Vec4 {
  f32 x = 0, y = 0, z = 0, w = 0;
};

struct Mat4 {
  f32 m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // row-major
  f32 at(int r, int c) const { return m[r * 4 + c]; }
  static Mat4 Of(std::initializer_list<f32> v) {
    Mat4 r;
    int i = 0;
    for (f32 e : v) r.m[i++] = e;
    return r;
  }
  // Mat4.MulM (math/mat4.go:201-220): each element ((a*b + c*d) + e*f) + g*h in float32
  Mat4 MulM(const Mat4& n) const {
    Mat4 r;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) r.m[i * 4 + j] = at(i, 0) * n.at(0, j) + at(i, 1) * n.at(1, j) + at(i, 2) * n.at(2, j) + at(i, 3) * n.at(3, j);
    return r;
  }
  Vec4 MulV(const Vec4& v) const {  // :224-230
    f32 o[4];
    for (int r = 0; r < 4; r++) o[r] = at(r, 0) * v.x + at(r, 1) * v.y + at(r, 2) * v.z + at(r, 3) * v.w;
    return {o[0], o[1], o[2], o[3]};
  }
  Mat4 T() const {  // :250-257
    Mat4 r;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) r.m[i * 4 + j] = at(j, i);
    return r;
  }
  f32 Det() const;  // :233-247
  Mat4 Inv() const;  // :260-283, throws where the reference panics (zero determinant)
};

// The cofactor expansions are term tables: each term is a signed product of elements "rc" (row, column) multiplied left
// to right, the terms summed left to right — the order fixes the float32 rounding (math/mat4.go:233-283).
namespace detail {
inline f32 eval_terms(const Mat4& m, const char* terms) {
  f32 acc = 0;
  bool first = true;
  const char* p = terms;
  while (*p) {
    while (*p == ' ') p++;
    if (!*p) break;
    const char sign = *p++;
    f32 prod = 0;
    bool pf = true;
    while (*p && *p != ' ') {
      const f32 e = m.at(p[0] - '0', p[1] - '0');
      p += 2;
      prod = pf ? e : prod * e;
      pf = false;
      if (*p == '.') p++;
    }
    if (first) acc = (sign == '+') ? prod : -prod;
    else acc = (sign == '+') ? acc + prod : acc - prod;
    first = false;
  }
  return acc;
}
constexpr const char* kDet =
    "+00.11.22.33 -00.11.23.32 +00.12.23.31 -00.12.21.33 +00.13.21.32 -00.13.22.31 "
    "-01.12.23.30 +01.12.20.33 -01.13.20.32 +01.13.22.30 -01.10.22.33 +01.10.23.32 "
    "+02.13.20.31 -02.13.21.30 +02.10.21.33 -02.10.23.31 +02.11.23.30 -02.11.20.33 "
    "-03.10.21.32 +03.10.22.31 -03.11.22.30 +03.11.20.32 -03.12.20.31 +03.12.21.30";
constexpr const char* kInv[16] = {
    "+12.23.31 -13.22.31 +13.21.32 -11.23.32 -12.21.33 +11.22.33", "+03.22.31 -02.23.31 -03.21.32 +01.23.32 +02.21.33 -01.22.33",
    "+02.13.31 -03.12.31 +03.11.32 -01.13.32 -02.11.33 +01.12.33", "+03.12.21 -02.13.21 -03.11.22 +01.13.22 +02.11.23 -01.12.23",
    "+13.22.30 -12.23.30 -13.20.32 +10.23.32 +12.20.33 -10.22.33", "+02.23.30 -03.22.30 +03.20.32 -00.23.32 -02.20.33 +00.22.33",
    "+03.12.30 -02.13.30 -03.10.32 +00.13.32 +02.10.33 -00.12.33", "+02.13.20 -03.12.20 +03.10.22 -00.13.22 -02.10.23 +00.12.23",
    "+11.23.30 -13.21.30 +13.20.31 -10.23.31 -11.20.33 +10.21.33", "+03.21.30 -01.23.30 -03.20.31 +00.23.31 +01.20.33 -00.21.33",
    "+01.13.30 -03.11.30 +03.10.31 -00.13.31 -01.10.33 +00.11.33", "+03.11.20 -01.13.20 -03.10.21 +00.13.21 +01.10.23 -00.11.23",
    "+12.21.30 -11.22.30 -12.20.31 +10.22.31 +11.20.32 -10.21.32", "+01.22.30 -02.21.30 +02.20.31 -00.22.31 -01.20.32 +00.21.32",
    "+02.11.30 -01.12.30 -02.10.31 +00.12.31 +01.10.32 -00.11.32", "+01.12.20 -02.11.20 +02.10.21 -00.12.21 -01.10.22 +00.11.22"};
}  // namespace detail

inline f32 Mat4::Det() const { return detail::eval_terms(*this, detail::kDet); }
inline Mat4 Mat4::Inv() const {
  const f32 d = Det();
  if (d == 0) throw std::domain_error("math: zero determinant");
  const f32 dinv = 1.0f / d;
  Mat4 r;
  for (int i = 0; i < 16; i++) r.m[i] = dinv * detail::eval_terms(*this, detail::kInv[i]);
  return r;
}

inline Mat4 ViewportMatrix(f32 w, f32 h) {  // math/math.go:270-277
  return Mat4::Of({w / 2.0f, 0, 0, w / 2.0f, 0, h / 2.0f, 0, h / 2.0f, 0, 0, 1, 0, 0, 0, 0, 1});
}
// Vec4.Apply (math/vec4.go:108-116) and Pos (:140-146)
inline Vec4 Apply(const Vec4& v, const Mat4& m) {
  f32 o[4];
  for (int r = 0; r < 4; r++) o[r] = FMA(m.at(r, 0), v.x, FMA(m.at(r, 1), v.y, FMA(m.at(r, 2), v.z, m.at(r, 3) * v.w)));
  return {o[0], o[1], o[2], o[3]};
}
inline Vec4 Pos(const Vec4& v) {
  if (v.w == 1 || v.w == 0) return {v.x, v.y, v.z, 1};
  const f32 inv = 1.0f / v.w;
  return {v.x * inv, v.y * inv, v.z * inv, 1};
}

struct Quaternion {  // math/quaternion.go
  f32 A = 1;
  Vec3 V;
  Quaternion Mul(const Quaternion& p) const {
    Quaternion r;
    r.A = A * p.A - V.Dot(p.V);
    const Vec3 c = V.Cross(p.V);
    r.V = {(p.V.x * A + V.x * p.A) + c.x, (p.V.y * A + V.y * p.A) + c.y, (p.V.z * A + V.z * p.A) + c.z};
    return r;
  }
  Mat4 ToRoMat() const {
    const f32 w = A, x = V.x, y = V.y, z = V.z, two = 2, one = 1;
    return Mat4::Of({one - two * y * y - two * z * z, two * x * y - two * z * w, two * x * z + two * y * w, 0,
                     two * x * y + two * z * w, one - two * x * x - two * z * z, two * y * z - two * x * w, 0,
                     two * x * z - two * y * w, two * y * z + two * x * w, one - two * x * x - two * y * y, 0, 0, 0, 0, 1});
  }
};

// math.TransformContext (math/context.go:18-175): scale/translate accumulate in a matrix, rotations in a quaternion;
// ModelMatrix = internal.MulM(rotation.ToRoMat()). `version` lets a renderer see that a model matrix moved.
class TransformContext {
 public:
  virtual ~TransformContext() = default;
  void ResetContext() { context_ = Mat4(); rotation_ = Quaternion(); internal_ = Mat4(); need_ = false; version_++; }
  const Mat4& ModelMatrix() {
    if (need_) { context_ = internal_.MulM(rotation_.ToRoMat()); need_ = false; }
    return context_;
  }
  void Scale(f32 sx, f32 sy, f32 sz) { internal_ = Mat4::Of({sx, 0, 0, 0, 0, sy, 0, 0, 0, 0, sz, 0, 0, 0, 0, 1}).MulM(internal_); touched(); }
  void Translate(f32 tx, f32 ty, f32 tz) { internal_ = Mat4::Of({1, 0, 0, tx, 0, 1, 0, ty, 0, 0, 1, tz, 0, 0, 0, 1}).MulM(internal_); touched(); }
  void Rotate(const Vec3& direction, f32 angle) {
    const Vec3 u = direction.Unit();
    const f32 half = angle * 0.5f, cosa = Cos(half), sina = Sin(half);
    Quaternion q;
    q.A = cosa;
    q.V = {sina * u.x, sina * u.y, sina * u.z};
    rotation_ = q.Mul(rotation_);
    touched();
  }
  void RotateX(f32 a) { Rotate({1, 0, 0}, a); }
  void RotateY(f32 a) { Rotate({0, 1, 0}, a); }
  void RotateZ(f32 a) { Rotate({0, 0, 1}, a); }
  uint64_t version() const { return version_; }

 private:
  void touched() { need_ = true; version_++; }
  Mat4 context_, internal_;
  Quaternion rotation_;
  bool need_ = false;
  uint64_t version_ = 1;
};
}  // namespace math

// ------------------------------------------------------------------------------------------------ camera
namespace camera {
using math::Mat4;
using math::Vec3;
inline Mat4 ViewMatrix(const Vec3& pos, const Vec3& target, const Vec3& up) {  // camera/camera.go:42-55
  const Vec3 l = (target - pos).Unit();
  const Vec3 lxu = l.Cross(up).Unit();
  const Vec3 u = lxu.Cross(l).Unit();
  return Mat4::Of({lxu.x, lxu.y, lxu.z, -lxu.Dot(pos), u.x, u.y, u.z, -u.Dot(pos), -l.x, -l.y, -l.z, l.Dot(pos), 0, 0, 0, 1});
}
struct Interface {
  Vec3 position{0, 0, 1}, target{0, 0, 0}, up{0, 1, 0};
  virtual ~Interface() = default;
  virtual Mat4 ProjMatrix() const = 0;
  virtual bool Perspect() const = 0;  // option.Perspect (render/options.go:49-56)
  Mat4 ViewMatrix() const { return camera::ViewMatrix(position, target, up); }
  Vec3 Position() const { return position; }
};
struct Perspective : Interface {  // camera/perspective.go:31-47,100-111; ViewFrustum = (fov, aspect, near, far)
  f32 fov = 60, aspect = 16.0f / 9.0f, near_ = 0.01f, far_ = 1000;
  Perspective() = default;
  Perspective(Vec3 pos, Vec3 tgt, Vec3 upv, f32 fov_, f32 aspect_, f32 n, f32 f) : fov(fov_), aspect(aspect_), near_(n), far_(f) {
    position = pos; target = tgt; up = upv;
  }
  bool Perspect() const override { return true; }
  Mat4 ProjMatrix() const override {
    const f32 fv = (fov * math::Pi) / 180.0f, n = near_, f = far_, t = math::Tan(fv / 2.0f);
    return Mat4::Of({-1.0f / (aspect * t), 0, 0, 0, 0, -1.0f / t, 0, 0, 0, 0, (n + f) / (n - f), (2.0f * n * f) / (n - f), 0, 0, 1, 0});
  }
};
struct Orthographic : Interface {  // camera/orthographic.go:32-50,106-119; ViewFrustum = (l, r, b, t, near, far)
  f32 left = -1, right = 1, bottom = -1, top = 1, near_ = 1, far_ = -1;
  Orthographic() = default;
  Orthographic(Vec3 pos, Vec3 tgt, Vec3 upv, f32 l, f32 r, f32 b, f32 t, f32 n, f32 f) : left(l), right(r), bottom(b), top(t), near_(n), far_(f) {
    position = pos; target = tgt; up = upv;
  }
  bool Perspect() const override { return false; }
  Mat4 ProjMatrix() const override {
    const f32 l = left, r = right, t = top, b = bottom, n = near_, f = far_;
    return Mat4::Of({2.0f / (r - l), 0, 0, (l + r) / (l - r), 0, 2.0f / (t - b), 0, (b + t) / (b - t), 0, 0, 2.0f / (n - f), (f + n) / (f - n), 0, 0, 0, 1});
  }
};
}  // namespace camera

// ------------------------------------------------------------------------------------------------ colour / images
struct RGBA {
  uint8_t r = 0, g = 0, b = 0, a = 0;
  uint32_t Pack() const { return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16) | ((uint32_t)a << 24); }
};
namespace color {
// color.FromValue (color/color.go:33-44): uint8(Round(v*255)), half away from zero
inline RGBA FromValue(f32 r, f32 g, f32 b, f32 a) {
  auto q = [](f32 v) { return (uint8_t)(int)std::round((double)(v * 255.0f)); };
  return {q(r), q(g), q(b), q(a)};
}
// FromLinear2sRGB with T = float32 through the 1024-entry float64 table (color/srgb.go:15-93)
inline f32 FromLinear2sRGB(f32 v) {
  static const std::vector<double> lut = [] {
    std::vector<double> t(1025);
    for (int i = 0; i < 1024; i++) {
      const double x = (double)i / 1024.0;
      t[i] = x <= 0.0031308 ? x * 12.92 : 1.055 * std::pow(x, 1.0 / 2.4) - 0.055;
    }
    t[1024] = t[1023];
    return t;
  }();
  if (v <= 0) return 0;
  if (v == 1) return 1;
  const f32 i = v * 1024.0f;
  const int ifloor = (int)(long long)i & 1023;
  const f32 v0 = (f32)lut[ifloor], v1 = (f32)lut[ifloor + 1], fr = i - (f32)ifloor;
  return v0 * (1.0f - fr) + v1 * fr;
}
}  // namespace color

namespace imageutil {
// imageutil.Resize (internal/imageutil/resize.go:16-164), both sizes given: two-pass bilinear with int16*256 coefficients
struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> pix;  // RGBA8, row 0 first
};
namespace detail {
inline void weights8(int dy, int filter_length, f32 scale, std::vector<int16_t>& coeffs, std::vector<int>& start, int& flen) {
  const f32 cs = (f32)std::ceil((double)scale);
  flen = filter_length * (int)(cs > 1.0f ? cs : 1.0f);
  const f32 inv = 1.0f / scale, ff = inv < 1.0f ? inv : 1.0f;
  coeffs.assign((size_t)dy * flen, 0);
  start.assign(dy, 0);
  for (int y = 0; y < dy; y++) {
    f32 interp = scale * ((f32)y + 0.5f) - 0.5f;
    start[y] = (int)interp - flen / 2 + 1;
    interp -= (f32)start[y];
    for (int i = 0; i < flen; i++) {
      f32 in = (interp - (f32)i) * ff;
      in = (f32)std::fabs((double)in);
      const f32 k = in <= 1 ? 1 - in : 0;
      coeffs[(size_t)y * flen + i] = (int16_t)(int)(k * 256);
    }
  }
}
inline void pass(const uint8_t* in, int in_w, int in_h, uint8_t* out, int dy, const std::vector<int16_t>& coeffs, const std::vector<int>& start, int flen) {
  const int maxX = in_w - 1;
  for (int x = 0; x < in_h; x++) {
    const uint8_t* row = in + (size_t)x * in_w * 4;
    for (int y = 0; y < dy; y++) {
      int32_t acc[4] = {0, 0, 0, 0}, sum = 0;
      for (int i = 0; i < flen; i++) {
        const int32_t c = coeffs[(size_t)y * flen + i];
        if (c == 0) continue;
        int xi = start[y] + i;
        if ((unsigned)xi < (unsigned)maxX) xi *= 4;
        else if (xi >= maxX) xi = 4 * maxX;
        else xi = 0;
        for (int ch = 0; ch < 4; ch++) acc[ch] += c * (int32_t)row[xi + ch];
        sum += c;
      }
      uint8_t* o = out + ((size_t)y * in_h + x) * 4;
      for (int ch = 0; ch < 4; ch++) {
        const int v = acc[ch] / sum;
        o[ch] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
      }
    }
  }
}
}  // namespace detail
inline Image Resize(int width, int height, const Image& img) {
  if (width == img.w && height == img.h) return img;
  Image out;
  out.w = width; out.h = height;
  out.pix.resize((size_t)width * height * 4);
  std::vector<uint8_t> temp((size_t)img.h * width * 4);
  std::vector<int16_t> coeffs;
  std::vector<int> start;
  int flen;
  detail::weights8(width, 2, (f32)img.w / (f32)width, coeffs, start, flen);
  detail::pass(img.pix.data(), img.w, img.h, temp.data(), width, coeffs, start, flen);
  detail::weights8(height, 2, (f32)img.h / (f32)height, coeffs, start, flen);
  detail::pass(temp.data(), img.h, width, out.pix.data(), height, coeffs, start, flen);
  return out;
}
}  // namespace imageutil

// ------------------------------------------------------------------------------------------------ light
namespace light {
enum Kind { kPoint = 0, kDirectional = 1, kAmbient = 2 };
struct Light {
  Kind kind = kPoint;
  f32 intensity = 1;
  RGBA color{255, 255, 255, 255};
  math::Vec3 position{1, 1, 1};   // Point: position; Directional: unused by the path
  math::Vec3 direction{0, -1, 0};  // Directional: unit direction (normalised at construction, light/directional.go:38-53)
  bool cast_shadow = false;
};
inline std::shared_ptr<Light> NewPoint(f32 intensity, RGBA color, math::Vec3 position, bool cast_shadow = false) {  // light/point.go:36-49
  auto l = std::make_shared<Light>();
  l->kind = kPoint; l->intensity = intensity; l->color = color; l->position = position; l->cast_shadow = cast_shadow;
  return l;
}
inline std::shared_ptr<Light> NewDirectional(f32 intensity, RGBA color, math::Vec3 direction, bool cast_shadow = false) {
  auto l = std::make_shared<Light>();
  l->kind = kDirectional; l->intensity = intensity; l->color = color; l->direction = direction.Unit(); l->position = {0, 0, 0}; l->cast_shadow = cast_shadow;
  return l;
}
inline std::shared_ptr<Light> NewAmbient(f32 intensity, RGBA color = {255, 255, 255, 255}) {  // light/ambient.go:37-48
  auto l = std::make_shared<Light>();
  l->kind = kAmbient; l->intensity = intensity; l->color = color;
  return l;
}
}  // namespace light

// ------------------------------------------------------------------------------------------------ material
namespace material {
// buffer.Texture (buffer/texture.go:28-69): RGBA8 image + mip chain: L = int(Log2(max(dx,dy))) + 1 levels, each resized
// FROM LEVEL 0 to (dx / 2^i, dy / 2^i)
struct Texture {
  std::vector<imageutil::Image> mipmap;
  bool use_mipmap = true;
  explicit Texture(const imageutil::Image& img, bool use_mip = true) : use_mipmap(use_mip) {
    mipmap.push_back(img);
    if (img.w == 1 && img.h == 1) return;
    const int mx = img.w > img.h ? img.w : img.h;
    const int L = (int)(f32)std::log2((double)mx) + 1;
    for (int i = 1; i < L; i++) mipmap.push_back(imageutil::Resize(img.w / (1 << i), img.h / (1 << i), img));
  }
  static std::shared_ptr<Texture> Uniform(RGBA c) {  // buffer.NewUniformTexture (buffer/texture.go:15-23)
    imageutil::Image im;
    im.w = im.h = 1;
    im.pix = {c.r, c.g, c.b, c.a};
    return std::make_shared<Texture>(im);
  }
};
struct BlinnPhong {  // material.NewBlinnPhong (material/material.go:54-70)
  std::shared_ptr<Texture> texture;
  RGBA diffuse = color::FromValue(0.5f, 0.5f, 0.5f, 1.0f), specular = color::FromValue(0.5f, 0.5f, 0.5f, 1.0f);
  f32 shininess = 1;
  bool flat_shading = false, ambient_occlusion = false, receive_shadow = false;
};
inline std::shared_ptr<BlinnPhong> Default() {  // material.Default (material/pool.go:15-29)
  static std::shared_ptr<BlinnPhong> d = [] {
    auto m = std::make_shared<BlinnPhong>();
    m->texture = Texture::Uniform({0, 0, 255, 255});
    m->diffuse = color::FromValue(0.7f, 0.7f, 0.7f, 1.0f);
    m->specular = color::FromValue(0.5f, 0.5f, 0.5f, 1.0f);
    m->shininess = 30;
    return m;
  }();
  return d;
}
}  // namespace material

// ------------------------------------------------------------------------------------------------ scene
namespace scene {
using math::Mat4;
using math::Vec3;
struct Object : math::TransformContext {
  virtual void AABB(Vec3& mn, Vec3& mx) const = 0;  // model space
};
// geometry.Geometry: a triangle soup + the materials it owns (geometry/geometry.go:26-59). pos/nor: [n][3][3] (Pos.W = 1,
// Nor.W = 0), uv: [n][3][2], col: [n][3] packed RGBA8, mat: [n] geometry-LOCAL material index (negative = vertex colour)
struct Geometry : Object {
  std::vector<f32> pos, nor, uv;
  std::vector<uint32_t> col;
  std::vector<int32_t> mat;
  std::vector<std::shared_ptr<material::BlinnPhong>> materials;
  size_t NumTriangles() const { return pos.size() / 9; }
  void AABB(Vec3& mn, Vec3& mx) const override {  // geometry/mesh/mesh_triangle.go:44-48
    mn = {INFINITY, INFINITY, INFINITY};
    mx = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = 0; i + 2 < pos.size(); i += 3) {
      mn.x = std::fmin(mn.x, pos[i]); mn.y = std::fmin(mn.y, pos[i + 1]); mn.z = std::fmin(mn.z, pos[i + 2]);
      mx.x = std::fmax(mx.x, pos[i]); mx.y = std::fmax(mx.y, pos[i + 1]); mx.z = std::fmax(mx.z, pos[i + 2]);
    }
  }
};
struct LightObject : Object {  // lights live in the scene graph too (scene/scene.go:35-52: their AABBs count for Center())
  std::shared_ptr<light::Light> l;
  void AABB(Vec3& mn, Vec3& mx) const override {
    if (l->kind == light::kPoint) mn = mx = l->position;  // light/point.go:70
    else mn = mx = Vec3{0, 0, 0};                          // light/directional.go:63, light/ambient.go:55
  }
};
struct Group : Object {  // scene.Group (scene/core.go:114-219)
  std::vector<std::shared_ptr<Object>> objects;
  void Add(std::shared_ptr<Object> o) { objects.push_back(std::move(o)); }
  // Group.iterObjects (scene/core.go:197-219): leaves get THIS group's model matrix; nested groups chain g.ModelMatrix().MulM(nested)
  void Iter(const std::function<void(Object*, const Mat4&)>& fn) {
    for (auto& o : objects) {
      if (auto* g = dynamic_cast<Group*>(o.get())) g->Iter([&](Object* obj, const Mat4& m) { fn(obj, ModelMatrix().MulM(m)); });
      else fn(o.get(), ModelMatrix());
    }
  }
  void AABB(Vec3& mn, Vec3& mx) const override {  // Group.AABB (scene/group.go:21-44): union of the leaves' model-space AABBs
    bool any = false;
    const_cast<Group*>(this)->Iter([&](Object* o, const Mat4&) {
      Vec3 a, b;
      o->AABB(a, b);
      if (!any) { mn = a; mx = b; any = true; }
      else {
        mn = {std::fmin(mn.x, a.x), std::fmin(mn.y, a.y), std::fmin(mn.z, a.z)};
        mx = {std::fmax(mx.x, b.x), std::fmax(mx.y, b.y), std::fmax(mx.z, b.z)};
      }
    });
    if (!any) mn = mx = Vec3{0, 0, 0};
  }
  void Normalize() {  // Group.Normalize (scene/group.go:47-64)
    const Mat4 m = ModelMatrix();
    Vec3 a, b;
    AABB(a, b);
    const math::Vec4 mn = m.MulV({a.x, a.y, a.z, 1}), mx = m.MulV({b.x, b.y, b.z, 1});
    const Vec3 center{(mn.x + mx.x) * 0.5f, (mn.y + mx.y) * 0.5f, (mn.z + mx.z) * 0.5f};
    const f32 radius = Vec3{mx.x - mn.x, mx.y - mn.y, mx.z - mn.z}.Len() / 2.0f, fac = 1.0f / radius;
    Translate(-center.x, -center.y, -center.z);
    Scale(fac, fac, fac);
  }
};
class Scene {  // scene.Scene (scene/core.go:30-111)
 public:
  Group root;
  void Add(std::shared_ptr<Object> o) { root.Add(std::move(o)); }
  void Add(std::shared_ptr<light::Light> l) {
    auto o = std::make_shared<LightObject>();
    o->l = std::move(l);
    root.Add(o);
  }
  // Scene.IterObjects (scene/core.go:86-111): root-level leaves get root.ModelMatrix(); groups root.ModelMatrix().MulM(chain)
  void IterObjects(const std::function<void(Object*, const Mat4&)>& fn) { root.Iter(fn); }
  Vec3 Center() {  // scene/scene.go:49-52
    Vec3 a, b;
    root.AABB(a, b);
    return {(a.x + b.x) * 0.5f, (a.y + b.y) * 0.5f, (a.z + b.z) * 0.5f};
  }
};
}  // namespace scene

// ------------------------------------------------------------------------------------------------ backend (the C ABI, dlopen'ed)
class Backend {
 public:
  // `prefix` = "prc_" for libpolyred_cuda.so; the parity tests load the CPU oracle ("orc_"), which exports the same ABI
  Backend(const std::string& lib_path, const std::string& prefix, int device) {
    lib_ = dlopen(lib_path.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib_) throw std::runtime_error(std::string("polyred: cannot load ") + lib_path + ": " + dlerror() + " (there is no CPU fallback)");
    auto sym = [&](const char* name, bool required = true) -> void* {
      void* p = dlsym(lib_, (prefix + name).c_str());
      if (!p && required) throw std::runtime_error("polyred: missing symbol " + prefix + name);
      return p;
    };
    close_ = (int32_t(*)(void*))sym("close");
    last_error_ = (const char* (*)(void*))sym("last_error");
    scene_upload_ = (int32_t(*)(void*, const prc_scene*))sym("scene_upload");
    shadow_reset_ = (int32_t(*)(void*))sym("shadow_reset");
    render_ = (int32_t(*)(void*, const prc_frame*, uint8_t*))sym("render");
    host_image_ = (int32_t(*)(void*, uint64_t*, uint64_t*))sym("host_image", false);
    get_timings_ = (int32_t(*)(void*, prc_timings*))sym("get_timings");
    read_shadowmap_ = (int32_t(*)(void*, uint32_t, float*))sym("read_shadowmap");
    if (prefix == "prc_") {
      auto open = (int32_t(*)(int32_t, void**))sym("open");
      if (open(device, &ctx_) != 0 || !ctx_) throw std::runtime_error("polyred: prc_open failed (is a CUDA device visible? there is no CPU fallback)");
    } else {
      auto open = (int32_t(*)(void**))sym("open");
      if (open(&ctx_) != 0 || !ctx_) throw std::runtime_error("polyred: backend open failed");
    }
  }
  ~Backend() {
    if (ctx_ && close_) close_(ctx_);
  }
  Backend(const Backend&) = delete;
  void Check(int32_t rc, const char* what) const {
    if (rc != 0) throw std::runtime_error(std::string("polyred: ") + what + " failed: " + (last_error_ ? last_error_(ctx_) : "?"));  // errors surface
  }
  void SceneUpload(const prc_scene& s) const { Check(scene_upload_(ctx_, &s), "scene_upload"); }
  void ShadowReset() const { Check(shadow_reset_(ctx_), "shadow_reset"); }
  void Render(const prc_frame& f, uint8_t* out) const { Check(render_(ctx_, &f, out), "render"); }
  prc_timings Timings() const {
    prc_timings t{};
    Check(get_timings_(ctx_, &t), "get_timings");
    return t;
  }
  void ReadShadowMap(uint32_t light, float* out) const { Check(read_shadowmap_(ctx_, light, out), "read_shadowmap"); }
  bool HasHostImage() const { return host_image_ != nullptr; }
  // the last frame rendered with rgba_out == NULL, in place in the library's page-locked double buffer (prc_host_image)
  const uint8_t* HostImage(uint64_t* bytes) const {
    uint64_t p = 0;
    Check(host_image_(ctx_, &p, bytes), "host_image");
    return reinterpret_cast<const uint8_t*>((uintptr_t)p);
  }

 private:
  void* lib_ = nullptr;
  void* ctx_ = nullptr;
  int32_t (*close_)(void*) = nullptr;
  const char* (*last_error_)(void*) = nullptr;
  int32_t (*scene_upload_)(void*, const prc_scene*) = nullptr;
  int32_t (*shadow_reset_)(void*) = nullptr;
  int32_t (*render_)(void*, const prc_frame*, uint8_t*) = nullptr;
  int32_t (*host_image_)(void*, uint64_t*, uint64_t*) = nullptr;
  int32_t (*get_timings_)(void*, prc_timings*) = nullptr;
  int32_t (*read_shadowmap_)(void*, uint32_t, float*) = nullptr;
};

// ------------------------------------------------------------------------------------------------ render
namespace render {
using math::Mat4;
using math::Vec3;
enum PixelFormatKind { PixelFormatRGBA = 0, PixelFormatBGRA = 1 };  // buffer/buffer.go:78-79
struct option {  // render/options.go:16-33
  int Width = 800, Height = 600, MSAA = 1;
  bool ShadowMap = false, GammaCorrect = false, Perspect = false;
  std::shared_ptr<scene::Scene> Scene;
  std::shared_ptr<camera::Interface> Camera;
  RGBA Background{0, 0, 0, 0};
  int Format = PixelFormatRGBA;
  bool Debug = false;
  int Workers = 0, BatchSize = 32;  // the CPU scheduler's knobs (options.go:109-120): accepted, there is no scheduler on this path
  bool hasBlendFunc = false;
  int cudaDevice = -1;
  std::string libPath, libPrefix = "prc_";
};
using Option = std::function<void(option&)>;
inline Option Size(int w, int h) { return [=](option& o) { o.Width = w; o.Height = h; }; }
inline Option Camera(std::shared_ptr<camera::Interface> c) { return [=](option& o) { o.Camera = c; o.Perspect = c->Perspect(); }; }
inline Option Scene(std::shared_ptr<scene::Scene> s) { return [=](option& o) { o.Scene = s; }; }
inline Option ShadowMap(bool e) { return [=](option& o) { o.ShadowMap = e; }; }
inline Option GammaCorrection(bool e) { return [=](option& o) { o.GammaCorrect = e; }; }
inline Option Background(RGBA c) { return [=](option& o) { o.Background = c; }; }
inline Option MSAA(int n) { return [=](option& o) { o.MSAA = n; }; }
inline Option PixelFormat(int f) { return [=](option& o) { o.Format = f; }; }
inline Option Debug(bool e) { return [=](option& o) { o.Debug = e; }; }          // options.go:101-107
inline Option Workers(int n) { return [=](option& o) { o.Workers = n; }; }      // options.go:115-120
inline Option BatchSize(int n) { return [=](option& o) { o.BatchSize = n; }; }  // options.go:109-113
// options.go:96-99. A blend function is host code called per pixel (raster_screen.go:83-85): it cannot run on this path, and it is
// rejected rather than ignored.
inline Option Blending(std::function<RGBA(RGBA, RGBA)> f) { return [=](option& o) { o.hasBlendFunc = (bool)f; }; }
// Backend selection, the analogue of render.GPU(dev) (render/options.go:103-110)
inline Option CUDA(int device, std::string lib_path = "libpolyred_cuda.so") {
  return [=](option& o) { o.cudaDevice = device; o.libPath = lib_path; o.libPrefix = "prc_"; };
}
// test hook: any library exporting the same ABI under another symbol prefix (the CPU oracle: "orc_")
inline Option BackendLibrary(std::string lib_path, std::string prefix) {
  return [=](option& o) { o.cudaDevice = 0; o.libPath = lib_path; o.libPrefix = prefix; };
}

struct Frame {  // the *image.RGBA Render() returns: row 0 = top, stride 4*w
  int w = 0, h = 0;
  std::vector<uint8_t> pix;
};
struct FrameView {  // the same frame read in place: valid until two frames later, like the reference's double buffer (raster.go:86,201-206)
  int w = 0, h = 0;
  const uint8_t* pix = nullptr;
};

class Renderer {
 public:
  explicit Renderer(std::initializer_list<Option> opts) {
    for (auto& f : opts) f(cfg_);
    if (cfg_.libPath.empty()) throw std::invalid_argument("render: no backend selected - pass render::CUDA(device); this library has no CPU renderer");
    validate();
    backend_ = std::make_unique<Backend>(cfg_.libPath, cfg_.libPrefix, cfg_.cudaDevice);
    if (cfg_.Scene && cfg_.ShadowMap) initShadowMaps();
  }
  // render/options.go:125-141
  void Options(std::initializer_list<Option> opts) {
    for (auto& f : opts) f(cfg_);
    validate();
    if (cfg_.Scene && cfg_.ShadowMap) initShadowMaps();
  }
  const option& Config() const { return cfg_; }

  // (*Renderer).Render (render/raster.go:155-199): ONE library call for the whole frame
  Frame Render() {
    if (!cfg_.Scene || !cfg_.Camera) throw std::invalid_argument("render: Scene and Camera are required");
    ensureUploaded();
    buildFrame();
    Frame out;
    out.w = cfg_.Width; out.h = cfg_.Height;
    out.pix.resize((size_t)out.w * out.h * 4);
    backend_->Render(frame_, out.pix.data());
    if (cfg_.Debug) debugReport();
    return out;
  }
  // Zero-copy variant: the frame stays in the library's page-locked buffer (what the Go shim wraps in an *image.RGBA)
  FrameView RenderView() {
    if (!cfg_.Scene || !cfg_.Camera) throw std::invalid_argument("render: Scene and Camera are required");
    if (!backend_->HasHostImage()) throw std::logic_error("render: this backend has no host-image export");
    ensureUploaded();
    buildFrame();
    backend_->Render(frame_, nullptr);
    uint64_t bytes = 0;
    FrameView v;
    v.w = cfg_.Width; v.h = cfg_.Height;
    v.pix = backend_->HostImage(&bytes);
    if (bytes != (uint64_t)v.w * v.h * 4) throw std::logic_error("render: host image size mismatch");
    if (cfg_.Debug) debugReport();
    return v;
  }
  // The picture passShadows saves under render.Debug(true) (render/shadow.go:98-118): pixel (i, j) = uint8(depths[i + (H-j-1)*W] * 255)
  // in R, G and B, alpha 255.
  Frame ShadowMapImage(int index) const {
    Frame img;
    img.w = cfg_.Width * cfg_.MSAA; img.h = cfg_.Height * cfg_.MSAA;
    std::vector<float> z((size_t)img.w * img.h);
    backend_->ReadShadowMap((uint32_t)index, z.data());
    img.pix.resize(z.size() * 4);
    for (int j = 0; j < img.h; j++)
      for (int i = 0; i < img.w; i++) {
        const uint8_t g = (uint8_t)(int)(z[(size_t)i + (size_t)(img.h - j - 1) * img.w] * 255.0f);  // depths lie in [0, 1]: Go's uint8(float32) truncates
        uint8_t* p = &img.pix[((size_t)j * img.w + i) * 4];
        p[0] = p[1] = p[2] = g;
        p[3] = 255;
      }
    return img;
  }
  // the uniforms of the last built frame (tests compare them with the other mirrors bit for bit)
  const prc_frame& LastFrame() const { return frame_; }
  const std::vector<prc_object_xf>& LastObjects() const { return objs_; }
  const std::vector<prc_light>& LastLights() const { return lights_; }
  const std::vector<std::vector<f32>>& LastShadowTrans() const { return shadow_trans_; }

 private:
  void validate() const {
    if (cfg_.MSAA < 1 || cfg_.MSAA > 8) throw std::invalid_argument("render: MSAA must be in 1..8");
    if (cfg_.Format != PixelFormatRGBA && cfg_.Format != PixelFormatBGRA) throw std::invalid_argument("render: unknown PixelFormat");
    if (cfg_.hasBlendFunc) throw std::invalid_argument("render: Blending is not on the CUDA path (SURVEY 8f-4)");
  }
  // render.Debug(true): the pass timings profiling.Timed prints (raster.go:156-161, 228, 277) and shadow-<i>.ppm per casting light
  // (the reference saves shadow-<i>.png, shadow.go:98-118; same pixels, a container that needs no PNG encoder)
  void debugReport() {
    const prc_timings t = backend_->Timings();
    std::printf("forward pass (shadow): %.3f ms\nforward pass (world): %.3f ms\ndeferred pass (shading): %.3f ms\nentire rendering: %.3f ms\n",
                t.shadow_ms, t.forward_ms, t.shade_ms, t.total_ms);
    if (!cfg_.ShadowMap) return;
    for (size_t i = 0; i < flat_.sources.size(); i++) {
      if (!flat_.sources[i]->cast_shadow) continue;
      const std::string file = "shadow-" + std::to_string(i) + ".ppm";
      std::printf("saving (shadow map)... %s\n", file.c_str());
      const Frame img = ShadowMapImage((int)i);
      if (FILE* f = std::fopen(file.c_str(), "wb")) {
        std::fprintf(f, "P6\n%d %d\n255\n", img.w, img.h);
        for (size_t k = 0; k < img.pix.size(); k += 4) std::fwrite(&img.pix[k], 1, 3, f);
        std::fclose(f);
      }
    }
  }
  struct Flat {  // flattened scene = prc_scene (render/raster.go:241-270), built once per membership
    std::vector<scene::Geometry*> geos;
    std::vector<f32> pos, nor, uv;
    std::vector<uint32_t> col;
    std::vector<int32_t> mat;
    std::vector<uint64_t> obj_start;
    std::vector<prc_material> mats;
    std::vector<uint32_t> tex_first, level_w, level_h;
    std::vector<uint64_t> level_off;
    std::vector<uint8_t> tex_data;
    std::vector<std::shared_ptr<light::Light>> sources, envs;
    std::vector<const void*> membership;
  };
  std::vector<const void*> membership() {
    std::vector<const void*> ids;
    cfg_.Scene->IterObjects([&](scene::Object* o, const Mat4&) {
      ids.push_back(o);
      if (auto* g = dynamic_cast<scene::Geometry*>(o)) ids.push_back(g->pos.data());
    });
    return ids;
  }
  void ensureUploaded() {
    auto ids = membership();
    if (flat_for_ != cfg_.Scene.get() || ids != flat_.membership) {
      flatten();
      flat_.membership = std::move(ids);
      flat_for_ = cfg_.Scene.get();
      prc_scene s{};
      s.abi_version = PRC_ABI_VERSION;
      s.n_tris = flat_.pos.size() / 9;
      s.pos = flat_.pos.data(); s.nor = flat_.nor.data(); s.uv = flat_.uv.data(); s.col = flat_.col.data(); s.mat = flat_.mat.data();
      s.n_objects = (uint32_t)flat_.geos.size(); s.n_materials = (uint32_t)flat_.mats.size();
      s.obj_tri_start = flat_.obj_start.data(); s.materials = flat_.mats.data();
      s.n_textures = (uint32_t)flat_.tex_first.size() - 1; s.n_tex_levels = (uint32_t)flat_.level_w.size();
      s.tex_first_level = flat_.tex_first.data(); s.level_w = flat_.level_w.data(); s.level_h = flat_.level_h.data();
      s.level_offset = flat_.level_off.data(); s.tex_data = flat_.tex_data.data(); s.tex_bytes = flat_.tex_data.size();
      backend_->SceneUpload(s);
    }
    if (shadow_reset_) { backend_->ShadowReset(); shadow_reset_ = false; }
  }
  void flatten() {
    flat_ = Flat();
    flat_.obj_start.push_back(0);
    std::vector<material::Texture*> textures;
    cfg_.Scene->IterObjects([&](scene::Object* o, const Mat4&) {
      if (auto* lo = dynamic_cast<scene::LightObject*>(o)) {  // Scene.Lights (scene/scene.go:35-47): traversal order
        (lo->l->kind == light::kAmbient ? flat_.envs : flat_.sources).push_back(lo->l);
        return;
      }
      auto* g = dynamic_cast<scene::Geometry*>(o);
      if (!g) return;
      const size_t n = g->NumTriangles();
      flat_.geos.push_back(g);
      flat_.pos.insert(flat_.pos.end(), g->pos.begin(), g->pos.end());
      auto padded = [&](std::vector<f32>& dst, const std::vector<f32>& src, size_t per) {
        if (src.size() == n * per) dst.insert(dst.end(), src.begin(), src.end());
        else dst.insert(dst.end(), n * per, 0.0f);
      };
      padded(flat_.nor, g->nor, 9);
      padded(flat_.uv, g->uv, 6);
      if (g->col.size() == n * 3) flat_.col.insert(flat_.col.end(), g->col.begin(), g->col.end());
      else flat_.col.insert(flat_.col.end(), n * 3, 0xFFFFFFFFu);
      const int32_t base = (int32_t)flat_.mats.size();  // flat material table (raster.go:252-262)
      for (size_t i = 0; i < n; i++) {
        const int32_t m = g->mat.size() == n ? g->mat[i] : 0;
        flat_.mat.push_back(m >= 0 ? m + base : m);
      }
      for (auto& m : g->materials) {
        prc_material pm{};
        if (!m || !m->texture) { pm.flags = PRC_MAT_NIL; pm.texture = -1; flat_.mats.push_back(pm); continue; }
        size_t ti = 0;
        while (ti < textures.size() && textures[ti] != m->texture.get()) ti++;  // textures deduplicated by identity
        if (ti == textures.size()) textures.push_back(m->texture.get());
        pm.diffuse_rgba = m->diffuse.Pack(); pm.specular_rgba = m->specular.Pack(); pm.shininess = m->shininess; pm.texture = (int32_t)ti;
        pm.flags = (m->flat_shading ? PRC_MAT_FLAT_SHADING : 0) | (m->ambient_occlusion ? PRC_MAT_AMBIENT_OCCLUSION : 0) |
                   (m->receive_shadow ? PRC_MAT_RECEIVE_SHADOW : 0) | (m->texture->use_mipmap ? 0 : PRC_MAT_NO_MIPMAP);
        flat_.mats.push_back(pm);
      }
      flat_.obj_start.push_back(flat_.pos.size() / 9);
    });
    flat_.tex_first.push_back(0);
    for (auto* t : textures) {
      for (auto& lv : t->mipmap) {
        flat_.level_w.push_back((uint32_t)lv.w); flat_.level_h.push_back((uint32_t)lv.h); flat_.level_off.push_back(flat_.tex_data.size());
        flat_.tex_data.insert(flat_.tex_data.end(), lv.pix.begin(), lv.pix.end());
      }
      flat_.tex_first.push_back((uint32_t)flat_.level_w.size());
    }
    if (flat_.mats.empty()) flat_.mats.push_back(prc_material{});
    if (flat_.level_w.empty()) { flat_.level_w.push_back(0); flat_.level_h.push_back(0); flat_.level_off.push_back(0); flat_.tex_data.assign(4, 0); }
  }
  // render/shadow.go:33-90: an orthographic light camera per casting point light, fitted to the view frustum
  void initShadowMaps() {
    light_cams_.clear();
    std::vector<std::shared_ptr<light::Light>> sources;
    cfg_.Scene->IterObjects([&](scene::Object* o, const Mat4&) {
      if (auto* lo = dynamic_cast<scene::LightObject*>(o))
        if (lo->l->kind != light::kAmbient) sources.push_back(lo->l);
    });
    light_cams_.resize(sources.size());
    const Vec3 center = cfg_.Scene->Center();
    for (size_t i = 0; i < sources.size(); i++) {
      auto& l = sources[i];
      if (!l->cast_shadow) continue;
      // shadow.go:67-86: only *light.Point gets a camera; a casting Directional leaves it nil and passShadows panics
      if (l->kind != light::kPoint) throw std::logic_error("render: a shadow-casting Directional light has no light camera in the reference (it panics)");
      const Mat4 tm = camera::ViewMatrix(l->position, center, {0, 1, 0}).MulM(cfg_.Camera->ViewMatrix().Inv()).MulM(cfg_.Camera->ProjMatrix().Inv());
      static const f32 cs[8][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {-1, 1, 1}, {-1, -1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, -1}};
      Vec3 mn{INFINITY, INFINITY, INFINITY}, mx{-INFINITY, -INFINITY, -INFINITY};
      for (auto& c : cs) {
        const math::Vec4 v = math::Pos(math::Apply({c[0], c[1], c[2], 1}, tm));
        mn = {std::fmin(mn.x, v.x), std::fmin(mn.y, v.y), std::fmin(mn.z, v.z)};
        mx = {std::fmax(mx.x, v.x), std::fmax(mx.y, v.y), std::fmax(mx.z, v.z)};
      }
      light_cams_[i] = std::make_shared<camera::Orthographic>(l->position, center, Vec3{0, 1, 0}, mn.x, mx.x, mn.y, mx.y, mx.z, mn.z - 2.0f);
    }
    shadow_reset_ = true;  // NewRenderer / Options re-create (zero) the shadow maps (shadow.go:87)
  }
  static void put(f32 dst[16], const Mat4& m) { std::memcpy(dst, m.m, 64); }
  void buildFrame() {
    const int W = cfg_.Width * cfg_.MSAA, H = cfg_.Height * cfg_.MSAA;  // resetBufs (raster.go:149)
    const Mat4 view = cfg_.Camera->ViewMatrix(), proj = cfg_.Camera->ProjMatrix(), vp = math::ViewportMatrix((f32)W, (f32)H);
    const Mat4 view_inv = view.Inv(), proj_inv = proj.Inv(), vp_inv = vp.Inv();
    const Mat4 pv = proj.MulM(view);
    const size_t nobj = flat_.geos.size();
    objs_.assign(nobj ? nobj : 1, prc_object_xf{});
    std::vector<Mat4> models(nobj);
    size_t k = 0;
    cfg_.Scene->IterObjects([&](scene::Object* o, const Mat4& chain) {
      auto* g = dynamic_cast<scene::Geometry*>(o);
      if (!g) return;
      models[k] = chain.MulM(g->ModelMatrix());              // raster.go:242
      put(objs_[k].normal, models[k].Inv().T());             // raster.go:243
      put(objs_[k].trans, pv.MulM(models[k]));               // raster.go:382
      k++;
    });
    lights_.assign(flat_.sources.size() ? flat_.sources.size() : 1, prc_light{});
    shadow_trans_.assign(flat_.sources.size(), {});
    for (size_t i = 0; i < flat_.sources.size(); i++) {
      auto& l = *flat_.sources[i];
      prc_light& pl = lights_[i];
      pl.kind = l.kind == light::kPoint ? PRC_LIGHT_POINT : PRC_LIGHT_DIRECTIONAL;
      const Vec3 v = l.kind == light::kPoint ? l.position : l.direction;
      pl.pos[0] = v.x; pl.pos[1] = v.y; pl.pos[2] = v.z;
      pl.intensity = l.intensity;
      pl.color_rgba = l.color.Pack();
      pl.cast_shadow = (l.cast_shadow && cfg_.ShadowMap) ? 1 : 0;
      if (pl.cast_shadow) {
        const auto& cam = light_cams_.at(i);
        const Mat4 lv = cam->ViewMatrix(), lp = cam->ProjMatrix(), lpv = lp.MulM(lv);
        put(pl.view, lv); put(pl.proj, lp);
        shadow_trans_[i].resize(nobj * 16);
        for (size_t o = 0; o < nobj; o++) std::memcpy(&shadow_trans_[i][o * 16], lpv.MulM(models[o]).m, 64);  // shadow.go:155
        pl.shadow_trans = shadow_trans_[i].data();
      }
    }
    ambient_.clear();
    for (auto& e : flat_.envs) ambient_.push_back(e->intensity);
    if (ambient_.empty()) ambient_.push_back(0);
    prc_frame& s = frame_;
    s = prc_frame{};
    s.abi_version = PRC_ABI_VERSION;
    s.flags = (cfg_.Perspect ? PRC_FRAME_PERSPECT : 0) | (cfg_.ShadowMap ? PRC_FRAME_SHADOWMAP : 0) | (cfg_.GammaCorrect ? PRC_FRAME_GAMMA : 0) |
              (cfg_.Format == PixelFormatBGRA ? PRC_FRAME_BGRA : 0);
    s.width = (uint32_t)W; s.height = (uint32_t)H; s.msaa = (uint32_t)cfg_.MSAA;
    s.n_objects = (uint32_t)nobj; s.n_lights = (uint32_t)flat_.sources.size(); s.n_ambient = (uint32_t)flat_.envs.size();
    s.background_rgba = cfg_.Background.Pack();
    s.objects = objs_.data(); s.lights = lights_.data(); s.ambient_intensity = ambient_.data();
    put(s.viewport, vp); put(s.viewport_inv, vp_inv); put(s.proj_inv, proj_inv); put(s.view_inv, view_inv);
    put(s.viewport_to_world, view_inv.MulM(proj_inv).MulM(vp_inv));  // raster.go:287
    const Vec3 c = cfg_.Camera->Position();
    s.cam_pos[0] = c.x; s.cam_pos[1] = c.y; s.cam_pos[2] = c.z;
    for (int i = 0; i < 256; i++)  // shader.GammaCorrection as a table (shader/gamma.go:13-18)
      s.gamma_lut[i] = (uint8_t)(long long)(color::FromLinear2sRGB((f32)i / 255.0f) * 255.0f + 0.5f);
    s.row0 = 0; s.row1 = (uint32_t)H;
  }

  option cfg_;
  std::unique_ptr<Backend> backend_;
  Flat flat_;
  const void* flat_for_ = nullptr;
  std::vector<std::shared_ptr<camera::Orthographic>> light_cams_;
  bool shadow_reset_ = false;
  prc_frame frame_{};
  std::vector<prc_object_xf> objs_;
  std::vector<prc_light> lights_;
  std::vector<std::vector<f32>> shadow_trans_;
  std::vector<f32> ambient_;
};

inline std::unique_ptr<Renderer> NewRenderer(std::initializer_list<Option> opts) { return std::make_unique<Renderer>(opts); }  // render/raster.go:84-143
}  // namespace render
}  // namespace polyred
