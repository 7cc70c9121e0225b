// capi.cpp — a flat C API over polyred_host.hpp, so the C++ host mirror can be driven (and checked against the Python
// mirror, bit for bit) from tests/test_cpp_host.py through ctypes. Handles are heap-allocated shared_ptr boxes.
#include <cstdio>
#include <cstring>

#include "polyred_host.hpp"

using namespace polyred;

namespace {
template <class T>
struct Box {
  std::shared_ptr<T> p;
};
void set_err(char* err, int n, const std::exception& e) {
  if (err && n > 0) std::snprintf(err, (size_t)n, "%s", e.what());
}
RGBA unpack(uint32_t c) { return RGBA{(uint8_t)c, (uint8_t)(c >> 8), (uint8_t)(c >> 16), (uint8_t)(c >> 24)}; }
}  // namespace

extern "C" {

void* pth_scene_new() { return new Box<scene::Scene>{std::make_shared<scene::Scene>()}; }
void* pth_group_new() { return new Box<scene::Object>{std::make_shared<scene::Group>()}; }

void* pth_texture_new(int w, int h, const uint8_t* rgba, int use_mipmap) {
  imageutil::Image im;
  im.w = w; im.h = h;
  im.pix.assign(rgba, rgba + (size_t)w * h * 4);
  return new Box<material::Texture>{std::make_shared<material::Texture>(im, use_mipmap != 0)};
}
// mip level `level` of a texture: returns width/height and copies the pixels when `out` is given
int pth_texture_level(void* tex, int level, int* w, int* h, uint8_t* out) {
  auto& t = *static_cast<Box<material::Texture>*>(tex)->p;
  if (level < 0 || level >= (int)t.mipmap.size()) return -1;
  *w = t.mipmap[level].w; *h = t.mipmap[level].h;
  if (out) std::memcpy(out, t.mipmap[level].pix.data(), t.mipmap[level].pix.size());
  return (int)t.mipmap.size();
}
void* pth_material_new(void* tex, uint32_t diffuse, uint32_t specular, float shininess, int flat, int ao, int recv) {
  auto m = std::make_shared<material::BlinnPhong>();
  if (tex) m->texture = static_cast<Box<material::Texture>*>(tex)->p;
  m->diffuse = unpack(diffuse); m->specular = unpack(specular); m->shininess = shininess;
  m->flat_shading = flat != 0; m->ambient_occlusion = ao != 0; m->receive_shadow = recv != 0;
  return new Box<material::BlinnPhong>{m};
}
void* pth_geometry_new(uint64_t n, const float* pos, const float* nor, const float* uv, const uint32_t* col, const int32_t* mat, void** materials, int nmat) {
  auto g = std::make_shared<scene::Geometry>();
  g->pos.assign(pos, pos + n * 9);
  if (nor) g->nor.assign(nor, nor + n * 9);
  if (uv) g->uv.assign(uv, uv + n * 6);
  if (col) g->col.assign(col, col + n * 3);
  if (mat) g->mat.assign(mat, mat + n);
  for (int i = 0; i < nmat; i++) g->materials.push_back(materials[i] ? static_cast<Box<material::BlinnPhong>*>(materials[i])->p : nullptr);
  return new Box<scene::Object>{g};
}
void pth_group_add(void* group, void* obj) {
  static_cast<scene::Group*>(static_cast<Box<scene::Object>*>(group)->p.get())->Add(static_cast<Box<scene::Object>*>(obj)->p);
}
void pth_scene_add(void* sc, void* obj) { static_cast<Box<scene::Scene>*>(sc)->p->Add(static_cast<Box<scene::Object>*>(obj)->p); }
void pth_scene_add_light(void* sc, int kind, float intensity, uint32_t color, float x, float y, float z, int cast) {
  auto& s = *static_cast<Box<scene::Scene>*>(sc)->p;
  if (kind == 0) s.Add(light::NewPoint(intensity, unpack(color), {x, y, z}, cast != 0));
  else if (kind == 1) s.Add(light::NewDirectional(intensity, unpack(color), {x, y, z}, cast != 0));
  else s.Add(light::NewAmbient(intensity, unpack(color)));
}
// op: 0 Scale(a,b,c)  1 Translate(a,b,c)  2 Rotate(axis (a,b,c), angle d)  3 ResetContext  4 Normalize (groups)
void pth_xf(void* obj, int op, float a, float b, float c, float d) {
  scene::Object* o = static_cast<Box<scene::Object>*>(obj)->p.get();
  if (op == 0) o->Scale(a, b, c);
  else if (op == 1) o->Translate(a, b, c);
  else if (op == 2) o->Rotate({a, b, c}, d);
  else if (op == 3) o->ResetContext();
  else if (op == 4) static_cast<scene::Group*>(o)->Normalize();
}
void pth_scene_root_xf(void* sc, int op, float a, float b, float c, float d) {
  auto& r = static_cast<Box<scene::Scene>*>(sc)->p->root;
  if (op == 0) r.Scale(a, b, c);
  else if (op == 1) r.Translate(a, b, c);
  else if (op == 2) r.Rotate({a, b, c}, d);
}
void pth_model_matrix(void* obj, float out[16]) { std::memcpy(out, static_cast<Box<scene::Object>*>(obj)->p->ModelMatrix().m, 64); }

void* pth_camera_perspective(const float pos[3], const float tgt[3], const float up[3], float fov, float aspect, float n, float f) {
  return new Box<camera::Interface>{std::make_shared<camera::Perspective>(math::Vec3{pos[0], pos[1], pos[2]}, math::Vec3{tgt[0], tgt[1], tgt[2]},
                                                                           math::Vec3{up[0], up[1], up[2]}, fov, aspect, n, f)};
}
void* pth_camera_orthographic(const float pos[3], const float tgt[3], const float up[3], float l, float r, float b, float t, float n, float f) {
  return new Box<camera::Interface>{std::make_shared<camera::Orthographic>(math::Vec3{pos[0], pos[1], pos[2]}, math::Vec3{tgt[0], tgt[1], tgt[2]},
                                                                            math::Vec3{up[0], up[1], up[2]}, l, r, b, t, n, f)};
}
void pth_camera_matrices(void* cam, float view[16], float proj[16]) {
  auto& c = *static_cast<Box<camera::Interface>*>(cam)->p;
  std::memcpy(view, c.ViewMatrix().m, 64);
  std::memcpy(proj, c.ProjMatrix().m, 64);
}

void* pth_renderer_new(int w, int h, void* cam, void* sc, int shadow, int gamma, int msaa, int format, uint32_t background, const char* lib,
                       const char* prefix, int device, char* err, int errlen) {
  try {
    auto r = render::NewRenderer({render::Size(w, h), render::Camera(static_cast<Box<camera::Interface>*>(cam)->p),
                                  render::Scene(static_cast<Box<scene::Scene>*>(sc)->p), render::ShadowMap(shadow != 0), render::GammaCorrection(gamma != 0),
                                  render::MSAA(msaa), render::PixelFormat(format), render::Background(unpack(background)),
                                  std::string(prefix) == "prc_" ? render::CUDA(device, lib) : render::BackendLibrary(lib, prefix)});
    return r.release();
  } catch (const std::exception& e) {
    set_err(err, errlen, e);
    return nullptr;
  }
}
int pth_render(void* r, uint8_t* out, char* err, int errlen) {
  try {
    render::Frame f = static_cast<render::Renderer*>(r)->Render();
    std::memcpy(out, f.pix.data(), f.pix.size());
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e);
    return -1;
  }
}
int pth_render_view(void* r, uint8_t* out, char* err, int errlen) {  // zero-copy path (prc_host_image), copied out for the test
  try {
    render::FrameView v = static_cast<render::Renderer*>(r)->RenderView();
    std::memcpy(out, v.pix, (size_t)v.w * v.h * 4);
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e);
    return -1;
  }
}
int pth_set_camera(void* r, void* cam, char* err, int errlen) {  // Options(Camera(c)): re-fits the light cameras, zeroes the shadow maps
  try {
    static_cast<render::Renderer*>(r)->Options({render::Camera(static_cast<Box<camera::Interface>*>(cam)->p)});
    return 0;
  } catch (const std::exception& e) {
    set_err(err, errlen, e);
    return -1;
  }
}
// render.Debug(true) / the shadow-map picture it saves (render/shadow.go:98-118); w*h*4 bytes
int pth_shadow_map_image(void* r, int index, uint8_t* out, char* err, int errlen) {
  try {
    render::Frame f = static_cast<render::Renderer*>(r)->ShadowMapImage(index);
    std::memcpy(out, f.pix.data(), f.pix.size());
    return 0;
  } catch (const std::exception& e) { set_err(err, errlen, e); return -1; }
}
int pth_set_debug(void* r, int enable, char* err, int errlen) {
  try {
    static_cast<render::Renderer*>(r)->Options({render::Debug(enable != 0), render::Workers(4), render::BatchSize(64)});
    return 0;
  } catch (const std::exception& e) { set_err(err, errlen, e); return -1; }
}
int pth_set_blending(void* r, char* err, int errlen) {  // must throw: Blending is rejected, not ignored
  try {
    static_cast<render::Renderer*>(r)->Options({render::Blending([](RGBA, RGBA b) { return b; })});
    return 0;
  } catch (const std::exception& e) { set_err(err, errlen, e); return -1; }
}
const prc_frame* pth_last_frame(void* r) { return &static_cast<render::Renderer*>(r)->LastFrame(); }
void pth_renderer_free(void* r) { delete static_cast<render::Renderer*>(r); }

// small math probes (bit-exactness of the C++ restatements against the other mirrors)
void pth_mat4_inv(const float a[16], float out[16]) { math::Mat4 m; std::memcpy(m.m, a, 64); std::memcpy(out, m.Inv().m, 64); }
void pth_mat4_mulm(const float a[16], const float b[16], float out[16]) { math::Mat4 m, n; std::memcpy(m.m, a, 64); std::memcpy(n.m, b, 64); std::memcpy(out, m.MulM(n).m, 64); }
float pth_mat4_det(const float a[16]) { math::Mat4 m; std::memcpy(m.m, a, 64); return m.Det(); }
void pth_gamma_lut(uint8_t out[256]) { for (int i = 0; i < 256; i++) out[i] = (uint8_t)(long long)(color::FromLinear2sRGB((float)i / 255.0f) * 255.0f + 0.5f); }
int pth_resize(const uint8_t* img, int iw, int ih, uint8_t* out, int ow, int oh) {
  imageutil::Image im; im.w = iw; im.h = ih; im.pix.assign(img, img + (size_t)iw * ih * 4);
  imageutil::Image o = imageutil::Resize(ow, oh, im);
  std::memcpy(out, o.pix.data(), o.pix.size());
  return 0;
}
}  // extern "C"
