// example.cpp — the call a C++ user makes: the reference's options plus one backend-selection option.
//   ./example [path/to/libpolyred_cuda.so] > frame.ppm
#include <cstdio>

#include "polyred_host.hpp"

using namespace polyred;

int main(int argc, char** argv) {
  const std::string lib = argc > 1 ? argv[1] : "../polyred_b200/csrc/libpolyred_cuda.so";
  auto s = std::make_shared<scene::Scene>();
  s->Add(light::NewPoint(5, {255, 255, 255, 255}, {-2, 2.5f, 6}));
  s->Add(light::NewAmbient(0.5f));
  auto g = std::make_shared<scene::Geometry>();  // one triangle with the default material
  g->pos = {-0.5f, -0.5f, 0, 0.5f, -0.5f, 0, 0, 0.5f, 0};
  g->nor = {0, 0, 1, 0, 0, 1, 0, 0, 1};
  g->uv = {0, 0, 1, 0, 0.5f, 1};
  g->materials = {material::Default()};
  g->RotateY(0.3f);
  s->Add(g);
  auto cam = std::make_shared<camera::Perspective>(math::Vec3{0, 0, 2}, math::Vec3{0, 0, 0}, math::Vec3{0, 1, 0}, 45.0f, 16.0f / 9.0f, 0.1f, 10.0f);
  try {
    auto r = render::NewRenderer({render::Camera(cam), render::Size(640, 360), render::Scene(s), render::ShadowMap(false), render::GammaCorrection(true),
                                  render::MSAA(2), render::CUDA(0, lib)});
    render::Frame f = r->Render();
    std::printf("P6\n%d %d\n255\n", f.w, f.h);
    for (size_t i = 0; i < f.pix.size(); i += 4) std::fwrite(&f.pix[i], 1, 3, stdout);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());  // errors surface; there is no CPU fallback
    return 1;
  }
  return 0;
}
