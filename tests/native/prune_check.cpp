// Brute-force validation of csrc/prc_prune.h against the reference pixel loop (oracle restatement of
// math.Barycoord + the inside test of render/raster.go:487-493). TEST INFRASTRUCTURE.
// usage: prune_check <triangles> <seed>   -> prints "violations=<n> checked=<n> pruned_ok=<n> accepted=<n> skipped_tests=<n>"
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../../oracle/pr_math.h"
#include "../../polyred_b200/csrc/prc_prune.h"

using namespace orc;

struct Stats { unsigned long long viol = 0, checked = 0, ok = 0, accepted = 0, full = 0, pruned = 0; };

static void check(const float v[6], Stats& st) {
  Vec2 a{v[0], v[1]}, b{v[2], v[3]}, c{v[4], v[5]};
  float mnx = Min3(a.x, b.x, c.x), mxx = Max3(a.x, b.x, c.x), mny = Min3(a.y, b.y, c.y), mxy = Max3(a.y, b.y, c.y);
  long long x0 = go_int(Round(mnx) - 1), x1 = go_int(Round(mxx) + 1), y0 = go_int(Round(mny) - 1), y1 = go_int(Round(mxy) + 1);
  if (x1 - x0 > 200 || y1 - y0 > 200) return;
  float Sabc = cross2z(b.x - a.x, b.y - a.y, c.x - a.x, c.y - a.y);
  st.checked++;
  bool ok = prune_ok(mnx, mny, mxx, mxy, Sabc);
  if (!ok) return;
  st.ok++;
  long long px0 = prune_first(mnx), px1 = prune_last(mxx), py0 = prune_first(mny), py1 = prune_last(mxy);
  st.full += (unsigned long long)((x1 - x0 + 1) * (y1 - y0 + 1));
  if (px1 >= px0 && py1 >= py0) st.pruned += (unsigned long long)((px1 - px0 + 1) * (py1 - py0 + 1));
  for (long long x = x0; x <= x1; x++)
    for (long long y = y0; y <= y1; y++) {
      float w[3];
      barycoord(Vec2{(float)x + 0.5f, (float)y + 0.5f}, a, b, c, w);
      if (w[0] < -Epsilon || w[1] < -Epsilon || w[2] < -Epsilon) continue;
      st.accepted++;
      if (x < px0 || x > px1 || y < py0 || y > py1) st.viol++;
    }
}

int main(int argc, char** argv) {
  unsigned long long n = argc > 1 ? strtoull(argv[1], 0, 10) : 1000000ULL;
  unsigned seed = argc > 2 ? atoi(argv[2]) : 1;
  int nth = std::thread::hardware_concurrency();
  if (nth < 1) nth = 1;
  std::vector<Stats> stats(nth);
  std::vector<std::thread> th;
  for (int t = 0; t < nth; t++)
    th.emplace_back([&, t]() {
      std::mt19937_64 rng(seed * 7919u + t);
      std::uniform_real_distribution<double> U(0, 1);
      Stats& st = stats[t];
      for (unsigned long long i = t; i < n; i += nth) {
        float v[6];
        int mode = (int)(rng() % 8);
        double base = (rng() % 4 == 0) ? 4000.0 * U(rng) : 64.0 * U(rng);
        double basey = (rng() % 4 == 0) ? 2000.0 * U(rng) : 64.0 * U(rng);
        double ext = std::pow(10.0, -3.0 + 4.8 * U(rng));  // 1e-3 .. 63
        auto rnd = [&]() { return (U(rng) - 0.5) * ext; };
        switch (mode) {
          case 0: case 1: case 2:  // general position
            for (int k = 0; k < 3; k++) { v[2 * k] = (float)(base + rnd()); v[2 * k + 1] = (float)(basey + rnd()); }
            break;
          case 3: {  // sliver: third vertex near the segment
            double x0 = base + rnd(), y0 = basey + rnd(), x1 = base + rnd(), y1 = basey + rnd(), s = U(rng);
            double off = std::pow(10.0, -8.0 + 7.0 * U(rng)) * (U(rng) < 0.5 ? -1 : 1);
            double nx = -(y1 - y0), ny = x1 - x0, nl = std::sqrt(nx * nx + ny * ny) + 1e-30;
            v[0] = (float)x0; v[1] = (float)y0; v[2] = (float)x1; v[3] = (float)y1;
            v[4] = (float)(x0 + s * (x1 - x0) + off * nx / nl); v[5] = (float)(y0 + s * (y1 - y0) + off * ny / nl);
            break; }
          case 4: {  // snapped to a binary grid (vertices on pixel centres / edges)
            double q = 1.0 / (1 << (rng() % 5));
            for (int k = 0; k < 3; k++) { v[2 * k] = (float)(std::floor((base + rnd()) / q) * q); v[2 * k + 1] = (float)(std::floor((basey + rnd()) / q) * q); }
            break; }
          case 5: {  // axis-aligned right triangle
            double x0 = base + rnd(), y0 = basey + rnd(), w = rnd(), h = rnd();
            v[0] = (float)x0; v[1] = (float)y0; v[2] = (float)(x0 + w); v[3] = (float)y0; v[4] = (float)x0; v[5] = (float)(y0 + h);
            break; }
          case 6: {  // one edge just touching a pixel centre
            double cx = std::floor(base) + 0.5, cy = std::floor(basey) + 0.5, e = std::pow(10.0, -9.0 + 8.0 * U(rng)) * (U(rng) < 0.5 ? -1 : 1);
            v[0] = (float)(cx + e); v[1] = (float)(cy + rnd()); v[2] = (float)(cx + e + rnd() * 1e-3); v[3] = (float)(cy + rnd()); v[4] = (float)(cx + rnd()); v[5] = (float)(cy + rnd());
            break; }
          default: {  // degenerate / repeated
            for (int k = 0; k < 3; k++) { v[2 * k] = (float)(base + rnd()); v[2 * k + 1] = (float)(basey + rnd()); }
            if (rng() & 1) { v[4] = v[0]; v[5] = v[1]; } else { v[4] = v[0] + (v[2] - v[0]) * 0.5f; v[5] = v[1] + (v[3] - v[1]) * 0.5f; }
            break; }
        }
        if (rng() & 1) { std::swap(v[2], v[4]); std::swap(v[3], v[5]); }  // both windings
        check(v, st);
      }
    });
  for (auto& t : th) t.join();
  Stats s;
  for (auto& x : stats) { s.viol += x.viol; s.checked += x.checked; s.ok += x.ok; s.accepted += x.accepted; s.full += x.full; s.pruned += x.pruned; }
  printf("violations=%llu checked=%llu pruned_ok=%llu accepted=%llu full_tests=%llu pruned_tests=%llu\n", s.viol, s.checked, s.ok, s.accepted, s.full, s.pruned);
  return s.viol ? 1 : 0;
}
