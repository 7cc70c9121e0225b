// Brute-force check of csrc/prc_pow.h against the oracle's literal restatement of Go's math.Pow.
// usage: pow_check <n_random> <seed>   prints "checked=... mismatches=..."
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include "../../oracle/pr_math.h"
#include "../../polyred_b200/csrc/prc_pow.h"

static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float from_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 1000000;
  std::mt19937_64 rng(argc > 2 ? atol(argv[2]) : 1);
  long checked = 0, bad = 0;
  auto check = [&](float x, float y) {
    if (!pow_int_unit_ok(x, y)) return;
    const float a = orc::Pow(x, y), b = pow_int_unit(x, y);
    checked++;
    if (bits(a) != bits(b)) { if (bad < 10) fprintf(stderr, "x=%a y=%g ref=%a got=%a\n", x, y, a, b); bad++; }
  };
  // every exponent up to 4096 against structured bases (powers of two +- 1 ulp, subnormals, near 1)
  for (int yi = 1; yi <= 4096; yi++)
    for (int e = 0; e < 150; e++)
      for (int d = -2; d <= 2; d++) {
        check(from_bits(bits(ldexpf(1.0f, -e)) + d), (float)yi);
        check(from_bits((uint32_t)(1 + (e * 7919u + (uint32_t)d + 2u) % 8388607u)), (float)yi);  // subnormal floats
      }
  // random bases over the whole (0,1] range (uniform in bits = log-uniform in value, and uniform in value)
  const float ys[] = {1, 2, 3, 4, 5, 7, 8, 10, 16, 25, 32, 50, 64, 100, 127, 128, 255, 256, 1000, 1024, 65536, 1048576};
  for (long i = 0; i < n; i++) {
    const float xb = from_bits((uint32_t)(rng() % 0x3F800001ull));
    const float xu = (float)((rng() >> 11) * (1.0 / 9007199254740992.0));
    const float y = (i & 1) ? ys[(rng() >> 8) % (sizeof(ys) / sizeof(ys[0]))] : (float)(1 + (rng() >> 8) % 300);
    check(xb, y);
    check(xu, y);
    check(1.0f - xu * 0.01f, y);
  }
  printf("checked=%ld mismatches=%ld\n", checked, bad);
  return bad != 0;
}
