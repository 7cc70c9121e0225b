// prc_pow.h — math.Pow for the specular term, integer exponents (host + device).
//
// shader/blinn_cpu.go:88 raises clamp(n.h, 0, 1) to the material's shininess through Go's math.Pow
// (math/pow.go): Frexp the base, then square-and-multiply over the bits of the integer part of the
// exponent on a mantissa in [.5, 1) with the binary exponent carried separately, Ldexp at the end. Every
// rescaling in that loop is a multiplication by a power of two, hence exact, so for a base in (0, 1] and a
// positive integer exponent the sequence of ROUNDED operations is exactly: double square-and-multiply on the
// unscaled value. The two can only part when a used intermediate leaves the normal double range, i.e. is
// below 2^-1022; every used intermediate is >= the final result (base <= 1), so the final result is then
// below 2^-1022 too and both convert to float32 zero. tests/native/pow_check.cpp checks this against the
// oracle's literal restatement on 10^8 bases, including subnormal floats and every exponent up to 4096.
#pragma once

#if defined(__CUDACC__)
#define PRC_POW_HD __host__ __device__ __forceinline__
#else
#define PRC_POW_HD inline
#endif

// true when pow_int_unit() applies
PRC_POW_HD bool pow_int_unit_ok(float x, float y) { return x > 0.0f && x <= 1.0f && y >= 1.0f && y <= 1048576.0f && y == (float)(unsigned int)y; }

PRC_POW_HD float pow_int_unit(float x, float y) {
  double x1 = (double)x, a1 = 1.0;
  for (unsigned int i = (unsigned int)y;;) {
    if (i & 1u) a1 = a1 * x1;  // a double product is never contracted
    i >>= 1;
    if (!i) break;
    x1 = x1 * x1;
  }
  return (float)a1;
}
