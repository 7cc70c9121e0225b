// prc_prune.h — exact-safe pruning of the reference's AABB+-1 pixel loop (host + device).
//
// drawClipped / drawDepth visit every pixel of int(Round(min)-1) .. int(Round(max)+1)
// (render/raster.go:473-493, render/shadow.go:186-203) and keep those whose three divided
// barycentrics are >= -1e-7. For a micro-triangle that is >= 9 Barycoord evaluations of which
// usually 0-2 can pass. A pixel centre lying outside the triangle's float bounding box by a
// margin M cannot pass unless (a) the barycentric tolerance reaches that far (|w| <= 1e-7 means a
// distance of ~1e-7 x triangle size) or (b) float cancellation in the cross products makes a
// clearly negative barycentric non-negative (only for slivers: error ~ 2^-23 |ap| L / Sabc).
// prune_ok() admits only triangles where both effects are orders of magnitude below M:
//     L = max bbox extent + 2 <= 64 px   and   |Sabc| >= 1.25e-4 L^3  (i.e. height >= 1.25e-4 L^2)
// Triangles failing the test (slivers, big ones, NaN/Inf) take the reference's full loop.
// The predicate is validated by brute force against the reference loop on 10^8+ random and
// adversarial triangles in tests/native/prune_check.cpp (run by tests/test_prune_native.py) and
// by the bit-exact G-buffer / shadow-map parity tests.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PRC_HD __host__ __device__ __forceinline__
#else
#define PRC_HD inline
#endif

#define PRC_PRUNE_MARGIN 0.03125f  // 1/32 px
#ifndef PRC_PRUNE_K
#define PRC_PRUNE_K 1.25e-4f
#endif

// mn/mx: float bbox of the three screen vertices; Sabc: the reference's cross product (2 x signed area)
PRC_HD bool prune_ok(float mnx, float mny, float mxx, float mxy, float Sabc) {
  float ex = mxx - mnx, ey = mxy - mny;
  float L = (ex > ey ? ex : ey) + 2.0f;
  if (!(L <= 64.0f)) return false;  // also false for NaN
  float a = fabsf(Sabc);
  return a >= PRC_PRUNE_K * L * L * L && a < 3.0e38f;
}
// first / last pixel index (inclusive) whose centre i+0.5 lies in [lo - M, hi + M]
PRC_HD int prune_first(float lo) { return (int)ceilf(lo - 0.5f - PRC_PRUNE_MARGIN); }
PRC_HD int prune_last(float hi) { return (int)floorf(hi - 0.5f + PRC_PRUNE_MARGIN); }
