"""CPU: the eight AO ray directions (material/ao.go:28-30) do not depend on whose cos / sin computes them.

The reference evaluates float32(math.Cos(float64(a))) with Go's pure-Go Cephes routines; the oracle uses libm and the CUDA
library std::cos on the host (csrc/polyred_cuda.cu prc_open) — no fixture pins Go's routines. For these eight fixed
arguments that does not matter: the exact values lie so far from every float32 rounding boundary that ANY routine with a relative
error below 2^-40 (Go, glibc and CUDA are all below 2^-52) rounds to the same float32. Checked here with 60-digit arithmetic.
(The per-step Atan has arbitrary arguments and stays unpinned: two faithful double routines can differ in the last bit, which
survives the rounding to float32 with probability ~2^-29 per call — DESIGN.md section 2.)"""
import numpy as np


def _angles():
    out, a, step = [], np.float32(0.0), np.float32(np.pi / 4)  # math.Pi / 4 as a float32 constant
    while a < np.float32(2 * np.pi) - np.float32(1e-4):       # for a := float32(0); a < TwoPi-1e-4; a += Pi/4
        out.append(a)
        a = np.float32(a + step)
    return out


def _margin_to_float32_boundary(x):
    """Relative distance of the exact value x (mpmath) to the nearest midpoint between two adjacent float32 numbers."""
    import mpmath as mp
    f = np.float32(float(x))
    lo, hi = np.nextafter(f, np.float32(-np.inf)), np.nextafter(f, np.float32(np.inf))
    mids = [(mp.mpf(float(f)) + mp.mpf(float(lo))) / 2, (mp.mpf(float(f)) + mp.mpf(float(hi))) / 2]
    return min(abs(x - m) for m in mids) / abs(x)


def test_ao_ray_directions_are_the_same_float32_for_any_faithful_cos_sin():
    import mpmath as mp
    mp.mp.dps = 60
    angles = _angles()
    assert len(angles) == 8
    for a in angles:
        for fn, npfn in ((mp.cos, np.cos), (mp.sin, np.sin)):
            exact = fn(mp.mpf(float(a)))
            if exact == 0:
                continue  # sin(0)
            got = np.float32(npfn(np.float64(a)))  # what the oracle / the library's host code computes
            assert got == np.float32(float(exact)), (a, fn)
            assert _margin_to_float32_boundary(exact) > mp.mpf(2) ** -40, (float(a), fn.__name__)
