"""CPU: brute-force validation of the exact-safe pixel-loop pruning (csrc/prc_prune.h) against the
reference's AABB+-1 loop on random and adversarial triangles (tests/native/prune_check.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prune_never_drops_a_pixel_the_reference_accepts(tmp_path):
    exe = tmp_path / "prune_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", str(exe), os.path.join(ROOT, "tests", "native", "prune_check.cpp")])
    out = subprocess.check_output([str(exe), "6000000", "11"], text=True)
    kv = dict(x.split("=") for x in out.split())
    assert int(kv["violations"]) == 0, out
    assert int(kv["pruned_ok"]) > 1_000_000 and int(kv["pruned_tests"]) < int(kv["full_tests"]), out


def test_integer_pow_shortcut_matches_go_pow(tmp_path):
    """csrc/prc_pow.h (plain double square-and-multiply for base in (0,1], integer exponent) against the oracle's literal
    restatement of math.Pow (Frexp / separate exponent / Ldexp), tests/native/pow_check.cpp."""
    exe = tmp_path / "pow_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", str(exe), os.path.join(ROOT, "tests", "native", "pow_check.cpp")])
    out = subprocess.check_output([str(exe), "3000000", "5"], text=True)
    kv = dict(x.split("=") for x in out.split())
    assert int(kv["mismatches"]) == 0 and int(kv["checked"]) > 10_000_000, out
