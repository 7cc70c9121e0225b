"""CPU: the device-group host logic (polyred_b200/csrc/prc_group.cpp: submit threads, strip partition and re-balancing, retry
vote, error handling, reconnects, view batches) against stubbed contexts — tests/native/group_host_check.cpp includes the group
source itself and defines the library calls it is built on as recording stubs. Run plain and under ThreadSanitizer (the worker
hand-off is a lock-free spin protocol; the first run found a worker reading g->workers while prc_group_open was still growing it)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "group_host_check.cpp")


def _build_and_run(tmp_path, name, extra):
    exe = tmp_path / name
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-pthread", "-Wno-subobject-linkage", *extra, "-o", str(exe), SRC], cwd=os.path.dirname(SRC))
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    return p


def test_group_host_logic_against_stub_contexts(tmp_path):
    p = _build_and_run(tmp_path, "group_host_check", [])
    assert p.returncode == 0 and "OK all checks passed" in p.stdout, p.stdout[-4000:] + p.stderr[-2000:]


def test_group_submit_threads_are_race_free_under_thread_sanitizer(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    try:
        p = _build_and_run(tmp_path, "group_host_check_tsan", ["-fsanitize=thread"])
    except subprocess.CalledProcessError:
        pytest.skip("ThreadSanitizer runtime not available")
    if "FATAL: ThreadSanitizer" in p.stderr and "unexpected memory mapping" in p.stderr:
        pytest.skip("ThreadSanitizer cannot run in this container (address-space layout)")
    assert "WARNING: ThreadSanitizer" not in p.stderr, p.stderr[-6000:]
    assert p.returncode == 0 and "OK all checks passed" in p.stdout, p.stdout[-4000:] + p.stderr[-2000:]
