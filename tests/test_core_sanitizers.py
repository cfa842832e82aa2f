"""Race / memory-error detection for the native control-plane core: the concurrency stress in
``core/csrc/stress_test.cpp`` is built with ThreadSanitizer and with AddressSanitizer + UBSan and must run clean.
The reference has no ``-race`` build at all; this is the C++ analogue."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "trainingjob_operator_b200", "core", "csrc")


@pytest.mark.parametrize("name,flags,env", [
    ("tsan", ["-fsanitize=thread"], {"TSAN_OPTIONS": "halt_on_error=1 second_deadlock_stack=1"}),
    ("asan_ubsan", ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
     {"ASAN_OPTIONS": "detect_leaks=1:abort_on_error=0", "UBSAN_OPTIONS": "halt_on_error=1:print_stacktrace=1"}),
])
def test_core_stress_runs_clean_under_sanitizers(tmp_path, name, flags, env):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / f"core_stress_{name}")
    build = subprocess.run([cxx, "-std=c++17", "-O1", "-g", *flags, "-pthread", os.path.join(CSRC, "stress_test.cpp"),
                            "-o", exe], capture_output=True, text=True, cwd=CSRC)
    if build.returncode != 0 and "sanitizer" in (build.stderr + build.stdout).lower() and "cannot find" in build.stderr:
        pytest.skip(f"sanitizer runtime not installed: {build.stderr[-200:]}")
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env={**os.environ, **env})
    out = run.stdout + run.stderr
    assert run.returncode == 0, out[-4000:]
    assert "core stress ok" in out
    assert "WARNING: ThreadSanitizer" not in out and "ERROR: AddressSanitizer" not in out and "runtime error" not in out, \
        out[-4000:]
