"""C++ host runtime under a multi-threaded stress test, plain and with ThreadSanitizer."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "csrc", "runtime", f) for f in ("test_host_runtime.cpp", "host_runtime.cpp")]


def _build_and_run(tmp_path, extra, env=None):
    exe = str(tmp_path / "t")
    err = ""
    # the sanitizer runtimes ship with the distro compiler, not necessarily with the first g++ on PATH
    for cxx in [os.environ.get("CXX"), "/usr/bin/g++", shutil.which("g++")]:
        if not cxx or not os.path.exists(cxx):
            continue
        r = subprocess.run([cxx, "-O1", "-g", "-std=c++17", "-pthread", *extra, *SRC, "-o", exe], capture_output=True, text=True)
        if r.returncode == 0:
            break
        err = r.stderr
    else:
        pytest.skip("toolchain cannot build this configuration: " + err[-300:])
    run = subprocess.run([exe], capture_output=True, text=True, timeout=300, env={**os.environ, **(env or {})})
    return run


def test_host_runtime_stress(tmp_path):
    run = _build_and_run(tmp_path, [])
    assert run.returncode == 0 and "OK" in run.stdout, run.stdout + run.stderr


def test_host_runtime_stress_tsan(tmp_path):
    run = _build_and_run(tmp_path, ["-fsanitize=thread"], env={"TSAN_OPTIONS": "halt_on_error=1 exitcode=66"})
    if "FATAL: ThreadSanitizer" in run.stderr and "unexpected memory mapping" in run.stderr:
        pytest.skip("TSAN cannot map its shadow memory in this container")
    assert "WARNING: ThreadSanitizer" not in run.stderr, run.stderr[-3000:]
    assert run.returncode == 0 and "OK" in run.stdout, run.stdout + run.stderr[-2000:]
