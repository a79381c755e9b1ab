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


def test_batch_take_preserves_order_and_blocks_like_take():
    import threading
    import time
    from asyncframework_b200.core import AsyncContext
    from asyncframework_b200.core.context import QueueClosed
    from asyncframework_b200.core.partial_result import PartialResult
    ctx = AsyncContext(4)
    for i in range(10):
        ctx.put(PartialResult({"i": i}, staleness=i, timestamp=i, batch_size=1, worker_id=i % 4))
    first = ctx.drain(4)
    assert [r.data["i"] for r in first] == [0, 1, 2, 3] and [r.worker_id for r in first] == [0, 1, 2, 3]
    assert [r.data["i"] for r in ctx.drain(100)] == [4, 5, 6, 7, 8, 9] and ctx.size() == 0
    t0 = time.perf_counter()
    assert ctx.drain(8, timeout=0.05) == [] and time.perf_counter() - t0 >= 0.04          # timeout, nothing queued
    got = []
    th = threading.Thread(target=lambda: got.extend(ctx.drain(8)))                           # blocks until the first result
    th.start()
    time.sleep(0.05)
    ctx.put(PartialResult("late", 0, 0, 1, 2))
    th.join(timeout=5)
    assert [r.data for r in got] == ["late"]
    ctx.close()
    with pytest.raises(QueueClosed):
        ctx.drain(1)
    assert not ctx._payloads                                                                    # no payload leaked
