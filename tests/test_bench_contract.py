"""bench.py contract (keys, types, arms) exercised on the CPU against the round simulator."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_reports_unavailable():
    d = _run("--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3")
    assert d["impl"] == "reference" and "unavailable" in d and "\n" not in d["unavailable"]


def test_bench_json_contract_on_simulator():
    d = _run("--device", "cpu", "--rows", "20000", "--steps", "5", "--warmup", "3", "--workers-per-gpu", "2")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["dtype"] == "bf16" and d["device"] == "cpu"
    assert abs(d["value"] - 5 / (d["ms_per_step"] * 5e-3)) < 1e-6 * d["value"]
    for k in ("model", "global_batch", "seq_len", "parallelism"):
        assert k in d["config"], k
    assert d["config"]["global_batch"] == 2000 and d["config"]["workers_per_gpu"] == 2
    assert d["gpu_launches"] == 2                                      # ONE persistent launch per worker for the 5 timed rounds
    assert d["config"]["dense_variant"] == "stream"                    # two workers share the GPU: serial persistent rounds (pipelined rounds own every SM)
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "iterations/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.0 < e["last_loss_per_row"] < 1.0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["config"]["accepted_updates_total"] == (3 + 5 + 3 + 5) * 2   # warm-up + timed + e2e warm-up + e2e, 2 workers


def test_time_to_loss_benchmark_runs_on_the_simulator():
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "benchmarks", "time_to_loss.py"), "--impl", "ours", "--target", "0.6",
                          "--eval-every", "5", "--max-iters", "60", "--gamma", "1e-4"], capture_output=True, text=True, timeout=240,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), cwd=root)
    assert res.returncode == 0, res.stderr[-1500:]
    line = [l for l in res.stdout.splitlines() if l.startswith("TIME_TO_LOSS ")][-1]
    r = json.loads(line[len("TIME_TO_LOSS "):])
    assert r["impl"] == "ours" and r["reached"] and r["ms_to_target"] > 0 and r["curve"][0][1] > r["curve"][-1][1]
