"""The daemons as a user starts them: ``python -m asyncframework_b200.deploy.master`` / ``.worker`` as separate OS processes,
an application submitted with ``python -m asyncframework_b200.submit --master spark://...`` (sbin/start-all.sh does the same)."""
import json
import os
import signal
import socket
import subprocess
import sys
import time
import urllib.request

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _wait_for(pred, timeout=40.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        try:
            if pred():
                return True
        except Exception:  # noqa: BLE001
            pass
        time.sleep(0.2)
    return False


def test_master_and_worker_daemons_run_a_submitted_reference_command_line(tmp_path):
    port, ui = _free_port(), _free_port()
    env = {**os.environ, "PYTHONPATH": ROOT, "CUDA_VISIBLE_DEVICES": ""}
    procs = []
    try:
        procs.append(subprocess.Popen([sys.executable, "-m", "asyncframework_b200.deploy.master", "--port", str(port),
                                       "--webui-port", str(ui), "--recovery-dir", str(tmp_path / "rec")],
                                      env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        state = lambda: json.load(urllib.request.urlopen(f"http://127.0.0.1:{ui}/json", timeout=2))  # noqa: E731
        assert _wait_for(lambda: state()["status"] == "ALIVE"), "master did not come up"
        for i in range(2):
            procs.append(subprocess.Popen([sys.executable, "-m", "asyncframework_b200.deploy.worker", f"spark://127.0.0.1:{port}",
                                           "--cores", "1", "--memory", "1024", "--work-dir", str(tmp_path / f"w{i}")],
                                          env=env, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        assert _wait_for(lambda: len(state()["workers"]) == 2), "workers did not register"
        # README.md:46-shaped submission (sync driver, 11 positional args) on synthetic data
        cmd = [sys.executable, "-m", "asyncframework_b200.submit", "--master", f"spark://127.0.0.1:{port}", "--name", "cli-it",
               "--class", "org.apache.async.SparkASGDSync", "--executor-memory", "512m", "async.jar",
               "x", "x", "6", "300", "2", "6", "0.5", "0.5", "3", "0", "42", "--backend=cpu", "--synthetic=linear", "--quiet=true"]
        out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=240)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
        assert "Elapsed time" in out.stdout
        assert _wait_for(lambda: [a["name"] for a in state()["completed_apps"]] == ["cli-it"] and not state()["active_apps"])
        done = state()["completed_apps"][0]
        assert done["state"] == "FINISHED" and len(done["executors"]) == 2
        assert all(w["cores_used"] == 0 for w in state()["workers"])
        # executor logs land in the worker's work dir: work/<app>/<exec>/{stdout,stderr}
        logs = [os.path.join(dp, f) for i in range(2) for dp, _, fs in os.walk(tmp_path / f"w{i}") for f in fs]
        assert sum(1 for l in logs if l.endswith("stderr")) == 2
        assert os.path.exists(tmp_path / "rec" / "master-state.json")
    finally:
        for p in reversed(procs):
            p.send_signal(signal.SIGINT)
        for p in procs:
            try:
                p.wait(10)
            except subprocess.TimeoutExpired:
                p.kill()
