"""Fused NVLink paths at 2+ GPUs: spawns torchrun on tests/mg_worker.py and checks its report."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]


@pytest.mark.timeout(600)
def test_fused_paths_on_all_visible_gpus():
    n = min(torch.cuda.device_count(), 8)
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "mg_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=580)
    line = [l for l in res.stdout.splitlines() if l.startswith("MG_RESULT ")]
    assert line, res.stdout[-2000:] + res.stderr[-4000:]
    out = json.loads(line[0][len("MG_RESULT "):])
    assert out["world"] == n and out["peer_read_ok"]
    assert out["sgd_sync_relerr"] < 2e-3 and out["saga_sync_relerr"] < 2e-3
    assert out["sgd_sync_replicas_identical"] and out["saga_sync_replicas_identical"]
    for key, iters in (("asgd_tawinf", 60 * n), ("asaga_tawinf", 60 * n), ("asgd_taw0", 15 * n)):
        assert out[key + "_k"] >= iters and out[key + "_finite"]
        assert out[key + "_avail"] == n                       # every worker idle again
        assert out[key + "_loss_last"] < out[key + "_loss_first"]
        assert out[key + "_workers_seen"] == list(range(n))
    assert out["asgd_taw0_stale_applied"] == 0                # no update applied with staleness > taw
    assert out["asgd_tawinf_accepted"] >= 60 * n
    assert out["versioned_broadcast_ok"]
    # BSP rounds are persistent (20 rounds = 1 launch) and use the NVLS multicast all-gather whenever the fabric offers it
    assert out["sgd_sync_variant"] == "stream" and out["sgd_sync_launches"] == 1
    assert out["sgd_sync_nvls"] == out["multicast"]
    # persistent worker loops: one launch, R rounds per worker, every round accepted exactly once
    R = 40
    for name, variant in (("pipelined", "stream-pipelined"), ("bucket", "stream"), ("ssp", "stream"), ("asaga_persist", "stream")):
        key = "p_" + name
        assert out[key + "_variant"] == variant and out[key + "_launches"] == 1
        assert out[key + "_k"] == R * n and out[key + "_clock"] == R * n and out[key + "_ring"] == R * n
        assert out[key + "_avail"] == n and out[key + "_finite"]
        assert out[key + "_workers_seen"] == list(range(n))
        assert out[key + "_loss_last"] < out[key + "_loss_first"]
    # ASYNCbarrier on the fused path: the masked worker sat every round out, the others ran theirs
    assert out["p_masked_k"] == R * (n - 1) and out["p_masked_workers_seen"] == list(range(n - 1))
    assert out["p_masked_skipped"] == [0] * (n - 1) + [R] and out["p_masked_avail"] == n


def _spawn_pair(mode, port):
    here = os.path.dirname(os.path.abspath(__file__))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   TORCH_NCCL_ASYNC_ERROR_HANDLING="0", NCCL_ASYNC_ERROR_HANDLING="0")
        procs.append(subprocess.Popen([sys.executable, os.path.join(here, "mg_elastic.py"), mode], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    try:
        out0, err0 = procs[0].communicate(timeout=150)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    rc1 = procs[1].wait(timeout=30)
    line = [l for l in out0.splitlines() if l.startswith("ELASTIC_RESULT ")]
    assert line, out0[-1500:] + err0[-3000:]
    return json.loads(line[0][len("ELASTIC_RESULT "):]), rc1


@pytest.mark.timeout(200)
def test_async_run_continues_with_p_minus_1_when_a_rank_dies():
    """SIGKILL one rank mid-run: the server's watchdog declares its worker dead on the device (alive = 0, gate threshold
    lowered) and the asynchronous run still reaches the target number of accepted updates with the surviving worker."""
    out, rc1 = _spawn_pair("async", 29571)
    assert rc1 == -9, out                                    # the rank really died by SIGKILL
    assert out["raised"] is None and out["k"] >= 200_000 and out["finite"], out
    assert out["dead"] == [1] and out["alive_flags"] == [1, 0] and out["num_alive"] == 1 and out["gate_dyn"] == 1


@pytest.mark.timeout(200)
def test_sync_run_aborts_cleanly_when_a_rank_dies():
    """BSP: a dead peer can never deliver its gradient slot — the watchdog stops the local rounds, the in-kernel flag wait
    observes the stop word, the kernel exits with the abort flag and the driver raises (no hang)."""
    out, rc1 = _spawn_pair("sync", 29572)
    assert rc1 == -9, out
    assert out["raised"] is not None and out["abort"] == 1 and out["dead"] == [1] and out["elapsed_s"] < 30, out
