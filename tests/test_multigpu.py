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
