"""Kill-a-rank worker (spawned twice by tests/test_multigpu.py::test_*_rank_dies_*, NOT under torchrun: torchrun tears the
whole group down when one child dies).  Rank 1 kills itself with SIGKILL in the middle of the run; rank 0 is the parameter
server and must

  * async mode (ASGD behind a beta = 1.0 wave gate, which needs EVERY live worker idle to open): notice the silent worker
    through the device heartbeats, declare it dead (alive = 0, idle count and gate threshold corrected on the server block),
    continue with P-1 and reach the target number of accepted updates           (HeartbeatReceiver.scala:193-220)
  * sync mode (BSP all-gather): stop its own rounds cleanly (the in-kernel flag wait observes the stop word) and raise.

Prints one line ``ELASTIC_RESULT {json}`` from rank 0.  No collective is issued after the death.
"""
import json
import os
import signal
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200.config import AsyncConfig  # noqa: E402
from asyncframework_b200.parallel.comm import init_distributed  # noqa: E402
from asyncframework_b200.parallel.fused_backend import FusedEngine  # noqa: E402
from asyncframework_b200.utils.watchdog import DeviceWatchdog  # noqa: E402
from tests.mg_worker import make_parts, shard_of  # noqa: E402


def main():
    mode = sys.argv[1]
    rank, local_rank, world = init_distributed()
    dev = torch.device("cuda", local_rank)
    d, n_per = 64, 4000
    target = 200_000            # ~1-2 s of rounds: the kill lands in the middle of the run
    if mode == "async":
        cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n_per * world, num_partitions=world, num_iterations=target,
                          step_size=0.4, batch_rate=0.3, loss="logistic", seed=42, backend="fused", barrier="bucket", bucket_ratio=1.0)
    else:
        cfg = AsyncConfig(algorithm="sgd_sync", num_features=d, num_rows=n_per * world, num_partitions=world, num_iterations=target,
                          step_size=0.4, batch_rate=0.3, loss="logistic", seed=42, backend="fused")
    parts = make_parts(world, n_per, d, 5, True)
    eng = FusedEngine(cfg, [shard_of(parts[rank], d, dev, rank)], n_per * world, device=dev)
    for rk in eng.kernels:
        rk.set(spin_limit_ns=int(8e9))
    eng.barrier()
    if rank == 1:
        eng.run_rounds(1_000_000)                     # a persistent worker loop that would run to the end of the job ...
        time.sleep(0.3)
        os.kill(os.getpid(), signal.SIGKILL)          # ... but the process dies mid-round, without saying goodbye
    # ---- rank 0: server + worker 0 + watchdog
    wd = DeviceWatchdog(eng, timeout_s=1.0).start(period_s=0.1, expect_running=True)
    t0 = time.time()
    out = {"mode": mode, "world": world}
    try:
        while time.time() - t0 < 60:
            eng.run_rounds(4000)
            eng.synchronize()
            eng.check_health()
            with eng.side_stream_ctx():
                k = eng.block.read_sctl()["k"] if eng.is_async else int(eng.kernels[0].read_ctl()["last_k_after"])
            if k >= target:
                break
        with eng.side_stream_ctx():
            st = eng.block.read_sctl()
            status = eng.block.read_status()
        out.update({"k": int(st["k"]) if eng.is_async else int(eng.kernels[0].read_ctl()["last_k_after"]), "dead": list(wd.dead),
                    "gate_dyn": st["gate_dyn"], "num_alive": st["num_alive"], "alive_flags": [int(x) for x in status["alive"]],
                    "elapsed_s": round(time.time() - t0, 2), "finite": bool(torch.isfinite(eng.block.w).all()), "raised": None})
    except RuntimeError as exc:
        out.update({"raised": str(exc)[:200], "dead": list(wd.dead), "elapsed_s": round(time.time() - t0, 2),
                    "abort": int(eng.kernels[0].read_ctl_raw()["abort"])})
    wd.stop()
    out["watchdog_error"] = repr(getattr(wd, "error", None))
    try:
        with eng.side_stream_ctx():
            st2 = eng.block.read_sctl()
            out["sctl"] = {k: (v if not isinstance(v, list) else v[:1]) for k, v in st2.items()}
            hb = eng.block.read_status()["heartbeat_ns"]
            out["hb_gap_ms"] = float(int(hb.max()) - int(hb.min())) * 1e-6
    except Exception as exc:  # noqa: BLE001
        out["sctl_error"] = repr(exc)[:200]
    print("ELASTIC_RESULT " + json.dumps(out), flush=True)
    os._exit(0)                                       # no collective teardown: the peer is gone


if __name__ == "__main__":
    main()
