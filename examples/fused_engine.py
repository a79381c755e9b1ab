#!/usr/bin/env python
"""The fused backend through its public API.  On a B200 box every round is one sm_100a kernel (gate + NVLink pull +
gradient + Hogwild push-and-apply); on a machine without a GPU the same object runs against the protocol simulator
(`ops/round_sim.py`), so this script runs anywhere.

    python examples/fused_engine.py                       # one process (1 GPU, or the CPU simulator)
    torchrun --nproc-per-node 8 examples/fused_engine.py  # 8 GPUs: rank 0's symmetric block is the parameter server

Reference loop being replaced: SparkASGDThread.scala:150-300 (driver thread + updater thread + Spark jobs).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from asyncframework_b200.config import AsyncConfig  # noqa: E402
from asyncframework_b200.core import AsyncContext  # noqa: E402
from asyncframework_b200.data.generators import shard_rows, synthetic_shard  # noqa: E402
from asyncframework_b200.parallel.comm import init_distributed  # noqa: E402
from asyncframework_b200.parallel.fused_backend import FusedEngine  # noqa: E402


def main(total_rows=200_000, d=784, workers_per_rank=2, rounds=40):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    use_cuda = torch.cuda.is_available()
    if world > 1:
        init_distributed()
    rank = dist.get_rank() if dist.is_initialized() else 0
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))) if use_cuda else torch.device("cpu")
    if not use_cuda:
        total_rows, d, rounds = 4000, 64, 10                      # the simulator is a spec, not a fast path
    P = world * workers_per_rank
    gamma = 8e-6 / P if use_cuda else 1e-3                        # bench.py's step size for the 784-pixel shape
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=total_rows, num_partitions=P, num_iterations=1 << 30, step_size=gamma,
                      batch_rate=0.1, loss="logistic", backend="fused", barrier="asp", seed=42)
    sizes = shard_rows(total_rows, P)
    ids = [p for p in range(P) if p % world == rank]
    shards = [synthetic_shard("mnist8m", sizes[p], p, P, cfg.seed, device, torch.float32, d, "binary01",
                              sum(sizes[:p])) for p in ids]
    eng = FusedEngine(cfg, shards, total_rows, ids, device=device)

    # 1. throughput path: enqueue rounds, no host synchronisation in between
    eng.run_rounds(rounds)
    eng.synchronize()
    eng.barrier()
    st = eng.server_state()
    if rank == 0:
        print(f"[{device}] {P} workers x {rounds} rounds -> k = {st['k']} accepted updates, clock = {st['clock']}, "
              f"|w| = {float(eng.weights().norm()):.4f}")

    # 2. end-to-end step: pinned H2D descriptor (step size) -> round -> pinned D2H result record
    r = eng.step(step_size=gamma / 2)
    if rank == 0:
        print("step():", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("accepted", "batch", "loss_sum", "staleness", "k_after")})

    # 3. the ASYNCcontext view of the same run: result envelopes from the device ring
    if rank == 0:
        ctx = AsyncContext(P)
        for res in eng.drain_results()[-3:]:
            ctx.put(res)
        while ctx.has_next():
            info = ctx.collect_all()
            print(f"  worker {info.worker_id}: staleness {info.staleness}, batch {info.batch_size}, accepted {info.data['accepted']}")
        print("  worker table (device mirror of workerState):", [int(x) for x in eng.worker_status()["num_tasks"]])
    eng.barrier()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
