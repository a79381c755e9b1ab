#!/usr/bin/env python
"""BASELINE.json configs 2-5 on the fused backend (run under torchrun for N > 1):

  mnist8m_asgd   ASGD logistic, dense 8.1M x 784, sampling 0.1
  epsilon_asaga  ASAGA (per-sample history), dense 400k x 2000, sampling 0.1
  rcv1_asgd      ASGD, CSR 677k x 47 236 (~74 nnz/row), sampling 0.05
  straggler      ASAGA vs synchronous SAGA under delay intensity {0.3, 0.6, 1.0}, mnist8m shape
  production     the reference's production-cluster experiment (README.md:72-74): 32 logical workers (4 per GPU at 8 GPUs),
                 25 % stragglers (delay -1: 80 % normal U[1.5,2.5), 20 % long-tail U[2.5,10) x the mean task time), beta = 0.7:
                 ASGD / ASAGA vs the synchronous variants at equal work

Each run reports device-timed iterations/s (max over ranks) and wall-clock to a target training loss
(full-data objective evaluated with the K12 kernel every `--eval-every` iterations).
One iteration = every worker completes one round (b*N rows in total).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from asyncframework_b200.algo.fused_driver import evaluate_snapshots  # noqa: E402
from asyncframework_b200.config import AsyncConfig  # noqa: E402
from asyncframework_b200.data.generators import SHAPES, shard_rows, synthetic_shard  # noqa: E402
from asyncframework_b200.parallel.comm import init_distributed, max_over_ranks  # noqa: E402
from asyncframework_b200.parallel.fused_backend import FusedEngine  # noqa: E402

CONFIGS = {
    # gamma is the per-ITERATION step; async algorithms apply P updates per iteration, so they run with
    # gamma / P per update (the reference's README does the same: async step = sync step / #partitions)
    "mnist8m_asgd": dict(shape="mnist8m", algo="asgd", loss="logistic", b=0.1, gamma=8e-6, rows=None),
    # unit-norm rows => L = 1/4 for the logistic loss: constant-step SAGA is stable up to ~1/(3L) per update (1.0 per update at
    # P = 8); ASGD decays as 1/sqrt(floor(k/P)+1), so it starts from a correspondingly larger step
    "epsilon_asaga": dict(shape="epsilon", algo="asaga", loss="logistic", b=0.1, gamma=8.0, rows=None),
    "rcv1_asgd": dict(shape="rcv1", algo="asgd", loss="logistic", b=0.05, gamma=32.0, rows=None),
}


_SHARD_CACHE = {}


def shards_for(shape, total, P, world, rank, dev, loss, row_dtype):
    """HBM-resident shards of this rank, generated once per (shape, P, dtype) and shared by every run of the sweep."""
    key = (shape, total, P, str(row_dtype))
    if key not in _SHARD_CACHE:
        _SHARD_CACHE.clear()                   # one resident dataset at a time
        torch.cuda.empty_cache()
        sizes = shard_rows(total, P)
        ids = [p for p in range(P) if p % world == rank]
        labels = "regression" if loss.startswith("least_squares") else "binary01"
        dense = SHAPES[shape][2] == "dense"
        _SHARD_CACHE[key] = ([synthetic_shard(shape, sizes[p], p, P, 42, dev, torch.float32, None, labels, sum(sizes[:p]),
                                              row_dtype=(row_dtype if dense else None)) for p in ids], ids)
    return _SHARD_CACHE[key]


def run_one(name, algo, shape, loss, b, gamma, rows, world, rank, dev, iters, eval_every, target_frac, coeff=0.0,
            wpg=1, warm_per_worker=20, rounds_per_batch=50, by_updates=False, row_dtype=None, barrier="asp", beta=0.0):
    P = world * wpg
    if algo in ("sgd_sync", "saga_sync") and wpg != 1:
        raise ValueError("BSP fused rounds run one worker per rank")
    total = int(rows or SHAPES[shape][0])
    shards, ids = shards_for(shape, total, P, world, rank, dev, loss, row_dtype)
    if algo in ("asgd", "asaga"):
        gamma = gamma / P
    is_async = algo in ("asgd", "asaga")
    # `by_updates`: an async run is finished when the SERVER has accepted iters*P updates (the reference's
    # termination rule, `k >= numIteration`): fast workers keep going while a straggler sleeps.
    warm_rounds = 3                      # untimed: lazy module loading, first-launch attribute calls, clocks
    max_updates = (iters + warm_rounds) * P if (by_updates and is_async) else (1 << 40)
    cfg = AsyncConfig(algorithm=algo, num_features=shards[0].num_features, num_rows=total, num_partitions=P,
                      num_iterations=max_updates, step_size=gamma, batch_rate=b, loss=loss, seed=42, backend="fused",
                      barrier=barrier if is_async else "asp", bucket_ratio=beta, coeff=coeff, warmup_updates_per_worker=warm_per_worker)
    eng = FusedEngine(cfg, shards, total, ids, device=dev)

    def objective():
        return evaluate_snapshots(eng, shards, [(0.0, eng.weights())] if rank == 0 else [(0.0, None)], loss, total)[0][1] \
            if rank == 0 or world > 1 else None

    def objective_all():
        snaps = [(0.0, eng.weights())]
        val = evaluate_snapshots(eng, shards, snaps, loss, total)
        return val[0][1] if val else None

    obj0 = objective_all()
    if world > 1:
        t = torch.tensor([obj0 if obj0 is not None else 0.0], device=dev, dtype=torch.float64)
        dist.broadcast(t, 0)
        obj0 = float(t.item())
    eng.run_rounds(warm_rounds if (by_updates and is_async) else 5)      # warm-up
    eng.synchronize()
    eng.barrier()
    if by_updates and is_async:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        eng.fork(e0)
        rounds = 0
        while True:
            eng.run_rounds(rounds_per_batch)
            rounds += rounds_per_batch
            if eng.server_state()["k"] >= max_updates or rounds > 50 * iters:
                break
        eng.join()
        e1.record()
        torch.cuda.synchronize()
        dev_ms = max_over_ranks(e0.elapsed_time(e1), dev)
        eng.barrier()
        obj = objective_all()
        if world > 1:
            t = torch.tensor([obj if obj is not None else 0.0], device=dev, dtype=torch.float64)
            dist.broadcast(t, 0)
            obj = float(t.item())
        st = eng.server_state()
        res = {"config": name, "algo": algo, "shape": shape, "n_gpus": world, "workers": P, "delay_intensity": coeff,
               "iterations": iters, "iterations_per_s": iters / (dev_ms * 1e-3), "ms_per_iteration": dev_ms / iters,
               "objective_start": obj0, "objective_end": obj, "accepted_updates": min(st["k"], max_updates),
               "local_rounds_rank0": rounds, "avg_delay_ms": st["avg_delay_ns"] * 1e-6, "mode": "fixed total updates (iters*P)",
               "variant": eng.kernels[0].variant_effective, "row_dtype": str(row_dtype or torch.float32), "barrier": barrier, "beta": beta}
        eng.check_health()
        eng.barrier()
        del eng
        torch.cuda.empty_cache()
        return res
    curve = [(0.0, obj0)]
    target = None
    t_target = None
    done = 0
    dev_ms = 0.0
    t0 = time.perf_counter()
    while done < iters:
        n = min(eval_every, iters - done)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.fork(e0)
        for _ in range(0, n, rounds_per_batch):
            eng.run_rounds(min(rounds_per_batch, n))
        eng.join()
        e1.record()
        torch.cuda.synchronize()
        dev_ms += e0.elapsed_time(e1)
        done += n
        eng.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        obj = objective_all()
        if world > 1:
            t = torch.tensor([obj if obj is not None else 0.0], device=dev, dtype=torch.float64)
            dist.broadcast(t, 0)
            obj = float(t.item())
        curve.append((wall, obj))
        if target is None:
            target = obj0 * target_frac
        if t_target is None and obj <= target:
            t_target = wall
            break
    dev_ms = max_over_ranks(dev_ms, dev)
    st = eng.server_state()
    res = {"config": name, "algo": algo, "shape": shape, "n_gpus": world, "workers": P, "delay_intensity": coeff,
           "iterations": done, "iterations_per_s": done / (dev_ms * 1e-3), "ms_per_iteration": dev_ms / done,
           "objective_start": obj0, "objective_end": curve[-1][1], "target": target, "wall_ms_to_target": t_target,
           "accepted_updates": st["k"] if eng.is_async else done, "avg_delay_ms": st["avg_delay_ns"] * 1e-6,
           "curve": [(round(a, 2), round(o, 6)) for a, o in curve],
           "variant": eng.kernels[0].variant_effective, "row_dtype": str(row_dtype or torch.float32), "launches": eng.launches}
    eng.check_health()
    eng.barrier()
    del eng
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=list(CONFIGS) + ["straggler", "production"])
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"], help="row storage of the mnist8m-shape shards (exact for integer pixels)")
    ap.add_argument("--gamma", type=float, default=0.0, help="override the per-iteration step size of the config")
    ap.add_argument("--iters", type=int, default=400)
    ap.add_argument("--eval-every", type=int, default=50)
    ap.add_argument("--target-frac", type=float, default=0.9, help="target = this fraction of the initial objective")
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--workers-per-gpu", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--delays", type=float, nargs="+", default=[0.0, 0.3, 0.6, 1.0])
    args = ap.parse_args()
    rank, local_rank, world = init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    results = []
    mnist_dt = torch.bfloat16 if args.dtype == "bf16" else None
    if args.config == "straggler":
        for coeff in args.delays:
            for algo, gamma in (("asaga", 3e-6), ("saga_sync", 3e-6)):
                results.append(run_one("straggler", algo, "mnist8m", "logistic", 0.1, gamma, args.rows or None, world, rank, dev,
                                       args.iters, args.iters, args.target_frac, coeff=coeff,
                                       wpg=args.workers_per_gpu if algo == "asaga" else 1, by_updates=True, row_dtype=mnist_dt))
                if rank == 0:
                    r = results[-1]
                    print(f"[straggler] {algo:10s} delay={coeff:.1f} equal-work it/s={r['iterations_per_s']:.1f} "
                          f"ms/it={r['ms_per_iteration']:.4f} obj {r['objective_start']:.4f}->{r['objective_end']:.4f} "
                          f"avg_delay_ms={r['avg_delay_ms']:.4f}", flush=True)
    elif args.config == "production":
        # README.md:72-74: 32 workers, 8 of them stragglers (delay -1), beta 0.7, b 0.1; ASYNC vs sync at equal work.
        # Three arms: the reference's shape (32 logical workers = 4 per GPU behind the beta = 0.7 wave gate), this engine's
        # native shape (one pipelined worker loop per GPU, fully asynchronous), and the BSP variants (one worker per GPU).
        # The straggler model marks 25 % of the workers of EVERY arm (indices 4c): 8 of 32, 2 of 8.
        wpg = max(1, 32 // world)
        arms = [("asgd", 8e-6, wpg, "bucket", 0.7), ("asaga", 3e-6, wpg, "bucket", 0.7),
                ("asgd", 8e-6, 1, "asp", 0.0), ("asaga", 3e-6, 1, "asp", 0.0),
                ("sgd_sync", 8e-6, 1, "asp", 0.0), ("saga_sync", 3e-6, 1, "asp", 0.0)]
        for algo, gamma, w_, bar_, beta_ in arms:
            results.append(run_one("production", algo, "mnist8m", "logistic", 0.1, gamma, args.rows or None, world, rank, dev,
                                   args.iters, args.iters, args.target_frac, coeff=-1.0, wpg=w_, by_updates=True,
                                   row_dtype=mnist_dt, barrier=bar_, beta=beta_))
            if rank == 0:
                r = results[-1]
                print(f"[production] {algo:10s} workers={r['workers']} equal-work it/s={r['iterations_per_s']:.1f} "
                      f"ms/it={r['ms_per_iteration']:.4f} obj {r['objective_start']:.4f}->{r['objective_end']:.4f} "
                      f"avg_delay_ms={r['avg_delay_ms']:.4f}", flush=True)
    else:
        c = CONFIGS[args.config]
        results.append(run_one(args.config, c["algo"], c["shape"], c["loss"], c["b"], args.gamma or c["gamma"], args.rows or c["rows"], world,
                               rank, dev, args.iters, args.eval_every, args.target_frac, wpg=args.workers_per_gpu,
                               row_dtype=mnist_dt if c["shape"] == "mnist8m" else None))
    if rank == 0:
        for r in results:
            print("CONFIG_RESULT " + json.dumps(r))
        if args.out:
            with open(args.out, "a") as f:
                for r in results:
                    f.write(json.dumps(r) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
