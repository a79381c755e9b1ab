#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ASGD logistic-regression iterations/second on the
mnist8m shape (dense 8.1 M x 784, sampling rate 0.1), device-timed, max over ranks.

    python bench.py --gpus N --steps K --warmup W            # N == 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *iteration* = one round in which every worker pulls w, computes the gradient of its
Bernoulli(0.1) sample (b*N/P rows) in a single fused kernel and pushes/applies it into the
server (P accepted updates, b*N rows in total) — the SparkASGDSync notion of an iteration, so
numbers are comparable across N (strong scaling: the 8.1 M-row dataset is fixed and sharded).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, FEATURES, BATCH_RATE = 8_100_000, 784, 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "nccl"])
    ap.add_argument("--algo", default="asgd", choices=["asgd", "asaga", "sgd_sync", "saga_sync"])
    ap.add_argument("--rows", type=int, default=ROWS)
    ap.add_argument("--features", type=int, default=FEATURES)
    ap.add_argument("--batch-rate", type=float, default=BATCH_RATE)
    ap.add_argument("--step-size", type=float, default=0.0, help="per-update step (0 = 8e-6 / #workers for async, 8e-6 for sync)")
    ap.add_argument("--loss", default="logistic")
    ap.add_argument("--shape", default="mnist8m")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"],
                    help="cpu = contract smoke-test against the pure-Python round simulator (NOT a benchmark)")
    ap.add_argument("--variant", default=None, choices=[None, "simple", "pipe", "ring", "persist", "pipelined"],
                    help="dense kernel variant (default: env/auto; persist = experimental one-launch-per-K-rounds kernel)")
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="row storage of the dense shards (math is fp32 either way); bf16 = experimental kernel, exact for integer pixels")
    ap.add_argument("--graph", type=int, default=0, help="replay a CUDA graph of this many rounds (0 = plain launches)")
    ap.add_argument("--stagger-us", type=float, default=0.0, help="one-time phase offset between co-located workers")
    ap.add_argument("--workers-per-gpu", type=int, default=0, help="logical workers (partitions) per GPU; 0 = auto")
    return ap.parse_args()


def reference_unavailable():
    if int(os.environ.get("RANK", "0")) != 0:      # under torchrun only rank 0 reports
        return 0
    print(json.dumps({"impl": "reference", "unavailable":
                      "reference is a JVM/Scala Apache Spark 2.3.2 fork (no setup.py/pyproject, `pip install /root/reference` "
                      "fails: not installable) and this image has no Java/Scala/sbt; it has no GPU code path"}))
    return 0


def main() -> int:
    args = parse()
    if args.impl == "reference":
        return reference_unavailable()
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world_env == 1:
        # launched as plain `python bench.py --gpus N`: re-exec under torchrun
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)

    import torch
    import torch.distributed as dist
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.data.generators import shard_rows, synthetic_shard
    from asyncframework_b200.parallel.comm import init_distributed, max_over_ranks, sum_over_ranks
    from asyncframework_b200.utils.clocks import ClockSampler

    if args.variant:
        os.environ["AFB_DENSE_VARIANT"] = args.variant
    rank, local_rank, world = init_distributed()
    from asyncframework_b200.parallel import _devapi as _dev
    dev = torch.device("cuda", local_rank) if args.device == "cuda" else torch.device("cpu")
    _dev.set_device(dev)
    # logical workers (= the reference's partitions).  The reference's clusters run 2 tasks per
    # executor (1 executor x 2 cores, README.md:56); two workers per GPU play the same role: one
    # worker's NVLink pull/push latency hides behind the other's HBM streaming.
    wpg = args.workers_per_gpu or (2 if (world > 1 and args.impl == "ours" and args.algo in ("asgd", "asaga")) else 1)
    if args.variant == "pipelined" and not args.workers_per_gpu:
        wpg = 1                                   # the pipelined persistent kernel owns every SM of its GPU
    P = world * wpg
    sizes = shard_rows(args.rows, P)
    ids = [p for p in range(P) if p % world == rank]
    t0 = time.time()
    shards = [synthetic_shard(args.shape, sizes[p], p, P, seed=42, device=dev, num_features=args.features,
                              labels="binary01" if args.loss != "least_squares" else "regression", row_offset=sum(sizes[:p]),
                              row_dtype=(torch.bfloat16 if args.dtype == "bf16" else None))
              for p in ids]
    shard = shards[0]
    _dev.synchronize(dev)
    gen_s = time.time() - t0

    def barrier():
        _dev.synchronize(dev)
        if world > 1:
            dist.barrier()

    if args.step_size <= 0:
        args.step_size = 8e-6 / P if args.algo in ("asgd", "asaga") else 8e-6
    cfg = AsyncConfig(algorithm=args.algo, num_features=args.features, num_rows=args.rows, num_partitions=P,
                      num_iterations=1 << 40, step_size=args.step_size, batch_rate=args.batch_rate, bucket_ratio=0.0,
                      barrier="asp", loss=args.loss, seed=42, backend="fused")

    if args.impl == "nccl":
        sys.path.insert(0, os.path.join(ROOT))
        from baseline.nccl_sync import NcclSyncSGD
        eng = NcclSyncSGD(shard.X, shard.y, args.features, args.rows, args.batch_rate, args.step_size, 42, args.loss,
                          saga=args.algo in ("asaga", "saga_sync"))
        run = lambda n: [eng.step() for _ in range(n)]      # noqa: E731
        launches_of = lambda: 0                              # noqa: E731
    else:
        from asyncframework_b200.parallel.fused_backend import FusedEngine
        eng = FusedEngine(cfg, shards, args.rows, ids, device=dev)
        if args.graph > 0:
            if args.steps % args.graph or max(args.warmup, 3) % args.graph:
                raise SystemExit("--steps and --warmup must be multiples of --graph")
            eng.build_graph(args.graph)
            run = lambda n: eng.run_graph(n // args.graph)    # noqa: E731
        else:
            run = eng.run_rounds
        launches_of = lambda: eng.launches                   # noqa: E731

    # ---- warm-up (untimed) -------------------------------------------------------------
    if args.impl == "ours" and args.stagger_us > 0:
        eng.stagger(args.stagger_us)
    run(max(args.warmup, 3))
    barrier()
    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = launches_of()
    e0, e1 = _dev.new_event(dev, enable_timing=True), _dev.new_event(dev, enable_timing=True)
    barrier()
    e0.record()
    if args.impl == "ours" and not args.graph:
        eng.fork(e0)                 # worker streams start after e0
    run(args.steps)
    if args.impl == "ours" and not args.graph:
        eng.join()                   # launching stream waits for every worker stream
    e1.record()
    _dev.synchronize(dev)
    ms = e0.elapsed_time(e1)
    barrier()
    launches = launches_of() - l0
    ms_max = max_over_ranks(ms, dev)
    launches_all = int(sum_over_ranks(launches, dev))

    # ---- end-to-end through the public API (H2D descriptor, round, D2H result, every step) ---
    e2e = None
    if args.impl == "ours" and not args.no_e2e:
        for _ in range(3):
            eng.step()
        barrier()
        t1 = time.perf_counter()
        e2, e3 = _dev.new_event(dev, enable_timing=True), _dev.new_event(dev, enable_timing=True)
        e2.record()
        last = None
        for _ in range(args.steps):
            r = eng.step(lag=1)          # H2D descriptor + round + D2H record EVERY step; results read one step behind
            last = r or last
        last = eng.flush_steps() or last
        e3.record()
        _dev.synchronize(dev)
        e2e_ms = max(e2.elapsed_time(e3), (time.perf_counter() - t1) * 1e3)
        barrier()
        e2e_ms = max_over_ranks(e2e_ms, dev)
        e2e = {"value": args.steps / (e2e_ms * 1e-3), "unit": "iterations/s",
               "h2d_bytes_per_step": eng.h2d_bytes_per_step, "d2h_bytes_per_step": eng.d2h_bytes_per_step,
               "ms_per_step": e2e_ms / args.steps,
               "last_loss_per_row": (last["loss_sum"] / max(last["batch"], 1)) if last else None}
    clocks = sampler.stop() if sampler is not None else None

    phases = None
    if args.impl == "ours" and eng.is_async:
        ctl = eng.kernels[0].read_ctl_raw()
        t = [int(x) for x in ctl["dbg_ns"]]
        names = ["gate", "pull_rtt", "compute", "clock_rtt", "reds_issue", "fence_rtt", "publish"]
        phases = {n: (t[i + 1] - t[i]) / 1e3 for i, n in enumerate(names)}
        # absolute phase stamps of the last round of every local worker, relative to worker 0's kernel entry
        t00 = t[0]
        phases["workers_abs_us"] = [[round((int(x) - t00) / 1e3, 1) for x in rk.read_ctl_raw()["dbg_ns"]] for rk in eng.kernels]
    if args.impl == "ours":
        st = eng.state()
        k_total = st["k"] if eng.is_async else eng.kernels[0].read_ctl()["last_k_after"]
    else:
        k_total = eng.k
    if rank == 0:
        bytes_per_gpu_step = sum(sizes[p] for p in ids) * args.batch_rate * (shard.ld if shard.is_dense else 0) * (shard.X.element_size() if shard.is_dense else 4)
        out = {
            "metric": "sgd_iterations_per_sec", "value": args.steps / (ms_max * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ("bf16 rows, fp32 accumulate" if args.dtype == "bf16" else "fp32"),
            "device": args.device,
            "data": "synthetic (mnist8m shape: integer pixels 0..255, ~19% non-zero, random-init true weights; no network for the real dataset)",
            "impl": args.impl,
            "config": {"model": f"{args.algo.upper()} {args.loss} regression (linear model, d={args.features})",
                       "dataset_shape": [args.rows, args.features], "global_batch": int(args.rows * args.batch_rate),
                       "sampling_rate": args.batch_rate, "seq_len": None, "parallelism": f"dp{world} parameter-server (rank 0 = server + worker)",
                       "updates_per_iteration": P, "accepted_updates_total": int(k_total),
                       "l2_policy": "inputs larger than L2: each round streams a fresh Bernoulli sample "
                                    f"(~{bytes_per_gpu_step / 1e6:.0f} MB/GPU) out of {sum(sh.nbytes() for sh in shards) / 1e9:.2f} GB of HBM-resident shards",
                       "workers_per_gpu": wpg, "cuda_graph_rounds": args.graph,
                       "pdl": (bool(eng.kernels[0].use_pdl) if args.impl == "ours" else None),
                       "seed_mode": (int(eng.kernels[0].params.seed_mode) if args.impl == "ours" else None), "dense_variant": (eng.kernels[0].variant_effective if args.impl == "ours" else None),
                       "e2e_inputs": "dataset cached in HBM (as the reference's .cache()); every step: pinned H2D round descriptor -> fused round(s) -> pinned D2H result records (host reads them one step behind)",
                       "datagen_s": round(gen_s, 2)},
            "hbm_gbps_per_gpu": bytes_per_gpu_step / (ms_max / args.steps * 1e-3) / 1e9,
            "gpu_launches": launches_all,
            "last_round_phases_us_rank0_worker0": phases,
            "clocks": clocks,
        }
        if e2e is not None:
            out["e2e"] = e2e
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
