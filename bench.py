#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): ASGD logistic-regression iterations/second on the
mnist8m shape (dense 8.1 M x 784, sampling rate 0.1), device-timed, max over ranks.

    python bench.py --gpus N --steps K --warmup W            # N == 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *iteration* = one round in which every worker pulls w, computes the gradient of its
Bernoulli(0.1) sample (b*N/P rows) in a single fused kernel and pushes/applies it into the
server (P accepted updates, b*N rows in total) — the SparkASGDSync notion of an iteration, so
numbers are comparable across N (strong scaling: the 8.1 M-row dataset is fixed and sharded).

The K timed iterations of the `ours` arm are ONE persistent kernel launch per worker (the device-side
worker loop of csrc/kernels/dense_stream.cu); the end-to-end number drives the same engine one step at a
time through its public API (pinned H2D round descriptor -> round -> pinned D2H result, every step).
The same process then runs the PyTorch + NCCL + cuBLAS arm (baseline/nccl_sync.py, the comparison bar
BASELINE.json names: the Spark/JVM reference has no GPU path) at the same N / K / W / dtype and reports
the ratios in `vs_baseline`.
"""
from __future__ import annotations

import argparse
import json
import math
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
    ap.add_argument("--no-baseline", action="store_true", help="skip the in-process NCCL arm (vs_baseline stays null)")
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"],
                    help="cpu = contract smoke-test against the pure-Python round simulator (NOT a benchmark)")
    ap.add_argument("--variant", default=None, choices=[None, "stream", "simple", "pipe", "ring"],
                    help="dense kernel variant (default stream: persistent TMA-ring kernel; the others are one-round A/B variants)")
    ap.add_argument("--no-pipeline", action="store_true", help="stream kernel: serial persistent rounds instead of pipelined ones")
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"],
                    help="row storage of the dense shards; parameters, gradient and all arithmetic are fp32 either way "
                         "(bf16 is exact for the integer pixels of the mnist shapes)")
    ap.add_argument("--graph", type=int, default=0, help="replay a CUDA graph of this many rounds (0 = plain launches)")
    ap.add_argument("--workers-per-gpu", type=int, default=0, help="logical workers (partitions) per GPU; 0 = auto (1)")
    ap.add_argument("--nccl-variant", default="both", choices=["both", "stock", "nosync"])
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
    # logical workers (= the reference's partitions).  One persistent, pipelined worker loop per GPU owns every SM: the
    # NVLink phases of round r hide behind round r+1's streaming inside the kernel, so co-located workers are not needed
    # (the one-round A/B variants still default to two workers per GPU, which hides one worker's NVLink latency behind the other).
    legacy = args.variant in ("simple", "pipe", "ring")
    wpg = args.workers_per_gpu or (2 if (legacy and world > 1 and args.impl == "ours" and args.algo in ("asgd", "asaga")) else 1)
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
    # random-init weights of the architecture (BASELINE.json: "random-init weights"): small gaussian, same on every arm
    gw = torch.Generator().manual_seed(1234)
    w0 = (1e-4 * torch.randn(args.features, generator=gw)).to(torch.float32)
    W = max(args.warmup, 3)

    def timed(run, steps, fork=None, join=None):
        """W warm-up steps were already run; time exactly `steps` steps with CUDA events on the launching stream."""
        barrier()
        e0, e1 = _dev.new_event(dev, enable_timing=True), _dev.new_event(dev, enable_timing=True)
        barrier()
        e0.record()
        if fork:
            fork(e0)
        run(steps)
        if join:
            join()
        e1.record()
        _dev.synchronize(dev)
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms, dev)

    def nccl_arm(variant):
        """The PyTorch + NCCL + cuBLAS arm at the same N / steps / warm-up / dtype: device-timed value + e2e."""
        from baseline.nccl_sync import NcclSyncSGD
        eng_b = NcclSyncSGD(shard.X, shard.y, args.features, args.rows, args.batch_rate, 8e-6, 42, args.loss,
                            saga=args.algo in ("asaga", "saga_sync"), variant=variant, w0=w0.to(dev))
        for _ in range(W):
            eng_b.step()
        ms = timed(lambda n: [eng_b.step() for _ in range(n)], args.steps)
        out = {"variant": variant, "value": args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps}
        if args.device == "cuda" and not args.no_e2e:
            for _ in range(3):
                eng_b.step_e2e()
            eng_b.flush_e2e()
            barrier()
            t1 = time.perf_counter()
            ea, eb = _dev.new_event(dev, enable_timing=True), _dev.new_event(dev, enable_timing=True)
            ea.record()
            last = None
            for _ in range(args.steps):
                r = eng_b.step_e2e()
                last = r if r is not None else last
            r = eng_b.flush_e2e()
            last = r if r is not None else last
            eb.record()
            _dev.synchronize(dev)
            ems = max(ea.elapsed_time(eb), (time.perf_counter() - t1) * 1e3)
            barrier()
            ems = max_over_ranks(ems, dev)
            out["e2e"] = {"value": args.steps / (ems * 1e-3), "ms_per_step": ems / args.steps,
                          "h2d_bytes_per_step": eng_b.h2d_bytes_per_step, "d2h_bytes_per_step": eng_b.d2h_bytes_per_step,
                          "last_loss": last}
        out["k"] = eng_b.k
        del eng_b
        return out

    if args.impl == "nccl":
        variants = ["stock", "nosync"] if args.nccl_variant == "both" else [args.nccl_variant]
        res = [nccl_arm(v) for v in variants]
        best = max(res, key=lambda r: r["value"])
        if rank == 0:
            print(json.dumps({"metric": "sgd_iterations_per_sec", "value": best["value"], "unit": "iterations/s", "n_gpus": world,
                              "steps": args.steps, "warmup": W, "ms_per_step": best["ms_per_step"], "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "impl": "nccl", "data": "synthetic",
                              "e2e": best.get("e2e"), "gpu_launches": 0, "variants": res,
                              "config": {"model": f"SGD_SYNC {args.loss} regression (linear model, d={args.features})",
                                         "dataset_shape": [args.rows, args.features], "global_batch": int(args.rows * args.batch_rate),
                                         "parallelism": f"dp{world} allreduce"}}))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    from asyncframework_b200.parallel.fused_backend import FusedEngine
    eng = FusedEngine(cfg, shards, args.rows, ids, device=dev, w0=w0.to(dev), pipeline=not args.no_pipeline)
    if args.graph > 0:
        if args.steps % args.graph or W % args.graph:
            raise SystemExit("--steps and --warmup must be multiples of --graph")
        eng.build_graph(args.graph)
        run = lambda n: eng.run_graph(n // args.graph)    # noqa: E731
    else:
        run = eng.run_rounds

    # ---- warm-up (untimed) -------------------------------------------------------------
    run(W)
    barrier()
    # ---- timed region: exactly K steps, CUDA events on the launching stream ----------------
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = eng.launches
    ms_max = timed(run, args.steps, fork=None if args.graph else eng.fork, join=None if args.graph else eng.join)
    launches = eng.launches - l0
    launches_all = int(sum_over_ranks(launches, dev))
    rounds_total = W + args.steps

    # ---- end-to-end through the public API (H2D descriptor, round, D2H result, every step) ---
    e2e = None
    if not args.no_e2e:
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
        rounds_total += 3 + args.steps
        e2e = {"value": args.steps / (e2e_ms * 1e-3), "unit": "iterations/s",
               "h2d_bytes_per_step": eng.h2d_bytes_per_step, "d2h_bytes_per_step": eng.d2h_bytes_per_step,
               "ms_per_step": e2e_ms / args.steps,
               "last_loss_per_row": (last["loss_sum"] / max(last["batch"], 1)) if last else None}
    clocks = sampler.stop() if sampler is not None else None

    phases = None
    if eng.is_async:
        ctl = eng.kernels[0].read_ctl_raw()
        t = [int(x) for x in ctl["dbg_ns"]]
        names = ["gate", "pull_rtt", "compute", "clock_rtt", "reds_issue", "fence_rtt", "publish"]
        phases = {n: (t[i + 1] - t[i]) / 1e3 for i, n in enumerate(names)}
    # ---- invariants of the run (the bench refuses to print a number for a broken run) ---------------------------
    eng.synchronize()
    if args.device == "cuda":
        eng.check_health()
    st = eng.state()
    k_total = st["k"] if eng.is_async else eng.kernels[0].read_ctl()["last_k_after"]
    k_all = int(max_over_ranks(float(eng.server_state()["k"]), dev)) if eng.is_async else int(k_total)
    expect_k = rounds_total * P if eng.is_async else rounds_total
    if k_all != expect_k:
        raise SystemExit(f"bench invariant violated: accepted updates {k_all} != rounds x workers {expect_k}")
    wfin = eng.weights()
    if not bool(torch.isfinite(wfin).all()):
        raise SystemExit("bench invariant violated: non-finite parameters")
    if e2e is not None and (e2e["last_loss_per_row"] is None or not math.isfinite(e2e["last_loss_per_row"])):
        raise SystemExit("bench invariant violated: non-finite loss")
    variant_eff = eng.kernels[0].variant_effective
    seed_mode = int(eng.kernels[0].params.seed_mode)
    shard_bytes = sum(sh.nbytes() for sh in shards)
    del eng

    # ---- the comparison arm, same process group, same N / steps / warm-up / dtype ----------------------------------
    vs = None
    if not args.no_baseline and args.device == "cuda":
        variants = ["stock", "nosync"] if args.nccl_variant == "both" else [args.nccl_variant]
        res = [nccl_arm(v) for v in variants]
        best = max(res, key=lambda r: r["value"])
        val = args.steps / (ms_max * 1e-3)
        vs = {"baseline": "PyTorch + NCCL + cuBLAS re-implementation of SparkASGDSync (baseline/nccl_sync.py), same box / N / steps / dtype",
              "value_ratio": val / best["value"], "baseline_value": best["value"], "baseline_variant": best["variant"],
              "e2e_ratio": (e2e["value"] / best["e2e"]["value"]) if (e2e and best.get("e2e")) else None,
              "baseline_e2e": best["e2e"]["value"] if best.get("e2e") else None,
              "variants": {r["variant"]: {"value": r["value"], "e2e": (r.get("e2e") or {}).get("value")} for r in res}}

    if rank == 0:
        esize = shard.X.element_size() if shard.is_dense else 4
        bytes_per_gpu_step = sum(sizes[p] for p in ids) * args.batch_rate * (shard.ld if shard.is_dense else 0) * esize
        out = {
            "metric": "sgd_iterations_per_sec", "value": args.steps / (ms_max * 1e-3), "unit": "iterations/s",
            "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "strong",
            # BASELINE.md publishes no number for this metric (the reference's only figure is a 4x wall-clock claim on a CPU
            # cluster); the ratio reported here is against the comparison arm BASELINE.json names, measured in this process
            "vs_baseline": vs,
            "dtype": ("bf16" if args.dtype == "bf16" else "fp32"),
            "dtype_detail": ("rows stored in bf16 (exact for the integer pixels), labels / parameters / gradient / accumulation fp32"
                             if args.dtype == "bf16" else "fp32 storage and arithmetic"),
            "device": args.device,
            "data": "synthetic (mnist8m shape: integer pixels 0..255, ~19% non-zero, labels from random true weights; random-init "
                    "model weights N(0, 1e-4^2); no network for the real dataset)",
            "impl": "ours",
            "config": {"model": f"{args.algo.upper()} {args.loss} regression (linear model, d={args.features})",
                       "dataset_shape": [args.rows, args.features], "global_batch": int(args.rows * args.batch_rate),
                       "sampling_rate": args.batch_rate, "seq_len": None, "parallelism": f"dp{world} parameter-server (rank 0 = server + worker)",
                       "updates_per_iteration": P, "accepted_updates_total": int(k_all),
                       "l2_policy": "inputs larger than L2: each round streams a fresh Bernoulli sample "
                                    f"(~{bytes_per_gpu_step / 1e6:.0f} MB/GPU) out of {shard_bytes / 1e9:.2f} GB of HBM-resident shards",
                       "workers_per_gpu": wpg, "cuda_graph_rounds": args.graph,
                       "seed_mode": seed_mode,
                       "seed_note": "sampling seed = seed + worker-local round + 1 (the reference uses seed + k + 1; documented deviation, "
                                    "same role: a fresh sample per submission)",
                       "dense_variant": variant_eff,
                       "timed_region": "K rounds per worker = one persistent kernel launch per worker (device-side worker loop)",
                       "e2e_inputs": "dataset cached in HBM (as the reference's .cache()); every step: pinned H2D round descriptor "
                                     "{step size, Bernoulli threshold, delay intensity} read by the kernel -> one fused round -> pinned D2H "
                                     "result records (host reads them one step behind)",
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
