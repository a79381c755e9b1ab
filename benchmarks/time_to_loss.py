#!/usr/bin/env python
"""Wall-clock (device time) to a target training loss: fused asynchronous engine vs the PyTorch + NCCL baseline
(BASELINE.md "wall-clock to target training loss"; the reference's own headline is of this kind: "4x faster than Spark
to equal error", Presentation/slides.pdf).

    python benchmarks/time_to_loss.py --shape mnist8m --target 0.60                      # 1 GPU
    torchrun --nproc-per-node 8 benchmarks/time_to_loss.py --shape mnist8m --target 0.60 [--impl both]

Both arms train logistic regression on the same synthetic shards (same generator, same seed) with the same per-iteration
step size and sampling rate; training time is measured with CUDA events around the training rounds only (max over ranks),
the full-data objective is evaluated between timed segments with the same plain-PyTorch code for both arms.  One iteration
= every worker processes one Bernoulli(b) sample of its shard.  Prints one JSON line per arm.
(On a machine without a GPU only `--impl ours` runs, against the protocol simulator, with a tiny shape — a smoke test.)
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

from asyncframework_b200.config import AsyncConfig  # noqa: E402
from asyncframework_b200.data.generators import SHAPES, shard_rows, synthetic_shard  # noqa: E402
from asyncframework_b200.parallel.comm import init_distributed  # noqa: E402
from asyncframework_b200.parallel.fused_backend import FusedEngine  # noqa: E402


def full_objective(shards, w, total_rows, world):
    """Mean logistic loss over every row of every shard (plain PyTorch, identical for both arms)."""
    dev = shards[0].y.device
    acc = torch.zeros((), dtype=torch.float64, device=dev)
    for s in shards:
        if s.is_dense and s.X.dtype != s.y.dtype:          # bf16 row storage: widen in chunks (the shard stays bf16 in HBM)
            wv = w[: s.num_features].float()
            for a in range(0, s.num_rows, 262144):
                m = s.X[a:a + 262144, : s.num_features].float() @ wv
                yy = s.y[a:a + 262144]
                acc += (torch.nn.functional.softplus(-m) + m * (1 - yy)).double().sum()
            continue
        margin = s.matvec(w[: s.num_features].to(s.y.dtype))
        acc += (torch.nn.functional.softplus(-margin) + margin * (1 - s.y)).double().sum()
    if world > 1:
        dist.all_reduce(acc)
    return float(acc) / total_rows


class Timer:
    """CUDA events on GPU, perf_counter on the CPU simulator."""

    def __init__(self, dev):
        self.cuda = dev.type == "cuda"
        self.ms = 0.0

    def __enter__(self):
        if self.cuda:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        else:
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        if self.cuda:
            self.e1.record()
            torch.cuda.synchronize()
            self.ms += self.e0.elapsed_time(self.e1)
        else:
            self.ms += (time.perf_counter() - self.t0) * 1e3
        return False


def max_ranks(v, dev, world):
    if world == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_ours(args, rank, world, dev, shards_by_worker, ids, total, P):
    cfg = AsyncConfig(algorithm="asgd", num_features=shards_by_worker[0].num_features, num_rows=total, num_partitions=P,
                      num_iterations=1 << 40, step_size=args.gamma / P, batch_rate=args.b, loss="logistic", seed=42, backend="fused",
                      barrier="asp")
    eng = FusedEngine(cfg, shards_by_worker, total, ids, device=dev)
    # untimed warm-up (module load, clocks), then the server state is reset so the timed run starts from w = 0
    eng.run_rounds(args.warmup)
    eng.synchronize()
    eng.barrier()
    eng.block.reset()
    eng._ring_seen = 0
    eng.synchronize()
    eng.barrier()
    curve = [(0.0, full_objective(shards_by_worker, eng.weights(), total, world))]
    timer = Timer(dev)
    done = 0
    while done < args.max_iters and curve[-1][1] > args.target:
        with timer:
            if timer.cuda:
                eng.fork(timer.e0)
            eng.run_rounds(args.eval_every)
            if timer.cuda:
                eng.join()
            else:
                eng.synchronize()
        done += args.eval_every
        eng.barrier()
        curve.append((max_ranks(timer.ms, dev, world), full_objective(shards_by_worker, eng.weights(), total, world)))
    return {"impl": "ours", "algo": "asgd (fused, asynchronous)", "workers": P, "iterations": done, "curve": curve}


def run_nccl(args, rank, world, dev, shards, total):
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from nccl_sync import NcclSyncSGD
    s = shards[0]
    opt = NcclSyncSGD(s.X, s.y, s.num_features, total, args.b, args.gamma, seed=42, loss="logistic", variant=args.nccl_variant)
    for _ in range(args.warmup):                  # untimed warm-up, then restart from w = 0
        opt.step()
    opt.w.zero_()
    opt.k = 0
    curve = [(0.0, full_objective(shards, opt.w, total, world))]
    timer = Timer(dev)
    done = 0
    while done < args.max_iters and curve[-1][1] > args.target:
        with timer:
            for _ in range(args.eval_every):
                opt.step()
        done += args.eval_every
        if world > 1:
            dist.barrier()
        curve.append((max_ranks(timer.ms, dev, world), full_objective(shards, opt.w, total, world)))
    return {"impl": "nccl", "algo": "sgd (PyTorch + NCCL + cuBLAS, synchronous)", "workers": world, "iterations": done, "curve": curve}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="mnist8m", choices=[k for k, v in SHAPES.items() if v[2] == "dense"])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--target", type=float, default=0.60, help="full-data mean logistic loss to reach (starts at ln 2 = 0.693)")
    ap.add_argument("--b", type=float, default=0.1)
    ap.add_argument("--gamma", type=float, default=8e-6, help="per-ITERATION step (async applies gamma / P per update)")
    ap.add_argument("--eval-every", type=int, default=25)
    ap.add_argument("--max-iters", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workers-per-gpu", type=int, default=0)
    ap.add_argument("--impl", default="both", choices=["both", "ours", "nccl"])
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"], help="row storage of BOTH arms (exact for the integer pixels)")
    ap.add_argument("--nccl-variant", default="nosync", choices=["stock", "nosync"])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rank, local_rank, world = init_distributed()
    cuda = torch.cuda.is_available()
    dev = torch.device("cuda", local_rank) if cuda else torch.device("cpu")
    if cuda:
        torch.cuda.set_device(dev)
    elif args.impl != "ours":
        raise SystemExit("the NCCL arm needs GPUs; use --impl ours for the simulator smoke run")
    total = int(args.rows or (SHAPES[args.shape][0] if cuda else 4000))
    results = []
    if args.impl in ("both", "ours"):
        wpg = args.workers_per_gpu or 1            # one persistent, pipelined worker loop per GPU
        P = world * wpg
        sizes = shard_rows(total, P)
        ids = [p for p in range(P) if p % world == rank]
        rdt = torch.bfloat16 if (args.dtype == "bf16" and cuda) else None
        shards = [synthetic_shard(args.shape, sizes[p], p, P, 42, dev, torch.float32, None, "binary01", sum(sizes[:p]), row_dtype=rdt) for p in ids]
        results.append(run_ours(args, rank, world, dev, shards, ids, total, P))
        del shards
    if args.impl in ("both", "nccl"):
        sizes = shard_rows(total, world)
        rdt = torch.bfloat16 if (args.dtype == "bf16" and cuda) else None
        shards = [synthetic_shard(args.shape, sizes[rank], rank, world, 42, dev, torch.float32, None, "binary01", sum(sizes[:rank]), row_dtype=rdt)]
        results.append(run_nccl(args, rank, world, dev, shards, total))
    if rank == 0:
        for r in results:
            reached = r["curve"][-1][1] <= args.target
            r.update({"shape": args.shape, "rows": total, "n_gpus": world, "target": args.target, "reached": reached,
                      "ms_to_target": r["curve"][-1][0] if reached else None, "final_loss": r["curve"][-1][1],
                      "curve": [(round(t, 3), round(o, 6)) for t, o in r["curve"]]})
            print("TIME_TO_LOSS " + json.dumps(r), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                for r in results:
                    f.write(json.dumps(r) + "\n")
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
