"""Multi-GPU worker script (launched by torchrun from tests/test_multigpu.py or by hand):
checks the fused NVLink paths against the NumPy oracle / invariants and prints one JSON line."""
import json
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200.algo.oracle import AsyncOracle, run_saga_sync_oracle, run_sgd_sync_oracle  # noqa: E402
from asyncframework_b200.config import AsyncConfig  # noqa: E402
from asyncframework_b200.data.batch import Shard, padded_ld  # noqa: E402
from asyncframework_b200.parallel.comm import allocate_symmetric, init_distributed  # noqa: E402
from asyncframework_b200.parallel.fused_backend import FusedEngine  # noqa: E402


def make_parts(P, n_per, d, seed, labels01):
    rng = np.random.RandomState(seed)
    wstar = rng.randn(d) * 0.5
    parts, off = [], 0
    for p in range(P):
        X = rng.randn(n_per, d) * 0.4
        if labels01:
            y = (rng.rand(n_per) < 1 / (1 + np.exp(-X @ wstar))).astype(np.float64)
        else:
            y = X @ wstar + 0.05 * rng.randn(n_per)
        parts.append((X.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64), np.arange(n_per) + off))
        off += n_per
    return parts


def shard_of(part, d, dev, p):
    X, y, g = part
    ld = padded_ld(d)
    Xp = torch.zeros(X.shape[0], ld, dtype=torch.float32)
    Xp[:, :d] = torch.from_numpy(X).float()
    return Shard(d, torch.from_numpy(y).float().to(dev), Xp.to(dev), gidx=torch.from_numpy(g).to(dev), partition=p)


def main():
    rank, local_rank, world = init_distributed()
    dev = torch.device("cuda", local_rank)
    out = {"world": world}
    d, n_per = 100, 3000

    # ---- 0. symmetric memory + P2P litmus -------------------------------------------------
    sb = allocate_symmetric(4096, dev)
    out["provider"] = sb.provider
    out["multicast"] = bool(sb.multicast_ptr)
    sb.local.view(torch.int32)[0] = 1000 + rank
    sb.barrier()
    if world > 1:
        from asyncframework_b200.ops import kernels
        import ctypes as C
        peer = (rank + 1) % world
        host = np.empty(4, dtype=np.uint8)
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy(C.c_void_p(host.ctypes.data), C.c_void_p(sb.ptrs[peer]), C.c_size_t(4), 2)
        out["peer_read_ok"] = int(host.view(np.int32)[0]) == 1000 + peer
        sb.barrier()
        # flag ping-pong latency between rank 0 and rank 1 (release/acquire at sys scope)
        if rank < 2:
            sb.local.zero_()
        sb.barrier()
        if rank < 2:
            ns = torch.zeros(1, dtype=torch.int64, device=dev)
            iters = 2000
            other = 1 - rank
            kernels().afb_p2p_pingpong(C.c_void_p(sb.local.data_ptr() + 64), C.c_void_p(sb.ptrs[other] + 64), iters,
                                       1 if rank == 0 else 0, C.c_void_p(ns.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            out["pingpong_rtt_us"] = float(ns.item()) / iters / 1e3
        sb.barrier()

    # ---- 1. BSP fused (all-gather over NVLink + replicated apply) vs the oracle ---------------
    for algo, loss, gamma, labels01 in (("sgd_sync", "logistic", 0.8, True), ("saga_sync", "least_squares", 0.05, False)):
        parts = make_parts(world, n_per, d, 7, labels01)
        cfg = AsyncConfig(algorithm=algo, num_features=d, num_rows=n_per * world, num_partitions=world, num_iterations=20,
                          step_size=gamma, batch_rate=0.3, loss=loss, seed=42, backend="fused")
        eng = FusedEngine(cfg, [shard_of(parts[rank], d, dev, rank)], n_per * world, device=dev)
        eng.run_rounds(20)
        eng.synchronize()
        eng.barrier()
        w = eng.weights().double().cpu().numpy()
        orc = AsyncOracle(parts, loss, 0.3, gamma, 10**9, 42)
        ref = (run_sgd_sync_oracle if algo == "sgd_sync" else run_saga_sync_oracle)(orc, 20)
        err = float(np.abs(w - ref.w).max() / max(1e-12, np.abs(ref.w).max()))
        # replicas must be bit-identical across ranks
        wt = eng.weights().clone()
        lst = [torch.empty_like(wt) for _ in range(world)]
        if world > 1:
            dist.all_gather(lst, wt)
        same = all(bool(torch.equal(lst[0], t)) for t in lst) if world > 1 else True
        out[f"{algo}_relerr"] = err
        out[f"{algo}_replicas_identical"] = same
        out[f"{algo}_nvls"] = bool(eng.kernels[0].nvls_bsp)          # all-gather through the NVLS multicast alias (multimem.st)
        out[f"{algo}_variant"] = eng.kernels[0].variant_effective
        out[f"{algo}_launches"] = eng.launches                        # 20 rounds = ONE persistent launch
        eng.check_health()
        del eng

    # ---- 2. async ASGD / ASAGA over NVLink: invariants + convergence ---------------------------
    for algo, loss, gamma, labels01, taw in (("asgd", "logistic", 0.8, True, 1 << 31), ("asaga", "least_squares", 0.05, False, 1 << 31),
                                             ("asgd", "logistic", 0.8, True, 0)):
        parts = make_parts(world, n_per, d, 11, labels01)
        iters = (60 if taw else 15) * world      # taw = 0 rejects most results at 8 workers: keep the target reachable
        cfg = AsyncConfig(algorithm=algo, num_features=d, num_rows=n_per * world, num_partitions=world, num_iterations=iters,
                          step_size=gamma, batch_rate=0.3, loss=loss, seed=42, backend="fused", taw=taw if taw < (1 << 31) else (1 << 31) - 1,
                          barrier="asp")
        eng = FusedEngine(cfg, [shard_of(parts[rank], d, dev, rank)], n_per * world, device=dev)
        losses = []
        nsteps = 0
        for _ in range(400):
            nsteps += 1
            r = eng.step()
            if r["accepted"] >= 0 and r["batch"] > 0:
                losses.append(r["loss_sum"] / r["batch"])
            if eng.server_state()["k"] >= iters:
                break
        eng.run_rounds(3)            # extra rounds after the limit must be no-ops
        eng.synchronize()
        eng.barrier()
        st = eng.server_state()
        w = eng.weights().double().cpu().numpy()
        key = f"{algo}_taw{0 if taw == 0 else 'inf'}"
        ns = torch.tensor([float(nsteps)], device=dev)
        lst_ns = [torch.zeros_like(ns) for _ in range(world)]
        if world > 1:
            dist.all_gather(lst_ns, ns)
        out[key + "_steps_per_rank"] = [int(t.item()) for t in lst_ns] if world > 1 else [nsteps]
        out[key + "_k"] = st["k"]
        out[key + "_clock"] = st["clock"]
        out[key + "_avail"] = st["avail_count"]
        out[key + "_finite"] = bool(np.isfinite(w).all())
        out[key + "_loss_first"] = losses[0] if losses else None
        out[key + "_loss_last"] = losses[-1] if losses else None
        if rank == 0:
            res = eng.drain_results()
            out[key + "_ring"] = len(res)
            out[key + "_accepted"] = sum(1 for r in res if r.data["accepted"])
            out[key + "_max_staleness"] = max((r.staleness for r in res), default=0)
            out[key + "_stale_applied"] = sum(1 for r in res if r.data["accepted"] and taw == 0 and r.staleness > 0)
            out[key + "_workers_seen"] = sorted({r.worker_id for r in res})
        eng.barrier()
        del eng

    # ---- 2b. persistent worker loops over NVLink: ONE launch = R rounds per worker ---------------------------------
    #   pipelined ASGD (asp), serial persistent ASGD behind a bucket gate, SSP gate, ASAGA, and a host-evaluated predicate
    #   that masks the last worker out (ASYNCbarrier on the fused path)
    R = 40
    for name, algo, loss, gamma, labels01, barrier, extra in (
            ("pipelined", "asgd", "logistic", 0.8, True, "asp", {}),
            ("bucket", "asgd", "logistic", 0.8, True, "bucket", {"bucket_ratio": 1.0}),    # gate = P: a round starts only while every worker is idle
            ("ssp", "asgd", "logistic", 0.8, True, "ssp", {"ssp_bound": 1 << 20}),
            ("asaga_persist", "asaga", "least_squares", 0.05, False, "asp", {}),
            ("masked", "asgd", "logistic", 0.8, True, "asp", {"mask": True})):
        parts = make_parts(world, n_per, d, 13, labels01)
        kw = {k: v for k, v in extra.items() if k != "mask"}
        cfg = AsyncConfig(algorithm=algo, num_features=d, num_rows=n_per * world, num_partitions=world, num_iterations=1 << 30,
                          step_size=gamma / world, batch_rate=0.3, loss=loss, seed=42, backend="fused", barrier=barrier, **kw)
        eng = FusedEngine(cfg, [shard_of(parts[rank], d, dev, rank)], n_per * world, device=dev, worker_mask=bool(extra.get("mask")))
        if extra.get("mask"):
            sel = eng.async_barrier(lambda st: True)                 # everybody ...
            if rank == 0:
                eng.set_worker_mask([q != world - 1 for q in range(world)])     # ... except the last worker
            eng.barrier()
        eng.run_rounds(R)
        eng.synchronize()
        eng.check_health()
        eng.barrier()
        st = eng.server_state()
        w = eng.weights().double().cpu().numpy()
        key = "p_" + name
        out[key + "_variant"] = eng.kernels[0].variant_effective
        out[key + "_launches"] = eng.launches
        out[key + "_k"] = st["k"]
        out[key + "_clock"] = st["clock"]
        out[key + "_avail"] = st["avail_count"]
        out[key + "_finite"] = bool(np.isfinite(w).all())
        ctl = eng.kernels[0].read_ctl()
        sk = torch.tensor([float(ctl["skipped_rounds"])], device=dev)
        lst_sk = [torch.zeros_like(sk) for _ in range(world)]
        if world > 1:
            dist.all_gather(lst_sk, sk)
        out[key + "_skipped"] = [int(t.item()) for t in lst_sk] if world > 1 else [int(sk.item())]
        if rank == 0:
            res = eng.drain_results()
            out[key + "_ring"] = len(res)
            out[key + "_workers_seen"] = sorted({r.worker_id for r in res})
            by_seq = [r.data["loss_sum"] / max(r.batch_size, 1) for r in res if r.batch_size > 0]
            out[key + "_loss_first"] = float(np.mean(by_seq[: world])) if by_seq else None
            out[key + "_loss_last"] = float(np.mean(by_seq[-world:])) if by_seq else None
            out[key + "_max_staleness"] = max((r.staleness for r in res), default=0)
        eng.barrier()
        del eng

    # ---- 3. versioned broadcast (AsyncBroadcast ring in symmetric memory, NVLS multicast when available) ---
    parts = make_parts(world, 256, d, 3, True)
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=256 * world, num_partitions=world, num_iterations=1 << 30,
                      step_size=0.5, batch_rate=0.5, loss="logistic", seed=1, backend="fused", barrier="asp", history_slots=3)
    eng = FusedEngine(cfg, [shard_of(parts[rank], d, dev, rank)], 256 * world, device=dev)
    vals = []
    for i in range(4):
        eng.run_rounds(2)
        eng.synchronize()
        eng.barrier()
        v = eng.publish_version()
        w_now = eng.weights().clone()
        if world > 1:
            dist.broadcast(w_now, 0)
        vals.append((v, w_now))
    ok = eng.version == 4
    for v, wv in vals[1:]:
        ok = ok and bool(torch.allclose(eng.version_value(v), wv))
    try:
        eng.version_value(0)
        ok = False
    except KeyError:
        pass
    flags = torch.tensor([1.0 if ok else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    out["versioned_broadcast_ok"] = bool(flags.item() == 1.0)
    out["versioned_broadcast_multicast"] = bool(eng.symm.multicast_ptr)
    eng.barrier()
    del eng

    if rank == 0:
        print("MG_RESULT " + json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
