"""Synchronous mini-batch SGD through the ASYNC API — the ``SparkASGDSync`` driver (Appendix A.3).

Reference: SparkASGDSync.scala:19-411.  Per iteration: broadcast ``w`` (:156), delay (:157-193),
Bernoulli(b) sample with ``seed+k+1``, ``ASYNCreduce`` (:211), block on ``ASYNCcollectAll`` until P
results (:234-266), ``w += (-gamma/sqrt(k+1)) * Σg/(b*N)`` (:268-272), snapshot every printerFreq.

CLI (11 positional): <path> <file> <d> <N> <P> <iters> <gamma> <b> <printerFreq> <coeff> <seed>
"""
from __future__ import annotations

import math

import torch

from ..config import AsyncConfig
from ..core import barriers
from ..core.context import AsyncContext
from ..optim.gradient import make_gradient
from .common import compute_dtype, load_points, make_scheduler, now_ms, partition_gradient, resolve_device, run_cli, say, start_watchdog, WorkerLost, HostCheckpointer
from .delay import DelayInjector
from .telemetry import MetricsLog, RunReport, WaitTimer, evaluate_objective


def run(cfg: AsyncConfig, points=None, scheduler=None) -> RunReport:
    if cfg.backend == "fused":
        from .fused_driver import run_fused
        return run_fused(cfg, points)
    if cfg.backend == "nccl":
        from .nccl_driver import run_nccl
        return run_nccl(cfg, points)
    if cfg.backend == "gloo":
        from ..parallel.cpu_backend import run_sync_gloo
        return run_sync_gloo(cfg)
    device = resolve_device(cfg)
    dtype = compute_dtype(cfg, device)
    own = scheduler is None
    sched = scheduler or make_scheduler(cfg, device)
    metrics = MetricsLog(cfg.metrics_path)
    try:
        if points is None:
            points = load_points(cfg, sched, device, dtype)
        P = cfg.num_partitions
        N = points.count()
        d = next(s for s in points.shards() if s is not None).num_features
        say(cfg, dict(enumerate(points.partition_sizes())))
        gradient = make_gradient(cfg.loss)
        w = torch.zeros(d, dtype=dtype, device=device)
        AC: AsyncContext = AsyncContext(P, sched)
        watchdog = start_watchdog(cfg, AC, metrics)
        AC.set_record_stat(True)
        points_indexed = points.zip_with_index().cache()
        ckpt = HostCheckpointer(cfg)
        ck = ckpt.resume(device, dtype)
        k0 = 0
        if ck is not None:
            w.copy_(ck["w"])
            k0 = int(ck["k"])
        opt_vars = [(0.0, w.detach().clone())]
        waits = WaitTimer()
        delay = DelayInjector(P, cfg.coeff, cfg.warmup_updates_per_worker)
        b, gamma = cfg.batch_rate, cfg.step_size
        start = now_ms()
        losses = []
        for k in range(k0, cfg.num_iterations):
            a = AC.broadcast(w)
            delay.maybe_calibrate(k)
            filtered = points_indexed.async_barrier(barriers.asp(), AC.stat)
            delayed = filtered.map_partitions_with_index(delay.hook(), preserves_batch=True)
            sampled = delayed.sample(False, b, cfg.seed + k + 1)

            def grad_fn(s, a=a):
                g, l, n, _ = partition_gradient(gradient, s, a.value)
                return [g, float(l), int(n)]
            sampled.map_batch(grad_fn).async_reduce(lambda x, y: [x[0].add_(y[0]), x[1] + y[1], x[2] + y[2]], AC)
            waits.on_submit(range(P), now_ms())
            acc = torch.zeros_like(w)
            loss_sum, cnt = 0.0, 0
            expected = len(filtered.selected_workers)
            got = 0
            while got < expected:
                info = AC.collect_all(timeout=0.05)
                AC.raise_if_failed()         # a task that exhausted its retries fails the (conceptually blocking) BSP job
                if watchdog.dead:
                    raise WorkerLost(f"worker(s) {watchdog.dead} stopped responding in iteration {k}: a BSP run cannot continue")
                if info is None:
                    # an empty sampled partition yields no result (quirk B.5): don't wait forever
                    if AC.size() == 0 and AC.dropped_empty + got >= expected:
                        break
                    continue
                now = now_ms()
                waits.on_finish(info.worker_id, now)
                delay.observe(k, waits.task_ms(info.worker_id, now))
                acc.add_(info.data[0].to(w.dtype))
                loss_sum += info.data[1]
                cnt += info.data[2]
                got += 1
            AC.dropped_empty = 0
            w.add_(acc, alpha=-gamma / math.sqrt(k + 1) / (b * N))
            losses.append(loss_sum / max(cnt, 1))
            ckpt.maybe_save(k + 1, lambda: dict(w=w.detach().clone()))
            if k % cfg.printer_freq == 0:
                say(cfg, f"Iteration {k} is finished")
                opt_vars.append((now_ms() - start, w.detach().clone()))
                metrics.emit("snapshot", k=k, elapsed_ms=now_ms() - start, loss=losses[-1])
        elapsed = now_ms() - start
        AC.set_mode(0)
        rep = RunReport("sgd_sync", P, iterations=cfg.num_iterations, elapsed_ms=elapsed, snapshots=opt_vars,
                        waiting_ms=dict(waits.waiting), accepted=cfg.num_iterations, final_weights=w.detach().clone())
        if waits.waiting and cfg.num_iterations:
            rep.avg_waiting_ms = sum(waits.waiting.values()) / (len(waits.waiting) * cfg.num_iterations * P)
        rep.extra["loss_history"] = losses
        rep.extra["avg_delay_ms"] = delay.avg_delay_ms
        if cfg.eval_objective:
            rep.objective = evaluate_objective(points, opt_vars + [(elapsed, w.detach().clone())], cfg.loss)
        metrics.emit("report", **{k2: v for k2, v in rep.to_dict().items() if k2 != "extra"})
        return rep
    finally:
        if 'ckpt' in locals():
            ckpt.close()
        if 'watchdog' in locals():
            watchdog.stop()
        metrics.close()
        if own:
            sched.stop()


def main(argv=None):
    return run_cli("sgd_sync", run, argv)


if __name__ == "__main__":
    main()
