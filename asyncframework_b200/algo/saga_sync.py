"""Synchronous SAGA — the ``SparkASAGASync`` driver (Appendix A.4).

Reference: SparkASAGASync.scala:17-443.  Each iteration: broadcast; every worker computes
``g_p = Σ (r_i − hist_i) x_i`` over its Bernoulli(b) sample; wait for P results merging the history
updates and ``G = Σ g_p`` (:258-292); ``w <- w − γ/(bN)·G − γ·ᾱ``; ``ᾱ <- ᾱ + G/N`` (:295-299).

CLI (11 positional): <path> <file> <d> <N> <P> <iters> <gamma> <b> <printerFreq> <coeff> <seed>
"""
from __future__ import annotations

import torch

from ..config import AsyncConfig
from ..core import barriers
from ..core.context import AsyncContext
from ..optim.gradient import make_gradient
from .asaga import saga_partition
from .common import compute_dtype, load_points, make_scheduler, now_ms, resolve_device, run_cli, say, start_watchdog, WorkerLost, HostCheckpointer
from .delay import DelayInjector
from .telemetry import MetricsLog, RunReport, WaitTimer, evaluate_objective


def run(cfg: AsyncConfig, points=None, scheduler=None) -> RunReport:
    if cfg.backend == "fused":
        from .fused_driver import run_fused
        return run_fused(cfg, points)
    if cfg.backend == "nccl":
        from .nccl_driver import run_nccl
        return run_nccl(cfg, points)
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
        sizes = points.partition_sizes()
        d = next(s for s in points.shards() if s is not None).num_features
        say(cfg, dict(enumerate(sizes)))
        gradient = make_gradient(cfg.loss)
        w = torch.zeros(d, dtype=dtype, device=device)
        alpha_bar = torch.zeros_like(w)
        hist = [torch.zeros(n, dtype=dtype, device=device) for n in sizes]
        starts = [sum(sizes[:p]) for p in range(P)]
        AC: AsyncContext = AsyncContext(P, sched)
        watchdog = start_watchdog(cfg, AC, metrics)
        AC.set_record_stat(True)
        points_indexed = points.zip_with_index().cache()
        ckpt = HostCheckpointer(cfg)
        ck = ckpt.resume(device, dtype)
        k0 = 0
        if ck is not None:
            w.copy_(ck["w"])
            if ck.get("alpha_bar") is not None:
                alpha_bar.copy_(ck["alpha_bar"])
            for p_, h_ in ck["history"].items():
                hist[p_].copy_(h_)
            k0 = int(ck["k"])
        opt_vars = [(0.0, w.detach().clone())]
        waits = WaitTimer()
        delay = DelayInjector(P, cfg.coeff, cfg.warmup_updates_per_worker)
        b, gamma = cfg.batch_rate, cfg.step_size
        start = now_ms()
        for k in range(k0, cfg.num_iterations):
            a = AC.broadcast(w)
            delay.maybe_calibrate(k)
            filtered = points_indexed.async_barrier(barriers.asp(), AC.stat)
            delayed = filtered.map_partitions_with_index(delay.hook(), preserves_batch=True)
            sampled = delayed.sample(False, b, cfg.seed + k + 1)

            def task(s, a=a):
                p = s.partition
                s.meta["rows"] = (s.gidx - starts[p]).to(torch.int64)
                return saga_partition(gradient, s, a.value, hist[p])
            sampled.map_batch(task).async_aggregate(
                lambda: None, lambda acc, r: r if acc is None else acc, lambda x, y: x if y is None else y, AC)
            waits.on_submit(range(P), now_ms())
            G = torch.zeros_like(w)
            expected, got = len(filtered.selected_workers), 0
            while got < expected:
                info = AC.collect_all(timeout=0.05)
                AC.raise_if_failed()         # a task that exhausted its retries fails the (conceptually blocking) BSP job
                if watchdog.dead:
                    raise WorkerLost(f"worker(s) {watchdog.dead} stopped responding in iteration {k}: a BSP run cannot continue")
                if info is None:
                    if AC.size() == 0 and AC.dropped_empty + got >= expected:
                        break
                    continue
                now = now_ms()
                waits.on_finish(info.worker_id, now)
                delay.observe(k, waits.task_ms(info.worker_id, now))
                rows, r_new, g, _ = info.data
                hist[info.worker_id][rows.to(device)] = r_new.to(dtype)
                G.add_(g.to(w.dtype))
                got += 1
            AC.dropped_empty = 0
            w.add_(G, alpha=-gamma / (b * N))
            w.add_(alpha_bar, alpha=-gamma)
            alpha_bar.add_(G, alpha=1.0 / N)
            ckpt.maybe_save(k + 1, lambda: dict(w=w.detach().clone(), alpha_bar=alpha_bar.detach().clone(),
                                                history={p_: h_.detach().clone() for p_, h_ in enumerate(hist)}))
            if k % cfg.printer_freq == 0:
                say(cfg, f"Iteration {k} is finished")
                opt_vars.append((now_ms() - start, w.detach().clone()))
                metrics.emit("snapshot", k=k, elapsed_ms=now_ms() - start)
        elapsed = now_ms() - start
        AC.set_mode(0)
        rep = RunReport("saga_sync", P, iterations=cfg.num_iterations, elapsed_ms=elapsed, snapshots=opt_vars,
                        waiting_ms=dict(waits.waiting), accepted=cfg.num_iterations, final_weights=w.detach().clone())
        if waits.waiting and cfg.num_iterations:
            rep.avg_waiting_ms = sum(waits.waiting.values()) / (len(waits.waiting) * cfg.num_iterations * P)
        if cfg.eval_objective:
            rep.objective = evaluate_objective(points, opt_vars + [(elapsed, w.detach().clone())], cfg.loss)
        metrics.emit("report", **rep.to_dict())
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
    return run_cli("saga_sync", run, argv)


if __name__ == "__main__":
    main()
