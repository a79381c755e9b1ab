"""Asynchronous SGD — the ``SparkASGDThread`` driver.

Reference: ASYNCsamples/src/main/scala/org/apache/async/SparkASGDThread.scala:20-438.
Two threads share the model: the **submitter** (main) gates on ``available >= floor(P*beta)``
(:230-234), snapshots/broadcasts ``w`` (:242), selects participants with ``ASYNCbarrier`` (:282-283),
injects delays (:284-309), samples with ``seed+k+1`` (:311), maps the gradient (:316-318) and calls
``ASYNCreduce`` (:324); the **updater** drains the context and applies
``w += (-gamma/sqrt(floor(k/P)+1)) * g/(b*N/P)`` when ``staleness <= taw`` (:150-223).

CLI (13 positional args, same order as the reference):
    python -m asyncframework_b200.algo.asgd <path> <file> <d> <N> <P> <iters> <gamma> <taw> <b> <beta> <printerFreq> <coeff> <seed>
plus ``--backend={cpu,fused}``, ``--loss=...``, ``--synthetic=...`` etc.
"""
from __future__ import annotations

import math
import threading
import time
from typing import Optional

import torch

from ..config import AsyncConfig
from ..core import barriers
from ..core.context import AsyncContext
from .common import (compute_dtype, load_points, make_scheduler, now_ms, partition_gradient, resolve_device, run_cli, say, start_watchdog, HostCheckpointer)
from .delay import DelayInjector
from .telemetry import MetricsLog, RunReport, WaitTimer, evaluate_objective
from ..optim.gradient import make_gradient


def run(cfg: AsyncConfig, points=None, scheduler=None) -> RunReport:
    if cfg.backend == "fused":
        from .fused_driver import run_fused
        return run_fused(cfg, points)
    if cfg.backend == "nccl":                      # two-sided collectives cannot express the asynchronous protocol: clear error
        from .nccl_driver import run_nccl
        return run_nccl(cfg, points)
    device = resolve_device(cfg)
    dtype = compute_dtype(cfg, device)
    own_sched = scheduler is None
    sched = scheduler or make_scheduler(cfg, device)
    metrics = MetricsLog(cfg.metrics_path)
    try:
        if points is None:
            points = load_points(cfg, sched, device, dtype)
        P = cfg.num_partitions
        N = points.count()
        d = next(s for s in points.shards() if s is not None).num_features
        say(cfg, dict(enumerate(points.partition_sizes())))           # "check how balanced the data is" (:77)
        gradient = make_gradient(cfg.loss)
        w = torch.zeros(d, dtype=dtype, device=device)
        AC: AsyncContext = AsyncContext(P, sched)
        watchdog = start_watchdog(cfg, AC, metrics)
        AC.set_record_stat(bool(cfg.extra.get("record_stat", True)))
        points_indexed = points.zip_with_index().cache()
        state = {"k": 0, "dropped": 0}
        ckpt = HostCheckpointer(cfg)
        ck = ckpt.resume(device, dtype)
        if ck is not None:
            w.copy_(ck["w"])
            state["k"] = int(ck["k"])
        opt_vars = [(0.0, w.detach().clone())]
        waits = WaitTimer()
        delay = DelayInjector(P, cfg.coeff, cfg.warmup_updates_per_worker * P)
        b, gamma, taw, iters = cfg.batch_rate, cfg.step_size, cfg.taw, cfg.num_iterations
        par_recs = b * N / P
        stal_hist = {}
        start = now_ms()
        wlock = threading.Lock() if cfg.extra.get("lock_w", False) else None   # Hogwild by default (:189 vs :242)

        def updater():
            try:
                _updater()
            except BaseException as exc:  # noqa: BLE001 — surfaced by the submitter loop instead of spinning on state['k'] forever
                state["error"] = exc

        def _updater():
            from .common import host_math_threads
            host_math_threads.pin()                        # small-vector math: one intra-op thread for this thread
            while state["k"] < iters:
                bsize = AC.size()
                if bsize == 0:
                    time.sleep(0.001)
                    continue
                for _ in range(bsize):
                    info = AC.collect_all(timeout=1.0)
                    if info is None:
                        break
                    k = state["k"]
                    if k >= iters:
                        break
                    stal_hist[info.staleness] = stal_hist.get(info.staleness, 0) + 1
                    now = now_ms()
                    if info.staleness <= taw:
                        waits.on_finish(info.worker_id, now)
                        delay.observe(k, waits.task_ms(info.worker_id, now))
                        g = info.data
                        eta = -gamma / math.sqrt(k // P + 1)                  # integer division (quirk B.8)
                        if wlock:
                            with wlock:
                                w.add_(g.to(w.dtype), alpha=eta / par_recs)
                        else:
                            w.add_(g.to(w.dtype), alpha=eta / par_recs)
                        if k % cfg.printer_freq == 0:
                            say(cfg, f"Iteration {k} is finished")
                            opt_vars.append((now_ms() - start, w.detach().clone()))
                            metrics.emit("snapshot", k=k, elapsed_ms=now_ms() - start)
                        state["k"] = k + 1
                        ckpt.maybe_save(k + 1, lambda: dict(w=w.detach().clone()))
                    else:
                        state["dropped"] += 1                                   # drop; worker is re-queued by STAT
        th = threading.Thread(target=updater, name="asgd-updater", daemon=True)
        th.start()

        need = math.floor(P * cfg.bucket_ratio)
        pred = barriers.bucket(P, cfg.bucket_ratio)
        last_k, resubmits, submitted = -1, 0, 0
        while state["k"] < iters:
            if state.get("error") is not None:
                raise RuntimeError(f"asgd updater thread failed: {state['error']!r}") from state["error"]
            if AC.lost and len(AC.lost) >= cfg.extra.get("max_lost_tasks", 4 * P):
                AC.raise_if_failed()                       # every retry of many tasks failed: stop instead of resubmitting forever
            k = state["k"]
            init_workers = P if submitted == 0 else AC.stat.available_workers()   # STAT is empty before the first submission
            if init_workers >= need:
                a = AC.broadcast(w)                                             # snapshot of w NOW (:242)
                delay.maybe_calibrate(k)
                filtered = points_indexed.async_barrier(pred, AC.stat)
                workers = list(filtered.selected_workers)
                if not workers:
                    time.sleep(0.001)
                    continue
                delayed = filtered.map_partitions_with_index(delay.hook(), preserves_batch=True)
                # seed = seed + k + 1 as in the reference; if k has not advanced since the last submission
                # (every sampled partition came back empty) perturb it, otherwise the identical empty sample
                # would be redrawn forever (tiny b*N/P only)
                resubmits = resubmits + 1 if k == last_k else 0
                last_k = k
                sampled = delayed.sample(False, b, cfg.seed + k + 1 + 7919 * resubmits)
                index_grad = sampled.map_batch(lambda s, a=a: partition_gradient(gradient, s, a.value)[0])
                index_grad.async_reduce(lambda x, y: x.add_(y), AC)
                waits.on_submit(workers, now_ms())
                submitted += 1
            else:
                time.sleep(0.001)
        th.join(timeout=10)
        elapsed = now_ms() - start
        AC.set_mode(0)
        rep = RunReport("asgd", P, iterations=state["k"], elapsed_ms=elapsed, snapshots=opt_vars,
                        waiting_ms=dict(waits.waiting), accepted=state["k"], dropped=state["dropped"],
                        staleness_hist=stal_hist, final_weights=w.detach().clone())
        if waits.waiting and state["k"]:
            rep.avg_waiting_ms = sum(waits.waiting.values()) / (len(waits.waiting) * state["k"])
        rep.extra["avg_delay_ms"] = delay.avg_delay_ms
        rep.extra["injected_delay_ms"] = delay.injected_ms
        if cfg.eval_objective:
            rep.objective = evaluate_objective(points, opt_vars + [(elapsed, w.detach().clone())], cfg.loss)
        rep.extra["dead_workers"] = list(watchdog.dead)
        metrics.emit("report", **rep.to_dict())
        return rep
    finally:
        if 'ckpt' in locals():
            ckpt.close()
        if 'watchdog' in locals():
            watchdog.stop()
        metrics.close()
        if own_sched:
            sched.stop()


def main(argv=None):
    return run_cli("asgd", run, argv)


if __name__ == "__main__":
    main()
