"""Asynchronous SAGA with per-sample history — the ``SparkASAGAThread`` driver (Appendix A.2).

Reference: SparkASAGAThread.scala:18-520.  State: ``w``, running mean of historical gradients
``alphaBar`` and the history table (``ScalarMap: sampleIdx -> last residual``, :114).  Per round the
driver broadcasts ``w`` with ``ASYNCbroadcast`` (:268), replays every selected worker's sampling
stream to ship only the needed history scalars (``sampledMap``, :280-294), workers emit
``L_p = [(i, r_i)]`` and ``g_p = Σ (r_i − hist_i) x_i`` (:373-386) through ``ASYNCaggregate``; the
updater accepts when ``k − ts <= taw`` (:184): ``hist[i] <- r_i`` (:199-201),
``w <- w − γ/(bN/P)·g_p − γ·ᾱ`` (ᾱ before its update, :210-212), ``ᾱ <- ᾱ + g_p/N`` (:213).
A discarded result does NOT touch the history (:224-231).

B200-first change: the history table is **sharded with the rows** (each worker owns its rows'
scalars) instead of living in a driver HashMap — the "partial history" transport disappears while
the math is identical; the accept/discard rule is preserved by committing a worker's new residuals
only when its result is accepted.
"""
from __future__ import annotations

import math
import threading
import time

import torch

from ..config import AsyncConfig
from ..core import barriers
from ..core.broadcast import AsyncBroadcast
from ..core.context import AsyncContext
from ..optim.gradient import make_gradient
from .common import compute_dtype, load_points, make_scheduler, now_ms, resolve_device, run_cli, say, start_watchdog, HostCheckpointer
from .delay import DelayInjector
from .telemetry import MetricsLog, RunReport, WaitTimer, evaluate_objective


def saga_partition(gradient, shard, w, hist_p):
    """Worker side of one SAGA task on an already-sampled shard.  ``hist_p`` is the worker's full
    history vector; ``shard.meta['rows']`` the sampled local positions.
    Returns (local rows, new residuals, g_p = Σ (r_i − hist_i) x_i, loss_sum)."""
    rows = shard.meta["rows"]
    h = hist_p[rows.to(hist_p.device)]
    g, loss, n, coeff = gradient.batch(shard, w, h.to(shard.device))
    return rows, coeff, g, float(loss)


def run(cfg: AsyncConfig, points=None, scheduler=None) -> RunReport:
    if cfg.backend == "fused":
        from .fused_driver import run_fused
        return run_fused(cfg, points)
    if cfg.backend == "nccl":                      # two-sided collectives cannot express the asynchronous protocol: clear error
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
        hist = [torch.zeros(n, dtype=dtype, device=device) for n in sizes]      # sharded ScalarMap
        AC: AsyncContext = AsyncContext(P, sched)
        watchdog = start_watchdog(cfg, AC, metrics)
        AC.set_record_stat(True)
        points_indexed = points.zip_with_index().cache()
        # remember each row's local position so sampled shards can address the history shard
        local_pos = {p: None for p in range(P)}
        state = {"k": 0, "dropped": 0}
        ckpt = HostCheckpointer(cfg)
        ck = ckpt.resume(device, dtype)
        if ck is not None:
            w.copy_(ck["w"])
            if ck.get("alpha_bar") is not None:
                alpha_bar.copy_(ck["alpha_bar"])
            for p_, h_ in ck["history"].items():
                hist[p_].copy_(h_)
            state["k"] = int(ck["k"])
        opt_vars = [(0.0, w.detach().clone())]
        waits = WaitTimer()
        delay = DelayInjector(P, cfg.coeff, cfg.warmup_updates_per_worker * P)
        b, gamma, taw, iters = cfg.batch_rate, cfg.step_size, cfg.taw, cfg.num_iterations
        par_recs = b * N / P
        stal_hist = {}
        starts = [sum(sizes[:p]) for p in range(P)]
        start = now_ms()

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
                    lag = k - info.timestamp                                   # ASAGA compares k − ts (quirk B.6)
                    stal_hist[lag] = stal_hist.get(lag, 0) + 1
                    now = now_ms()
                    if lag <= taw:
                        waits.on_finish(info.worker_id, now)
                        delay.observe(k, waits.task_ms(info.worker_id, now))
                        rows, r_new, g, _ = info.data
                        hist[info.worker_id][rows.to(device)] = r_new.to(dtype)                # K7
                        w.add_(g.to(w.dtype), alpha=-gamma / par_recs)                          # K6
                        w.add_(alpha_bar, alpha=-gamma)
                        alpha_bar.add_(g.to(w.dtype), alpha=1.0 / N)
                        if k % cfg.printer_freq == 0:
                            say(cfg, f"Iteration {k} is finished")
                            opt_vars.append((now_ms() - start, w.detach().clone()))
                            metrics.emit("snapshot", k=k, elapsed_ms=now_ms() - start)
                        state["k"] = k + 1
                        ckpt.maybe_save(k + 1, lambda: dict(w=w.detach().clone(), alpha_bar=alpha_bar.detach().clone(),
                                                            history={p_: h_.detach().clone() for p_, h_ in enumerate(hist)}))
                    else:
                        state["dropped"] += 1                                   # history untouched
        th = threading.Thread(target=updater, name="asaga-updater", daemon=True)
        th.start()

        need = math.floor(P * cfg.bucket_ratio)
        pred = barriers.bucket(P, cfg.bucket_ratio)
        last_k, resubmits, submitted = -1, 0, 0
        while state["k"] < iters:
            if state.get("error") is not None:
                raise RuntimeError(f"asaga updater thread failed: {state['error']!r}") from state["error"]
            if AC.lost and len(AC.lost) >= cfg.extra.get("max_lost_tasks", 4 * P):
                AC.raise_if_failed()                       # every retry of many tasks failed: stop instead of resubmitting forever
            k = state["k"]
            init_workers = P if submitted == 0 else AC.stat.available_workers()   # STAT is empty before the first submission
            if init_workers >= need:
                a = AsyncBroadcast(AC, w, slots=cfg.history_slots)                # versioned broadcast (:268)
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

                def task(s, a=a):
                    p = s.partition
                    s.meta["rows"] = (s.gidx - starts[p]).to(torch.int64)
                    return saga_partition(gradient, s, a.value, hist[p])
                sampled.map_batch(task).async_aggregate(
                    lambda: None, lambda acc, r: r if acc is None else acc, lambda x, y: x if y is None else y, AC)
                waits.on_submit(workers, now_ms())
                submitted += 1
            else:
                time.sleep(0.001)
        th.join(timeout=10)
        elapsed = now_ms() - start
        AC.set_mode(0)
        rep = RunReport("asaga", P, iterations=state["k"], elapsed_ms=elapsed, snapshots=opt_vars,
                        waiting_ms=dict(waits.waiting), accepted=state["k"], dropped=state["dropped"],
                        staleness_hist=stal_hist, final_weights=w.detach().clone())
        if waits.waiting and state["k"]:
            rep.avg_waiting_ms = sum(waits.waiting.values()) / (len(waits.waiting) * state["k"])
        rep.extra["alpha_bar_norm"] = float(alpha_bar.norm())
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
        if own:
            sched.stop()


def main(argv=None):
    return run_cli("asaga", run, argv)


if __name__ == "__main__":
    main()
