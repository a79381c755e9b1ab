"""MLlib mini-batch SGD baseline — the ``SparkSGDMLLIB`` driver (Appendix A.5).

Reference: SparkSGDMLLIB.scala:13-137: ``LinearRegressionWithSGD.train(points, iters, step, b)``
(LeastSquaresGradient, SimpleUpdater step/√i, true sampled count, seed 42+i, convergence tol 0.001),
then objective-vs-time from the weight history ``LinearRegressionWithSGD.h`` (snapshots at i == 1
and every 100th iteration).

CLI (8 positional): <path> <file> <d> <N> <P> <iters> <step> <b>
"""
from __future__ import annotations

from ..config import AsyncConfig
from ..models.glm import LinearRegressionWithSGD, LogisticRegressionWithSGD
from .common import compute_dtype, load_points, make_scheduler, now_ms, resolve_device, run_cli
from .telemetry import MetricsLog, RunReport, evaluate_objective


def run(cfg: AsyncConfig, points=None, scheduler=None) -> RunReport:
    device = resolve_device(cfg)
    dtype = compute_dtype(cfg, device)
    own = scheduler is None
    sched = scheduler or make_scheduler(cfg, device)
    metrics = MetricsLog(cfg.metrics_path)
    try:
        if points is None:
            points = load_points(cfg, sched, device, dtype)
        points = points.cache()
        start_wall_ms = __import__("time").time() * 1e3
        start = now_ms()
        if cfg.loss == "logistic":
            model = LogisticRegressionWithSGD.train(points, cfg.num_iterations, cfg.step_size, cfg.batch_rate,
                                                    reg_param=cfg.reg_param)
        else:
            model = LinearRegressionWithSGD.train(points, cfg.num_iterations, cfg.step_size, cfg.batch_rate)
        elapsed = now_ms() - start
        hist = getattr(model, "weight_history", LinearRegressionWithSGD.h)
        snaps = [(float(t - start_wall_ms), w) for t, w in hist] + [(elapsed, model.weights)]
        rep = RunReport("sgd_mllib", cfg.num_partitions, iterations=len(getattr(model, "loss_history", [])),
                        elapsed_ms=elapsed, snapshots=snaps, final_weights=model.weights)
        rep.extra["loss_history"] = list(getattr(model, "loss_history", []))
        if cfg.eval_objective:
            rep.objective = evaluate_objective(points, snaps, "least_squares" if cfg.loss != "logistic" else "logistic")
        metrics.emit("report", **{k: v for k, v in rep.to_dict().items() if k != "extra"})
        return rep
    finally:
        metrics.close()
        if own:
            sched.stop()


def main(argv=None):
    return run_cli("sgd_mllib", run, argv)


if __name__ == "__main__":
    main()
