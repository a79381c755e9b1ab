"""Multi-process-on-one-box tests (the `local-cluster` analogue): 2 CPU workers over gloo.
Includes BASELINE config 1: SparkASGDSync logistic regression on 1k x 20 synthetic LIBSVM."""
import numpy as np
import pytest
import torch

from asyncframework_b200.config import AsyncConfig
from asyncframework_b200.parallel.cpu_backend import spawn_local


def _libsvm_logistic(tmp_path, n=1000, d=20):
    from asyncframework_b200.core import Scheduler
    from asyncframework_b200.data import LogisticRegressionDataGenerator, save_as_libsvm_file
    s = Scheduler(1)
    try:
        ds = LogisticRegressionDataGenerator.generate_logistic_dataset(s, n, d, 1.0, 1, seed=3)
        save_as_libsvm_file(ds, str(tmp_path / "data"))
    finally:
        s.stop()
    return str(tmp_path) + "/", "data"


@pytest.mark.timeout(180)
def test_baseline_config1_sync_logistic_two_cpu_workers(tmp_path):
    path, fname = _libsvm_logistic(tmp_path)
    cfg = AsyncConfig(algorithm="sgd_sync", path=path, file=fname, num_features=20, num_rows=1000, num_partitions=2,
                      num_iterations=80, step_size=2.0, batch_rate=0.3, printer_freq=20, loss="logistic", backend="gloo",
                      layout="dense")
    out = spawn_local(cfg, world=2, mode="sync")
    losses = out["loss_history"]
    assert len(losses) == 80 and losses[-1] < 0.6 * losses[0]
    assert out["objective"][0][1] < 0.45                                  # mean logistic loss at the final model
    # equals the single-process oracle on the same partitioning
    from asyncframework_b200.algo.oracle import AsyncOracle, run_sgd_sync_oracle
    from asyncframework_b200.core import Scheduler
    from asyncframework_b200.data import load_libsvm_file
    s = Scheduler(1)
    try:
        ds = load_libsvm_file(s, path + fname, 20, 2, layout="dense").repartition(2).zip_with_index()
        parts = [(sh.X[:, :20].numpy(), sh.y.numpy(), sh.gidx.numpy()) for sh in ds.shards()]
    finally:
        s.stop()
    ref = run_sgd_sync_oracle(AsyncOracle(parts, "logistic", 0.3, 2.0, 10**9, 42), 80)
    assert np.allclose(np.asarray(out["w"]), ref.w, rtol=1e-8, atol=1e-10)


@pytest.mark.timeout(180)
@pytest.mark.parametrize("alg", ["asgd", "asaga"])
def test_async_parameter_server_over_gloo(alg):
    cfg = AsyncConfig(algorithm=alg, synthetic="linear", num_rows=1200, num_features=10, num_partitions=3,
                      num_iterations=300, step_size=1.0 if alg == "asgd" else 0.05, batch_rate=0.2, backend="gloo")
    out = spawn_local(cfg, world=3, mode="async")
    assert out["iterations"] >= 300 and out["objective"][0][1] < 0.05
