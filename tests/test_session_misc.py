import json

import numpy as np
import pytest
import torch

from asyncframework_b200 import AsyncSession, Vectors
from asyncframework_b200.linalg import converters
from asyncframework_b200.utils.metrics import ConsoleSink, CsvSink, JsonlSink, MetricsRegistry, SchedulerSource


def test_session_master_urls_and_ops(tmp_path):
    with pytest.raises(ValueError):
        AsyncSession("mesos://host:5050")
    from asyncframework_b200.rpc import RpcError
    with pytest.raises(RpcError, match="cannot connect"):      # standalone master URL is understood; nobody listens there
        AsyncSession("spark://127.0.0.1:1")
    with AsyncSession("local[3,2]") as sc:
        assert sc.default_parallelism == 3 and sc.scheduler.max_failures == 2
        assert sc.parallelize(range(10)).map(lambda x: x * x).reduce(lambda a, b: a + b) == 285
        p = tmp_path / "t.txt"
        p.write_text("a b\nc\n\nd e f\n")
        assert sc.text_file(str(p), 2).flat_map(lambda l: l.split()).count() == 6
        (tmp_path / "d.libsvm").write_text("1 1:2.0 3:1.0\n0 2:5.0\n")
        ds = sc.load_libsvm_file(str(tmp_path / "d.libsvm"))
        assert ds.count() == 2 and ds.num_features == 3
        ctx = sc.async_context()
        sc.set_mode(1)
        assert sc.mode == 1 and sc.broadcast([1, 2]).value == [1, 2]
        ab = sc.async_broadcast(ctx, torch.ones(2))
        assert ab.value.tolist() == [1.0, 1.0]
        sc.set_mode(0)
    assert AsyncSession("local").default_parallelism == 1


def test_converters_roundtrip():
    sv = Vectors.sparse(5, [1, 3], [2.0, -1.0])
    t = converters.to_torch(sv)
    assert t.tolist() == [0, 2, 0, -1, 0] and np.array_equal(converters.to_numpy(sv), t.numpy())
    assert converters.from_torch(t, sparse=True) == sv and converters.to_sparse([0.0, 1.0]).indices.tolist() == [1]
    assert converters.to_dense(sv).values.tolist() == [0, 2, 0, -1, 0]


def test_metrics_registry_and_sinks(tmp_path, scheduler):
    from asyncframework_b200.data import ShardedDataset
    lines = []
    reg = MetricsRegistry().add_sink(ConsoleSink(lines.append)).add_sink(CsvSink(str(tmp_path / "csv"))).add_sink(
        JsonlSink(str(tmp_path / "m.jsonl")))
    SchedulerSource(reg, scheduler)
    reg.gauge("answer", lambda: 42)
    ShardedDataset.from_records(scheduler, list(range(8)), 4).count()
    snap = reg.report()
    assert snap["scheduler.jobs"] == 1 and snap["scheduler.tasks"] == 4 and snap["answer"] == 42
    assert snap["scheduler.task_run_ms.count"] == 4
    assert any("scheduler.tasks" in l for l in lines)
    assert (tmp_path / "csv" / "scheduler.tasks.csv").read_text().splitlines()[1].endswith("4.0")
    assert json.loads((tmp_path / "m.jsonl").read_text())["answer"] == 42


def test_launcher_runs_driver_inline(capsys):
    from asyncframework_b200 import launch
    rc = launch.main(["--gpus", "1", "sgd_sync", "x", "x", "10", "400", "2", "15", "1.0", "0.3", "5", "0", "42",
                      "--synthetic=linear", "--backend=cpu"])
    assert rc == 0


def test_compat_layer_and_transcribed_example():
    import importlib.util
    import os
    from asyncframework_b200 import compat
    assert issubclass(compat.ASYNCcontext, compat._AsyncContext) and hasattr(compat.RDD, "ASYNCreduce")
    for name in ("gettaskResult", "getStaleness", "getbatchSize", "getWorkerID"):
        assert hasattr(compat.RDDPartialRes, name)
    for name in ("getStaleness", "getAvailability", "getNumTasks", "getAvailableWorkers", "getMaxStaleness"):
        assert hasattr(compat.workerState, name)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("spark_asgd_sync", os.path.join(here, "examples", "spark_asgd_sync.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(numIteration=25) < 0.1            # record-level closures, the reference's own spelling


def test_mlutils_k_fold_and_text_loaders(tmp_path):
    """MLUtils.kFold: the k validation sets partition the data, each training set is the complement; loadVectors /
    loadLabeledPoints read what saveAsTextFile wrote (MLUtilsSuite "kFold", "loadVectors", "loadLabeledPoints")."""
    from asyncframework_b200 import AsyncSession
    from asyncframework_b200.data.libsvm import MLUtils
    from asyncframework_b200.linalg.vectors import LabeledPoint, Vectors
    with AsyncSession("local[2]") as sc:
        data = sc.parallelize(range(1, 101), 3)
        for k in (2, 5, 10):
            folds = MLUtils.kFold(data, k, 42)
            assert len(folds) == k
            seen = []
            for train, val in folds:
                v, t = val.collect(), train.collect()
                assert sorted(v + t) == list(range(1, 101)) and not set(v) & set(t)
                assert 0 < len(v) < 100 if k < 10 else len(v) < 100
                seen += v
            assert sorted(seen) == list(range(1, 101))                    # validation sets: a partition of the data
            assert [val.collect() for _, val in MLUtils.k_fold(data, k, 42)] == [val.collect() for _, val in folds]   # seeded
        vecs = [Vectors.dense(1.0, 2.0), Vectors.sparse(4, [0, 3], [1.5, -2.0])]
        sc.parallelize(vecs, 2).save_as_text_file(str(tmp_path / "vecs"))
        assert sorted(map(str, MLUtils.loadVectors(sc, str(tmp_path / "vecs")).collect())) == sorted(map(str, vecs))
        pts = [LabeledPoint(1.0, vecs[0]), LabeledPoint(0.0, vecs[1])]
        sc.parallelize(pts, 2).save_as_text_file(str(tmp_path / "pts"))
        back = sorted(MLUtils.loadLabeledPoints(sc, str(tmp_path / "pts")).collect(), key=lambda lp: lp.label)
        assert [(lp.label, str(lp.features)) for lp in back] == [(0.0, str(vecs[1])), (1.0, str(vecs[0]))]


def test_svm_data_generator_cli_writes_loadable_labeled_points(tmp_path):
    """SVMDataGenerator: separable-with-noise data; its command line writes text that loadLabeledPoints reads back and an SVM fits."""
    from asyncframework_b200 import AsyncSession
    from asyncframework_b200.data.generators import SVMDataGenerator
    from asyncframework_b200.data.libsvm import MLUtils
    X, y, w = SVMDataGenerator.generate_svm_arrays(2000, 3, seed=1)
    agree = ((X @ w >= 0) == (y > 0.5)).mean()
    assert agree > 0.93 and 0.3 < y.mean() < 0.7                     # labels follow the true hyperplane up to the 0.1 noise
    out = tmp_path / "svm"
    assert SVMDataGenerator.main(["local[2]", str(out), "300", "4", "3"]) == 0
    assert SVMDataGenerator.main(["local[2]"]) == 1
    with AsyncSession("local[2]") as sc:
        pts = MLUtils.load_labeled_points(sc, str(out)).collect()
        assert len(pts) == 300 and all(p.features.size == 4 and p.label in (0.0, 1.0) for p in pts)
