import pytest

from asyncframework_b200.config import AsyncConfig, parse_cli, usage


def test_async_positional_order_matches_reference():
    # README.md:46 quick-start line
    argv = "data/ mnist8m.scale 784 8100000 64 16000 1.5625e-3 20000000 0.01 0.7 200 -1 42".split()
    cfg = parse_cli("asgd", argv)
    assert (cfg.path, cfg.file, cfg.num_features, cfg.num_rows, cfg.num_partitions) == ("data/", "mnist8m.scale", 784, 8100000, 64)
    assert cfg.num_iterations == 16000 and cfg.step_size == pytest.approx(1.5625e-3) and cfg.taw == 20000000
    assert cfg.batch_rate == 0.01 and cfg.bucket_ratio == 0.7 and cfg.printer_freq == 200
    assert cfg.coeff == -1 and cfg.production_cluster and cfg.seed == 42
    assert cfg.data_path == "data/mnist8m.scale"


def test_sync_and_mllib_layouts():
    cfg = parse_cli("sgd_sync", "p/ mnist 780 60000 16 600 8e-7 0.1 6 0 42".split())
    assert cfg.batch_rate == 0.1 and cfg.printer_freq == 6 and cfg.coeff == 0 and cfg.seed == 42
    cfg = parse_cli("sgd_mllib", "p mnist 780 60000 8 2000 2e-7 0.1".split())
    assert cfg.data_path == "p/mnist" and cfg.num_iterations == 2000


def test_overrides_and_arity_check():
    cfg = parse_cli("asaga", "p/ f 10 100 4 50 0.1 5 0.2 0.8 10 0.3 1 --backend=cpu --loss logistic --history_slots=4".split())
    assert cfg.backend == "cpu" and cfg.loss == "logistic" and cfg.history_slots == 4 and cfg.coeff == 0.3
    with pytest.raises(SystemExit):
        parse_cli("asgd", ["only", "three", "args"])
    assert "13 positional" in usage("asgd")


def test_validation():
    with pytest.raises(ValueError):
        AsyncConfig(batch_rate=0.0).validate()
    with pytest.raises(ValueError):
        AsyncConfig(coeff=-0.5).validate()
    c = AsyncConfig.from_dict({"num_partitions": 4, "unknown_key": 3})
    assert c.extra["unknown_key"] == 3
