"""Drivers and the general API on a single GPU (fused backend with several logical workers, and the
in-process scheduler over HBM-resident shards)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sample_rows_kernel_matches_cpu_hash():
    from asyncframework_b200.core.sampler import bernoulli_mask
    from asyncframework_b200.ops.sampling import sample_rows
    g = (torch.arange(100003, dtype=torch.int64) * 3 + 17).cuda()
    for seed, frac in [(1, 0.1), (2**40 + 5, 0.37), (9, 1.0), (9, 0.0)]:
        rows = sample_rows(g.numel(), frac, seed, g)
        ref = np.nonzero(bernoulli_mask(seed, g.cpu(), frac))[0]
        assert np.array_equal(rows.cpu().numpy(), ref)
    rows = sample_rows(5000, 0.25, 77, None, row_base=123456, device="cuda")
    ref = np.nonzero(bernoulli_mask(77, np.arange(5000) + 123456, 0.25))[0]
    assert np.array_equal(rows.cpu().numpy(), ref)


@pytest.mark.parametrize("alg,step,P", [("asgd", 1.0, 4), ("asaga", 0.05, 4), ("sgd_sync", 1.0, 1), ("saga_sync", 0.3, 1)])
def test_fused_driver_single_gpu(alg, step, P):
    from asyncframework_b200 import algo
    from asyncframework_b200.config import AsyncConfig
    cfg = AsyncConfig(algorithm=alg, synthetic="linear", num_rows=4000, num_features=20, num_partitions=P,
                      num_iterations=400 if P > 1 else 120, step_size=step, batch_rate=0.2, printer_freq=100, backend="fused",
                      loss="least_squares", bucket_ratio=0.5, barrier="bucket", rounds_per_graph=8, layout="dense")
    rep = algo.run(cfg)
    objs = [o for _, o in rep.objective]
    assert rep.iterations >= cfg.num_iterations
    assert objs[0] > 0.3 and objs[-1] < 0.05
    if P > 1:
        assert len(rep.waiting_ms) == P and sum(rep.staleness_hist.values()) >= cfg.num_iterations


def test_fused_driver_delay_injection_and_taw():
    from asyncframework_b200 import algo
    from asyncframework_b200.config import AsyncConfig
    cfg = AsyncConfig(algorithm="asgd", synthetic="linear", num_rows=8000, num_features=20, num_partitions=4,
                      num_iterations=600, step_size=1.0, batch_rate=0.2, printer_freq=300, backend="fused",
                      bucket_ratio=0.5, barrier="bucket", coeff=1.0, warmup_updates_per_worker=20, taw=1, layout="dense")
    rep = algo.run(cfg)
    assert rep.objective[-1][1] < 0.05
    assert rep.extra["avg_delay_ms"] > 0                    # calibrated on the device after 20*P updates
    # worker 0 is the straggler: it completes fewer rounds than the others
    tasks = rep.extra["avg_task_ms"]
    assert tasks[0] > 1.3 * min(tasks[p] for p in (1, 2, 3))
    assert rep.dropped >= 0 and max(rep.staleness_hist) >= 1


def test_inprocess_api_on_hbm_shards(scheduler):
    """The general closure API (async_barrier/sample/map_batch/async_reduce) over CUDA-resident shards."""
    from asyncframework_b200 import algo
    from asyncframework_b200.config import AsyncConfig
    cfg = AsyncConfig(algorithm="sgd_sync", synthetic="linear", num_rows=3000, num_features=16, num_partitions=3,
                      num_iterations=40, step_size=1.0, batch_rate=0.3, printer_freq=20, backend="auto")
    rep = algo.run(cfg)
    assert rep.final_weights.is_cuda and rep.objective[-1][1] < 0.05


def test_glm_training_on_gpu_uses_fused_kernels(scheduler):
    from asyncframework_b200.data import LinearDataGenerator, ShardedDataset
    from asyncframework_b200.models import LinearRegressionWithSGD
    X, y, w = LinearDataGenerator.generate_linear_arrays(3000, 12, eps=0.05, seed=2)
    ds = ShardedDataset.from_arrays(scheduler, X, y, 2, dtype=torch.float32, device="cuda")
    m = LinearRegressionWithSGD.train(ds, iterations=150, step=1.0, mini_batch_fraction=0.5, convergence_tol=0.0)
    assert m.weights.is_cuda and np.abs(m.weights.cpu().numpy() - w).max() < 0.08


def test_fused_engine_feeds_async_context():
    """collect_all()/stat on the fused backend: the device result ring pumps into a host AsyncContext."""
    import time
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.core import AsyncContext
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.parallel.fused_backend import FusedEngine
    P = 3
    shards = [synthetic_shard("epsilon", 2000, p, P, 5, "cuda", num_features=64, labels="regression", row_offset=2000 * p) for p in range(P)]
    cfg = AsyncConfig(algorithm="asgd", num_features=64, num_rows=6000, num_partitions=P, num_iterations=30, step_size=0.5,
                      batch_rate=0.25, loss="least_squares", barrier="asp", backend="fused")
    eng = FusedEngine(cfg, shards, 6000, list(range(P)), device="cuda")
    ctx = AsyncContext(P)
    eng.start_result_pump(ctx)
    eng.run_rounds(12)
    eng.synchronize()
    got = []
    t0 = time.time()
    while len(got) < 30 and time.time() - t0 < 10:
        r = ctx.collect_all(timeout=0.5)
        if r is not None:
            got.append(r)
    eng.stop_result_pump()
    assert len(got) == 30 and sum(1 for r in got if r.data["accepted"]) == 30       # exactly num_iterations accepted
    assert sorted({r.worker_id for r in got}) == [0, 1, 2]
    assert all(r.batch_size > 0 and r.staleness >= 0 and r.finish_ms >= r.submit_ms for r in got)
    assert [r.data["k_after"] for r in got] == list(range(1, 31))
    assert ctx.stat.available_workers() == P and ctx.current_time >= 30


def test_cuda_graph_replay_matches_plain_launches():
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.parallel.fused_backend import FusedEngine

    def make():
        sh = synthetic_shard("epsilon", 4000, 0, 1, 5, "cuda", num_features=128, labels="regression")
        cfg = AsyncConfig(algorithm="sgd_sync", num_features=128, num_rows=4000, num_partitions=1, num_iterations=1 << 30,
                          step_size=0.5, batch_rate=0.25, loss="least_squares", backend="fused")
        return FusedEngine(cfg, [sh], 4000, [0], device="cuda")
    a, b = make(), make()
    a.run_rounds(12)
    b.build_graph(4)
    b.run_graph(3)
    a.synchronize(); b.synchronize()
    assert a.launches == 1 and b.launches == 3          # persistent rounds: 12 rounds = one launch; a 4-round graph replayed 3 times
    assert torch.allclose(a.weights(), b.weights(), rtol=1e-4, atol=1e-6)
    assert a.kernels[0].read_ctl()["round"] == b.kernels[0].read_ctl()["round"] == 12


def test_fused_checkpoint_and_resume(tmp_path):
    from asyncframework_b200 import algo
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.utils import load_checkpoint
    base = dict(algorithm="asaga", synthetic="linear", num_rows=4000, num_features=20, num_partitions=2, step_size=0.05,
                batch_rate=0.2, printer_freq=100, backend="fused", loss="least_squares", barrier="asp", layout="dense",
                rounds_per_graph=5)
    cfg = AsyncConfig(num_iterations=200, checkpoint_dir=str(tmp_path / "ck"), checkpoint_every=100, **base)
    rep = algo.run(cfg)
    ck = load_checkpoint(str(tmp_path / "ck"))
    assert ck["k"] >= 100 and ck["alpha_bar"] is not None and sorted(ck["history"]) == [0, 1]
    cfg2 = AsyncConfig(num_iterations=300, resume=str(tmp_path / "ck"), **base)      # num_iterations is a TOTAL: 100 more than the checkpoint
    rep2 = algo.run(cfg2)
    # resumed run starts from the checkpointed model: its first objective is far below the cold-start one
    assert rep2.objective[0][1] < 0.5 * rep.objective[0][1]
    assert rep2.objective[-1][1] < 0.05
