"""The fused backend's host logic against the pure-Python round simulator (no GPU needed):
the simulator itself is checked against the NumPy oracle, then FusedEngine / run_fused / the result
pump / checkpointing run on top of it."""
import math
import os
import time

import numpy as np
import pytest
import torch

from asyncframework_b200 import algo
from asyncframework_b200.algo.oracle import AsyncOracle, run_asaga_oracle, run_asgd_oracle, run_saga_sync_oracle, run_sgd_sync_oracle
from asyncframework_b200.config import AsyncConfig
from asyncframework_b200.core import AsyncContext
from asyncframework_b200.data.batch import Shard, padded_ld
from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
from asyncframework_b200.optim.gradient import LOSS_IDS
from asyncframework_b200.parallel.fused_backend import FusedEngine


def _parts(P, n, d, seed, labels01):
    rng = np.random.RandomState(seed)
    wstar = rng.randn(d) * 0.5
    out, off = [], 0
    for p in range(P):
        X = (rng.randn(n, d) * 0.4).astype(np.float32).astype(np.float64)
        y = ((rng.rand(n) < 1 / (1 + np.exp(-X @ wstar))).astype(np.float64) if labels01
             else (X @ wstar + 0.05 * rng.randn(n)).astype(np.float32).astype(np.float64))
        out.append((X, y, np.arange(n) + off))
        off += n
    return out


def _shard(part, d, p):
    X, y, g = part
    Xp = torch.zeros(X.shape[0], padded_ld(d), dtype=torch.float32)
    Xp[:, :d] = torch.from_numpy(X).float()
    return Shard(d, torch.from_numpy(y).float(), Xp, gidx=torch.from_numpy(g), partition=p)


@pytest.mark.parametrize("algo_name,loss,gamma", [("asgd", "logistic", 0.8), ("asaga", "least_squares", 0.05)])
def test_simulated_round_matches_oracle_single_worker(algo_name, loss, gamma):
    d, n = 30, 800
    parts = _parts(1, n, d, 3, loss == "logistic")
    sh = _shard(parts[0], d, 0)
    layout = ServerLayout(sh.ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, "cpu")
    blk.reset()
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo=algo_name, loss_id=LOSS_IDS[loss], num_workers=1,
                     worker_id=0, total_rows=n, batch_rate=0.3, gamma=gamma, seed=42, seed_mode=0, sim_server_block=blk)
    assert rk.sim
    for _ in range(6):
        rk.launch()
    events = []
    for _ in range(6):
        events += [("submit", [0]), ("complete", 0)]
    orc = AsyncOracle(parts, loss, 0.3, gamma, 10**9, 42)
    ref = (run_asgd_oracle if algo_name == "asgd" else run_asaga_oracle)(orc, events, 10**6)
    assert np.allclose(blk.w[:d].double().numpy(), ref.w, rtol=2e-4, atol=1e-6)
    st = blk.read_sctl()
    assert st["k"] == 6 and st["clock"] == 6 and st["avail_count"] == 1
    ring = blk.read_ring()
    assert [int(r["seq"]) for r in ring[:6]] == [1, 2, 3, 4, 5, 6]


def test_simulated_bsp_matches_sync_oracle():
    d, n = 20, 600
    for algo_name, loss, gamma, fn in (("sgd_sync", "logistic", 0.7, run_sgd_sync_oracle), ("saga_sync", "least_squares", 0.05, run_saga_sync_oracle)):
        parts = _parts(1, n, d, 5, loss == "logistic")
        cfg = AsyncConfig(algorithm=algo_name, num_features=d, num_rows=n, num_partitions=1, num_iterations=12, step_size=gamma,
                          batch_rate=0.4, loss=loss, seed=42, backend="fused")
        eng = FusedEngine(cfg, [_shard(parts[0], d, 0)], n, [0], device="cpu")
        eng.run_rounds(12)
        ref = fn(AsyncOracle(parts, loss, 0.4, gamma, 10**9, 42), 12)
        assert np.allclose(eng.weights().double().numpy(), ref.w, rtol=2e-4, atol=1e-6)


def test_fused_engine_protocol_on_simulator():
    """4 logical workers interleaved on one simulated server: clock/k/staleness/availability/ring invariants,
    the max_iters stop, the taw filter and the AsyncContext result pump."""
    P, d, n = 4, 24, 500
    parts = _parts(P, n, d, 9, True)
    shards = [_shard(parts[p], d, p) for p in range(P)]
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n * P, num_partitions=P, num_iterations=50, step_size=0.8,
                      batch_rate=0.3, loss="logistic", seed=1, backend="fused", barrier="asp", history_slots=2)
    eng = FusedEngine(cfg, shards, n * P, list(range(P)), device="cpu")
    ctx = AsyncContext(P)
    eng.start_result_pump(ctx, period_s=0.001)
    losses = []
    for _ in range(20):
        r = eng.step()
        if r["batch"]:
            losses.append(r["loss_sum"] / r["batch"])
    st = eng.server_state()
    assert st["k"] == 50 and st["clock"] == 50 and st["avail_count"] == P          # stops exactly at max_iters
    assert losses[-1] < losses[0]
    got = []
    t0 = time.time()
    while len(got) < 50 and time.time() - t0 < 5:
        r = ctx.collect_all(timeout=0.2)
        if r is not None:
            got.append(r)
    eng.stop_result_pump()
    assert len(got) == 50 and [r.data["k_after"] for r in got] == list(range(1, 51))
    assert sorted({r.worker_id for r in got}) == list(range(P))
    assert all(r.staleness == 0 for r in got)            # sequential launches: nobody merges between a pull and its push
    assert ctx.stat.available_workers() == P
    v0 = eng.publish_version()
    eng.block.w.add_(1.0)
    v1 = eng.publish_version()
    assert (v0, v1) == (0, 1) and torch.allclose(eng.version_value(v1) - eng.version_value(v0), torch.ones(d))
    with pytest.raises(RuntimeError):
        eng.build_graph(2)


def test_staleness_filter_and_saga_rollback_on_simulator():
    """A result whose staleness exceeds taw must tick the clock, NOT the accepted-update counter, leave w
    untouched and roll the worker's SAGA history back (SparkASAGAThread.scala:184-231)."""
    P, d, n = 2, 10, 300
    parts = _parts(P, n, d, 4, False)
    layout = ServerLayout(padded_ld(d), P, ring_size=32)
    blk = ServerBlock.allocate(layout, "cpu")
    blk.reset()

    def mk(p):
        return RoundKernel(_shard(parts[p], d, p), ServerPointers(blk.buf.data_ptr(), layout), algo="asaga",
                           loss_id=LOSS_IDS["least_squares"], num_workers=P, worker_id=p, total_rows=n * P, batch_rate=0.5,
                           gamma=0.05, seed=7, taw=0, sim_server_block=blk)
    a, b = mk(0), mk(1)
    assert a.rollback and b.rollback
    a.launch()                                        # accepted: k 0 -> 1
    w_after_a = blk.w.clone()
    hist_before = b.hist.clone()
    b.params.taw = -1                                 # a sequential model cannot interleave: force "too stale"
    b.launch()
    st = blk.read_sctl()
    assert st["k"] == 1 and st["clock"] == 2          # merged (clock ticks) but not accepted
    assert torch.equal(blk.w, w_after_a)              # w untouched
    assert torch.equal(b.hist, hist_before)           # history rolled back
    assert b.read_ctl()["last_accepted"] == 0 and int(blk.read_ring()[1]["accepted"]) == 0
    assert st["avail_count"] == P                     # the worker is re-queued (available again)


@pytest.mark.parametrize("alg,step,P", [("asgd", 1.0, 4), ("asaga", 0.05, 4), ("sgd_sync", 1.0, 1), ("saga_sync", 0.3, 1)])
def test_fused_driver_on_simulator(alg, step, P, tmp_path):
    cfg = AsyncConfig(algorithm=alg, synthetic="linear", num_rows=2000, num_features=12, num_partitions=P,
                      num_iterations=300 if P > 1 else 100, step_size=step, batch_rate=0.2, printer_freq=100, backend="fused",
                      loss="least_squares", barrier="asp", rounds_per_graph=5, layout="dense",
                      checkpoint_dir=str(tmp_path / "ck"), checkpoint_every=100)
    rep = algo.run(cfg)
    objs = [o for _, o in rep.objective]
    assert rep.iterations >= cfg.num_iterations and objs[0] > 0.3 and objs[-1] < 0.05
    from asyncframework_b200.utils import load_checkpoint
    ck = load_checkpoint(str(tmp_path / "ck"))
    assert ck["k"] >= 100 and ck["w"].numel() == 12
    import json
    latest = json.load(open(tmp_path / "ck" / "latest.json"))          # published by the background writer AFTER the file
    assert latest["k"] == ck["k"] and not list((tmp_path / "ck").glob("*.tmp"))
    if alg in ("asaga", "saga_sync"):
        assert set(ck["history"]) == set(range(P)) and all(h.numel() > 0 for h in ck["history"].values())


def test_device_watchdog_on_simulated_status_table():
    """DeviceWatchdog flags a worker that went busy and then silent (heartbeat older than the newest by > timeout)."""
    from asyncframework_b200.ops._signatures import WORKER_SLOT_DTYPE
    from asyncframework_b200.utils import DeviceWatchdog
    P, d, n = 3, 8, 200
    parts = _parts(P, n, d, 2, True)
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n * P, num_partitions=P, num_iterations=1 << 30, step_size=0.5,
                      batch_rate=0.5, loss="logistic", backend="fused", barrier="asp")
    eng = FusedEngine(cfg, [_shard(parts[p], d, p) for p in range(P)], n * P, list(range(P)), device="cpu")
    eng.run_rounds(2)
    wd = DeviceWatchdog(eng, timeout_s=1e-3)
    assert wd.check_once() == []
    st = eng.block.status.numpy().view(WORKER_SLOT_DTYPE)
    st["available"][1] = 0                                   # worker 1 started a round ...
    st["heartbeat_ns"][1] = int(st["heartbeat_ns"].max()) - 10_000_000      # ... 10 ms before everybody's last sign of life
    assert wd.check_once() == [1] and wd.dead == [1]
    assert wd.check_once() == []                             # reported once


def test_persistent_variant_is_n_back_to_back_rounds_on_the_simulator():
    """launch_rounds(n) of the persistent stream variant must be protocol-identical to n single launches."""
    P, d, n = 2, 12, 300
    parts = _parts(P, n, d, 5, False)
    cfg = AsyncConfig(algorithm="asaga", num_features=d, num_rows=n * P, num_partitions=P, num_iterations=1 << 30, step_size=0.05,
                      batch_rate=0.3, loss="least_squares", backend="fused", barrier="asp")
    outs = []
    for variant in ("simple", "stream"):
        os.environ["AFB_DENSE_VARIANT"] = variant
        try:
            eng = FusedEngine(cfg, [_shard(parts[p], d, p) for p in range(P)], n * P, list(range(P)), device="cpu")
        finally:
            os.environ.pop("AFB_DENSE_VARIANT", None)
        assert all(rk.persist == (variant == "stream") for rk in eng.kernels)
        eng.run_rounds(6)
        st = eng.server_state()
        outs.append((eng.weights().clone(), st["k"], st["clock"], eng.launches))
    # the simulator runs workers in launch order: simple interleaves w0,w1,w0,w1..., persist runs w0 x6 then w1 x6, so only
    # the protocol counters (not the weights) must agree; the launch count shows the one-launch-per-worker structure
    assert outs[0][1] == outs[1][1] == 12 and outs[0][2] == outs[1][2] == 12
    assert outs[0][3] == 12 and outs[1][3] == 2
    assert torch.isfinite(outs[1][0]).all()


def test_pipelined_variant_semantics_on_the_simulator():
    """Pipelined persistent rounds: same protocol counters as n single launches, but round i+1 is computed on parameters
    that lack the worker's own round-i update (one own update in flight) — checked against a NumPy transcription."""
    d, n, b, gamma, seed, rounds = 10, 400, 0.5, 0.4, 3, 5
    parts = _parts(1, n, d, 9, True)
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=gamma,
                      batch_rate=b, loss="logistic", backend="fused", barrier="asp", seed=seed)
    eng = FusedEngine(cfg, [_shard(parts[0], d, 0)], n, [0], device="cpu")      # the default: stream kernel, pipelined ASGD
    rk = eng.kernels[0]
    assert rk.pipelined and not rk.persist and rk.variant_effective == "stream-pipelined"
    eng.run_rounds(rounds)
    st = eng.server_state()
    assert st["k"] == rounds and st["clock"] == rounds and eng.launches == 1
    # reference: w_pulled(0) = w0; w_pulled(i+1) = server BEFORE update i  (=> update i is computed on w_{i-1})
    from asyncframework_b200.core.sampler import bernoulli_mask
    X, y, gid = parts[0]
    X, y = torch.from_numpy(X), torch.from_numpy(y)
    w_srv = torch.zeros(d, dtype=torch.float64)
    w_pull = w_srv.clone()
    for i in range(rounds):
        rows = torch.from_numpy(np.nonzero(bernoulli_mask(seed + i + 1, gid, b))[0])
        m = X[rows] @ w_pull
        g = X[rows].T @ (torch.sigmoid(m) - y[rows])
        before = w_srv.clone()
        w_srv = w_srv - gamma / math.sqrt(i + 1) * g / (b * n)
        w_pull = before
    assert torch.allclose(eng.weights()[:d].double(), w_srv, rtol=1e-4, atol=1e-6)
    # an ineligible request (ASAGA) falls back to the barrier-persistent variant and says so
    cfg2 = AsyncConfig(algorithm="asaga", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=0.05,
                       batch_rate=b, loss="least_squares", backend="fused", barrier="asp")
    eng2 = FusedEngine(cfg2, [_shard(parts[0], d, 0)], n, [0], device="cpu")
    assert eng2.kernels[0].variant_effective == "stream"
    # ... and so does ASGD behind a device gate (bucket) or with pipelining switched off
    cfg3 = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=gamma,
                       batch_rate=b, loss="logistic", backend="fused", barrier="asp", seed=seed)
    eng3 = FusedEngine(cfg3, [_shard(parts[0], d, 0)], n, [0], device="cpu", pipeline=False)
    assert eng3.kernels[0].variant_effective == "stream" and not eng3.kernels[0].pipelined


def test_persistent_csr_variant_on_the_simulator():
    """Wide CSR shards run the persistent one-kernel round (csr_stream.cu) by default; on the simulator that is n rounds."""
    from asyncframework_b200.data.generators import synthetic_shard
    cfg = AsyncConfig(algorithm="asgd", num_features=300, num_rows=600, num_partitions=2, num_iterations=1 << 30, step_size=0.5,
                      batch_rate=0.3, loss="logistic", backend="fused", barrier="asp")
    shards = [synthetic_shard("rcv1", 300, p, 2, seed=3, device="cpu", num_features=300, row_offset=300 * p) for p in range(2)]
    os.environ["AFB_SPLIT_MIN_LD"] = "64"
    try:
        eng = FusedEngine(cfg, shards, 600, [0, 1], device="cpu")
    finally:
        os.environ.pop("AFB_SPLIT_MIN_LD", None)
    assert all(rk.persist_csr and rk.variant_effective == "csr-stream" for rk in eng.kernels)
    eng.run_rounds(4)
    st = eng.server_state()
    assert st["k"] == 8 and eng.launches == 2 and torch.isfinite(eng.weights()).all()


def test_bf16_row_storage_on_the_simulator_is_exact_for_integer_pixels():
    """bf16 ROW STORAGE, fp32 math: for the mnist shapes (integer pixels) the rounds are bit-identical to fp32 storage."""
    from asyncframework_b200.data.generators import synthetic_shard
    outs = []
    for row_dtype in (None, torch.bfloat16):
        sh = synthetic_shard("mnist", 500, 0, 1, seed=3, device="cpu", num_features=784, row_dtype=row_dtype)
        assert sh.y.dtype == torch.float32 and sh.X.dtype == (row_dtype or torch.float32) and sh.X.shape[1] == 784
        cfg = AsyncConfig(algorithm="asgd", num_features=784, num_rows=500, num_partitions=1, num_iterations=1 << 30, step_size=1e-4,
                          batch_rate=0.3, loss="logistic", backend="fused", barrier="asp")
        eng = FusedEngine(cfg, [sh], 500, [0], device="cpu")
        assert eng.kernels[0].bf16_rows == (row_dtype is not None)
        assert eng.kernels[0].variant_effective == "stream-pipelined"        # one kernel family for fp32 and bf16 rows
        eng.run_rounds(4)
        outs.append(eng.weights().clone())
    assert torch.equal(outs[0], outs[1]) and outs[0].abs().sum() > 0


@pytest.mark.parametrize("alg,step", [("sgd_sync", 1.0), ("saga_sync", 0.3), ("asgd", 1.0)])
def test_fused_resume_is_exact_and_num_iterations_is_a_total(alg, step, tmp_path):
    """A run interrupted at k = 40 and resumed to a total of 80 ends on exactly the weights of an uninterrupted 80-iteration run:
    the accepted-update counter, the logical clock and the per-worker round counters are restored, so the step-size schedule
    (gamma / sqrt(round + 1), gamma / sqrt(floor(k / P) + 1)) and the seed + round + 1 sampling sequence continue."""
    from asyncframework_b200.utils import load_checkpoint
    base = dict(algorithm=alg, synthetic="linear", num_rows=1500, num_features=10, num_partitions=1, step_size=step, batch_rate=0.3,
                printer_freq=1000, backend="fused", loss="least_squares", barrier="asp", rounds_per_graph=4, layout="dense",
                eval_objective=False, quiet=True)
    full = algo.run(AsyncConfig(num_iterations=80, **base))
    part = algo.run(AsyncConfig(num_iterations=40, checkpoint_dir=str(tmp_path / "ck"), checkpoint_every=40, **base))
    ck = load_checkpoint(str(tmp_path / "ck"))
    assert ck["k"] == 40 and ck["extra"]["rounds"] == [40]
    assert not torch.equal(part.final_weights, full.final_weights)
    resumed = algo.run(AsyncConfig(num_iterations=80, resume=str(tmp_path / "ck"), **base))
    assert resumed.iterations == 80
    assert torch.allclose(resumed.final_weights, full.final_weights, rtol=0, atol=0), (resumed.final_weights - full.final_weights).abs().max()
    # num_iterations is a TOTAL: resuming with a target already reached runs nothing more
    again = algo.run(AsyncConfig(num_iterations=40, resume=str(tmp_path / "ck"), **base))
    assert torch.equal(again.final_weights, part.final_weights)


def test_predicate_mask_and_declare_dead_on_the_simulator():
    """ASYNCbarrier on the fused path (host-evaluated predicate -> device enable mask) and the continue-with-P-1 bookkeeping:
    masked workers skip their rounds, a worker declared dead leaves the idle count / the mask and lowers the bucket threshold."""
    P, d, n = 4, 8, 300
    parts = _parts(P, n, d, 2, True)
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n * P, num_partitions=P, num_iterations=1 << 30, step_size=0.5,
                      batch_rate=0.5, loss="logistic", backend="fused", barrier="bucket", bucket_ratio=0.75)
    eng = FusedEngine(cfg, [_shard(parts[p], d, p) for p in range(P)], n * P, list(range(P)), device="cpu", worker_mask=True)
    assert eng.gate_threshold == 3 and all(rk.params.use_mask == 1 and not rk.pipelined for rk in eng.kernels)
    eng.run_rounds(2)
    assert eng.server_state()["k"] == 2 * P
    # predicate over the mirrored WorkerState table: only workers 0 and 2 take part in the next rounds
    sel = eng.async_barrier(lambda st: st.worker_id % 2 == 0 if hasattr(st, "worker_id") else True)
    eng.set_worker_mask([True, False, True, False])
    eng.run_rounds(3)
    st = eng.server_state()
    assert st["k"] == 2 * P + 3 * 2
    skipped = [int(rk.read_ctl()["skipped_rounds"]) for rk in eng.kernels]
    assert skipped == [0, 3, 0, 3]
    # worker 3 is declared dead while idle: the idle count drops, the threshold follows floor(P_alive * beta), its bit stays clear
    eng.set_worker_mask([True] * P)
    eng.declare_dead(3)
    st = eng.server_state()
    assert st["avail_count"] == P - 1 and st["num_alive"] == 3 and st["gate_dyn"] == 2
    assert int(eng.worker_status()[3]["alive"]) == 0 and (st["enable_mask"][0] & 0b1000) == 0
    eng.run_rounds(2)
    assert eng.server_state()["k"] == 2 * P + 3 * 2 + 2 * 3          # the three live workers keep going


def test_regularised_fused_rounds_on_the_simulator_match_the_updater_formulas():
    """reg_param / updater reach the fused backend (K8 fused into the apply): BSP rounds with the squared-L2 / L1 updaters
    shrink the weights exactly like SquaredL2Updater / L1Updater (Updater.scala:114-158)."""
    d, n = 10, 400
    parts = _parts(1, n, d, 4, False)
    outs = {}
    for upd, lam in (("simple", 0.0), ("squared_l2", 0.5), ("l1", 0.05)):
        cfg = AsyncConfig(algorithm="sgd_sync", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=0.2,
                          batch_rate=0.5, loss="least_squares", backend="fused", updater=upd, reg_param=lam)
        eng = FusedEngine(cfg, [_shard(parts[0], d, 0)], n, [0], device="cpu")
        assert eng.kernels[0].params.reg_type == {"simple": 0, "squared_l2": 1, "l1": 2}[upd]
        eng.run_rounds(25)
        outs[upd] = eng.weights().clone()
    assert outs["squared_l2"].norm() < 0.9 * outs["simple"].norm()
    assert int((outs["l1"] == 0).sum()) >= int((outs["simple"] == 0).sum()) and outs["l1"].abs().sum() < outs["simple"].abs().sum()


class _FakeE2eLib:
    """Stands in for the CUDA side-stream plumbing (csrc/kernels/e2e_step.cu) so the host logic of FusedEngine.step() — slot
    rotation, one descriptor copy per step, result copy into the pinned slot the parser reads, first-step fallback — runs on CPU."""

    def __init__(self, eng, fail_first_pre=False):
        self.eng, self.calls, self.fail_first_pre, self.destroyed = eng, [], fail_first_pre, False

    def afb_e2e_pre(self, ctx, slot, desc_dev, desc_host, nbytes, stream):
        self.calls.append(("pre", slot, desc_dev - self.eng._e2e_ptrs[0], nbytes))
        if self.fail_first_pre and len(self.calls) == 1:
            return 2
        ds = (desc_dev - self.eng._e2e_ptrs[0]) // nbytes
        self.eng._desc_dev[ds].copy_(self.eng._desc_host[ds])         # what the H2D copy does on the device
        return 0

    def afb_e2e_post(self, ctx, slot, wctl, snap, host, nbytes, stream):
        self.calls.append(("post", slot, nbytes))
        self.eng._ctl_host[slot].copy_(self.eng.wctl_all)            # what snapshot kernel + D2H do on the device
        return 0

    def afb_e2e_wait(self, ctx, slot):
        self.calls.append(("wait", slot))
        return 0

    def afb_e2e_sync(self, ctx):
        return 0

    def afb_e2e_destroy(self, ctx):
        self.destroyed = True


def _engine_with_fake_side_streams(fail_first_pre=False):
    d, n = 16, 400
    parts = _parts(1, n, d, 3, True)
    cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=0.5,
                      batch_rate=0.3, loss="logistic", seed=2, backend="fused", barrier="asp")
    eng = FusedEngine(cfg, [_shard(parts[0], d, 0)], n, [0], device="cpu")
    assert eng._e2e is None                                           # CPU engines never create the CUDA context
    eng._e2e_ptrs = (1 << 20, 2 << 20, 3 << 20, [4 << 20, 5 << 20], [6 << 20, 7 << 20])
    eng._snap = torch.zeros(2, 1, eng.wctl_all.shape[1], dtype=torch.uint8)
    eng._e2e_lib = _FakeE2eLib(eng, fail_first_pre)
    eng._e2e = 1
    return eng


@pytest.mark.parametrize("lag", [0, 1])
def test_step_with_side_stream_plumbing_rotates_slots_and_returns_every_record(lag):
    eng = _engine_with_fake_side_streams()
    lib = eng._e2e_lib
    ref = FusedEngine(eng.cfg, [eng.kernels[0].shard], eng.total_rows, [0], device="cpu")
    got, want = [], []
    for i in range(7):
        r = eng.step(step_size=0.5 / (i + 1), lag=lag)
        if r is not None:
            got.append((r["k"], r["batch"], r["accepted"]))
        w = ref.step(step_size=0.5 / (i + 1))
        want.append((w["k"], w["batch"], w["accepted"]))
    last = eng.flush_steps()
    if last is not None:
        got.append((last["k"], last["batch"], last["accepted"]))
    assert got == want and torch.allclose(eng.weights(), ref.weights())
    pres = [c for c in lib.calls if c[0] == "pre"]
    posts = [c for c in lib.calls if c[0] == "post"]
    assert len(pres) == 7 and len(posts) == 7                         # exactly one H2D and one D2H per step
    assert [c[1] for c in pres] == [i & 1 for i in range(7)] and [c[1] for c in posts] == [i & 1 for i in range(7)]
    assert [c[2] for c in pres] == [16 * i for i in range(7)] and all(c[3] == 16 for c in pres)     # descriptor ring slot = round % slots
    assert all(c[2] == eng.wctl_all.shape[1] for c in posts)
    eng._e2e = None                                                   # nothing to destroy for the fake


def test_first_step_falls_back_to_the_launching_stream_when_the_side_stream_copy_fails():
    eng = _engine_with_fake_side_streams(fail_first_pre=True)
    lib = eng._e2e_lib
    with pytest.warns(UserWarning, match="side-stream descriptor copy failed"):
        r = eng.step(step_size=0.1)
    assert r["k"] == 1 and r["accepted"] == 1 and eng._e2e is None and lib.destroyed
    r = eng.step(step_size=0.1)                                        # later steps use the same-stream path
    assert r["k"] == 2 and [c[0] for c in lib.calls] == ["pre"]
