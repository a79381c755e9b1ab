"""GPU tests of the default round kernels: the persistent TMA-ring stream kernel (csrc/kernels/dense_stream.cu: serial and
pipelined rounds, fp32 and bf16 row storage), the one-kernel CSR round (csr_stream.cu), the tensor-core evaluation with a loss
epilogue, and the protocol features fused into them (regularisation, enable mask, SSP gate, round descriptors, bounded spins).
Numerics are compared against the one-round register kernel, the pure-Python round simulator and fp64 oracles."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


def _dense_shard(n, d, seed=0, labels01=True, device="cuda"):
    from asyncframework_b200.data.batch import Shard, padded_ld
    g = torch.Generator().manual_seed(seed)
    ld = padded_ld(d)
    X = torch.zeros(n, ld)
    X[:, :d] = torch.randn(n, d, generator=g) * 0.3
    y = (torch.rand(n, generator=g) < 0.5).float() if labels01 else torch.randn(n, generator=g)
    gidx = torch.arange(n, dtype=torch.int64) + 1000
    return Shard(d, y.to(device), X.to(device), gidx=gidx.to(device))


def _kernel(sh, *, algo, loss, n, b, gamma, seed, variant=None, dev="cuda", **kw):
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    layout = ServerLayout(sh.ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, dev)
    blk.reset()
    if algo in ("sgd_sync", "saga_sync"):
        kw.setdefault("local_block", blk)
        kw.setdefault("peer_blocks", [ServerPointers(blk.buf.data_ptr(), layout)])
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo=algo, loss_id=LOSS_IDS[loss], num_workers=1,
                     worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed, variant=variant,
                     sim_server_block=blk if dev == "cpu" else None, **kw)
    return rk, blk


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 784), ("asaga", "least_squares", 200), ("asgd", "logistic", 1000),
                                         ("sgd_sync", "logistic", 784), ("saga_sync", "least_squares", 2000), ("asgd", "hinge", 36)])
def test_persistent_stream_rounds_match_single_round_launches(algo, loss, d):
    """n rounds in ONE persistent launch of the TMA-ring kernel == n launches of the one-round register kernel (single
    worker: deterministic up to fp32 summation order)."""
    n, b, gamma, seed, rounds = 6000, 0.25, 0.05, 5, 7
    sh = _dense_shard(n, d, seed=3, labels01=(loss != "least_squares"))
    res = {}
    for variant in ("simple", "stream"):
        kw = dict(max_iters=rounds - 1) if algo in ("asgd", "asaga") else {}
        rk, blk = _kernel(sh, algo=algo, loss=loss, n=n, b=b, gamma=gamma, seed=seed, variant=variant, pipeline=False, **kw)
        assert rk.stream == (variant == "stream") and not rk.pipelined
        rk.launch_rounds(rounds)              # async: the last round hits k >= max_iters: the persistent grid must exit cleanly
        torch.cuda.synchronize()
        rk.check_abort()
        st = blk.read_sctl()
        ctl = rk.read_ctl()
        res[variant] = (blk.w.clone(), blk.abar.clone(), st["k"], st["clock"], ctl["rounds_done"], rk.launches, ctl["last_batch"],
                        None if rk.hist is None else rk.hist.clone())
    a, c = res["simple"], res["stream"]
    if algo in ("asgd", "asaga"):
        assert a[2] == c[2] == rounds - 1 and a[3] == c[3]
    assert a[4] == c[4]
    assert c[5] == 1 and a[5] == rounds
    assert torch.allclose(a[0], c[0], rtol=1e-4, atol=1e-6) and torch.allclose(a[1], c[1], rtol=1e-4, atol=1e-6)
    if a[7] is not None:
        assert torch.allclose(a[7], c[7], rtol=1e-4, atol=1e-6)


def test_persistent_launch_refuses_non_resident_grid():
    sh = _dense_shard(2000, 784, seed=3)
    rk, _ = _kernel(sh, algo="asgd", loss="logistic", n=2000, b=0.5, gamma=0.1, seed=1, pipeline=False)
    with pytest.raises(RuntimeError):
        rk.launch_rounds(4, share=2)          # full-capacity grid x 2 co-running grids cannot be resident
    torch.cuda.synchronize()


@pytest.mark.parametrize("d", [784, 200, 1000])
def test_pipelined_rounds_match_the_simulator_semantics(d):
    """Single worker => deterministic: n pipelined rounds on the GPU == the simulator's transcription of the same schedule
    (round i+1 computed on the parameters pulled before update i landed).  The pull-ahead is switched off so that the
    parameters a round pulls do not depend on timing."""
    n, b, gamma, seed, rounds = 6000, 0.25, 0.05, 5, 8
    sh = _dense_shard(n, d, seed=3, labels01=True)
    res = {}
    os.environ["AFB_PULL_AHEAD_CHUNKS"] = "0"
    try:
        for dev in ("cuda", "cpu"):
            s = sh if dev == "cuda" else sh.to("cpu")
            rk, blk = _kernel(s, algo="asgd", loss="logistic", n=n, b=b, gamma=gamma, seed=seed, dev=dev, max_iters=rounds - 2)
            assert rk.pipelined
            rk.launch_rounds(rounds)              # k reaches max_iters inside the launch: the grid must drain and exit
            if dev == "cuda":
                torch.cuda.synchronize()
                rk.check_abort()
            st = blk.read_sctl()
            res[dev] = (blk.w.cpu().clone(), st["k"], st["clock"], st["avail_count"], rk.read_ctl()["rounds_done"])
    finally:
        os.environ.pop("AFB_PULL_AHEAD_CHUNKS", None)
    g, c = res["cuda"], res["cpu"]
    assert g[1] == rounds - 2 and g[3] == 1          # accepted updates, worker idle again
    assert g[1] <= c[1] + 1 and abs(g[2] - c[2]) <= 1  # the device may start one more round before it sees k >= max_iters
    assert torch.allclose(g[0], c[0], rtol=2e-3, atol=1e-5)


def test_pipelined_rounds_with_pull_ahead_keep_the_protocol_invariants():
    """With the pull-ahead armed the parameters a round sees depend on timing (that is the point), so only the invariants are
    checked: every round accepted exactly once, clock == k, worker idle at the end, finite model, loss going down."""
    n, d, b, rounds = 400_000, 784, 0.1, 24
    from asyncframework_b200.data.generators import synthetic_shard
    sh = synthetic_shard("mnist8m", n, 0, 1, seed=2, device="cuda")
    rk, blk = _kernel(sh, algo="asgd", loss="logistic", n=n, b=b, gamma=2e-6, seed=7)
    assert rk.pipelined and rk.params.pull_ahead_chunks > 0
    rk.launch_rounds(rounds)
    torch.cuda.synchronize()
    rk.check_abort()
    st = blk.read_sctl()
    ring = blk.read_ring()
    assert st["k"] == rounds and st["clock"] == rounds and st["avail_count"] == 1
    seqs = sorted(int(r["seq"]) for r in ring if int(r["seq"]))
    assert seqs == list(range(1, rounds + 1))
    losses = [float(r["loss_sum"]) / max(int(r["batch"]), 1) for r in sorted(ring[:rounds], key=lambda r: int(r["seq"]))]
    assert all(math.isfinite(x) for x in losses) and losses[-1] < losses[0]
    assert bool(torch.isfinite(blk.w).all())


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 784), ("asaga", "least_squares", 200), ("asgd", "logistic", 2000)])
def test_bf16_row_storage_matches_fp32_rows_on_bf16_exact_data(algo, loss, d):
    """Rows that are exactly representable in bf16 (integer pixels) must give the same rounds from bf16 row storage as from
    fp32 rows (same kernel family, both accumulate in fp32; only the summation order inside a row differs)."""
    from asyncframework_b200.data.batch import Shard
    n, b, gamma, seed, rounds = 6000, 0.25, 1e-6, 5, 5
    g = torch.Generator().manual_seed(1)
    ld = (d + 7) // 8 * 8
    X = torch.zeros(n, ld)
    X[:, :d] = (torch.randint(0, 256, (n, d), generator=g) * (torch.rand(n, d, generator=g) < 0.2)).float()
    y = (torch.rand(n, generator=g) < 0.5).float() if loss == "logistic" else torch.randn(n, generator=g)
    gidx = torch.arange(n, dtype=torch.int64)
    res = {}
    for dt in (torch.float32, torch.bfloat16):
        sh = Shard(d, y.cuda(), X.to(dt).cuda(), gidx=gidx.cuda())
        rk, blk = _kernel(sh, algo=algo, loss=loss, n=n, b=b, gamma=gamma, seed=seed, pipeline=False)
        assert rk.bf16_rows == (dt == torch.bfloat16) and rk.stream
        rk.launch_rounds(rounds)
        torch.cuda.synchronize()
        rk.check_abort()
        st = blk.read_sctl()
        res[dt] = (blk.w.clone(), st["k"], rk.read_ctl()["last_batch"])
    a, c = res[torch.float32], res[torch.bfloat16]
    assert a[1] == c[1] == rounds and a[2] == c[2]
    assert torch.allclose(a[0], c[0], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 47236), ("asaga", "least_squares", 6000)])
def test_persistent_csr_round_matches_the_split_round(algo, loss, d):
    """Wide sparse model: n rounds in ONE persistent launch with in-kernel grid barriers (multi-CTA pull / push) == n split
    rounds (gate, pull, grad, push kernels).  Single worker; float atomics reorder sums, hence the tolerance."""
    from asyncframework_b200.data.generators import synthetic_shard
    n, b, gamma, seed, rounds = 20000, 0.2, 0.05, 5, 6
    sh = synthetic_shard("rcv1", n, 0, 1, seed=3, device="cuda", num_features=d,
                         labels="binary01" if loss == "logistic" else "regression")
    res = {}
    for variant in ("simple", "stream"):
        rk, blk = _kernel(sh, algo=algo, loss=loss, n=n, b=b, gamma=gamma, seed=seed, variant=variant, max_iters=rounds - 1)
        assert rk.stream_csr == (variant == "stream") and rk.split == (variant == "simple")
        rk.launch_rounds(rounds)
        torch.cuda.synchronize()
        rk.check_abort()
        st = blk.read_sctl()
        res[variant] = (blk.w.clone(), blk.abar.clone(), st["k"], st["clock"], st["avail_count"], rk.launches,
                        None if rk.hist is None else rk.hist.clone())
    a, c = res["simple"], res["stream"]
    assert a[2] == c[2] == rounds - 1 and a[3] == c[3] and c[4] == 1 and c[5] == 1
    assert torch.allclose(a[0], c[0], rtol=1e-3, atol=1e-5) and torch.allclose(a[1], c[1], rtol=1e-3, atol=1e-6)
    if a[6] is not None:
        assert torch.allclose(a[6], c[6], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("loss", ["logistic", "hinge", "least_squares"])
@pytest.mark.parametrize("T", [8, 20, 40])
def test_tensor_core_evaluation_with_loss_epilogue_matches_fp64(loss, T):
    """csrc/kernels/eval_tc.cu: tcgen05 GEMM + loss epilogue vs an fp64 oracle and vs the exact CUDA-core kernel, on bf16-exact
    data (where the tf32 product is exact, so only the fp32 loss arithmetic differs)."""
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS, make_gradient
    sh = synthetic_shard("mnist8m", 20_000, 0, 1, seed=5, device="cuda", labels="binary01")     # integer pixels: exact in tf32
    W = (torch.randn(sh.num_features, T, device="cuda") * 1e-3)
    exact = eval_loss(sh, W, LOSS_IDS[loss], use_tensor_cores=False)
    tc = eval_loss(sh, W, LOSS_IDS[loss])                      # the default for dense shards with >= 8 snapshots
    torch.cuda.synchronize()
    s64 = sh.to("cpu", torch.float64)
    M = s64.X[:, : sh.num_features] @ W.double().cpu()
    _, L = make_gradient(loss).coefficients(M, s64.y.double()[:, None].expand_as(M))
    oracle = L.sum(0)
    assert torch.allclose(tc.cpu(), oracle, rtol=2e-4, atol=1e-2), (tc[:4], oracle[:4])
    assert torch.allclose(tc.cpu(), exact.cpu(), rtol=2e-4, atol=1e-2)


@pytest.mark.parametrize("algo,reg_type", [("sgd_sync", "l2"), ("sgd_sync", "l1"), ("saga_sync", "l2"), ("asgd", "l2"), ("asgd", "l1"),
                                           ("asaga", "l2")])
def test_regularisation_is_fused_into_the_apply(algo, reg_type):
    """K8 on the fused path: SquaredL2Updater (w <- (1 - eta*lambda) w - eta g) and L1Updater (soft threshold by eta*lambda)
    folded into the BSP replicated apply / the Hogwild push (mllib/optimization/Updater.scala:114-158), against a NumPy
    fp64 transcription driven by the same Bernoulli samples."""
    from asyncframework_b200.core.sampler import bernoulli_mask
    n, d, b, gamma, lam, seed, rounds = 4000, 60, 0.3, 0.05, 0.5 if reg_type == "l2" else 0.02, 11, 6
    sh = _dense_shard(n, d, seed=5, labels01=False)
    rk, blk = _kernel(sh, algo=algo, loss="least_squares", n=n, b=b, gamma=gamma, seed=seed, reg_param=lam, reg_type=reg_type,
                      pipeline=False)
    assert rk.params.reg_type == (1 if reg_type == "l2" else 2)
    rk.launch_rounds(rounds)
    torch.cuda.synchronize()
    rk.check_abort()
    X = sh.X[:, :d].double().cpu()
    y = sh.y.double().cpu()
    gid = sh.gidx.cpu().numpy()
    w = torch.zeros(d, dtype=torch.float64)
    abar = torch.zeros(d, dtype=torch.float64)
    hist = torch.zeros(n, dtype=torch.float64)
    saga = algo in ("asaga", "saga_sync")
    soft = lambda t, s: torch.sign(t) * torch.clamp(t.abs() - s, min=0.0)   # noqa: E731
    for i in range(rounds):
        rows = torch.from_numpy(np.nonzero(bernoulli_mask(seed + i + 1, gid, b))[0])
        r = X[rows] @ w - y[rows]
        eff = r - hist[rows] if saga else r
        if saga:
            hist[rows] = r
        g = X[rows].T @ eff
        step = gamma / math.sqrt(i + 1) if algo in ("sgd_sync", "asgd") else gamma     # P = 1: floor(k/P) = i
        if algo in ("sgd_sync", "saga_sync"):
            if reg_type == "l2":
                w = w * (1 - step * lam)
            w = w - step * g / (b * n) - (step * abar if saga else 0)
            if reg_type == "l1":
                w = soft(w, step * lam)
        else:
            delta = -step * g / (b * n) - (step * abar if saga else 0)
            if reg_type == "l2":
                delta = delta - step * lam * w
                w = w + delta
            else:
                w = soft(w + delta, step * lam)
        if saga:
            abar = abar + g / n
    assert torch.allclose(blk.w[:d].double().cpu(), w, rtol=2e-4, atol=2e-6), (blk.w[:6], w[:6])
    if reg_type == "l1" and algo == "sgd_sync":
        assert int((blk.w[:d] == 0).sum()) == int((w == 0).sum())       # the prox produces exact zeros


def test_enable_mask_skips_rounds_and_ssp_gate_opens():
    """ASYNCbarrier on the device: a worker whose enable bit is clear sits its rounds out (no update, no clock tick, the
    persistent kernel moves on); once enabled it runs.  The SSP gate (max staleness <= bound) opens when the table allows it."""
    n, d = 3000, 100
    sh = _dense_shard(n, d, seed=8)
    rk, blk = _kernel(sh, algo="asgd", loss="logistic", n=n, b=0.3, gamma=0.05, seed=3, use_mask=True, ssp_bound=0)
    assert not rk.pipelined and rk.params.use_mask == 1 and rk.params.ssp_bound == 0
    blk.set_enable_mask([False])
    rk.launch_rounds(5)
    torch.cuda.synchronize()
    rk.check_abort()
    st, ctl = blk.read_sctl(), rk.read_ctl()
    assert st["k"] == 0 and st["clock"] == 0 and ctl["skipped_rounds"] == 5 and ctl["round"] == 5 and ctl["last_accepted"] == -2
    assert float(blk.w.abs().sum()) == 0.0 and st["avail_count"] == 1
    blk.set_enable_mask([True])
    rk.launch_rounds(4)
    torch.cuda.synchronize()
    rk.check_abort()
    st, ctl = blk.read_sctl(), rk.read_ctl()
    assert st["k"] == 4 and ctl["round"] == 9 and ctl["last_accepted"] == 1 and float(blk.w.abs().sum()) > 0


def test_bounded_spin_aborts_instead_of_hanging():
    """A gate that can never open (bucket threshold 2 with a single worker) must end the kernel through the SpinGuard: the
    abort flag is raised and the host sees it, instead of a hung GPU."""
    n, d = 2000, 64
    sh = _dense_shard(n, d, seed=8)
    rk, blk = _kernel(sh, algo="asgd", loss="logistic", n=n, b=0.3, gamma=0.05, seed=3, gate_threshold=2, spin_limit_s=0.2)
    rk.launch_rounds(3)
    torch.cuda.synchronize()
    assert int(rk.read_ctl_raw()["abort"]) == 1
    with pytest.raises(RuntimeError):
        rk.check_abort()
    assert blk.read_sctl()["k"] == 0


def test_step_descriptor_is_consumed_by_the_kernel():
    """FusedEngine.step(): the pinned H2D round descriptor carries the step size / batch rate of THAT round and the kernel
    reads it — doubling the step doubles the update, halving the batch rate halves the sampled rows."""
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.parallel.fused_backend import FusedEngine
    n, d = 50_000, 96
    sh = _dense_shard(n, d, seed=4, labels01=True)
    upd, batches = [], []
    for gamma_mult, rate in ((1.0, 0.2), (2.0, 0.2), (1.0, 0.1)):
        cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=0.01,
                          batch_rate=0.2, loss="logistic", backend="fused", barrier="asp", seed=9)
        eng = FusedEngine(cfg, [sh], n, [0], device="cuda")
        r = eng.step(step_size=0.01 * gamma_mult, batch_rate=rate)
        upd.append(eng.weights().clone())
        batches.append(r["batch"])
        eng.check_health()
        assert eng.h2d_bytes_per_step == 16
    assert torch.allclose(upd[1], 2.0 * upd[0], rtol=1e-5, atol=1e-9)
    assert abs(batches[2] / batches[0] - 0.5) < 0.05 and batches[0] == batches[1]


@pytest.mark.parametrize("lag", [0, 1])
def test_overlapped_step_copies_return_the_same_records_as_same_stream_copies(lag, monkeypatch):
    """csrc/kernels/e2e_step.cu: with the per-step H2D / D2H copies on side streams (default) every step returns exactly the
    record the same-stream plumbing returns (AFB_E2E_OVERLAP=0), lagged or not, and the models end up identical."""
    from asyncframework_b200.config import AsyncConfig
    from asyncframework_b200.parallel.fused_backend import FusedEngine
    n, d = 60_000, 128
    sh = _dense_shard(n, d, seed=6, labels01=True)
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("AFB_E2E_OVERLAP", mode)
        cfg = AsyncConfig(algorithm="asgd", num_features=d, num_rows=n, num_partitions=1, num_iterations=1 << 30, step_size=0.02,
                          batch_rate=0.2, loss="logistic", backend="fused", barrier="asp", seed=11)
        eng = FusedEngine(cfg, [sh], n, [0], device="cuda")
        assert (eng._e2e is not None) == (mode == "1")
        recs = []
        for i in range(40):
            r = eng.step(step_size=0.02 / (1 + i), batch_rate=0.1 + 0.01 * (i % 7), lag=lag)
            if r is not None:
                recs.append((r["k"], r["batch"], r["accepted"], round(r["loss_sum"], 3)))
        last = eng.flush_steps()
        if last is not None:
            recs.append((last["k"], last["batch"], last["accepted"], round(last["loss_sum"], 3)))
        eng.synchronize()
        eng.check_health()
        out[mode] = (recs, eng.weights().clone())
    assert len(out["1"][0]) == 40 and [r[0] for r in out["1"][0]] == list(range(1, 41))      # one record per step, in order
    assert [r[:3] for r in out["1"][0]] == [r[:3] for r in out["0"][0]]
    for a, b in zip(out["1"][0], out["0"][0]):                       # the CTA merge order is not deterministic: low bits differ
        assert abs(a[3] - b[3]) <= 1e-4 * abs(b[3]) + 1e-2
    assert torch.allclose(out["1"][1], out["0"][1], rtol=1e-4, atol=1e-7)


def test_objective_evaluation_accepts_bf16_row_storage():
    """Shards stored with bf16 rows are evaluated through the same kernels (widened chunk by chunk) with identical results."""
    from asyncframework_b200.data.batch import Shard
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = synthetic_shard("mnist8m", 300_000, 0, 1, seed=5, device="cuda", labels="binary01", row_dtype=torch.bfloat16)
    ref = Shard(sh.num_features, sh.y, sh.X.to(torch.float32))
    W = torch.randn(sh.num_features, 12, device="cuda") * 1e-3
    a = eval_loss(sh, W, LOSS_IDS["logistic"])
    b = eval_loss(ref, W, LOSS_IDS["logistic"])
    torch.cuda.synchronize()
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
