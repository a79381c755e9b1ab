"""Numerics of every hand-written CUDA kernel against a plain PyTorch fp32/fp64 reference."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dense_shard(n, d, seed=0, labels01=True, device="cuda"):
    from asyncframework_b200.data.batch import Shard, padded_ld
    g = torch.Generator().manual_seed(seed)
    ld = padded_ld(d)
    X = torch.zeros(n, ld)
    X[:, :d] = torch.randn(n, d, generator=g) * 0.3
    y = (torch.rand(n, generator=g) < 0.5).float() if labels01 else torch.randn(n, generator=g)
    gidx = torch.arange(n, dtype=torch.int64) + 1000
    return Shard(d, y.to(device), X.to(device), gidx=gidx.to(device))


def _ref_grad(shard, w, loss_name, rows=None, hist=None):
    from asyncframework_b200.optim.gradient import make_gradient
    s = shard.to("cpu", torch.float64) if rows is None else shard.to("cpu", torch.float64).select(rows)
    g = make_gradient(loss_name)
    margin = s.matvec(w.double().cpu())
    coeff, loss = g.coefficients(margin, s.y.double())
    eff = coeff if hist is None else coeff - hist.double().cpu()[rows if rows is not None else slice(None)]
    return s.rmatvec(eff), float(loss.sum()), s.num_rows, coeff


def test_abi_sizes():
    from asyncframework_b200.ops import kernels
    from asyncframework_b200.ops import _signatures as S
    import ctypes as C
    lib = kernels()
    assert lib.afb_abi_sizeof(0) == C.sizeof(S.RoundParams)
    assert lib.afb_abi_sizeof(1) == S.WORKER_CTL_DTYPE.itemsize
    assert lib.afb_abi_sizeof(2) == S.SERVER_CTL_DTYPE.itemsize
    assert lib.afb_abi_sizeof(3) == S.WORKER_SLOT_DTYPE.itemsize
    assert lib.afb_abi_sizeof(4) == S.RESULT_REC_DTYPE.itemsize
    assert lib.afb_abi_sizeof(5) == C.sizeof(S.SyncPeers)


@pytest.fixture(params=["simple", "pipe", "ring"])
def dense_variant(request, monkeypatch):
    monkeypatch.setenv("AFB_DENSE_VARIANT", request.param)
    return request.param


@pytest.mark.parametrize("d", [20, 128, 784, 1000, 2000])
@pytest.mark.parametrize("loss", ["least_squares", "logistic", "hinge", "least_squares_half"])
def test_dense_grad_full(d, loss, dense_variant):
    from asyncframework_b200.ops.grad import grad
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = _dense_shard(3000, d, seed=d)
    w = torch.randn(d, generator=torch.Generator().manual_seed(1)) * 0.1
    g, l, n = grad(sh, w.cuda(), LOSS_IDS[loss])
    gr, lr, nr, _ = _ref_grad(sh, w, loss)
    assert n == nr
    assert abs(l - lr) <= 2e-4 * max(1.0, abs(lr))
    assert torch.allclose(g.cpu().double(), gr, rtol=2e-4, atol=2e-3 * math.sqrt(d))


def test_dense_grad_sampled_matches_cpu_sampler(dense_variant):
    from asyncframework_b200.core.sampler import bernoulli_mask
    from asyncframework_b200.ops.grad import grad
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = _dense_shard(5000, 784, seed=3)
    w = torch.randn(784) * 0.05
    for seed, frac in [(43, 0.1), (1234567, 0.5), (7, 0.013)]:
        g, l, n = grad(sh, w.cuda(), LOSS_IDS["logistic"], fraction=frac, seed=seed)
        rows = torch.from_numpy(np.nonzero(bernoulli_mask(seed, sh.gidx.cpu(), frac))[0])
        gr, lr, nr, _ = _ref_grad(sh, w, "logistic", rows)
        assert n == nr == rows.numel()
        assert torch.allclose(g.cpu().double(), gr, rtol=2e-4, atol=5e-3)
        assert abs(l - lr) <= 2e-4 * max(1.0, abs(lr))


def test_dense_grad_saga_history(dense_variant):
    from asyncframework_b200.ops.grad import grad
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = _dense_shard(2003, 100, seed=5, labels01=False)
    w = torch.randn(100) * 0.1
    hist = torch.randn(2003).cuda() * 0.2
    h0 = hist.clone()
    g, l, n = grad(sh, w.cuda(), LOSS_IDS["least_squares"], history=hist)
    gr, lr, nr, coeff = _ref_grad(sh, w, "least_squares", None, h0)
    assert torch.allclose(g.cpu().double(), gr, rtol=2e-4, atol=5e-3)
    assert torch.allclose(hist.cpu().double(), coeff, rtol=1e-4, atol=1e-4)   # K7: hist <- new residual


@pytest.mark.parametrize("loss", ["least_squares", "logistic"])
def test_csr_grad(loss):
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.grad import grad
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = synthetic_shard("rcv1", 4000, 0, 1, seed=3, device="cuda", num_features=5000)
    w = torch.randn(5000) * 0.3
    g, l, n = grad(sh, w.cuda(), LOSS_IDS[loss], fraction=0.3, seed=99)
    from asyncframework_b200.core.sampler import bernoulli_mask
    rows = torch.from_numpy(np.nonzero(bernoulli_mask(99, sh.gidx.cpu(), 0.3))[0])
    gr, lr, nr, _ = _ref_grad(sh, w, loss, rows)
    assert n == nr
    assert torch.allclose(g.cpu().double(), gr, rtol=2e-4, atol=2e-3)
    assert abs(l - lr) <= 2e-4 * max(1.0, abs(lr))


def test_level1_and_updaters():
    from asyncframework_b200.ops import level1, updaters
    x = torch.randn(47236).cuda()
    y = torch.randn(47236).cuda()
    y0 = y.clone()
    level1.axpy_(0.37, x, y)
    assert torch.allclose(y, y0 + 0.37 * x, atol=1e-6)
    assert abs(level1.dot(x, y) - float(torch.dot(x.double(), y.double()))) < 1e-2
    z = x.clone()
    level1.scal_(-2.5, z)
    assert torch.allclose(z, -2.5 * x)
    w, g = torch.randn(1003).cuda(), torch.randn(1003).cuda()
    assert torch.allclose(updaters.simple_update(w, g, 0.1), w - 0.1 * g, atol=1e-6)
    w1, l1 = updaters.l1_update(w, g, 0.1, 0.5)
    t = w - 0.1 * g
    ref = torch.sign(t) * torch.clamp(t.abs() - 0.05, min=0)
    assert torch.allclose(w1, ref, atol=1e-6) and abs(l1 - float(ref.abs().sum())) < 1e-2
    w2, sq = updaters.l2_update(w, g, 0.1, 0.5)
    ref2 = w * (1 - 0.05) - 0.1 * g
    assert torch.allclose(w2, ref2, atol=1e-6) and abs(sq - float((ref2.double() ** 2).sum())) < 1e-2


def test_asgd_round_single_gpu_semantics(dense_variant):
    """One worker, async mode: every launch ticks clock and k, applies w += -gamma/sqrt(k+1)*g/parRecs."""
    from asyncframework_b200.core.sampler import bernoulli_mask
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, d, b, gamma, seed = 4000, 784, 0.25, 0.5, 42
    sh = _dense_shard(n, d, seed=11)
    layout = ServerLayout(sh.ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, "cuda")
    blk.reset()
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo="asgd", loss_id=LOSS_IDS["logistic"],
                     num_workers=1, worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed, max_iters=3)
    w_ref = torch.zeros(d, dtype=torch.float64)
    for k in range(3):
        rk.launch()
        torch.cuda.synchronize()
        rows = torch.from_numpy(np.nonzero(bernoulli_mask(seed + k + 1, sh.gidx.cpu(), b))[0])
        g, _, cnt, _ = _ref_grad(sh, w_ref, "logistic", rows)
        w_ref = w_ref - gamma / math.sqrt(k // 1 + 1) * g / (b * n / 1)
        ctl = rk.read_ctl()
        assert ctl["last_accepted"] == 1 and ctl["last_batch"] == cnt and ctl["last_k_after"] == k + 1
        assert torch.allclose(blk.w[:d].cpu().double(), w_ref, rtol=1e-3, atol=1e-5)
    st = blk.read_sctl()
    assert st["k"] == 3 and st["clock"] == 3 and st["avail_count"] == 1
    rk.launch()                      # k >= max_iters: the kernel must stop without touching w
    torch.cuda.synchronize()
    assert rk.read_ctl()["last_accepted"] == -1
    assert torch.allclose(blk.w[:d].cpu().double(), w_ref, rtol=1e-3, atol=1e-5)
    ring = blk.read_ring()
    assert [int(r["seq"]) for r in ring[:3]] == [1, 2, 3] and all(int(r["accepted"]) == 1 for r in ring[:3])


def test_asaga_round_single_gpu_semantics(dense_variant):
    from asyncframework_b200.core.sampler import bernoulli_mask
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, d, b, gamma, seed = 3000, 200, 0.2, 0.05, 7
    sh = _dense_shard(n, d, seed=13, labels01=False)
    layout = ServerLayout(sh.ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, "cuda")
    blk.reset()
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo="asaga", loss_id=LOSS_IDS["least_squares"],
                     num_workers=1, worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed)
    w = torch.zeros(d, dtype=torch.float64)
    abar = torch.zeros(d, dtype=torch.float64)
    hist = torch.zeros(n, dtype=torch.float64)
    for k in range(4):
        rk.launch()
        torch.cuda.synchronize()
        rows = torch.from_numpy(np.nonzero(bernoulli_mask(seed + k + 1, sh.gidx.cpu(), b))[0])
        g, _, cnt, coeff = _ref_grad(sh, w, "least_squares", rows, hist)
        hist[rows] = coeff
        w = w - gamma * (g / (b * n) + abar)
        abar = abar + g / n
        assert torch.allclose(blk.w[:d].cpu().double(), w, rtol=2e-3, atol=1e-5)
        assert torch.allclose(blk.abar[:d].cpu().double(), abar, rtol=2e-3, atol=1e-6)
    assert torch.allclose(rk.hist.cpu().double(), hist, rtol=1e-3, atol=1e-4)
