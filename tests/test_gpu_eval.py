"""Objective-evaluation kernels (K12) vs a PyTorch fp64 reference."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(shard, W, loss):
    from asyncframework_b200.optim.gradient import make_gradient
    s = shard.to("cpu", torch.float64)
    Wd = W.double().cpu()
    M = torch.stack([s.matvec(Wd[:, t]) for t in range(Wd.shape[1])], 1)
    _, L = make_gradient(loss).coefficients(M, s.y[:, None].expand_as(M))
    return L.sum(0)


@pytest.mark.parametrize("loss", ["least_squares", "logistic"])
@pytest.mark.parametrize("T", [1, 5, 40])
def test_eval_dense_cuda_core(loss, T):
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = synthetic_shard("epsilon", 3001, 0, 1, seed=2, device="cuda", num_features=300,
                         labels="binary01" if loss == "logistic" else "regression")
    W = torch.randn(300, T, device="cuda") * 0.3
    out = eval_loss(sh, W, LOSS_IDS[loss], use_tensor_cores=False)
    ref = _ref(sh, W, loss)
    assert torch.allclose(out.cpu(), ref, rtol=2e-4, atol=1e-3)


def test_eval_csr():
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = synthetic_shard("rcv1", 2500, 0, 1, seed=4, device="cuda", num_features=3000, labels="regression")
    W = torch.randn(3000, 7, device="cuda") * 0.2
    out = eval_loss(sh, W, LOSS_IDS["least_squares"])
    assert torch.allclose(out.cpu(), _ref(sh, W, "least_squares"), rtol=2e-4, atol=1e-3)


@pytest.mark.skipif(os.environ.get("AFB_TEST_TC", "1") != "1", reason="tcgen05 path disabled")
@pytest.mark.parametrize("T,d,n", [(16, 784, 5000), (40, 784, 1000), (9, 100, 300), (32, 2000, 4097)])
def test_eval_tcgen05_gemm(T, d, n):
    """tcgen05/TMEM GEMM variant: exact-ish on integer-pixel data (tf32 represents it exactly),
    tf32-level tolerance on generic data."""
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    shape = "mnist8m" if d == 784 else "epsilon"
    sh = synthetic_shard(shape, n, 0, 1, seed=6, device="cuda", num_features=d, labels="regression")
    scale = 1e-3 if shape == "mnist8m" else 0.3
    W = torch.randn(d, T, device="cuda") * scale
    out = eval_loss(sh, W, LOSS_IDS["least_squares"], use_tensor_cores=True)
    ref = _ref(sh, W, "least_squares")
    tol = 2e-5 if shape == "mnist8m" else 3e-3
    assert torch.allclose(out.cpu(), ref, rtol=tol, atol=tol * float(ref.abs().max()))
