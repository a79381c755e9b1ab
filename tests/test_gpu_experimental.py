"""Opt-in GPU tests of csrc/experimental (persistent multi-round kernel).  Skipped unless AFB_TEST_EXPERIMENTAL=1: the kernel
spins on device memory between rounds, so it is only ever run under an explicit `timeout` until it has been validated on
hardware (`AFB_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_gpu_experimental.py -x -q`)."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("AFB_TEST_EXPERIMENTAL", "0") != "1", reason="set AFB_TEST_EXPERIMENTAL=1")]


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 784), ("asaga", "least_squares", 200), ("asgd", "logistic", 1000)])
def test_persistent_rounds_match_single_round_launches(algo, loss, d):
    """n rounds in ONE persistent launch == n launches of the one-round kernel (single worker: deterministic)."""
    from tests.test_gpu_kernels import _dense_shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, b, gamma, seed, rounds = 6000, 0.25, 0.05, 5, 7
    sh = _dense_shard(n, d, seed=3, labels01=(loss == "logistic"))
    res = {}
    for variant in ("simple", "persist"):
        layout = ServerLayout(sh.ld, 1, ring_size=64)
        blk = ServerBlock.allocate(layout, "cuda")
        blk.reset()
        rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo=algo, loss_id=LOSS_IDS[loss], num_workers=1,
                         worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed, variant=variant, max_iters=rounds - 1)
        assert rk.persist == (variant == "persist")
        rk.launch_rounds(rounds)              # the last round hits k >= max_iters: the persistent grid must exit cleanly
        torch.cuda.synchronize()
        st = blk.read_sctl()
        res[variant] = (blk.w.clone(), blk.abar.clone(), st["k"], st["clock"], rk.read_ctl()["rounds_done"], rk.launches)
    a, c = res["simple"], res["persist"]
    assert a[2] == c[2] == rounds - 1 and a[3] == c[3]
    assert c[5] == 1 and a[5] == rounds
    assert torch.allclose(a[0], c[0], rtol=1e-4, atol=1e-6) and torch.allclose(a[1], c[1], rtol=1e-4, atol=1e-6)


def test_persistent_launch_refuses_non_resident_grid():
    from tests.test_gpu_kernels import _dense_shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = _dense_shard(2000, 784, seed=3)
    layout = ServerLayout(sh.ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, "cuda")
    blk.reset()
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo="asgd", loss_id=LOSS_IDS["logistic"], num_workers=1,
                     worker_id=0, total_rows=2000, batch_rate=0.5, gamma=0.1, seed=1, variant="persist")
    with pytest.raises(RuntimeError):
        rk.launch_rounds(4, share=2)          # full-capacity grid x 2 co-running grids cannot be resident
    torch.cuda.synchronize()


@pytest.mark.parametrize("d", [784, 200, 1000])
def test_pipelined_rounds_match_the_simulator_semantics(d):
    """Single worker => deterministic: n pipelined rounds on the GPU == the simulator's transcription of the same schedule
    (round i+1 computed on the parameters pulled before update i landed)."""
    from tests.test_gpu_kernels import _dense_shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, b, gamma, seed, rounds = 6000, 0.25, 0.05, 5, 8
    sh = _dense_shard(n, d, seed=3, labels01=True)
    res = {}
    for dev in ("cuda", "cpu"):
        s = sh if dev == "cuda" else sh.to("cpu")
        layout = ServerLayout(s.ld, 1, ring_size=64)
        blk = ServerBlock.allocate(layout, dev)
        blk.reset()
        rk = RoundKernel(s, ServerPointers(blk.buf.data_ptr(), layout), algo="asgd", loss_id=LOSS_IDS["logistic"], num_workers=1,
                         worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed, variant="pipelined", max_iters=rounds - 2,
                         sim_server_block=blk if dev == "cpu" else None)
        assert rk.pipelined
        rk.launch_rounds(rounds)              # k reaches max_iters inside the launch: the grid must drain and exit
        if dev == "cuda":
            torch.cuda.synchronize()
        st = blk.read_sctl()
        res[dev] = (blk.w.cpu().clone(), st["k"], st["clock"], st["avail_count"], rk.read_ctl()["rounds_done"])
    g, c = res["cuda"], res["cpu"]
    assert g[1] == rounds - 2 and g[3] == 1          # accepted updates, worker idle again
    assert g[1] <= c[1] + 1 and abs(g[2] - c[2]) <= 1  # the device may start one more round before it sees k >= max_iters
    assert torch.allclose(g[0], c[0], rtol=2e-3, atol=1e-5)


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 47236), ("asaga", "least_squares", 6000)])
def test_persistent_csr_round_matches_the_split_round(algo, loss, d):
    """Wide sparse model: n rounds in ONE persistent launch with in-kernel grid barriers (multi-CTA pull / push) == n split
    rounds (gate, pull, grad, push kernels).  Single worker; float atomics reorder sums, hence the tolerance."""
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, b, gamma, seed, rounds = 20000, 0.2, 0.05, 5, 6
    sh = synthetic_shard("rcv1", n, 0, 1, seed=3, device="cuda", num_features=d,
                         labels="binary01" if loss == "logistic" else "regression")
    ld = (d + 3) // 4 * 4
    res = {}
    for variant in ("simple", "persist"):
        layout = ServerLayout(ld, 1, ring_size=64)
        blk = ServerBlock.allocate(layout, "cuda")
        blk.reset()
        rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo=algo, loss_id=LOSS_IDS[loss], num_workers=1,
                         worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed, variant=variant, max_iters=rounds - 1)
        assert rk.persist_csr == (variant == "persist")
        rk.launch_rounds(rounds)
        torch.cuda.synchronize()
        st = blk.read_sctl()
        res[variant] = (blk.w.clone(), blk.abar.clone(), st["k"], st["clock"], st["avail_count"], rk.launches,
                        None if rk.hist is None else rk.hist.clone())
    a, c = res["simple"], res["persist"]
    assert a[2] == c[2] == rounds - 1 and a[3] == c[3] and c[4] == 1 and c[5] == 1
    assert torch.allclose(a[0], c[0], rtol=1e-3, atol=1e-5) and torch.allclose(a[1], c[1], rtol=1e-3, atol=1e-6)
    if a[6] is not None:
        assert torch.allclose(a[6], c[6], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("algo,loss,d", [("asgd", "logistic", 784), ("asaga", "least_squares", 200), ("asgd", "logistic", 2000)])
def test_bf16_row_storage_kernel_matches_fp32_kernel_on_bf16_exact_data(algo, loss, d):
    """Rows that are exactly representable in bf16 (integer pixels) must give the same rounds from the bf16-storage kernel
    as from the fp32 kernel (both accumulate in fp32; only the summation order inside a row differs)."""
    from asyncframework_b200.data.batch import Shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
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
        layout = ServerLayout(ld, 1, ring_size=64)
        blk = ServerBlock.allocate(layout, "cuda")
        blk.reset()
        rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo=algo, loss_id=LOSS_IDS[loss], num_workers=1,
                         worker_id=0, total_rows=n, batch_rate=b, gamma=gamma, seed=seed)
        assert rk.bf16_rows == (dt == torch.bfloat16)
        for _ in range(rounds):
            rk.launch()
        torch.cuda.synchronize()
        st = blk.read_sctl()
        res[dt] = (blk.w.clone(), st["k"], rk.read_ctl()["last_batch"])
    a, c = res[torch.float32], res[torch.bfloat16]
    assert a[1] == c[1] == rounds and a[2] == c[2]
    assert torch.allclose(a[0], c[0], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("loss", ["logistic", "hinge", "least_squares_half"])
@pytest.mark.parametrize("T", [8, 20, 40])
def test_tensor_core_evaluation_with_loss_epilogue_matches_the_exact_kernel(loss, T, monkeypatch):
    """csrc/experimental/eval_tc_loss.cu: tcgen05 GEMM + logistic / hinge epilogue vs the exact CUDA-core kernel, on
    bf16-exact data (where the tf32 product is exact, so only the fp32 loss arithmetic differs)."""
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    loss_id = LOSS_IDS.get(loss, 3)
    sh = synthetic_shard("mnist8m", 20_000, 0, 1, seed=5, device="cuda", labels="binary01")     # integer pixels: exact in tf32
    W = (torch.randn(sh.num_features, T, device="cuda") * 1e-3)
    exact = eval_loss(sh, W, loss_id, use_tensor_cores=False)
    monkeypatch.setenv("AFB_EVAL_TC_LOSS", "1")
    tc = eval_loss(sh, W, loss_id, use_tensor_cores=True)
    torch.cuda.synchronize()
    assert torch.allclose(tc.cpu(), exact.cpu(), rtol=2e-4, atol=1e-2), (tc[:4], exact[:4])
