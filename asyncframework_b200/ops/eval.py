"""Objective evaluation (K12) on HBM-resident shards: Σ_i loss(x_i·w_t, y_i) for T snapshots in one
pass.  Dense shards with >= 8 snapshots go to the tcgen05 GEMM kernel for every objective (eval_tc.cu: squared error,
logistic, hinge, half-squared applied in the TMEM epilogue; X consumed as tf32, parameters split hi/lo => ~1e-3 relative
accuracy on generic data, exact on integer/bf16-representable data); set AFB_EVAL_TC=0 to force the exact-fp32 CUDA-core
kernels (eval.cu), which also serve CSR shards and small evaluations."""
from __future__ import annotations

import os

import torch

from ..data.batch import Shard
from . import check, current_stream_ptr, kernels


def eval_loss(shard: Shard, W: torch.Tensor, loss_id: int, use_tensor_cores=None) -> torch.Tensor:
    """``W``: [d, T] snapshots.  Returns float64[T] loss sums over the shard's rows."""
    if shard.is_dense and shard.X.dtype == torch.bfloat16:
        # bf16 row storage: the evaluation kernels consume fp32 tiles — widen the shard in chunks (evaluation is not a
        # hot path; the shard itself stays bf16 in HBM).  The chunk row stride must be a multiple of 4 floats.
        d, T = W.shape
        out = torch.zeros(T, dtype=torch.float64, device=shard.device)
        step = 262_144
        for a in range(0, shard.num_rows, step):
            b = min(shard.num_rows, a + step)
            sub = Shard(shard.num_features, shard.y[a:b].contiguous(), shard.X[a:b].to(torch.float32))
            out += eval_loss(sub, W, loss_id, use_tensor_cores)
        return out
    lib = kernels()
    d, T = W.shape
    ld = shard.ld if shard.is_dense else (shard.num_features + 3) // 4 * 4
    Wt = torch.zeros(T, ld, dtype=torch.float32, device=shard.device)
    Wt[:, :d] = W.t().to(torch.float32)
    out = torch.zeros(T, dtype=torch.float64, device=shard.device)
    if shard.is_dense:
        if use_tensor_cores is None:
            use_tensor_cores = os.environ.get("AFB_EVAL_TC", "1") == "1" and T >= 8 and shard.num_rows >= 128
        if use_tensor_cores:
            # tcgen05 / TMEM / TMA GEMM with the loss applied in the TMEM -> register epilogue (csrc/kernels/eval_tc.cu)
            scratch = torch.empty(64 * ld, dtype=torch.float32, device=shard.device)
            check(lib.afb_eval_loss_tc(shard.X.data_ptr(), shard.y.data_ptr(), shard.num_rows, ld, d, Wt.data_ptr(), T, loss_id,
                                       out.data_ptr(), scratch.data_ptr(), current_stream_ptr()), "eval_loss_tc")
            return out
        check(lib.afb_eval_mse_dense(shard.X.data_ptr(), shard.y.data_ptr(), shard.num_rows, ld, d, Wt.data_ptr(), T,
                                     loss_id, out.data_ptr(), current_stream_ptr()), "eval_mse_dense")
    else:
        check(lib.afb_eval_mse_csr(shard.indptr.data_ptr(), shard.indices.data_ptr(), shard.values.data_ptr(),
                                   shard.y.data_ptr(), shard.num_rows, ld, Wt.data_ptr(), T, loss_id, out.data_ptr(),
                                   current_stream_ptr()), "eval_mse_csr")
    return out
