#!/usr/bin/env python
"""K12 objective-evaluation kernels: tcgen05/TMEM GEMM vs exact CUDA-core kernel (single GPU).
Reports ms, effective HBM GB/s (X is read once per pass of <=32 snapshots) and useful TFLOP/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from asyncframework_b200.data.generators import synthetic_shard  # noqa: E402
from asyncframework_b200.ops.eval import eval_loss  # noqa: E402


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--T", type=int, nargs="+", default=[16, 32, 64])
    ap.add_argument("--only", default=None, choices=[None, "tc", "cc"])
    args = ap.parse_args()
    sh = synthetic_shard("mnist8m", args.rows, 0, 1, seed=1, device="cuda", labels="regression")
    d = sh.num_features
    out = []
    for T in args.T:
        W = torch.randn(d, T, device="cuda") * 1e-3
        rec = {"rows": args.rows, "d": d, "T": T, "x_bytes": sh.X.numel() * 4}
        passes = (T + 31) // 32
        if args.only in (None, "tc"):
            ms = timeit(lambda: eval_loss(sh, W, 0, use_tensor_cores=True))
            rec["tc_ms"] = ms
            rec["tc_hbm_gbps"] = rec["x_bytes"] * passes / ms / 1e6
            rec["tc_useful_tflops"] = 2.0 * args.rows * d * T / ms / 1e9
            rec["tc_issued_tflops_tf32"] = 2.0 * args.rows * 800 * (2 * (16 if T <= 16 else 32)) * passes / ms / 1e9
        if args.only in (None, "cc"):
            ms = timeit(lambda: eval_loss(sh, W, 0, use_tensor_cores=False), iters=2)
            rec["cc_ms"] = ms
            rec["cc_useful_tflops"] = 2.0 * args.rows * d * T / ms / 1e9
        if "tc_ms" in rec and "cc_ms" in rec:
            a = eval_loss(sh, W, 0, use_tensor_cores=True)
            b = eval_loss(sh, W, 0, use_tensor_cores=False)
            rec["max_rel_diff"] = float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
            rec["speedup"] = rec["cc_ms"] / rec["tc_ms"]
        out.append(rec)
        print("EVAL_RESULT " + json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
