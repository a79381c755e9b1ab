"""Tiny repro driver for the stream kernel (run under compute-sanitizer on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_gpu_stream import _dense_shard, _kernel  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "serial"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 784
n = 3000
sh = _dense_shard(n, d, seed=3)
if mode == "bf16":
    from asyncframework_b200.data.batch import Shard
    ld = (d + 7) // 8 * 8
    X = torch.zeros(n, ld)
    X[:, :d] = torch.randint(0, 255, (n, d)).float()
    sh = Shard(d, sh.y, X.to(torch.bfloat16).cuda(), gidx=sh.gidx)
rk, blk = _kernel(sh, algo="asgd", loss="logistic", n=n, b=0.25, gamma=0.05, seed=5, pipeline=(mode == "pipe"))
print("variant", rk.variant_effective, "grid", rk.grid, flush=True)
rk.launch_rounds(3)
torch.cuda.synchronize()
print("k", blk.read_sctl()["k"], "abort", int(rk.read_ctl_raw()["abort"]), flush=True)
