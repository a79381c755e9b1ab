"""Small single-GPU drivers for `ncu` captures of kernels that bench.py does not launch (run under scripts/gpu_runs/r2_final_n1.sh)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
what = sys.argv[1]
if what in ("csr", "csr_split"):
    # rcv1-shape worker of an 8-GPU run: 84 675 rows x 47 236 features, ~74 nnz/row, sampling 0.05
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.round import RoundKernel, ServerBlock, ServerLayout, ServerPointers
    from asyncframework_b200.optim.gradient import LOSS_IDS
    n, d = 84_675, 47_236
    sh = synthetic_shard("rcv1", n, 0, 8, seed=3, device="cuda", num_features=d, labels="binary01")
    ld = (d + 3) // 4 * 4
    layout = ServerLayout(ld, 1, ring_size=64)
    blk = ServerBlock.allocate(layout, "cuda")
    blk.reset()
    rk = RoundKernel(sh, ServerPointers(blk.buf.data_ptr(), layout), algo="asgd", loss_id=LOSS_IDS["logistic"], num_workers=1, worker_id=0,
                     total_rows=n, batch_rate=0.05, gamma=0.5, seed=5, variant="simple" if what == "csr_split" else None)
    for _ in range(3):
        rk.launch_rounds(8)
    torch.cuda.synchronize()
    if len(sys.argv) > 2 and sys.argv[2] == "time":
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rk.launch_rounds(400)
        e1.record()
        torch.cuda.synchronize()
        print(f"CSR_ROUND_US grid={rk.grid} variant={rk.variant_effective} {e0.elapsed_time(e1) * 1e3 / 400:.2f}")
    print("k", blk.read_sctl()["k"], rk.variant_effective)
elif what == "eval":
    from asyncframework_b200.data.generators import synthetic_shard
    from asyncframework_b200.ops.eval import eval_loss
    from asyncframework_b200.optim.gradient import LOSS_IDS
    sh = synthetic_shard("mnist8m", 2_000_000, 0, 1, seed=5, device="cuda", labels="binary01")
    W = torch.randn(sh.num_features, 32, device="cuda") * 1e-3
    for _ in range(3):
        out = eval_loss(sh, W, LOSS_IDS["logistic"])
    torch.cuda.synchronize()
    print(out[:3].tolist())
