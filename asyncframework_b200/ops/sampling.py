"""Bernoulli row sampling as a standalone op (K11) on HBM-resident shards: stream-compacts the
selected local row positions with the same counter hash the fused kernels evaluate on the fly
(csrc/kernels/misc.cu: count -> scan -> write).  Backs ``ShardedDataset.sample`` on CUDA shards."""
from __future__ import annotations

from typing import Optional

import torch

from ..core.sampler import threshold24
from . import check, current_stream_ptr, kernels


def sample_rows(n_rows: int, fraction: float, seed: int, gidx: Optional[torch.Tensor] = None, row_base: int = 0,
                device=None) -> torch.Tensor:
    """int64 tensor of the selected local row positions (ascending)."""
    dev = gidx.device if gidx is not None else torch.device(device)
    if n_rows == 0:
        return torch.zeros(0, dtype=torch.int64, device=dev)
    nblocks = (n_rows + 255) // 256
    counts = torch.empty(nblocks, dtype=torch.int32, device=dev)
    out = torch.empty(n_rows, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    g = gidx.contiguous() if gidx is not None else None
    seed_signed = seed if seed < (1 << 63) else seed - (1 << 64)
    check(kernels().afb_sample_rows_v2(g.data_ptr() if g is not None else None, int(row_base), int(n_rows), int(seed_signed),
                                       threshold24(fraction), counts.data_ptr(), out.data_ptr(), total.data_ptr(),
                                       current_stream_ptr()), "sample_rows")
    return out[: int(total.item())]
