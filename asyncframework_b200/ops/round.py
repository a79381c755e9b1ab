"""Launcher for the fused round kernels (csrc/kernels/dense_round.cu, csr_round.cu).

One ``RoundKernel`` = one logical worker bound to one HBM-resident ``Shard``.  It owns the
worker-local scratch (pulled ``w`` copy, gradient accumulator, control block, SAGA history) and
holds raw pointers to the *server block* — a byte region laid out by ``ServerLayout`` that lives
either in ordinary device memory (single process) or in rank 0's symmetric memory, in which case
the pointers are NVLink peer pointers and the same kernel performs the push/pull over NVSwitch.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch

from ..core.sampler import threshold24
from ..data.batch import Shard
from . import check, kernels
from ._signatures import (REG_TYPES, RESULT_REC_DTYPE, ROUND_DESC_DTYPE, RoundParams, SERVER_CTL_DTYPE, SyncPeers,
                          WORKER_CTL_DTYPE, WORKER_SLOT_DTYPE)

ALGO = {"grad_only": 0, "asgd": 1, "asaga": 2, "sgd_sync": 3, "saga_sync": 4}
NEVER_STALE = 1 << 40


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


@dataclass
class ServerLayout:
    """Byte layout of the server block.  Identical on every rank (symmetric allocation)."""
    ld: int
    num_workers: int
    ring_size: int = 4096
    history_slots: int = 0

    def __post_init__(self):
        off = 0
        self.off_w = off; off = _align(off + self.ld * 4)
        self.off_abar = off; off = _align(off + self.ld * 4)
        self.off_sctl = off; off = _align(off + SERVER_CTL_DTYPE.itemsize)
        self.off_status = off; off = _align(off + self.num_workers * WORKER_SLOT_DTYPE.itemsize)
        self.off_ring = off; off = _align(off + self.ring_size * RESULT_REC_DTYPE.itemsize)
        self.slot_stride = self.ld + 4
        self.off_gstage = off; off = _align(off + 2 * self.num_workers * self.slot_stride * 4)
        self.off_flags = off; off = _align(off + 2 * self.num_workers * 4)
        # AsyncBroadcast: ring of K parameter versions + {version counter, kernel scratch} words
        self.off_version = off; off = _align(off + 16)
        self.off_hist_w = off; off = _align(off + max(self.history_slots, 0) * self.ld * 4)
        self.nbytes = off


class ServerBlock:
    """Typed views over a server block that lives on *this* device (rank 0's copy is the
    authoritative parameter server; in BSP modes every rank's copy holds its local replica)."""

    def __init__(self, layout: ServerLayout, buf: torch.Tensor):
        assert buf.dtype == torch.uint8 and buf.numel() >= layout.nbytes
        self.layout = layout
        self.buf = buf
        L = layout
        self.w = buf[L.off_w:L.off_w + L.ld * 4].view(torch.float32)
        self.abar = buf[L.off_abar:L.off_abar + L.ld * 4].view(torch.float32)
        self.sctl = buf[L.off_sctl:L.off_sctl + SERVER_CTL_DTYPE.itemsize]
        self.status = buf[L.off_status:L.off_status + L.num_workers * WORKER_SLOT_DTYPE.itemsize]
        self.ring = buf[L.off_ring:L.off_ring + L.ring_size * RESULT_REC_DTYPE.itemsize]
        self.gstage = buf[L.off_gstage:L.off_gstage + 2 * L.num_workers * L.slot_stride * 4].view(torch.float32)
        self.flags = buf[L.off_flags:L.off_flags + 2 * L.num_workers * 4].view(torch.int32)
        self.version_words = buf[L.off_version:L.off_version + 16].view(torch.int32)
        self.hist_w = (buf[L.off_hist_w:L.off_hist_w + L.history_slots * L.ld * 4].view(torch.float32).view(L.history_slots, L.ld)
                       if L.history_slots > 0 else None)

    @classmethod
    def allocate(cls, layout: ServerLayout, device) -> "ServerBlock":
        return cls(layout, torch.zeros(layout.nbytes, dtype=torch.uint8, device=device))

    def reset(self, w0: Optional[torch.Tensor] = None) -> None:
        self.buf.zero_()
        if w0 is not None:
            self.w[: w0.numel()].copy_(w0.to(torch.float32))
        st = np.zeros(self.layout.num_workers, dtype=WORKER_SLOT_DTYPE)
        st["available"] = 1
        st["alive"] = 1
        self.status.copy_(torch.from_numpy(st.view(np.uint8)).to(self.buf.device))
        sc = np.zeros(1, dtype=SERVER_CTL_DTYPE)
        sc["avail_count"] = self.layout.num_workers
        sc["num_alive"] = self.layout.num_workers
        sc["enable_mask"] = np.uint64(0xFFFFFFFFFFFFFFFF)          # every worker enabled until a predicate says otherwise
        self.sctl.copy_(torch.from_numpy(sc.view(np.uint8)).to(self.buf.device))

    # host-side reads (sync D2H of a few hundred bytes)
    def read_sctl(self) -> Dict[str, int]:
        a = self.sctl.cpu().numpy().view(SERVER_CTL_DTYPE)[0]
        ck = int(a["clock_k"])
        return {"clock": ck & 0xFFFFFFFF, "k": ck >> 32, "avail_count": int(a["avail_count"]), "stop": int(a["stop"]),
                "avg_delay_ns": int(a["avg_delay_ns"]), "cul_count": int(a["cul_count"]), "gate_dyn": int(a["gate_dyn"]),
                "num_alive": int(a["num_alive"]), "enable_mask": [int(x) for x in a["enable_mask"]]}

    def _poke(self, field: str, values: np.ndarray) -> None:
        """Overwrite ONE field of the server control word in place (in-flight counters next to it are not clobbered)."""
        off = self.layout.off_sctl + SERVER_CTL_DTYPE.fields[field][1]
        raw = np.ascontiguousarray(values).view(np.uint8)
        self.buf[off:off + raw.size].copy_(torch.from_numpy(raw.copy()).to(self.buf.device))

    def set_enable_mask(self, enabled) -> None:
        """ASYNCbarrier bridge: bit q of the mask = worker q may start rounds (workers with a clear bit skip theirs)."""
        words = np.zeros(4, dtype=np.uint64)
        for q, on in enumerate(enabled):
            if on:
                words[q >> 6] |= np.uint64(1) << np.uint64(q & 63)
        self._poke("enable_mask", words)

    def set_gate_dyn(self, threshold: int, num_alive: int) -> None:
        self._poke("gate_dyn", np.asarray([int(threshold), int(num_alive)], dtype=np.int32))

    def add_avail(self, delta: int) -> None:
        """Host-side correction of the idle-worker counter (watchdog: a worker that died while idle no longer counts)."""
        a = self.sctl.cpu().numpy().view(SERVER_CTL_DTYPE)[0]
        self._poke("avail_count", np.asarray([int(a["avail_count"]) + int(delta)], dtype=np.int32))

    def set_progress(self, k: int, clock: int) -> None:
        """Restore the accepted-update counter and the logical clock (checkpoint resume)."""
        self._poke("clock_k", np.asarray([(int(k) << 32) | (int(clock) & 0xFFFFFFFF)], dtype=np.uint64))

    def read_status(self) -> np.ndarray:
        return self.status.cpu().numpy().view(WORKER_SLOT_DTYPE).copy()

    def read_ring(self) -> np.ndarray:
        return self.ring.cpu().numpy().view(RESULT_REC_DTYPE).copy()

    def request_stop(self) -> None:
        # only the stop word is rewritten so in-flight counters are not clobbered
        self._poke("stop", np.asarray([1], dtype=np.int32))


@dataclass
class ServerPointers:
    """Raw addresses of a server block as seen from the launching rank (local or peer-mapped)."""
    base: int
    layout: ServerLayout

    @property
    def w(self): return self.base + self.layout.off_w
    @property
    def abar(self): return self.base + self.layout.off_abar
    @property
    def sctl(self): return self.base + self.layout.off_sctl
    @property
    def status(self): return self.base + self.layout.off_status
    @property
    def ring(self): return self.base + self.layout.off_ring
    @property
    def gstage(self): return self.base + self.layout.off_gstage
    @property
    def flags(self): return self.base + self.layout.off_flags


class _SimLib:
    """Stand-in for libafb_kernels.so when a RoundKernel runs against the CPU simulator."""

    def afb_dense_round_max_ld(self): return 1 << 30
    def afb_dense_stream_max_ld(self): return 1 << 30
    def afb_dense_stream_ctas_per_sm(self, ld, bf16=0): return 2
    def afb_dense_round_ctas_per_sm(self, ld): return 2
    def afb_dense_round_ring_max_ld(self): return 0
    def afb_dense_round_pipe_max_ld(self): return 0
    afb_dense_round_launch = afb_csr_round_launch = afb_dense_round_ring_launch = afb_dense_round_pipe_launch = None


class RoundKernel:
    """One logical worker's fused rounds.

    ``launch_rounds(n)`` enqueues ``n`` rounds as ONE persistent kernel (csrc/kernels/dense_stream.cu for dense shards,
    csr_stream.cu for wide CSR models): the worker loop runs on the device, rounds are separated by flags in device memory.
    ``launch()`` enqueues a single round.  Dense kernel variants (``variant=`` / ``AFB_DENSE_VARIANT``):

    * ``stream`` (default) — TMA bulk-copy row rings + persistent rounds; ASGD without a gate additionally PIPELINES its
      rounds (round r+1 streams while round r's NVLink tail is in flight; ``pipeline=False`` / ``AFB_PIPELINE=0`` turns
      that off);
    * ``simple`` — the register double-buffered one-round kernel (dense_round.cu; also used for SAGA rollback rounds);
    * ``ring`` / ``pipe`` — earlier one-round TMA variants kept for A/B measurements.
    """

    def __init__(self, shard: Shard, server: ServerPointers, *, algo: str, loss_id: int, num_workers: int,
                 worker_id: int, total_rows: int, batch_rate: float, gamma: float, seed: int,
                 taw: int = NEVER_STALE, max_iters: int = 1 << 40, gate_threshold: int = 0,
                 delay_mode: int = 0, delay_coeff: float = 0.0, warmup_updates: int = 0,
                 local_block: Optional[ServerBlock] = None, peer_blocks: Optional[List[ServerPointers]] = None,
                 grid: Optional[int] = None, seed_mode: Optional[int] = None, rollback: Optional[bool] = None,
                 variant: Optional[str] = None, wctl: Optional[torch.Tensor] = None, sim_server_block=None,
                 reg_param: float = 0.0, reg_type: Optional[str] = None, ssp_bound: int = -1, use_mask: bool = False,
                 pipeline: Optional[bool] = None, multicast_base: int = 0, spin_limit_s: Optional[float] = None):
        # CPU shards run against the pure-Python round simulator (ops/round_sim.py): same byte layouts and
        # protocol, used by the CPU test-suite; CUDA shards always run the hand-written kernels.
        self.sim = shard.device.type != "cuda"
        if self.sim and os.environ.get("AFB_ALLOW_SIM", "1") != "1":
            raise ValueError("RoundKernel needs an HBM-resident shard")
        # bf16 ROW STORAGE (dense shards only): X is bf16, everything else (labels, parameters, gradient, accumulation)
        # stays fp32.  Half the HBM bytes per round.
        self.bf16_rows = shard.is_dense and shard.X.dtype == torch.bfloat16
        if self.bf16_rows:
            if shard.y.dtype != torch.float32:
                raise TypeError("bf16 row storage keeps the labels in fp32")
            if shard.X.shape[1] % 8 != 0:
                raise ValueError("bf16 rows need ld % 8 == 0 (16-byte chunks)")
        elif shard.dtype != torch.float32:
            raise TypeError("fused kernels are fp32 (storage and accumulate), or bf16 row storage with fp32 math")
        self.lib = kernels() if not self.sim else _SimLib()
        self.shard = shard
        self.device = shard.device
        self.algo = algo
        self.num_workers = num_workers
        self.worker_id = worker_id
        d = shard.num_features
        ld = server.layout.ld
        self.ld = ld
        self.is_dense = shard.is_dense
        if self.is_dense:
            if shard.X.shape[1] != ld or not shard.X.is_contiguous():
                raise ValueError(f"dense shard row stride {shard.X.shape[1]} != server ld {ld}")
            if ld > self.lib.afb_dense_round_max_ld():
                raise ValueError(f"dense fused kernel supports ld <= {self.lib.afb_dense_round_max_ld()}; use CSR layout")
        dev = self.device
        self.w_local = torch.zeros(ld, dtype=torch.float32, device=dev)
        self.abar_local = torch.zeros(ld, dtype=torch.float32, device=dev)
        # worker-local gradient: `g_copies` replicas (replica 0 is the one every consumer reads); the stream kernels spread
        # the per-CTA merges over the replicas, every other kernel uses replica 0 only
        self.g_copies = int(os.environ.get("AFB_G_COPIES", "16")) if shard.is_dense else 1
        self.g_all = torch.zeros(self.g_copies, ld, dtype=torch.float32, device=dev)
        self.g_local = self.g_all[0]
        # control block: caller may pass a slice of a contiguous per-rank array (one D2H for all workers)
        self.wctl = wctl if wctl is not None else torch.zeros(WORKER_CTL_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        assert self.wctl.numel() == WORKER_CTL_DTYPE.itemsize and self.wctl.is_contiguous()
        saga = algo in ("asaga", "saga_sync")
        self.hist = torch.zeros(shard.num_rows, dtype=torch.float32, device=dev) if saga else None
        if rollback is None:
            rollback = saga and taw < NEVER_STALE and algo == "asaga"
        self.rollback = bool(rollback)
        self.hist_prev = torch.zeros(shard.num_rows, dtype=torch.float32, device=dev) if self.rollback else None
        self.hist_tag = torch.full((shard.num_rows,), -1, dtype=torch.int32, device=dev) if self.rollback else None
        self.local_block = local_block
        self.peers_buf = None
        p = RoundParams()
        if self.is_dense:
            p.X = shard.X.data_ptr()
        else:
            p.indptr = shard.indptr.data_ptr(); p.indices = shard.indices.data_ptr(); p.values = shard.values.data_ptr()
        p.y = shard.y.data_ptr()
        contiguous_gidx = False
        if shard.gidx is not None and shard.num_rows > 0:
            g0 = int(shard.gidx[0])
            contiguous_gidx = bool((shard.gidx[-1] - g0 + 1 == shard.num_rows).item()) and bool(
                (shard.gidx[1:] - shard.gidx[:-1] == 1).all().item()) if shard.num_rows > 1 else True
            p.row_base = g0
        if shard.gidx is not None and not contiguous_gidx:
            self._gidx = shard.gidx.contiguous()
            p.gidx = self._gidx.data_ptr()
        elif shard.gidx is None:
            p.row_base = worker_id << 40
        if self.hist is not None:
            p.hist = self.hist.data_ptr()
        if self.rollback:
            p.hist_prev = self.hist_prev.data_ptr(); p.hist_tag = self.hist_tag.data_ptr()
        p.ld = ld
        p.w_server = server.w; p.abar_server = server.abar; p.sctl = server.sctl
        p.status = server.status; p.ring = server.ring
        p.w_local = self.w_local.data_ptr(); p.abar_local = self.abar_local.data_ptr()
        p.g_local = self.g_all.data_ptr(); p.wctl = self.wctl.data_ptr(); p.g_copies = self.g_copies
        p.n_rows = shard.num_rows; p.d = d; p.algo = ALGO[algo]; p.ring_size = server.layout.ring_size
        p.num_workers = num_workers; p.worker_id = worker_id
        p.thr24 = threshold24(batch_rate); p.gate_threshold = int(gate_threshold)
        p.seed = int(seed); p.taw = int(min(taw, NEVER_STALE)); p.max_iters = int(max_iters)
        p.gamma = float(gamma)
        if algo in ("sgd_sync", "saga_sync"):
            # BSP replica lives in the rank-local server block so rank 0's copy is the model
            if local_block is None or peer_blocks is None:
                raise ValueError("sync algorithms need local_block and peer_blocks")
            p.w_local = local_block.w.data_ptr(); p.abar_local = local_block.abar.data_ptr()
            p.sctl = local_block.sctl.data_ptr()
            p.inv_par_recs = 1.0 / (batch_rate * total_rows)
            sp = SyncPeers()
            for q, pb in enumerate(peer_blocks):
                sp.gstage[q] = pb.gstage; sp.flags[q] = pb.flags
            if multicast_base and os.environ.get("AFB_NVLS_BSP", "1") == "1":
                # NVLS: the all-gather of the BSP round is ONE multimem.st stream through the multicast alias of the
                # symmetric block (same offsets on every rank) instead of P peer-store loops
                sp.mc_gstage = multicast_base + server.layout.off_gstage
                sp.mc_flags = multicast_base + server.layout.off_flags
            self._sync_peers_host = sp
            self.peers_buf = torch.from_numpy(np.frombuffer(bytes(sp), dtype=np.uint8).copy()).to(dev)
            p.peers = self.peers_buf.data_ptr()
            p.w_server = local_block.w.data_ptr()
        else:
            p.inv_par_recs = 1.0 / (batch_rate * total_rows / num_workers)
        self.nvls_bsp = bool(algo in ("sgd_sync", "saga_sync") and multicast_base and os.environ.get("AFB_NVLS_BSP", "1") == "1")
        p.inv_n = 1.0 / total_rows
        p.delay_coeff = float(delay_coeff); p.delay_mode = int(delay_mode); p.warmup_updates = int(warmup_updates)
        if seed_mode is None:
            # 0: seed + k_read + 1 (the reference's `seed + k + 1`, k read from the server at pull time)
            # 1: seed + round + 1  (worker-local round counter: lets the first row fetch overlap the pull)
            seed_mode = {"asgd": 0, "asaga": 0, "sgd_sync": 1, "saga_sync": 1, "grad_only": 2}[algo]
            if algo in ("asgd", "asaga") and os.environ.get("AFB_SEED_MODE", "round") == "round":
                seed_mode = 1
        p.seed_mode = int(seed_mode); p.loss_id = int(loss_id)
        # K8 fused into the apply (Updater.scala:114-158); ASYNCbarrier predicates on the device
        p.reg_param = float(reg_param); p.reg_type = int(REG_TYPES[reg_type] if not isinstance(reg_type, int) else reg_type)
        if p.reg_type and p.reg_param <= 0.0:
            p.reg_type = 0
        p.ssp_bound = int(ssp_bound) if algo in ("asgd", "asaga") else -1
        p.use_mask = 1 if (use_mask and algo in ("asgd", "asaga")) else 0
        if spin_limit_s is None:
            spin_limit_s = float(os.environ.get("AFB_SPIN_LIMIT_S", "20"))
        p.spin_limit_ns = int(spin_limit_s * 1e9)
        # stream kernel: which ring stages are filled by the TMA unit (one cp.async.bulk per row, ~64 cycles of TMA service
        # per operation) and which by per-lane 16-byte cp.async through the LSU path; alternating stages keeps BOTH copy
        # engines busy (profiles/README.md: all-TMA caps at 5.2 TB/s on 3 KB rows)
        # Measured at N = 1 (profiles/README.md): bf16 rows: all-TMA 3583 it/s > alternating 3499 > all-LSU 3434 (a 1.5 KB row
        # amortises the per-operation TMA cost); fp32 rows: all-LSU 2278 > alternating 2178 > all-TMA 2038 (TMA cap 5.2 TB/s).
        default_mask = "0xFF" if self.bf16_rows else "0x00"
        p.tma_mask = int(os.environ.get("AFB_TMA_MASK", default_mask), 0)
        self.params = p
        self.desc_dev = None                     # set by FusedEngine for the end-to-end path (RoundDesc ring)
        self.num_sms = torch.cuda.get_device_properties(dev).multi_processor_count if not self.sim else 148
        # ---- kernel selection ------------------------------------------------------------------------------------
        variant = variant or os.environ.get("AFB_DENSE_VARIANT", "stream")
        if variant in ("persist", "pipelined"):          # round-1 spellings of the persistent kernels
            variant = "stream"
        self.stream = False                              # persistent TMA-ring kernel (dense)
        self.stream_csr = False                          # persistent one-kernel CSR round
        self._launch = None
        if self.is_dense:
            aligned = (shard.X.data_ptr() % 16 == 0) and (shard.y.data_ptr() % 16 == 0)
            if self.bf16_rows and not (aligned and not self.rollback):
                raise ValueError("bf16 row storage runs on the stream kernel only (16-byte aligned shard, no SAGA rollback rounds)")
            if self.bf16_rows or (variant == "stream" and not self.rollback and aligned and ld <= self.lib.afb_dense_stream_max_ld()):
                self.variant = "stream"
                self.stream = True
                self.ctas_per_sm = self.lib.afb_dense_stream_ctas_per_sm(ld, int(self.bf16_rows))
            elif variant == "ring" and ld <= self.lib.afb_dense_round_ring_max_ld():
                self.variant = "ring"
                self._launch = self.lib.afb_dense_round_ring_launch
                self.ctas_per_sm = 2
            elif variant == "pipe" and ld <= self.lib.afb_dense_round_pipe_max_ld():
                self.variant = "pipe"
                self._launch = self.lib.afb_dense_round_pipe_launch
                self.ctas_per_sm = 2
            else:
                self.variant = "simple"
                self._launch = self.lib.afb_dense_round_launch
                self.ctas_per_sm = self.lib.afb_dense_round_ctas_per_sm(ld)
        else:
            self.variant = "csr"
            self._launch = self.lib.afb_csr_round_launch
            self.ctas_per_sm = 2
        self.grid = int(grid) if grid else self.ctas_per_sm * self.num_sms
        # pipelined rounds: ASGD, no device gate, worker-local sampling seed
        if pipeline is None:
            pipeline = os.environ.get("AFB_PIPELINE", "1") == "1"
        self.pipelined = bool(self.stream and pipeline and algo == "asgd" and int(gate_threshold) == 0 and p.ssp_bound < 0
                              and not p.use_mask and int(p.seed_mode) in (1, 2))
        if self.pipelined:
            nbytes = 192 if self.sim else self.lib.afb_pipe_state_bytes()
            self._pipe_state = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self._pipe_g2 = torch.zeros(2, self.g_copies, ld, dtype=torch.float32, device=dev)
            self._pipe_w2 = torch.zeros(2, ld, dtype=torch.float32, device=dev)
            # pull round r+1's parameters when ~6 us of round r remain: chunks/us = HBM bytes/us / (row bytes * 64 * b)
            row_bytes = ld * shard.X.element_size()
            per_us = 6.5e6 / (row_bytes * 64.0 * max(batch_rate, 1e-6))
            p.pull_ahead_chunks = int(os.environ.get("AFB_PULL_AHEAD_CHUNKS", max(1, int(6.0 * per_us))))
        # wide models (CSR, d = 47 236): async rounds run as ONE persistent kernel with in-kernel grid barriers
        # (multi-CTA pull and push); rollback rounds keep the split launch sequence gate -> pull -> grad -> push
        wide = (not self.is_dense) and algo in ("asgd", "asaga") and ld >= int(os.environ.get("AFB_SPLIT_MIN_LD", "4096"))
        self.stream_csr = bool(wide and not self.rollback and variant != "simple" and (self.sim or hasattr(self.lib, "afb_csr_stream_launch")))
        if self.stream_csr and not self.sim:
            self._grid_bar = torch.zeros(self.lib.afb_grid_bar_bytes(), dtype=torch.uint8, device=dev)
            if not grid:
                # a wide-model round is a few thousand sampled rows: its cost is the three in-kernel grid barriers (one
                # same-address arrival per CTA each) and the NVLink round trips, not the arithmetic — size the grid by the
                # work (~64 sampled rows per 16-warp CTA, every thread still moves one float4 of the pull / push)
                want = int(math.ceil(batch_rate * shard.num_rows / 64.0))
                need_comm = int(math.ceil(ld / 4 / 512.0))           # CTAs needed for one float4 per thread
                self.grid = int(os.environ.get("AFB_CSR_GRID", max(32, min(self.grid, max(want, min(need_comm, self.num_sms))))))
        self.launches = 0
        self.rounds_launched = 0
        if self.sim:
            from .round_sim import RoundSim
            self._sim = RoundSim(self, sim_server_block)
        # programmatic dependent launch (simple dense variant only; measured -3 % at N=8 in round 1: opt-in)
        self.use_pdl = (not self.sim and self.variant == "simple" and not self.rollback and algo in ("asgd", "asaga", "sgd_sync", "saga_sync")
                        and os.environ.get("AFB_PDL", "0") == "1")
        self.split = (not self.sim) and wide and not self.stream_csr
        if self.split:
            pg = RoundParams.from_buffer_copy(bytes(p))
            pg.algo = ALGO["grad_only"]
            pg.seed_mode = 3
            self.params_grad = pg
            self.push_ctl = torch.zeros(32, dtype=torch.uint8, device=dev)
            self.comm_grid = max(1, min(self.num_sms, (ld // 4 + 255) // 256))

    # back-compat spellings used by older call sites / tests
    @property
    def persist(self) -> bool:
        return self.stream and not self.pipelined

    @property
    def persist_csr(self) -> bool:
        return self.stream_csr

    # ------------------------------------------------------------------
    def launch(self, stream=None) -> None:
        """Enqueue ONE round."""
        if self.sim:
            self._sim.launch()
            self.launches += 1
            self.rounds_launched += 1
            return
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        sp = int(s.cuda_stream)
        if self.stream or self.stream_csr:
            self.launch_rounds(1, stream)
            return
        if self.split:
            check(self.lib.afb_gate_launch(C.byref(self.params), sp), "gate")
            check(self.lib.afb_pull_w_launch(C.byref(self.params), self.comm_grid, sp), "pull_w")
            check(self._launch(C.byref(self.params_grad), self.grid, sp), "grad (split round)")
            check(self.lib.afb_push_apply_launch(C.byref(self.params), self.push_ctl.data_ptr(), self.comm_grid, sp), "push_apply")
            self.launches += 4
            self.rounds_launched += 1
            return
        if self.use_pdl and not torch.cuda.is_current_stream_capturing():
            self.params.pdl = 1
            self.params.round_host = self.rounds_launched
        else:
            self.params.pdl = 0
        check(self._launch(C.byref(self.params), self.grid, sp), "round kernel launch")
        self.launches += 1
        self.rounds_launched += 1
        if self.rollback:
            check(self.lib.afb_saga_rollback_launch(self.hist.data_ptr(), self.hist_prev.data_ptr(), self.hist_tag.data_ptr(),
                                                    self.shard.num_rows, self.wctl.data_ptr(), sp), "saga rollback")
            self.launches += 1

    def launch_rounds(self, n: int, stream=None, share: int = 1) -> None:
        """Enqueue ``n`` consecutive rounds: ONE persistent kernel with the stream kernels (``share`` = persistent grids
        co-running on this GPU; the launcher refuses a grid that cannot be fully co-resident, because the in-kernel round
        barrier would deadlock), ``n`` launches with the legacy one-round variants."""
        if n <= 0:
            return
        if self.sim:
            if self.pipelined and n > 1:
                self._sim.launch_pipelined(n)
            else:
                for _ in range(n):                   # protocol-wise a persistent launch IS n back-to-back rounds
                    self._sim.launch()
            self.launches += 1 if (self.stream or self.stream_csr) else n
            self.rounds_launched += n
            return
        if not (self.stream or self.stream_csr):
            for _ in range(n):
                self.launch(stream)
            return
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        sp = int(s.cuda_stream)
        self.params.pdl = 0
        if self.stream_csr:
            check(self.lib.afb_csr_stream_launch(C.byref(self.params), int(n), self.grid, int(share), self._grid_bar.data_ptr(), sp),
                  "persistent CSR round kernel launch")
        elif self.pipelined and n > 1:
            if share != 1:
                raise ValueError("pipelined rounds need the whole GPU: one logical worker per GPU")
            check(self.lib.afb_dense_stream_pipelined_launch(C.byref(self.params), int(n), self.grid, self._pipe_state.data_ptr(),
                                                             self._pipe_g2.data_ptr(), self._pipe_w2.data_ptr(), int(self.bf16_rows), sp),
                  "pipelined round kernel launch")
        else:
            check(self.lib.afb_dense_stream_launch(C.byref(self.params), int(n), self.grid, int(share), int(self.bf16_rows), sp),
                  "persistent round kernel launch")
        self.launches += 1
        self.rounds_launched += n

    def set(self, **kw) -> None:
        for k, v in kw.items():
            setattr(self.params, k, v)
            if self.split and k not in ("algo", "seed_mode"):
                setattr(self.params_grad, k, v)

    def attach_desc(self, desc_dev: torch.Tensor, slots: int) -> None:
        """Point the kernel at a device ring of ``RoundDesc`` (end-to-end path): slot ``round % slots`` overrides the step
        size, the batch rate and the delay coefficient of that round."""
        self.desc_dev = desc_dev
        self.set(desc=desc_dev.data_ptr(), desc_slots=int(slots))

    def detach_desc(self) -> None:
        self.desc_dev = None
        self.set(desc=None, desc_slots=0)

    def read_ctl_raw(self):
        return self.wctl.cpu().numpy().view(WORKER_CTL_DTYPE)[0]

    def read_ctl(self) -> Dict[str, float]:
        a = self.read_ctl_raw()
        return {n: (float(a[n]) if a[n].dtype.kind == "f" else int(a[n])) for n in WORKER_CTL_DTYPE.names if n != "dbg_ns"}

    def check_abort(self) -> None:
        """Raise if a bounded in-kernel spin expired in a previous launch (call after synchronising)."""
        if int(self.read_ctl_raw()["abort"]):
            raise RuntimeError(f"worker {self.worker_id}: a device-side wait expired (spin limit {self.params.spin_limit_ns / 1e9:.0f} s); "
                               "the engine state is invalid")

    def restore_progress(self, rounds_done: int) -> None:
        """Checkpoint resume: continue the worker-local round counter (sampling seed / BSP step schedule)."""
        a = self.wctl.cpu().numpy().view(WORKER_CTL_DTYPE).copy()
        a["round"] = int(rounds_done)
        a["rounds_done"] = int(rounds_done)
        self.wctl.copy_(torch.from_numpy(a.view(np.uint8)).to(self.device))

    @property
    def variant_effective(self) -> str:
        """The kernel ``launch_rounds`` actually uses."""
        if self.stream:
            return "stream-pipelined" if self.pipelined else "stream"
        if self.stream_csr:
            return "csr-stream"
        return self.variant

    @property
    def w_view(self) -> torch.Tensor:
        """The parameter vector this worker computes with (BSP: the local replica)."""
        if self.local_block is not None and self.algo in ("sgd_sync", "saga_sync"):
            return self.local_block.w
        return self.w_local
