"""Pure-Python model of one fused round — an executable specification of the kernel protocol in
``csrc/include/afb_round_impl.cuh`` + the dense/CSR gradient bodies.

It operates on the SAME byte layouts as the CUDA kernels (``ServerBlock`` regions, ``WorkerCtl``,
``ResultRec`` ring, ``WorkerSlot`` table) but with CPU tensors, so the whole host side of the fused
backend (``FusedEngine``, ``algo/fused_driver.py``, checkpointing, the AsyncContext result pump) runs —
and is tested — without a GPU.  One ``launch()`` == one kernel launch of one logical worker, executed
synchronously.  Differences from the device (by construction of a sequential model): a launch never
overlaps another, so staleness only arises from the interleaving of launches of different workers;
the bucket gate cannot block (it is evaluated once) and delays are accounted, not slept.
"""
from __future__ import annotations

import math
import time
from typing import Optional

import numpy as np
import torch

from ..core.sampler import hash_u24, mix64_int
from ._signatures import RESULT_REC_DTYPE, ROUND_DESC_DTYPE, SERVER_CTL_DTYPE, WORKER_CTL_DTYPE, WORKER_SLOT_DTYPE

ALGO_GRAD_ONLY, ALGO_ASGD, ALGO_ASAGA, ALGO_SYNC_SGD, ALGO_SYNC_SAGA = 0, 1, 2, 3, 4


def _now_ns() -> int:
    return time.perf_counter_ns()


def _soft(t: torch.Tensor, shrink: float) -> torch.Tensor:
    return torch.sign(t) * torch.clamp(t.abs() - shrink, min=0.0)


def _coeff_loss(loss_id: int, margin: torch.Tensor, y: torch.Tensor):
    """fp32 transcription of afb::loss_coeff."""
    if loss_id == 0:
        c = margin - y
        return c, c * c
    if loss_id == 3:
        c = margin - y
        return c, 0.5 * c * c
    if loss_id == 1:
        e = torch.exp(-margin.abs())
        sig = torch.where(margin >= 0, 1.0 / (1.0 + e), e / (1.0 + e))
        l1p = torch.log1p(e) + torch.clamp(-margin, min=0.0)
        return sig - y, torch.where(y > 0, l1p, l1p + margin)
    s = 2.0 * y - 1.0
    active = (1.0 - s * margin) > 0
    return torch.where(active, -s, torch.zeros_like(s)), torch.where(active, 1.0 - s * margin, torch.zeros_like(s))


class RoundSim:
    def __init__(self, rk, server_block=None):
        self.rk = rk
        self.server = server_block            # ServerBlock holding the parameter server state (rank 0's)
        self.avg_delay_accounted_ns = 0

    # structured views over the shared byte regions -------------------------------------------
    def _wc(self):
        return self.rk.wctl.numpy().view(WORKER_CTL_DTYPE)[0:1]

    def _sctl(self, block):
        return block.sctl.numpy().view(SERVER_CTL_DTYPE)[0:1]

    def launch_pipelined(self, n: int) -> None:
        """``n`` rounds of the pipelined persistent kernel (csrc/kernels/dense_stream.cu, PIPE = true): round i+1's pull is
        issued before round i's update has been applied, so each round computes on parameters that lack the worker's own
        previous update (everything else — clock, tickets, ring, status — is identical to ``n`` single launches)."""
        w_next = None
        for _ in range(n):
            before = self.server.w.clone()           # what the next round's puller reads: the server BEFORE this round's apply
            self.launch(w_pulled=w_next)
            w_next = before

    def launch(self, w_pulled: Optional[torch.Tensor] = None) -> None:
        rk, p = self.rk, self.rk.params
        wc = self._wc()
        algo = int(p.algo)
        is_async = algo in (ALGO_ASGD, ALGO_ASAGA)
        rnd = int(wc["round"][0])
        d = rk.shard.num_features
        t_entry = _now_ns()
        # round descriptor (afb::round_desc): slot rnd % S overrides the launch scalars when its seq matches
        gamma, thr24, delay_coeff = float(p.gamma), int(p.thr24), float(p.delay_coeff)
        if rk.desc_dev is not None and int(p.desc_slots) > 0:
            dsc = rk.desc_dev.numpy().view(ROUND_DESC_DTYPE).reshape(-1)[rnd % int(p.desc_slots)]
            if int(dsc["seq"]) == rnd + 1:
                gamma, thr24, delay_coeff = float(dsc["gamma"]), int(dsc["thr24"]), float(dsc["delay_coeff"])
        # ---------------------------------------------------------------- phase 0: gate + pull
        if algo != ALGO_GRAD_ONLY:
            if is_async:
                srv = self.server
                sc = self._sctl(srv)
                ck = int(sc["clock_k"][0])
                if (ck >> 32) >= int(p.max_iters) or int(sc["stop"][0]):
                    wc["proceed"] = 0
                    wc["last_accepted"] = -1
                    wc["round"] = rnd + 1
                    return
                if int(p.use_mask):
                    word = int(sc["enable_mask"][0][(int(p.worker_id) >> 6) & 3])
                    if not (word >> (int(p.worker_id) & 63)) & 1:      # masked by the host-evaluated ASYNCbarrier predicate
                        wc["proceed"] = 2
                        wc["last_accepted"] = -2
                        wc["skipped_rounds"] += 1
                        wc["round"] = rnd + 1
                        return
                sc["avail_count"] -= 1
                st = srv.status.numpy().view(WORKER_SLOT_DTYPE)
                me = st[int(p.worker_id):int(p.worker_id) + 1]
                if int(wc["last_finish_ns"][0]):
                    wc["wait_ns"] += t_entry - int(wc["last_finish_ns"][0])
                me["available"] = 0
                me["busy_since_ns"] = t_entry
                me["heartbeat_ns"] = t_entry
                me["wait_ns"] = wc["wait_ns"][0]
                rk.w_local.copy_(srv.w if w_pulled is None else w_pulled)  # K10
            else:
                ck = rnd
            wc["ts"] = ck & 0xFFFFFFFF
            wc["k_read"] = (ck >> 32) if is_async else rnd
            wc["proceed"] = 1
            wc["t_start_ns"] = t_entry
            avg = int(self._sctl(self.server if is_async else rk.local_block)["avg_delay_ns"][0])
            if int(p.delay_mode) != 0 and avg:
                mult = delay_coeff
                if int(p.delay_mode) >= 2:
                    u = float(hash_u24(0, np.asarray([int(p.worker_id)]))[0]) / float(1 << 24)
                    mult = (2.5 + 7.5 * u) if int(p.delay_mode) == 3 else (1.5 + u)
                self.avg_delay_accounted_ns += int(mult * avg)
        seed_eff = int(p.seed)
        if algo == ALGO_GRAD_ONLY and int(p.seed_mode) != 3:
            pass
        elif int(p.seed_mode) == 0:
            seed_eff += int(wc["k_read"][0]) + 1
        elif int(p.seed_mode) == 1:
            seed_eff += rnd + 1
        # ---------------------------------------------------------------- phase 1: fused gradient
        sh = rk.shard
        if sh.is_dense and sh.X.dtype == torch.bfloat16:       # bf16 row storage: widened exactly, math in fp32 like the kernel
            if getattr(self, "_f32_shard", None) is None:
                from ..data.batch import Shard
                self._f32_shard = Shard(sh.num_features, sh.y, sh.X.to(torch.float32), gidx=sh.gidx, partition=sh.partition)
            sh = self._f32_shard
        gidx = sh.gidx.numpy() if (sh.gidx is not None and int(p.gidx or 0)) else (np.arange(sh.num_rows, dtype=np.int64) + int(p.row_base))
        rows = torch.from_numpy(np.nonzero(hash_u24(seed_eff, gidx) < thr24)[0])
        w = (rk.w_view if algo in (ALGO_SYNC_SGD, ALGO_SYNC_SAGA) else rk.w_local)[:d].to(torch.float32)
        g = torch.zeros(rk.ld, dtype=torch.float32)
        loss_sum, batch = 0.0, int(rows.numel())
        saga = rk.hist is not None and int(p.hist or 0) != 0
        if batch:
            sub = sh.select(rows)
            margin = sub.matvec(w).to(torch.float32)
            coeff, loss = _coeff_loss(int(p.loss_id), margin, sub.y.to(torch.float32))
            eff = coeff
            if saga:
                h = rk.hist[rows]
                eff = coeff - h
                if rk.rollback:
                    rk.hist_prev[rows] = h
                    rk.hist_tag[rows] = rnd
                rk.hist[rows] = coeff                                                      # K7
            g[:d] = sub.rmatvec(eff).to(torch.float32)
            loss_sum = float(loss.sum())
        t_compute = _now_ns()
        # ---------------------------------------------------------------- phase 3
        if algo == ALGO_GRAD_ONLY:
            rk.g_local.add_(g)
            wc["last_loss"] = loss_sum
            wc["last_batch"] = batch
            wc["rounds_done"] += 1
            wc["round"] = rnd + 1
            return
        if is_async:
            srv = self.server
            sc = self._sctl(srv)
            old = int(sc["clock_k"][0])
            ts = int(wc["ts"][0])
            clock_old, k_old = old & 0xFFFFFFFF, old >> 32
            staleness = clock_old - ts
            if int(p.taw) >= (1 << 31):
                sc["clock_k"] = old + (1 << 32) + 1
                accept = k_old < int(p.max_iters)
            else:
                cmp_v = (k_old - ts) if algo == ALGO_ASAGA else staleness
                accept = cmp_v <= int(p.taw)
                new = old + 1
                if accept:
                    if k_old >= int(p.max_iters):
                        accept = False
                    new += (1 << 32)
                sc["clock_k"] = new
            step = gamma / math.sqrt(k_old // int(p.num_workers) + 1) if algo == ALGO_ASGD else gamma
            if accept:
                # hogwild_apply: delta = -step (g/parRecs [+ abar] [+ lambda w_pulled]); L1: prox against the CURRENT server value
                delta = g * (-step * float(p.inv_par_recs))
                if algo == ALGO_ASAGA:
                    delta = delta - step * srv.abar                                            # abar BEFORE its update (K6)
                    srv.abar.add_(g, alpha=float(p.inv_n))
                if int(p.reg_type) == 1:
                    delta = delta - step * float(p.reg_param) * rk.w_local
                elif int(p.reg_type) == 2:
                    delta = _soft(srv.w + delta, step * float(p.reg_param)) - srv.w
                srv.w.add_(delta)                                                              # K5 (Hogwild red.add)
            elif saga and rk.rollback:
                tagged = rk.hist_tag == rnd
                rk.hist[tagged] = rk.hist_prev[tagged]                                         # saga_rollback_kernel
            now = _now_ns()
            ring = srv.ring.numpy().view(RESULT_REC_DTYPE)
            rec = ring[clock_old % int(p.ring_size):clock_old % int(p.ring_size) + 1]
            k_after = k_old + (1 if accept else 0)
            rec["worker"] = int(p.worker_id); rec["accepted"] = int(accept); rec["ts"] = ts; rec["staleness"] = staleness
            rec["k_after"] = k_after; rec["batch"] = batch; rec["loss_sum"] = loss_sum; rec["step"] = step if accept else 0.0
            rec["t_start_ns"] = int(wc["t_start_ns"][0]); rec["t_end_ns"] = now; rec["seq"] = clock_old + 1
            st = srv.status.numpy().view(WORKER_SLOT_DTYPE)
            me = st[int(p.worker_id):int(p.worker_id) + 1]
            ntasks = int(wc["rounds_done"][0]) + 1
            task_ns = now - int(wc["t_start_ns"][0])
            wc["avg_task_ms"] += (task_ns * 1e-6 - float(wc["avg_task_ms"][0])) / ntasks
            wc["last_finish_ns"] = now
            me["staleness"] = staleness; me["num_tasks"] = ntasks; me["last_task_ms"] = task_ns * 1e-6
            me["avg_task_ms"] = wc["avg_task_ms"][0]; me["heartbeat_ns"] = now; me["last_finish_ns"] = now; me["available"] = 1
            if accept and int(p.warmup_updates) > 0 and k_after <= int(p.warmup_updates):
                sc["cul_time_ns"] += task_ns
                sc["cul_count"] += 1
                if k_after == int(p.warmup_updates) and int(sc["cul_count"][0]):
                    sc["avg_delay_ns"] = int(sc["cul_time_ns"][0]) // int(sc["cul_count"][0])
            sc["avail_count"] += 1
            wc["last_accepted"] = int(accept); wc["last_batch"] = batch; wc["last_loss"] = loss_sum
            wc["last_step"] = step if accept else 0.0; wc["last_staleness"] = staleness; wc["last_k_after"] = k_after
            wc["last_task_ns"] = task_ns; wc["rounds_done"] += 1
            dbg = wc["dbg_ns"][0]
            dbg[:] = [t_entry, t_entry, t_entry, t_compute, t_compute, now, now, now]
            wc["round"] = rnd + 1
            return
        # ---- BSP (single process: P == 1, the all-gather degenerates to the worker's own slot) ----
        blk = rk.local_block
        step = gamma / math.sqrt(rnd + 1) if algo == ALGO_SYNC_SGD else gamma
        a = step * float(p.inv_par_recs)
        if int(p.reg_type) == 1:
            blk.w.mul_(1.0 - step * float(p.reg_param))                                        # SquaredL2Updater (Updater.scala:151-158)
        if algo == ALGO_SYNC_SGD:
            blk.w.add_(g, alpha=-a)
        else:
            ab = blk.abar.clone()
            blk.w.add_(g, alpha=-a)
            blk.w.add_(ab, alpha=-step)
            blk.abar.add_(g, alpha=float(p.inv_n))
        if int(p.reg_type) == 2:
            blk.w.copy_(_soft(blk.w, step * float(p.reg_param)))                               # L1Updater (Updater.scala:114-129)
        sc = self._sctl(blk)
        if rnd < int(p.warmup_updates):
            sc["cul_time_ns"] += _now_ns() - t_entry
            sc["cul_count"] += 1
            if rnd + 1 == int(p.warmup_updates):
                sc["avg_delay_ns"] = int(sc["cul_time_ns"][0]) // max(int(sc["cul_count"][0]), 1)
        wc["last_accepted"] = 1; wc["last_batch"] = batch; wc["last_loss"] = loss_sum; wc["last_step"] = step
        wc["last_staleness"] = 0; wc["last_k_after"] = rnd + 1; wc["last_task_ns"] = _now_ns() - t_entry
        wc["rounds_done"] += 1
        wc["round"] = rnd + 1
