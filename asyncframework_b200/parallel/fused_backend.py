"""The fused backend: every worker round is ONE hand-written kernel that pulls ``w`` over NVLink,
computes the sampled gradient in a single pass and pushes/applies it into the server's HBM.

``FusedEngine`` is the public API of this backend (what bench.py, the drivers and users call):

    engine = FusedEngine(cfg, local_shards, total_rows)      # one process per GPU (torchrun) or 1 GPU
    engine.run_rounds(n)      # enqueue n rounds per local worker; no host sync on the hot path
    engine.step()             # end-to-end step: H2D round descriptor -> round -> D2H result
    engine.weights()          # current model (rank 0's server block; BSP: the local replica)

Mapping to the reference: a logical worker = an RDD partition (numPart), rank 0's server block =
the driver's ``w`` / ``alphaBar`` / ``ScalarMap`` owner, the device-side clock/availability/result
ring = ASYNCcontext (core/rdd/ASYNCcontext.scala) + workerState, and the gate/staleness logic =
ASYNCbarrier + the updater thread's ``staleness <= taw`` filter (SparkASGDThread.scala:150-223).
"""
from __future__ import annotations

import math
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from ..config import AsyncConfig
from ..core.partial_result import PartialResult
from ..data.batch import Shard
from ..ops._signatures import RESULT_REC_DTYPE, ROUND_DESC_DTYPE, WORKER_CTL_DTYPE, WORKER_SLOT_DTYPE
from ..core.sampler import threshold24
from ..ops.round import NEVER_STALE, RoundKernel, ServerBlock, ServerLayout, ServerPointers
from ..optim.gradient import LOSS_IDS
from . import _devapi as _dev
from .comm import SymmetricBuffer, allocate_symmetric

_CUDART = None


def _cudart():
    global _CUDART
    if _CUDART is None:
        import ctypes as C
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                _CUDART = C.CDLL(name)
                break
            except OSError:
                continue
        if _CUDART is None:
            raise RuntimeError("libcudart not found")
    return _CUDART


ALGO_NAMES = {"asgd": "asgd", "asaga": "asaga", "sgd_sync": "sgd_sync", "saga_sync": "saga_sync"}


def straggler_sets(num_partitions: int):
    """Production-cluster straggler model (SparkASGDThread.scala:121-138): 25 % of the partitions
    (indices 4c) are stragglers; the first 20 % of those are long-tail, the rest 'normal'."""
    length = int(round(0.25 * num_partitions))
    n_normal = int(round(0.8 * length))
    n_long = length - n_normal
    long_tail = [c * 4 for c in range(length) if c < n_long]
    normal = [c * 4 for c in range(length) if c >= n_long]
    return normal, long_tail


class FusedEngine:
    def __init__(self, cfg: AsyncConfig, local_shards: Sequence[Shard], total_rows: int,
                 local_worker_ids: Optional[Sequence[int]] = None, device=None, group=None,
                 ring_size: int = 4096, symm_provider: str = "auto", w0: Optional[torch.Tensor] = None,
                 worker_mask: Optional[bool] = None, pipeline: Optional[bool] = None):
        self.cfg = cfg
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device) if device is not None else local_shards[0].device
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.P = int(cfg.num_partitions)
        if local_worker_ids is None:
            local_worker_ids = [p for p in range(self.P) if p % self.world == self.rank]
        self.local_worker_ids = list(local_worker_ids)
        if len(self.local_worker_ids) != len(local_shards):
            raise ValueError("one shard per local worker expected")
        self.total_rows = int(total_rows)
        first = local_shards[0]
        self.d = first.num_features
        ld = first.ld if first.is_dense else (self.d + 3) // 4 * 4
        self.layout = ServerLayout(ld, self.P, ring_size, history_slots=int(cfg.history_slots))
        # ---- symmetric server block (rank 0's copy is the parameter server) ----------------
        self.symm: SymmetricBuffer = allocate_symmetric(self.layout.nbytes, self.device, group, symm_provider)
        self.block = ServerBlock(self.layout, self.symm.local)
        self.block.reset(w0)
        self.symm.barrier()
        algo = cfg.algorithm
        if algo not in ALGO_NAMES:
            raise ValueError(f"fused backend does not implement {algo!r}")
        self.algo = algo
        self.is_async = algo in ("asgd", "asaga")
        server = ServerPointers(self.symm.ptrs[0], self.layout)
        # BSP: worker q's staging lives on the rank hosting q
        peer_blocks = None
        if not self.is_async:
            if len(self.local_worker_ids) != 1 and self.world > 1:
                raise ValueError("BSP fused mode runs one worker per rank")
            peer_blocks = [ServerPointers(self.symm.ptrs[q % self.world], self.layout) for q in range(self.P)]
            if self.world == 1 and self.P > 1:
                raise ValueError("BSP fused mode on a single process supports P == 1 (use the in-process scheduler for P > 1)")
        gate = 0
        if self.is_async and cfg.barrier in ("bucket", "bsp"):
            beta = 1.0 if cfg.barrier == "bsp" else cfg.bucket_ratio
            gate = int(math.floor(self.P * beta))
            if gate <= 1:
                gate = 0          # >= 1 idle worker is always true for the asker itself
        self.gate_threshold = gate
        self.beta = (1.0 if cfg.barrier == "bsp" else cfg.bucket_ratio) if gate else 0.0
        # device-side ASYNCbarrier beyond the bucket rule: SSP on the live status table, host-evaluated predicate mask
        ssp_bound = int(cfg.ssp_bound) if (self.is_async and cfg.barrier == "ssp") else -1
        if worker_mask is None:
            worker_mask = bool(cfg.extra.get("worker_mask", False))
        self.worker_mask = bool(worker_mask and self.is_async)
        # K8 on the fused path: the regulariser of the reference's updaters (Updater.scala:114-158) folded into the apply
        reg_type = {"l1": "l1", "squared_l2": "l2", "l2": "l2"}.get(str(cfg.updater).lower()) if cfg.reg_param > 0 else None
        normal, long_tail = straggler_sets(self.P) if cfg.production_cluster else ([], [])
        warm = cfg.warmup_updates_per_worker * (self.P if self.is_async else 1)
        n_local = len(self.local_worker_ids)
        sms = _dev.device_sms(self.device)
        self.kernels: List[RoundKernel] = []
        self.streams: List[torch.cuda.Stream] = []
        self.wctl_all = torch.zeros(n_local, WORKER_CTL_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        for li, (shard, wid) in enumerate(zip(local_shards, self.local_worker_ids)):
            if cfg.production_cluster:
                dmode = 3 if wid in long_tail else (2 if wid in normal else 0)
            else:
                dmode = 1 if (cfg.coeff > 0 and wid == 0) else 0
            grid = None
            if n_local > 1:
                # co-resident logical workers split the resident-CTA capacity (2 CTAs/SM): with two
                # workers per GPU each gets one CTA per SM, so one worker's NVLink latency phases
                # overlap the other's HBM streaming
                grid = max(1, (2 * sms) // n_local)
            rk = RoundKernel(shard, server, algo=algo, loss_id=LOSS_IDS[cfg.loss], num_workers=self.P,
                             worker_id=wid, total_rows=self.total_rows, batch_rate=cfg.batch_rate,
                             gamma=cfg.step_size, seed=cfg.seed,
                             taw=cfg.taw if cfg.taw < (1 << 31) - 1 else NEVER_STALE,
                             max_iters=cfg.num_iterations if self.is_async else (1 << 40),
                             gate_threshold=gate, delay_mode=dmode, delay_coeff=max(cfg.coeff, 0.0),
                             warmup_updates=warm, local_block=self.block, peer_blocks=peer_blocks, grid=grid,
                             wctl=self.wctl_all[li], sim_server_block=self.block if self.device.type != "cuda" else None,
                             reg_param=cfg.reg_param, reg_type=reg_type, ssp_bound=ssp_bound, use_mask=self.worker_mask,
                             pipeline=(pipeline if n_local == 1 else False),      # pipelined rounds own every SM of their GPU
                             multicast_base=int(getattr(self.symm, "multicast_ptr", 0) or 0) if self.world > 1 else 0)
            self.kernels.append(rk)
            self.streams.append(_dev.new_stream(self.device) if n_local > 1 else _dev.current_stream(self.device))
        # pinned staging for the end-to-end path (double-buffered: step i+1 is submitted before the
        # result of step i is read, so the per-step host work never idles the GPU)
        # end-to-end path: a ring of per-round descriptors {step size, Bernoulli threshold, delay intensity, seq}; step()
        # writes slot (round % S) in pinned memory and copies it H2D, the round kernel READS it (afb::round_desc)
        self.desc_slots = 64
        self._desc_host = _dev.pin(torch.zeros(self.desc_slots, ROUND_DESC_DTYPE.itemsize, dtype=torch.uint8), self.device)
        self._desc_np = self._desc_host.numpy().view(ROUND_DESC_DTYPE).reshape(-1)
        self._desc_dev = torch.zeros(self.desc_slots, ROUND_DESC_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        for rk in self.kernels:
            rk.attach_desc(self._desc_dev, self.desc_slots)
        self.dead_workers: List[int] = []
        self._ctl_host = _dev.pin(torch.zeros(2, n_local, WORKER_CTL_DTYPE.itemsize, dtype=torch.uint8), self.device)
        self._ctl_np = self._ctl_host.numpy()
        self._step_events = [_dev.new_event(self.device), _dev.new_event(self.device)]
        self._step_pending = [False, False]
        # overlapped end-to-end plumbing (csrc/kernels/e2e_step.cu): the per-step H2D / D2H copies ride on side streams, so
        # the launching stream carries nothing but rounds.  One local worker on a GPU only; AFB_E2E_OVERLAP=0 keeps both
        # copies on the launching stream (the previous behaviour, and the one multi-worker engines use).
        import os as _os
        self._e2e = None
        self._e2e_lib = None
        if (self.device.type == "cuda" and n_local == 1 and not self.kernels[0].sim and _os.environ.get("AFB_E2E_OVERLAP", "1") != "0"
                and hasattr(self.kernels[0].lib, "afb_e2e_create")):
            self._e2e_lib = self.kernels[0].lib
            self._e2e = self._e2e_lib.afb_e2e_create() or None
        if self._e2e is None and self._e2e_lib is not None:
            import warnings
            warnings.warn("afb_e2e_create failed (side streams / events): FusedEngine.step() keeps its copies on the launching stream")
            self._e2e_lib = None
        if self._e2e is not None:
            self._snap = torch.zeros(2, n_local, WORKER_CTL_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
            self._e2e_ptrs = (self._desc_dev.data_ptr(), self._desc_host.data_ptr(), self.wctl_all.data_ptr(),
                              [self._snap[0].data_ptr(), self._snap[1].data_ptr()],
                              [self._ctl_host[0].data_ptr(), self._ctl_host[1].data_ptr()])
        self._step_idx = 0
        self._ring_seen = 0
        self.rounds_enqueued = 0
        self._graph = None
        self._graph_rounds = 0
        self._graph_launches = 0
        self._graph_kernel_launches = 0
        self.symm.barrier()

    # ------------------------------------------------------------------ hot path
    def run_rounds(self, n: int) -> None:
        """Enqueue ``n`` rounds on every local worker.  No host synchronisation."""
        if all(rk.stream or rk.stream_csr or rk.sim for rk in self.kernels):
            # persistent worker loops: ONE resident kernel per local worker runs its n rounds
            for rk, st in zip(self.kernels, self.streams):
                rk.launch_rounds(n, st, share=len(self.kernels))
        else:
            for _ in range(n):
                for rk, st in zip(self.kernels, self.streams):
                    rk.launch(st)
        self.rounds_enqueued += n

    @property
    def launches(self) -> int:
        return sum(rk.launches for rk in self.kernels) + self._graph_launches

    # ------------------------------------------------------------------ CUDA graphs
    def build_graph(self, rounds: int) -> None:
        """Capture ``rounds`` rounds of every local worker into one CUDA graph.  Everything that
        changes from round to round (round counter, seeds, step size via k, stop condition) lives in
        device memory, so a replay is exactly ``run_rounds(rounds)`` with one host call — this removes
        the launch-bound regime for small rounds (epsilon / rcv1 shapes)."""
        if self.device.type != "cuda":
            raise RuntimeError("CUDA graphs need a CUDA device")
        self.synchronize()
        cap = torch.cuda.Stream(device=self.device)
        g = torch.cuda.CUDAGraph()
        before = [rk.launches for rk in self.kernels]
        cap.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.graph(g, stream=cap):
            fork = torch.cuda.Event()
            fork.record(cap)
            side = []
            for rk, st in zip(self.kernels, self._graph_streams(cap)):
                if st != cap:
                    st.wait_event(fork)
                rk.launch_rounds(rounds, st, share=len(self.kernels))
                if st != cap:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    side.append(ev)
            for ev in side:
                cap.wait_event(ev)
        self._graph = g
        self._graph_rounds = rounds
        self._graph_kernel_launches = sum(rk.launches - b for rk, b in zip(self.kernels, before))
        for rk, b in zip(self.kernels, before):          # capture does not execute: undo the launch counters
            rk.launches = b

    def _graph_streams(self, cap):
        if len(self.kernels) == 1:
            return [cap]
        if not hasattr(self, "_gstreams"):
            self._gstreams = [cap] + [torch.cuda.Stream(device=self.device) for _ in self.kernels[1:]]
        self._gstreams[0] = cap
        return self._gstreams

    def run_graph(self, replays: int = 1) -> None:
        """Replay the captured graph on the current stream (``replays * rounds`` rounds per worker)."""
        if self._graph is None:
            raise RuntimeError("build_graph() first")
        for _ in range(replays):
            self._graph.replay()
        self.rounds_enqueued += replays * self._graph_rounds
        self._graph_launches += replays * self._graph_kernel_launches

    def synchronize(self) -> None:
        for st in self.streams:
            st.synchronize()
        if self._e2e is not None:
            from ..ops import check
            check(self._e2e_lib.afb_e2e_sync(self._e2e), "end-to-end side streams")

    def _wait_step(self, slot: int) -> None:
        """Block until the result record of the step that used ``slot`` is in pinned host memory."""
        if self._e2e is not None:
            from ..ops import check
            check(self._e2e_lib.afb_e2e_wait(self._e2e, slot), "end-to-end result wait")
        else:
            self._step_events[slot].synchronize()

    def __del__(self):
        try:
            if getattr(self, "_e2e", None):
                self._e2e_lib.afb_e2e_destroy(self._e2e)
                self._e2e = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def stagger(self, period_us: float) -> None:
        """Offset the local workers' round cadence by ``period/num_local`` each (one-time device-side
        sleep on their streams) so their NVLink round trips and HBM streaming phases interleave instead
        of staying in lock-step."""
        from ..ops import check, kernels
        n = len(self.kernels)
        for j, st in enumerate(self.streams):
            if j == 0 or n < 2:
                continue
            check(kernels().afb_sleep(int(period_us * 1e3 * j / n), int(st.cuda_stream)), "stagger")

    def fork(self, event) -> None:
        """Make every worker stream wait for ``event`` (recorded on the launching stream)."""
        cur = _dev.current_stream(self.device)
        for st in self.streams:
            if st != cur:
                st.wait_event(event)

    def join(self) -> None:
        """Make the launching stream wait for everything enqueued on the worker streams."""
        cur = _dev.current_stream(self.device)
        for st in self.streams:
            if st != cur:
                ev = _dev.new_event(self.device)
                ev.record(st)
                cur.wait_event(ev)

    # ------------------------------------------------------------------ end-to-end step
    def _parse_result(self, slot: int) -> Dict[str, float]:
        ctl = self._ctl_np[slot].view(WORKER_CTL_DTYPE).reshape(-1)
        return {"loss_sum": float(ctl["last_loss"].sum()), "batch": int(ctl["last_batch"].sum()),
                "accepted": int((ctl["last_accepted"] == 1).sum()), "k": int(ctl["last_k_after"].max()),
                "staleness": int(ctl["last_staleness"].max())}

    def step(self, step_size: Optional[float] = None, lag: int = 0, batch_rate: Optional[float] = None,
             delay_coeff: Optional[float] = None) -> Optional[Dict[str, float]]:
        """One user-visible step: copy this round's descriptor (step size, batch rate, delay
        coefficient, round id) from pinned host memory to the device, run one fused round per local
        worker, copy every worker's result record back to pinned memory.

        ``lag=0`` waits for this step's record and returns it.  ``lag=1`` returns the record of the
        PREVIOUS step (None on the first call): the copies still happen every step, but the host reads
        results one step behind so the device never waits for Python (call ``flush_steps()`` at the end)."""
        slot = self._step_idx & 1
        if self._step_pending[slot]:                 # slot reuse: its previous occupant must have been read
            self._wait_step(slot)
        rnd = self.kernels[0].rounds_launched        # == the device round counter of every local worker (lock-step launches)
        ds = rnd % self.desc_slots
        d = self._desc_np[ds:ds + 1]
        d["gamma"] = self.kernels[0].params.gamma if step_size is None else float(step_size)
        d["thr24"] = self.kernels[0].params.thr24 if batch_rate is None else threshold24(float(batch_rate))
        d["delay_coeff"] = max(self.cfg.coeff, 0.0) if delay_coeff is None else float(delay_coeff)
        d["seq"] = rnd + 1
        cur = _dev.current_stream(self.device)
        if self._e2e is not None and self._step_idx == 0:
            # first step: if the side-stream copy cannot even be enqueued (nothing has been launched yet), fall back to the
            # launching-stream plumbing for good instead of failing the run
            rc = self._e2e_lib.afb_e2e_pre(self._e2e, slot, self._e2e_ptrs[0] + ds * ROUND_DESC_DTYPE.itemsize,
                                           self._e2e_ptrs[1] + ds * ROUND_DESC_DTYPE.itemsize, ROUND_DESC_DTYPE.itemsize, int(cur.cuda_stream))
            if rc != 0:
                import warnings
                warnings.warn(f"side-stream descriptor copy failed (cuda error {rc}): step() keeps its copies on the launching stream")
                self._e2e_lib.afb_e2e_destroy(self._e2e)
                self._e2e, self._e2e_lib = None, None
            first_pre_done = self._e2e is not None
        else:
            first_pre_done = False
        if self._e2e is not None:
            # H2D on the descriptor stream (the round waits for it through an event), round on the launching stream, result
            # snapshot behind it, D2H on the result stream: the launching stream never waits for a copy engine
            from ..ops import check
            lib, ctx, sp = self._e2e_lib, self._e2e, int(cur.cuda_stream)
            desc_dev, desc_host, wctl, snaps, hosts = self._e2e_ptrs
            nb = ROUND_DESC_DTYPE.itemsize
            if not first_pre_done:
                check(lib.afb_e2e_pre(ctx, slot, desc_dev + ds * nb, desc_host + ds * nb, nb, sp), "end-to-end descriptor copy")
            self.kernels[0].launch(self.streams[0])
            check(lib.afb_e2e_post(ctx, slot, wctl, snaps[slot], hosts[slot], self._snap[slot].numel(), sp), "end-to-end result copy")
        else:
            self._desc_dev[ds].copy_(self._desc_host[ds], non_blocking=True)            # H2D (pinned): this round's descriptor
            multi = len(self.kernels) > 1
            if multi:
                fork = _dev.new_event(self.device)
                fork.record(cur)
            for rk, st in zip(self.kernels, self.streams):
                if multi and st != cur:
                    st.wait_event(fork)
                rk.launch(st)
            if multi:
                self.join()
            self._ctl_host[slot].copy_(self.wctl_all, non_blocking=True)                # D2H (pinned), all local workers
            self._step_events[slot].record(cur)
        self.rounds_enqueued += 1
        self._step_pending[slot] = True
        self._step_idx += 1
        if lag == 0:
            self._wait_step(slot)
            self._step_pending[slot] = False
            return self._parse_result(slot)
        prev = slot ^ 1
        if self._step_pending[prev]:
            self._wait_step(prev)
            self._step_pending[prev] = False
            return self._parse_result(prev)
        return None

    def flush_steps(self) -> Optional[Dict[str, float]]:
        """Wait for (and return) the newest outstanding lagged step result."""
        last = None
        for slot in ((self._step_idx & 1), (self._step_idx & 1) ^ 1):
            if self._step_pending[slot]:
                self._wait_step(slot)
                self._step_pending[slot] = False
                last = self._parse_result(slot)
        return last

    @property
    def h2d_bytes_per_step(self) -> int:
        return int(ROUND_DESC_DTYPE.itemsize)

    @property
    def d2h_bytes_per_step(self) -> int:
        return int(self._ctl_host[0].numel())

    # ------------------------------------------------------------------ side stream (reads / pokes while rounds are running)
    def side_stream_ctx(self):
        """Context manager: torch ops inside run on a dedicated NON-BLOCKING stream, so the watchdog / result pump can read
        and poke the server block while a persistent round kernel occupies the launching stream."""
        import contextlib
        if self.device.type != "cuda":
            return contextlib.nullcontext()
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = _dev.new_stream(self.device)
        return _dev.stream_ctx(self._side_stream, self.device)

    # ------------------------------------------------------------------ health, elasticity, resume
    def check_health(self) -> None:
        """Raise if any local worker's kernel gave up on a bounded device-side wait (call after synchronising)."""
        for rk in self.kernels:
            if not rk.sim:
                rk.check_abort()

    def set_worker_mask(self, enabled: Sequence[bool]) -> None:
        """ASYNCbarrier(f, STAT) on the fused path (core/rdd/RDD.scala:1050-1064): the host evaluates the predicate over the
        mirrored status table (``worker_states()``) and writes the selection; a worker whose bit is clear sits its rounds out
        (an empty task in the reference).  Rank 0 owns the mask (workers read it from the server block)."""
        if not self.worker_mask:
            raise RuntimeError("construct the engine with worker_mask=True (cfg.extra['worker_mask']) to use predicate gating")
        if len(enabled) != self.P:
            raise ValueError("one flag per logical worker expected")
        if self.rank == 0:
            self.block.set_enable_mask([bool(x) for x in enabled])

    def async_barrier(self, predicate) -> List[int]:
        """Evaluate ``predicate(WorkerState) -> bool`` on the live device status table and select the workers for the next
        rounds; returns the selected worker ids (rank 0; other ranks return [])."""
        if self.rank != 0:
            return []
        states = self.worker_states()
        sel = [bool(predicate(states[q])) and q not in self.dead_workers for q in range(self.P)]
        self.set_worker_mask(sel)
        return [q for q, on in enumerate(sel) if on]

    def worker_states(self):
        """The device status table as ``WorkerState`` objects (staleness, average task time, availability, task count, with
        the table-wide queries ``available_workers()`` / ``max_staleness()`` the reference exposes on every record)."""
        from ..core.worker_state import StatTable
        st = self.worker_status()
        table = StatTable(self.P)
        for q in range(self.P):
            table.set_state(q, int(st[q]["staleness"]), float(st[q]["avg_task_ms"]), bool(st[q]["available"]) and bool(st[q]["alive"]),
                            int(st[q]["num_tasks"]))
        return [table[q] for q in range(self.P)]

    def declare_dead(self, worker: int) -> None:
        """Continue with P-1 (the Spark side is HeartbeatReceiver.scala:193-220 expiring an executor): mark the
        worker dead in the server's status table, take it out of the idle count if it died idle, lower the bucket threshold
        to floor(P_alive * beta) and clear its enable bit.  Asynchronous modes only; rank 0 (the server) applies it."""
        if worker in self.dead_workers:
            return
        self.dead_workers.append(int(worker))
        if self.rank != 0 or not self.is_async:
            return
        st = self.block.read_status()
        was_idle = bool(st[worker]["available"])
        st[worker]["alive"] = 0
        st[worker]["available"] = 0
        off = self.layout.off_status + worker * WORKER_SLOT_DTYPE.itemsize
        self.block.buf[off:off + 8].copy_(torch.from_numpy(st[worker:worker + 1].view(np.uint8)[:8].copy()).to(self.device))
        if was_idle:
            self.block.add_avail(-1)
        alive = self.P - len(self.dead_workers)
        self.block.set_gate_dyn(max(1, int(math.floor(alive * self.beta))) if self.gate_threshold else 0, alive)
        if self.worker_mask:
            self.block.set_enable_mask([q not in self.dead_workers for q in range(self.P)])

    def request_stop_local(self) -> None:
        """Stop THIS rank's rounds (BSP: a peer died — every surviving rank's watchdog calls this; the in-kernel flag
        waits observe the local stop word and abort cleanly)."""
        self.block.request_stop()

    def restore_progress(self, k: int, clock: Optional[int] = None, rounds: Optional[Sequence[int]] = None) -> None:
        """Checkpoint resume: restore the accepted-update counter / logical clock of the server and the per-worker round
        counters, so the step-size schedule (gamma/sqrt(floor(k/P)+1), rsqrt(round+1)) and the seed+round+1 sampling
        sequence continue where the checkpoint left off."""
        if self.rank == 0 and self.is_async:
            self.block.set_progress(int(k), int(k if clock is None else clock))
        for i, rk in enumerate(self.kernels):
            r = int(rounds[i]) if rounds is not None else (int(k) // max(self.P, 1) if self.is_async else int(k))
            if rk.sim:
                a = rk.wctl.numpy().view(WORKER_CTL_DTYPE)
                a["round"] = r
                a["rounds_done"] = r
            else:
                rk.restore_progress(r)
            rk.rounds_launched = r
        self.rounds_enqueued = max((rk.rounds_launched for rk in self.kernels), default=0)

    # ------------------------------------------------------------------ state
    def weights(self) -> torch.Tensor:
        """Current model: rank 0's server ``w`` (peers read it through the symmetric mapping)."""
        self.synchronize()
        if self.is_async and self.rank != 0 and self.world > 1:
            host = self._peer_read(self.symm.ptrs[0] + self.layout.off_w, self.layout.ld * 4).view(np.float32)
            return torch.from_numpy(host.copy()).to(self.device)[: self.d]
        return self.block.w[: self.d].clone()

    def alpha_bar(self) -> torch.Tensor:
        self.synchronize()
        return self.block.abar[: self.d].clone()

    def state(self) -> Dict[str, int]:
        self.synchronize()
        return self.block.read_sctl()

    def _peer_read(self, addr: int, nbytes: int) -> np.ndarray:
        """Copy ``nbytes`` (multiple of 4) from a peer-mapped device address to the host: an in-kernel
        NVLink load into a local bounce buffer (the path the round kernels use), then a local D2H copy."""
        from ..ops import check, kernels
        n = (nbytes + 3) // 4 * 4
        if getattr(self, "_bounce", None) is None or self._bounce.numel() < n:
            self._bounce = torch.empty(max(n, 4096), dtype=torch.uint8, device=self.device)
        st = torch.cuda.current_stream(self.device)
        check(kernels().afb_peer_copy(self._bounce.data_ptr(), addr, n, int(st.cuda_stream)), "peer_copy")
        return self._bounce[:n].cpu().numpy()[:nbytes].copy()

    def server_state(self) -> Dict[str, int]:
        """Clock / accepted-update counter of the parameter server (rank 0), readable from any rank."""
        self.synchronize()
        if self.rank == 0 or self.world == 1 or not self.is_async:
            return self.block.read_sctl()
        from ..ops._signatures import SERVER_CTL_DTYPE
        a = self._peer_read(self.symm.ptrs[0] + self.layout.off_sctl, SERVER_CTL_DTYPE.itemsize).view(SERVER_CTL_DTYPE)[0]
        ck = int(a["clock_k"])
        return {"clock": ck & 0xFFFFFFFF, "k": ck >> 32, "avail_count": int(a["avail_count"]), "stop": int(a["stop"]),
                "avg_delay_ns": int(a["avg_delay_ns"]), "cul_count": int(a["cul_count"])}

    def worker_status(self) -> np.ndarray:
        return self.block.read_status()

    def drain_results(self) -> List[PartialResult]:
        """New entries of the device result ring as ``PartialResult`` envelopes (rank 0)."""
        ring = self.block.read_ring()
        out: List[PartialResult] = []
        n = self.layout.ring_size
        while True:
            rec = ring[self._ring_seen % n]
            if int(rec["seq"]) != self._ring_seen + 1:
                break
            out.append(PartialResult(
                data={"accepted": bool(rec["accepted"]), "loss_sum": float(rec["loss_sum"]), "step": float(rec["step"]),
                      "k_after": int(rec["k_after"])},
                staleness=int(rec["staleness"]), timestamp=int(rec["ts"]), batch_size=int(rec["batch"]),
                worker_id=int(rec["worker"]), submit_ms=float(rec["t_start_ns"]) * 1e-6,
                finish_ms=float(rec["t_end_ns"]) * 1e-6))
            self._ring_seen += 1
        return out

    # ------------------------------------------------------------------ versioned broadcast (A8/A9)
    def publish_version(self) -> int:
        """``ASYNCbroadcast``: rank 0 publishes the current ``w`` as a new version into slot
        ``version % K`` of EVERY rank's history ring — one NVLS multicast store stream when the fabric
        supports it (``multimem.st``), peer stores otherwise — and bumps every rank's version counter.
        Returns the new version id.  Collective: call on all ranks."""
        import ctypes as C
        from ..ops import check, kernels
        K = self.layout.history_slots
        if K <= 0:
            raise ValueError("history_slots == 0: versioned broadcast disabled")
        self.synchronize()
        self.barrier()
        version = int(self.block.version_words[0].item())
        slot = version % K
        if self.device.type != "cuda":
            # simulator: every rank == this process; the "multicast" is a tensor copy + version bump
            self.block.hist_w[slot].copy_(self.block.w)
            self.block.version_words[0] += 1
            return version
        if self.rank == 0:
            lib = kernels()
            st = torch.cuda.current_stream(self.device).cuda_stream
            off = self.layout.off_hist_w + slot * self.layout.ld * 4
            done = self.block.version_words[2:3]
            if self.symm.multicast_ptr:
                check(lib.afb_broadcast_mc(self.block.w.data_ptr(), self.symm.multicast_ptr + off, self.layout.ld,
                                           self.symm.multicast_ptr + self.layout.off_version, done.data_ptr(), st), "broadcast_mc")
            else:
                dst = (C.c_void_p * self.world)(*[p + off for p in self.symm.ptrs])
                flg = (C.c_void_p * self.world)(*[p + self.layout.off_version for p in self.symm.ptrs])
                check(lib.afb_broadcast_p2p(self.block.w.data_ptr(), dst, flg, self.world, self.layout.ld, done.data_ptr(), st),
                      "broadcast_p2p")
            torch.cuda.synchronize(self.device)
        self.barrier()
        return version

    def version_value(self, version: int) -> torch.Tensor:
        """``ASYNCbroadcast.value(index)``: a previously published version, read from the LOCAL ring."""
        K = self.layout.history_slots
        cur = int(self.block.version_words[0].item())
        if not (cur - K <= version < cur):
            raise KeyError(f"version {version} not in the ring (current {cur}, slots {K})")
        return self.block.hist_w[version % K, : self.d].clone()

    @property
    def version(self) -> int:
        return int(self.block.version_words[0].item())

    # ------------------------------------------------------------------ AsyncContext bridge
    def start_result_pump(self, ctx, period_s: float = 0.002):
        """Feed a host ``AsyncContext`` from the device: a daemon thread on rank 0 drains the result
        ring into ``ctx`` (so ``ctx.collect_all()`` / ``size()`` / ``has_next()`` work on the fused
        backend exactly like on the in-process backend), mirrors the device worker-status table into
        ``ctx.stat`` and keeps ``ctx.current_time`` equal to the device logical clock."""
        import threading

        stop = threading.Event()
        side = _dev.new_stream(self.device)

        def pump():
            _dev.set_device(self.device)
            while not stop.is_set():
                with _dev.stream_ctx(side, self.device):
                    ring = self.block.ring.cpu().numpy().view(RESULT_REC_DTYPE)
                    status = self.block.read_status()
                n = self.layout.ring_size
                while True:
                    rec = ring[self._ring_seen % n]
                    if int(rec["seq"]) != self._ring_seen + 1:
                        break
                    ctx.put(PartialResult(
                        data={"accepted": bool(rec["accepted"]), "loss_sum": float(rec["loss_sum"]), "step": float(rec["step"]),
                              "k_after": int(rec["k_after"])},
                        staleness=int(rec["staleness"]), timestamp=int(rec["ts"]), batch_size=int(rec["batch"]),
                        worker_id=int(rec["worker"]), submit_ms=float(rec["t_start_ns"]) * 1e-6,
                        finish_ms=float(rec["t_end_ns"]) * 1e-6))
                    self._ring_seen += 1
                    ctx.current_time = self._ring_seen
                for p in range(self.P):
                    ctx.stat.set_state(p, int(status[p]["staleness"]), float(status[p]["avg_task_ms"]),
                                       bool(status[p]["available"]), int(status[p]["num_tasks"]))
                stop.wait(period_s)

        th = threading.Thread(target=pump, name="afb-result-pump", daemon=True)
        th.start()
        self._pump = (th, stop)
        return th

    def stop_result_pump(self) -> None:
        if getattr(self, "_pump", None):
            self._pump[1].set()
            self._pump[0].join(timeout=5)
            self._pump = None

    def request_stop(self) -> None:
        if self.rank == 0:
            self.block.request_stop()

    def barrier(self) -> None:
        self.symm.barrier()
