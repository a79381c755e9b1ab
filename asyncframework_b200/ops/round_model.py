"""CTA-level interleaving model of the fused round protocol (race detection for the *algorithm*).

``ops/round_sim.py`` is the sequential spec of one round.  This module models what the simulator cannot: the
*concurrency inside a grid* — G CTAs of one worker racing through puller election, dynamic chunk scheduling,
the grid-level reduction, the early clock ticket, last-CTA election, the reset of the per-round counters and
(persistent rounds, ``csrc/kernels/dense_stream.cu::stream_round_serial``) the in-kernel round barrier on
``WorkerCtl.round`` — with several workers sharing one server.  Every CTA is a Python generator that yields
after each access to shared state, in the order the CUDA code performs them (``csrc/include/afb_round_impl.cuh``,
``csrc/kernels/dense_round.cu``); ``explore`` runs them under a seeded random / bursty scheduler and checks:

* every row chunk of every executed round is processed exactly once (no chunk lost or duplicated by the
  static-first-chunk + one-ahead dynamic fetch scheme, no counter reset racing a late fetch);
* the server's ``w`` equals the sum of the contributions of the accepted rounds (no gradient leaks from one
  round into the next through ``g_local``, none is lost);
* ``k`` = number of accepted rounds, ``clock`` = number of finished rounds, ``avail_count`` returns to P;
* all per-round counters are zero at the end; no deadlock (every generator spinning with no progress).

It is sequentially consistent (one access at a time), so it checks protocol logic, not the PTX memory model —
the fences/scopes are covered by the litmus test (`afb_p2p_pingpong`) and compute-sanitizer racecheck on the GPU.
Mutations (``bug=...``) remove one synchronisation edge each; the tests assert the checker catches every one.
"""
from __future__ import annotations

import random
from dataclasses import dataclass, field
from typing import Dict, Generator, List, Optional

SPIN = "spin"


@dataclass
class Server:
    P: int
    max_iters: int = 1 << 40
    taw: int = 1 << 40
    clock: int = 0
    k: int = 0
    w: int = 0
    avail: int = 0
    stop: int = 0
    ring: List[tuple] = field(default_factory=list)

    def __post_init__(self):
        self.avail = self.P


@dataclass
class WorkerCtl:
    chunk_ctr: int = 0
    done_ctas: int = 0
    pull_ticket: int = 0
    pull_flag: int = 0
    round: int = 0
    proceed: int = 0
    batch_count: int = 0
    ts: int = 0
    g: int = 0
    tk_flag: int = 0
    tk_accept: int = 0
    tk_clock_old: int = 0
    rounds_done: int = 0
    last_accepted: int = 0


def contrib(worker: int, rnd: int, chunk: int) -> int:
    """Distinct per (worker, round, chunk) so a sum identifies exactly which chunks were accumulated."""
    return 1 + chunk + 1009 * rnd + 1000003 * worker


class WorkerModel:
    """One logical worker: G CTAs x `rounds` rounds in ONE persistent launch, or one launch per round."""

    def __init__(self, sv: Server, wid: int, G: int, n_chunks: int, bug: Optional[str] = None, sync_peers: Optional[list] = None):
        self.sv, self.wid, self.G, self.n_chunks, self.bug = sv, wid, G, n_chunks, bug
        self.wc = WorkerCtl()
        # BSP mode (ALGO_SYNC_*): `sync_peers` is the list of all workers; each owns a staging buffer
        # stage[parity][q] and arrival flags flag[parity][q] that every peer writes over NVLink, and a local replica
        self.sync_peers = sync_peers
        self.stage = [[0] * sv.P for _ in range(2)]
        self.flag = [[0] * sv.P for _ in range(2)]
        self.w_replica = 0
        self.covered: Dict[int, List[int]] = {}
        self.accepted_rounds: List[int] = []
        self.executed_rounds: List[int] = []
        self.staleness: List[int] = []

    # ------------------------------------------------------------------ one round of one CTA
    def _round(self, cid: int, rnd: int, wait: bool) -> Generator:
        sv, wc, G = self.sv, self.wc, self.G
        # prologue on immutable inputs: first chunk is static (chunk id == CTA id), nothing shared is touched
        c = cid
        if wait and self.bug != "no_round_barrier":
            while wc.round != rnd:                     # round barrier of the persistent kernel
                yield SPIN
        # ---- round_begin_issue: puller election
        t = wc.pull_ticket; wc.pull_ticket = t + 1     # atomicAdd
        yield
        puller = (t == 0)
        is_async = self.sync_peers is None
        ck = (rnd, rnd)                                 # BSP: the local round counter plays the clock
        if puller and is_async:
            ck = (sv.clock, sv.k)                       # one 64-bit load
            yield
        # ---- first dynamic chunk id requested while the pull is in flight
        nxt = wc.chunk_ctr + G; wc.chunk_ctr += 1      # atomicAdd (+ static_chunks)
        yield
        # ---- round_begin_complete
        if puller:
            proceed = 0 if (is_async and (ck[1] >= sv.max_iters or sv.stop)) else 1
            if proceed and is_async:
                sv.avail -= 1
                yield
            wc.ts = ck[0]
            wc.proceed = proceed
            yield
            wc.pull_flag = rnd + 1                      # release
            yield
        else:
            while wc.pull_flag != rnd + 1:
                yield SPIN
        if not wc.proceed:
            t = wc.done_ctas; wc.done_ctas = t + 1
            yield
            if t == G - 1:
                wc.done_ctas = 0; wc.pull_ticket = 0; wc.chunk_ctr = 0
                wc.last_accepted = -1
                yield
                wc.round = rnd + 1
                yield
            return False
        if cid == 0:
            self.executed_rounds.append(rnd)
        cov = self.covered.setdefault(rnd, [0] * self.n_chunks)
        # ---- gradient: chunks, the next id always requested one ahead
        local = 0
        while c < self.n_chunks:
            cov[c] += 1
            local += contrib(self.wid, rnd, c)
            yield
            c = nxt
            if c >= self.n_chunks:
                break
            nxt = wc.chunk_ctr + G; wc.chunk_ctr += 1
            yield
        wc.g += local                                   # red.gpu.add into g_local
        yield
        # ---- round_finish
        wc.batch_count += 1
        yield
        ticket = wc.done_ctas; wc.done_ctas = ticket + 1
        yield
        if ticket == 0 and is_async:                     # early clock ticket (first CTA to finish)
            old_clock, old_k = sv.clock, sv.k
            if sv.taw >= (1 << 31):
                sv.clock += 1; sv.k += 1                 # single 64-bit atomic (never-stale fast path)
                yield
                accept = 1 if old_k < sv.max_iters else 0  # k may overshoot max_iters; the overshoot is never applied
            else:
                sv.clock += 1                            # first atomic: clock only
                yield
                accept = 1 if (old_clock - wc.ts) <= sv.taw else 0
                if accept:
                    old_k = sv.k; sv.k += 1              # second atomic: k
                    yield
                    if old_k >= sv.max_iters:
                        accept = 0
            self.staleness.append(old_clock - wc.ts)
            wc.tk_accept = accept; wc.tk_clock_old = old_clock
            yield
            wc.tk_flag = rnd + 1
            yield
        if ticket != G - 1:
            return True
        if not is_async:
            yield from self._bsp_tail(rnd)
            return True
        # ---- last CTA: push-and-apply, reset, publish
        while wc.tk_flag != rnd + 1:
            yield SPIN
        accept = wc.tk_accept
        g = wc.g
        yield
        if accept:
            sv.w += g                                    # Hogwild reductions into the server
            self.accepted_rounds.append(rnd)
            yield
        if self.bug != "late_g_zero":
            wc.g = 0
            yield
        sv.ring.append((wc.tk_clock_old, self.wid, accept))
        sv.avail += 1
        yield
        wc.last_accepted = accept
        wc.rounds_done += 1
        wc.done_ctas = 0; wc.pull_ticket = 0
        yield
        wc.chunk_ctr = 0; wc.batch_count = 0
        yield
        wc.round = rnd + 1                               # release: everything above is visible first
        yield
        if self.bug == "late_g_zero":
            wc.g = 0                                     # (mutation) zeroing g_local after publication wipes the next round's partial sums
            yield
        return True

    def _bsp_tail(self, rnd: int) -> Generator:
        """Last CTA, BSP: all-gather of g into every peer's staging slot (double-buffered by round parity), arrival
        flags, wait for all P flags, fixed-order reduction, replicated apply."""
        wc, P = self.wc, self.sv.P
        par = 0 if self.bug == "single_buffer" else (rnd & 1)
        g = wc.g
        yield
        for q in range(P):
            self.sync_peers[q].stage[par][self.wid] = g
            yield
        for q in range(P):
            self.sync_peers[q].flag[par][self.wid] = rnd + 1
            yield
        for q in range(P):
            while self.flag[par][q] != rnd + 1:
                yield SPIN
        tot = 0
        for q in range(P):
            tot += self.stage[par][q]
            yield
        self.w_replica += tot
        self.accepted_rounds.append(rnd)
        wc.g = 0
        yield
        wc.last_accepted = 1
        wc.rounds_done += 1
        wc.done_ctas = 0; wc.pull_ticket = 0
        yield
        wc.chunk_ctr = 0; wc.batch_count = 0
        yield
        wc.round = rnd + 1
        yield

    # ------------------------------------------------------------------ pipelined persistent rounds (dense_stream.cu::stream_round_pipe)
    def _pipe_state(self):
        return WorkerCtl()

    def pipelined_cta(self, cid: int, rounds: int) -> Generator:
        """Persistent grid in which round r+1's pull and gradient phase overlap round r's NVLink tail: every piece of
        per-round state (tickets, chunk counter, g_local, pulled w / clock, early ticket) is double-buffered by round
        parity; a CTA may start round r once round r-2 has been PUBLISHED (its buffers are clean), and the last CTA to
        finish round r applies it while everybody else is already streaming round r+1.  The worker therefore computes
        round r+1 on parameters that do not contain its own round-r update (one own update in flight)."""
        sv, wc, G = self.sv, self.wc, self.G
        if not hasattr(self, "pb"):
            self.pb = [self._pipe_state(), self._pipe_state()]
        round0 = wc.round
        yield
        for it in range(rounds):
            rnd = round0 + it
            S = self.pb[0 if self.bug == "pipe_single_buffer" else (rnd & 1)]
            c = cid
            if it >= 2 and self.bug != "pipe_no_reuse_wait":
                while wc.round < rnd - 1:                  # round r-2 published => parity buffers are clean
                    yield SPIN
            t = S.pull_ticket; S.pull_ticket = t + 1
            yield
            puller = (t == 0)
            ck = None
            if puller:
                ck = (sv.clock, sv.k)
                yield
            nxt = S.chunk_ctr + G; S.chunk_ctr += 1
            yield
            if puller:
                S.ts = ck[0]
                S.proceed = 0 if (ck[1] >= sv.max_iters or sv.stop) else 1
                yield
                S.pull_flag = rnd + 1
                yield
            else:
                while S.pull_flag != rnd + 1:
                    yield SPIN
            if not S.proceed:
                t = S.done_ctas; S.done_ctas = t + 1
                yield
                if t == G - 1:
                    while wc.round != rnd:                 # the previous round's tail publishes first
                        yield SPIN
                    S.done_ctas = 0; S.pull_ticket = 0; S.chunk_ctr = 0
                    wc.last_accepted = -1
                    yield
                    wc.round = rnd + 1
                    yield
                return
            if cid == 0:
                self.executed_rounds.append(rnd)
            cov = self.covered.setdefault(rnd, [0] * self.n_chunks)
            local = 0
            while c < self.n_chunks:
                cov[c] += 1
                local += contrib(self.wid, rnd, c)
                yield
                c = nxt
                if c >= self.n_chunks:
                    break
                nxt = S.chunk_ctr + G; S.chunk_ctr += 1
                yield
            S.g += local
            yield
            S.batch_count += 1
            yield
            ticket = S.done_ctas; S.done_ctas = ticket + 1
            yield
            if ticket == 0:
                old_clock, old_k = sv.clock, sv.k
                sv.clock += 1; sv.k += 1
                yield
                S.tk_accept = 1 if old_k < sv.max_iters else 0
                S.tk_clock_old = old_clock
                self.staleness.append(old_clock - S.ts)
                yield
                S.tk_flag = rnd + 1
                yield
            if ticket != G - 1:
                continue
            # ---- tail of round rnd, overlapping everybody else's round rnd+1
            while wc.round != rnd:                         # tails are applied and published in round order
                yield SPIN
            while S.tk_flag != rnd + 1:
                yield SPIN
            accept = S.tk_accept
            g = S.g
            yield
            if accept:
                sv.w += g
                self.accepted_rounds.append(rnd)
                yield
            S.g = 0
            yield
            sv.ring.append((S.tk_clock_old, self.wid, accept))
            yield
            wc.last_accepted = accept
            wc.rounds_done += 1
            S.done_ctas = 0; S.pull_ticket = 0
            yield
            S.chunk_ctr = 0; S.batch_count = 0
            yield
            wc.round = rnd + 1
            yield

    # ------------------------------------------------------------------ launch shapes
    def persistent_cta(self, cid: int, rounds: int) -> Generator:
        round0 = self.wc.round                           # same value in every CTA (see the kernel's comment)
        yield
        for it in range(rounds):
            alive = yield from self._round(cid, round0 + it, wait=(it > 0))
            if not alive:
                return

    def single_round_cta(self, cid: int) -> Generator:
        rnd = self.wc.round
        yield
        yield from self._round(cid, rnd, wait=False)

    # ------------------------------------------------------------------ end-of-run invariants
    def check(self) -> None:
        wc = self.wc
        for rnd in self.executed_rounds:
            cov = self.covered[rnd]
            bad = [i for i, v in enumerate(cov) if v != 1]
            assert not bad, f"worker {self.wid} round {rnd}: chunks processed != once: {[(i, cov[i]) for i in bad][:6]}"
        for rnd in self.covered:
            assert rnd in self.executed_rounds, f"worker {self.wid}: chunk work in a round that was not started ({rnd})"
        for st in [wc] + list(getattr(self, "pb", [])):
            assert st.g == 0 and st.done_ctas == 0 and st.pull_ticket == 0 and st.chunk_ctr == 0 and st.batch_count == 0, \
                f"worker {self.wid}: per-round state not reset: {st}"

    def expected_w(self) -> int:
        return sum(contrib(self.wid, r, c) for r in self.accepted_rounds for c in range(self.n_chunks))


class Deadlock(AssertionError):
    pass


def run_schedule(gens: List[Generator], rng: random.Random, burst: int = 1, max_steps: int = 2_000_000) -> int:
    """Advance the generators one shared-memory access at a time.  `burst` > 1 lets the chosen generator run up to that
    many consecutive steps (adversarial: one CTA races far ahead / another is descheduled for long).  A full sweep in
    which every live generator only spins is a deadlock."""
    live = list(gens)
    steps = 0
    spinning = set()
    while live:
        g = rng.choice(live)
        for _ in range(rng.randint(1, burst)):
            try:
                r = next(g)
            except StopIteration:
                live.remove(g)
                spinning.discard(id(g))
                spinning.clear()
                break
            steps += 1
            if r == SPIN:
                spinning.add(id(g))
                if len(spinning) == len(live):
                    raise Deadlock(f"all {len(live)} live CTAs are spinning")
                break                                    # a spinning CTA gives up the rest of its burst
            spinning.clear()
        if steps > max_steps:
            raise Deadlock("step budget exhausted")
    return steps


def explore(P: int = 2, G: int = 3, n_chunks: int = 7, rounds: int = 4, mode: str = "persist", seed: int = 0, burst: int = 1,
            max_iters: int = 1 << 40, bug: Optional[str] = None, algo: str = "async", taw: int = 1 << 40) -> Dict[str, int]:
    """Run one random schedule of P workers x G CTAs x `rounds` rounds and check every invariant."""
    rng = random.Random(seed)
    sv = Server(P, max_iters=max_iters, taw=taw)
    workers: List[WorkerModel] = []
    for wid in range(P):
        workers.append(WorkerModel(sv, wid, G, n_chunks, bug, sync_peers=workers if algo == "sync" else None))
    steps = 0
    if mode == "persist":
        gens = [w.persistent_cta(c, rounds) for w in workers for c in range(G)]
        steps = run_schedule(gens, rng, burst)
    elif mode == "pipelined":
        gens = [w.pipelined_cta(c, rounds) for w in workers for c in range(G)]
        steps = run_schedule(gens, rng, burst)
    elif mode == "launch":
        # one grid per round per worker; grids of one worker are stream-ordered (kernel boundary = full barrier),
        # different workers' grids interleave freely
        def worker_stream(w: WorkerModel) -> Generator:
            for _ in range(rounds):
                ctas = [w.single_round_cta(c) for c in range(G)]
                live = list(ctas)
                while live:
                    g = rng.choice(live)
                    try:
                        r = next(g)
                    except StopIteration:
                        live.remove(g)
                        continue
                    yield r if len(live) == 1 else None   # a spin inside a grid with other runnable CTAs is not a deadlock signal
        steps = run_schedule([worker_stream(w) for w in workers], rng, burst)
    else:
        raise ValueError(mode)
    for w in workers:
        w.check()
    if algo == "sync":
        exp = sum(contrib(w.wid, r, c) for w in workers for r in range(rounds) for c in range(n_chunks))
        for w in workers:
            assert w.accepted_rounds == list(range(rounds)), f"worker {w.wid} ran rounds {w.accepted_rounds}"
            assert w.w_replica == exp, f"replica of worker {w.wid} diverged: {w.w_replica} != {exp}"
        return {"steps": steps, "accepted": rounds * P, "finished": rounds * P, "k": rounds}
    exp_w = sum(w.expected_w() for w in workers)
    accepted = sum(len(w.accepted_rounds) for w in workers)
    finished = sum(len(w.executed_rounds) for w in workers)
    assert sv.w == exp_w, f"server w {sv.w} != sum of accepted contributions {exp_w}"
    assert sv.clock == finished, f"clock {sv.clock} != finished rounds {finished}"
    if taw >= (1 << 31):
        assert accepted == min(finished, max_iters), f"accepted {accepted} != min(finished {finished}, max_iters {max_iters})"
    else:
        assert accepted <= max_iters and sv.k >= accepted
        for w in workers:
            assert all(st >= 0 for st in w.staleness)
        n_fresh = sum(1 for w in workers for st in w.staleness if st <= taw)
        assert accepted == min(n_fresh, max_iters) or max_iters < n_fresh, f"accepted {accepted} but {n_fresh} results were within taw"
    assert sv.avail == P, f"avail_count {sv.avail} != P"
    assert sorted(r[0] for r in sv.ring) == list(range(finished)), "result-ring sequence numbers are not a permutation of the clock"
    return {"steps": steps, "accepted": accepted, "finished": finished, "k": sv.k}


# ---------------------------------------------------------------------------------------------------------------------
# Device-side ASYNCbarrier (csrc/include/afb_round_impl.cuh::gate_wait_warp): bucket rule with WAVE semantics, the watchdog's
# dynamic threshold, the SSP rule.  One generator per worker = the elected gate warp of its pulling CTA; every access to
# shared state is one step, in the order of the CUDA code.
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class GateState:
    P: int
    threshold: int                    # RoundParams.gate_threshold = floor(P * beta); 0 = no bucket gate
    ssp_bound: int = -1               # RoundParams.ssp_bound; < 0 = off
    wave: int = 0                     # ServerCtl.wave
    avail: int = 0                    # ServerCtl.avail_count
    gate_dyn: int = 0                 # ServerCtl.gate_dyn (watchdog override; 0 = none)
    clock: int = 0
    staleness: List[int] = field(default_factory=list)      # WorkerSlot.staleness (last reported)
    runs: List[tuple] = field(default_factory=list)         # (worker, wave joined, avail seen by the opener or None)
    opened: List[tuple] = field(default_factory=list)       # (wave opened, avail read that justified it, threshold used)

    def __post_init__(self):
        self.avail = self.P
        self.staleness = [0] * self.P


def gate_worker(gs: GateState, wid: int, rounds: int, work_steps: int = 3, dies_after: Optional[int] = None,
                bug: Optional[str] = None) -> Generator:
    """``rounds`` rounds of one worker: gate -> busy (avail_count - 1) -> work -> finish (clock, staleness, avail_count + 1).
    ``dies_after``: the worker stops for good after that many rounds WITHOUT becoming available again mid-round (a rank that
    was killed while busy).  Mutations: ``no_wave_check`` (threshold test only: the pre-wave gate), ``cas_blind`` (opens the
    next wave without the compare)."""
    wave_seen = 0
    last_fail = None
    for r in range(rounds):
        if dies_after is not None and r == dies_after:
            gs.avail -= 1                                 # dies busy: it never reports back
            return
        # ---- gate_wait_warp
        while True:
            open_ = True
            if gs.threshold > 0:
                thr = gs.threshold
                dyn = gs.gate_dyn; yield
                if 0 < dyn < thr:
                    thr = dyn
                w = gs.wave; yield
                if bug != "no_wave_check" and w != wave_seen:
                    wave_seen = w                         # a wave opened since this worker last ran: join it
                else:
                    a = gs.avail; yield
                    if a >= thr:
                        if bug == "no_wave_check":
                            pass                          # the round-1 gate: run whenever the threshold holds
                        elif bug == "cas_blind":
                            gs.wave = w + 1; wave_seen = w + 1
                            gs.opened.append((w + 1, a, thr)); yield
                        else:
                            old = gs.wave                 # atom.cas.acq_rel.sys(wave, w, w + 1)
                            if old == w:
                                gs.wave = w + 1
                                gs.opened.append((w + 1, a, thr))
                            wave_seen = w + 1 if old == w else old
                            yield
                    else:
                        open_ = False
            if open_ and gs.ssp_bound >= 0:
                m = 0
                for q in range(gs.P):                     # one sweep of the status table (not atomic as a whole)
                    m = max(m, gs.staleness[q]); yield
                if m > gs.ssp_bound:
                    open_ = False
            if open_:
                break
            # a failed attempt is a "spin" for the deadlock detector only if nothing it depends on has changed since the
            # previous failed attempt (its reads may be stale with respect to workers that have just finished)
            snap = (gs.wave, gs.avail, gs.gate_dyn, tuple(gs.staleness))
            yield SPIN if snap == last_fail else None
            last_fail = snap
        last_fail = None
        gs.runs.append((wid, wave_seen))
        # ---- pull: the worker becomes busy, remembers the clock it read
        ts = gs.clock
        gs.avail -= 1; yield
        for _ in range(work_steps):
            yield
        # ---- finish: logical clock, own staleness, available again
        st = gs.clock - ts
        gs.clock += 1; yield
        gs.staleness[wid] = st; yield
        gs.avail += 1; yield


def explore_gate(P: int = 4, threshold: int = 2, rounds: int = 5, seed: int = 0, burst: int = 1, ssp_bound: int = -1,
                 dead: Optional[Dict[int, int]] = None, watchdog_after: Optional[int] = None, bug: Optional[str] = None,
                 work_steps: int = 3) -> GateState:
    """One random schedule of P gate warps.  ``dead`` = {worker: rounds completed before it dies}; ``watchdog_after`` = number
    of scheduler steps after which the watchdog declares the dead workers dead (lowers ``gate_dyn`` to threshold - #dead, like
    ``FusedEngine.declare_dead``).  Checks: a worker joins each wave at most once; every wave was opened on a threshold that
    held when it was read; the clock counts every finished round; everybody idle at the end; no deadlock."""
    rng = random.Random(seed)
    gs = GateState(P, threshold, ssp_bound)
    dead = dead or {}
    gens = [gate_worker(gs, w, rounds, work_steps, dead.get(w), bug) for w in range(P)]
    if watchdog_after is not None:
        def watchdog() -> Generator:
            for _ in range(watchdog_after):
                yield
            gs.gate_dyn = max(1, threshold - len(dead)); yield
        gens.append(watchdog())
    run_schedule(gens, rng, burst)
    if threshold > 0 and bug is None:
        per_wave: Dict[tuple, int] = {}
        for wid, wave in gs.runs:
            per_wave[(wid, wave)] = per_wave.get((wid, wave), 0) + 1
        assert all(c == 1 for c in per_wave.values()), f"a worker ran twice in one wave: {[k for k, c in per_wave.items() if c > 1]}"
        assert [o[0] for o in gs.opened] == list(range(1, len(gs.opened) + 1)), f"waves were not opened one at a time: {gs.opened}"
        assert all(a >= thr for _, a, thr in gs.opened), f"a wave was opened below its threshold: {gs.opened}"
    finished = sum(min(rounds, dead.get(w, rounds)) for w in range(P))
    assert gs.clock == finished, f"clock {gs.clock} != finished rounds {finished}"
    assert gs.avail == P - len(dead), f"avail_count {gs.avail} != live workers {P - len(dead)}"
    return gs
