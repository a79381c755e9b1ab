"""Protocol-level race detection: random/bursty interleavings of the CTA-level model of the fused round
(asyncframework_b200/ops/round_model.py), for the one-launch-per-round kernel and the persistent multi-round kernel."""
import pytest

from asyncframework_b200.ops.round_model import Deadlock, explore


@pytest.mark.parametrize("mode", ["persist", "launch"])
@pytest.mark.parametrize("burst", [1, 6, 40])
def test_random_schedules_keep_every_invariant(mode, burst):
    for seed in range(60):
        r = explore(P=2, G=3, n_chunks=7, rounds=4, mode=mode, seed=seed, burst=burst)
        assert r["finished"] == 8 and r["accepted"] == 8


@pytest.mark.parametrize("mode", ["persist", "launch"])
def test_shapes_fewer_chunks_than_ctas_single_cta_many_workers(mode):
    for seed in range(25):
        explore(P=1, G=4, n_chunks=2, rounds=3, mode=mode, seed=seed, burst=5)      # some CTAs never get a chunk
        explore(P=3, G=1, n_chunks=5, rounds=3, mode=mode, seed=seed, burst=5)      # puller == ticket holder == last CTA
        explore(P=4, G=2, n_chunks=9, rounds=2, mode=mode, seed=seed, burst=9)
        explore(P=1, G=5, n_chunks=0, rounds=2, mode=mode, seed=seed, burst=3)      # empty shard


@pytest.mark.parametrize("mode", ["persist", "launch"])
def test_stop_at_max_iters_leaves_clean_state(mode):
    """k reaches max_iters in the middle of the run: later rounds must stop at the pull, apply nothing, reset everything —
    and a persistent grid must exit instead of spinning on the next round."""
    for seed in range(60):
        r = explore(P=2, G=3, n_chunks=5, rounds=6, mode=mode, seed=seed, burst=7, max_iters=5)
        assert r["accepted"] == 5 and r["finished"] >= 5


def _fails_somewhere(**kw):
    for seed in range(300):
        try:
            explore(seed=seed, **kw)
        except (AssertionError, Deadlock):
            return seed
    return None


def test_checker_catches_a_missing_round_barrier():
    """Without the spin on WorkerCtl.round a fast CTA enters round r+1 while round r's last CTA is still applying /
    resetting: tickets, chunk counters and g_local get corrupted.  The model must expose that."""
    assert _fails_somewhere(P=1, G=3, n_chunks=6, rounds=3, mode="persist", burst=12, bug="no_round_barrier") is not None


def test_checker_catches_g_local_zeroed_after_publication():
    """Every reset must precede the publication of the round word: zeroing g_local afterwards wipes partial sums the
    next round's fast CTAs have already flushed."""
    assert _fails_somewhere(P=1, G=3, n_chunks=6, rounds=3, mode="persist", burst=12, bug="late_g_zero") is not None


@pytest.mark.parametrize("mode", ["persist", "launch"])
def test_bsp_allgather_replicas_identical_under_any_schedule(mode):
    """BSP rounds: every worker's replica ends up with the same, complete sum although fast workers run up to one round
    ahead (staging slots and flags are double-buffered by round parity)."""
    for seed in range(40):
        explore(P=3, G=2, n_chunks=5, rounds=5, mode=mode, seed=seed, burst=1 + seed % 25, algo="sync")
        explore(P=2, G=1, n_chunks=3, rounds=6, mode=mode, seed=seed, burst=30, algo="sync")


def test_checker_catches_single_buffered_bsp_staging():
    """Without the parity double-buffer a worker that is one round ahead overwrites a slot a slow peer has not reduced yet."""
    assert _fails_somewhere(P=3, G=1, n_chunks=3, rounds=5, mode="persist", burst=30, bug="single_buffer", algo="sync") is not None


@pytest.mark.parametrize("mode", ["persist", "launch"])
def test_finite_taw_two_atomic_path(mode):
    """Finite taw: clock tick, staleness gate, then a second atomic on k.  Rejected rounds apply nothing, still publish a
    ring record, free the worker and reset; accepted == number of results whose staleness was within taw."""
    saw_reject = False
    for seed in range(60):
        r = explore(P=3, G=2, n_chunks=4, rounds=4, mode=mode, seed=seed, burst=1 + seed % 20, taw=0)
        saw_reject |= r["accepted"] < r["finished"]
    assert saw_reject


@pytest.mark.parametrize("burst", [1, 8, 60])
def test_pipelined_rounds_overlapping_tail(burst):
    """Design check for the pipelined persistent kernel: round r+1 streams while round r's last CTA applies; state is
    double-buffered by parity; a CTA may enter round r only after round r-2 has been published."""
    for seed in range(60):
        r = explore(P=2, G=3, n_chunks=7, rounds=6, mode="pipelined", seed=seed, burst=burst)
        assert r["finished"] == 12 and r["accepted"] == 12
        explore(P=1, G=4, n_chunks=3, rounds=5, mode="pipelined", seed=seed, burst=burst)
        explore(P=3, G=1, n_chunks=2, rounds=4, mode="pipelined", seed=seed, burst=burst)
        r = explore(P=2, G=3, n_chunks=5, rounds=7, mode="pipelined", seed=seed, burst=burst, max_iters=5)
        assert r["accepted"] == 5


def test_checker_catches_pipelined_buffer_reuse_bugs():
    assert _fails_somewhere(P=1, G=3, n_chunks=6, rounds=5, mode="pipelined", burst=25, bug="pipe_no_reuse_wait") is not None
    assert _fails_somewhere(P=1, G=3, n_chunks=6, rounds=5, mode="pipelined", burst=25, bug="pipe_single_buffer") is not None


@pytest.mark.parametrize("P,threshold,burst", [(4, 2, 1), (4, 4, 3), (8, 5, 2), (3, 1, 4), (6, 0, 1)])
def test_wave_gate_keeps_its_invariants_under_random_schedules(P, threshold, burst):
    """Device ASYNCbarrier (gate_wait_warp): a worker joins every wave at most once, waves open one at a time and only on a
    threshold that held when it was read, nobody starves (beta = 1 included), for 150 random schedules per shape."""
    from asyncframework_b200.ops.round_model import explore_gate
    for seed in range(150):
        gs = explore_gate(P=P, threshold=threshold, rounds=6, seed=seed, burst=burst)
        assert len(gs.runs) == 6 * P
        if threshold:
            assert 6 <= len(gs.opened) <= 6 * P            # at least one wave per round of the slowest worker


def test_wave_gate_differs_from_a_bare_threshold_test():
    """The mutation that drops the wave word (the round-1 gate: run whenever avail_count >= threshold) lets a fast worker run
    again and again while the others have not started — the checker sees a worker twice in one wave under the real protocol
    bookkeeping, i.e. the invariant is not vacuous; a blind wave increment (no compare) double-opens waves."""
    from asyncframework_b200.ops.round_model import explore_gate
    laps = 0
    for seed in range(60):
        gs = explore_gate(P=4, threshold=2, rounds=6, seed=seed, burst=6, bug="no_wave_check")
        firsts = [i for i, (w, _) in enumerate(gs.runs) if w == 0]
        laps = max(laps, max(b - a for a, b in zip(firsts, firsts[1:])) if len(firsts) > 1 else 0)
        by_worker = {}
        for w, _ in gs.runs[:8]:
            by_worker[w] = by_worker.get(w, 0) + 1
        if max(by_worker.values()) >= 4:                    # one worker took half of the first eight slots: no wave discipline
            break
    else:
        pytest.fail("the bare threshold gate never let a worker run ahead: the mutation is not observable")
    hit = 0
    for seed in range(200):
        gs = explore_gate(P=4, threshold=2, rounds=4, seed=seed, burst=2, bug="cas_blind")
        waves = [o[0] for o in gs.opened]
        hit += waves != list(range(1, len(waves) + 1))
    assert hit > 0                                           # two openers both "opened" the same wave somewhere


def test_gate_with_a_dead_worker_needs_the_watchdog_and_then_continues_with_p_minus_1():
    """beta = 1 with one rank killed mid-run: without the watchdog's dynamic threshold the survivors wait forever (the model
    reports the deadlock); once `gate_dyn` is lowered (FusedEngine.declare_dead) they finish all their rounds."""
    from asyncframework_b200.ops.round_model import Deadlock, explore_gate
    with pytest.raises(Deadlock):
        explore_gate(P=4, threshold=4, rounds=5, seed=1, dead={2: 2})
    for seed in range(80):
        gs = explore_gate(P=4, threshold=4, rounds=5, seed=seed, burst=2, dead={2: 2}, watchdog_after=40)
        assert gs.gate_dyn == 3 and sum(1 for w, _ in gs.runs if w != 2) == 15 and sum(1 for w, _ in gs.runs if w == 2) == 2


def test_ssp_gate_follows_the_reference_rule_including_its_wedge():
    """SSP as the reference defines it (workerState.getMaxStaleness: max over the workers' LAST REPORTED staleness <= s).  With a
    bound no staleness can exceed, every schedule completes.  With a tight bound the rule can wedge — a worker that reported a
    large staleness and now waits keeps everybody out, itself included — which the model reports as a deadlock (on the device
    the bounded spin turns it into an abort): the bound has to leave room for the natural staleness of P concurrent workers."""
    from asyncframework_b200.ops.round_model import Deadlock, explore_gate
    P, rounds = 3, 5
    for seed in range(80):
        gs = explore_gate(P=P, threshold=0, rounds=rounds, seed=seed, burst=2, ssp_bound=(P - 1) * rounds)
        assert len(gs.runs) == P * rounds and gs.clock == P * rounds and max(gs.staleness) <= (P - 1) * rounds
    wedged = 0
    for seed in range(40):
        try:
            explore_gate(P=P, threshold=0, rounds=rounds, seed=seed, burst=2, ssp_bound=0)
        except Deadlock:
            wedged += 1
    assert wedged > 0
