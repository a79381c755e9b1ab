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
