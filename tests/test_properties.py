"""Property-based tests (hypothesis) for the data layer invariants."""
import os

import numpy as np
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from asyncframework_b200.core import Scheduler
from asyncframework_b200.core.sampler import bernoulli_mask, hash_u24, threshold24
from asyncframework_b200.data import ShardedDataset, load_libsvm_file, save_as_libsvm_file
from asyncframework_b200.data.batch import Shard, concat_shards, padded_ld

_SCHED = Scheduler(2)
COMMON = dict(deadline=None, max_examples=25, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])


def _random_sparse(rng, n, d, density):
    X = rng.randn(n, d) * (rng.rand(n, d) < density)
    y = rng.randint(0, 3, size=n).astype(np.float64)
    return X, y


@settings(**COMMON)
@given(n=st.integers(1, 60), d=st.integers(1, 25), density=st.floats(0.0, 1.0), P=st.integers(1, 5), seed=st.integers(0, 10**6))
def test_libsvm_save_load_roundtrip(tmp_path_factory, n, d, density, P, seed):
    rng = np.random.RandomState(seed)
    X, y = _random_sparse(rng, n, d, density)
    ds = ShardedDataset.from_arrays(_SCHED, X, y, P)
    out = str(tmp_path_factory.mktemp("svm") / "o")
    save_as_libsvm_file(ds, out)
    back = load_libsvm_file(_SCHED, out, num_features=d, num_partitions=P, layout="dense")
    got = torch.cat([s.X[:, :d] for s in back.shards()])
    lab = torch.cat([s.y for s in back.shards()])
    exp = torch.cat([s.X[:, :d] for s in ds.shards()])
    explab = torch.cat([s.y for s in ds.shards()])
    assert got.shape == exp.shape and torch.equal(got, exp) and torch.equal(lab, explab)     # %.17g round-trips doubles


@settings(**COMMON)
@given(n=st.integers(0, 200), P0=st.integers(1, 6), P1=st.integers(1, 7))
def test_repartition_preserves_rows_and_balances(n, P0, P1):
    X = np.arange(n * 3, dtype=np.float64).reshape(n, 3)
    y = np.arange(n, dtype=np.float64)
    if n == 0:
        return
    ds = ShardedDataset.from_arrays(_SCHED, X, y, P0).repartition(P1)
    sizes = ds.partition_sizes()
    assert sum(sizes) == n and max(sizes) - min(sizes) <= P0          # round-robin from each input partition
    labels = sorted(torch.cat([s.y for s in ds.shards()]).tolist())
    assert labels == list(map(float, range(n)))
    zi = ds.zip_with_index()
    assert sorted(torch.cat([s.gidx for s in zi.shards()]).tolist()) == list(range(n))


@settings(**COMMON)
@given(n=st.integers(1, 80), d=st.integers(1, 20), density=st.floats(0.05, 0.9), seed=st.integers(0, 10**6))
def test_csr_dense_equivalence(n, d, density, seed):
    rng = np.random.RandomState(seed)
    X, y = _random_sparse(rng, n, d, density)
    Xp = torch.zeros(n, padded_ld(d), dtype=torch.float64)
    Xp[:, :d] = torch.from_numpy(X)
    dense = Shard(d, torch.from_numpy(y), Xp, gidx=torch.arange(n))
    csr = dense.to_csr()
    w = torch.from_numpy(rng.randn(d))
    r = torch.from_numpy(rng.randn(n))
    assert torch.allclose(dense.matvec(w), csr.matvec(w), atol=1e-12) and torch.allclose(dense.rmatvec(r), csr.rmatvec(r), atol=1e-12)
    rows = torch.from_numpy(np.sort(rng.choice(n, size=rng.randint(0, n + 1), replace=False)))
    a, b = dense.select(rows), csr.select(rows)
    assert torch.equal(a.gidx, b.gidx) and torch.allclose(a.matvec(w), b.matvec(w), atol=1e-12)
    assert torch.equal(b.to_dense().X, a.X)
    cat = concat_shards([csr.select(rows), csr])
    assert cat.num_rows == rows.numel() + n and torch.allclose(cat.matvec(w)[rows.numel():], csr.matvec(w), atol=1e-12)


@settings(**COMMON)
@given(seed=st.integers(0, 2**63 - 1), frac=st.floats(0.0, 1.0), n=st.integers(1, 3000))
def test_sampler_properties(seed, frac, n):
    idx = np.arange(n) * 7 + 3
    h = hash_u24(seed, idx)
    assert h.min() >= 0 and h.max() < (1 << 24)
    m = bernoulli_mask(seed, idx, frac)
    assert np.array_equal(m, h < threshold24(frac))
    # monotone in the fraction: a row selected at rate f stays selected at any higher rate
    assert not (m & ~bernoulli_mask(seed, idx, min(1.0, frac + 0.1))).any()
    # membership depends only on (seed, global index): any sub-range gives the same answer
    assert np.array_equal(bernoulli_mask(seed, idx[n // 2:], frac), m[n // 2:])


@settings(**COMMON)
@given(vals=st.lists(st.integers(-1000, 1000), min_size=1, max_size=200), P=st.integers(1, 9), depth=st.integers(1, 4))
def test_tree_aggregate_equals_aggregate(vals, P, depth):
    ds = ShardedDataset.from_records(_SCHED, vals, P)
    seq, comb = (lambda a, x: a + x), (lambda a, b: a + b)
    assert ds.tree_aggregate(lambda: 0, seq, comb, depth=depth) == ds.aggregate(lambda: 0, seq, comb) == sum(vals)
    assert ds.reduce(lambda a, b: a + b) == sum(vals)


@settings(**COMMON)
@given(recs=st.lists(st.tuples(st.integers(-20, 20), st.integers(-5, 5)), max_size=120), P=st.integers(1, 4),
       spill=st.integers(1, 40), batch=st.integers(1, 9), mode=st.sampled_from(["plain", "sorted", "sorted_desc", "combine", "combine_sorted"]))
def test_external_sorter_equals_the_in_memory_answer(tmp_path_factory, recs, P, spill, batch, mode):
    """Whatever the spill cadence, batch size and partition count: partitioned (+ ordered) (+ combined) output of the spillable
    sorter == the plain in-memory computation."""
    from asyncframework_b200.storage.external_sorter import Aggregator, ExternalSorter, SpillConfig
    cfg = SpillConfig(force_spill_elements=spill, batch_size=batch, local_dir=str(tmp_path_factory.mktemp("spill")))
    part = lambda k: k % P                              # noqa: E731
    agg = Aggregator.from_reduce(lambda a, b: a + b) if mode.startswith("combine") else None
    key = (lambda k: k) if "sorted" in mode else None
    reverse = mode == "sorted_desc"
    with ExternalSorter(P, part, agg, key=key, reverse=reverse, config=cfg) as s:
        s.insert_all(recs)
        out = {p: list(it) for p, it in s.partitioned_iterator()}
    for p in range(P):
        mine = [(k, v) for k, v in recs if k % P == p]
        if agg is not None:
            want = {}
            for k, v in mine:
                want[k] = want.get(k, 0) + v
            assert dict(out[p]) == want and len(out[p]) == len(want)
            if key is not None:
                assert [k for k, _ in out[p]] == sorted(want)
        elif key is not None:
            assert [k for k, _ in out[p]] == sorted((k for k, _ in mine), reverse=reverse)
            assert sorted(out[p]) == sorted(mine)
        else:
            assert sorted(out[p]) == sorted(mine)
