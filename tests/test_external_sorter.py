"""Spillable shuffle collections (reference: util/collection/ExternalSorter.scala:89, ExternalAppendOnlyMap.scala)."""
import glob
import os
import random

import pytest

from asyncframework_b200.storage import external_sorter as xs
from asyncframework_b200.storage.external_sorter import Aggregator, ExternalSorter, SpillConfig, external_combine, external_sort


def _cfg(tmp_path, n=50, **kw):
    return SpillConfig(force_spill_elements=n, batch_size=7, local_dir=str(tmp_path), **kw)


def test_partitioned_output_without_ordering_keeps_every_record(tmp_path):
    recs = [(i, i * i) for i in range(1000)]
    random.Random(1).shuffle(recs)
    with ExternalSorter(4, lambda k: k % 4, config=_cfg(tmp_path)) as s:
        s.insert_all(recs)
        assert s.spill_count >= 19 and len(glob.glob(str(tmp_path / "afb-spill-*.run"))) == s.spill_count
        out = {p: list(it) for p, it in s.partitioned_iterator()}
        assert sorted(out) == [0, 1, 2, 3]
        for p, rows in out.items():
            assert all(k % 4 == p for k, _ in rows)
        assert sorted(kv for rows in out.values() for kv in rows) == sorted(recs)
        m = s.metrics()
        assert m["disk_bytes_spilled"] > 0 and m["memory_bytes_spilled"] > 0 and m["records"] == 1000
    assert glob.glob(str(tmp_path / "afb-spill-*.run")) == []          # stop() removed the runs


@pytest.mark.parametrize("reverse", [False, True])
def test_sorted_runs_merge_into_a_total_order_per_partition(tmp_path, reverse):
    rng = random.Random(2)
    recs = [(rng.randrange(10**6), i) for i in range(3000)]
    with ExternalSorter(3, lambda k: k % 3, key=lambda k: k, reverse=reverse, config=_cfg(tmp_path, 200)) as s:
        s.insert_all(recs)
        assert s.spill_count >= 14
        for p, it in s.partitioned_iterator():
            keys = [k for k, _ in it]
            assert keys == sorted((k for k, _ in recs if k % 3 == p), reverse=reverse)


def test_combiner_folds_values_across_memory_and_spilled_runs(tmp_path):
    rng = random.Random(3)
    recs = [(f"key{rng.randrange(300)}", 1) for _ in range(20000)]
    want = {}
    for k, v in recs:
        want[k] = want.get(k, 0) + v
    with ExternalSorter(5, None, Aggregator.from_reduce(lambda a, b: a + b), config=_cfg(tmp_path, 100)) as s:
        s.partitioner = lambda k: hash(k) % 5
        s.insert_all(recs)
        assert s.spill_count > 10
        got = dict(s.iterator())
    assert got == want


def test_hash_collisions_are_disambiguated_by_equality(tmp_path, monkeypatch):
    monkeypatch.setattr(xs, "_stable_key_hash", lambda k: len(str(k)))          # every same-length key collides
    recs = [(k, 1) for k in ["aa", "bb", "cc", "aa", "ddd", "eee", "bb"] * 40]
    got = dict(external_combine(recs, Aggregator.from_reduce(lambda a, b: a + b), _cfg(tmp_path, 3)))
    assert got == {"aa": 80, "bb": 80, "cc": 40, "ddd": 40, "eee": 40}


def test_grouping_and_external_sort_helpers(tmp_path):
    recs = [(i % 10, i) for i in range(500)]
    seen = {}
    groups = dict(external_combine(recs, Aggregator.grouping(), _cfg(tmp_path, 40), on_metrics=seen.update))
    assert seen["spills"] >= 10
    assert {k: sorted(v) for k, v in groups.items()} == {k: [i for i in range(500) if i % 10 == k] for k in range(10)}
    out = list(external_sort([(k, None) for k in [5, 3, 9, 1, 7] * 30], key=lambda k: -k, config=_cfg(tmp_path, 11)))
    assert [k for k, _ in out] == sorted([5, 3, 9, 1, 7] * 30, reverse=True)


def test_memory_threshold_triggers_spills_from_the_size_estimate(tmp_path):
    cfg = SpillConfig(memory_threshold=64 << 10, batch_size=100, local_dir=str(tmp_path), sample_every=16)
    with ExternalSorter(1, None, None, key=lambda k: k, config=cfg) as s:
        s.insert_all((i, "x" * 200) for i in range(5000))
        assert s.spill_count >= 5 and s.peak_memory > 64 << 10
        assert [k for k, _ in s.iterator()] == list(range(5000))


def test_dataset_ops_spill_on_a_process_cluster_and_report_task_metrics(tmp_path):
    """reduce_by_key / group_by_key / sort_by_key / sort_by / join through the block-manager shuffle with a forced spill
    threshold: results equal the in-memory answers and the spill counters reach the task metrics."""
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.core.env import RuntimeLimits
    from asyncframework_b200.data.dataset import ShardedDataset
    limits = RuntimeLimits(shuffle_spill_elements=64, shuffle_spill_batch=16, local_dir=str(tmp_path))
    sch = ClusterScheduler(2, limits=limits)
    ends = []
    sch.listeners.add(lambda ev: ends.append(ev["metrics"]) if ev["event"] == "task_end" else None)
    try:
        rng = random.Random(5)
        recs = [(rng.randrange(97), rng.randrange(1000)) for _ in range(4000)]
        ds = ShardedDataset.from_records(sch, recs, 4)
        want = {}
        for k, v in recs:
            want[k] = want.get(k, 0) + v
        assert dict(ds.reduce_by_key(lambda a, b: a + b, 3).collect()) == want
        groups = dict(ds.group_by_key(3).collect())
        assert {k: sorted(v) for k, v in groups.items()} == {k: sorted(v for kk, v in recs if kk == k) for k in want}
        assert [k for k, _ in ds.sort_by_key(True, 3).collect()] == sorted(k for k, _ in recs)
        assert ds.sort_by(lambda kv: (kv[1], kv[0]), False, 3).collect() == sorted(recs, key=lambda kv: (kv[1], kv[0]), reverse=True)
        small = ShardedDataset.from_records(sch, [(k, str(k)) for k in range(0, 97, 5)], 2)
        joined = sorted(ds.join(small, 3).collect())
        assert joined == sorted((k, (v, str(k))) for k, v in recs if k % 5 == 0)
        assert glob.glob(str(tmp_path / "afb-spill-*.run")) == []            # every sorter cleaned up after itself
        spilled = [m for m in ends if m.get("spills", 0) > 0]
        assert spilled and all(m["disk_bytes_spilled"] > 0 and m["memory_bytes_spilled"] > 0 for m in spilled)
        assert max(m["spills"] for m in spilled) >= 10                       # 1000 records per map task / 64 per run
    finally:
        sch.stop()


def test_in_process_master_uses_the_process_default_config(tmp_path, monkeypatch):
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.data.dataset import ShardedDataset
    monkeypatch.setattr(xs, "DEFAULT_CONFIG", _cfg(tmp_path, 25))
    counted = []
    orig = xs.ExternalSorter.spill

    def spy(self):
        counted.append(self._n)
        return orig(self)
    monkeypatch.setattr(xs.ExternalSorter, "spill", spy)
    sch = Scheduler(2)
    try:
        ds = ShardedDataset.from_records(sch, [(i % 13, i) for i in range(600)], 3)
        assert dict(ds.reduce_by_key(lambda a, b: a + b).collect()) == {k: sum(i for i in range(600) if i % 13 == k) for k in range(13)}
        assert [k for k, _ in ds.sort_by_key().collect()] == sorted(i % 13 for i in range(600))
        assert counted                                                        # the reduce-side sort really spilled
    finally:
        sch.stop()
