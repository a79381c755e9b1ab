"""Wider RDD surface (PairRDDFunctionsSuite / RDDSuite / SortingSuite / ShuffleSuite scenarios) on the in-process master,
and the block-manager shuffle on executor processes."""
import math
import os

import pytest

from asyncframework_b200 import AsyncSession


@pytest.fixture(scope="module")
def sc():
    with AsyncSession("local[3]") as s:
        yield s


def test_joins_cogroup_and_set_operations(sc):
    a = sc.parallelize([(1, "a"), (2, "b"), (2, "bb"), (3, "c")], 3)
    b = sc.parallelize([(2, "x"), (3, "y"), (3, "yy"), (4, "z")], 2)
    assert sorted(a.join(b).collect()) == [(2, ("b", "x")), (2, ("bb", "x")), (3, ("c", "y")), (3, ("c", "yy"))]
    assert sorted(a.left_outer_join(b).collect())[0] == (1, ("a", None))
    assert (4, (None, "z")) in a.right_outer_join(b).collect()
    assert sorted(k for k, _ in a.full_outer_join(b).collect()) == [1, 2, 2, 3, 3, 4]
    cg = dict(a.cogroup(b).collect())
    assert sorted(cg[2][0]) == ["b", "bb"] and cg[2][1] == ["x"] and cg[1] == (["a"], []) and cg[4] == ([], ["z"])
    assert sorted(a.subtract_by_key(b).collect()) == [(1, "a")]
    x, y = sc.parallelize(range(10), 3), sc.parallelize(range(5, 15), 2)
    assert sorted(x.subtract(y).collect()) == [0, 1, 2, 3, 4] and sorted(x.intersection(y).collect()) == [5, 6, 7, 8, 9]
    assert sorted(sc.parallelize([1, 2], 2).cartesian(sc.parallelize("ab", 2)).collect()) == [(1, "a"), (1, "b"), (2, "a"), (2, "b")]
    assert a.lookup(2) == ["b", "bb"] and a.collect_as_map()[3] == "c"


def test_combine_aggregate_fold_by_key_and_partition_by(sc):
    kv = sc.parallelize([(i % 4, i) for i in range(40)], 4)
    assert dict(kv.aggregate_by_key(lambda: (0, 0), lambda a, v: (a[0] + v, a[1] + 1), lambda a, b: (a[0] + b[0], a[1] + b[1])).collect()) == \
        {k: (sum(i for i in range(40) if i % 4 == k), 10) for k in range(4)}
    assert dict(kv.fold_by_key(lambda: 0, lambda a, b: a + b).collect())[1] == sum(range(1, 40, 4))
    assert dict(kv.combine_by_key(lambda v: [v], lambda c, v: c + [v], lambda a, b: a + b).map_values(len).collect()) == {k: 10 for k in range(4)}
    pb = kv.partition_by(2, lambda k: k % 2)
    assert [sorted({k for k, _ in part}) for part in pb.glom().collect()] == [[0, 2], [1, 3]]


def test_sort_by_key_uses_a_range_partitioner(sc):
    import random
    rng = random.Random(3)
    data = [(rng.randrange(10_000), i) for i in range(2000)]
    srt = sc.parallelize(data, 4).sort_by_key(num_partitions=4)
    parts = srt.glom().collect()
    assert [k for k, _ in srt.collect()] == sorted(k for k, _ in data)                 # concatenation IS the total order
    assert all(len(p) > 200 for p in parts)                                             # roughly balanced ranges
    assert all(parts[i][-1][0] <= parts[i + 1][0][0] for i in range(3))
    desc = sc.parallelize(data, 4).sort_by_key(ascending=False).keys().collect()
    assert desc == sorted((k for k, _ in data), reverse=True)


def test_numeric_actions_and_misc(sc, tmp_path):
    ds = sc.parallelize([1.0, 2.0, 3.0, 4.0, 10.0], 3)
    st = ds.stats()
    assert st["count"] == 5 and st["mean"] == 4.0 and math.isclose(st["variance"], 10.0) and st["min"] == 1.0 and st["max"] == 10.0
    assert math.isclose(ds.stdev(), math.sqrt(10.0)) and math.isclose(st["sample_variance"], 12.5)
    assert ds.histogram([0, 2, 4, 10]) == ([0, 2, 4, 10], [1, 2, 2]) and ds.histogram(3)[1] == [3, 1, 1]
    assert ds.fold(lambda: 0.0, lambda a, b: a + b) == 20.0 and ds.take_ordered(2) == [1.0, 2.0]
    assert not ds.is_empty() and sc.parallelize([], 2).is_empty()
    a, b = sc.parallelize(range(1000), 4).random_split([0.7, 0.3], seed=5)
    na, nb = a.count(), b.count()
    assert na + nb == 1000 and 620 < na < 780 and not set(a.collect()) & set(b.collect())
    out = str(tmp_path / "txt")
    sc.parallelize(range(10), 3).save_as_text_file(out)
    assert sorted(os.listdir(out)) == ["_SUCCESS", "part-00000", "part-00001", "part-00002"]
    assert sorted(int(l) for l in sc.text_file(out + "/part-00001").collect() + sc.text_file(out + "/part-00000").collect()
                  + sc.text_file(out + "/part-00002").collect()) == list(range(10))
    assert list(sc.parallelize(range(7), 3).to_local_iterator()) == list(range(7))
    lin = sc.parallelize(range(4), 2).map(lambda x: x).filter(lambda x: True).cache().set_name("final")
    dbg = lin.to_debug_string().splitlines()
    assert dbg[0].startswith("final [2 partitions, cached]") and len(dbg) == 3 and dbg[2].lstrip().startswith("+- records")
    assert lin.count() == 4 and lin._cached_records is not None
    assert lin.unpersist()._cached_records is None and lin.count() == 4


def test_shuffle_goes_through_executor_block_managers_on_cluster_masters():
    with AsyncSession("local-cluster[2]") as sc:
        kv = sc.parallelize([(i % 10, i) for i in range(1000)], 4)
        red = kv.reduce_by_key(lambda a, b: a + b, 3)
        assert dict(red.collect()) == {k: sum(i for i in range(1000) if i % 10 == k) for k in range(10)}
        sh = red.parent
        assert sh.shuffle_id is not None and sum(sh.map_output_bytes) > 0
        st = sc.scheduler.storage_status()
        blocks = {b: m.executor_id for m, s in st.items() for b in s["blocks"] if b.startswith(f"shuffle_{sh.shuffle_id}_")}
        assert blocks and set(blocks.values()) == {"0", "1"}                         # written where the map tasks ran
        assert all(b.count("_") == 3 for b in blocks)
        # the driver never saw the records: only MapStatus sizes came back from the map stage
        j = sc.parallelize([(k, str(k)) for k in range(10)], 2).join(red)
        assert sorted(j.collect())[3] == (3, ("3", sum(i for i in range(1000) if i % 10 == 3)))
        # fetch failure: a map output disappears -> the reducer recomputes that bucket from lineage
        victim = sorted(blocks)[0]
        assert sc.scheduler.env.bm_master.remove_block(victim) == 1
        assert dict(red.collect())[7] == sum(i for i in range(1000) if i % 10 == 7)
        assert sh.remove_shuffle() >= len(blocks) - 1
        assert not any(b.startswith(f"shuffle_{sh.shuffle_id}_") for s in sc.scheduler.storage_status().values() for b in s["blocks"])


def test_sampling_with_replacement_io_helpers_and_shell(sc, tmp_path):
    import subprocess
    import sys
    import numpy as np
    from asyncframework_b200.core.sampler import poisson_counts
    c = poisson_counts(7, np.arange(200_000), 1.5)
    assert abs(c.mean() - 1.5) < 0.02 and abs(c.var() - 1.5) < 0.05 and c.max() > 5          # Poisson(1.5)
    assert (poisson_counts(7, np.arange(100), 1.5) == c[:100]).all()                          # replayable per row
    ds = sc.parallelize(range(5000), 4)
    boot = ds.sample(True, 1.0, seed=3).collect()
    assert 4600 < len(boot) < 5400 and len(set(boot)) < len(boot) and max(np.bincount(boot)) >= 3
    assert ds.sample(True, 1.0, seed=3).collect() == boot                                     # deterministic in the seed
    assert sc.range(10).sum() == 45 and sc.range(2, 10, 2).collect() == [2, 4, 6, 8] and sc.empty_dataset().count() == 0
    assert sc.union([sc.range(3), sc.range(3, 5), sc.range(5, 6)]).collect() == [0, 1, 2, 3, 4, 5]
    d = tmp_path / "txt"
    d.mkdir()
    (d / "a.txt").write_text("x\ny\n")
    (d / "b.txt").write_text("z\n")
    assert sorted(os.path.basename(k) for k, _ in sc.whole_text_files(str(d)).collect()) == ["a.txt", "b.txt"]
    out = str(tmp_path / "pk")
    sc.parallelize([{"i": i} for i in range(7)], 3).save_as_pickle_file(out)
    back = sc.pickle_file(out)
    assert back.num_partitions == 3 and sorted(r["i"] for r in back.collect()) == list(range(7))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "asyncframework_b200.shell", "--master", "local[2]", "--", "-c",
                          "print('SUM', sc.parallelize(range(10)).map(lambda x: x * x).reduce(lambda a, b: a + b))"],
                         capture_output=True, text=True, timeout=120, cwd=root, env={**os.environ, "PYTHONPATH": root})
    assert res.returncode == 0 and "SUM 285" in res.stdout, res.stderr[-800:]


def test_narrow_coalesce_zip_partitions_pipe_job_groups_and_files(sc, tmp_path):
    import threading
    from asyncframework_b200.core.scheduler import JobCancelled
    from asyncframework_b200.core.session import SparkFiles
    ds = sc.parallelize(range(12), 6)
    co = ds.coalesce(2, shuffle=False)
    assert co.num_partitions == 2 and co.glom().collect() == [[0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11]]      # contiguous, order kept
    assert ds.coalesce(50, shuffle=False).num_partitions == 6
    zp = ds.zip_partitions(ds.map(lambda x: x * 10), lambda a, b: [sum(a) + sum(b)])
    assert zp.collect() == [(2 * p + 2 * p + 1) * 11 for p in range(6)]
    assert sc.parallelize(["b", "a", "c", "a"], 2).pipe("sort").collect() == ["a", "b", "a", "c"]
    with pytest.raises(Exception, match="status"):
        sc.parallelize(["x"], 1).pipe(["false"]).collect()
    # job groups: cancel only the jobs of one group
    sc.set_job_group("slow-group")
    gate = threading.Event()
    w1 = sc.scheduler.submit_job(ds, lambda t, it: gate.wait(10))
    sc.set_job_group("other")
    w2 = sc.scheduler.submit_job(ds, lambda t, it: sum(it))
    sc.clear_job_group()
    assert w1.job_id in sc.scheduler.active_job_ids("slow-group")
    assert sc.cancel_job_group("slow-group") == 1
    gate.set()
    with pytest.raises(JobCancelled):
        w1.wait(30)
    assert w2.wait(30)
    # files
    f = tmp_path / "lookup.txt"
    f.write_text("forty-two")
    sc.add_file(str(f))
    assert ds.map(lambda x, files=sc.files: open(SparkFiles.get("lookup.txt", files)).read()).distinct().collect() == ["forty-two"]
    assert sc.version and sc.application_id.startswith("local-") and sc.ui_web_url is None


def test_persist_levels_spill_to_the_executor_disk_store_and_files_reach_executor_processes(tmp_path):
    from asyncframework_b200.conf import AsyncConf
    from asyncframework_b200.core.session import SparkFiles
    from asyncframework_b200.storage import StorageLevel
    conf = AsyncConf(load_defaults=False).set("spark.executor.memory", "1m").set("spark.memory.fraction", 0.5)   # 512 KB of storage memory
    with AsyncSession("local-cluster[2]", conf=conf) as sc:
        big = sc.parallelize([bytes(2000)] * 800, 4)                                  # ~400 KB per partition, 2 partitions per executor
        kept = big.map(lambda b: b).persist(StorageLevel.MEMORY_AND_DISK)
        assert kept.count() == 800 and kept.get_storage_level() == StorageLevel.MEMORY_AND_DISK
        st = sc.scheduler.storage_status()
        mine = {b: v for s in st.values() for b, v in s["blocks"].items() if b.startswith(f"rdd_{kept.ds_id}_")}
        assert len(mine) == 4 and any(disk > 0 for _, disk in mine.values())          # what did not fit in memory is on disk
        assert kept.count() == 800
        mem_only = big.map(lambda b: b + b"").cache()
        assert mem_only.count() == 800
        st = sc.scheduler.storage_status()
        mo = {b: v for s in st.values() for b, v in s["blocks"].items() if b.startswith(f"rdd_{mem_only.ds_id}_")}
        assert all(disk == 0 for _, disk in mo.values())                               # MEMORY_ONLY never spills
        f = tmp_path / "side.txt"
        f.write_text("side-data")
        sc.add_file(str(f))
        out = sc.parallelize(range(4), 2).map(lambda x, files=sc.files: (os.getpid(), open(SparkFiles.get("side.txt", files)).read())).collect()
        assert {v for _, v in out} == {"side-data"} and os.getpid() not in {p for p, _ in out}


def test_partitioner_classes_status_tracker_and_memory_status(sc):
    from asyncframework_b200.data import HashPartitioner, RangePartitioner
    kv = sc.parallelize([(i, i * i) for i in range(100)], 4)
    hp = kv.partition_by(3, HashPartitioner(3))
    assert all({k % 3 for k, _ in part} == {p} for p, part in enumerate(hp.glom().collect())) and HashPartitioner(3) == HashPartitioner(3)
    rp = RangePartitioner(4, kv)
    parts = kv.partition_by(4, rp).glom().collect()
    assert all(max(k for k, _ in parts[i]) <= min(k for k, _ in parts[i + 1]) for i in range(3)) and all(len(p) > 10 for p in parts)
    assert RangePartitioner(3, bounds=[10, 20])(5) == 0 and RangePartitioner(3, bounds=[10, 20])(10) == 0 and RangePartitioner(3, bounds=[10, 20])(25) == 2
    sc.start_ui()
    n = kv.count()
    tr = sc.status_tracker
    assert n == 100 and tr.get_active_job_ids() == []
    last = sc.status_store.job_list()[0]["job_id"]
    assert tr.get_job_info(last)["status"] == "SUCCEEDED" and len(tr.get_executor_infos()) == 3
    assert sc.get_executor_memory_status() == {}                      # in-process master: no block managers
    with AsyncSession("local-cluster[2]") as c2:
        assert c2.parallelize(range(4), 2).count() == 4               # executors (and their block managers) are up once a job ran
        ms = c2.get_executor_memory_status()
        assert len(ms) == 3 and all(free <= mx for mx, free in ms.values())


def test_tree_reduce_and_hyperloglog_count_approx_distinct(sc):
    ds = sc.parallelize(range(1, 101), 7)
    assert ds.tree_reduce(lambda a, b: a + b) == 5050 and ds.tree_reduce(max, depth=3) == 100
    with pytest.raises(ValueError):
        sc.parallelize([], 2).tree_reduce(lambda a, b: a + b)
    import random
    rng = random.Random(1)
    data = [rng.randrange(20_000) for _ in range(60_000)]
    true = len(set(data))
    est = sc.parallelize(data, 4).count_approx_distinct(0.02)
    assert abs(est - true) / true < 0.06, (est, true)                    # 3 sigma at relative_sd = 0.02
    assert sc.parallelize(["a", "b", "a", "c"] * 10, 2).count_approx_distinct() == 3     # small range: linear counting is exact-ish
    assert ds.id() == ds.ds_id


def test_group_by_flat_map_values_stratified_sampling_and_sorted_repartition(sc):
    """PairRDDFunctions / OrderedRDDFunctions leftovers: groupBy, flatMapValues, sampleByKey, repartitionAndSortWithinPartitions."""
    ds = sc.parallelize(range(60), 4)
    groups = dict(ds.group_by(lambda x: x % 3, 2).collect())
    assert {k: sorted(v) for k, v in groups.items()} == {r: [x for x in range(60) if x % 3 == r] for r in range(3)}
    pairs = sc.parallelize([("a", 2), ("b", 0), ("c", 3)], 2)
    assert sorted(pairs.flat_map_values(lambda n: range(n)).collect()) == [("a", 0), ("a", 1), ("c", 0), ("c", 1), ("c", 2)]
    strat = sc.parallelize([("x", i) for i in range(2000)] + [("y", i) for i in range(2000)], 4)
    s1 = strat.sample_by_key(False, {"x": 0.5, "y": 0.05}, seed=7).collect()
    assert s1 == strat.sample_by_key(False, {"x": 0.5, "y": 0.05}, seed=7).collect()          # deterministic in the seed
    nx, ny = sum(1 for k, _ in s1 if k == "x"), sum(1 for k, _ in s1 if k == "y")
    assert 900 < nx < 1100 and 60 < ny < 140
    s2 = strat.sample_by_key(True, {"x": 2.0, "y": 0.0}, seed=7).collect()
    assert 3700 < len(s2) < 4300 and all(k == "x" for k, _ in s2)
    with pytest.raises(Exception, match="no sampling fraction"):
        strat.sample_by_key(False, {"x": 0.5}, seed=1).collect()
    import random
    rng = random.Random(3)
    recs = [(rng.randrange(1000), i) for i in range(500)]
    out = sc.parallelize(recs, 5).repartition_and_sort_within_partitions(3, partitioner=lambda k: k % 3, ascending=False)
    parts = out.glom().collect()
    assert len(parts) == 3
    for p, rows in enumerate(parts):
        ks = [k for k, _ in rows]
        assert all(k % 3 == p for k in ks) and ks == sorted(ks, reverse=True)
    assert sorted(kv for rows in parts for kv in rows) == sorted(recs)


def test_sample_statistics_count_approx_and_local_checkpoint(sc):
    import statistics
    import time
    xs = [float(i * i % 17) for i in range(200)]
    ds = sc.parallelize(xs, 4)
    assert abs(ds.sample_variance() - statistics.variance(xs)) < 1e-9 and abs(ds.sample_stdev() - statistics.stdev(xs)) < 1e-9
    assert math.isnan(sc.parallelize([1.0], 1).sample_variance())
    full = ds.count_approx(timeout=30.0)
    assert full == {"mean": 200.0, "low": 200.0, "high": 200.0, "confidence": 1.0, "final": True, "partitions_done": 4}

    def slow(x):
        if x >= 150:                     # the last partition takes much longer than the timeout
            time.sleep(0.02)
        return x
    part = sc.parallelize(range(200), 4).map(slow).count_approx(timeout=0.25, confidence=0.9)
    assert part["final"] is False and 1 <= part["partitions_done"] < 4
    assert part["low"] <= 200.0 <= part["high"] * 1.5 and part["low"] >= 50 * part["partitions_done"]
    lc = sc.parallelize(range(40), 4).map(lambda x: x + 1)
    assert not lc.is_checkpointed()
    lc.local_checkpoint()
    assert lc.is_checkpointed() and lc.get_checkpoint_file() is None and lc.reduce(lambda a, b: a + b) == sum(range(1, 41))
