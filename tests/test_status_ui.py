"""Status store, REST/HTML UI, history server, metric sinks (AppStatusListenerSuite / UISeleniumSuite /
HistoryServerSuite / MetricsSystemSuite scenarios of the reference, behaviour level)."""
import json
import socket
import socketserver
import threading
import time
import urllib.error
import urllib.request

import pytest

from asyncframework_b200 import AsyncSession
from asyncframework_b200.utils.metrics import (MetricsRegistry, MetricsSystem, PrometheusFileSink, StatsdSink, GraphiteSink,
                                                prometheus_text)
from asyncframework_b200.utils.status_server import HistoryServer


def _get(url):
    with urllib.request.urlopen(url, timeout=5) as r:
        body = r.read().decode()
        return json.loads(body) if "json" in r.headers["Content-Type"] else body


def test_live_ui_rest_api_over_local_scheduler(tmp_path):
    log = str(tmp_path / "app-1.jsonl")
    with AsyncSession("local[2]", app_name="ui-test") as sc:
        ui = sc.start_ui(event_log=log)
        app_id = sc.status_store.app_id
        base = f"{ui.url}/api/v1/applications"
        ds = sc.parallelize(range(40), 4)
        assert ds.map(lambda x: x + 1).reduce(lambda a, b: a + b) == 820
        with pytest.raises(Exception):
            ds.map(lambda x: 1 // 0).collect()
        ctx = sc.async_context(4)
        ds.async_reduce(lambda a, b: a + b, ctx)
        for _ in range(4):
            ctx.collect()
        apps = _get(base)
        assert apps[0]["id"] == app_id and apps[0]["name"] == "ui-test"
        jobs = _get(f"{base}/{app_id}/jobs")
        assert len(jobs) == 3 and {j["status"] for j in jobs} == {"SUCCEEDED", "FAILED"}
        failed = _get(f"{base}/{app_id}/jobs?status=failed")
        assert len(failed) == 1 and failed[0]["num_failed_tasks"] >= 1 and "ZeroDivisionError" in failed[0]["error"]
        ok = [j for j in jobs if j["status"] == "SUCCEEDED"][0]
        detail = _get(f"{base}/{app_id}/jobs/{ok['job_id']}")
        assert detail["num_completed_tasks"] == 4 and len(detail["tasks"]) == 4 and detail["task_run_ms"]["max"] >= 0
        ex = _get(f"{base}/{app_id}/executors")
        assert sum(e["completed_tasks"] for e in ex) == 8 and all(e["active"] for e in ex)
        workers = _get(f"{base}/{app_id}/async/workers")
        assert len(workers) == 4 and all(w["available"] and w["clock"] == 4 for w in workers)
        env = _get(f"{base}/{app_id}/environment")
        assert env["master"] == "local[2]" and env["conf"]["batch_rate"] == 0.1
        m = _get(f"{ui.url}/metrics/json")
        assert m["scheduler.jobs"] == 3 and m["scheduler.tasks"] == 8 and m["scheduler.task_failures"] >= 1
        prom = _get(f"{ui.url}/metrics/prometheus")
        assert "# TYPE afb_scheduler_jobs gauge" in prom and "afb_scheduler_jobs 3" in prom
        page = _get(ui.url + "/")
        assert "<table" in page and "ui-test" in page and "ASYNC workers" in page
        jp = _get(f"{ui.url}/jobs/{ok['job_id']}")
        assert "Tasks" in jp and "SUCCESS" in jp and "http-equiv='refresh'" in jp
        assert "Executors" in _get(ui.url + "/executors")
        with pytest.raises(urllib.error.HTTPError) as e404:
            _get(f"{base}/{app_id}/jobs/999999")
        assert e404.value.code == 404
    # the event log written next to the live UI is what the history server replays
    hs = HistoryServer(str(tmp_path))
    try:
        apps = _get(f"{hs.url}/api/v1/applications")
        assert apps[0]["id"] == "app-1" and apps[0]["jobs"] == 3 and apps[0]["tasks"] == 8 and apps[0]["failed_attempts"] >= 1
        jobs = _get(f"{hs.url}/api/v1/applications/app-1/jobs")
        assert sum(1 for j in jobs if j.get("failed")) == 1
        assert "Applications" in _get(hs.url + "/")
    finally:
        hs.stop()


def test_ui_reports_block_manager_storage_for_cluster_masters():
    with AsyncSession("local-cluster[2]", app_name="ui-cluster") as sc:
        ui = sc.start_ui()
        base = f"{ui.url}/api/v1/applications/{sc.status_store.app_id}"
        ds = sc.parallelize(range(40), 4).map(lambda x: x * 2).cache()
        assert ds.count() == 40
        ex = _get(f"{base}/executors")
        assert ex[0]["id"] == "driver" and {str(e["id"]) for e in ex} == {"driver", "0", "1"}
        assert all(e["rdd_blocks"] == 2 and e["memory_used"] > 0 for e in ex if e["id"] != "driver")
        rdd = _get(f"{base}/storage/rdd")
        assert rdd[0]["num_cached_partitions"] == 4 and rdd[0]["executors"] == ["0", "1"]
        job = _get(f"{base}/jobs")[0]
        assert job["task_binary_bytes"] > 0 and job["result_bytes"] > 0
        dump = _get(f"{base}/executors/1/threads")
        assert any("run_tasks" in t["stack"] for t in dump) and any("heartbeat" in t["name"] for t in dump)
        assert any("MainThread" == t["name"] for t in _get(f"{base}/executors/driver/threads"))
        t = _get(f"{base}/jobs/{job['job_id']}")["tasks"][0]
        assert {"serialize_ms", "deserialize_ms", "result_bytes"} <= set(t["metrics"])      # TaskMetrics


def test_prometheus_statsd_graphite_sinks_and_metrics_properties(tmp_path):
    snap = {"scheduler.jobs": 3.0, "a-b/c": 1.5}
    text = prometheus_text(snap)
    assert "afb_a_b_c 1.5" in text and text.endswith("\n")

    udp = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)
    udp.bind(("127.0.0.1", 0))
    udp.settimeout(3)
    got_tcp = []

    class H(socketserver.StreamRequestHandler):
        def handle(self):
            got_tcp.append(self.rfile.read().decode())
    tcp = socketserver.TCPServer(("127.0.0.1", 0), H)
    threading.Thread(target=tcp.serve_forever, daemon=True).start()

    props = tmp_path / "metrics.properties"
    props.write_text(f"""# like conf/metrics.properties.template
*.sink.statsd.class=org.apache.spark.metrics.sink.StatsdSink
*.sink.statsd.host=127.0.0.1
*.sink.statsd.port={udp.getsockname()[1]}
*.sink.statsd.period=0.05
*.sink.graphite.class=GraphiteSink
*.sink.graphite.host=127.0.0.1
*.sink.graphite.port={tcp.server_address[1]}
*.sink.prom.class=prometheus
*.sink.prom.path={tmp_path / 'metrics.prom'}
*.sink.csv.class=CsvSink
*.sink.csv.directory={tmp_path / 'csv'}
""")
    ms = MetricsSystem.from_properties(MetricsSystem.load_properties(str(props)))
    assert ms.period == 0.05 and len(ms.registry.sinks) == 4
    ms.registry.inc("jobs", 2)
    ms.registry.observe("task_ms", 4.0)
    ms.start()
    datagram = udp.recv(4096).decode()
    assert "afb.jobs:2|g" in datagram and "afb.task_ms.mean:4|g" in datagram
    ms.stop()
    t0 = time.time()
    while not got_tcp and time.time() - t0 < 3:
        time.sleep(0.02)
    assert any(line.startswith("afb.jobs 2 ") for line in got_tcp[0].splitlines())
    assert "afb_jobs 2" in (tmp_path / "metrics.prom").read_text()
    assert (tmp_path / "csv" / "jobs.csv").read_text().splitlines()[0] == "t,value"
    tcp.shutdown()
    tcp.server_close()
    udp.close()
    g = GraphiteSink("127.0.0.1", 1)                       # nothing listens: the sink must swallow the error
    g.write(time.time(), {"x": 1.0})
    assert g.errors == 1
    with pytest.raises(ValueError, match="unknown metrics sink"):
        MetricsSystem.from_properties({"*.sink.x.class": "NoSuchSink"})


@pytest.mark.parametrize("master", ["local[2]", "local-cluster[2]"])
def test_task_profiler_collects_cprofile_stats_per_job(master, tmp_path):
    """ProfilerSuite (python/pyspark/tests.py): stats of every task come back through an accumulator, per job."""
    import pstats
    from asyncframework_b200.conf import AsyncConf

    def busy(x):
        return sum(i * i for i in range(2000)) + x
    with AsyncSession(master, conf=AsyncConf(load_defaults=False).set("spark.python.profile", "true")) as sc:
        ds = sc.parallelize(range(40), 4)
        assert ds.map(busy).count() == 40
        assert ds.map(lambda x: x).reduce(lambda a, b: a + b) == 780
        jobs = sc.profiler.jobs()
        assert len(jobs) == 2
        st = sc.profiler.stats(jobs[0])
        names = {f[2] for f in st.stats}
        assert "busy" in names
        calls = sum(v[1] for f, v in st.stats.items() if f[2] == "busy")
        assert calls == 40                                                  # 4 tasks merged, each record exactly once
        text = sc.show_profiles(top=5)
        assert f"Profile of job {jobs[0]}" in text and "cumulative" in text
        sc.dump_profiles(str(tmp_path))
        loaded = pstats.Stats(str(tmp_path / f"job_{jobs[0]}.pstats"))
        assert any(f[2] == "busy" for f in loaded.stats)


def test_status_store_aggregates_spills_killed_attempts_and_allocation_events():
    """AppStatusListener counterparts for the round-2 scheduler events: spill bytes of a job's tasks, `task_killed` (the losing
    copy of a speculated task) and `executors_resized` (dynamic allocation) show up in the job view and the summary."""
    from asyncframework_b200.utils.status import AppStatusStore
    st = AppStatusStore("t")
    st({"event": "job_start", "job": 1, "partitions": [0, 1]})
    for p, spill in ((0, 1000), (1, 0)):
        st({"event": "task_start", "job": 1, "partition": p, "attempt": 0, "executor": p})
        st({"event": "task_end", "job": 1, "partition": p, "executor": p, "run_ms": 3.0,
            "metrics": {"result_bytes": 10, "memory_bytes_spilled": 4 * spill, "disk_bytes_spilled": spill, "spills": int(spill > 0)}})
    st({"event": "task_killed", "job": 1, "partition": 1, "attempt": 0, "executor": 1, "reason": "interrupted", "ran_ms": 12.0})
    st({"event": "job_end", "job": 1})
    st({"event": "executors_resized", "from": 1, "to": 4, "reason": "dynamic_allocation"})
    j = st.job(1)
    assert j["status"] == "SUCCEEDED" and j["memory_bytes_spilled"] == 4000 and j["disk_bytes_spilled"] == 1000 and j["num_killed_tasks"] == 1
    s = st.summary()
    assert s["killed_task_attempts"] == 1 and s["disk_bytes_spilled"] == 1000
    assert [(e["from"], e["to"]) for e in s["executor_allocation_events"]] == [(1, 4)]
