"""Auxiliary subsystems: checkpoint/resume, heartbeat watchdog, trace recorder."""
import json
import time

import pytest
import torch

from asyncframework_b200.core import AsyncContext, Scheduler
from asyncframework_b200.data import ShardedDataset
from asyncframework_b200.utils import HeartbeatWatchdog, TraceRecorder, load_checkpoint, save_checkpoint
from asyncframework_b200.utils.checkpoint import load_model_text, save_model_text


def test_checkpoint_roundtrip(tmp_path):
    w = torch.randn(10)
    p = save_checkpoint(str(tmp_path / "ck" / "rank0.pt"), w=w, k=123, clock=150, alpha_bar=w * 2,
                        history={0: torch.arange(5.0), 3: torch.ones(2)}, config={"seed": 42})
    ck = load_checkpoint(p)
    assert ck["k"] == 123 and ck["clock"] == 150 and torch.equal(ck["w"], w) and torch.equal(ck["alpha_bar"], w * 2)
    assert sorted(ck["history"]) == [0, 3] and ck["config"]["seed"] == 42
    assert load_checkpoint(str(tmp_path / "ck"))["k"] == 123             # directory form picks rank file
    save_model_text(str(tmp_path / "m.txt"), w, 0.5)
    w2, b = load_model_text(str(tmp_path / "m.txt"))
    assert torch.allclose(w2.float(), w) and b == 0.5


def test_watchdog_marks_silent_busy_worker_dead():
    ctx = AsyncContext(3)
    for p in range(3):
        ctx.stat.task_done(p, 0, 1.0)
    ctx.mark_submitted(1)                       # busy and silent
    dead = []
    wd = HeartbeatWatchdog(ctx.stat, timeout_s=0.05, period_s=0.01, on_dead=dead.append).start()
    t0 = time.time()
    while not dead and time.time() - t0 < 2:
        ctx.stat.heartbeat(0)
        time.sleep(0.01)
    wd.stop()
    assert dead == [1] and not ctx.stat[1].alive
    assert ctx.stat.available_workers() == 2    # the run continues with P-1


def test_trace_recorder(tmp_path):
    s = Scheduler(2)
    try:
        tr = TraceRecorder().attach(s)
        ShardedDataset.from_records(s, list(range(10)), 4).count()
        tr.dump(str(tmp_path / "trace.json"))
        d = json.load(open(tmp_path / "trace.json"))
        assert len(d["traceEvents"]) == 4 and d["counters"]["tasks"] == 4 and d["counters"]["jobs"] == 1
    finally:
        s.stop()


def test_event_log_replay(tmp_path):
    from asyncframework_b200.utils.eventlog import EventLog, replay, summarize
    s = Scheduler(2, max_failures=3)
    try:
        log = EventLog(str(tmp_path / "events.jsonl")).attach(s)
        ds = ShardedDataset.from_records(s, list(range(12)), 3)
        ds.count()
        s.fault_injector = lambda t: (_ for _ in ()).throw(RuntimeError("x")) if (t.partition_id == 2 and t.attempt == 0) else None
        ds.map(lambda x: x + 1).count()
        log.close()
        jobs = replay(str(tmp_path / "events.jsonl"))
        assert len(jobs) == 2
        a, b = [jobs[k] for k in sorted(jobs)]
        assert a["tasks"] == 3 and a["failures"] == 0 and a["finished"] and a["executors"] == [0, 1]
        assert b["tasks"] == 3 and b["failures"] == 1 and b["finished"]
        assert len(summarize(str(tmp_path / "events.jsonl"))) == 2
    finally:
        s.stop()


def test_async_checkpoint_writer_orders_and_reports_errors(tmp_path):
    from asyncframework_b200.utils.checkpoint import AsyncCheckpointWriter
    wr = AsyncCheckpointWriter(max_pending=2)
    published = []
    for k in (10, 20, 30):
        wr.submit(None, after=lambda k=k: published.append(k), path=str(tmp_path / "ck" / "rank0.pt"), w=torch.full((4,), float(k)), k=k)
    wr.flush()
    assert published == [10, 20, 30] and wr.written == 3
    ck = load_checkpoint(str(tmp_path / "ck" / "rank0.pt"))
    assert ck["k"] == 30 and float(ck["w"][0]) == 30.0
    wr.submit(None, path=str(tmp_path / "ck"), w=torch.zeros(1), k=1)          # a directory: the write fails ...
    with pytest.raises(RuntimeError):
        wr.close()                                                            # ... and the failure surfaces on close


def test_log4j_style_logging_configuration(tmp_path):
    import logging
    from asyncframework_b200.utils import logging as afb_logging
    logf = tmp_path / "driver.log"
    props = tmp_path / "log4j.properties"
    props.write_text(f"""log4j.rootCategory=WARN, file
log4j.appender.file=org.apache.log4j.FileAppender
log4j.appender.file.File={logf}
log4j.appender.file.layout.ConversionPattern=%p %c: %m%n
log4j.logger.asyncframework_b200.gd=DEBUG
""")
    base = afb_logging.configure(str(props))
    try:
        logging.getLogger("asyncframework_b200.core").info("hidden")
        logging.getLogger("asyncframework_b200.core").warning("shown")
        logging.getLogger("asyncframework_b200.gd").debug("gd-debug")
        for h in base.handlers:
            h.flush()
        text = logf.read_text()
        assert "WARNING asyncframework_b200.core: shown" in text and "hidden" not in text and "gd-debug" in text
        afb_logging.set_log_level("OFF")
        logging.getLogger("asyncframework_b200.core").error("silenced")
        assert "silenced" not in logf.read_text()
        import pytest as _pt
        with _pt.raises(ValueError):
            afb_logging.set_log_level("LOUD")
    finally:
        afb_logging.configure(props={"log4j.rootCategory": "ERROR, console"})
        logging.getLogger("asyncframework_b200.gd").setLevel(logging.NOTSET)
