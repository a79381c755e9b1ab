"""HTTP status endpoints: live application UI / REST API, history server, master status page.

Reference: the web UI (core/ui/SparkUI.scala:39, 6.3 k LoC of servlets) and the REST API under
``/api/v1/applications/[app-id]/{jobs,jobs/[id],executors,storage/rdd,environment}`` (core/status/api/v1), the
``MetricsServlet`` (``/metrics/json``), the standalone master page (``MasterWebUI`` ``/json``) and the History Server
(core/deploy/history/FsHistoryProvider.scala:80 — lists event logs in a directory and replays them).

One small threaded HTTP server serves a route table of ``path -> callable() -> JSON-able``; :class:`StatusServer`
binds the routes of a live :class:`~asyncframework_b200.utils.status.AppStatusStore` (+ Prometheus text exposition of
a :class:`MetricsRegistry`), :class:`HistoryServer` those of an event-log directory.  ``/`` renders a plain HTML
overview (jobs, executors, storage, ASYNC worker table) built from the same JSON.
"""
from __future__ import annotations

import html
import json
import os
import re
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Callable, Dict, List, Optional, Tuple

Route = Callable[..., Any]


class _Handler(BaseHTTPRequestHandler):
    routes: List[Tuple[re.Pattern, Route]] = []
    server_version = "afb-status/1.0"

    def log_message(self, fmt, *args):  # silence the default stderr access log
        pass

    def do_GET(self):  # noqa: N802
        path, _, query = self.path.partition("?")
        params = dict(kv.split("=", 1) for kv in query.split("&") if "=" in kv)
        for pat, fn in self.routes:
            m = pat.fullmatch(path)
            if m is None:
                continue
            try:
                body = fn(*m.groups(), **params) if params else fn(*m.groups())
            except KeyError as exc:
                self._send(404, "application/json", json.dumps({"error": f"not found: {exc}"}))
                return
            except Exception as exc:  # noqa: BLE001
                self._send(500, "application/json", json.dumps({"error": repr(exc)}))
                return
            if isinstance(body, tuple):                      # (content type, text)
                self._send(200, body[0], body[1])
            else:
                self._send(200, "application/json", json.dumps(body, default=_jsonable, indent=1))
            return
        self._send(404, "application/json", json.dumps({"error": f"no route for {path}", "routes": [p.pattern for p, _ in self.routes]}))

    def _send(self, code: int, ctype: str, text: str) -> None:
        data = text.encode()
        self.send_response(code)
        self.send_header("Content-Type", ctype + "; charset=utf-8")
        self.send_header("Content-Length", str(len(data)))
        self.end_headers()
        self.wfile.write(data)


def _jsonable(o):
    if isinstance(o, (set, frozenset)):
        return sorted(o, key=str)
    if hasattr(o, "_asdict"):
        return o._asdict()
    if hasattr(o, "tolist"):
        return o.tolist()
    return repr(o)


class JsonServer:
    def __init__(self, host: str, port: int, routes: Dict[str, Route]):
        handler = type("Handler", (_Handler,), {"routes": [(re.compile(p), fn) for p, fn in routes.items()]})
        self.httpd = ThreadingHTTPServer((host, port), handler)
        self.httpd.daemon_threads = True
        self.host, self.port = host, self.httpd.server_address[1]
        self.thread = threading.Thread(target=self.httpd.serve_forever, daemon=True, name=f"status-http-{self.port}")
        self.thread.start()

    @property
    def url(self) -> str:
        return f"http://{self.host}:{self.port}"

    def stop(self) -> None:
        self.httpd.shutdown()
        self.httpd.server_close()


def serve_json(host: str, port: int, routes: Dict[str, Route]) -> JsonServer:
    return JsonServer(host, port, routes)


def _table(title: str, rows: List[dict], cols: Optional[List[str]] = None) -> str:
    if not rows:
        return f"<h3>{html.escape(title)}</h3><p><i>none</i></p>"
    cols = cols or list(rows[0].keys())
    head = "".join(f"<th>{html.escape(c)}</th>" for c in cols)
    body = "".join("<tr>" + "".join(f"<td>{html.escape(str(r.get(c, '')))}</td>" for c in cols) + "</tr>" for r in rows)
    return f"<h3>{html.escape(title)}</h3><table border=1 cellpadding=3 style='border-collapse:collapse'><tr>{head}</tr>{body}</table>"


class StatusServer:
    """``SparkUI`` + REST API of one running application."""

    def __init__(self, store, registry=None, host: str = "127.0.0.1", port: int = 0):
        self.store, self.registry = store, registry
        app = re.escape(store.app_id)
        base = "/api/v1/applications"
        routes: Dict[str, Route] = {
            "/": self._index,
            "/jobs/(\\d+)": self._job_page,
            "/executors": self._executors_page,
            base: lambda: [store.application_info()],
            f"{base}/{app}": store.application_info,
            f"{base}/{app}/jobs": lambda status=None: store.job_list(status),
            f"{base}/{app}/jobs/(\\d+)": self._job,
            f"{base}/{app}/executors": store.executor_list,
            f"{base}/{app}/allexecutors": store.executor_list,
            f"{base}/{app}/executors/([^/]+)/threads": self._threads,
            f"{base}/{app}/storage/rdd": store.rdd_storage,
            f"{base}/{app}/environment": lambda: store.environment,
            f"{base}/{app}/async/workers": store.async_workers,
            f"{base}/{app}/summary": store.summary,
            "/metrics/json": lambda: (registry.snapshot() if registry is not None else {}),
            "/metrics/prometheus": self._prometheus,
        }
        self.server = JsonServer(host, port, routes)

    @property
    def url(self) -> str:
        return self.server.url

    def _job(self, job_id: str):
        j = self.store.job(int(job_id))
        if j is None:
            raise KeyError(f"job {job_id}")
        return j

    def _threads(self, executor_id: str):
        sched = self.store._scheduler
        if hasattr(sched, "thread_dump"):
            return sched.thread_dump(executor_id)
        if executor_id != "driver":
            raise KeyError(f"executor {executor_id} (executors of an in-process master are threads of the driver: ask for 'driver')")
        from ..core.cluster import thread_dump
        return thread_dump()

    def _prometheus(self):
        from .metrics import prometheus_text
        return ("text/plain; version=0.0.4", prometheus_text(self.registry.snapshot() if self.registry is not None else {}))

    @staticmethod
    def _page(title: str, body: str, refresh: int = 5) -> tuple:
        nav = "<p><a href='/'>overview</a> &middot; <a href='/executors'>executors</a> &middot; <a href='/metrics/json'>metrics</a></p>"
        return ("text/html", f"<html><head><title>{html.escape(title)}</title><meta http-equiv='refresh' content='{refresh}'></head>"
                             f"<body><h2>{html.escape(title)}</h2>{nav}{body}</body></html>")

    def _job_page(self, job_id: str):
        j = self._job(job_id)
        tasks = [{"partition": t.get("partition"), "status": t.get("status"), "executor": t.get("executor"), "attempt": t.get("attempt"),
                  "run_ms": round(t.get("run_ms", 0.0), 3) if "run_ms" in t else "", **{k: round(v, 3) if isinstance(v, float) else v
                                                                                      for k, v in (t.get("metrics") or {}).items()},
                  "error": t.get("error", "")} for t in j.pop("tasks", [])]
        summary = [{k: v for k, v in j.items() if k != "task_run_ms"}]
        return self._page(f"Job {job_id}", _table("Summary", summary) + _table("Task run time (ms)", [j.get("task_run_ms") or {}])
                          + _table("Tasks", tasks))

    def _executors_page(self):
        rows = self.store.executor_list()
        links = [{**r, "threads": f"/api/v1/applications/{self.store.app_id}/executors/{r['id']}/threads"} for r in rows]
        return self._page("Executors", _table("Executors", links) + _table("Cached datasets", self.store.rdd_storage()))

    def _index(self):
        s = self.store
        info = s.summary()
        jobs = [{**{k: j.get(k) for k in ("job_id", "status", "num_tasks", "num_completed_tasks", "num_failed_tasks", "num_killed_tasks",
                                          "disk_bytes_spilled", "duration_ms")},
                 "details": f"/jobs/{j['job_id']}"} for j in s.job_list()[:50]]
        page = (f"<html><head><title>{html.escape(s.app_name)}</title></head><body><h2>{html.escape(s.app_name)} "
                f"<small>({html.escape(s.app_id)})</small></h2><p>{html.escape(json.dumps(info['jobs_by_status']))} &middot; "
                f"{info['completed_tasks']} tasks &middot; {info['active_executors']}/{info['executors']} executors</p>"
                + _table("Jobs (latest 50)", jobs) + _table("Executors", s.executor_list())
                + _table("Cached datasets", s.rdd_storage()) + _table("ASYNC workers", s.async_workers())
                + _table("Executor allocation (dynamic)", info.get("executor_allocation_events") or [])
                + "<p>REST: <code>/api/v1/applications</code>, metrics: <code>/metrics/json</code>, <code>/metrics/prometheus</code></p></body></html>")
        return ("text/html", page)

    def stop(self) -> None:
        self.server.stop()


class HistoryServer:
    """``FsHistoryProvider``: every ``*.jsonl`` event log under ``log_dir`` is one application; its per-job summary is
    rebuilt by replaying the log (utils/eventlog.py::replay) on demand and cached by (size, mtime)."""

    def __init__(self, log_dir: str, host: str = "127.0.0.1", port: int = 0):
        self.log_dir = log_dir
        self._cache: Dict[str, tuple] = {}
        self.server = JsonServer(host, port, {
            "/": self._index,
            "/api/v1/applications": self.applications,
            "/api/v1/applications/([^/]+)/jobs": self.jobs,
            "/api/v1/applications/([^/]+)/jobs/(-?\\d+)": self.job,
        })

    @property
    def url(self) -> str:
        return self.server.url

    def _logs(self) -> Dict[str, str]:
        out = {}
        for f in sorted(os.listdir(self.log_dir)) if os.path.isdir(self.log_dir) else []:
            if f.endswith((".jsonl", ".eventlog")):
                out[os.path.splitext(f)[0]] = os.path.join(self.log_dir, f)
        return out

    def _replayed(self, app: str) -> Dict[int, dict]:
        from .eventlog import replay
        path = self._logs()[app]                       # KeyError -> 404
        st = os.stat(path)
        key = (st.st_size, st.st_mtime_ns)
        hit = self._cache.get(app)
        if hit is None or hit[0] != key:
            hit = (key, replay(path))
            self._cache[app] = hit
        return hit[1]

    def applications(self) -> List[dict]:
        out = []
        for app, path in self._logs().items():
            jobs = self._replayed(app)
            st = os.stat(path)
            out.append({"id": app, "log": path, "log_bytes": st.st_size, "last_updated": st.st_mtime, "jobs": len(jobs),
                        "completed": all(j["finished"] for j in jobs.values()) if jobs else False,
                        "tasks": sum(j["tasks"] for j in jobs.values()), "failed_attempts": sum(j["failures"] for j in jobs.values())})
        return out

    def jobs(self, app: str) -> List[dict]:
        return [{"job_id": jid, **j} for jid, j in sorted(self._replayed(app).items())]

    def job(self, app: str, job_id: str) -> dict:
        return {"job_id": int(job_id), **self._replayed(app)[int(job_id)]}

    def _index(self):
        return ("text/html", "<html><body><h2>History</h2>" + _table("Applications", self.applications()) + "</body></html>")

    def stop(self) -> None:
        self.server.stop()


def main(argv=None) -> None:
    import argparse
    import time
    ap = argparse.ArgumentParser(description="history server over a directory of event logs")
    ap.add_argument("log_dir")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=18080)
    a = ap.parse_args(argv)
    hs = HistoryServer(a.log_dir, a.host, a.port)
    print(f"History server at {hs.url}", flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        hs.stop()


if __name__ == "__main__":
    main()
