"""``AsyncSession`` — the SparkContext-shaped entry point.

Bundles what ``new SparkContext(conf)`` sets up for the ASYNC samples (core/SparkContext.scala:
432-571): the scheduler ("local[N]" executors), broadcast factory, dataset constructors
(``parallelize``, ``text_file``, ``load_libsvm_file``), the blocking/non-blocking job mode switch
(``set_mode``, SparkContext.scala:87-104) and log level.  Master URL grammar follows
SparkContext.scala:2805-2811: ``local``, ``local[N]``, ``local[*]``, ``local[N,F]`` (F = max task
failures), ``local-cluster[n,c,m]`` (executor processes), ``spark://host:port`` (standalone master); ``fused`` selects the multi-GPU fused backend (one process per GPU, see FusedEngine).
"""
from __future__ import annotations

import os
import re
from typing import Any, Iterable, Optional, Sequence

import torch

from ..config import AsyncConfig
from .broadcast import AsyncBroadcast, Broadcast
from .context import AsyncContext
from .scheduler import Scheduler

_LOCAL_N = re.compile(r"^local\[(\*|\d+)(?:\s*,\s*(\d+))?\]$")
_LOCAL_CLUSTER = re.compile(r"^local-cluster\[\s*(\d+)\s*(?:,\s*(\d+)\s*(?:,\s*(\d+)\s*)?)?\]$")


_added_files: dict = {}            # file name -> Broadcast of its bytes (sc.addFile)


class SparkFiles:
    """``SparkFiles.get(name)`` inside a task: local path of a file shipped with ``sc.add_file``."""
    _dir: Optional[str] = None
    _written: set = set()
    _lock = __import__("threading").Lock()
    registry = _added_files           # pickled BY VALUE inside closures that call SparkFiles.get through `files=`

    @classmethod
    def get_root_directory(cls) -> str:
        import tempfile
        if cls._dir is None:
            cls._dir = tempfile.mkdtemp(prefix="afb-files-")
            import atexit
            import shutil
            atexit.register(shutil.rmtree, cls._dir, True)
        return cls._dir

    @classmethod
    def get(cls, name: str, files: Optional[dict] = None) -> str:
        reg = files if files is not None else cls.registry
        if name not in reg:
            raise FileNotFoundError(f"{name} was not added with sc.add_file (in a cluster closure pass files=sc.files)")
        path = os.path.join(cls.get_root_directory(), name)
        with cls._lock:                                # executor THREADS of an in-process master share this directory
            if name not in cls._written or not os.path.exists(path):
                tmp = f"{path}.{os.getpid()}.tmp"
                with open(tmp, "wb") as f:
                    f.write(reg[name].value)
                os.replace(tmp, path)                   # readers never see a half-written file
                cls._written.add(name)
        return path


class _StatusTracker:
    """``SparkStatusTracker`` (core/SparkStatusTracker.scala): ``get_active_job_ids``, ``get_job_ids_for_group``, ``get_job_info``."""

    def __init__(self, sc):
        self.sc = sc

    def get_active_job_ids(self):
        return self.sc.scheduler.active_job_ids()

    def get_job_ids_for_group(self, group_id):
        return self.sc.scheduler.active_job_ids(group_id)

    def get_job_info(self, job_id: int):
        """Needs ``sc.start_ui()`` / ``spark.ui.enabled`` (the status store records the listener events); ``None`` when unknown."""
        store = getattr(self.sc, "status_store", None)
        return store.job(job_id) if store is not None else None

    def get_executor_infos(self):
        store = getattr(self.sc, "status_store", None)
        return store.executor_list() if store is not None else [{"id": e} for e in range(self.sc.scheduler.num_executors)]


class AsyncSession:
    def __init__(self, master: Optional[str] = None, app_name: Optional[str] = None, device=None,
                 conf=None, num_executors: Optional[int] = None, cores_per_executor: Optional[int] = None,
                 executor_memory_mb: Optional[int] = None, runtime_conf=None):
        """``conf`` may be the algorithm-level :class:`AsyncConfig` or the runtime :class:`AsyncConf` (``new
        SparkContext(conf)``); explicit arguments win over ``spark.master`` / ``spark.app.name`` / ``spark.executor.*``."""
        from ..conf import AsyncConf
        if isinstance(conf, AsyncConf):
            runtime_conf, conf = conf, None
        rc = runtime_conf if runtime_conf is not None else AsyncConf()
        self.runtime_conf = rc.validate()
        master = master or rc.get("spark.master")
        app_name = app_name or rc.get("spark.app.name")
        num_executors = num_executors or (rc.get("spark.executor.instances") or None)
        cores_per_executor = cores_per_executor or rc.get("spark.executor.cores")
        executor_memory_mb = executor_memory_mb or (max(1, rc.get("spark.executor.memory") >> 20) if rc.contains("spark.executor.memory") else 512)
        self.master = master
        self.app_name = app_name
        self.conf = conf or AsyncConfig()
        n, max_fail = self._parse_master(master)
        if rc.contains("spark.task.maxFailures"):
            max_fail = rc.get("spark.task.maxFailures")
        self.device = torch.device(device) if device is not None else None
        cluster_kw = dict(max_failures=max_fail, name=app_name, limits=rc.runtime_limits(),
                          block_first_async_job=rc.get("spark.async.blockFirstJob"),
                          speculation=rc.get("spark.speculation"), speculation_multiplier=rc.get("spark.speculation.multiplier"),
                          speculation_quantile=rc.get("spark.speculation.quantile"), speculation_interval=rc.get("spark.speculation.interval"),
                          blacklist=rc.get("spark.blacklist.enabled"),
                          blacklist_max_failures=rc.get("spark.blacklist.application.maxFailedTasksPerExecutor"),
                          blacklist_timeout=rc.get("spark.blacklist.timeout"),
                          max_task_binary_bytes=rc.get("spark.rpc.message.maxSize"))
        if master.startswith("local-cluster"):
            from .cluster import ClusterScheduler
            m = _LOCAL_CLUSTER.match(master)
            if m and m.group(2):                          # local-cluster[n, coresPerExecutor, memoryMB]
                cluster_kw["cores_per_executor"] = int(m.group(2))
            self.scheduler = ClusterScheduler(n, recover_lost_executors=rc.contains("spark.async.recoverLostExecutors")
                                              and rc.get("spark.async.recoverLostExecutors"), **cluster_kw)
        elif master.startswith(("spark://", "async://")):
            # standalone cluster manager (deploy/master.py, deploy/worker.py): executors are forked by Worker daemons on
            # any host and talk to this driver over TCP; a lost executor is replaced by the Master and its tasks re-run
            from ..deploy.client import StandaloneTransport
            from .cluster import ClusterScheduler
            transport = StandaloneTransport(master, num_executors, cores_per_executor, executor_memory_mb, app_name,
                                            register_timeout=max(30.0, rc.get("spark.network.timeout") / 2))
            self.scheduler = ClusterScheduler(num_executors or 1, transport=transport,
                                              recover_lost_executors=rc.get("spark.async.recoverLostExecutors"), **cluster_kw)
        else:
            self.scheduler = Scheduler(n, max_failures=max_fail, device=self.device, name=app_name,
                                       block_first_async_job=rc.get("spark.async.blockFirstJob"))
        self._stopped = False
        if rc.get("spark.dynamicAllocation.enabled"):
            # ExecutorAllocationManager (SparkContext.scala:552-566 starts it when the cluster manager can resize the app)
            from .allocation import AllocationPolicy
            if not hasattr(getattr(self.scheduler, "transport", None), "request_total_executors"):
                raise ValueError(f"spark.dynamicAllocation.enabled needs a cluster manager (spark://...), not master {master!r}")
            self.scheduler.enable_dynamic_allocation(AllocationPolicy.from_conf(rc))
        if rc.get("spark.ui.enabled") or rc.get("spark.eventLog.enabled"):
            log = None
            if rc.get("spark.eventLog.enabled"):
                os.makedirs(rc.get("spark.eventLog.dir"), exist_ok=True)
                import time as _t
                log = os.path.join(rc.get("spark.eventLog.dir"), f"{app_name}-{int(_t.time() * 1000)}.jsonl")
            self.start_ui(port=rc.get("spark.ui.port") if rc.get("spark.ui.enabled") else 0, event_log=log)
        if rc.get("spark.python.profile"):
            self.enable_profiling()
        if rc.get("spark.metrics.conf"):
            from ..utils.metrics import MetricsSystem, SchedulerSource
            self.metrics_system = MetricsSystem.from_properties(MetricsSystem.load_properties(rc.get("spark.metrics.conf")),
                                                                getattr(self, "metrics", None))
            if getattr(self, "metrics", None) is None:
                self.metrics = self.metrics_system.registry
                SchedulerSource(self.metrics, self.scheduler)
            self.metrics_system.start()

    @staticmethod
    def _parse_master(master: str):
        if master == "local":
            return 1, 1
        m = _LOCAL_N.match(master)
        if m:
            n = (os.cpu_count() or 1) if m.group(1) == "*" else int(m.group(1))
            return max(1, n), int(m.group(2)) if m.group(2) else 4
        m = _LOCAL_CLUSTER.match(master)
        if m:                                  # local-cluster[numExecutors, coresPerExecutor, memoryMB]
            return max(1, int(m.group(1))), 4
        if master in ("fused", "gloo"):
            return 1, 4
        if re.match(r"^(spark|async)://[^:/,]+:\d+(,[^:/,]+:\d+)*$", master):
            return 1, 4
        raise ValueError(f"Could not parse Master URL: '{master}'")

    # ---- datasets -----------------------------------------------------------------------
    @property
    def default_parallelism(self) -> int:
        """``spark.default.parallelism`` when set, else the number of executors (``sc.defaultParallelism``)."""
        n = self.runtime_conf.get("spark.default.parallelism")
        return n if n and n > 0 else self.scheduler.num_executors

    def parallelize(self, data: Iterable[Any], num_slices: Optional[int] = None):
        from ..data.dataset import ShardedDataset
        return ShardedDataset.from_records(self.scheduler, list(data), num_slices or self.default_parallelism)

    def range(self, start: int, end: Optional[int] = None, step: int = 1, num_slices: Optional[int] = None):
        """``sc.range``."""
        if end is None:
            start, end = 0, start
        return self.parallelize(range(start, end, step), num_slices)

    def empty_dataset(self):
        """``sc.emptyRDD``."""
        return self.parallelize([], 1)

    def union(self, datasets):
        """``sc.union(rdds)``."""
        out = datasets[0]
        for d in datasets[1:]:
            out = out.union(d)
        return out

    def whole_text_files(self, path: str, min_partitions: Optional[int] = None):
        """``sc.wholeTextFiles``: one ``(file path, content)`` record per file of a directory."""
        from ..data.libsvm import _input_files
        recs = []
        for f in _input_files(path):
            with open(f) as fh:
                recs.append((f, fh.read()))
        return self.parallelize(recs, min_partitions or min(self.default_parallelism, max(1, len(recs))))

    def pickle_file(self, directory: str):
        """``sc.pickleFile``: read back what ``ShardedDataset.save_as_pickle_file`` wrote (one partition per part file)."""
        import pickle as _pickle
        from ..data.dataset import ShardedDataset
        parts = sorted(f for f in os.listdir(directory) if f.startswith("part-"))

        def compute(p, t, parts=parts, directory=directory):
            with open(os.path.join(directory, parts[p]), "rb") as fh:
                return iter(_pickle.load(fh))
        return ShardedDataset(self.scheduler, max(1, len(parts)), compute if parts else (lambda p, t: iter(())), None, None, "pickle_file")

    def text_file(self, path: str, min_partitions: Optional[int] = None):
        """Line records of a file or a directory of part files (``sc.textFile``, SparkContext.scala:837)."""
        from ..data.libsvm import _input_files
        lines = []
        for f in _input_files(path):
            with open(f) as fh:
                lines.extend(l.rstrip("\n") for l in fh)
        return self.parallelize(lines, min_partitions or self.default_parallelism)

    def load_libsvm_file(self, path: str, num_features: int = -1, min_partitions: Optional[int] = None, **kw):
        from ..data.libsvm import load_libsvm_file
        return load_libsvm_file(self.scheduler, path, num_features, min_partitions or self.default_parallelism, **kw)

    # ---- broadcast / async context -----------------------------------------------------------
    def broadcast(self, value: Any) -> Broadcast:
        """In-process masters share the snapshot by reference (like Spark's local mode); cluster masters write a
        piece-wise :class:`TorrentBroadcast` into the driver's block manager and ship only the handle."""
        if getattr(self.scheduler, "env", None) is not None:
            from .torrent_broadcast import TorrentBroadcast
            return TorrentBroadcast(value, self.scheduler.env)
        return Broadcast(value)

    get_broadcast_with_id = broadcast     # ``BroadcastManager.getBroadcastwithID`` of the fork: an (unused) alias of newBroadcast

    # ``sc.numFinishedPar`` of the fork (SparkContext.scala:90-99): a plain counter the samples may use
    num_finished_par = 0

    def set_num_finished_par(self, i: int) -> None:
        self.num_finished_par = int(i)

    def get_num_finished_par(self) -> int:
        return self.num_finished_par

    def long_accumulator(self, name: Optional[str] = None):
        """``sc.longAccumulator``: tasks ``add``, the driver reads ``.value`` (updates of successful tasks only, exactly once)."""
        from .accumulator import NumericAccumulator
        return NumericAccumulator(0, name)

    def double_accumulator(self, name: Optional[str] = None):
        from .accumulator import NumericAccumulator
        return NumericAccumulator(0.0, name)

    def collection_accumulator(self, name: Optional[str] = None):
        """``sc.collectionAccumulator``: tasks append elements, the driver reads the list."""
        from .accumulator import CollectionAccumulator
        return CollectionAccumulator(name)

    def accumulator(self, zero, name: Optional[str] = None, op=None):
        """General accumulator: ``op`` must be associative and commutative (default ``+``)."""
        from .accumulator import Accumulator
        return Accumulator(zero, name, op)

    def async_context(self, num_workers: Optional[int] = None) -> AsyncContext:
        ctx = AsyncContext(num_workers or self.default_parallelism, self.scheduler)
        if getattr(self, "status_store", None) is not None:
            self.status_store.watch_context(ctx)
        return ctx

    def async_broadcast(self, ctx: AsyncContext, value: Any, slots: int = 8) -> AsyncBroadcast:
        return AsyncBroadcast(ctx, value, slots)

    # ---- job groups, files, identity ------------------------------------------------------------------------
    def set_job_group(self, group_id: Optional[str], description: str = "") -> None:
        """``sc.setJobGroup``: jobs submitted by THIS thread from now on belong to ``group_id`` (``None`` clears it)."""
        from .scheduler import set_job_group
        set_job_group(group_id)

    def clear_job_group(self) -> None:
        self.set_job_group(None)

    def cancel_job_group(self, group_id: str) -> int:
        return self.scheduler.cancel_job_group(group_id)

    def add_file(self, path: str) -> None:
        """``sc.addFile``: make a driver-side file available to every task as ``SparkFiles.get(name)`` (the bytes travel as a
        broadcast — torrent on cluster masters — and are written once per executor process)."""
        with open(path, "rb") as f:
            data = f.read()
        name = os.path.basename(path)
        _added_files[name] = self.broadcast(data)

    def get_executor_memory_status(self) -> dict:
        """``sc.getExecutorMemoryStatus``: ``{"executor-id host:port": (max storage memory, remaining)}`` on cluster masters."""
        st = self.scheduler.storage_status() if hasattr(self.scheduler, "storage_status") else {}
        return {f"{m.executor_id} {m.host}:{m.port}": (s["max_memory"], s["max_memory"] - s["mem_used"]) for m, s in st.items()}

    @property
    def status_tracker(self):
        """``sc.statusTracker``: low-level job status without the UI."""
        return _StatusTracker(self)

    @property
    def files(self) -> dict:
        """The added files (name -> broadcast); capture it in closures that run on executor processes:
        ``lambda x, files=sc.files: open(SparkFiles.get("lookup.txt", files))``."""
        return dict(_added_files)

    @property
    def application_id(self) -> str:
        t = getattr(self.scheduler, "transport", None)
        return getattr(t, "app_id", None) or f"local-{id(self):x}"

    @property
    def ui_web_url(self) -> Optional[str]:
        return self.ui.url if getattr(self, "ui", None) is not None else None

    @property
    def version(self) -> str:
        from .. import __version__
        return __version__

    # ---- job mode / misc ---------------------------------------------------------------------
    def set_mode(self, mode: int) -> None:
        self.scheduler.set_mode(mode)

    @property
    def mode(self) -> int:
        return self.scheduler.mode

    def set_log_level(self, level: str) -> None:
        from ..utils.logging import set_log_level
        set_log_level(level)

    def add_listener(self, fn) -> None:
        self.scheduler.listeners.add(fn)

    # ---- observability (web UI / REST / metrics) ---------------------------------------------------------
    def start_ui(self, host: str = "127.0.0.1", port: int = 0, event_log: Optional[str] = None):
        """Attach the status store + metrics registry to the scheduler's listener bus and serve them over HTTP
        (``spark.ui.port``; REST under ``/api/v1/applications``, ``/metrics/json``, ``/metrics/prometheus``).  ``event_log``
        additionally appends every event to a JSONL file the :class:`HistoryServer` can replay (``spark.eventLog.enabled``)."""
        from ..utils.metrics import MetricsRegistry, SchedulerSource
        from ..utils.status import AppStatusStore
        from ..utils.status_server import StatusServer
        if getattr(self, "ui", None) is not None:
            return self.ui
        app_id = getattr(getattr(self.scheduler, "transport", None), "app_id", None)
        self.status_store = AppStatusStore(self.app_name, app_id).attach(
            self.scheduler, {"master": self.master, "app_name": self.app_name, "conf": self.conf.to_dict(),
                             "num_executors": self.scheduler.num_executors})
        self.metrics = MetricsRegistry()
        SchedulerSource(self.metrics, self.scheduler)
        if event_log:
            from ..utils.eventlog import EventLog
            self._event_log = EventLog(event_log).attach(self.scheduler)
        self.ui = StatusServer(self.status_store, self.metrics, host, port)
        return self.ui

    def enable_profiling(self):
        """``spark.python.profile``: profile every task; read the result with ``sc.profiler.show_profiles()``."""
        from ..utils.profiler import TaskProfiler
        self.profiler = TaskProfiler().attach(self.scheduler)
        return self.profiler

    def show_profiles(self, top: int = 10) -> str:
        return self.profiler.show_profiles(top)

    def dump_profiles(self, directory: str) -> None:
        self.profiler.dump_profiles(directory)

    def stop(self) -> None:
        if not self._stopped:
            if getattr(self, "metrics_system", None) is not None:
                self.metrics_system.stop()
            if getattr(self, "ui", None) is not None:
                self.status_store.stop()
                self.ui.stop()
            if getattr(self, "_event_log", None) is not None:
                self._event_log.close()
            self.scheduler.stop()
            self._stopped = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
