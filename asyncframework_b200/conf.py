"""``AsyncConf`` — string-keyed configuration with typed entries.

Reference: ``SparkConf`` (core/SparkConf.scala:82-119 set/get, :244-413 typed getters with time / size suffixes,
deprecated-key translation), typed ``ConfigBuilder`` entries (core/internal/config/package.scala, e.g.
``spark.task.maxFailures`` :155), ``spark-defaults.conf`` loading by ``SparkSubmitArguments`` and the
``conf/*.template`` files.

``AsyncConfig`` (config.py) stays the algorithm-level dataclass the five drivers parse from the reference's
positional CLIs; ``AsyncConf`` is the *runtime* configuration: scheduler, executors, RPC, storage, broadcast,
event log, UI, metrics.  Keys keep the reference's ``spark.*`` spelling so an existing ``spark-defaults.conf`` /
``--conf`` line works unchanged; unknown keys are kept verbatim (and forwarded to executors) like SparkConf does.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

_SIZE = {"b": 1, "k": 1 << 10, "kb": 1 << 10, "m": 1 << 20, "mb": 1 << 20, "g": 1 << 30, "gb": 1 << 30, "t": 1 << 40, "tb": 1 << 40}
_TIME = {"us": 1e-6, "ms": 1e-3, "s": 1.0, "m": 60.0, "min": 60.0, "h": 3600.0, "d": 86400.0}


def parse_size(text, default_unit: str = "b") -> int:
    """``"512m"`` / ``"1g"`` / ``"4096"`` -> bytes (JavaUtils.byteStringAs)."""
    if isinstance(text, (int, float)):
        return int(text * _SIZE[default_unit])
    m = re.fullmatch(r"\s*(-?[\d.]+)\s*([a-zA-Z]*)\s*", str(text))
    if not m or m.group(2).lower() not in _SIZE | {"": 1}:
        raise ValueError(f"invalid size string {text!r} (use e.g. 512m, 2g)")
    return int(float(m.group(1)) * (_SIZE[m.group(2).lower()] if m.group(2) else _SIZE[default_unit]))


def parse_time(text, default_unit: str = "s") -> float:
    """``"120s"`` / ``"10ms"`` / ``"2min"`` / ``"5"`` -> seconds (JavaUtils.timeStringAs)."""
    if isinstance(text, (int, float)):
        return float(text) * _TIME[default_unit]
    m = re.fullmatch(r"\s*(-?[\d.]+)\s*([a-zA-Z]*)\s*", str(text))
    if not m or m.group(2).lower() not in _TIME | {"": 1}:
        raise ValueError(f"invalid time string {text!r} (use e.g. 100ms, 120s, 2min)")
    return float(m.group(1)) * (_TIME[m.group(2).lower()] if m.group(2) else _TIME[default_unit])


def _parse_bool(text) -> bool:
    if isinstance(text, bool):
        return text
    low = str(text).strip().lower()
    if low in ("true", "1", "yes", "on"):
        return True
    if low in ("false", "0", "no", "off"):
        return False
    raise ValueError(f"invalid boolean {text!r}")


@dataclass(frozen=True)
class ConfigEntry:
    key: str
    parse: Callable[[Any], Any]
    default: Any
    doc: str
    check: Optional[Callable[[Any], bool]] = None
    alternatives: Tuple[str, ...] = ()          # deprecated spellings still honoured

    def read(self, raw: Any) -> Any:
        try:
            v = self.parse(raw)
        except (ValueError, TypeError) as exc:
            raise ValueError(f"{self.key}: {exc}") from None
        if self.check is not None and not self.check(v):
            raise ValueError(f"{self.key}: value {raw!r} is out of range")
        return v


ENTRIES: Dict[str, ConfigEntry] = {}


def _entry(key, parse, default, doc, check=None, alternatives=()) -> ConfigEntry:
    e = ConfigEntry(key, parse, default, doc, check, tuple(alternatives))
    ENTRIES[key] = e
    return e


_pos = lambda v: v > 0      # noqa: E731
_nonneg = lambda v: v >= 0  # noqa: E731

MASTER = _entry("spark.master", str, "local[*]", "master URL: local, local[N], local[N,F], local-cluster[n,c,m], spark://host:port, fused, gloo")
APP_NAME = _entry("spark.app.name", str, "asyncframework_b200", "application name shown by the master and the UI")
TASK_MAX_FAILURES = _entry("spark.task.maxFailures", int, 4, "attempts per task before the job fails (internal/config/package.scala:155)", _pos)
EXECUTOR_INSTANCES = _entry("spark.executor.instances", int, 0, "executors to request from the standalone master (0 = one per alive worker)", _nonneg)
EXECUTOR_CORES = _entry("spark.executor.cores", int, 1, "cores per executor", _pos)
EXECUTOR_MEMORY = _entry("spark.executor.memory", lambda t: parse_size(t, "m"), 1 << 30, "memory per executor (README.md:46 uses 15g)", _pos)
DRIVER_MEMORY = _entry("spark.driver.memory", lambda t: parse_size(t, "m"), 1 << 30, "driver memory (README.md:46 uses 20g)", _pos)
CORES_MAX = _entry("spark.cores.max", int, 0, "cap on the application's total cores (0 = unlimited)", _nonneg)
MEMORY_FRACTION = _entry("spark.memory.fraction", float, 0.6, "share of executor memory usable for block storage", lambda v: 0 < v <= 1)
MAX_DIRECT_RESULT = _entry("spark.task.maxDirectResultSize", parse_size, 1 << 20, "results above this travel through the block manager (Executor.scala:140-141)", _pos)
MAX_RESULT_SIZE = _entry("spark.driver.maxResultSize", parse_size, 1 << 30, "abort a job whose serialized results exceed this (0 = unlimited)", _nonneg)
BROADCAST_BLOCK = _entry("spark.broadcast.blockSize", lambda t: parse_size(t, "k"), 4 << 20, "TorrentBroadcast piece size (TorrentBroadcast.scala:81)", _pos)
BROADCAST_INLINE = _entry("spark.async.broadcast.inlineThreshold", parse_size, 32 << 10,
                          "broadcast values up to this serialized size travel inside the handle (0 = always use torrent pieces)", _nonneg)
BROADCAST_COMPRESS = _entry("spark.broadcast.compress", _parse_bool, True, "compress broadcast pieces")
BROADCAST_CHECKSUM = _entry("spark.broadcast.checksum", _parse_bool, True, "Adler32 per piece")
IO_CODEC = _entry("spark.io.compression.codec", lambda s: str(s).lower().rsplit(".", 1)[-1].replace("compressioncodec", ""), "lz4",
                  "lz4 | zstd | zlib | lzma | bz2 | none (CompressionCodec.scala:58)")
SERIALIZER = _entry("spark.serializer", lambda s: "pickle" if "java" in str(s).lower() or "kryo" in str(s).lower() else str(s).lower(), "pickle",
                    "data serializer: pickle | cloudpickle (JavaSerializer / KryoSerializer map to pickle)")
RPC_MAX_SIZE = _entry("spark.rpc.message.maxSize", lambda t: parse_size(t, "m"), 128 << 20,
                      "largest task binary (closure + lineage, source data excluded) the driver will ship; MiB when no unit", _pos)
NETWORK_TIMEOUT = _entry("spark.network.timeout", parse_time, 120.0, "RPC ask / block fetch timeout", _pos)
HEARTBEAT_INTERVAL = _entry("spark.executor.heartbeatInterval", parse_time, 10.0, "liveness heartbeat period (Executor.scala:770-815)", _pos)
WORKER_TIMEOUT = _entry("spark.worker.timeout", parse_time, 60.0, "master declares a silent worker dead after this", _pos)
LOCAL_DIR = _entry("spark.local.dir", str, "", "scratch directory of the disk block store (empty = system temp)")
AUTH = _entry("spark.authenticate", _parse_bool, False, "require the shared secret on every RPC connection")
AUTH_SECRET = _entry("spark.authenticate.secret", str, "", "the shared secret")
CRYPTO = _entry("spark.network.crypto.enabled", _parse_bool, False, "AES-256-GCM on every RPC / block-transfer frame (needs spark.authenticate)")
SPECULATION = _entry("spark.speculation", _parse_bool, False, "re-launch slow tasks speculatively (TaskSetManager.scala:64-70)")
SPECULATION_MULTIPLIER = _entry("spark.speculation.multiplier", float, 1.5, "how many times slower than the median a task must be", _pos)
SPECULATION_QUANTILE = _entry("spark.speculation.quantile", float, 0.75, "fraction of tasks that must finish before speculating", lambda v: 0 < v <= 1)
SPECULATION_INTERVAL = _entry("spark.speculation.interval", lambda t: parse_time(t, "ms"), 0.1, "how often to look for slow tasks", _pos)
BLACKLIST = _entry("spark.blacklist.enabled", _parse_bool, False, "stop scheduling on executors that keep failing (BlacklistTracker.scala:50)")
BLACKLIST_MAX_FAILURES = _entry("spark.blacklist.application.maxFailedTasksPerExecutor", int, 2, "failed tasks before an executor is blacklisted", _pos)
BLACKLIST_TIMEOUT = _entry("spark.blacklist.timeout", parse_time, 3600.0, "how long an executor stays blacklisted", _pos)
EVENT_LOG = _entry("spark.eventLog.enabled", _parse_bool, False, "append every scheduler event to a JSONL file")
EVENT_LOG_DIR = _entry("spark.eventLog.dir", str, "/tmp/afb-events", "directory of the event logs (history server input)")
UI_ENABLED = _entry("spark.ui.enabled", _parse_bool, False, "serve the live status UI / REST API")
UI_PORT = _entry("spark.ui.port", int, 4040, "port of the live UI (0 = any free port)", _nonneg)
PROFILE = _entry("spark.python.profile", _parse_bool, False, "run every task under cProfile; sc.show_profiles() / sc.dump_profiles(dir)")
METRICS_CONF = _entry("spark.metrics.conf", str, "", "path of a metrics.properties file")
LOCALITY_WAIT = _entry("spark.locality.wait", parse_time, 3.0, "accepted for compatibility: shard->executor affinity is hard here, nothing waits")
DEFAULT_PARALLELISM = _entry("spark.default.parallelism", int, 0, "default number of partitions (0 = number of executors)", _nonneg)
BLOCK_FIRST_ASYNC = _entry("spark.async.blockFirstJob", _parse_bool, True, "reproduce the reference's blocking first async job (SURVEY B.4)")
SHUFFLE_MEMORY = _entry("spark.shuffle.spill.memoryThreshold", parse_size, 256 << 20,
                        "estimated bytes a shuffle sorter / combine map holds before it spills a sorted run (ExternalSorter.scala:89)", _pos)
SHUFFLE_FORCE_SPILL = _entry("spark.shuffle.spill.numElementsForceSpillThreshold", int, 0, "spill after this many in-memory records (0 = off)", _nonneg)
SHUFFLE_SPILL_BATCH = _entry("spark.shuffle.spill.batchSize", int, 10000, "records per serialized batch of a spill file", _pos)
SHUFFLE_SPILL_COMPRESS = _entry("spark.shuffle.spill.compress", _parse_bool, True, "compress spill files")
DYN_ALLOC = _entry("spark.dynamicAllocation.enabled", _parse_bool, False, "grow / shrink the executor set with the task backlog (ExecutorAllocationManager.scala:228)")
DYN_ALLOC_MIN = _entry("spark.dynamicAllocation.minExecutors", int, 1, "lower bound of the executor target", _pos)
DYN_ALLOC_MAX = _entry("spark.dynamicAllocation.maxExecutors", int, 1 << 30, "upper bound of the executor target", _pos)
DYN_ALLOC_INITIAL = _entry("spark.dynamicAllocation.initialExecutors", int, 0, "initial executor target (0 = minExecutors)", _nonneg)
DYN_ALLOC_BACKLOG = _entry("spark.dynamicAllocation.schedulerBacklogTimeout", parse_time, 1.0, "backlog must persist this long before executors are added", _pos)
DYN_ALLOC_SUSTAINED = _entry("spark.dynamicAllocation.sustainedSchedulerBacklogTimeout", parse_time, 1.0, "interval of the following (doubling) additions", _pos)
DYN_ALLOC_IDLE = _entry("spark.dynamicAllocation.executorIdleTimeout", parse_time, 60.0, "remove an executor that ran no task for this long", _pos)
RECOVER_EXECUTORS = _entry("spark.async.recoverLostExecutors", _parse_bool, True, "replace a lost executor and re-run its tasks instead of failing the job")

DEPRECATED = {      # old key -> (new key, version it changed in), SparkConf.deprecatedConfigs / configsWithAlternatives
    "spark.akka.timeout": ("spark.network.timeout", "1.4"),
    "spark.core.connection.ack.wait.timeout": ("spark.network.timeout", "1.4"),
    "spark.kryoserializer.buffer.mb": ("spark.kryoserializer.buffer", "1.4"),
    "spark.executor.heartbeat.interval": ("spark.executor.heartbeatInterval", "2.0"),
}


class AsyncConf:
    def __init__(self, load_defaults: bool = True):
        self._settings: Dict[str, str] = {}
        self.warnings: List[str] = []
        if load_defaults:
            path = self.defaults_file()
            if path is not None:
                self.load_properties_file(path)
            for k, v in os.environ.items():                # java system properties `-Dspark.*` become env AFB_CONF_spark_x_y
                if k.startswith("AFB_CONF_"):
                    self.set(k[len("AFB_CONF_"):].replace("__", "-").replace("_", "."), v)

    # ---- files ---------------------------------------------------------------------------------------------------
    @staticmethod
    def defaults_file() -> Optional[str]:
        cand = []
        if os.environ.get("AFB_CONF_DIR"):
            cand.append(os.path.join(os.environ["AFB_CONF_DIR"], "async-defaults.conf"))
        cand.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "conf", "async-defaults.conf"))
        for c in cand:
            if os.path.isfile(c):
                return c
        return None

    def load_properties_file(self, path: str) -> "AsyncConf":
        """``spark-defaults.conf`` syntax: ``key value`` or ``key=value`` per line, ``#`` comments."""
        with open(path) as f:
            for ln, line in enumerate(f, 1):
                line = line.strip()
                if not line or line.startswith("#"):
                    continue
                m = re.match(r"([^\s=:]+)\s*[=:\s]\s*(.*)$", line)
                if not m:
                    raise ValueError(f"{path}:{ln}: cannot parse {line!r}")
                self.set_if_missing(m.group(1), m.group(2).strip())
        return self

    # ---- set / get ------------------------------------------------------------------------------------------------
    def set(self, key: str, value: Any) -> "AsyncConf":
        if key is None or value is None:
            raise ValueError("null key or value")
        if key in DEPRECATED:
            new, ver = DEPRECATED[key]
            self.warnings.append(f"The configuration key '{key}' has been deprecated as of version {ver}; use '{new}'")
            key = new
        entry = ENTRIES.get(key)
        if entry is not None:
            entry.read(value)                              # validate eagerly: a bad value fails where it is set
        self._settings[key] = str(value) if not isinstance(value, bool) else ("true" if value else "false")
        return self

    def set_all(self, items: Iterable[Tuple[str, Any]]) -> "AsyncConf":
        for k, v in (items.items() if isinstance(items, dict) else items):
            self.set(k, v)
        return self

    def set_if_missing(self, key: str, value: Any) -> "AsyncConf":
        if key not in self._settings:
            self.set(key, value)
        return self

    def set_master(self, master: str) -> "AsyncConf":
        return self.set("spark.master", master)

    def set_app_name(self, name: str) -> "AsyncConf":
        return self.set("spark.app.name", name)

    def remove(self, key: str) -> "AsyncConf":
        self._settings.pop(key, None)
        return self

    def contains(self, key: str) -> bool:
        return key in self._settings

    def get(self, key: str, default: Any = None) -> Any:
        """Typed value for registered keys (the entry's default when unset), the raw string otherwise."""
        entry = ENTRIES.get(key)
        if key in self._settings:
            return entry.read(self._settings[key]) if entry is not None else self._settings[key]
        if entry is not None:
            for alt in entry.alternatives:
                if alt in self._settings:
                    return entry.read(self._settings[alt])
            return entry.default if default is None else default
        if default is None:
            raise KeyError(key)
        return default

    def get_raw(self, key: str, default: Optional[str] = None) -> Optional[str]:
        return self._settings.get(key, default)

    def get_int(self, key: str, default: int = 0) -> int:
        return int(self.get(key, default))

    def get_double(self, key: str, default: float = 0.0) -> float:
        return float(self.get(key, default))

    def get_boolean(self, key: str, default: bool = False) -> bool:
        v = self.get(key, default)
        return v if isinstance(v, bool) else _parse_bool(v)

    def get_size_as_bytes(self, key: str, default: Any = 0) -> int:
        v = self._settings.get(key)
        return parse_size(default if v is None else v)

    def get_time_as_seconds(self, key: str, default: Any = 0) -> float:
        v = self._settings.get(key)
        return parse_time(default if v is None else v)

    def get_all(self) -> List[Tuple[str, str]]:
        return sorted(self._settings.items())

    def get_all_with_prefix(self, prefix: str) -> List[Tuple[str, str]]:
        return sorted((k[len(prefix):], v) for k, v in self._settings.items() if k.startswith(prefix))

    def clone(self) -> "AsyncConf":
        c = AsyncConf(load_defaults=False)
        c._settings = dict(self._settings)
        return c

    def to_debug_string(self) -> str:
        return "\n".join(f"{k}={'*********(redacted)' if 'secret' in k.lower() or 'password' in k.lower() else v}"
                         for k, v in self.get_all())

    def validate(self) -> "AsyncConf":
        """``SparkConf.validateSettings``: cross-field checks."""
        if self.get("spark.authenticate") and not self.get("spark.authenticate.secret"):
            raise ValueError("spark.authenticate is on but spark.authenticate.secret is not set")
        if self.get("spark.network.crypto.enabled") and not self.get("spark.authenticate"):
            raise ValueError("spark.network.crypto.enabled requires spark.authenticate=true and a secret")
        if self.get("spark.task.maxDirectResultSize") > self.get("spark.driver.maxResultSize") > 0:
            raise ValueError("spark.task.maxDirectResultSize must not exceed spark.driver.maxResultSize")
        for k, v in self._settings.items():
            if k in ENTRIES:
                ENTRIES[k].read(v)
        return self

    # ---- bridges ----------------------------------------------------------------------------------------------------
    def runtime_limits(self):
        """The :class:`RuntimeLimits` the driver / executors' AsyncEnv is built from."""
        from .core.env import RuntimeLimits
        g = self.get
        return RuntimeLimits(
            max_direct_result_size=g("spark.task.maxDirectResultSize"), max_result_size=g("spark.driver.maxResultSize"),
            broadcast_block_size=g("spark.broadcast.blockSize"), broadcast_compress=g("spark.broadcast.compress"),
            broadcast_inline_threshold=g("spark.async.broadcast.inlineThreshold"),
            broadcast_checksum=g("spark.broadcast.checksum"),
            storage_memory=int(g("spark.executor.memory") * g("spark.memory.fraction")),
            compression_codec=g("spark.io.compression.codec"), serializer=g("spark.serializer"),
            local_dir=g("spark.local.dir") or None, network_timeout=g("spark.network.timeout"),
            heartbeat_interval=g("spark.executor.heartbeatInterval"),
            auth_secret=(g("spark.authenticate.secret") or None) if g("spark.authenticate") else None,
            encrypt=bool(g("spark.network.crypto.enabled") and g("spark.authenticate")),
            shuffle_memory=g("spark.shuffle.spill.memoryThreshold"),
            shuffle_spill_elements=g("spark.shuffle.spill.numElementsForceSpillThreshold"),
            shuffle_spill_batch=g("spark.shuffle.spill.batchSize"), shuffle_spill_compress=g("spark.shuffle.spill.compress"))

    @staticmethod
    def describe() -> str:
        """The documented key table (``docs/configuration.md`` of the reference)."""
        rows = [f"{e.key:<58} {str(e.default):<22} {e.doc}" for e in sorted(ENTRIES.values(), key=lambda e: e.key)]
        return "\n".join([f"{'key':<58} {'default':<22} meaning", "-" * 120] + rows)
