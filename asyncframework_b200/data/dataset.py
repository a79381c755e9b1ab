"""``ShardedDataset`` — the partitioned, lazily-evaluated dataset with the ASYNC actions.

Covers the slice of the RDD API the ASYNC samples use (core/rdd/RDD.scala): ``map`` :370,
``repartition`` :420-475, ``sample`` :488-503, ``mapPartitionsWithIndex`` :847, ``zipWithIndex``
:1527 (+ ZippedWithIndexRDD.scala:43-69), ``cache`` :207, ``count`` :1392, ``reduce`` :1017-1042,
``aggregate`` :1252, ``treeAggregate`` :1358-1388 — and the ASYNC additions ``ASYNCbarrier``
:1050-1077, ``ASYNCreduce`` :1087-1171, ``ASYNCaggregate`` :1268-1345.

Two evaluation granularities share one lineage graph:
  * record level  — iterators of Python records (``LabeledPoint`` …): full API generality,
    used by tests, tiny data and arbitrary user closures;
  * batch level   — a partition is one columnar ``Shard`` (dense/CSR tensors, possibly in HBM);
    ``sample``/``zip_with_index``/``async_barrier`` stay vectorised and ``map_batch`` hands the
    whole shard to torch code or to the fused CUDA kernels.  This is the fast path.
"""
from __future__ import annotations

import itertools
import math
import os
import threading
import time
from typing import Any, Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ..core import sampler as _sampler
from ..core.partial_result import PartialResult
from ..linalg.vectors import DenseVector, LabeledPoint, SparseVector, Vector
from .batch import Shard, concat_shards, padded_ld

class _Empty:
    """Picklable singleton marker for "no value" (survives a trip through an executor process)."""
    _inst = None

    def __new__(cls):
        if cls._inst is None:
            cls._inst = super().__new__(cls)
        return cls._inst

    def __reduce__(self):
        return (_Empty, ())

    def __repr__(self):
        return "<EMPTY>"


_EMPTY = _Empty()
_ds_ids = itertools.count(1)
# executor-side block store: cached partitions keyed by (dataset id, partition, kind); lives in whichever
# process evaluates the partition (BlockManager MEMORY_ONLY analogue for the process-based scheduler)
_BLOCK_STORE: dict = {}


def _cache_get(ds_id: int, p: int, kind: str):
    """(found, value) from the executor's BlockManager when this process has an AsyncEnv (cluster executors: the block is
    ``rdd_<dataset>_<partition>``, accounted, evictable and visible in the driver's storage status), else from the
    plain per-process dict."""
    from ..core.env import AsyncEnv
    env = AsyncEnv.get()
    if env is not None and not env.is_driver:
        return env.block_manager.get_local_values(f"rdd_{ds_id}_{p}{kind}")
    key = (ds_id, p, kind)
    return (True, _BLOCK_STORE[key]) if key in _BLOCK_STORE else (False, None)


def _cache_put(ds_id: int, p: int, kind: str, value, level=None) -> None:
    from ..core.env import AsyncEnv
    env = AsyncEnv.get()
    if env is not None and not env.is_driver:
        from ..storage.block_manager import StorageLevel
        # MEMORY_ONLY (``cache()``): a partition that does not fit is recomputed next time; ``persist(MEMORY_AND_DISK)`` /
        # ``DISK_ONLY`` spill to the executor's disk store instead (RDD.persist, storage/StorageLevel.scala)
        env.block_manager.put_single(f"rdd_{ds_id}_{p}{kind}", value, level or StorageLevel.MEMORY_ONLY)
        return
    _BLOCK_STORE[(ds_id, p, kind)] = value

# ---- ship-once source data (cluster mode) -------------------------------------------------------------------
# A dataset's SOURCE (the parallelised records / the HBM or host shards) sits at the root of its lineage, and the
# lineage is part of every job's task binary.  Spark ships a ParallelCollectionRDD's data with every task; for an
# iterative optimiser that submits one job per round that would re-send the whole training set each round.  Here the
# source is wrapped in a SourceBlock: while a ClusterScheduler pickles a task binary, a block that every target
# executor already holds is replaced by its id (a few bytes) and resolved from the executor-side store — the
# "task size should be small" property the reference's cluster tests check (LinearRegressionClusterSuite,
# mllib/regression/LinearRegressionSuite.scala:152-166).
_SOURCE_STORE: dict = {}
_block_ids = itertools.count(1)
_shuffle_ids = itertools.count(0)


def _spill_config():
    """Spill thresholds of the running task: the executor's ``RuntimeLimits`` (``spark.shuffle.spill.*``), else the process default."""
    from ..core.env import AsyncEnv
    from ..storage import external_sorter as xs
    e = AsyncEnv.get()
    return xs.SpillConfig.from_limits(e.limits) if e is not None else xs.DEFAULT_CONFIG


def _note_spill(t, m: dict) -> None:
    """Fold a sorter's counters into the task metrics (TaskMetrics.memoryBytesSpilled / diskBytesSpilled / peakExecutionMemory)."""
    tm = getattr(t, "metrics", None)
    if not isinstance(tm, dict):
        return
    tm["memory_bytes_spilled"] = tm.get("memory_bytes_spilled", 0) + m["memory_bytes_spilled"]
    tm["disk_bytes_spilled"] = tm.get("disk_bytes_spilled", 0) + m["disk_bytes_spilled"]
    tm["spills"] = tm.get("spills", 0) + m["spills"]
    tm["peak_execution_memory"] = max(tm.get("peak_execution_memory", 0), m["peak_memory"])
_pickle_ctx = threading.local()


def _stable_hash(key) -> int:
    """Partitioner hash that is identical in every process (Python's ``hash`` of str/bytes is salted per process, so executor
    processes would disagree on a key's partition)."""
    if isinstance(key, int) and not isinstance(key, bool):
        return key & 0x7FFFFFFFFFFFFFFF
    import pickle
    import zlib
    return zlib.crc32(pickle.dumps(key, protocol=4))


def _with_rows(sub: Shard, parent: Shard, rows: torch.Tensor) -> Shard:
    """Record the sampled rows' positions in the UNSAMPLED partition (``meta['rows']``): per-row state that lives beside the
    partition — the SAGA history table — is indexed with it.  Positions compose through nested samples."""
    prev = parent.meta.get("rows")
    sub.meta["rows"] = rows.to(torch.int64) if prev is None else prev.to(rows.device)[rows]
    return sub


class HashPartitioner:
    """``new HashPartitioner(n)``: process-independent hash of the key modulo ``n`` (Partitioner.scala:86)."""

    def __init__(self, num_partitions: int):
        self.num_partitions = int(num_partitions)

    def __call__(self, key) -> int:
        return _stable_hash(key) % self.num_partitions

    def __eq__(self, other) -> bool:
        return isinstance(other, HashPartitioner) and other.num_partitions == self.num_partitions

    __hash__ = None  # type: ignore[assignment]


class RangePartitioner:
    """``new RangePartitioner(n, rdd)``: boundaries from a sample of the keys; partition q holds keys in (bound[q-1], bound[q]]
    (Partitioner.scala:127; what ``sort_by_key`` builds internally)."""

    def __init__(self, num_partitions: int, keyed_dataset=None, ascending: bool = True, bounds: Optional[list] = None, key=lambda k: k):
        self.num_partitions, self.ascending, self.key = int(num_partitions), ascending, key
        if bounds is None:
            n = max(keyed_dataset.count(), 1)
            sample = sorted(key(k) for k in keyed_dataset.keys().sample(False, min(1.0, max(0.01, 20.0 * num_partitions / n)), 17).collect())
            bounds = [sample[len(sample) * (i + 1) // num_partitions] for i in range(num_partitions - 1)] if sample and num_partitions > 1 else []
        self.bounds = list(bounds)

    def __call__(self, k) -> int:
        import bisect
        q = bisect.bisect_left(self.bounds, self.key(k))
        return q if self.ascending else (len(self.bounds) - q)


class BlockMissing(RuntimeError):
    """An executor was sent a block reference for data it does not hold (e.g. a replaced executor)."""


def drop_source_block(block_id: int) -> bool:
    """Executor side of the cleaner: forget a shipped source block."""
    return _SOURCE_STORE.pop(block_id, None) is not None


class SourceBlock:
    __slots__ = ("block_id", "data", "__weakref__")

    def __init__(self, data, block_id: Optional[int] = None):
        self.block_id = block_id if block_id is not None else ((os.getpid() << 32) | next(_block_ids))
        self.data = data

    @staticmethod
    def _with_data(block_id, data):
        _SOURCE_STORE[block_id] = data
        return SourceBlock(data, block_id)

    @staticmethod
    def _from_store(block_id):
        try:
            return SourceBlock(_SOURCE_STORE[block_id], block_id)
        except KeyError:
            raise BlockMissing(f"source block {block_id:#x} is not in this executor's store") from None

    def __reduce__(self):
        ctx = getattr(_pickle_ctx, "ctx", None)
        if ctx is not None and ctx.holds(self.block_id):
            return (SourceBlock._from_store, (self.block_id,))
        if ctx is not None:
            ctx.shipping(self.block_id, self)
        return (SourceBlock._with_data, (self.block_id, self.data))


class ShippingContext:
    """Installed by ClusterScheduler.submit_job around cloudpickle.dumps: decides per block whether the data or only the
    reference travels, and records what was shipped."""

    def __init__(self, held_by_all_targets: Callable[[int], bool]):
        self.holds = held_by_all_targets
        self.shipped: List[int] = []
        self.blocks: dict = {}

    def shipping(self, block_id: int, block=None) -> None:
        self.shipped.append(block_id)
        if block is not None:
            self.blocks[block_id] = block

    def __enter__(self):
        _pickle_ctx.ctx = self
        return self

    def __exit__(self, *exc):
        _pickle_ctx.ctx = None
        return False


class ShardedDataset:
    def __init__(self, scheduler, num_partitions: int,
                 compute: Optional[Callable[[int, Any], Iterator]] = None,
                 batch: Optional[Callable[[int, Any], Optional[Shard]]] = None,
                 parent: Optional["ShardedDataset"] = None, name: str = "dataset"):
        self.scheduler = scheduler
        self.num_partitions = int(num_partitions)
        self._compute = compute
        self._batch = batch
        self.parent = parent
        self.name = name
        self._cached_batches: Optional[List[Optional[Shard]]] = None
        self._cached_records: Optional[List[list]] = None
        self._want_cache = False
        self._cache_lock = threading.Lock()
        self.selected_workers: Optional[List[int]] = None   # set by async_barrier
        self.ds_id = (os.getpid() << 32) | next(_ds_ids)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_shards(cls, scheduler, shards: Sequence[Shard], name: str = "shards") -> "ShardedDataset":
        shards = list(shards)
        for p, s in enumerate(shards):
            s.partition = p
        blk = SourceBlock(shards)

        def batch(p, tctx):
            return blk.data[p]

        def compute(p, tctx):
            s = blk.data[p]
            return s.indexed_records() if s.gidx is not None else s.records()

        return cls(scheduler, len(shards), compute, batch, None, name)

    @classmethod
    def from_records(cls, scheduler, records: Sequence[Any], num_partitions: int, name: str = "records"):
        """``sc.parallelize``: contiguous slices."""
        records = list(records)
        n = len(records)
        bounds = [n * i // num_partitions for i in range(num_partitions + 1)]
        blk = SourceBlock([records[bounds[i]:bounds[i + 1]] for i in range(num_partitions)])
        return cls(scheduler, num_partitions, lambda p, t: iter(blk.data[p]), None, None, name)

    @classmethod
    def from_arrays(cls, scheduler, X, y, num_partitions: int, dtype=torch.float64, device=None,
                    name: str = "arrays") -> "ShardedDataset":
        X = torch.as_tensor(X, dtype=dtype)
        y = torch.as_tensor(y, dtype=dtype)
        n, d = X.shape
        ld = padded_ld(d)
        shards = []
        for p in range(num_partitions):   # round-robin like repartition
            rows = torch.arange(p, n, num_partitions) if p < n else torch.zeros(0, dtype=torch.int64)
            Xp = torch.zeros(rows.numel(), ld, dtype=dtype)
            Xp[:, :d] = X[rows]
            shards.append(Shard(d, y[rows].clone(), Xp, partition=p).to(device))
        return cls.from_shards(scheduler, shards, name)

    # ------------------------------------------------------------------ evaluation
    @property
    def batch_capable(self) -> bool:
        return self._batch is not None

    # -- pickling (task binaries shipped to executor processes): the scheduler and locks stay behind ----
    def __getstate__(self):
        st = dict(self.__dict__)
        st["scheduler"] = None
        st["_cache_lock"] = None
        st["_cached_batches"] = None
        st["_cached_records"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._cache_lock = threading.Lock()

    def iterator(self, p: int, tctx=None) -> Iterator:
        if self._cached_records is not None and self._cached_records[p] is not None:
            return iter(self._cached_records[p])
        if self._want_cache:
            found, recs = _cache_get(self.ds_id, p, "r")
            if found:
                return iter(recs)
        if self._compute is not None:
            it = self._compute(p, tctx)
        else:
            b = self.batch(p, tctx)
            it = iter(()) if b is None else (b.indexed_records() if b.gidx is not None else b.records())
        if self._want_cache and not self.batch_capable:
            recs = list(it)
            with self._cache_lock:
                if self._cached_records is None:
                    self._cached_records = [None] * self.num_partitions
                self._cached_records[p] = recs
            _cache_put(self.ds_id, p, "r", recs, getattr(self, "_storage_level", None))
            return iter(recs)
        return it

    def batch(self, p: int, tctx=None) -> Optional[Shard]:
        if self._batch is None:
            raise TypeError(f"dataset {self.name!r} has no batch-level view (a record-level map broke the columnar lineage)")
        if self._cached_batches is not None and self._cached_batches[p] is not _EMPTY:
            return self._cached_batches[p]
        if self._want_cache:
            found, blk = _cache_get(self.ds_id, p, "b")
            if found:
                return blk
        b = self._batch(p, tctx)
        if self._want_cache:
            with self._cache_lock:
                if self._cached_batches is None:
                    self._cached_batches = [_EMPTY] * self.num_partitions
                self._cached_batches[p] = b
            _cache_put(self.ds_id, p, "b", b, getattr(self, "_storage_level", None))
        return b

    def cache(self) -> "ShardedDataset":
        """Mark for caching; materialised lazily by the first job (RDD.scala:207)."""
        self._want_cache = True
        return self

    def persist(self, level=None) -> "ShardedDataset":
        """``rdd.persist(StorageLevel)``: like ``cache()`` but executor processes store the partitions at ``level``
        (``MEMORY_AND_DISK`` / ``DISK_ONLY`` spill instead of recomputing; in-process masters keep references either way)."""
        self._storage_level = level
        return self.cache()

    def get_storage_level(self):
        from ..storage.block_manager import StorageLevel
        if not self._want_cache:
            return StorageLevel.NONE
        return getattr(self, "_storage_level", None) or StorageLevel.MEMORY_ONLY

    def checkpoint(self, directory: str) -> "ShardedDataset":
        """``rdd.checkpoint()`` (RDD.scala:1773, ReliableCheckpointRDD.scala:121-210): every partition is materialised by a job
        and written to ``directory/part-<p>.pt`` (columnar shard when the dataset is batch-capable, records otherwise); the
        returned dataset reads those files and has NO parent — the lineage is truncated, so later jobs (and task binaries in
        cluster mode) no longer carry or recompute it.  Eager, unlike Spark's mark-then-write-after-the-next-job."""
        os.makedirs(directory, exist_ok=True)
        batchy = self.batch_capable

        def write(tctx, it):
            p = tctx.partition_id
            payload = {"batch": self.batch(p, tctx)} if batchy else {"records": list(it)}
            if batchy and payload["batch"] is not None:
                payload["batch"] = payload["batch"].to("cpu")
            tmp = os.path.join(directory, f"part-{p:05d}.pt.tmp")
            torch.save(payload, tmp)
            os.replace(tmp, os.path.join(directory, f"part-{p:05d}.pt"))
            return p
        self._run(write)
        with open(os.path.join(directory, "_SUCCESS"), "w") as f:
            f.write(f"{self.num_partitions} {int(batchy)}\n")
        return ShardedDataset.from_checkpoint(self.scheduler, directory)

    def local_checkpoint(self) -> "ShardedDataset":
        """``rdd.localCheckpoint()`` (RDD.scala:1600-1640): materialise the partitions once at ``MEMORY_AND_DISK`` in the
        executors' block stores and serve every later job from there, without reliable storage (lost with its executor —
        the reference trades fault tolerance for speed the same way).  Unlike :meth:`checkpoint` the parent chain is kept as
        the recomputation fallback for a lost block."""
        from ..storage.block_manager import StorageLevel
        self.persist(StorageLevel.MEMORY_AND_DISK)
        self._run(lambda t, it: sum(1 for _ in it))              # materialise now
        self._local_checkpointed = True
        return self

    def is_checkpointed(self) -> bool:
        """``rdd.isCheckpointed``: true for datasets re-opened from a reliable checkpoint and for locally checkpointed ones."""
        return bool(getattr(self, "_checkpoint_dir", None) or getattr(self, "_local_checkpointed", False))

    def get_checkpoint_file(self) -> Optional[str]:
        return getattr(self, "_checkpoint_dir", None)

    @classmethod
    def from_checkpoint(cls, scheduler, directory: str) -> "ShardedDataset":
        """Re-open a dataset written by :meth:`checkpoint` (``sc.checkpointFile``)."""
        with open(os.path.join(directory, "_SUCCESS")) as f:
            n, batchy = (int(t) for t in f.read().split())

        def load(p):
            return torch.load(os.path.join(directory, f"part-{p:05d}.pt"), weights_only=False)

        def compute(p, tctx):
            d = load(p)
            if "records" in d:
                return iter(d["records"])
            s = d["batch"]
            return iter(()) if s is None else (s.indexed_records() if s.gidx is not None else s.records())

        def batch(p, tctx):
            s = load(p)["batch"]
            if s is not None:
                s.partition = p
            return s
        ds = cls(scheduler, n, compute, batch if batchy else None, None, f"checkpoint:{os.path.basename(directory)}")
        ds._checkpoint_dir = directory
        return ds

    def _derive(self, compute=None, batch=None, name="") -> "ShardedDataset":
        ds = ShardedDataset(self.scheduler, self.num_partitions, compute, batch, self, name or self.name)
        ds.selected_workers = self.selected_workers
        return ds

    # ------------------------------------------------------------------ transformations
    def map(self, f: Callable[[Any], Any]) -> "ShardedDataset":
        return self._derive(lambda p, t: (f(r) for r in self.iterator(p, t)), None, "map")

    def filter(self, f: Callable[[Any], bool]) -> "ShardedDataset":
        return self._derive(lambda p, t: (r for r in self.iterator(p, t) if f(r)), None, "filter")

    def flat_map(self, f) -> "ShardedDataset":
        return self._derive(lambda p, t: itertools.chain.from_iterable(f(r) for r in self.iterator(p, t)), None, "flat_map")

    def map_partitions_with_index(self, f: Callable[[int, Iterator], Iterator],
                                  preserves_batch: bool = False) -> "ShardedDataset":
        """``f(index, iterator) -> iterator``.  With ``preserves_batch=True`` the function is a
        side-effect hook (e.g. the delay injector): it is invoked once per partition with an
        empty iterator on the batch path and the shard passes through untouched."""
        batch = None
        if preserves_batch and self.batch_capable:
            def batch(p, t):  # noqa: E306
                b = self.batch(p, t)
                if b is not None:
                    for _ in f(p, iter(())):
                        pass
                return b
        return self._derive(lambda p, t: f(p, self.iterator(p, t)), batch, "map_partitions_with_index")

    def map_partitions(self, f: Callable[[Iterator], Iterator]) -> "ShardedDataset":
        return self._derive(lambda p, t: f(self.iterator(p, t)), None, "map_partitions")

    def map_batch(self, f: Callable[[Shard], Any]) -> "ShardedDataset":
        """Batch-level map: one output record ``f(shard)`` per non-empty partition."""
        def compute(p, t):
            b = self.batch(p, t)
            if b is None or b.num_rows == 0:
                return iter(())
            return iter((f(b),))
        return self._derive(compute, None, "map_batch")

    def map_shards(self, f: Callable[[Shard], Shard]) -> "ShardedDataset":
        """Columnar transformation ``Shard -> Shard`` that KEEPS the batch-level lineage (feature scaling, bias append, dtype /
        device moves …): the result is still batch-capable, so ``sample`` / ``async_grad`` / the fused kernels apply to it."""
        if not self.batch_capable:
            raise TypeError(f"dataset {self.name!r} has no batch-level view to transform")

        def batch(p, t):
            b = self.batch(p, t)
            if b is None:
                return None
            out = f(b)
            out.partition = p
            return out

        def compute(p, t):
            b = batch(p, t)
            return iter(()) if b is None else (b.indexed_records() if b.gidx is not None else b.records())
        return self._derive(compute, batch, "map_shards")

    def to(self, device=None, dtype=None) -> "ShardedDataset":
        """Move every shard to ``device`` / ``dtype`` where the partition is EVALUATED and cache it there: in an executor process
        whose worker pinned a GPU (``deploy.worker --gpu I``) ``ds.to("cuda")`` makes the partition HBM-resident in that executor,
        so later ``sample`` / ``async_grad`` rounds run the fused CUDA kernels on it."""
        return self.map_shards(lambda s: s.to(device, dtype)).cache()

    def glom(self) -> "ShardedDataset":
        return self._derive(lambda p, t: iter((list(self.iterator(p, t)),)), None, "glom")

    def sample(self, with_replacement: bool, fraction: float, seed: Optional[int] = None) -> "ShardedDataset":
        """Bernoulli sample without replacement (the only mode the reference's drivers use).
        Membership = hash(seed, global row index) < fraction — see core/sampler.py; an unseeded
        call draws a fresh seed like the reference's unseeded ``.sample(false, 0.1)`` (B.9)."""
        if seed is None:
            seed = int.from_bytes(np.random.bytes(8), "little")
        seed = int(seed)
        if with_replacement:
            return self._sample_with_replacement(fraction, seed)

        def batch(p, t):
            b = self.batch(p, t)
            if b is None:
                return None
            if b.device.type == "cuda":
                from ..ops.sampling import sample_rows          # K11 compaction kernel, same hash
                rows = sample_rows(b.num_rows, fraction, seed, b.gidx, p << 40, b.device)
                return _with_rows(b.select(rows), b, rows)
            idx = b.gidx if b.gidx is not None else (torch.arange(b.num_rows, dtype=torch.int64) + (p << 40))
            mask = _sampler.bernoulli_mask(seed, idx, fraction)
            rows = torch.from_numpy(np.nonzero(mask)[0])
            return _with_rows(b.select(rows), b, rows)

        def compute(p, t):
            # record-level path: membership of a CHUNK of records is evaluated with one vectorised hash (a NumPy call per record
            # costs more than the hash itself); same rows as the batch-level path, the hash is a pure function of the row index
            thr = _sampler.threshold24(fraction)
            it = enumerate(self.iterator(p, t))
            while True:
                chunk = list(itertools.islice(it, 4096))
                if not chunk:
                    return
                gis = np.fromiter((rec[1] if (isinstance(rec, tuple) and len(rec) == 2 and isinstance(rec[1], int)) else (pos + (p << 40))
                                   for pos, rec in chunk), dtype=np.int64, count=len(chunk))
                keep = _sampler.hash_u24(seed, gis) < thr
                for (pos, rec), k in zip(chunk, keep):
                    if k:
                        yield rec

        return self._derive(compute, batch if self.batch_capable else None, "sample")

    def _sample_with_replacement(self, fraction: float, seed: int) -> "ShardedDataset":
        """``sample(withReplacement = true, fraction)``: every row appears Poisson(fraction) times (PoissonSampler); the counts
        are a pure function of (seed, global row index), see core/sampler.py::poisson_counts.  Not used by the ASYNC drivers
        (they sample without replacement) but part of the RDD contract (bootstrap resampling)."""
        def batch(p, t):
            b = self.batch(p, t)
            if b is None:
                return None
            idx = b.gidx if b.gidx is not None else (torch.arange(b.num_rows, dtype=torch.int64) + (p << 40))
            counts = torch.from_numpy(_sampler.poisson_counts(seed, idx, fraction))
            rows = torch.repeat_interleave(torch.arange(b.num_rows, dtype=torch.int64), counts)
            return _with_rows(b.select(rows), b, rows)

        def compute(p, t):
            for pos, rec in enumerate(self.iterator(p, t)):
                gi = rec[1] if (isinstance(rec, tuple) and len(rec) == 2 and isinstance(rec[1], int)) else (pos + (p << 40))
                for _ in range(int(_sampler.poisson_counts(seed, np.asarray([gi]), fraction)[0])):
                    yield rec
        return self._derive(compute, batch if self.batch_capable else None, "sample_with_replacement")

    def zip_with_index(self) -> "ShardedDataset":
        """Global index = partition offset + position; runs a count job over the partitions to
        get the offsets (ZippedWithIndexRDD.scala:43-56)."""
        sizes = self.partition_sizes()
        starts = [0] * self.num_partitions
        for p in range(1, self.num_partitions):
            starts[p] = starts[p - 1] + sizes[p - 1]

        def batch(p, t):
            b = self.batch(p, t)
            if b is None:
                return None
            g = torch.arange(b.num_rows, dtype=torch.int64, device=b.device) + starts[p]
            return Shard(b.num_features, b.y, b.X, b.indptr, b.indices, b.values, g, p, dict(b.meta))

        def compute(p, t):
            def strip(r):
                return r[0] if (isinstance(r, tuple) and len(r) == 2 and isinstance(r[0], LabeledPoint)) else r
            return ((strip(r), starts[p] + i) for i, r in enumerate(self.iterator(p, t)))

        return self._derive(compute, batch if self.batch_capable else None, "zip_with_index")

    def repartition(self, num_partitions: int) -> "ShardedDataset":
        """Round-robin redistribution (``coalesce(shuffle=true)``, RDD.scala:420-475); this is the
        one all-to-all of the pipeline and happens once at load time."""
        P = int(num_partitions)
        if self.batch_capable:
            olds = [self.batch(p) for p in range(self.num_partitions)]
            olds = [o for o in olds if o is not None]
            buckets: List[List[Shard]] = [[] for _ in range(P)]
            start = 0
            for o in olds:
                for q in range(P):
                    first = (q - start) % P
                    if first >= o.num_rows:
                        continue
                    rows = torch.arange(first, o.num_rows, P)
                    if rows.numel():
                        buckets[q].append(o.select(rows))
                start = (start + o.num_rows) % P
            d = olds[0].num_features if olds else 0
            shards = []
            for q in range(P):
                if buckets[q]:
                    shards.append(concat_shards(buckets[q], q))
                else:
                    proto = olds[0]
                    shards.append(proto.select(torch.zeros(0, dtype=torch.int64)))
            return ShardedDataset.from_shards(self.scheduler, shards, "repartition")
        parts: List[list] = [[] for _ in range(P)]
        i = 0
        for p in range(self.num_partitions):
            for r in self.iterator(p):
                parts[i % P].append(r)
                i += 1
        return ShardedDataset(self.scheduler, P, lambda p, t: iter(parts[p]), None, None, "repartition")

    # ------------------------------------------------------------------ ASYNC: barrier
    def async_barrier(self, predicate: Callable[[Any], bool], stat) -> "ShardedDataset":
        """Select the workers that take part in the next round (``ASYNCbarrier``,
        RDD.scala:1050-1077): partitions whose STAT entry satisfies ``predicate(state) and
        state.available`` — or that have no entry yet.  Non-selected partitions evaluate to empty.
        The selection is stored on the returned dataset (not in a process-global, quirk B.2)."""
        selected: List[int] = []
        for p in range(self.num_partitions):
            st = stat.get(p)
            if st is None:
                selected.append(p)
            elif st.alive and st.available and predicate(st):
                selected.append(p)
        sel = set(selected)

        def compute(p, t):
            return self.iterator(p, t) if p in sel else iter(())

        batch = None
        if self.batch_capable:
            def batch(p, t):  # noqa: E306
                return self.batch(p, t) if p in sel else None
        ds = self._derive(compute, batch, "async_barrier")
        ds.selected_workers = selected
        return ds

    ASYNCbarrier = async_barrier

    # ------------------------------------------------------------------ the wider RDD surface (record path; data-prep scale)
    def union(self, other: "ShardedDataset") -> "ShardedDataset":
        """``rdd.union``: partitions of ``self`` followed by the partitions of ``other`` (no shuffle)."""
        n1 = self.num_partitions
        a, b = self, other
        ds = ShardedDataset(self.scheduler, n1 + other.num_partitions,
                            lambda p, t: a.iterator(p, t) if p < n1 else b.iterator(p - n1, t), None, None, "union")
        ds._union_parents = (a, b)  # type: ignore[attr-defined]  (keeps both lineages alive / picklable)
        return ds

    def zip(self, other: "ShardedDataset") -> "ShardedDataset":
        """``rdd.zip``: pairs records position by position; both sides need the same partitioning."""
        if other.num_partitions != self.num_partitions:
            raise ValueError("zip: both datasets must have the same number of partitions")

        def compute(p, t):
            xs, ys = list(self.iterator(p, t)), list(other.iterator(p, t))
            if len(xs) != len(ys):
                raise ValueError(f"zip: partition {p} has {len(xs)} vs {len(ys)} records")
            return iter(list(zip(xs, ys)))
        ds = self._derive(compute, None, "zip")
        ds._zip_other = other  # type: ignore[attr-defined]
        return ds

    def key_by(self, f) -> "ShardedDataset":
        return self.map(lambda r: (f(r), r))

    def keys(self) -> "ShardedDataset":
        return self.map(lambda kv: kv[0])

    def values(self) -> "ShardedDataset":
        return self.map(lambda kv: kv[1])

    def map_values(self, f) -> "ShardedDataset":
        return self.map(lambda kv: (kv[0], f(kv[1])))

    def flat_map_values(self, f) -> "ShardedDataset":
        """``rdd.flatMapValues`` (PairRDDFunctions.scala:768): keys are kept, every value expands to zero or more values."""
        return self.flat_map(lambda kv: [(kv[0], v) for v in f(kv[1])])

    def group_by(self, f, num_partitions: Optional[int] = None) -> "ShardedDataset":
        """``rdd.groupBy(f)`` = ``map(x => (f(x), x)).groupByKey()`` (RDD.scala:685-705)."""
        return self.key_by(f).group_by_key(num_partitions)

    def sample_by_key(self, with_replacement: bool, fractions: dict, seed: Optional[int] = None) -> "ShardedDataset":
        """``rdd.sampleByKey`` (stratified sampling, PairRDDFunctions.scala:262-284): every key has its own sampling rate;
        Bernoulli per record without replacement, Poisson counts with replacement.  Deterministic in ``(seed, partition,
        position)`` like the other samplers here (a retried task reproduces its sample)."""
        if seed is None:
            seed = int.from_bytes(np.random.bytes(8), "little")
        seed = int(seed)
        missing = object()

        def compute(p, t):
            import random
            rng = random.Random((seed * 0x9E3779B97F4A7C15 + p) & ((1 << 64) - 1))
            for kv in self.iterator(p, t):
                f = fractions.get(kv[0], missing)
                if f is missing:
                    raise KeyError(f"sample_by_key: no sampling fraction for key {kv[0]!r}")
                if with_replacement:
                    # Poisson(f) copies, Knuth's method (f is a sampling rate: small)
                    L, k, prod = math.exp(-float(f)), 0, rng.random()
                    while prod > L:
                        k += 1
                        prod *= rng.random()
                    for _ in range(k):
                        yield kv
                elif rng.random() < f:
                    yield kv
        return self._derive(compute, None, "sample_by_key")

    def foreach(self, f) -> None:
        self._run(lambda t, it: [f(r) for r in it] and None)

    def _shuffle(self, num_partitions: Optional[int], map_side, partitioner=None, aggregator=None) -> "ShardedDataset":
        """Hash (or custom-partitioner) shuffle.

        In-process masters: every map task returns its output bucketed by target partition and the driver regroups — the
        executors share the driver's memory, nothing is copied twice.  Cluster masters (an :class:`AsyncEnv` exists): a real
        shuffle (SortShuffleManager / MapOutputTracker / ShuffleBlockFetcherIterator) — map tasks push their
        records through a spillable :class:`~asyncframework_b200.storage.external_sorter.ExternalSorter` (partition, optional
        map-side combine with ``aggregator``; sorted runs go to disk once the estimated footprint passes
        ``spark.shuffle.spill.memoryThreshold``, ExternalSorter.scala:89), then serialise + compress ONE partition's bucket at
        a time into a ``shuffle_<id>_<map>_<reduce>`` block of THEIR BlockManager and return only the sizes (``MapStatus``);
        the driver keeps that table (``MapOutputTracker``) and ships it inside the reduce-side lineage; reduce tasks fetch
        their column of blocks from the executors that hold them over the bulk chunk path (local blocks are read directly).
        A block that has vanished with its executor is recomputed from lineage by the reducer (fetch-failure recovery)."""
        P = int(num_partitions or self.num_partitions)
        part_of = partitioner or (lambda k: _stable_hash(k) % P)
        env = getattr(self.scheduler, "env", None)

        def bucketize(it):
            buckets: List[list] = [[] for _ in range(P)]
            if aggregator is None:
                for k, v in map_side(it):
                    buckets[part_of(k)].append((k, v))
                return buckets
            acc: dict = {}
            for k, v in map_side(it):
                acc[k] = aggregator.merge_value(acc[k], v) if k in acc else aggregator.create_combiner(v)
            for k, c in acc.items():
                buckets[part_of(k)].append((k, c))
            return buckets

        if env is None:
            outs = self._run(lambda t, it: bucketize(it))
            merged = [[kv for b in outs for kv in b[q]] for q in range(P)]
            blk = SourceBlock(merged)
            return ShardedDataset(self.scheduler, P, lambda p, t: iter(blk.data[p]), None, None, "shuffled")

        sid = next(_shuffle_ids)
        parent = self

        def map_task(t, it):
            from ..core.env import AsyncEnv
            from ..storage.block_manager import StorageLevel
            from ..storage.external_sorter import ExternalSorter, SpillConfig
            e = AsyncEnv.get()
            sizes = [0] * P
            with ExternalSorter(P, part_of, aggregator, config=SpillConfig.from_limits(e.limits)) as sorter:
                sorter.insert_all(map_side(it))
                for q, records in sorter.partitioned_iterator():
                    b = list(records)                         # one bucket at a time (the sorter streams spilled runs)
                    if not b:
                        continue
                    data = e.codec.compress(e.serializer.dumps(b))
                    e.block_manager.put_bytes(f"shuffle_{sid}_{t.partition_id}_{q}", data, StorageLevel.MEMORY_AND_DISK_SER)
                    sizes[q] = len(data)
                _note_spill(t, sorter.metrics())
            return sizes
        statuses = self._run(map_task)                       # MapOutputTracker: statuses[m][q] = bytes of block (m, q)
        num_maps = self.num_partitions

        def reduce_side(p, t):
            from ..core.env import AsyncEnv
            e = AsyncEnv.get()
            for m in range(num_maps):
                if statuses[m][p] == 0:
                    continue
                data = e.block_manager.get_bytes(f"shuffle_{sid}_{m}_{p}") if e is not None else None
                if data is None:                              # FetchFailed: the map output died with its executor
                    yield from bucketize(parent.iterator(m, t))[p]
                    continue
                yield from e.serializer.loads(e.codec.decompress(data))
        ds = ShardedDataset(self.scheduler, P, reduce_side, None, None, f"shuffled[{sid}]")
        ds.shuffle_id = sid
        ds.map_output_bytes = [sum(s) for s in statuses]
        return ds

    def remove_shuffle(self) -> int:
        """Drop this dataset's shuffle blocks on every executor (``ContextCleaner.doCleanupShuffle``)."""
        sid, env = getattr(self, "shuffle_id", None), getattr(self.scheduler, "env", None)
        if sid is None or env is None:
            return 0
        n = 0
        for st in env.bm_master.storage_status().values():
            for b in st["blocks"]:
                if b.startswith(f"shuffle_{sid}_"):
                    n += env.bm_master.remove_block(b)
        return n

    def reduce_by_key(self, f, num_partitions: Optional[int] = None) -> "ShardedDataset":
        """``rdd.reduceByKey`` with map-side combining."""
        return self.combine_by_key(lambda v: v, f, f, num_partitions, _name="reduce_by_key")

    def group_by_key(self, num_partitions: Optional[int] = None) -> "ShardedDataset":
        """``rdd.groupByKey``: no map-side combine (it would not shrink the data, PairRDDFunctions.scala:501-511); the reduce
        side groups through a spillable map (ExternalAppendOnlyMap)."""
        from ..storage.external_sorter import Aggregator, external_combine
        sh = self._shuffle(num_partitions, lambda it: it)

        def compute(p, t):
            return external_combine(sh.iterator(p, t), Aggregator.grouping(), _spill_config(), on_metrics=lambda m: _note_spill(t, m))
        return sh._derive(compute, None, "group_by_key")

    def distinct(self, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.map(lambda r: (r, None)).reduce_by_key(lambda a, b: a, num_partitions).keys()

    def count_by_key(self) -> dict:
        return dict(self.map_values(lambda v: 1).reduce_by_key(lambda a, b: a + b).collect())

    def count_by_value(self) -> dict:
        return dict(self.map(lambda r: (r, 1)).reduce_by_key(lambda a, b: a + b).collect())

    def sort_by(self, key=lambda r: r, ascending: bool = True, num_partitions: Optional[int] = None) -> "ShardedDataset":
        """``rdd.sortBy`` = ``keyBy(f).sortByKey().values`` (RDD.scala:622-628): range-partitioned, spillable, nothing is
        collected on the driver except the key sample of the range partitioner."""
        return self.map(lambda r: (key(r), r)).sort_by_key(ascending, num_partitions).values()

    def top(self, n: int, key=lambda r: r) -> list:
        import heapq
        parts = self._run(lambda t, it: heapq.nlargest(n, it, key=key))
        return heapq.nlargest(n, [r for part in parts for r in part], key=key)

    def take_sample(self, with_replacement: bool, num: int, seed: int = 42) -> list:
        import random
        recs = self.collect()
        rng = random.Random(seed)
        return [rng.choice(recs) for _ in range(num)] if with_replacement else rng.sample(recs, min(num, len(recs)))

    def sum(self):
        return self.aggregate(lambda: 0, lambda a, x: a + x, lambda a, b: a + b)

    def max(self, key=lambda r: r):
        return self.reduce(lambda a, b: a if key(a) >= key(b) else b)

    def min(self, key=lambda r: r):
        return self.reduce(lambda a, b: a if key(a) <= key(b) else b)

    def mean(self) -> float:
        s, n = self.aggregate(lambda: (0.0, 0), lambda a, x: (a[0] + x, a[1] + 1), lambda a, b: (a[0] + b[0], a[1] + b[1]))
        if n == 0:
            raise ValueError("empty collection")
        return s / n

    def coalesce(self, num_partitions: int, shuffle: bool = True) -> "ShardedDataset":
        """``rdd.coalesce(n, shuffle = true)`` == repartition (the form the reference uses, RDD.scala:420-475).  With
        ``shuffle=False`` the parent partitions are merged locally into ``n`` contiguous groups (narrow dependency: no data
        leaves its executor group, like ``CoalescedRDD``); asking for more partitions than exist is then a no-op."""
        if shuffle:
            return self.repartition(num_partitions)
        n = max(1, min(int(num_partitions), self.num_partitions))
        P = self.num_partitions
        groups = [list(range(P * g // n, P * (g + 1) // n)) for g in range(n)]
        parent = self

        def compute(p, t):
            return itertools.chain.from_iterable(parent.iterator(q, t) for q in groups[p])
        return ShardedDataset(self.scheduler, n, compute, None, self, "coalesce")

    def zip_partitions(self, other: "ShardedDataset", f) -> "ShardedDataset":
        """``rdd.zipPartitions(other)(f)``: ``f(iterator_a, iterator_b) -> iterator`` per co-indexed partition pair."""
        if other.num_partitions != self.num_partitions:
            raise ValueError("zip_partitions: both datasets must have the same number of partitions")
        ds = self._derive(lambda p, t: f(self.iterator(p, t), other.iterator(p, t)), None, "zip_partitions")
        ds._zip_other = other  # type: ignore[attr-defined]
        return ds

    def pipe(self, command, env: Optional[dict] = None) -> "ShardedDataset":
        """``rdd.pipe(cmd)``: every partition is streamed line by line through an external process; its stdout lines are the
        result (PipedRDD).  ``command`` is a string (split like a shell word list) or an argv list."""
        import shlex
        import subprocess
        argv = shlex.split(command) if isinstance(command, str) else list(command)

        def compute(p, t):
            data = "".join(str(r) + "\n" for r in self.iterator(p, t))
            res = subprocess.run(argv, input=data, capture_output=True, text=True, env={**os.environ, **(env or {})})
            if res.returncode != 0:
                raise RuntimeError(f"Subprocess exited with status {res.returncode}. Command ran: {' '.join(argv)}: {res.stderr[-300:]}")
            return iter(res.stdout.splitlines())
        return self._derive(compute, None, "pipe")

    def zip_with_unique_id(self) -> "ShardedDataset":
        """``rdd.zipWithUniqueId``: ids k, k+P, k+2P, ... in partition k — unique without the counting job zip_with_index needs."""
        P = self.num_partitions
        return self._derive(lambda p, t: ((r, p + i * P) for i, r in enumerate(self.iterator(p, t))), None, "zip_with_unique_id")

    # ------------------------------------------------------------------ more pair / set operations (PairRDDFunctions)
    def partition_by(self, num_partitions: int, partitioner=None) -> "ShardedDataset":
        """``rdd.partitionBy``: ``partitioner(key) -> partition`` (default: the process-independent hash)."""
        return self._shuffle(num_partitions, lambda it: it, partitioner)

    def combine_by_key(self, create, merge_value, merge_combiners, num_partitions: Optional[int] = None,
                       _name: str = "combine_by_key") -> "ShardedDataset":
        """``combineByKey`` with map-side combining — the primitive under reduce/aggregate/fold-ByKey.  Both sides combine
        through a spillable map (``Aggregator.combineValuesByKey`` / ``combineCombinersByKey``)."""
        from ..storage.external_sorter import Aggregator, external_combine
        agg = Aggregator(create, merge_value, merge_combiners)
        sh = self._shuffle(num_partitions, lambda it: it, aggregator=agg)

        def compute(p, t):
            return external_combine(sh.iterator(p, t), agg, _spill_config(), values_are_combiners=True,
                                    on_metrics=lambda m: _note_spill(t, m))
        return sh._derive(compute, None, _name)

    def aggregate_by_key(self, zero_factory, seq_op, comb_op, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.combine_by_key(lambda v: seq_op(zero_factory(), v), seq_op, comb_op, num_partitions)

    def fold_by_key(self, zero_factory, f, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.combine_by_key(lambda v: f(zero_factory(), v), f, f, num_partitions)

    def cogroup(self, other: "ShardedDataset", num_partitions: Optional[int] = None) -> "ShardedDataset":
        """``(k, (values of self, values of other))`` for every key of either side."""
        P = int(num_partitions or max(self.num_partitions, other.num_partitions))
        tagged = self.map(lambda kv: (kv[0], (0, kv[1]))).union(other.map(lambda kv: (kv[0], (1, kv[1]))))
        sh = tagged._shuffle(P, lambda it: it)

        def mv(c, sv):
            c[sv[0]].append(sv[1])
            return c

        def mc(a, b):
            a[0].extend(b[0])
            a[1].extend(b[1])
            return a
        from ..storage.external_sorter import Aggregator, external_combine
        agg = Aggregator(lambda sv: mv(([], []), sv), mv, mc)

        def compute(p, t):
            return external_combine(sh.iterator(p, t), agg, _spill_config(), on_metrics=lambda m: _note_spill(t, m))
        return sh._derive(compute, None, "cogroup")

    def join(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.cogroup(other, num_partitions).flat_map(lambda kv: [(kv[0], (a, b)) for a in kv[1][0] for b in kv[1][1]])

    def left_outer_join(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.cogroup(other, num_partitions).flat_map(
            lambda kv: [(kv[0], (a, b)) for a in kv[1][0] for b in (kv[1][1] or [None])])

    def right_outer_join(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.cogroup(other, num_partitions).flat_map(
            lambda kv: [(kv[0], (a, b)) for b in kv[1][1] for a in (kv[1][0] or [None])])

    def full_outer_join(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.cogroup(other, num_partitions).flat_map(
            lambda kv: [(kv[0], (a, b)) for a in (kv[1][0] or [None]) for b in (kv[1][1] or [None])])

    def subtract_by_key(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.cogroup(other, num_partitions).flat_map(lambda kv: [(kv[0], a) for a in kv[1][0]] if not kv[1][1] else [])

    def subtract(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.map(lambda r: (r, None)).subtract_by_key(other.map(lambda r: (r, None)), num_partitions).keys()

    def intersection(self, other, num_partitions: Optional[int] = None) -> "ShardedDataset":
        return self.map(lambda r: (r, None)).cogroup(other.map(lambda r: (r, None)), num_partitions) \
            .filter(lambda kv: kv[1][0] and kv[1][1]).keys()

    def cartesian(self, other: "ShardedDataset") -> "ShardedDataset":
        """``rdd.cartesian``: partition (i, j) pairs partition i of self with partition j of other."""
        n2 = other.num_partitions
        left, right = self, other

        def compute(p, t):
            rs = list(right.iterator(p % n2, t))
            return ((a, b) for a in left.iterator(p // n2, t) for b in rs)
        return ShardedDataset(self.scheduler, self.num_partitions * n2, compute, None, None, "cartesian")

    def sort_by_key(self, ascending: bool = True, num_partitions: Optional[int] = None, key=lambda k: k) -> "ShardedDataset":
        """``sortByKey`` with a **range partitioner** built from a sample of the keys (RangePartitioner.sketch): partition q
        holds keys between boundaries q-1 and q, each partition is sorted locally, so reading the partitions in order is the
        total order — no driver-side sort of the data."""
        P = int(num_partitions or self.num_partitions)
        partitioner = RangePartitioner(P, self, ascending, key=key)
        sh = self._shuffle(P, lambda it: it, partitioner)
        from ..storage.external_sorter import external_sort

        def compute(p, t):                                     # spillable local sort (ExternalSorter with an ordering)
            return external_sort(sh.iterator(p, t), key, not ascending, _spill_config(), on_metrics=lambda m: _note_spill(t, m))
        return sh._derive(compute, None, "sort_by_key")

    def repartition_and_sort_within_partitions(self, num_partitions: Optional[int] = None, partitioner=None, ascending: bool = True,
                                               key=lambda k: k) -> "ShardedDataset":
        """``repartitionAndSortWithinPartitions`` (OrderedRDDFunctions.scala:71-79): shuffle by the given partitioner (hash by
        default) and sort every output partition by key — the sort happens in the spillable reduce-side sorter, not after it."""
        from ..storage.external_sorter import external_sort
        sh = self._shuffle(num_partitions, lambda it: it, partitioner)

        def compute(p, t):
            return external_sort(sh.iterator(p, t), key, not ascending, _spill_config(), on_metrics=lambda m: _note_spill(t, m))
        return sh._derive(compute, None, "repartition_and_sort_within_partitions")

    def tree_reduce(self, f, depth: int = 2):
        """``rdd.treeReduce``: reduce inside the partitions, then combine the partials in a ``depth``-level tree."""
        def seq(acc, x):
            return x if acc is _EMPTY else f(acc, x)

        def comb(a, b):
            return b if a is _EMPTY else (a if b is _EMPTY else f(a, b))
        out = self.tree_aggregate(lambda: _EMPTY, seq, comb, depth)
        if out is _EMPTY:
            raise ValueError("empty collection")
        return out

    def count_approx_distinct(self, relative_sd: float = 0.05) -> int:
        """``rdd.countApproxDistinct`` with HyperLogLog (one register array per partition, merged by max): the standard error is
        ``1.04 / sqrt(m)``, so ``m = 2^ceil(log2((1.04 / relative_sd)^2))`` registers; small-range (linear counting) correction."""
        p_bits = max(4, min(18, math.ceil(math.log2((1.04 / max(relative_sd, 1e-4)) ** 2))))
        m = 1 << p_bits

        def seq(reg, x):
            h = _sampler.mix64_int(_stable_hash(x) * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019)
            idx = h >> (64 - p_bits)
            rest = (h << p_bits) & ((1 << 64) - 1)
            rank = (64 - p_bits + 1) if rest == 0 else (64 - rest.bit_length() + 1)
            if rank > reg[idx]:
                reg[idx] = rank
            return reg
        reg = self.aggregate(lambda: np.zeros(m, dtype=np.uint8), seq, np.maximum)
        alpha = 0.7213 / (1 + 1.079 / m) if m >= 128 else {16: 0.673, 32: 0.697, 64: 0.709}[m]
        est = alpha * m * m / float(np.sum(np.power(2.0, -reg.astype(np.float64))))
        zeros = int((reg == 0).sum())
        if est <= 2.5 * m and zeros:
            est = m * math.log(m / zeros)
        return int(round(est))

    def id(self) -> int:
        """``rdd.id``: the unique id blocks of this dataset are stored under (``rdd_<id>_<partition>``)."""
        return self.ds_id

    def lookup(self, k) -> list:
        return self.filter(lambda kv: kv[0] == k).values().collect()

    def collect_as_map(self) -> dict:
        """``collectAsMap`` (PairRDDFunctions.scala:742; the samples build their partition-size table with it)."""
        return dict(self.collect())

    def fold(self, zero_factory, f):
        return self.aggregate(zero_factory, f, f)

    def take_ordered(self, n: int, key=lambda r: r) -> list:
        import heapq
        parts = self._run(lambda t, it: heapq.nsmallest(n, it, key=key))
        return heapq.nsmallest(n, [r for part in parts for r in part], key=key)

    def is_empty(self) -> bool:
        return not any(self._run(lambda t, it: any(True for _ in it)))

    def stats(self) -> dict:
        """``rdd.stats()`` (StatCounter): count, mean, variance (population), stdev, min, max in one pass, merged with the
        numerically stable parallel formula."""
        def seq(a, x):
            n, mean, m2, lo, hi = a
            n += 1
            d = x - mean
            mean += d / n
            m2 += d * (x - mean)
            return (n, mean, m2, min(lo, x), max(hi, x))

        def comb(a, b):
            if a[0] == 0:
                return b
            if b[0] == 0:
                return a
            n = a[0] + b[0]
            d = b[1] - a[1]
            return (n, a[1] + d * b[0] / n, a[2] + b[2] + d * d * a[0] * b[0] / n, min(a[3], b[3]), max(a[4], b[4]))
        n, mean, m2, lo, hi = self.aggregate(lambda: (0, 0.0, 0.0, float("inf"), float("-inf")), seq, comb)
        var = m2 / n if n else float("nan")
        return {"count": n, "mean": mean if n else float("nan"), "variance": var, "stdev": var ** 0.5 if n else float("nan"),
                "min": lo, "max": hi, "sample_variance": m2 / (n - 1) if n > 1 else float("nan")}

    def variance(self) -> float:
        return self.stats()["variance"]

    def stdev(self) -> float:
        return self.stats()["stdev"]

    def sample_variance(self) -> float:
        """``rdd.sampleVariance()``: the unbiased estimator (divides by n - 1; StatCounter.sampleVariance)."""
        st = self.stats()
        n = st["count"]
        return float("nan") if n <= 1 else st["variance"] * n / (n - 1)

    def sample_stdev(self) -> float:
        return math.sqrt(self.sample_variance())

    def count_approx(self, timeout: float, confidence: float = 0.95) -> dict:
        """``rdd.countApprox(timeout, confidence)`` (RDD.scala:1185-1198, partial/CountEvaluator.scala): start the counting job,
        wait at most ``timeout`` seconds and answer from the partitions that have finished — ``{"mean", "low", "high",
        "confidence", "final", "partitions_done"}``.  With every partition in, the answer is exact (``low == high``); with a
        fraction p of them, the total is extrapolated as ``sum / p`` with a normal interval of variance ``sum (1 - p) / p^2``
        (the evaluator's large-count branch).  The job keeps running in the background; its late results are ignored."""
        import statistics
        import threading
        total = self.num_partitions
        lock = threading.Lock()
        got: dict = {}

        def handler(i, r):
            with lock:
                got[i] = int(r)
        waiter = self.scheduler.submit_job(self, lambda t, it: sum(1 for _ in it), None, handler)
        try:
            waiter.wait(max(0.0, float(timeout)))
        except Exception:  # noqa: BLE001  (a failed partition simply does not contribute; count() is the strict form)
            pass
        with lock:
            done, part_sum = len(got), sum(got.values())
        if done == total:
            return {"mean": float(part_sum), "low": float(part_sum), "high": float(part_sum), "confidence": 1.0, "final": True,
                    "partitions_done": done}
        if done == 0:
            return {"mean": 0.0, "low": 0.0, "high": float("inf"), "confidence": float(confidence), "final": False, "partitions_done": 0}
        p = done / total
        mean = part_sum / p
        sd = math.sqrt(part_sum * (1.0 - p)) / p
        z = statistics.NormalDist().inv_cdf(0.5 + min(max(confidence, 0.0), 0.999999) / 2.0)
        return {"mean": mean, "low": max(float(part_sum), mean - z * sd), "high": mean + z * sd, "confidence": float(confidence),
                "final": False, "partitions_done": done}

    def histogram(self, buckets):
        """``rdd.histogram(n | boundaries)`` -> (boundaries, counts); the last bucket is closed on the right."""
        import bisect
        if isinstance(buckets, int):
            st = self.stats()
            lo, hi = st["min"], st["max"]
            bounds = [lo + (hi - lo) * i / buckets for i in range(buckets + 1)] if hi > lo else [lo, lo]
        else:
            bounds = list(buckets)
        nb = len(bounds) - 1

        def seq(acc, x):
            if bounds[0] <= x <= bounds[-1]:
                acc[min(bisect.bisect_right(bounds, x) - 1, nb - 1)] += 1
            return acc
        counts = self.aggregate(lambda: [0] * nb, seq, lambda a, b: [x + y for x, y in zip(a, b)])
        return bounds, counts

    def random_split(self, weights: Sequence[float], seed: int = 17) -> List["ShardedDataset"]:
        """``rdd.randomSplit``: disjoint samples whose sizes follow ``weights`` (one hash draw per record)."""
        tot = float(sum(weights))
        cuts = [sum(weights[:i + 1]) / tot for i in range(len(weights))]
        indexed = self.zip_with_unique_id()

        def pick(lo, hi):
            return indexed.filter(lambda ri: lo <= (_stable_hash((seed, ri[1])) % 1_000_003) / 1_000_003 < hi).map(lambda ri: ri[0])
        return [pick(0.0 if i == 0 else cuts[i - 1], cuts[i] if i < len(cuts) - 1 else 1.1) for i in range(len(cuts))]

    def save_as_text_file(self, directory: str) -> None:
        """``saveAsTextFile``: one ``part-<p>`` file per partition + ``_SUCCESS`` (readable back with ``sc.text_file``)."""
        os.makedirs(directory, exist_ok=True)

        def write(t, it):
            with open(os.path.join(directory, f"part-{t.partition_id:05d}"), "w") as f:
                for r in it:
                    f.write(str(r) + "\n")
            return t.partition_id
        self._run(write)
        open(os.path.join(directory, "_SUCCESS"), "w").close()

    def save_as_pickle_file(self, directory: str) -> None:
        """``saveAsPickleFile`` / ``saveAsObjectFile``: one pickled list per partition (``sc.pickle_file`` reads it back)."""
        import pickle as _pickle
        os.makedirs(directory, exist_ok=True)

        def write(t, it):
            with open(os.path.join(directory, f"part-{t.partition_id:05d}"), "wb") as f:
                _pickle.dump(list(it), f, protocol=_pickle.HIGHEST_PROTOCOL)
            return t.partition_id
        self._run(write)
        open(os.path.join(directory, "_SUCCESS"), "w").close()

    def to_local_iterator(self):
        """``toLocalIterator``: one partition at a time (one job each), so the driver never holds more than a partition."""
        for p in range(self.num_partitions):
            yield from self._run(lambda t, it: list(it), [p])[0]

    def set_name(self, name: str) -> "ShardedDataset":
        self.name = name
        return self

    def to_debug_string(self) -> str:
        """``rdd.toDebugString``: the lineage, one line per ancestor."""
        lines, ds, depth = [], self, 0
        while ds is not None:
            flags = [f"{ds.num_partitions} partitions"] + (["cached"] if ds._want_cache else []) + (["columnar"] if ds.batch_capable else [])
            lines.append(" " * (2 * depth) + ("+- " if depth else "") + f"{ds.name} [{', '.join(flags)}]")
            ds, depth = ds.parent, depth + 1
        return "\n".join(lines)

    def unpersist(self) -> "ShardedDataset":
        """Drop the cached partitions everywhere (driver lists, executor dicts, executor BlockManagers)."""
        self._want_cache = False
        with self._cache_lock:
            self._cached_batches = None
            self._cached_records = None
        for k in [k for k in list(_BLOCK_STORE) if k[0] == self.ds_id]:
            _BLOCK_STORE.pop(k, None)
        env = getattr(self.scheduler, "env", None)
        if env is not None:
            env.bm_master.remove_rdd(self.ds_id)
        return self

    # ------------------------------------------------------------------ blocking actions
    def _run(self, func, partitions=None):
        # blocking for THIS job only; the scheduler-wide mode (sc.set_mode) is left alone, so actions issued from several
        # driver threads cannot see each other's temporary state
        return self.scheduler.run_job(self, func, partitions, blocking=True)

    def collect(self) -> list:
        out = self._run(lambda t, it: list(it))
        return [r for part in out for r in part]

    def count(self) -> int:
        return sum(self.partition_sizes())

    def partition_sizes(self) -> List[int]:
        """``mapPartitionsWithIndex{(i,rows)=>(i,rows.size)}`` histogram (SparkASGDThread.scala:77)."""
        if self.batch_capable:
            def f(t, it):
                b = self.batch(t.partition_id, t)
                return 0 if b is None else b.num_rows
        else:
            def f(t, it):
                return sum(1 for _ in it)
        return [int(x) for x in self._run(f)]

    def first(self):
        for p in range(self.num_partitions):
            for r in self.iterator(p):
                return r
        raise ValueError("empty dataset")

    def take(self, n: int) -> list:
        out = []
        for p in range(self.num_partitions):
            for r in self.iterator(p):
                out.append(r)
                if len(out) >= n:
                    return out
        return out

    def reduce(self, f: Callable[[Any, Any], Any]):
        def part(t, it):
            it = iter(it)
            try:
                acc = next(it)
            except StopIteration:
                return _EMPTY
            for r in it:
                acc = f(acc, r)
            return acc
        vals = [v for v in self._run(part) if v is not _EMPTY]
        if not vals:
            raise ValueError("empty collection")
        acc = vals[0]
        for v in vals[1:]:
            acc = f(acc, v)
        return acc

    def aggregate(self, zero_factory, seq_op, comb_op):
        """``zero_factory`` is a callable returning a fresh zero (the reference clones the zero
        value per task via serialisation)."""
        zf = zero_factory if callable(zero_factory) else (lambda: _clone(zero_factory))

        def part(t, it):
            acc = zf()
            for r in it:
                acc = seq_op(acc, r)
            return acc
        acc = zf()
        for v in self._run(part):
            acc = comb_op(acc, v)
        return acc

    def tree_aggregate(self, zero_factory, seq_op, comb_op, depth: int = 2):
        """Hierarchical reduce (RDD.scala:1358-1388): partials are combined in ``depth`` levels."""
        zf = zero_factory if callable(zero_factory) else (lambda: _clone(zero_factory))

        def part(t, it):
            acc = zf()
            for r in it:
                acc = seq_op(acc, r)
            return acc
        partials = self._run(part)
        scale = max(int(math.ceil(len(partials) ** (1.0 / max(depth, 1)))), 2)
        while len(partials) > scale + math.ceil(len(partials) / scale):
            groups = max(len(partials) // scale, 1)
            merged = []
            for g in range(groups):
                grp = partials[g::groups]
                acc = grp[0]
                for v in grp[1:]:
                    acc = comb_op(acc, v)
                merged.append(acc)
            partials = merged
        acc = partials[0]
        for v in partials[1:]:
            acc = comb_op(acc, v)
        return acc

    def foreach_partition(self, f) -> None:
        self._run(lambda t, it: f(it))

    # ------------------------------------------------------------------ ASYNC actions
    def _async_submit(self, ctx, part_fn, envelope_ts_only: bool, empty_raises: bool):
        """Shared machinery of async_reduce / async_aggregate (RDD.scala:1087-1171, 1268-1345)."""
        sched = self.scheduler
        selected = list(range(self.num_partitions)) if self.selected_workers is None else list(self.selected_workers)
        if not selected:
            return None   # "if the same w is passed here, we shouldn't run the job again" (:1095-1097)
        start = time.perf_counter() * 1e3
        ts = ctx.current_time
        ctx.last_time = ts
        for p in selected:
            ctx.mark_submitted(p)      # busy until their result returns (:1138-1142)
        sel = set(selected)

        def merge(index: int, result):
            count, value = result
            with ctx.merge_lock:
                if value is not _EMPTY:
                    end = time.perf_counter() * 1e3
                    staleness = ctx.current_time - ts
                    free = ctx.mark_finished(index) if index in sel else True
                    ctx.stat.task_done(index, staleness, end - start)
                    if not free:
                        ctx.stat.set_available(index, False)
                    ctx.add_to_current_time(1)
                    # publish LAST: whoever dequeues the result already sees the ticked clock and the
                    # refreshed STAT entry (the reference enqueues first, RDD.scala:1155-1158)
                    ctx.put(PartialResult(value, staleness, ts, count, index, start, end))
                else:
                    if index in sel:
                        free = ctx.mark_finished(index)
                        ctx.dropped_empty += 1
                        if empty_raises:
                            ctx.stat.set_available(index, free)
                            raise ValueError(f"async_aggregate: selected partition {index} produced no records")
                    else:
                        free = ctx.inflight(index) == 0
                    if free:
                        ctx.stat.set_available(index, True)

        def lost(index: int, exc):   # a failed async task must not strand its worker (quirk B.15)
            if index in sel:
                ctx.mark_finished(index)
            ctx.stat.set_available(index, True)
            from ..core.scheduler import JobCancelled
            if exc is not None and not isinstance(exc, JobCancelled):
                ctx.record_lost(index, exc)

        sched.set_mode(1)   # ASYNCreduce always switches the scheduler to non-blocking (RDD.scala:1167)
        # only the selected partitions do work; running the no-op tasks of the others is pure
        # overhead (the reference runs them and they just flip availability) — skip them.
        return sched.run_job(self, part_fn, selected, merge, lost, blocking=False)

    def async_reduce(self, f: Callable[[Any, Any], Any], ctx):
        """Non-blocking reduce: each selected partition's ``reduceLeft(f)`` lands in ``ctx`` as it
        finishes, tagged with staleness / worker id."""
        record_stat = ctx.record_stat

        def part(t, it):
            it = iter(it)
            try:
                acc = next(it)
            except StopIteration:
                return (0, _EMPTY)
            n = 1
            for r in it:
                acc = f(acc, r)
                n += 1
            return (n if record_stat else -(2**31), acc)
        return self._async_submit(ctx, part, envelope_ts_only=False, empty_raises=False)

    def async_aggregate(self, zero_factory, seq_op, comb_op, ctx, strict_empty: bool = False):
        """Non-blocking aggregate; the envelope carries the submit timestamp (and, unlike the
        reference, also the staleness).  ``strict_empty`` reproduces the reference's
        ``None.get`` failure on an empty sampled partition (RDD.scala:1333, quirk B.5)."""
        zf = zero_factory if callable(zero_factory) else (lambda: _clone(zero_factory))
        record_stat = ctx.record_stat

        def part(t, it):
            acc = zf()
            n = 0
            for r in it:
                acc = seq_op(acc, r)
                n += 1
            if n == 0:
                return (0, _EMPTY)
            return (n if record_stat else -(2**31), acc)
        return self._async_submit(ctx, part, envelope_ts_only=True, empty_raises=strict_empty)

    def async_grad(self, loss, w_bcast, ctx, history=None):
        """Fused fast path of the ASYNC drivers: for every selected (and already sampled) partition compute the
        mini-batch gradient of ``loss`` at ``w_bcast.value`` in ONE pass over the shard — the fused CUDA kernel on HBM-resident
        shards, the vectorised torch path on CPU shards — and deliver ``[g, loss_sum, n]`` (``[g, loss_sum, n, coeff, rows]`` with
        ``history``: the SAGA-corrected gradient Σ (r_i − hist_i) x_i, the new residuals and their local row positions) to ``ctx`` as each worker finishes, tagged with staleness / worker id.
        ``loss`` is a name (``"logistic"`` …) or a :class:`Gradient`; ``history`` maps partition -> per-row tensor."""
        from ..optim.gradient import Gradient, make_gradient
        grad = loss if isinstance(loss, Gradient) else make_gradient(loss)

        def fn(s, w_bcast=w_bcast, grad=grad, history=history):
            from ..algo.common import partition_gradient
            h = rows = None
            if history is not None:
                hp = history[s.partition]
                rows = s.meta.get("rows")                      # local positions of the sampled rows (set by ``sample``)
                h = hp if rows is None else hp[rows.to(hp.device)]
            g, l, n, coeff = partition_gradient(grad, s, w_bcast.value, None if h is None else h.to(s.device))
            return [g, float(l), int(n)] if history is None else [g, float(l), int(n), coeff, rows]

        def merge(x, y):                              # a partition is ONE shard, so this only runs for multi-shard partitions
            out = [x[0].add_(y[0]), x[1] + y[1], x[2] + y[2]]
            return out if len(x) == 3 else out + [torch.cat([x[3], y[3]]), None]
        return self.map_batch(fn).async_reduce(merge, ctx)

    ASYNCreduce = async_reduce
    ASYNCaggregate = async_aggregate

    # ------------------------------------------------------------------ misc
    def get_num_partitions(self) -> int:
        return self.num_partitions

    def shards(self) -> List[Optional[Shard]]:
        return [self.batch(p) for p in range(self.num_partitions)]

    def __repr__(self) -> str:
        return f"ShardedDataset({self.name}, P={self.num_partitions}, batch={self.batch_capable})"


def _clone(z):
    if isinstance(z, torch.Tensor):
        return z.clone()
    if isinstance(z, DenseVector):
        return z.copy()
    if isinstance(z, tuple):
        return tuple(_clone(v) for v in z)
    if isinstance(z, list):
        return [_clone(v) for v in z]
    if isinstance(z, dict):
        return {k: _clone(v) for k, v in z.items()}
    return z
