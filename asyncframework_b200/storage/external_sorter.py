"""Spillable sort / aggregation for the shuffle.

Reference behaviour: ``util/collection/ExternalSorter.scala:89`` (map side: partition -> optional combine -> optional
key order, spilling sorted runs to disk when the in-memory collection outgrows its share of execution memory, merging the
runs per partition when the map output is written) and ``ExternalAppendOnlyMap.scala`` (reduce side: the same idea for
the fetched stream).  The design here is its own:

* records live in ONE python container per sorter — a ``dict`` keyed by ``(partition, key)`` when a combiner is given
  (combining happens on insert, like ``PartitionedAppendOnlyMap.changeValue``), a flat list otherwise;
* the footprint is *estimated* (``SizeEstimator`` cannot exist for arbitrary python objects either): every
  ``sample_every`` inserts the serialized size of the newest record is measured and a running mean is kept, so the check
  costs one ``pickle.dumps`` per few hundred records;
* a spill writes ONE file: for every partition in order a sequence of length-prefixed, optionally compressed batches of
  ``batch_size`` records (``spark.shuffle.spill.batchSize``), plus an in-memory index of ``(offset, n_batches)`` per
  partition — reading a partition back streams one batch at a time, so the merge holds one batch per run;
* inside a partition a run is sorted by the caller's key order if there is one, otherwise by a *stable hash* of the key
  (``ExternalSorter.keyComparator``'s trick: any total order works for combining; equal hashes are disambiguated by ``==``
  while merging);
* the merge is ``heapq.merge`` over the runs (+ the in-memory remainder), followed — with a combiner — by a pass that
  folds neighbouring equal keys (equal hash buckets are folded pairwise by equality).

Counters mirror the task metrics the reference reports (``memoryBytesSpilled`` / ``diskBytesSpilled`` /
``peakExecutionMemory``).
"""
from __future__ import annotations

import heapq
import os
import pickle
import struct
import tempfile
import zlib
from dataclasses import dataclass
from typing import Any, Callable, Iterable, Iterator, List, Optional, Tuple


@dataclass
class Aggregator:
    """``Aggregator[K, V, C]`` of the reference (createCombiner / mergeValue / mergeCombiners)."""
    create_combiner: Callable[[Any], Any]
    merge_value: Callable[[Any, Any], Any]
    merge_combiners: Callable[[Any, Any], Any]

    @staticmethod
    def from_reduce(f: Callable[[Any, Any], Any]) -> "Aggregator":
        return Aggregator(lambda v: v, f, f)

    @staticmethod
    def grouping() -> "Aggregator":
        def mv(c, v):
            c.append(v)
            return c

        def mc(a, b):
            a.extend(b)
            return a
        return Aggregator(lambda v: [v], mv, mc)


@dataclass
class SpillConfig:
    memory_threshold: int = 256 << 20        # bytes of estimated in-memory records before a spill
    force_spill_elements: int = 0            # spark.shuffle.spill.numElementsForceSpillThreshold (0 = off)
    batch_size: int = 10000                  # spark.shuffle.spill.batchSize
    compress: bool = True                    # spark.shuffle.spill.compress
    local_dir: Optional[str] = None          # spark.local.dir
    sample_every: int = 256

    @staticmethod
    def from_limits(limits) -> "SpillConfig":
        if limits is None:
            return SpillConfig()
        return SpillConfig(memory_threshold=int(getattr(limits, "shuffle_memory", 256 << 20)),
                           force_spill_elements=int(getattr(limits, "shuffle_spill_elements", 0)),
                           batch_size=int(getattr(limits, "shuffle_spill_batch", 10000)),
                           compress=bool(getattr(limits, "shuffle_spill_compress", True)),
                           local_dir=getattr(limits, "local_dir", None))


# process-wide default for sorters created outside an AsyncEnv (in-process masters); tests override it
DEFAULT_CONFIG = SpillConfig()


def _stable_key_hash(k) -> int:
    from ..data.dataset import _stable_hash
    return _stable_hash(k)


class _Run:
    """One spill file: per-partition (offset, batches) index; ``read(p)`` streams partition p's records."""

    def __init__(self, path: str, index: List[Tuple[int, int]], compress: bool, nbytes: int):
        self.path, self.index, self.compress, self.nbytes = path, index, compress, nbytes

    def read(self, p: int) -> Iterator[tuple]:
        off, nb = self.index[p]
        if nb == 0:
            return
        with open(self.path, "rb") as f:
            f.seek(off)
            for _ in range(nb):
                (n,) = struct.unpack("<Q", f.read(8))
                blob = f.read(n)
                if self.compress:
                    blob = zlib.decompress(blob)
                yield from pickle.loads(blob)

    def delete(self) -> None:
        try:
            os.unlink(self.path)
        except OSError:
            pass


class ExternalSorter:
    """Partition (+ combine) (+ order) a stream of ``(k, v)`` records with bounded memory.

    ``partitioner``: ``k -> int`` in ``[0, num_partitions)``.  ``aggregator``: combine values of equal keys.  ``key``: total
    order on keys inside a partition (``None``: unordered output, or hash order when combining); ``reverse`` flips it."""

    def __init__(self, num_partitions: int, partitioner: Optional[Callable[[Any], int]] = None, aggregator: Optional[Aggregator] = None,
                 key: Optional[Callable[[Any], Any]] = None, reverse: bool = False, config: Optional[SpillConfig] = None):
        self.P = int(num_partitions)
        self.partitioner = partitioner or (lambda k: 0)
        self.agg = aggregator
        self.key, self.reverse = key, reverse
        self.cfg = config or DEFAULT_CONFIG
        self._map: dict = {}
        self._buf: list = []
        self._n = 0                             # entries held in memory (distinct keys when combining)
        self._reads = 0                         # records inserted since the last spill (``elementsRead`` of Spillable.scala)
        self._avg = 64.0                        # running mean of the serialized entry size (the combiner's, when combining)
        self._samples = 0
        self.runs: List[_Run] = []
        self.spill_count = 0
        self.memory_bytes_spilled = 0
        self.disk_bytes_spilled = 0
        self.peak_memory = 0
        self.records_in = 0

    # ---- insert --------------------------------------------------------------------------------------------------
    def _estimate(self) -> int:
        return int(self._n * self._avg * 2.5)    # live python objects weigh a few times their pickled size

    def _sample(self, rec) -> None:
        try:
            sz = len(pickle.dumps(rec, protocol=pickle.HIGHEST_PROTOCOL))
        except Exception:  # noqa: BLE001  (unpicklable records cannot spill either; the error surfaces in spill())
            sz = 64
        self._samples += 1
        self._avg += (sz - self._avg) / min(self._samples, 64)

    def insert_all(self, records: Iterable[Tuple[Any, Any]]) -> "ExternalSorter":
        agg, part = self.agg, self.partitioner
        every = self.cfg.sample_every
        for k, v in records:
            self.records_in += 1
            if agg is not None:
                pk = (part(k), k)
                if pk in self._map:
                    self._map[pk] = agg.merge_value(self._map[pk], v)
                else:
                    self._map[pk] = agg.create_combiner(v)
                    self._n += 1
            else:
                self._buf.append((part(k), k, v))
                self._n += 1
            self._reads += 1
            if self.records_in % every == 1:           # a growing combiner (grouping) is measured as it is NOW
                self._sample((k, self._map[pk]) if agg is not None else (k, v))
            if self._should_spill():
                self.spill()
        return self

    def _should_spill(self) -> bool:
        if self._n == 0:
            return False
        if self.cfg.force_spill_elements and self._reads >= self.cfg.force_spill_elements:
            return True
        if self._reads % 32:
            return False
        est = self._estimate()
        if est > self.peak_memory:
            self.peak_memory = est
        return est > self.cfg.memory_threshold

    # ---- ordering inside a partition ---------------------------------------------------------------------------------
    def _order_of(self, k):
        if self.key is not None:
            return self.key(k)
        return _stable_key_hash(k)

    @property
    def _ordered(self) -> bool:
        return self.key is not None or self.agg is not None

    def _sorted_memory(self) -> List[Tuple[int, Any, Any]]:
        """The in-memory collection as a list of (partition, k, c) sorted by (partition, order)."""
        if self.agg is not None:
            items = [(pk[0], pk[1], c) for pk, c in self._map.items()]
        else:
            items = self._buf
        if self._ordered:
            if self.key is not None and self.reverse:
                # partitions ascending, keys descending: two stable passes
                items = sorted(items, key=lambda r: self.key(r[1]), reverse=True)
                items.sort(key=lambda r: r[0])
            else:
                items = sorted(items, key=lambda r: (r[0], self._order_of(r[1])))
        else:
            items = sorted(items, key=lambda r: r[0])
        return items

    # ---- spill -----------------------------------------------------------------------------------------------------
    def spill(self) -> None:
        if self._n == 0:
            return
        items = self._sorted_memory()
        d = self.cfg.local_dir or tempfile.gettempdir()
        os.makedirs(d, exist_ok=True)
        fd, path = tempfile.mkstemp(prefix="afb-spill-", suffix=".run", dir=d)
        index: List[Tuple[int, int]] = [(0, 0)] * self.P
        B = max(1, self.cfg.batch_size)
        with os.fdopen(fd, "wb") as f:
            i, n = 0, len(items)
            while i < n:
                p = items[i][0]
                j = i
                while j < n and items[j][0] == p:
                    j += 1
                off, nb = f.tell(), 0
                for s in range(i, j, B):
                    blob = pickle.dumps([(k, c) for (_, k, c) in items[s:min(j, s + B)]], protocol=pickle.HIGHEST_PROTOCOL)
                    if self.cfg.compress:
                        blob = zlib.compress(blob, 1)
                    f.write(struct.pack("<Q", len(blob)))
                    f.write(blob)
                    nb += 1
                index[p] = (off, nb)
                i = j
            nbytes = f.tell()
        self.runs.append(_Run(path, index, self.cfg.compress, nbytes))
        self.spill_count += 1
        self.memory_bytes_spilled += self._estimate()
        self.disk_bytes_spilled += nbytes
        self._map, self._buf, self._n, self._reads = {}, [], 0, 0

    # ---- read-out ----------------------------------------------------------------------------------------------------
    def _combine_sorted(self, it: Iterator[Tuple[Any, Any]]) -> Iterator[Tuple[Any, Any]]:
        """Fold equal keys of a stream sorted by ``_order_of`` (equal orders may hide different keys: bucket by ==)."""
        mc = self.agg.merge_combiners
        cur_order, bucket = None, []                   # bucket: [[k, c], ...] sharing one order value
        have = False
        for k, c in it:
            o = self._order_of(k)
            if have and o == cur_order:
                for kc in bucket:
                    if kc[0] == k:
                        kc[1] = mc(kc[1], c)
                        break
                else:
                    bucket.append([k, c])
                continue
            if have:
                for kc in bucket:
                    yield kc[0], kc[1]
            cur_order, bucket, have = o, [[k, c]], True
        if have:
            for kc in bucket:
                yield kc[0], kc[1]

    def partitioned_iterator(self) -> Iterator[Tuple[int, Iterator[Tuple[Any, Any]]]]:
        """``(partition, iterator of (k, c))`` for every partition in order (ExternalSorter.partitionedIterator)."""
        mem = self._sorted_memory()
        # slice the in-memory list per partition
        bounds, i, n = {}, 0, len(mem)
        while i < n:
            p, j = mem[i][0], i
            while j < n and mem[j][0] == p:
                j += 1
            bounds[p] = (i, j)
            i = j
        for p in range(self.P):
            lo, hi = bounds.get(p, (0, 0))
            streams = [((k, c) for (_, k, c) in mem[lo:hi])] if hi > lo else []
            streams += [r.read(p) for r in self.runs if r.index[p][1]]
            if not streams:
                yield p, iter(())
                continue
            if not self._ordered:
                def chain(ss=streams):
                    for s in ss:
                        yield from s
                yield p, chain()
                continue
            if len(streams) == 1:
                merged = streams[0]
            else:
                merged = heapq.merge(*streams, key=lambda kc: self._order_of(kc[0]), reverse=self.reverse and self.key is not None)
            yield p, (self._combine_sorted(merged) if self.agg is not None and self.runs else merged)

    def iterator(self) -> Iterator[Tuple[Any, Any]]:
        for _, it in self.partitioned_iterator():
            yield from it

    def stop(self) -> None:
        for r in self.runs:
            r.delete()
        self.runs = []
        self._map, self._buf, self._n = {}, [], 0

    def metrics(self) -> dict:
        return {"spills": self.spill_count, "memory_bytes_spilled": self.memory_bytes_spilled,
                "disk_bytes_spilled": self.disk_bytes_spilled, "peak_memory": max(self.peak_memory, self._estimate()),
                "records": self.records_in}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()


def external_sort(records: Iterable[Tuple[Any, Any]], key: Callable[[Any], Any], reverse: bool = False,
                  config: Optional[SpillConfig] = None, on_metrics: Optional[Callable[[dict], None]] = None) -> Iterator[Tuple[Any, Any]]:
    """Sort ``(k, v)`` records by ``key(k)`` with bounded memory (the reduce side of ``sortByKey``)."""
    s = ExternalSorter(1, None, None, key=key, reverse=reverse, config=config)
    try:
        s.insert_all(records)
        yield from s.iterator()
        if on_metrics is not None:
            on_metrics(s.metrics())
    finally:
        s.stop()


def external_combine(records: Iterable[Tuple[Any, Any]], aggregator: Aggregator, config: Optional[SpillConfig] = None,
                     values_are_combiners: bool = False, on_metrics: Optional[Callable[[dict], None]] = None) -> Iterator[Tuple[Any, Any]]:
    """Combine the values of equal keys with bounded memory (``ExternalAppendOnlyMap``).  ``values_are_combiners``: the input
    already holds map-side combiners, so they are merged with ``merge_combiners``."""
    agg = aggregator
    if values_are_combiners:
        agg = Aggregator(lambda c: c, aggregator.merge_combiners, aggregator.merge_combiners)
    s = ExternalSorter(1, None, agg, config=config)
    try:
        s.insert_all(records)
        yield from s.iterator()
        if on_metrics is not None:
            on_metrics(s.metrics())
    finally:
        s.stop()
