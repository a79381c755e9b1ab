"""Per-process block store + driver-side location master (the reference's BlockManager stack, S7).

Reference: ``BlockManager`` (core/storage/BlockManager.scala:118 — ``getLocalBytes`` :573-589, ``getRemoteBytes``,
``putBytes``, ``putSingle``), ``MemoryStore`` (storage/memory/MemoryStore.scala:80, LRU eviction),
``DiskStore`` (storage/DiskStore.scala:42), ``BlockInfoManager`` (storage/BlockInfoManager.scala:115, per-block
read/write locks), ``BlockManagerMaster`` + ``BlockManagerMasterEndpoint`` (storage/BlockManagerMaster.scala:31 —
who holds which block) and ``NettyBlockTransferService`` (network/netty/NettyBlockTransferService.scala:45 —
``fetchBlocks`` over ChunkFetchRequest frames).

What lives here in this framework: broadcast pieces (``broadcast_<id>_piece<i>``), reassembled broadcast values
(``broadcast_<id>``), task results too large for the status message (``taskresult_<tid>``) and cached dataset
partitions of executor processes (``rdd_<dataset>_<partition>``).  Training shards of the fused backend do NOT live
here — they are HBM-resident tensors owned by the worker process (``data/batch.py::Shard``), and parameters live in
symmetric device memory (``ops/round.py::ServerBlock``); this store is the host-side substrate around them.

Blocks are either *serialized* (bytes — what travels between processes) or *deserialized* (a Python object; served
to remote readers by pickling on demand).  Memory is accounted against ``max_memory``; when a put does not fit, the
least-recently-used blocks are evicted — dropped to the disk store when their storage level allows it, otherwise
removed (a cached partition is then recomputed from lineage, like MEMORY_ONLY in the reference).
"""
from __future__ import annotations

import hashlib
import os
import random
import shutil
import tempfile
import threading
import time
import uuid
from collections import OrderedDict
from dataclasses import dataclass
from typing import Any, Callable, Dict, Iterable, List, NamedTuple, Optional, Tuple

import cloudpickle
import pickle

from ..rpc import RpcCallContext, RpcEndpoint, RpcEndpointRef, RpcEnv, RpcError


# ------------------------------------------------------------------------------------------------ ids, levels
def rdd_block_id(dataset_id: int, partition: int, kind: str = "") -> str:
    return f"rdd_{dataset_id}_{partition}{kind}"


def broadcast_block_id(broadcast_id: int, piece: Optional[int] = None) -> str:
    return f"broadcast_{broadcast_id}" if piece is None else f"broadcast_{broadcast_id}_piece{piece}"


def task_result_block_id(task_id: int) -> str:
    return f"taskresult_{task_id}"


@dataclass(frozen=True)
class StorageLevel:
    use_disk: bool
    use_memory: bool
    deserialized: bool
    replication: int = 1

    @property
    def valid(self) -> bool:
        return self.use_disk or self.use_memory


StorageLevel.NONE = StorageLevel(False, False, False)                    # type: ignore[attr-defined]
StorageLevel.MEMORY_ONLY = StorageLevel(False, True, True)               # type: ignore[attr-defined]
StorageLevel.MEMORY_ONLY_SER = StorageLevel(False, True, False)          # type: ignore[attr-defined]
StorageLevel.MEMORY_AND_DISK = StorageLevel(True, True, True)            # type: ignore[attr-defined]
StorageLevel.MEMORY_AND_DISK_SER = StorageLevel(True, True, False)       # type: ignore[attr-defined]
StorageLevel.DISK_ONLY = StorageLevel(True, False, False)                # type: ignore[attr-defined]
StorageLevel.MEMORY_ONLY_2 = StorageLevel(False, True, True, 2)          # type: ignore[attr-defined]


class BlockManagerId(NamedTuple):
    executor_id: str
    host: str
    port: int


def estimate_size(value: Any) -> int:
    """``SizeEstimator.estimate`` for what this framework caches: tensors / arrays by their buffers, shards by their
    tensors, containers recursively (sampled for long lists), anything else by its pickled length."""
    import numpy as np
    import torch
    if isinstance(value, (bytes, bytearray, memoryview)):
        return len(value)
    if isinstance(value, torch.Tensor):
        return value.numel() * value.element_size()
    if isinstance(value, np.ndarray):
        return int(value.nbytes)
    if hasattr(value, "nbytes") and callable(getattr(value, "nbytes")):
        try:
            return int(value.nbytes())
        except Exception:  # noqa: BLE001
            pass
    if hasattr(value, "__dict__") and any(isinstance(v, torch.Tensor) for v in vars(value).values()):
        return 64 + sum(estimate_size(v) for v in vars(value).values() if isinstance(v, (torch.Tensor, np.ndarray)))
    if isinstance(value, (list, tuple)) and len(value) > 64:
        sample = random.Random(len(value)).sample(range(len(value)), 32)
        return 56 + int(len(value) * sum(estimate_size(value[i]) for i in sample) / 32)
    if isinstance(value, (list, tuple, set, frozenset)):
        return 56 + sum(estimate_size(v) for v in value)
    if isinstance(value, dict):
        return 232 + sum(estimate_size(k) + estimate_size(v) for k, v in value.items())
    if isinstance(value, (int, float, bool)) or value is None:
        return 28
    try:
        return len(pickle.dumps(value, protocol=pickle.HIGHEST_PROTOCOL))
    except Exception:  # noqa: BLE001
        return len(cloudpickle.dumps(value))


# ------------------------------------------------------------------------------------------------ locks
class BlockInfoManager:
    """Reader/writer locks per block: many concurrent readers or one writer; ``remove`` needs the write lock.
    (BlockInfoManager.scala:115 — task-attempt bookkeeping is reduced to thread identity.)"""

    def __init__(self):
        self._cv = threading.Condition()
        self._readers: Dict[str, int] = {}
        self._writer: Dict[str, int] = {}

    def lock_for_reading(self, block_id: str, blocking: bool = True, timeout: Optional[float] = None) -> bool:
        with self._cv:
            end = None if timeout is None else time.monotonic() + timeout
            while block_id in self._writer:
                if not blocking:
                    return False
                left = None if end is None else end - time.monotonic()
                if left is not None and left <= 0:
                    return False
                self._cv.wait(left)
            self._readers[block_id] = self._readers.get(block_id, 0) + 1
            return True

    def lock_for_writing(self, block_id: str, blocking: bool = True, timeout: Optional[float] = None) -> bool:
        me = threading.get_ident()
        with self._cv:
            end = None if timeout is None else time.monotonic() + timeout
            while block_id in self._writer or self._readers.get(block_id, 0) > 0:
                if self._writer.get(block_id) == me:
                    raise RuntimeError(f"thread already holds the write lock of {block_id}")
                if not blocking:
                    return False
                left = None if end is None else end - time.monotonic()
                if left is not None and left <= 0:
                    return False
                self._cv.wait(left)
            self._writer[block_id] = me
            return True

    def unlock(self, block_id: str) -> None:
        with self._cv:
            if self._writer.get(block_id) == threading.get_ident():
                del self._writer[block_id]
            elif self._readers.get(block_id, 0) > 0:
                self._readers[block_id] -= 1
                if self._readers[block_id] == 0:
                    del self._readers[block_id]
            else:
                raise RuntimeError(f"unlock of {block_id} without a lock")
            self._cv.notify_all()

    def downgrade(self, block_id: str) -> None:
        """write lock -> read lock (a put that wants to keep reading the block it just wrote)."""
        with self._cv:
            assert self._writer.get(block_id) == threading.get_ident()
            del self._writer[block_id]
            self._readers[block_id] = self._readers.get(block_id, 0) + 1
            self._cv.notify_all()

    def is_locked(self, block_id: str) -> bool:
        with self._cv:
            return block_id in self._writer or self._readers.get(block_id, 0) > 0


# ------------------------------------------------------------------------------------------------ stores
class _MemEntry:
    __slots__ = ("value", "size", "deserialized", "level")

    def __init__(self, value, size, deserialized, level):
        self.value, self.size, self.deserialized, self.level = value, size, deserialized, level


class MemoryStore:
    def __init__(self, max_memory: int):
        self.max_memory = int(max_memory)
        self.used = 0
        self.entries: "OrderedDict[str, _MemEntry]" = OrderedDict()
        self.lock = threading.RLock()

    def contains(self, block_id: str) -> bool:
        with self.lock:
            return block_id in self.entries

    def get(self, block_id: str) -> Optional[_MemEntry]:
        with self.lock:
            e = self.entries.get(block_id)
            if e is not None:
                self.entries.move_to_end(block_id)          # LRU touch
            return e

    def put(self, block_id: str, entry: _MemEntry, evict: Callable[[str, _MemEntry], None],
            can_evict: Callable[[str], bool]) -> bool:
        """False when the block cannot fit even after evicting every evictable block."""
        with self.lock:
            old = self.entries.pop(block_id, None)
            if old is not None:
                self.used -= old.size
            if entry.size > self.max_memory:
                return False
            if self.used + entry.size > self.max_memory:
                need = self.used + entry.size - self.max_memory
                victims, freed = [], 0
                # like the reference, never evict blocks of the same RDD to make room for one of its partitions
                prefix = block_id.rsplit("_", 1)[0] if block_id.startswith("rdd_") else None
                for bid, e in self.entries.items():
                    if freed >= need:
                        break
                    if (prefix is not None and bid.startswith(prefix + "_")) or not can_evict(bid):
                        continue
                    victims.append((bid, e))
                    freed += e.size
                if freed < need:
                    return False
                for bid, e in victims:
                    del self.entries[bid]
                    self.used -= e.size
                    evict(bid, e)
            self.entries[block_id] = entry
            self.used += entry.size
            return True

    def remove(self, block_id: str) -> bool:
        with self.lock:
            e = self.entries.pop(block_id, None)
            if e is None:
                return False
            self.used -= e.size
            return True

    def clear(self) -> None:
        with self.lock:
            self.entries.clear()
            self.used = 0


class DiskStore:
    """Files under ``<root>/blockmgr-<uuid>/<2-hex-digit subdir>/<block id>`` (DiskBlockManager's hashed layout)."""

    def __init__(self, root: Optional[str] = None, subdirs: int = 64):
        base = root or tempfile.gettempdir()
        self.dir = os.path.join(base, f"blockmgr-{uuid.uuid4().hex[:12]}")      # created on the first spill, not before:
        self.subdirs = subdirs                                                   # most block managers never touch the disk
        self.sizes: Dict[str, int] = {}
        self.lock = threading.Lock()

    def _path(self, block_id: str) -> str:
        h = int(hashlib.md5(block_id.encode()).hexdigest()[:8], 16) % self.subdirs
        d = os.path.join(self.dir, f"{h:02x}")
        os.makedirs(d, exist_ok=True)
        return os.path.join(d, block_id)

    def put_bytes(self, block_id: str, data) -> None:
        path = self._path(block_id)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(data)
        os.replace(tmp, path)
        with self.lock:
            self.sizes[block_id] = len(data)

    def get_bytes(self, block_id: str) -> Optional[bytes]:
        if not self.contains(block_id):
            return None
        try:
            with open(self._path(block_id), "rb") as f:
                return f.read()
        except FileNotFoundError:
            return None

    def contains(self, block_id: str) -> bool:
        with self.lock:
            return block_id in self.sizes

    def size(self, block_id: str) -> int:
        with self.lock:
            return self.sizes.get(block_id, 0)

    def remove(self, block_id: str) -> bool:
        with self.lock:
            if self.sizes.pop(block_id, None) is None:
                return False
        try:
            os.remove(self._path(block_id))
        except FileNotFoundError:
            pass
        return True

    @property
    def used(self) -> int:
        with self.lock:
            return sum(self.sizes.values())

    def stop(self) -> None:
        shutil.rmtree(self.dir, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ master
class BlockManagerMasterEndpoint(RpcEndpoint):
    """Driver-side registry: block id -> holders, per-manager memory status; fans removal commands out to the
    managers' slave endpoints (BlockManagerMasterEndpoint.scala)."""
    NAME = "BlockManagerMaster"

    def __init__(self):
        self.managers: Dict[BlockManagerId, dict] = {}
        self.locations: Dict[str, Dict[BlockManagerId, Tuple[int, int]]] = {}
        self.slaves: Dict[BlockManagerId, RpcEndpointRef] = {}
        self._values_cached_everywhere: set = set()          # (reserved: broadcasts whose values may be cached without blocks)

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        op = msg[0]
        if op == "register":
            _, bm_id, max_mem, slave_ref = msg
            bm_id = BlockManagerId(*bm_id)
            self.managers[bm_id] = {"max_memory": max_mem, "mem_used": 0, "disk_used": 0, "registered": time.time()}
            self.slaves[bm_id] = slave_ref
            ctx.reply(True)
        elif op == "update":
            _, bm_id, block_id, mem_size, disk_size = msg
            bm_id = BlockManagerId(*bm_id)
            if bm_id not in self.managers:
                ctx.reply(False)               # unknown manager: it must re-register (BlockManager.reregister)
                return
            holders = self.locations.setdefault(block_id, {})
            old = holders.pop(bm_id, (0, 0))
            st = self.managers[bm_id]
            st["mem_used"] += mem_size - old[0]
            st["disk_used"] += disk_size - old[1]
            if mem_size or disk_size:
                holders[bm_id] = (mem_size, disk_size)
            elif not holders:
                self.locations.pop(block_id, None)
            ctx.reply(True)
        elif op == "locations":
            ctx.reply(list(self.locations.get(msg[1], {}).keys()))
        elif op == "locations_multi":
            ctx.reply([list(self.locations.get(b, {}).keys()) for b in msg[1]])
        elif op == "peers":
            me = BlockManagerId(*msg[1])
            ctx.reply([m for m in self.managers if m != me and m.executor_id != "driver"])
        elif op == "memory_status":
            ctx.reply({m: (s["max_memory"], s["max_memory"] - s["mem_used"]) for m, s in self.managers.items()})
        elif op == "storage_status":
            ctx.reply({m: {"max_memory": s["max_memory"], "mem_used": s["mem_used"], "disk_used": s["disk_used"],
                           "blocks": {b: h[m] for b, h in self.locations.items() if m in h}}
                       for m, s in self.managers.items()})
        elif op == "remove_executor":
            self._remove_executor(msg[1])
            ctx.reply(True)
        elif op in ("remove_block", "remove_rdd", "remove_broadcast"):
            ctx.reply(self._fan_out(op, msg[1], msg[2] if len(msg) > 2 else True))
        elif op == "contains":
            ctx.reply(bool(self.locations.get(msg[1])))
        else:
            ctx.send_failure(RpcError(f"BlockManagerMaster: unknown op {op!r}"))

    def _fan_out(self, op: str, key, include_driver: bool) -> int:
        # only the managers that actually hold matching blocks are asked (a per-round broadcast that one executor fetched costs
        # one or two round trips to clean up, not one per process of the cluster)
        if op == "remove_block":
            match = lambda b: b == key  # noqa: E731
        elif op == "remove_rdd":
            match = lambda b: b.startswith(f"rdd_{key}_")  # noqa: E731
        else:
            match = lambda b: b == f"broadcast_{key}" or b.startswith(f"broadcast_{key}_")  # noqa: E731
        holders = {m for b, h in self.locations.items() if match(b) for m in h}
        if op == "remove_broadcast":
            holders |= {m for m in self.slaves if m.executor_id != "driver"} if key in self._values_cached_everywhere else set()
        futs = []
        for bm_id, ref in list(self.slaves.items()):
            if bm_id not in holders or (not include_driver and bm_id.executor_id == "driver"):
                continue
            try:
                futs.append(ref.ask((op, key)))
            except RpcError:
                pass
        n = 0
        for f in futs:
            try:
                n += int(f.result(10) or 0)
            except Exception:  # noqa: BLE001  (a dead manager is cleaned up by remove_executor)
                pass
        # the slaves do not report these removals back one by one: drop the location entries here
        if op == "remove_block":
            gone = [key] if key in self.locations else []
        elif op == "remove_rdd":
            gone = [b for b in self.locations if b.startswith(f"rdd_{key}_")]
        else:
            gone = [b for b in self.locations if b == f"broadcast_{key}" or b.startswith(f"broadcast_{key}_")]
        for b in gone:
            for bm_id, (mem, disk) in list(self.locations[b].items()):
                if not include_driver and bm_id.executor_id == "driver":
                    continue
                st = self.managers.get(bm_id)
                if st is not None:
                    st["mem_used"] -= mem
                    st["disk_used"] -= disk
                del self.locations[b][bm_id]
            if not self.locations[b]:
                del self.locations[b]
        return n

    def _remove_executor(self, executor_id: str) -> None:
        for bm_id in [m for m in self.managers if m.executor_id == executor_id]:
            self.managers.pop(bm_id, None)
            self.slaves.pop(bm_id, None)
            for block_id in [b for b, h in self.locations.items() if bm_id in h]:
                self.locations[block_id].pop(bm_id, None)
                if not self.locations[block_id]:
                    del self.locations[block_id]

    def on_disconnected(self, address) -> None:
        for bm_id in [m for m in self.managers if (m.host, m.port) == (address.host, address.port)]:
            self._remove_executor(bm_id.executor_id)


class _SlaveEndpoint(RpcEndpoint):
    def __init__(self, bm: "BlockManager"):
        self.bm = bm

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        op, key = msg
        if op == "remove_block":
            ctx.reply(int(self.bm.remove_block(key, tell_master=False)))
        elif op == "remove_rdd":
            ctx.reply(self.bm.remove_rdd(key, tell_master=False))
        elif op == "remove_broadcast":
            ctx.reply(self.bm.remove_broadcast(key, tell_master=False))
        else:
            ctx.send_failure(RpcError(f"unknown slave op {op!r}"))


# ------------------------------------------------------------------------------------------------ manager
class BlockManager:
    def __init__(self, executor_id: str, rpc_env: Optional[RpcEnv] = None, master: Optional[RpcEndpointRef] = None,
                 max_memory: int = 1 << 30, disk_root: Optional[str] = None, fetch_timeout: float = 60.0):
        self.executor_id = str(executor_id)
        self.rpc_env = rpc_env
        self.master = master
        self.memory = MemoryStore(max_memory)
        self.disk = DiskStore(disk_root)
        self.info = BlockInfoManager()
        self.levels: Dict[str, StorageLevel] = {}
        self.fetch_timeout = fetch_timeout
        self.remote_bytes_read = 0
        self.remote_fetches = 0
        self.evictions = 0
        self.broadcast_removed_hooks: List[Callable[[int], None]] = []   # e.g. drop the cached deserialized value
        self._lock = threading.RLock()
        addr = rpc_env.address if rpc_env is not None and rpc_env.address is not None else ("localhost", 0)
        self.id = BlockManagerId(self.executor_id, addr[0], addr[1])
        if rpc_env is not None:
            if rpc_env.address is not None:
                rpc_env.set_stream_manager(self._serve_chunk)
            self._slave_name = f"BlockManagerEndpoint-{self.executor_id}-{uuid.uuid4().hex[:6]}"
            self._slave_ref = rpc_env.setup_endpoint(self._slave_name, _SlaveEndpoint(self))
            self._slave_ref = RpcEndpointRef(rpc_env, rpc_env.address, self._slave_name)
        if master is not None:
            self.register()

    # ---- master RPCs ------------------------------------------------------------------------------------
    def register(self) -> None:
        self.master.ask_sync(("register", tuple(self.id), self.memory.max_memory, self._slave_ref))

    def _tell_master(self, block_id: str) -> None:
        if self.master is None:
            return
        e = self.memory.get(block_id)
        ok = self.master.ask_sync(("update", tuple(self.id), block_id, e.size if e else 0, self.disk.size(block_id)))
        if ok is False:                                   # the master forgot us (restart / expiry): re-register + report all
            self.register()
            for b in self.block_ids():
                e = self.memory.get(b)
                self.master.ask_sync(("update", tuple(self.id), b, e.size if e else 0, self.disk.size(b)))

    def get_locations(self, block_id: str) -> List[BlockManagerId]:
        if self.master is None:
            return []
        return [BlockManagerId(*l) for l in self.master.ask_sync(("locations", block_id))]

    # ---- puts -----------------------------------------------------------------------------------------------
    def _evicted(self, block_id: str, e: _MemEntry) -> None:
        self.evictions += 1
        if e.level.use_disk and not self.disk.contains(block_id):
            self.disk.put_bytes(block_id, e.value if not e.deserialized else cloudpickle.dumps(e.value))
        elif not e.level.use_disk:
            self.levels.pop(block_id, None)
        try:
            self._tell_master(block_id)
        except RpcError:
            pass

    def _put(self, block_id: str, value, deserialized: bool, level: StorageLevel, tell_master: bool) -> bool:
        if not level.valid:
            raise ValueError("storage level must use memory or disk")
        self.info.lock_for_writing(block_id)
        stored = False
        try:
            size = estimate_size(value) if deserialized else len(value)
            if level.use_memory:
                entry = _MemEntry(value if (deserialized and level.deserialized) or not deserialized else None,
                                  size, deserialized and level.deserialized, level)
                if deserialized and not level.deserialized:        # *_SER levels keep bytes in memory
                    entry.value = cloudpickle.dumps(value)
                    entry.size = len(entry.value)
                stored = self.memory.put(block_id, entry, self._evicted, lambda b: not self.info.is_locked(b))
            if not stored and level.use_disk:
                self.disk.put_bytes(block_id, cloudpickle.dumps(value) if deserialized else value)
                stored = True
            if stored:
                self.levels[block_id] = level
        finally:
            self.info.unlock(block_id)
        if stored and tell_master:
            self._tell_master(block_id)
        if stored and level.replication > 1:
            self._replicate(block_id, value, deserialized, level)
        return stored

    def put_bytes(self, block_id: str, data, level: StorageLevel = StorageLevel.MEMORY_AND_DISK_SER,  # type: ignore[attr-defined]
                  tell_master: bool = True) -> bool:
        return self._put(block_id, bytes(data) if not isinstance(data, bytes) else data, False, level, tell_master)

    def put_single(self, block_id: str, value: Any, level: StorageLevel = StorageLevel.MEMORY_ONLY,  # type: ignore[attr-defined]
                   tell_master: bool = True) -> bool:
        return self._put(block_id, value, True, level, tell_master)

    def _replicate(self, block_id: str, value, deserialized: bool, level: StorageLevel) -> None:
        if self.master is None:
            return
        peers = [BlockManagerId(*p) for p in self.master.ask_sync(("peers", tuple(self.id)))]
        random.Random(hash(block_id) & 0xFFFF).shuffle(peers)
        data = cloudpickle.dumps(value) if deserialized else value
        single = StorageLevel(level.use_disk, level.use_memory, False, 1)
        done = 0
        for p in peers:
            if done >= level.replication - 1:
                break
            try:
                ref = self.rpc_env.endpoint_ref((p.host, p.port), "BlockUpload")
                if ref.ask_sync(("upload", block_id, data, single), self.fetch_timeout):
                    done += 1
            except RpcError:
                continue

    # ---- gets -----------------------------------------------------------------------------------------------
    def get_local_bytes(self, block_id: str) -> Optional[bytes]:
        if not self.info.lock_for_reading(block_id, timeout=self.fetch_timeout):
            return None
        try:
            e = self.memory.get(block_id)
            if e is not None:
                return cloudpickle.dumps(e.value) if e.deserialized else e.value
            return self.disk.get_bytes(block_id)
        finally:
            self.info.unlock(block_id)

    def get_local_values(self, block_id: str) -> Tuple[bool, Any]:
        """(found, value).  Blocks stored as bytes by ``put_single`` on a *_SER / disk level are unpickled."""
        if not self.info.lock_for_reading(block_id, timeout=self.fetch_timeout):
            return False, None
        try:
            e = self.memory.get(block_id)
            if e is not None:
                return True, (e.value if e.deserialized else pickle.loads(e.value))
            data = self.disk.get_bytes(block_id)
            if data is None:
                return False, None
            value = pickle.loads(data)
        finally:
            self.info.unlock(block_id)
        lvl = self.levels.get(block_id)
        if lvl is not None and lvl.use_memory and lvl.deserialized:      # promote back into memory (MEMORY_AND_DISK read path)
            entry = _MemEntry(value, estimate_size(value), True, lvl)
            self.memory.put(block_id, entry, self._evicted, lambda b: not self.info.is_locked(b))
        return True, value

    def get_remote_bytes(self, block_id: str, locations: Optional[Iterable[BlockManagerId]] = None) -> Optional[bytes]:
        """Fetch from a random holder; peers first, the driver last (TorrentBroadcast.readBlocks relies on this so the
        swarm, not the driver, serves most pieces)."""
        if self.rpc_env is None:
            return None
        locs = [BlockManagerId(*l) for l in (locations if locations is not None else self.get_locations(block_id))]
        locs = [l for l in locs if l != self.id]
        random.shuffle(locs)
        locs.sort(key=lambda l: l.executor_id == "driver")
        for loc in locs:
            try:
                data = self.rpc_env.fetch_chunk((loc.host, loc.port), block_id, self.fetch_timeout)
            except RpcError:
                continue
            self.remote_bytes_read += len(data)
            self.remote_fetches += 1
            return data
        return None

    def get_bytes(self, block_id: str) -> Optional[bytes]:
        data = self.get_local_bytes(block_id)
        return data if data is not None else self.get_remote_bytes(block_id)

    def contains(self, block_id: str) -> bool:
        return self.memory.contains(block_id) or self.disk.contains(block_id)

    def block_ids(self) -> List[str]:
        with self.memory.lock:
            ids = set(self.memory.entries)
        with self.disk.lock:
            ids |= set(self.disk.sizes)
        return sorted(ids)

    def _serve_chunk(self, chunk_id: str) -> Optional[bytes]:
        return self.get_local_bytes(chunk_id)

    # ---- removal ---------------------------------------------------------------------------------------------
    def remove_block(self, block_id: str, tell_master: bool = True) -> bool:
        if not self.contains(block_id):
            return False
        self.info.lock_for_writing(block_id)
        try:
            a = self.memory.remove(block_id)
            b = self.disk.remove(block_id)
            self.levels.pop(block_id, None)
        finally:
            self.info.unlock(block_id)
        if tell_master and (a or b):
            self._tell_master(block_id)
        return a or b

    def remove_rdd(self, dataset_id: int, tell_master: bool = True) -> int:
        ids = [b for b in self.block_ids() if b.startswith(f"rdd_{dataset_id}_")]
        return sum(1 for b in ids if self.remove_block(b, tell_master))

    def remove_broadcast(self, broadcast_id: int, tell_master: bool = True) -> int:
        pre = f"broadcast_{broadcast_id}"
        ids = [b for b in self.block_ids() if b == pre or b.startswith(pre + "_")]
        for hook in self.broadcast_removed_hooks:
            hook(int(broadcast_id))
        return sum(1 for b in ids if self.remove_block(b, tell_master))

    # ---- status ----------------------------------------------------------------------------------------------
    def memory_status(self) -> Tuple[int, int]:
        return self.memory.max_memory, self.memory.max_memory - self.memory.used

    def stop(self) -> None:
        self.memory.clear()
        self.disk.stop()
        if self.rpc_env is not None:
            self.rpc_env.stop_endpoint(self._slave_name)


class BlockUploadEndpoint(RpcEndpoint):
    """Receives replicated blocks (``NettyBlockRpcServer`` UploadBlock)."""
    NAME = "BlockUpload"

    def __init__(self, bm: BlockManager):
        self.bm = bm

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        _, block_id, data, level = msg
        ctx.reply(self.bm.put_bytes(block_id, data, level))


class BlockManagerMaster:
    """Driver-side convenience wrapper over the master endpoint ref (BlockManagerMaster.scala:31)."""

    def __init__(self, ref: RpcEndpointRef, timeout: float = 30.0):
        self.ref, self.timeout = ref, timeout

    def get_locations(self, block_id: str) -> List[BlockManagerId]:
        return [BlockManagerId(*l) for l in self.ref.ask_sync(("locations", block_id), self.timeout)]

    def contains(self, block_id: str) -> bool:
        return self.ref.ask_sync(("contains", block_id), self.timeout)

    def remove_block(self, block_id: str) -> int:
        return self.ref.ask_sync(("remove_block", block_id), self.timeout)

    def remove_rdd(self, dataset_id: int) -> int:
        return self.ref.ask_sync(("remove_rdd", dataset_id), self.timeout)

    def remove_broadcast(self, broadcast_id: int, remove_from_driver: bool = True) -> int:
        return self.ref.ask_sync(("remove_broadcast", broadcast_id, remove_from_driver), self.timeout)

    def remove_executor(self, executor_id: str) -> None:
        self.ref.ask_sync(("remove_executor", str(executor_id)), self.timeout)

    def memory_status(self) -> dict:
        return self.ref.ask_sync(("memory_status",), self.timeout)

    def storage_status(self) -> dict:
        return self.ref.ask_sync(("storage_status",), self.timeout)
