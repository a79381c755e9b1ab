"""``TorrentBroadcast`` for the cluster deployments — piece-wise, checksummed, peer-to-peer.

Reference (core/broadcast/TorrentBroadcast.scala): the driver serialises + compresses the value into
``spark.broadcast.blockSize`` = 4 MiB pieces stored in its BlockManager with Adler32 checksums (``writeBlocks``
:133-157); an executor's first ``.value`` fetches the pieces **in random order from the driver or any peer that
already has them**, stores each fetched piece in its own BlockManager and reports it to the master — so it becomes a
seeder (:160-201, :190) — reassembles (``unBlockifyObject`` :314-329) and caches the object by broadcast id
(``BroadcastManager.cachedValues``, :224-270).  The ASYNC fork makes the id mutable (``Broadcast.getBroadcastID /
changeBroadcastID``, Broadcast.scala:73-80,137-149; TorrentBroadcast.scala:85-112) so that
``ASYNCbroadcast.value(index)`` can re-read an *older* version out of the block store.

This file is that mechanism over :mod:`asyncframework_b200.storage` + :mod:`asyncframework_b200.rpc`.  The handle
that travels inside a task binary is tiny (id, piece count, checksums); the bytes move over the bulk chunk-fetch
path, peers first.  A per-broadcast ``_meta`` block lets a handle whose id was *changed* on the executor
(``change_broadcast_id``) discover the piece count / checksums of the other version.

On the fused multi-GPU backend none of this is used: ``w`` is pulled over NVLink inside the round kernel and
published through the NVLS multicast ring (``parallel/fused_backend.py::FusedEngine.publish_version``).
"""
from __future__ import annotations

import pickle
import random
import threading
import zlib
from typing import Any, List, Optional, Tuple

from ..storage.block_manager import StorageLevel, broadcast_block_id
from .broadcast import Broadcast, _bcast_ids, _snapshot
from .env import AsyncEnv


class BroadcastCorrupt(RuntimeError):
    pass


def blockify(data: bytes, block_size: int) -> List[bytes]:
    return [data[i:i + block_size] for i in range(0, len(data), block_size)] or [b""]


def _meta_id(bid: int) -> str:
    return f"broadcast_{bid}_meta"


class TorrentBroadcast(Broadcast):
    """Driver: ``TorrentBroadcast(value)`` writes the pieces.  Executors receive the handle by unpickling."""

    def __init__(self, value: Any, env: Optional[AsyncEnv] = None, snapshot: bool = True, allow_inline: bool = True):
        env = env or AsyncEnv.get()
        if env is None:
            raise RuntimeError("TorrentBroadcast needs an AsyncEnv (cluster deployment); use Broadcast for local[N]")
        self.id = next(_bcast_ids)
        self._changed_id: Optional[int] = None
        self._value = _snapshot(value) if snapshot else value
        self._valid = True
        self._lock = threading.Lock()
        self._inline: Optional[bytes] = None
        if allow_inline and env.limits.broadcast_inline_threshold > 0:
            # parameter-sized values (w of mnist8m = 6 KB, epsilon = 16 KB) ride inside the handle, i.e. inside the task binary:
            # zero block-manager round trips per round and nothing to clean up.  Versioned broadcasts (AsyncBroadcast) never
            # inline — their older versions must stay addressable by id in the block store (A9).
            data = env.serializer.dumps(self._value)
            if len(data) <= env.limits.broadcast_inline_threshold:
                self._inline = data
                self.num_blocks, self.checksums, self.compressed = 0, [], False
                return
        self.num_blocks, self.checksums, self.compressed = self._write_blocks(env, self._value)
        with env.broadcast_lock:
            env.broadcast_cache[self.id] = self._value
        # ContextCleaner: once this handle is unreachable, drop the pieces and cached values on every block manager
        from .cleaner import ContextCleaner
        ContextCleaner.get().register(self, "broadcast", _make_remover(env, self.id))

    # ------------------------------------------------------------------ driver side
    def _write_blocks(self, env: AsyncEnv, value: Any) -> Tuple[int, List[int], bool]:
        lim = env.limits
        data = env.serializer.dumps(value)
        if lim.broadcast_compress:
            data = env.codec.compress(data)
        pieces = blockify(data, lim.broadcast_block_size)
        sums = [zlib.adler32(p) & 0xFFFFFFFF for p in pieces] if lim.broadcast_checksum else []
        bm = env.block_manager
        for i, p in enumerate(pieces):
            if not bm.put_bytes(broadcast_block_id(self.id, i), p, StorageLevel.MEMORY_AND_DISK_SER):  # type: ignore[attr-defined]
                raise RuntimeError(f"failed to store piece {i} of broadcast {self.id} in the driver's block manager")
        meta = pickle.dumps((len(pieces), sums, lim.broadcast_compress))
        bm.put_bytes(_meta_id(self.id), meta, StorageLevel.MEMORY_AND_DISK_SER)  # type: ignore[attr-defined]
        return len(pieces), sums, lim.broadcast_compress

    # ------------------------------------------------------------------ pickling: only the handle travels
    def __reduce__(self):
        if self._inline is not None:
            return (_rebuild_inline, (self.id, self._inline))
        return (_rebuild, (self.id, self.num_blocks, self.checksums, self.compressed))

    # ------------------------------------------------------------------ mutable id (A9)
    def get_broadcast_id(self) -> int:
        return self.id if self._changed_id is None else self._changed_id

    def change_broadcast_id(self, new_id: int) -> None:
        """Point this handle at another broadcast's blocks (``changeBroadcastID``); ``value`` then reads THAT version."""
        self._changed_id = None if int(new_id) == self.id else int(new_id)

    # ------------------------------------------------------------------ value
    @property
    def value(self) -> Any:
        if not self._valid:
            raise RuntimeError(f"Attempted to use Broadcast({self.id}) after it was destroyed")
        bid = self.get_broadcast_id()
        if bid == self.id and self._value is not None:
            return self._value
        v = read_broadcast(bid, (self.num_blocks, self.checksums, self.compressed) if bid == self.id else None)
        if bid == self.id:
            self._value = v
        return v

    def unpersist(self, blocking: bool = False) -> None:
        """Remove the executors' copies (pieces + cached values); the driver keeps its pieces so the next ``.value``
        on an executor re-fetches (Broadcast.unpersist)."""
        env = AsyncEnv.get()
        if self._inline is None and env is not None and env.is_driver and env.bm_master is not None:
            env.bm_master.remove_broadcast(self.id, remove_from_driver=False)

    def destroy(self) -> None:
        env = AsyncEnv.get()
        if self._inline is None and env is not None and env.is_driver and env.bm_master is not None:
            env.bm_master.remove_broadcast(self.id, remove_from_driver=True)
            with env.broadcast_lock:
                env.broadcast_cache.pop(self.id, None)
        self._valid = False
        self._value = None


def _make_remover(env: AsyncEnv, bid: int):
    def remove() -> None:
        if env._stopped:
            return
        env.bm_master.remove_broadcast(bid, remove_from_driver=True)
        with env.broadcast_lock:
            env.broadcast_cache.pop(bid, None)
    return remove


def _rebuild_inline(bid: int, data: bytes) -> TorrentBroadcast:
    tb = TorrentBroadcast.__new__(TorrentBroadcast)
    tb.id, tb._changed_id, tb._valid, tb._inline = bid, None, True, data
    tb.num_blocks, tb.checksums, tb.compressed = 0, [], False
    tb._lock = threading.Lock()
    import pickle as _p
    tb._value = _p.loads(data)
    return tb


def _rebuild(bid: int, num_blocks: int, checksums: List[int], compressed: bool) -> TorrentBroadcast:
    tb = TorrentBroadcast.__new__(TorrentBroadcast)
    tb.id, tb._changed_id, tb._valid, tb._inline = bid, None, True, None
    tb.num_blocks, tb.checksums, tb.compressed = num_blocks, checksums, compressed
    tb._lock = threading.Lock()
    env = AsyncEnv.get()
    tb._value = None
    if env is not None:
        with env.broadcast_lock:
            tb._value = env.broadcast_cache.get(bid)
    return tb


def read_broadcast(bid: int, meta: Optional[Tuple[int, List[int], bool]] = None) -> Any:
    """Executor-side ``readBroadcastBlock``: cached value, else local pieces, else the swarm."""
    env = AsyncEnv.get()
    if env is None:
        raise RuntimeError("no AsyncEnv in this process: a TorrentBroadcast handle can only be read inside a cluster deployment")
    with env.broadcast_lock:
        if bid in env.broadcast_cache:
            return env.broadcast_cache[bid]
        fetch_lock = env.broadcast_fetch_locks.setdefault(bid, threading.Lock())
    with fetch_lock:                                   # concurrent tasks of one executor fetch a broadcast ONCE (TorrentBroadcast.synchronized)
        with env.broadcast_lock:
            if bid in env.broadcast_cache:
                return env.broadcast_cache[bid]
        try:
            return _fetch_broadcast(env, bid, meta)
        finally:
            with env.broadcast_lock:
                env.broadcast_fetch_locks.pop(bid, None)


def _fetch_broadcast(env: AsyncEnv, bid: int, meta: Optional[Tuple[int, List[int], bool]]) -> Any:
    bm = env.block_manager
    if meta is None:
        raw = bm.get_bytes(_meta_id(bid))
        if raw is None:
            raise KeyError(f"broadcast {bid}: no metadata block anywhere (destroyed, evicted or never existed)")
        meta = pickle.loads(raw)
    num_blocks, checksums, compressed = meta
    pieces: List[Optional[bytes]] = [None] * num_blocks
    order = list(range(num_blocks))
    random.shuffle(order)                                  # random piece order spreads the load over the seeders
    for i in order:
        pid = broadcast_block_id(bid, i)
        data = bm.get_local_bytes(pid)
        if data is None:
            data = bm.get_remote_bytes(pid)
            if data is None:
                raise KeyError(f"Failed to get {pid} of broadcast {bid}")
            if checksums and (zlib.adler32(data) & 0xFFFFFFFF) != checksums[i]:
                raise BroadcastCorrupt(f"corrupt remote block {pid}: checksum {zlib.adler32(data) & 0xFFFFFFFF:#x} "
                                       f"!= {checksums[i]:#x}")
            # store + report: this process now seeds the piece
            bm.put_bytes(pid, data, StorageLevel.MEMORY_AND_DISK_SER)  # type: ignore[attr-defined]
        pieces[i] = data
    blob = b"".join(pieces)  # type: ignore[arg-type]
    if compressed:
        blob = env.codec.decompress(blob)
    value = env.serializer.loads(blob)
    with env.broadcast_lock:
        env.broadcast_cache[bid] = value
    return value


def drop_cached_value(bid: int) -> None:
    env = AsyncEnv.get()
    if env is not None:
        with env.broadcast_lock:
            env.broadcast_cache.pop(bid, None)
