"""BlockManager stack, codecs and TorrentBroadcast (BlockManagerSuite / MemoryStoreSuite / BroadcastSuite /
CompressionCodecSuite scenarios of the reference, core/src/test/scala/org/apache/spark/{storage,broadcast,io})."""
import os
import threading
import time

import numpy as np
import pytest
import torch

from asyncframework_b200.core.env import AsyncEnv, RuntimeLimits
from asyncframework_b200.core.serializer import CompressionCodec, Serializer
from asyncframework_b200.core.torrent_broadcast import BroadcastCorrupt, TorrentBroadcast, read_broadcast
from asyncframework_b200.storage import (BlockInfoManager, BlockManager, DiskStore, StorageLevel, estimate_size)


# ------------------------------------------------------------------------------------------------ codecs
@pytest.mark.parametrize("name", ["none", "zlib", "lz4", "zstd", "lzma", "bz2"])
def test_codec_roundtrip(name):
    codec = CompressionCodec.create(name)
    rng = np.random.default_rng(0)
    for payload in (b"", b"a", b"abc" * 10000, rng.integers(0, 255, 100_000, dtype=np.uint8).tobytes(),
                    np.zeros(1 << 20, dtype=np.float32).tobytes()):
        blob = codec.compress(payload)
        assert CompressionCodec.decompress(blob) == payload          # self-describing: no codec needed to read
    if name != "none":
        assert len(codec.compress(b"\0" * 100000)) < 5000


def test_native_codecs_are_bound_and_detect_corruption():
    assert {"lz4", "zstd"} <= set(CompressionCodec.available()), "liblz4 / libzstd should be present in this image"
    for name in ("lz4", "zstd", "zlib"):
        blob = bytearray(CompressionCodec.create(name).compress(bytes(range(256)) * 64))
        blob[-5] ^= 0xFF
        blob[len(blob) // 2] ^= 0x55
        with pytest.raises(Exception):
            CompressionCodec.decompress(bytes(blob))


def test_serializers():
    pk, cp = Serializer.create("pickle"), Serializer.create("cloudpickle")
    assert pk.loads(pk.dumps({"a": torch.arange(3)}))["a"].tolist() == [0, 1, 2]
    k = 5
    f = cp.loads(cp.dumps(lambda x: x + k))
    assert f(1) == 6
    assert pk.loads(pk.dumps(lambda x: x * 2))(4) == 8                # falls back to by-value pickling
    with pytest.raises(ValueError):
        Serializer.create("nope")


# ------------------------------------------------------------------------------------------------ stores
def test_block_info_manager_read_write_locks():
    m = BlockInfoManager()
    assert m.lock_for_reading("b") and m.lock_for_reading("b")
    assert not m.lock_for_writing("b", blocking=False)
    m.unlock("b")
    m.unlock("b")
    assert m.lock_for_writing("b")
    assert not m.lock_for_reading("b", blocking=False)
    got = []
    t = threading.Thread(target=lambda: (m.lock_for_reading("b"), got.append(1), m.unlock("b")))
    t.start()
    time.sleep(0.05)
    assert not got
    m.downgrade("b")                      # writer -> reader lets the waiting reader in
    t.join(2)
    assert got == [1]
    m.unlock("b")
    with pytest.raises(RuntimeError):
        m.unlock("b")


def test_memory_store_lru_eviction_and_disk_fallback(tmp_path):
    bm = BlockManager("e", max_memory=3000, disk_root=str(tmp_path))
    assert bm.put_bytes("a", b"x" * 1000, StorageLevel.MEMORY_ONLY_SER)
    assert bm.put_bytes("b", b"y" * 1000, StorageLevel.MEMORY_AND_DISK_SER)
    assert bm.put_bytes("c", b"z" * 1000, StorageLevel.MEMORY_ONLY_SER)
    assert bm.get_local_bytes("a") == b"x" * 1000             # touch a -> b is now least recently used
    assert bm.put_bytes("d", b"w" * 1000, StorageLevel.MEMORY_ONLY_SER)
    assert not bm.memory.contains("b") and bm.disk.contains("b")          # dropped to disk, not lost
    assert bm.get_local_bytes("b") == b"y" * 1000
    assert bm.put_bytes("e", b"v" * 1000, StorageLevel.MEMORY_ONLY_SER)   # evicts c (memory only -> gone)
    assert not bm.contains("c") and bm.get_local_bytes("c") is None
    assert bm.evictions == 2
    assert not bm.put_bytes("huge", b"h" * 5000, StorageLevel.MEMORY_ONLY_SER)      # can never fit
    assert bm.put_bytes("huge", b"h" * 5000, StorageLevel.MEMORY_AND_DISK_SER)      # ... unless disk is allowed
    assert bm.disk.contains("huge") and bm.get_local_bytes("huge") == b"h" * 5000
    assert bm.remove_block("huge") and not bm.contains("huge")
    d = bm.disk.dir
    bm.stop()
    assert not os.path.exists(d)


def test_put_single_levels_and_rdd_blocks_do_not_evict_siblings(tmp_path):
    bm = BlockManager("e", max_memory=10_000, disk_root=str(tmp_path))
    t = torch.arange(500, dtype=torch.float64)                           # 4000 bytes
    assert estimate_size(t) == 4000
    assert bm.put_single("rdd_7_0", t, StorageLevel.MEMORY_ONLY)
    assert bm.put_single("rdd_7_1", t + 1, StorageLevel.MEMORY_ONLY)
    assert not bm.put_single("rdd_7_2", t + 2, StorageLevel.MEMORY_ONLY)   # would have to evict its own RDD's blocks
    assert bm.contains("rdd_7_0") and bm.contains("rdd_7_1")
    assert bm.put_single("rdd_8_0", t + 3, StorageLevel.MEMORY_ONLY)       # another RDD may evict them
    found, v = bm.get_local_values("rdd_8_0")
    assert found and v is not None and torch.equal(v, t + 3)
    assert bm.put_single("s", {"k": [1, 2, 3]}, StorageLevel.DISK_ONLY)
    assert bm.get_local_values("s") == (True, {"k": [1, 2, 3]})
    assert bm.put_single("m", [1, 2], StorageLevel.MEMORY_ONLY_SER)
    assert isinstance(bm.memory.get("m").value, bytes) and bm.get_local_values("m") == (True, [1, 2])
    assert bm.remove_rdd(7) + bm.remove_rdd(8) >= 1
    bm.stop()


# ------------------------------------------------------------------------------------------------ master + transfer
def _cluster(n=2, **lim):
    limits = RuntimeLimits(**lim)
    driver = AsyncEnv.create_driver(limits)
    execs = [AsyncEnv.create_executor(str(i), driver.rpc_env.address, limits) for i in range(n)]
    AsyncEnv.set(driver)
    return driver, execs


def _stop(driver, execs):
    for e in execs:
        e.stop()
    driver.stop()
    AsyncEnv.set(None)


def test_master_tracks_locations_and_remote_fetch_and_removal():
    driver, (e0, e1) = _cluster(2)
    try:
        assert e0.block_manager.put_bytes("blk", b"payload" * 100)
        locs = driver.bm_master.get_locations("blk")
        assert [l.executor_id for l in locs] == ["0"]
        assert e1.block_manager.get_local_bytes("blk") is None
        assert e1.block_manager.get_bytes("blk") == b"payload" * 100          # chunk fetch from executor 0
        assert e1.block_manager.remote_fetches == 1
        st = driver.bm_master.storage_status()
        assert st[e0.block_manager.id]["blocks"]["blk"][0] == 700 and st[e0.block_manager.id]["mem_used"] == 700
        assert driver.bm_master.remove_block("blk") == 1
        assert not e0.block_manager.contains("blk") and not driver.bm_master.contains("blk")
        assert driver.bm_master.storage_status()[e0.block_manager.id]["mem_used"] == 0
        # replication: MEMORY_ONLY_2 uploads a copy to a peer
        assert e0.block_manager.put_single("rdd_1_0", [1, 2, 3], StorageLevel.MEMORY_ONLY_2)
        assert {l.executor_id for l in driver.bm_master.get_locations("rdd_1_0")} == {"0", "1"}
        assert driver.bm_master.remove_rdd(1) == 2
        # executor loss drops its locations
        e1.block_manager.put_bytes("only-on-1", b"x")
        driver.bm_master.remove_executor("1")
        assert driver.bm_master.get_locations("only-on-1") == []
        # ... and an update from a forgotten manager makes it re-register and re-report everything
        e1.block_manager.put_bytes("again", b"y")
        assert {l.executor_id for l in driver.bm_master.get_locations("only-on-1")} == {"1"}
    finally:
        _stop(driver, [e0, e1])


def test_torrent_broadcast_pieces_checksums_and_swarm():
    driver, execs = _cluster(3, broadcast_block_size=64 * 1024, compression_codec="lz4")
    try:
        rng = np.random.default_rng(1)
        w = torch.from_numpy(rng.standard_normal(100_000))                  # 800 KB, incompressible -> 13 pieces
        bc = TorrentBroadcast(w)
        assert bc.num_blocks >= 12 and len(bc.checksums) == bc.num_blocks
        import cloudpickle
        handle_bytes = cloudpickle.dumps(bc)
        assert len(handle_bytes) < 2000                                     # only the handle travels in a task binary
        e0, e1, e2 = execs
        AsyncEnv.set(e0)
        h0 = cloudpickle.loads(handle_bytes)
        assert torch.equal(h0.value, w)
        assert e0.block_manager.remote_fetches == bc.num_blocks             # all from the driver (only seeder)
        # executor 0 now seeds every piece: the next reader prefers peers over the driver
        assert all({l.executor_id for l in driver.bm_master.get_locations(f"broadcast_{bc.id}_piece{i}")} == {"driver", "0"}
                   for i in range(bc.num_blocks))
        sent_by_driver_before = driver.rpc_env.bytes_sent
        AsyncEnv.set(e1)
        h1 = cloudpickle.loads(handle_bytes)
        assert torch.equal(h1.value, w)
        assert driver.rpc_env.bytes_sent - sent_by_driver_before < 100_000    # control traffic only: pieces came from executor 0
        assert h1.value is cloudpickle.loads(handle_bytes).value             # cached by id in the executor's env
        # corruption is detected by the Adler32 of the piece
        AsyncEnv.set(e2)
        bad = cloudpickle.loads(handle_bytes)
        bad.checksums = list(bad.checksums)
        bad.checksums[3] ^= 1
        with pytest.raises(BroadcastCorrupt):
            _ = bad.value
        # unpersist removes executor copies, the driver keeps seeding; destroy removes everything
        AsyncEnv.set(driver)
        bc.unpersist()
        assert not e0.block_manager.contains(f"broadcast_{bc.id}_piece0") and bc.id not in e0.broadcast_cache
        AsyncEnv.set(e0)
        assert torch.equal(cloudpickle.loads(handle_bytes).value, w)
        AsyncEnv.set(driver)
        bc.destroy()
        assert driver.bm_master.get_locations(f"broadcast_{bc.id}_piece0") == []
        with pytest.raises(RuntimeError, match="destroyed"):
            _ = bc.value
        AsyncEnv.set(e2)
        e2.broadcast_cache.clear()
        with pytest.raises(KeyError):
            read_broadcast(bc.id)
    finally:
        _stop(driver, execs)


def test_mutable_broadcast_id_reads_older_versions_from_the_block_store():
    """A9: ``changeBroadcastID`` on the executor re-reads another version (what ASYNCbroadcast.value(index) does)."""
    driver, (e0,) = _cluster(1)
    try:
        import cloudpickle
        versions = [TorrentBroadcast(torch.full((10,), float(v)), allow_inline=False) for v in range(3)]   # what AsyncBroadcast does
        handle = cloudpickle.dumps(versions[2])
        AsyncEnv.set(e0)
        h = cloudpickle.loads(handle)
        assert float(h.value[0]) == 2.0 and h.get_broadcast_id() == versions[2].id
        h.change_broadcast_id(versions[0].id)
        assert h.get_broadcast_id() == versions[0].id and float(h.value[0]) == 0.0
        h.change_broadcast_id(versions[2].id)                                 # restore, like the reference does
        assert float(h.value[0]) == 2.0
    finally:
        _stop(driver, [e0])


def test_concurrent_readers_in_one_executor_fetch_a_broadcast_once():
    import cloudpickle
    driver, (e0,) = _cluster(1, broadcast_block_size=32 * 1024)
    try:
        bc = TorrentBroadcast(torch.arange(100_000, dtype=torch.float64))        # ~25 pieces
        handle = cloudpickle.dumps(bc)
        AsyncEnv.set(e0)
        out, errs = [], []

        def reader():
            try:
                out.append(cloudpickle.loads(handle).value)
            except BaseException as exc:  # noqa: BLE001
                errs.append(exc)
        ts = [threading.Thread(target=reader) for _ in range(6)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs and len(out) == 6 and all(o is out[0] for o in out)       # one assembled object shared by all readers
        assert e0.block_manager.remote_fetches == bc.num_blocks                   # every piece crossed the wire exactly once
    finally:
        _stop(driver, [e0])


def test_parameter_sized_broadcasts_travel_inside_the_handle():
    """w of mnist8m is 6 KB: no pieces, no block-manager round trips, nothing to clean up; big values still go through the swarm."""
    import cloudpickle
    driver, (e0,) = _cluster(1)
    try:
        small = TorrentBroadcast(torch.arange(784, dtype=torch.float64))
        big = TorrentBroadcast(torch.arange(100_000, dtype=torch.float64))
        assert small.num_blocks == 0 and big.num_blocks >= 1
        assert not driver.block_manager.contains(f"broadcast_{small.id}_meta") and driver.block_manager.contains(f"broadcast_{big.id}_meta")
        hs, hb = cloudpickle.dumps(small), cloudpickle.dumps(big)
        assert 6000 < len(hs) < 8000 and len(hb) < 2000
        AsyncEnv.set(e0)
        before = e0.block_manager.remote_fetches
        assert torch.equal(cloudpickle.loads(hs).value, torch.arange(784, dtype=torch.float64))
        assert e0.block_manager.remote_fetches == before                         # nothing fetched for the inlined one
        assert float(cloudpickle.loads(hb).value[-1]) == 99_999.0 and e0.block_manager.remote_fetches > before
        AsyncEnv.set(driver)
        small.unpersist()
        small.destroy()
        with pytest.raises(RuntimeError, match="destroyed"):
            _ = small.value
        never = TorrentBroadcast(torch.zeros(4), AsyncEnv.create_driver(RuntimeLimits(broadcast_inline_threshold=0)))
        assert never.num_blocks == 1
        AsyncEnv.get().stop()
    finally:
        _stop(driver, [e0])
