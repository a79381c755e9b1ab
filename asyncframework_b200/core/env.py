"""``AsyncEnv`` — the per-process runtime bundle (``SparkEnv`` analogue, core/SparkEnv.scala:217-400).

The reference's ``SparkEnv`` holds, per JVM: RpcEnv (:249), serializer (:283) / closure serializer (:289),
BroadcastManager (:302), BlockTransferService (:339), BlockManager (:349) (+ MapOutputTracker, ShuffleManager,
MemoryManager, MetricsSystem).  Here, per OS process of a *cluster* deployment (``local-cluster[n]`` or
``spark://host:port``): an :class:`RpcEnv`, a :class:`BlockManager` registered with the driver's
:class:`BlockManagerMasterEndpoint`, the serializer / compression codec pair and the typed runtime limits
(``max_direct_result_size`` …).  The in-process ``local[N]`` scheduler and the fused multi-GPU backend do not
need one: their tasks share the driver's address space / exchange data over NVLink.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from typing import Optional

from ..rpc import RpcEndpointRef, RpcEnv
from ..storage.block_manager import (BlockManager, BlockManagerMaster, BlockManagerMasterEndpoint,
                                     BlockUploadEndpoint)
from .serializer import CompressionCodec, Serializer


@dataclass
class RuntimeLimits:
    """Typed entries of core/internal/config/package.scala that govern the result / broadcast paths."""
    max_direct_result_size: int = 1 << 20          # spark.task.maxDirectResultSize (Executor.scala:140-141)
    max_result_size: int = 1 << 30                 # spark.driver.maxResultSize (TaskSetManager.scala:706-712); 0 = unlimited
    broadcast_block_size: int = 4 << 20            # spark.broadcast.blockSize (TorrentBroadcast.scala:81)
    broadcast_inline_threshold: int = 32 << 10     # payloads up to this size travel INSIDE the handle (no pieces, no block-manager RPCs)
    broadcast_compress: bool = True                # spark.broadcast.compress
    broadcast_checksum: bool = True                # spark.broadcast.checksum
    storage_memory: int = 1 << 30                  # executor storage memory (UnifiedMemoryManager storage region)
    compression_codec: str = "zlib"                # spark.io.compression.codec
    serializer: str = "pickle"                     # spark.serializer
    local_dir: Optional[str] = None                # spark.local.dir
    network_timeout: float = 120.0                 # spark.network.timeout (also: executor heartbeat expiry)
    heartbeat_interval: float = 10.0               # spark.executor.heartbeatInterval
    auth_secret: Optional[str] = None              # spark.authenticate.secret
    encrypt: bool = False                          # spark.network.crypto.enabled (AES-GCM on every RPC / block-transfer frame)
    shuffle_memory: int = 256 << 20                # estimated bytes a shuffle sorter / combine map may hold before it spills
    shuffle_spill_elements: int = 0                # spark.shuffle.spill.numElementsForceSpillThreshold (0 = off)
    shuffle_spill_batch: int = 10000               # spark.shuffle.spill.batchSize
    shuffle_spill_compress: bool = True            # spark.shuffle.spill.compress


class AsyncEnv:
    _instance: Optional["AsyncEnv"] = None
    _lock = threading.Lock()

    def __init__(self, executor_id: str, rpc_env: RpcEnv, block_manager: BlockManager, limits: RuntimeLimits,
                 is_driver: bool, bm_master: Optional[BlockManagerMaster] = None):
        self.executor_id = executor_id
        self.rpc_env = rpc_env
        self.block_manager = block_manager
        self.limits = limits
        self.is_driver = is_driver
        self.bm_master = bm_master
        self.serializer = Serializer.create(limits.serializer)
        self.codec = CompressionCodec.create(limits.compression_codec)
        self.broadcast_cache: dict = {}             # id -> value (BroadcastManager.cachedValues)
        self.broadcast_lock = threading.Lock()
        self.broadcast_fetch_locks: dict = {}       # id -> lock held while one thread assembles that broadcast
        self._stopped = False
        block_manager.broadcast_removed_hooks.append(self._drop_broadcast_value)

    def _drop_broadcast_value(self, bid: int) -> None:
        if not self.is_driver:                     # the driver's handle keeps its own reference until destroy()
            with self.broadcast_lock:
                self.broadcast_cache.pop(bid, None)

    # ---- process-global accessor (SparkEnv.get / SparkEnv.set) ---------------------------------------------------
    @classmethod
    def get(cls) -> Optional["AsyncEnv"]:
        return cls._instance

    @classmethod
    def set(cls, env: Optional["AsyncEnv"]) -> None:
        with cls._lock:
            cls._instance = env

    # ---- factories -------------------------------------------------------------------------------------------------
    @classmethod
    def create_driver(cls, limits: Optional[RuntimeLimits] = None, host: str = "127.0.0.1", port: int = 0,
                      name: str = "driver") -> "AsyncEnv":
        limits = limits or RuntimeLimits()
        rpc = RpcEnv(name, host, port, secret=limits.auth_secret, encrypt=limits.encrypt)
        master_ep = BlockManagerMasterEndpoint()
        master_ref = rpc.setup_endpoint(BlockManagerMasterEndpoint.NAME, master_ep)
        bm = BlockManager("driver", rpc, master_ref, limits.storage_memory, limits.local_dir, limits.network_timeout)
        rpc.setup_endpoint(BlockUploadEndpoint.NAME, BlockUploadEndpoint(bm))
        env = cls("driver", rpc, bm, limits, True, BlockManagerMaster(master_ref))
        env.master_endpoint = master_ep
        env._prev = cls._instance if (cls._instance is not None and not cls._instance._stopped) else None
        cls.set(env)
        return env

    @classmethod
    def create_executor(cls, executor_id: str, driver_address, limits: Optional[RuntimeLimits] = None,
                        host: str = "127.0.0.1", port: int = 0) -> "AsyncEnv":
        limits = limits or RuntimeLimits()
        rpc = RpcEnv(f"executor-{executor_id}", host, port, secret=limits.auth_secret, encrypt=limits.encrypt)
        master_ref: RpcEndpointRef = rpc.endpoint_ref(driver_address, BlockManagerMasterEndpoint.NAME)
        bm = BlockManager(str(executor_id), rpc, master_ref, limits.storage_memory, limits.local_dir, limits.network_timeout)
        rpc.setup_endpoint(BlockUploadEndpoint.NAME, BlockUploadEndpoint(bm))
        env = cls(str(executor_id), rpc, bm, limits, False, BlockManagerMaster(master_ref))
        cls.set(env)
        return env

    def stop(self) -> None:
        if self._stopped:
            return
        self._stopped = True
        try:
            self.block_manager.stop()
        finally:
            self.rpc_env.shutdown()
            if AsyncEnv._instance is self:       # several drivers in one process (tests): fall back to the outer one
                prev = getattr(self, "_prev", None)
                AsyncEnv.set(prev if prev is not None and not prev._stopped else None)
