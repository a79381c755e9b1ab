"""Host-side block storage (BlockManager / MemoryStore / DiskStore / BlockManagerMaster analogue)."""
from .block_manager import (BlockInfoManager, BlockManager, BlockManagerId, BlockManagerMaster,  # noqa: F401
                            BlockManagerMasterEndpoint, BlockUploadEndpoint, DiskStore, MemoryStore, StorageLevel,
                            broadcast_block_id, estimate_size, rdd_block_id, task_result_block_id)
