"""asyncframework_b200 — a Blackwell-native asynchronous-optimization engine.

Same capabilities and API surface as ASYNCframework/ASYNCframework (a Spark 2.3.2 fork adding
ASYNCcontext / ASYNCreduce / ASYNCaggregate / ASYNCbarrier / ASYNCbroadcast and the ASGD / ASAGA
drivers), rebuilt B200-first: PyTorch at the API layer, a C++ host runtime, and hand-written
sm_100a CUDA kernels whose gradient push / parameter pull run over NVLink from inside the kernel.
"""
from .config import AsyncConfig, parse_cli
from .core import (AsyncBroadcast, AsyncContext, AsyncSession, Broadcast, PartialResult, Scheduler, StatTable,
                   WorkerState, barriers)
from .data import (MLUtils, Shard, ShardedDataset, load_libsvm_file, save_as_libsvm_file,
                   synthetic_dataset)
from .linalg import DenseVector, LabeledPoint, SparseVector, Vectors, blas_util

__version__ = "0.1.0"

__all__ = ["AsyncConfig", "parse_cli", "AsyncBroadcast", "AsyncContext", "AsyncSession", "Broadcast", "PartialResult",
           "Scheduler", "StatTable", "WorkerState", "barriers", "MLUtils", "Shard", "ShardedDataset",
           "load_libsvm_file", "save_as_libsvm_file", "synthetic_dataset", "DenseVector", "LabeledPoint",
           "SparseVector", "Vectors", "blas_util", "__version__"]
