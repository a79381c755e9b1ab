from .context import AsyncContext, QueueClosed
from .partial_result import PartialResult
from .worker_state import StatTable, WorkerState
from .broadcast import AsyncBroadcast, Broadcast
from .accumulator import Accumulator
from .scheduler import JobCancelled, JobFailed, JobWaiter, Scheduler, TaskContext
from .cluster import ClusterScheduler
from .session import AsyncSession
from . import barriers, sampler

__all__ = ["AsyncContext", "QueueClosed", "PartialResult", "StatTable", "WorkerState", "AsyncBroadcast",
           "Accumulator", "Broadcast", "JobCancelled", "JobFailed", "JobWaiter", "Scheduler", "TaskContext", "AsyncSession", "ClusterScheduler", "barriers", "sampler"]
