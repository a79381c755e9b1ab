"""Reference-spelling compatibility layer: the Scala names of the ASYNC API on top of this framework, so a
program written against ASYNCframework/ASYNCframework can be transcribed line by line.

    from asyncframework_b200.compat import *
    sc = SparkContext("local[8]")
    points = MLUtils.loadLibSVMFile(sc, path).repartition(numPart)
    AC = ASYNCcontext(numPart, sc)
    ...
    pointsIndexed.ASYNCbarrier(f, AC.STAT).sample(False, b, seed + k + 1).map(...).ASYNCreduce(comOp, AC)
    info = AC.ASYNCcollectAll(); info.getStaleness(); info.gettaskResult(); info.getWorkerID()

Reference symbols: core/rdd/ASYNCcontext.scala, workerState.scala, RDDPartialRes.scala, RDD.scala:1044-1345,
broadcast/ASYNCbroadcast.scala, mllib/BLASUtil.scala, mllib/util/MLUtils.scala.
"""
from __future__ import annotations

from typing import Any, Optional

from .conf import AsyncConf as _AsyncConf
from .core.broadcast import AsyncBroadcast as _AsyncBroadcast
from .core.context import AsyncContext as _AsyncContext
from .core.partial_result import PartialResult as RDDPartialRes
from .core.session import AsyncSession as _AsyncSession
from .core.worker_state import WorkerState as workerState
from .data.dataset import HashPartitioner, RangePartitioner  # noqa: F401
from .data.dataset import ShardedDataset as RDD
from .data.libsvm import MLUtils as _MLUtils
from .linalg import DenseVector, LabeledPoint, SparseVector, Vectors  # noqa: F401
from .linalg.blas_util import axpy_op as axpyOp, dot_op as dotOp, scal_op as scalOp  # noqa: F401
from .linalg import DenseMatrix, Matrices, SparseMatrix  # noqa: F401
from .models import (LassoWithSGD, LinearRegressionWithSGD, LogisticRegressionWithLBFGS, LogisticRegressionWithSGD,  # noqa: F401
                     RidgeRegressionWithSGD, SVMWithSGD)
from .optim import LBFGS  # noqa: F401
from .storage import StorageLevel  # noqa: F401

_checkpoint_root = {"dir": None}
from .optim import (GradientDescent, HingeGradient, L1Updater, LeastSquaresGradient, LogisticGradient,  # noqa: F401
                    SimpleUpdater, SquaredL2Updater)


class SparkConf(_AsyncConf):
    """``new SparkConf().setMaster(...).setAppName(...).set(k, v)`` (core/SparkConf.scala:82-119)."""

    def __init__(self, loadDefaults: bool = True):
        super().__init__(load_defaults=loadDefaults)

    def setMaster(self, m: str): return self.set_master(m)
    def setAppName(self, n: str): return self.set_app_name(n)
    def setAll(self, items): return self.set_all(items)
    def setIfMissing(self, k, v): return self.set_if_missing(k, v)
    def getAll(self): return self.get_all()
    def getInt(self, k, d=0): return self.get_int(k, d)
    def getDouble(self, k, d=0.0): return self.get_double(k, d)
    def getBoolean(self, k, d=False): return self.get_boolean(k, d)
    def getSizeAsBytes(self, k, d=0): return self.get_size_as_bytes(k, d)
    def getTimeAsSeconds(self, k, d=0): return self.get_time_as_seconds(k, d)
    def toDebugString(self): return self.to_debug_string()


class SparkContext(_AsyncSession):
    """``new SparkContext(conf)`` / ``new SparkContext(master, appName)`` — ``setLogLevel``, ``set_mode``, ``broadcast``,
    ``parallelize``, ``textFile``, ``setCheckpointDir``, ``getConf``, ``defaultParallelism``."""

    def __init__(self, master=None, appName: Optional[str] = None, conf=None, **kw):
        if isinstance(master, _AsyncConf):          # new SparkContext(conf)
            master, conf = None, master
        super().__init__(master, appName, conf=conf, **kw)
        self._checkpoint_dir: Optional[str] = None

    def getConf(self):
        return self.runtime_conf.clone()

    @property
    def defaultParallelism(self) -> int:
        return self.default_parallelism

    def setCheckpointDir(self, directory: str) -> None:
        """``sc.setCheckpointDir`` (SparkContext.scala:2337): ``rdd.checkpoint()`` without an argument writes below it."""
        import os
        os.makedirs(directory, exist_ok=True)
        self._checkpoint_dir = directory
        _checkpoint_root["dir"] = directory

    def setLogLevel(self, level: str) -> None:
        self.set_log_level(level)

    def textFile(self, path: str, minPartitions: Optional[int] = None):
        return self.text_file(path, minPartitions)


class ASYNCcontext(_AsyncContext):
    """``new ASYNCcontext[T]()``; pass the SparkContext (or a scheduler) so ``set_mode`` reaches it."""

    def __init__(self, num_workers: int, sc: Any = None):
        super().__init__(num_workers, getattr(sc, "scheduler", sc))

    def getSize(self) -> int: return self.size()
    def hasNext(self) -> bool: return self.has_next()
    def getCurrentTime(self) -> int: return self.current_time
    def setCurrentTime(self, t: int) -> None: self.current_time = t
    def add2currentTime(self, t: int) -> None: self.add_to_current_time(t)
    def setLastTime(self, t: int) -> None: self.last_time = t
    def isOld(self) -> bool: return self.is_old()
    def setRecordStat(self, b: bool) -> None: self.set_record_stat(b)
    def getRecordStat(self) -> bool: return self.record_stat


class ASYNCbroadcast(_AsyncBroadcast):
    """``new ASYNCbroadcast(sc, w)`` needs the context that owns the version ring: pass the ASYNCcontext."""

    def __init__(self, ctx, value, slots: int = 8):
        super().__init__(ctx, value, slots)

    def getIndex(self) -> int: return self.index
    def setID(self, i: int) -> None: self.set_id(i)


class MLUtils(_MLUtils):
    @staticmethod
    def loadLibSVMFile(sc, path: str, numFeatures: int = -1, minPartitions: Optional[int] = None, **kw):
        return sc.load_libsvm_file(path, numFeatures, minPartitions, **kw)

    @staticmethod
    def saveAsLibSVMFile(data, dir_path: str) -> None:
        from .data.libsvm import save_as_libsvm_file
        save_as_libsvm_file(data, dir_path)

    kFold = staticmethod(_MLUtils.k_fold)
    loadVectors = staticmethod(_MLUtils.load_vectors)
    loadLabeledPoints = staticmethod(_MLUtils.load_labeled_points)
    appendBias = staticmethod(_MLUtils.append_bias)


# --- PySpark keyword spellings on the train() entry points ---------------------------------------------------------
_KW = {"miniBatchFraction": "mini_batch_fraction", "initialWeights": "initial_weights", "regParam": "reg_param", "regType": "reg_type",
       "validateData": "validate_data", "convergenceTol": "convergence_tol", "numIterations": "iterations", "stepSize": "step"}


def _camel_train(cls):
    """``LinearRegressionWithSGD.train(data, iterations=100, step=1.0, miniBatchFraction=1.0, initialWeights=None, regParam=0.0,
    regType=None, intercept=False, validateData=True, convergenceTol=0.001)`` (python/pyspark/mllib/regression.py:229)."""
    orig = cls.__dict__["train"].__func__
    if getattr(orig, "_camel", False):
        return

    def train(klass, data, *args, **kw):
        return orig(klass, data, *args, **{_KW.get(k, k): v for k, v in kw.items()})
    train._camel = True          # type: ignore[attr-defined]
    train.__doc__ = orig.__doc__
    cls.train = classmethod(train)


for _c in (LinearRegressionWithSGD, LogisticRegressionWithSGD, SVMWithSGD, LogisticRegressionWithLBFGS):
    if "train" in _c.__dict__:
        _camel_train(_c)


# --- camelCase accessors on the shared classes (added once, idempotent) -----------------------------------
def _alias(cls, **names):
    for new, old in names.items():
        if not hasattr(cls, new):
            setattr(cls, new, getattr(cls, old) if isinstance(old, str) else old)


_alias(RDDPartialRes, gettaskResult="get_task_result", getStaleness="get_staleness", getbatchSize="get_batch_size",
       getWorkerID="get_worker_id", getTimestamp="get_timestamp")
_alias(workerState, getStaleness=lambda self: self.staleness, getAvailability=lambda self: self.available,
       getNumTasks=lambda self: self.num_tasks, getAverageTaskTime=lambda self: self.avg_task_time_ms,
       getAvailableWorkers=lambda self: self.available_workers(), getMaxStaleness=lambda self: self.max_staleness(),
       setAvailability=lambda self, v: setattr(self, "available", v))
_alias(RDD, zipWithIndex="zip_with_index", mapPartitionsWithIndex="map_partitions_with_index", getNumPartitions="get_num_partitions",
       treeAggregate="tree_aggregate", flatMap="flat_map", mapPartitions="map_partitions", foreachPartition="foreach_partition",
       reduceByKey="reduce_by_key", groupByKey="group_by_key", keyBy="key_by", mapValues="map_values", countByKey="count_by_key",
       countByValue="count_by_value", sortBy="sort_by", takeSample="take_sample", zipWithUniqueId="zip_with_unique_id",
       partitionBy="partition_by", combineByKey="combine_by_key", aggregateByKey="aggregate_by_key", foldByKey="fold_by_key",
       leftOuterJoin="left_outer_join", rightOuterJoin="right_outer_join", fullOuterJoin="full_outer_join",
       subtractByKey="subtract_by_key", sortByKey="sort_by_key", collectAsMap="collect_as_map", takeOrdered="take_ordered",
       isEmpty="is_empty", randomSplit="random_split", saveAsTextFile="save_as_text_file", toLocalIterator="to_local_iterator",
       setName="set_name", toDebugString="to_debug_string", zipPartitions="zip_partitions", getStorageLevel="get_storage_level",
       flatMapValues="flat_map_values", groupBy="group_by", sampleByKey="sample_by_key", sampleVariance="sample_variance",
       sampleStdev="sample_stdev", countApprox="count_approx", localCheckpoint="local_checkpoint", isCheckpointed="is_checkpointed",
       getCheckpointFile="get_checkpoint_file", repartitionAndSortWithinPartitions="repartition_and_sort_within_partitions",
       treeReduce="tree_reduce", countApproxDistinct="count_approx_distinct", saveAsPickleFile="save_as_pickle_file", mapShards="map_shards", ASYNCgrad="async_grad")
_alias(SparkContext, setJobGroup="set_job_group", clearJobGroup="clear_job_group", cancelJobGroup="cancel_job_group",
       getExecutorMemoryStatus="get_executor_memory_status", statusTracker=property(lambda self: self.status_tracker),
       addFile="add_file", wholeTextFiles="whole_text_files", pickleFile="pickle_file", emptyRDD="empty_dataset",
       applicationId=property(lambda self: self.application_id), uiWebUrl=property(lambda self: self.ui_web_url))
_alias(SparkContext, longAccumulator="long_accumulator", doubleAccumulator="double_accumulator", collectionAccumulator="collection_accumulator",
       cancelJob=lambda self, job_id: self.scheduler.cancel_job(job_id), cancelAllJobs=lambda self: self.scheduler.cancel_all_jobs())

def _checkpoint(self, directory: Optional[str] = None):
    import os
    d = directory or (_checkpoint_root["dir"] and os.path.join(_checkpoint_root["dir"], f"rdd-{self.ds_id & 0xFFFFFFFF}"))
    if not d:
        raise RuntimeError("Checkpoint directory has not been set in the SparkContext")
    return type(self).checkpoint_to(self, d)


if not hasattr(RDD, "checkpoint_to"):
    RDD.checkpoint_to = RDD.checkpoint                 # the explicit-directory form stays available
    RDD.checkpoint = _checkpoint

__all__ = ["HashPartitioner", "RangePartitioner", "SparkConf", "StorageLevel", "Matrices", "DenseMatrix", "SparseMatrix", "LBFGS", "LogisticRegressionWithLBFGS",
           "RidgeRegressionWithSGD", "LassoWithSGD", "SparkContext", "ASYNCcontext", "ASYNCbroadcast", "RDDPartialRes", "workerState", "RDD", "MLUtils", "Vectors",
           "DenseVector", "SparseVector", "LabeledPoint", "axpyOp", "dotOp", "scalOp", "GradientDescent", "LeastSquaresGradient",
           "LogisticGradient", "HingeGradient", "SimpleUpdater", "L1Updater", "SquaredL2Updater", "LinearRegressionWithSGD",
           "LogisticRegressionWithSGD", "SVMWithSGD"]
