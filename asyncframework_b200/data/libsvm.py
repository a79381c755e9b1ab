"""MLlib-compatible LIBSVM ingest / export.

Parity: ``MLUtils.loadLibSVMFile(sc, path[, numFeatures[, minPartitions]])``
(mllib/util/MLUtils.scala:71-89,155-167; record grammar :97-148) and ``saveAsLibSVMFile``
(:176-190): file **or directory of part files**, blank / ``#`` lines skipped, 1-based strictly
ascending indices, label-only rows allowed, auto feature count = max index + 1.

The parser is the multithreaded C++ mmap parser in csrc/runtime/libsvm_parser.cpp; the pure
Python ``parse_libsvm_record`` exists for single-record use and as the test oracle.
The reference always produces sparse vectors; here the dense-vs-CSR storage decision is made at
load time from the measured density (``layout="auto"``).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import _native
from ..linalg.vectors import LabeledPoint, SparseVector, Vectors
from .batch import Shard, padded_ld
from .dataset import ShardedDataset


def parse_libsvm_record(line: str) -> Tuple[float, np.ndarray, np.ndarray]:
    """(label, 0-based indices, values) for one line — mirrors MLUtils.parseLibSVMRecord."""
    items = line.split(" ")
    label = float(items[0])
    idx: List[int] = []
    val: List[float] = []
    prev = -1
    for item in items[1:]:
        if not item:
            continue
        i_s, v_s = item.split(":")
        i = int(i_s) - 1
        if not i > prev:
            raise ValueError(f"indices should be one-based and in ascending order; found current={i}, previous={prev}; line=\"{line}\"")
        prev = i
        idx.append(i)
        val.append(float(v_s))
    return label, np.asarray(idx, dtype=np.int32), np.asarray(val, dtype=np.float64)


def _input_files(path: str) -> List[str]:
    p = Path(path)
    if p.is_dir():
        files = sorted(str(f) for f in p.iterdir() if f.is_file() and not f.name.startswith(("_", ".")))
        if not files:
            raise FileNotFoundError(f"no part files under {path}")
        return files
    if not p.exists():
        raise FileNotFoundError(path)
    return [str(p)]


def parse_libsvm_file_csr(path: str, num_threads: Optional[int] = None):
    """Parse one file into CSR numpy arrays with the native parser:
    returns (labels f64[n], indptr i64[n+1], indices i32[nnz], values f64[nnz], max_index)."""
    rt = _native.runtime()
    err = C.create_string_buffer(1024)
    nt = num_threads or min(32, os.cpu_count() or 1)
    h = rt.afb_libsvm_parse(path.encode(), nt, err, len(err))
    if not h:
        raise ValueError(err.value.decode() or f"failed to parse {path}")
    try:
        n, nnz = rt.afb_libsvm_rows(h), rt.afb_libsvm_nnz(h)
        labels = np.empty(n, dtype=np.float64)
        indptr = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(nnz, dtype=np.int32)
        values = np.empty(nnz, dtype=np.float64)
        rt.afb_libsvm_copy(h, labels.ctypes.data, indptr.ctypes.data, indices.ctypes.data, values.ctypes.data)
        return labels, indptr, indices, values, int(rt.afb_libsvm_max_index(h))
    finally:
        rt.afb_libsvm_free(h)


def load_libsvm_csr(path: str, num_threads: Optional[int] = None):
    parts = [parse_libsvm_file_csr(f, num_threads) for f in _input_files(path)]
    if len(parts) == 1:
        return parts[0]
    labels = np.concatenate([p[0] for p in parts])
    indptrs, off = [parts[0][1]], parts[0][1][-1]
    for p in parts[1:]:
        indptrs.append(p[1][1:] + off)
        off += p[1][-1]
    return (labels, np.concatenate(indptrs), np.concatenate([p[2] for p in parts]),
            np.concatenate([p[3] for p in parts]), max(p[4] for p in parts))


def load_libsvm_file(scheduler, path: str, num_features: int = -1, num_partitions: Optional[int] = None,
                     dtype: torch.dtype = torch.float64, layout: str = "auto", density_threshold: float = 0.25,
                     device=None) -> ShardedDataset:
    """``MLUtils.loadLibSVMFile``: returns a batch-capable ``ShardedDataset`` whose records are
    ``LabeledPoint(label, SparseVector)``.  ``num_partitions`` plays ``minPartitions`` (contiguous
    splits, like Hadoop input splits); call ``.repartition(P)`` for the round-robin shuffle the
    drivers apply."""
    labels, indptr, indices, values, max_index = load_libsvm_csr(path)
    d = int(num_features) if num_features and num_features > 0 else max_index + 1
    if max_index >= d:
        raise ValueError(f"feature index {max_index} out of range for num_features={d}")
    n = labels.shape[0]
    P = int(num_partitions) if num_partitions else max(1, getattr(scheduler, "num_executors", 1))
    P = max(1, min(P, max(n, 1)))
    nnz = values.shape[0]
    density = nnz / max(1, n * max(d, 1))
    dense = layout == "dense" or (layout == "auto" and density >= density_threshold)
    bounds = [n * i // P for i in range(P + 1)]
    shards = []
    t_labels, t_indptr = torch.from_numpy(labels), torch.from_numpy(indptr)
    t_idx, t_val = torch.from_numpy(indices), torch.from_numpy(values)
    for p in range(P):
        a, b = bounds[p], bounds[p + 1]
        s, e = int(indptr[a]), int(indptr[b])
        sh = Shard(d, t_labels[a:b].to(dtype), None, (t_indptr[a:b + 1] - s).clone(), t_idx[s:e].clone(),
                   t_val[s:e].to(dtype), partition=p)
        if dense:
            sh = sh.to_dense()
        shards.append(sh.to(device) if device is not None else sh)
    ds = ShardedDataset.from_shards(scheduler, shards, f"libsvm:{os.path.basename(path)}")
    ds.num_features = d  # type: ignore[attr-defined]
    ds.density = density  # type: ignore[attr-defined]
    return ds


def _fmt(v: float) -> str:
    return repr(float(v))


def save_as_libsvm_file(data, dir_path: str) -> None:
    """``MLUtils.saveAsLibSVMFile``: a directory with one ``part-NNNNN`` file per partition and a
    ``_SUCCESS`` marker; ``label i1:v1 i2:v2 ...`` with 1-based indices over the *active* entries."""
    os.makedirs(dir_path, exist_ok=False)
    rt = _native.runtime()
    if isinstance(data, ShardedDataset) and data.batch_capable:
        for p in range(data.num_partitions):
            sh = data.batch(p)
            out = os.path.join(dir_path, f"part-{p:05d}")
            if sh is None:
                open(out, "w").close()
                continue
            c = sh.to_csr() if sh.is_dense else sh
            labels = np.ascontiguousarray(c.y.detach().cpu().double().numpy())
            indptr = np.ascontiguousarray(c.indptr.detach().cpu().numpy().astype(np.int64))
            indices = np.ascontiguousarray(c.indices.detach().cpu().numpy().astype(np.int32))
            values = np.ascontiguousarray(c.values.detach().cpu().double().numpy())
            err = C.create_string_buffer(512)
            rc = rt.afb_libsvm_write(out.encode(), labels.shape[0], labels.ctypes.data, indptr.ctypes.data,
                                     indices.ctypes.data, values.ctypes.data, err, len(err))
            if rc != 0:
                raise IOError(err.value.decode())
    else:
        parts = ([list(data.iterator(p)) for p in range(data.num_partitions)]
                 if isinstance(data, ShardedDataset) else [list(data)])
        for p, recs in enumerate(parts):
            with open(os.path.join(dir_path, f"part-{p:05d}"), "w") as f:
                for rec in recs:
                    lp = rec[0] if isinstance(rec, tuple) else rec
                    sv = lp.features.to_sparse() if hasattr(lp.features, "to_sparse") else lp.features
                    toks = [_fmt(lp.label)] + [f"{int(i) + 1}:{_fmt(v)}" for i, v in zip(sv.indices.tolist(), sv.values.tolist())]
                    f.write(" ".join(toks) + "\n")
    open(os.path.join(dir_path, "_SUCCESS"), "w").close()


class MLUtils:
    """Namespace with the reference's spellings."""
    loadLibSVMFile = staticmethod(load_libsvm_file)
    saveAsLibSVMFile = staticmethod(save_as_libsvm_file)
    load_libsvm_file = staticmethod(load_libsvm_file)
    save_as_libsvm_file = staticmethod(save_as_libsvm_file)

    @staticmethod
    def append_bias(v):
        """MLUtils.appendBias (MLUtils.scala:255): trailing 1.0 feature."""
        from ..linalg.vectors import DenseVector
        if isinstance(v, SparseVector):
            idx = torch.cat([v.indices, torch.tensor([v.size], dtype=torch.int32)])
            val = torch.cat([v.values, torch.ones(1, dtype=v.values.dtype)])
            return SparseVector(v.size + 1, idx, val, dtype=v.values.dtype)
        return DenseVector(torch.cat([v.values, torch.ones(1, dtype=v.values.dtype)]), dtype=v.values.dtype)

    @staticmethod
    def k_fold(ds, num_folds: int, seed: int = 42):
        """``MLUtils.kFold`` (MLUtils.scala:232-249): ``num_folds`` pairs ``(training, validation)``; fold f's validation set is
        the records whose per-record uniform draw falls in ``[(f-1)/k, f/k)`` (BernoulliCellSampler), its training set the
        complement — the k validation sets partition the data.  The draw is a counter hash of (seed, partition, position), so
        every fold sees the same draw for the same record and a retried task reproduces it."""
        from ..core.sampler import uniform01
        import numpy as np
        k = int(num_folds)
        if k < 2:
            raise ValueError("k_fold needs at least two folds")

        def cell(lo, hi, complement):
            def compute(p, t):
                recs = list(ds.iterator(p, t))
                if not recs:
                    return iter(())
                u = uniform01(int(seed) * 1_000_003 + p, np.arange(len(recs), dtype=np.int64))
                inside = (u >= lo) & (u < hi)
                keep = ~inside if complement else inside
                return (r for r, kp in zip(recs, keep.tolist()) if kp)
            return ds._derive(compute, None, "k_fold")
        return [(cell((f - 1) / k, f / k, True), cell((f - 1) / k, f / k, False)) for f in range(1, k + 1)]

    @staticmethod
    def load_vectors(sc, path: str, min_partitions=None):
        """``MLUtils.loadVectors`` (MLUtils.scala:197-205): text files written by ``RDD[Vector].saveAsTextFile``."""
        from ..linalg.vectors import Vectors
        return sc.text_file(path, min_partitions).filter(lambda l: l.strip() != "").map(Vectors.parse)

    @staticmethod
    def load_labeled_points(sc, path: str, min_partitions=None):
        """``MLUtils.loadLabeledPoints`` (MLUtils.scala:215-224): text files written by ``RDD[LabeledPoint].saveAsTextFile``
        (also the pre-1.0 ``label,f0 f1 ...`` lines)."""
        from ..linalg.vectors import LabeledPoint
        return sc.text_file(path, min_partitions).filter(lambda l: l.strip() != "").map(LabeledPoint.parse)

    kFold = k_fold
    loadVectors = load_vectors
    loadLabeledPoints = load_labeled_points

    @staticmethod
    def log1p_exp(x: float) -> float:
        """MLUtils.log1pExp (MLUtils.scala:548-554): numerically stable log(1+e^x)."""
        import math
        return x + math.log1p(math.exp(-x)) if x > 0 else math.log1p(math.exp(x))
