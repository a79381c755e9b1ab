"""Micro-benchmarks of the C++ host runtime (no GPU needed): LIBSVM ingest throughput (mmap multithreaded parser vs a
pure-Python/NumPy parser of the same format), result-queue and worker-table operation rates.

    python benchmarks/host_runtime_bench.py [--rows 200000] [--features 784] [--out profiles/host_runtime.json]

K14 of SURVEY.md §2.7 (the reference parses LIBSVM with per-line Scala string splitting,
mllib/util/MLUtils.scala:74-131) and the latency-critical host pieces of §2.4 (iv)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200 import _native  # noqa: E402
from asyncframework_b200.data.libsvm import parse_libsvm_file_csr, parse_libsvm_record  # noqa: E402


def write_dense_libsvm(path: str, rows: int, d: int, density: float, seed: int = 0) -> int:
    rng = np.random.default_rng(seed)
    with open(path, "w") as f:
        for start in range(0, rows, 4096):
            n = min(4096, rows - start)
            mask = rng.random((n, d)) < density
            vals = rng.integers(1, 256, size=(n, d))
            labels = rng.integers(0, 10, size=n)
            lines = []
            for i in range(n):
                idx = np.nonzero(mask[i])[0]
                lines.append(str(labels[i]) + " " + " ".join(f"{j + 1}:{vals[i, j]}" for j in idx))
            f.write("\n".join(lines) + "\n")
    return os.path.getsize(path)


def bench_parser(rows: int, d: int, density: float):
    out = {}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "data.libsvm")
        nbytes = write_dense_libsvm(path, rows, d, density)
        out["file_mb"] = round(nbytes / 1e6, 1)
        parse_libsvm_file_csr(path, num_threads=os.cpu_count() or 1)          # warm-up: page cache, library load
        for nt in sorted({1, 2, 4, os.cpu_count() or 1}):
            ts = []
            for _ in range(7):
                t0 = time.perf_counter()
                labels, indptr, indices, values, mx = parse_libsvm_file_csr(path, num_threads=nt)
                ts.append(time.perf_counter() - t0)
            ts.sort()
            out[f"native_{nt}t_mb_s"] = round(nbytes / 1e6 / ts[len(ts) // 2], 1)      # median of 7 (file -> CSR numpy arrays)
        out["rows"], out["nnz"] = int(len(labels)), int(len(values))
        # pure-Python reference of the same grammar on a slice of the file (it is ~100x slower)
        with open(path) as f:
            lines = [next(f) for _ in range(min(rows, 5000))]
        sl_bytes = sum(len(x) for x in lines)
        t0 = time.perf_counter()
        for ln in lines:
            parse_libsvm_record(ln)
        out["python_mb_s"] = round(sl_bytes / 1e6 / (time.perf_counter() - t0), 2)
    return out


def bench_queue(n_producers: int = 4, per_producer: int = 1_000_000):
    """C++ threads on both sides (`afb_queue_bench`): the rate the queue itself sustains, single take vs batch take."""
    rt = _native.runtime()
    total = n_producers * per_producer
    return {"producers": n_producers, "records": total,
            "single_take_mrec_s": round(total / rt.afb_queue_bench(n_producers, per_producer, 1) / 1e6, 2),
            "batch64_take_mrec_s": round(total / rt.afb_queue_bench(n_producers, per_producer, 64) / 1e6, 2)}


def bench_table(P: int = 64, ops: int = 300_000):
    rt = _native.runtime()
    tb = rt.afb_table_create(P)
    t0 = time.perf_counter()
    for i in range(ops):
        rt.afb_table_task_done(tb, i % P, 1, 0.5)
    dt_w = time.perf_counter() - t0
    t0 = time.perf_counter()
    for i in range(ops):
        rt.afb_table_available_workers(tb)
    dt_r = time.perf_counter() - t0
    rt.afb_table_destroy(tb)
    return {"P": P, "task_done_mops_s": round(ops / dt_w / 1e6, 3), "available_workers_scan_mops_s": round(ops / dt_r / 1e6, 3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000)
    ap.add_argument("--features", type=int, default=784)
    ap.add_argument("--density", type=float, default=0.25)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = {"host": {"cpus": os.cpu_count()}, "libsvm_parser": bench_parser(a.rows, a.features, a.density),
           "result_queue": bench_queue(), "worker_table": bench_table()}
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
