"""ctypes bindings for the in-tree native libraries.

``runtime()`` -> libafb_runtime.so (always required; built on demand with g++).
``kernels()`` -> libafb_kernels.so (CUDA; required whenever a CUDA device is present — we fail
loudly rather than silently falling back to eager PyTorch on a GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

LIBDIR = Path(__file__).resolve().parent / "lib"
_lock = threading.Lock()
_runtime: Optional[C.CDLL] = None
_kernels: Optional[C.CDLL] = None


class NativeLibraryError(RuntimeError):
    pass


class Record(C.Structure):
    _fields_ = [("payload_id", C.c_int64), ("ts", C.c_int64), ("staleness", C.c_int64),
                ("batch_size", C.c_int64), ("worker", C.c_int32), ("flags", C.c_int32),
                ("t_submit_ms", C.c_double), ("t_finish_ms", C.c_double)]


def _sig(lib, name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


def runtime() -> C.CDLL:
    global _runtime
    if _runtime is not None:
        return _runtime
    with _lock:
        if _runtime is not None:
            return _runtime
        path = LIBDIR / "libafb_runtime.so"
        if not path.exists() or os.environ.get("AFB_REBUILD"):
            from .build import build_runtime
            build_runtime()
        lib = C.CDLL(str(path))
        vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        _sig(lib, "afb_libsvm_parse", vp, [C.c_char_p, C.c_int, C.c_char_p, C.c_int])
        _sig(lib, "afb_libsvm_rows", i64, [vp])
        _sig(lib, "afb_libsvm_nnz", i64, [vp])
        _sig(lib, "afb_libsvm_max_index", i32, [vp])
        _sig(lib, "afb_libsvm_copy", None, [vp, vp, vp, vp, vp])
        _sig(lib, "afb_libsvm_free", None, [vp])
        _sig(lib, "afb_libsvm_write", C.c_int, [C.c_char_p, i64, vp, vp, vp, vp, C.c_char_p, C.c_int])
        _sig(lib, "afb_queue_create", vp, [])
        _sig(lib, "afb_queue_destroy", None, [vp])
        _sig(lib, "afb_queue_put", None, [vp, C.POINTER(Record)])
        _sig(lib, "afb_queue_take", C.c_int, [vp, C.POINTER(Record), f64])
        _sig(lib, "afb_queue_take_batch", C.c_int, [vp, C.POINTER(Record), C.c_int, f64])
        _sig(lib, "afb_queue_bench", f64, [C.c_int, i64, C.c_int])
        _sig(lib, "afb_queue_size", i64, [vp])
        _sig(lib, "afb_queue_close", None, [vp])
        _sig(lib, "afb_queue_reopen", None, [vp])
        _sig(lib, "afb_clock_create", vp, [])
        _sig(lib, "afb_clock_destroy", None, [vp])
        _sig(lib, "afb_clock_get", i64, [vp])
        _sig(lib, "afb_clock_add", i64, [vp, i64])
        _sig(lib, "afb_clock_set", None, [vp, i64])
        _sig(lib, "afb_table_create", vp, [C.c_int])
        _sig(lib, "afb_table_destroy", None, [vp])
        _sig(lib, "afb_table_size", C.c_int, [vp])
        _sig(lib, "afb_table_set_state", None, [vp, C.c_int, i64, f64, C.c_int, i64])
        _sig(lib, "afb_table_task_done", i64, [vp, C.c_int, i64, f64])
        _sig(lib, "afb_table_set_available", None, [vp, C.c_int, C.c_int])
        _sig(lib, "afb_table_set_alive", None, [vp, C.c_int, C.c_int])
        _sig(lib, "afb_table_heartbeat", None, [vp, C.c_int])
        _sig(lib, "afb_table_has_entry", C.c_int, [vp, C.c_int])
        _sig(lib, "afb_table_available", C.c_int, [vp, C.c_int])
        _sig(lib, "afb_table_alive", C.c_int, [vp, C.c_int])
        _sig(lib, "afb_table_staleness", i64, [vp, C.c_int])
        _sig(lib, "afb_table_avg_ms", f64, [vp, C.c_int])
        _sig(lib, "afb_table_num_tasks", i64, [vp, C.c_int])
        _sig(lib, "afb_table_heartbeat_age_ms", f64, [vp, C.c_int])
        _sig(lib, "afb_table_available_workers", C.c_int, [vp])
        _sig(lib, "afb_table_max_staleness", i64, [vp])
        _sig(lib, "afb_now_ms", f64, [])
        _runtime = lib
        return lib


def kernels_path() -> Path:
    return LIBDIR / "libafb_kernels.so"


def kernels() -> C.CDLL:
    """Load the CUDA kernel library; raises NativeLibraryError if it is missing."""
    global _kernels
    if _kernels is not None:
        return _kernels
    with _lock:
        if _kernels is not None:
            return _kernels
        path = kernels_path()
        if not path.exists():
            raise NativeLibraryError(
                f"{path} is missing: run `python -m asyncframework_b200.build` (nvcc sm_100a). "
                "asyncframework_b200 never falls back to eager PyTorch for its CUDA hot ops.")
        try:
            lib = C.CDLL(str(path))
        except OSError as e:  # e.g. libcuda.so.1 absent on the CPU-only build box
            raise NativeLibraryError(f"cannot load {path}: {e}") from e
        from .ops import _signatures
        _signatures.declare(lib)
        _kernels = lib
        return lib


def kernels_available() -> bool:
    try:
        kernels()
        return True
    except NativeLibraryError:
        return False
