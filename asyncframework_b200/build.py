"""In-tree native build: every C++/CUDA source is compiled into ``asyncframework_b200/lib``.

* ``libafb_runtime.so``  - g++:  LIBSVM parser, result queue, logical clock, worker table.
* ``libafb_kernels.so``  - nvcc ``-gencode arch=compute_100a,code=sm_100a -lineinfo``: every CUDA
  kernel (level-1, fused dense/CSR gradient, round kernels with in-kernel NVLink push/pull,
  tcgen05 evaluation GEMM, sampler, updaters, comm helpers).

Both expose a plain C ABI consumed with ctypes (no torch headers => seconds to build, no ABI
coupling; the .so files travel to the GPU box with the snapshot).  nvcc cross-compiles sm_100a
without a GPU, so this runs on the CPU box too.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path
from typing import List, Optional

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "csrc"
LIBDIR = Path(__file__).resolve().parent / "lib"

RUNTIME_SOURCES = ["runtime/libsvm_parser.cpp", "runtime/host_runtime.cpp"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def kernel_sources() -> List[Path]:
    return sorted((CSRC / "kernels").glob("*.cu"))


def _digest(paths: List[Path], extra: str = "") -> str:
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(p.name.encode())
        h.update(p.read_bytes())
    for p in sorted((CSRC / "include").glob("*")):
        h.update(p.read_bytes())
    return h.hexdigest()


def _run(cmd: List[str], verbose: bool) -> None:
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError(f"build command failed ({res.returncode}): {' '.join(cmd)}")
    if verbose and res.stdout.strip():
        print(res.stdout)


def find_nvcc() -> Optional[str]:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if cand and os.path.exists(cand) else None


def build_runtime(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(parents=True, exist_ok=True)
    out = LIBDIR / "libafb_runtime.so"
    srcs = [CSRC / s for s in RUNTIME_SOURCES]
    stamp = LIBDIR / ".runtime.sha"
    dg = _digest(srcs)
    if out.exists() and not force and stamp.exists() and stamp.read_text() == dg:
        return out
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    _run([cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall",
          *[str(s) for s in srcs], "-o", str(out)], verbose)
    stamp.write_text(dg)
    return out


def build_kernels(force: bool = False, verbose: bool = False, keep_ptx: bool = False) -> Optional[Path]:
    LIBDIR.mkdir(parents=True, exist_ok=True)
    out = LIBDIR / "libafb_kernels.so"
    srcs = kernel_sources()
    if not srcs:
        return None
    stamp = LIBDIR / ".kernels.sha"
    dg = _digest(srcs, "v1")
    if out.exists() and not force and stamp.exists() and stamp.read_text() == dg:
        return out
    nvcc = find_nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found; cannot build CUDA kernels")
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    common = ["-O3", "-std=c++17", "-lineinfo", *NVCC_ARCH, "-Xcompiler", "-fPIC",
              "-I", str(CSRC / "include"), "--expt-relaxed-constexpr", "-Xptxas", "-v"]
    objs = []
    procs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        odg = objdir / (s.stem + ".sha")
        sdg = _digest([s], "obj")
        objs.append(o)
        if o.exists() and odg.exists() and odg.read_text() == sdg and not force:
            continue
        cmd = [nvcc, *common, "-c", str(s), "-o", str(o)]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        procs.append((s, o, odg, sdg, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log_lines = []
    for s, o, odg, sdg, p in procs:
        txt, _ = p.communicate()
        log_lines.append(f"==== {s.name} ====\n{txt}")
        if p.returncode != 0:
            sys.stderr.write(txt)
            raise RuntimeError(f"nvcc failed for {s}")
        odg.write_text(sdg)
    if log_lines:
        (ROOT / "build" / "ptxas.log").write_text("\n".join(log_lines))
    _run([nvcc, "-shared", *NVCC_ARCH, "-o", str(out), *[str(o) for o in objs]], verbose)
    stamp.write_text(dg)
    return out


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_runtime(force=force, verbose=verbose)
    build_kernels(force=force, verbose=verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built:", sorted(p.name for p in LIBDIR.glob("*.so")))
