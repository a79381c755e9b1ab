"""Communication layer: symmetric HBM buffers whose peer (and NVLS multicast) addresses are
handed to the hand-written kernels.

This replaces the reference's Netty RPC + TorrentBroadcast + task-result transfer
(core/rpc/netty/*, core/network/netty/NettyBlockTransferService.scala,
core/broadcast/TorrentBroadcast.scala) with plain loads/stores/reductions on NVLink-mapped
addresses issued from inside the compute kernels.  ``torch.distributed`` is used only for the
bootstrap rendezvous (handle exchange, barriers).

Providers:
  * ``symm_mem`` — ``torch.distributed._symmetric_memory`` (CUDA VMM + fd passing); also yields the
    NVLS multicast address when the fabric supports it.
  * ``ipc``      — our own CUDA-IPC path (csrc/kernels/misc.cu: cudaIpcGetMemHandle /
    cudaIpcOpenMemHandle with lazy peer access); handles travel over the process group's store.
  * ``local``    — single process: every "peer" pointer is the local pointer (one GPU, or several
    logical workers sharing a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class SymmetricBuffer:
    local: torch.Tensor                 # uint8 view of this rank's copy
    ptrs: List[int]                     # address of every rank's copy as mapped into THIS process
    rank: int
    world_size: int
    multicast_ptr: int = 0              # NVLS multicast alias (0 = unavailable)
    provider: str = "local"
    _keep: list = field(default_factory=list)

    @property
    def nbytes(self) -> int:
        return int(self.local.numel())

    def barrier(self) -> None:
        if self.world_size > 1:
            torch.cuda.synchronize()
            dist.barrier()


class _RawCudaBuffer:
    """cudaMalloc'ed buffer exposed through __cuda_array_interface__ (so cudaIpcGetMemHandle sees
    an allocation base, not a slice of torch's caching allocator pool)."""

    def __init__(self, nbytes: int, device: torch.device):
        self.nbytes = nbytes
        self.device = device
        rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
        self._rt = rt
        ptr = C.c_void_p()
        with torch.cuda.device(device):
            rc = rt.cudaMalloc(C.byref(ptr), C.c_size_t(nbytes))
            if rc != 0:
                raise RuntimeError(f"cudaMalloc({nbytes}) failed: {rc}")
            rt.cudaMemset(ptr, 0, C.c_size_t(nbytes))
        self.ptr = int(ptr.value)
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}

    def tensor(self) -> torch.Tensor:
        return torch.as_tensor(self, device=self.device)


def _alloc_local(nbytes: int, device) -> SymmetricBuffer:
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    return SymmetricBuffer(t, [t.data_ptr()], 0, 1, 0, "local")


def _alloc_symm_mem(nbytes: int, device, group) -> SymmetricBuffer:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
    hdl = symm_mem.rendezvous(t, group=group if group is not None else dist.group.WORLD)
    t.zero_()
    ptrs = [int(p) for p in hdl.buffer_ptrs]
    mc = 0
    try:
        if getattr(hdl, "has_multicast_support", False) if not callable(getattr(hdl, "has_multicast_support", None)) else hdl.has_multicast_support:
            mc = int(hdl.multicast_ptr or 0)
    except Exception:
        mc = 0
    buf = SymmetricBuffer(t, ptrs, dist.get_rank(group), dist.get_world_size(group), mc, "symm_mem", [hdl])
    return buf


def _alloc_ipc(nbytes: int, device, group) -> SymmetricBuffer:
    from .._native import kernels
    lib = kernels()
    raw = _RawCudaBuffer(nbytes, torch.device(device))
    t = raw.tensor()
    handle = (C.c_ubyte * 64)()
    rc = lib.afb_ipc_get_handle(C.c_void_p(raw.ptr), handle)
    if rc != 0:
        raise RuntimeError(f"cudaIpcGetMemHandle failed: {rc}")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    gathered: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(gathered, bytes(handle), group=group)
    ptrs: List[int] = []
    for r in range(world):
        if r == rank:
            ptrs.append(raw.ptr)
            continue
        hb = (C.c_ubyte * 64).from_buffer_copy(gathered[r])
        out = C.c_void_p()
        rc = lib.afb_ipc_open_handle(hb, C.byref(out))
        if rc != 0:
            raise RuntimeError(f"cudaIpcOpenMemHandle(rank {r}) failed: {rc}")
        ptrs.append(int(out.value))
    dist.barrier(group=group)
    return SymmetricBuffer(t, ptrs, rank, world, 0, "ipc", [raw])


def allocate_symmetric(nbytes: int, device, group=None, provider: str = "auto") -> SymmetricBuffer:
    """Allocate ``nbytes`` on every rank and map all copies into this process."""
    nbytes = (int(nbytes) + 4095) // 4096 * 4096
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return _alloc_local(nbytes, device)
    provider = os.environ.get("AFB_SYMM_PROVIDER", provider)
    errors = []
    if provider in ("auto", "symm_mem"):
        try:
            return _alloc_symm_mem(nbytes, device, group)
        except Exception as e:  # noqa: BLE001
            errors.append(f"symm_mem: {e!r}")
            if provider == "symm_mem":
                raise
    try:
        return _alloc_ipc(nbytes, device, group)
    except Exception as e:  # noqa: BLE001
        errors.append(f"ipc: {e!r}")
        raise RuntimeError("no symmetric-memory provider worked: " + "; ".join(errors))


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world) from the torchrun env; initialises the process group if needed."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    return rank, local_rank, world


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
