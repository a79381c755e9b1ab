"""Thin device shim: on CUDA devices every helper returns the real torch.cuda object (identical behaviour
to calling torch.cuda directly); on the CPU it returns inert stand-ins so the fused engine's host logic
(round bookkeeping, drivers, checkpointing, result pump) can run against the pure-Python round simulator
(`ops/round_sim.py`) in the CPU test-suite."""
from __future__ import annotations

import contextlib

import torch


class NullStream:
    cuda_stream = 0

    def synchronize(self): pass
    def wait_event(self, event): pass
    def wait_stream(self, stream): pass
    def record_event(self, event=None): return event or NullEvent()


class NullEvent:
    def __init__(self, enable_timing: bool = False):
        import time
        self._t = time.perf_counter()

    def record(self, stream=None):
        import time
        self._t = time.perf_counter()

    def synchronize(self): pass
    def wait(self, stream=None): pass
    def query(self): return True

    def elapsed_time(self, other) -> float:
        return (other._t - self._t) * 1e3


_NULL_STREAM = NullStream()


def is_cuda(device) -> bool:
    return torch.device(device).type == "cuda"


def current_stream(device):
    return torch.cuda.current_stream(device) if is_cuda(device) else _NULL_STREAM


def new_stream(device):
    return torch.cuda.Stream(device=device) if is_cuda(device) else NullStream()


def new_event(device, enable_timing: bool = False):
    return torch.cuda.Event(enable_timing=enable_timing) if is_cuda(device) else NullEvent(enable_timing)


def pin(t: torch.Tensor, device) -> torch.Tensor:
    return t.pin_memory() if is_cuda(device) else t


def device_sms(device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count if is_cuda(device) else 148


def synchronize(device) -> None:
    if is_cuda(device):
        torch.cuda.synchronize(device)


def set_device(device) -> None:
    if is_cuda(device):
        torch.cuda.set_device(device)


def stream_ctx(stream, device):
    return torch.cuda.stream(stream) if is_cuda(device) else contextlib.nullcontext()
