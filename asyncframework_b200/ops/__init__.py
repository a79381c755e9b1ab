"""Hand-written sm_100a CUDA kernels and their Python launchers.

Every hot op of SURVEY.md §2.7 has a kernel in csrc/kernels/ compiled into
``asyncframework_b200/lib/libafb_kernels.so`` (see ``asyncframework_b200.build``).  The launchers
here take torch CUDA tensors, pass raw pointers + the current stream through ctypes, and never
fall back to eager PyTorch: if the library is missing on a GPU box they raise.
"""
from .._native import kernels, kernels_available  # noqa: F401


def check(rc: int, what: str) -> None:
    if rc != 0:
        import torch
        raise RuntimeError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)
