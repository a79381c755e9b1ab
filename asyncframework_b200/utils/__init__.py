from .checkpoint import load_checkpoint, save_checkpoint
from .clocks import ClockSampler
from .tracing import TraceRecorder, cuda_timer, nvtx_range
from .watchdog import DeviceWatchdog, HeartbeatWatchdog

__all__ = ["load_checkpoint", "save_checkpoint", "ClockSampler", "TraceRecorder", "cuda_timer", "nvtx_range",
           "DeviceWatchdog", "HeartbeatWatchdog"]
