"""GPU clock / throttle-reason sampler (NVML in a background thread) for benchmark hygiene."""
from __future__ import annotations

import statistics
import threading
import time
from typing import Dict, List, Optional

_REASONS = {
    0x0000000000000001: "gpu_idle",
    0x0000000000000002: "applications_clocks_setting",
    0x0000000000000004: "sw_power_cap",
    0x0000000000000008: "hw_slowdown",
    0x0000000000000010: "sync_boost",
    0x0000000000000020: "sw_thermal_slowdown",
    0x0000000000000040: "hw_thermal_slowdown",
    0x0000000000000080: "hw_power_brake_slowdown",
    0x0000000000000100: "display_clock_setting",
}
BAD = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


class ClockSampler:
    def __init__(self, device_index: int = 0, period_s: float = 0.01):
        self.device_index = device_index
        self.period_s = period_s
        self.samples: List[int] = []
        self.power: List[float] = []
        self.reasons: Dict[str, int] = {}
        self.max_mhz: Optional[int] = None
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._ok = True
        except Exception:  # noqa: BLE001
            self._ok = False

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                except Exception:  # noqa: BLE001
                    pass
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:  # noqa: BLE001
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in _REASONS.items():
                    if mask & bit:
                        self.reasons[name] = self.reasons.get(name, 0) + 1
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period_s)

    def start(self) -> "ClockSampler":
        if self._ok:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def stop(self) -> Dict[str, object]:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
        reasons = sorted(r for r in self.reasons if r not in ("gpu_idle",))
        out = {
            "sm_mhz": (statistics.median(self.samples) if self.samples else None),
            "sm_max_mhz": self.max_mhz,
            "reasons": reasons,
            "samples": len(self.samples),
            "power_w_max": (max(self.power) if self.power else None),
        }
        out["rejected"] = bool(BAD & set(reasons))
        return out
