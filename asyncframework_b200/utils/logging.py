"""Logging setup: the reference silences log4j with ``sc.setLogLevel("ERROR")`` and
prints user-facing progress with println; we use the std ``logging`` tree under
``asyncframework_b200`` with the same default level."""
from __future__ import annotations

import logging
import os


def set_log_level(level: str = "ERROR") -> None:
    logging.getLogger("asyncframework_b200").setLevel(getattr(logging, level.upper(), logging.ERROR))


def setup(level: str = None) -> logging.Logger:
    level = level or os.environ.get("AFB_LOG_LEVEL", "ERROR")
    logging.basicConfig(format="%(asctime)s %(levelname)s %(name)s: %(message)s")
    set_log_level(level)
    return logging.getLogger("asyncframework_b200")
