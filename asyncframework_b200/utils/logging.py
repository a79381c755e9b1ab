"""Logging setup.

The reference logs through a ``Logging`` trait over slf4j/log4j (core/internal/Logging.scala) configured by
``conf/log4j.properties`` (template: ``log4j.rootCategory=INFO, console`` + per-logger levels), and its samples silence it
with ``sc.setLogLevel("ERROR")`` while printing user-facing progress with println.  Here: the std ``logging`` tree under
``asyncframework_b200`` with the same default level, ``set_log_level`` (= ``sc.setLogLevel``), and ``configure(path)`` which
reads the same *properties* syntax — root level + appenders (``console`` / ``file``), a conversion pattern, and
``log4j.logger.<name>=<LEVEL>`` lines — so an existing ``log4j.properties`` keeps working for the parts that make sense.
"""
from __future__ import annotations

import logging
import os
import sys
from typing import Dict, Optional

_LEVELS = {"ALL": logging.NOTSET + 1, "TRACE": 5, "DEBUG": logging.DEBUG, "INFO": logging.INFO, "WARN": logging.WARNING,
           "WARNING": logging.WARNING, "ERROR": logging.ERROR, "FATAL": logging.CRITICAL, "OFF": logging.CRITICAL + 10}
DEFAULT_FORMAT = "%(asctime)s %(levelname)s %(name)s: %(message)s"


def _level(name: str) -> int:
    try:
        return _LEVELS[name.strip().upper()]
    except KeyError:
        raise ValueError(f"unknown log level {name!r} (use one of {sorted(_LEVELS)})") from None


def set_log_level(level: str = "ERROR") -> None:
    """``sc.setLogLevel`` (SparkContext.scala:370-377): valid levels ALL, DEBUG, ERROR, FATAL, INFO, OFF, TRACE, WARN."""
    logging.getLogger("asyncframework_b200").setLevel(_level(level))


def setup(level: Optional[str] = None) -> logging.Logger:
    level = level or os.environ.get("AFB_LOG_LEVEL", "ERROR")
    logging.basicConfig(format=DEFAULT_FORMAT)
    set_log_level(level)
    return logging.getLogger("asyncframework_b200")


def _pattern_to_format(pattern: str) -> str:
    """The common log4j conversion characters: %d date, %p level, %c logger, %m message, %t thread, %n newline."""
    import re
    out = pattern
    out = re.sub(r"%d(\{[^}]*\})?", "%(asctime)s", out)
    out = re.sub(r"%-?\d*p", "%(levelname)s", out)
    out = re.sub(r"%c(\{\d+\})?", "%(name)s", out)
    out = out.replace("%m", "%(message)s").replace("%t", "%(threadName)s").replace("%n", "")
    return out


def configure(path: Optional[str] = None, props: Optional[Dict[str, str]] = None) -> logging.Logger:
    """Apply a ``log4j.properties``-style configuration to the framework's logger tree.  Looks for
    ``$AFB_CONF_DIR/log4j.properties`` then ``conf/log4j.properties`` when no path is given; returns the root framework logger."""
    if props is None:
        if path is None:
            root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            for cand in (os.path.join(os.environ.get("AFB_CONF_DIR", ""), "log4j.properties"), os.path.join(root, "conf", "log4j.properties")):
                if cand and os.path.isfile(cand):
                    path = cand
                    break
        props = {}
        if path is not None:
            with open(path) as f:
                for line in f:
                    line = line.strip()
                    if line and not line.startswith(("#", "!")) and "=" in line:
                        k, v = line.split("=", 1)
                        props[k.strip()] = v.strip()
    base = logging.getLogger("asyncframework_b200")
    root_spec = props.get("log4j.rootCategory") or props.get("log4j.rootLogger") or "ERROR, console"
    parts = [p.strip() for p in root_spec.split(",")]
    base.setLevel(_level(parts[0]))
    for h in list(base.handlers):
        if getattr(h, "_afb_configured", False):
            base.removeHandler(h)
    for app in parts[1:] or ["console"]:
        cls = props.get(f"log4j.appender.{app}", "ConsoleAppender")
        fmt = _pattern_to_format(props.get(f"log4j.appender.{app}.layout.ConversionPattern", "%d %p %c: %m%n"))
        if "File" in cls:
            handler: logging.Handler = logging.FileHandler(props.get(f"log4j.appender.{app}.File", "afb.log"))
        else:
            target = props.get(f"log4j.appender.{app}.target", "System.err")
            handler = logging.StreamHandler(sys.stdout if target.endswith("out") else sys.stderr)
        handler.setFormatter(logging.Formatter(fmt))
        handler._afb_configured = True  # type: ignore[attr-defined]
        base.addHandler(handler)
    base.propagate = False
    for k, v in props.items():
        if k.startswith("log4j.logger."):
            name = k[len("log4j.logger."):]
            # reference logger names (org.apache.spark.*) address the whole framework tree
            target = "asyncframework_b200" if name.startswith(("org.apache.spark", "org.apache.async")) else name
            logging.getLogger(target).setLevel(_level(v.split(",")[0]))
    return base
