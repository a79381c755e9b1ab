"""``spark-shell`` / ``pyspark`` analogue: an interactive interpreter with ``sc`` (an :class:`AsyncSession`) predefined.

    python -m asyncframework_b200.shell --master local[4]
    python -m asyncframework_b200.shell --master spark://10.0.0.1:7077 --conf spark.ui.enabled=true

(repl/ of the reference; the same option resolution as ``asyncframework_b200.submit``.)
"""
from __future__ import annotations

import code
import sys

from . import submit


def main(argv=None) -> int:
    ns = submit.build_parser().parse_args((argv if argv is not None else sys.argv[1:]))
    conf = submit.resolve_conf(ns)
    from . import AsyncSession, barriers
    from .compat import ASYNCbroadcast, ASYNCcontext, MLUtils, Vectors
    import torch
    sc = AsyncSession(conf=conf)
    banner = (f"asyncframework_b200 shell — master {sc.master}, app {sc.app_name!r}, {sc.default_parallelism} executors"
              + (f", UI {sc.ui.url}" if getattr(sc, "ui", None) else "") + "\nsession available as 'sc'.")
    ns_vars = {"sc": sc, "torch": torch, "barriers": barriers, "ASYNCcontext": ASYNCcontext, "ASYNCbroadcast": ASYNCbroadcast,
               "MLUtils": MLUtils, "Vectors": Vectors}
    try:
        if ns.app:                                        # `shell -- -c "expr"`: evaluate and exit (scripting / tests)
            args = [a for a in ns.app if a != "--"]
            if args and args[0] == "-c":
                exec(args[1], ns_vars)                    # noqa: S102
                return 0
        code.interact(banner=banner, local=ns_vars)
    finally:
        sc.stop()
    return 0


if __name__ == "__main__":
    sys.exit(main())
