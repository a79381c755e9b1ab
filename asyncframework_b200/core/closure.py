""""Task not serializable" diagnostics (``ClosureCleaner.ensureSerializable``, core/util/ClosureCleaner.scala:34 and the
``SerializationDebugger`` that prints the reference chain to the offending object).

cloudpickle ships closures by value: captured variables, referenced globals and — a trap — sub-modules it believes a referenced
module attribute belongs to (``from ..ops import x`` inside a nested function makes it chase ``torch.ops``).  When a task binary
cannot be pickled, :func:`explain_pickle_failure` walks the object graph (closure cells, referenced globals, defaults, attributes,
containers) and returns the chain of references that leads to the first unpicklable leaf, so the error names the culprit
instead of ``TypeError: cannot pickle '_thread.lock' object``.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

import cloudpickle


class TaskNotSerializable(TypeError):
    """Raised by the cluster schedulers when a job's (lineage, function) pair cannot be shipped to executor processes."""


def _picklable(obj: Any) -> Optional[BaseException]:
    try:
        cloudpickle.dumps(obj)
        return None
    except BaseException as exc:  # noqa: BLE001
        return exc


def _children(obj: Any) -> List[Tuple[str, Any]]:
    out: List[Tuple[str, Any]] = []
    code = getattr(obj, "__code__", None)
    if code is not None:
        if getattr(obj, "__closure__", None):
            out += [(f"<closure {n}>", c.cell_contents) for n, c in zip(code.co_freevars, obj.__closure__) if _has_contents(c)]
        g = getattr(obj, "__globals__", {})
        names = set(code.co_names)
        for const in code.co_consts:                       # nested functions / lambdas / comprehensions
            if hasattr(const, "co_names"):
                names |= set(const.co_names)
        out += [(f"<global {n}>", g[n]) for n in sorted(names) if n in g and not isinstance(g[n], type(cloudpickle))]
        out += [(f"<default {i}>", d) for i, d in enumerate(obj.__defaults__ or ())]
        out += [(f"<kwdefault {k}>", v) for k, v in (obj.__kwdefaults__ or {}).items()]
        return out
    if isinstance(obj, dict):
        return [(f"[{k!r}]", v) for k, v in obj.items()]
    if isinstance(obj, (list, tuple, set, frozenset)):
        return [(f"[{i}]", v) for i, v in enumerate(obj)]
    if hasattr(obj, "__getstate__") and not isinstance(obj, type):
        try:
            st = obj.__getstate__()
            if isinstance(st, dict):
                return [(f".{k}", v) for k, v in st.items()]
        except Exception:  # noqa: BLE001
            pass
    if hasattr(obj, "__dict__") and not isinstance(obj, type):
        return [(f".{k}", v) for k, v in vars(obj).items()]
    return []


def _has_contents(cell) -> bool:
    try:
        cell.cell_contents
        return True
    except ValueError:
        return False


def explain_pickle_failure(obj: Any, name: str = "task", max_depth: int = 12) -> Optional[str]:
    """``None`` when ``obj`` pickles; otherwise a multi-line description of the reference chain to the unpicklable leaf."""
    exc = _picklable(obj)
    if exc is None:
        return None
    chain: List[str] = [f"{name}: {type(obj).__name__}"]
    seen = {id(obj)}
    cur = obj
    for _ in range(max_depth):
        bad = None
        for label, child in _children(cur):
            if id(child) in seen:
                continue
            if _picklable(child) is not None:
                bad = (label, child)
                break
        if bad is None:
            break
        label, cur = bad
        seen.add(id(cur))
        desc = getattr(cur, "__qualname__", None) or type(cur).__name__
        chain.append(f"{label}: {type(cur).__module__}.{desc}" if not hasattr(cur, "__code__") else f"{label}: function {desc}")
    lines = [f"{'  ' * i}{'-> ' if i else ''}{c}" for i, c in enumerate(chain)]
    hint = ""
    if type(cur).__name__ == "_Ops" or "torch._ops" in type(cur).__module__:
        hint = ("\nhint: `from <pkg>.ops import x` inside a nested function makes cloudpickle follow `torch.ops`; import the leaf module "
                "(`from <pkg>.ops.x import f`) or call a module-level helper instead.")
    elif "lock" in type(cur).__name__.lower():
        hint = "\nhint: a threading lock was captured; reach it through a module-level function (pickled by reference) or drop it in __getstate__."
    return f"{exc!r}\nreference chain to the object that cannot be pickled:\n" + "\n".join(lines) + hint
