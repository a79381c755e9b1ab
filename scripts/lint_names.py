#!/usr/bin/env python
"""Tiny static checker (no pyflakes in this image): flags names that are read inside a function but are
assigned/imported *later* in that same function without a prior binding (the UnboundLocalError class),
and module-level names that are used but never bound/imported."""
import ast
import builtins
import sys


class FuncCheck(ast.NodeVisitor):
    def __init__(self, path):
        self.path = path
        self.problems = []

    def check_function(self, fn):
        # names bound anywhere in this function body (excluding nested defs)
        bound_lines = {}
        params = {a.arg for a in fn.args.args + fn.args.kwonlyargs + fn.args.posonlyargs}
        if fn.args.vararg: params.add(fn.args.vararg.arg)
        if fn.args.kwarg: params.add(fn.args.kwarg.arg)
        globals_decl = set()

        def walk(node):
            for ch in ast.iter_child_nodes(node):
                if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef, ast.Lambda, ast.ClassDef)):
                    if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                        bound_lines.setdefault(ch.name, ch.lineno)
                    continue
                if isinstance(ch, (ast.Global, ast.Nonlocal)):
                    globals_decl.update(ch.names)
                if isinstance(ch, (ast.Import, ast.ImportFrom)):
                    for a in ch.names:
                        nm = (a.asname or a.name).split(".")[0]
                        bound_lines[nm] = min(bound_lines.get(nm, 10**9), ch.lineno)
                if isinstance(ch, ast.Name) and isinstance(ch.ctx, (ast.Store, ast.Del)):
                    bound_lines[ch.id] = min(bound_lines.get(ch.id, 10**9), ch.lineno)
                if isinstance(ch, ast.ExceptHandler) and ch.name:
                    bound_lines[ch.name] = min(bound_lines.get(ch.name, 10**9), ch.lineno)
                walk(ch)
        walk(fn)

        def walk_loads(node):
            for ch in ast.iter_child_nodes(node):
                if isinstance(ch, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                    continue
                if isinstance(ch, ast.Name) and isinstance(ch.ctx, ast.Load):
                    if ch.id in bound_lines and ch.id not in params and ch.id not in globals_decl and ch.lineno < bound_lines[ch.id]:
                        self.problems.append(f"{self.path}:{ch.lineno}: '{ch.id}' used before its local binding at line {bound_lines[ch.id]} in {fn.name}()")
                walk_loads(ch)
        walk_loads(fn)

    def visit_FunctionDef(self, node):
        self.check_function(node)
        self.generic_visit(node)

    visit_AsyncFunctionDef = visit_FunctionDef


def main(paths):
    import os
    files = []
    for p in paths:
        if os.path.isdir(p):
            for r, _, fs in os.walk(p):
                files += [os.path.join(r, f) for f in fs if f.endswith(".py")]
        else:
            files.append(p)
    bad = 0
    for f in sorted(files):
        tree = ast.parse(open(f).read(), f)
        c = FuncCheck(f)
        c.visit(tree)
        for pr in c.problems:
            print(pr)
            bad += 1
    print(f"checked {len(files)} files, {bad} problem(s)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:] or ["asyncframework_b200", "bench.py", "benchmarks", "baseline", "tests", "__graft_entry__.py"]))
