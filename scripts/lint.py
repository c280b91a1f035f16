#!/usr/bin/env python
"""`make lint` (reference: scripts/cpp_lint.py via `make lint`).  No third-party linters are assumed:
  * Python: every file must compile; imports that are never used (and not re-exported via `# noqa`) are reported;
    lines longer than 130 columns (170 in tests / benchmarks) and tabs are reported;
  * C++ / CUDA: lines longer than 150 columns, tabs and trailing whitespace are reported.
Exit code 1 if anything is reported."""
import ast
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PY_DIRS = ["poseidon_b200", "tests", "benchmarks", "scripts", "bench.py", "__graft_entry__.py"]
CC_DIRS = ["csrc", "csrc_host"]
SKIP = {"_ext", "_ext_exp", "__pycache__", "wt"}


def files(dirs, exts):
    for d in dirs:
        p = os.path.join(ROOT, d)
        if os.path.isfile(p):
            yield p
            continue
        for base, sub, names in os.walk(p):
            sub[:] = [s for s in sub if s not in SKIP]
            for n in names:
                if n.endswith(exts):
                    yield os.path.join(base, n)


def unused_imports(tree, src_lines):
    imported = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            for a in node.names:
                imported[(a.asname or a.name).split(".")[0]] = node.lineno
        elif isinstance(node, ast.ImportFrom):
            for a in node.names:
                if a.name != "*":
                    imported[a.asname or a.name] = node.lineno
    used = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.Name):
            used.add(node.id)
        elif isinstance(node, ast.Attribute):
            pass
    # names listed in __all__ or referenced inside string annotations count as used
    text = "\n".join(src_lines)
    out = []
    for name, line in imported.items():
        if name in used or name == "annotations":
            continue
        if "noqa" in src_lines[line - 1]:
            continue
        # multi-line import statements: honour a noqa anywhere in the statement
        j = line - 1
        stmt = src_lines[j]
        while stmt.count("(") > stmt.count(")") and j + 1 < len(src_lines):
            j += 1
            stmt += src_lines[j]
        if "noqa" in stmt:
            continue
        if f'"{name}"' in text or f"'{name}'" in text:
            continue
        out.append((line, name))
    return out


def main():
    problems = []
    for f in files(PY_DIRS, (".py",)):
        rel = os.path.relpath(f, ROOT)
        src = open(f, encoding="utf-8").read()
        lines = src.split("\n")
        try:
            tree = ast.parse(src, filename=rel)
        except SyntaxError as e:
            problems.append(f"{rel}:{e.lineno}: syntax error: {e.msg}")
            continue
        if not rel.endswith("__init__.py"):
            for line, name in unused_imports(tree, lines):
                problems.append(f"{rel}:{line}: unused import '{name}'")
        for i, l in enumerate(lines, 1):
            limit = 130 if rel.startswith("poseidon_b200") or "/" not in rel else 170      # tests / benchmarks: tables
            if len(l) > limit and "http" not in l and "noqa" not in l:
                problems.append(f"{rel}:{i}: line too long ({len(l)} > {limit})")
            if "\t" in l:
                problems.append(f"{rel}:{i}: tab character")
    for f in files(CC_DIRS, (".cu", ".cuh", ".cpp", ".h")):
        rel = os.path.relpath(f, ROOT)
        for i, l in enumerate(open(f, encoding="utf-8").read().split("\n"), 1):
            if len(l) > 150:
                problems.append(f"{rel}:{i}: line too long ({len(l)} > 150)")
            if "\t" in l:
                problems.append(f"{rel}:{i}: tab character")
            if l != l.rstrip():
                problems.append(f"{rel}:{i}: trailing whitespace")
    for p in problems:
        print(p)
    print(f"lint: {len(problems)} problem(s)")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
