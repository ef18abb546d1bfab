"""A linter-free safety net (no pyflakes / ruff in the image): every global name a function of the package, bench.py,
__graft_entry__.py or tools/ loads must be defined at module level or be a builtin.  Catches typos in code paths that only run
on the GPU box (N > 1 branches of bench.py, error paths)."""
import ast
import builtins
import dis
import os
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _files():
    out = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for d in ("specfem3d_b200", "oracle", "tools", "tests", os.path.join("tests", "golden")):
        p = os.path.join(ROOT, d)
        out += sorted(os.path.join(p, f) for f in os.listdir(p) if f.endswith(".py"))
    return out


def _module_names(tree):
    names = set(dir(builtins)) | {"__file__", "__name__", "__doc__", "__builtins__", "__spec__", "__package__", "__annotations__", "__class__", "__qualname__", "__module__"}
    for node in ast.walk(tree):  # conservative: any name bound anywhere in the file by a module-level-capable statement
        if isinstance(node, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
            names.add(node.name)
        elif isinstance(node, (ast.Import, ast.ImportFrom)):
            for a in node.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.Name) and isinstance(node.ctx, (ast.Store, ast.Del)):
            names.add(node.id)
        elif isinstance(node, ast.Global):
            names.update(node.names)
    return names


def _code_objects(code):
    yield code
    for c in code.co_consts:
        if isinstance(c, types.CodeType):
            yield from _code_objects(c)


@pytest.mark.parametrize("path", _files(), ids=lambda p: os.path.relpath(p, ROOT))
def test_no_undefined_global_names(path):
    src = open(path).read()
    tree = ast.parse(src, path)
    known = _module_names(tree)
    missing = set()
    for code in _code_objects(compile(src, path, "exec")):
        for ins in dis.get_instructions(code):
            if ins.opname in ("LOAD_GLOBAL", "LOAD_NAME") and ins.argval not in known:
                missing.add((ins.argval, code.co_name))
    assert not missing, sorted(missing)
