"""CPU: the product package never imports, links or executes anything under oracle/, and has no
numpy fallback for amplitude work."""
import ast
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pennylane_b200")


def _py_files():
    for d, _, fs in os.walk(PKG):
        for f in fs:
            if f.endswith(".py"):
                yield os.path.join(d, f)


def test_no_oracle_imports_in_package():
    for path in _py_files():
        tree = ast.parse(open(path).read(), path)
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                assert not (n == "oracle" or n.startswith("oracle.")), f"{path} imports {n}"
        src = open(path).read()
        assert "import oracle" not in src and "from oracle" not in src, path
        assert "oracle/" not in src and "'oracle'" not in src and '"oracle"' not in src, path


def test_importing_package_does_not_load_oracle():
    code = ("import sys; sys.path.insert(0, %r); import pennylane_b200, pennylane_b200.adjoint, "
            "pennylane_b200.simulate; assert not any(m == 'oracle' or m.startswith('oracle.') "
            "for m in sys.modules), 'oracle imported'") % ROOT
    subprocess.run([sys.executable, "-c", code], check=True)


def test_csrc_has_no_reference_to_oracle():
    for f in os.listdir(os.path.join(PKG, "csrc")):
        if f.endswith((".cu", ".cuh", ".h")):
            src = open(os.path.join(PKG, "csrc", f)).read()
            for line in src.splitlines():
                if "#include" in line:
                    assert "oracle" not in line.lower(), (f, line)
