"""Static conformance of ``pennylane_b200/pl_plugin.py`` (the genuine ``pennylane.devices.Device``
subclass) against the reference SOURCES.  PennyLane cannot be imported in this container
(autograd / autoray / rustworkx are absent), so the plugin cannot be executed here; what can be
checked without importing anything is checked with ``ast``:

  * every name the plugin imports from ``pennylane...`` is defined (or re-exported) by the module
    of that path in /root/reference;
  * every ``Device`` method the plugin overrides has the reference's parameter list
    (devices/device_api.py);
  * ``preprocess_transforms`` adds the same transforms, in the same order, with the same keyword
    names, under the same conditions as ``DefaultQubit.preprocess_transforms``
    (devices/default_qubit.py:611-679) and ``_add_adjoint_transforms`` (:315-349), minus
    ``validate_multiprocessing_workers``;
  * ``setup_execution_config`` / ``_setup_mcm_config`` resolve the same fields.

Skipped where the reference tree is not present (the GPU box)."""
import ast
import os

import pytest

REF = "/root/reference/pennylane"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "pennylane_b200", "pl_plugin.py")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present")


def _parse(path):
    with open(path) as f:
        return ast.parse(f.read())


def _module_file(dotted):
    rel = dotted.split(".")[1:]
    base = os.path.join(REF, *rel)
    if os.path.isdir(base):
        return os.path.join(base, "__init__.py")
    return base + ".py"


def _top_level_names(tree, path):
    """Names bound at module level: defs, classes, assignments, imports (incl. ``from x import *``
    resolved one level for package __init__ files)."""
    names = set()
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            names.add(node.name)
        elif isinstance(node, ast.Assign):
            for t in node.targets:
                for n in ast.walk(t):
                    if isinstance(n, ast.Name):
                        names.add(n.id)
        elif isinstance(node, ast.AnnAssign) and isinstance(node.target, ast.Name):
            names.add(node.target.id)
        elif isinstance(node, ast.Import):
            for a in node.names:
                names.add((a.asname or a.name).split(".")[0])
        elif isinstance(node, ast.ImportFrom):
            for a in node.names:
                if a.name == "*":
                    pkg = os.path.dirname(path)
                    sub = os.path.join(pkg, *(node.module or "").split("."))
                    for cand in (sub + ".py", os.path.join(sub, "__init__.py")):
                        if os.path.exists(cand):
                            names |= _top_level_names(_parse(cand), cand)
                else:
                    names.add(a.asname or a.name)
        elif isinstance(node, (ast.If, ast.Try)):
            for sub in ast.walk(node):
                if isinstance(sub, ast.ImportFrom):
                    names |= {a.asname or a.name for a in sub.names}
    # submodules of a package are importable names too
    if os.path.basename(path) == "__init__.py":
        for f in os.listdir(os.path.dirname(path)):
            names.add(f[:-3] if f.endswith(".py") else f)
    return names


def test_every_pennylane_import_exists_in_the_reference():
    tree = _parse(PLUGIN)
    checked = 0
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("pennylane") \
                and node.level == 0:
            path = _module_file(node.module)
            assert os.path.exists(path), f"{node.module}: no such module in the reference"
            have = _top_level_names(_parse(path), path)
            for a in node.names:
                assert a.name in have, f"{node.module} does not define {a.name}"
                checked += 1
    assert checked >= 25


def _class(tree, name):
    return next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)


def _methods(cls):
    out = {}
    for n in cls.body:
        if isinstance(n, ast.FunctionDef):
            out.setdefault(n.name, n)          # first definition (overloads come later)
    return out


def _params(fn):
    a = fn.args
    return [x.arg for x in a.posonlyargs + a.args] + ([f"*{a.vararg.arg}"] if a.vararg else []) + \
        [x.arg for x in a.kwonlyargs]


def test_overridden_device_methods_have_the_reference_signature():
    ref = _methods(_class(_parse(os.path.join(REF, "devices", "device_api.py")), "Device"))
    dq = _methods(_class(_parse(os.path.join(REF, "devices", "default_qubit.py")), "DefaultQubit"))
    ours = _methods(_class(_parse(PLUGIN), "B200QubitDevice"))
    api = ["preprocess_transforms", "setup_execution_config", "execute", "supports_derivatives",
           "compute_derivatives", "execute_and_compute_derivatives", "compute_jvp",
           "execute_and_compute_jvp", "compute_vjp", "execute_and_compute_vjp"]
    for name in api:
        assert name in ours, f"plugin does not implement {name}"
        src = dq.get(name) or ref[name]
        # the last overload / implementation in the reference class
        impl = [n for n in _class(_parse(os.path.join(REF, "devices", "device_api.py")), "Device").body
                if isinstance(n, ast.FunctionDef) and n.name == name][-1]
        assert _params(ours[name]) == _params(impl), (name, _params(ours[name]), _params(impl))
        assert _params(ours[name]) == _params(src), (name, _params(ours[name]), _params(src))
    # properties / attributes PennyLane reads on a device
    body_src = open(PLUGIN).read()
    for attr in ("_debugger", "_state_cache", "_rng", "name"):
        assert attr in body_src


def _add_transform_calls(fn):
    """[(transform name, sorted keyword names, guarding condition source)] in source order; a call
    to ``_add_adjoint_transforms`` is reported as such."""
    out = []

    def visit(nodes, cond):
        for node in nodes:
            if isinstance(node, ast.If):
                c = ast.unparse(node.test)
                visit(node.body, cond + [c])
                visit(node.orelse, cond + [f"not ({c})"])
                continue
            for sub in ast.walk(node):
                if isinstance(sub, ast.Call):
                    f = sub.func
                    if isinstance(f, ast.Attribute) and f.attr == "add_transform":
                        name = ast.unparse(sub.args[0])
                        out.append((name, sorted(k.arg for k in sub.keywords), tuple(cond)))
                    elif isinstance(f, ast.Name) and f.id == "_add_adjoint_transforms":
                        out.append(("_add_adjoint_transforms", sorted(k.arg for k in sub.keywords), tuple(cond)))
    visit(fn.body, [])
    return out


def test_preprocess_transforms_adds_the_reference_pipeline():
    dq = _methods(_class(_parse(os.path.join(REF, "devices", "default_qubit.py")), "DefaultQubit"))
    ours = _methods(_class(_parse(PLUGIN), "B200QubitDevice"))
    ref_calls = [c for c in _add_transform_calls(dq["preprocess_transforms"])
                 if c[0] != "validate_multiprocessing_workers"]
    our_calls = _add_transform_calls(ours["preprocess_transforms"])
    assert [c[0] for c in our_calls] == [c[0] for c in ref_calls]
    for (n1, k1, c1), (n2, k2, c2) in zip(our_calls, ref_calls):
        assert k1 == k2, (n1, k1, k2)
        assert c1 == c2, (n1, c1, c2)
    # the adjoint part is the reference's own helper, which must still be what INTEGRATION.md says
    tree = _parse(os.path.join(REF, "devices", "default_qubit.py"))
    helper = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_add_adjoint_transforms")
    assert [c[0] for c in _add_transform_calls(helper)] == [
        "no_sampling", "decompose", "validate_observables", "validate_measurements",
        "adjoint_state_measurements", "broadcast_expand", "validate_adjoint_trainable_params"]
    assert _params(helper) == ["program", "device_vjp", "device_wires", "target_gates"]


def test_execution_config_resolution_matches_the_reference():
    dq = _methods(_class(_parse(os.path.join(REF, "devices", "default_qubit.py")), "DefaultQubit"))
    ours = _methods(_class(_parse(PLUGIN), "B200QubitDevice"))

    def keys(fn):
        out = []
        for n in ast.walk(fn):
            if isinstance(n, ast.Subscript) and isinstance(n.value, ast.Name) and n.value.id == "updated_values" \
                    and isinstance(n.slice, ast.Constant):
                out.append(n.slice.value)
        return sorted(set(out))

    ref_keys = [k for k in keys(dq["setup_execution_config"]) if k != "convert_to_numpy"]   # JAX only
    assert keys(ours["setup_execution_config"]) == ref_keys
    # _setup_mcm_config: same supported methods, same messages
    ref_src = ast.unparse(dq["_setup_mcm_config"]).replace("default.qubit", "b200.qubit")
    our_src = ast.unparse(ours["_setup_mcm_config"])
    strip = lambda s: "".join(s.split())       # noqa: E731
    import re
    drop_doc = lambda s: re.sub(r'"""[^"]*"""', "", s)   # noqa: E731
    assert strip(drop_doc(our_src)).replace("mcm_config,tape", "X") == \
        strip(drop_doc(re.sub(r":\s*MCMConfig|:\s*QuantumScript|->\s*MCMConfig", "", ref_src))).replace("mcm_config,tape", "X")
