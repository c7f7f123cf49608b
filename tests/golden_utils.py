"""Build operator / measurement objects from the golden-fixture specs (tests/golden)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def load_cases():
    return json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))["cases"]


def cx(d):
    return np.asarray(d["re"], dtype=float) + 1j * np.asarray(d["im"], dtype=float)


def build_op(spec):
    from pennylane_b200 import ops as q

    name = spec["name"]
    if name in ("QubitUnitary", "Hermitian"):
        return getattr(q, name)(cx(spec["matrix"]), wires=spec["wires"])
    if name in ("Hamiltonian", "LinearCombination"):
        return q.LinearCombination(spec["coeffs"], [build_op(o) for o in spec["ops"]])
    if name == "Sum":
        return q.Sum(*[build_op(o) for o in spec["ops"]])
    if name == "Prod":
        return q.Prod(*[build_op(o) for o in spec["ops"]])
    cls = getattr(q, name)
    return cls(*spec.get("params", []), wires=spec["wires"], **spec.get("hyper", {}))


def build_mp(spec):
    import pennylane_b200 as qb

    kind = spec["kind"]
    if kind == "state":
        return qb.state()
    if "obs" in spec:
        return getattr(qb, kind)(build_op(spec["obs"]))
    if kind == "mutual_info":
        return qb.mutual_info(spec["wires0"], spec["wires1"])
    if kind in ("density_matrix", "vn_entropy", "purity"):
        return getattr(qb, kind)(spec["wires"])
    return getattr(qb, kind)(wires=spec["wires"])


def build_tape(case, shots=None):
    import pennylane_b200 as qb

    return qb.QuantumScript([build_op(o) for o in case["ops"]],
                            [build_mp(m) for m in case["measurements"]], shots=shots,
                            trainable_params=case.get("trainable"))
