"""ORACLE (test infrastructure only — never imported by ``pennylane_b200``).

Gate matrices for the CPU restatement of default.qubit, written independently of
``pennylane_b200/ops.py``: rotation-like gates are built as ``expm(i * theta * G)`` from the
generators the reference documents, everything else from the literal matrices in the reference
files cited per entry.  Operators are consumed duck-typed (``name``, ``wires``, ``data``,
``hyperparameters``, optional ``base`` / ``control_wires`` / ``control_values``), so genuine
PennyLane operators and the repo's mirror classes both work.

Reference: pennylane/ops/qubit/{non_parametric_ops,parametric_ops_single_qubit,
parametric_ops_multi_qubit,qchem_ops,matrix_ops,observables}.py, pennylane/ops/op_math/
{controlled_ops,controlled,adjoint}.py, pennylane/ops/identity.py.
"""
import numpy as np
from scipy.linalg import expm

I2 = np.eye(2, dtype=complex)
PX = np.array([[0, 1], [1, 0]], dtype=complex)
PY = np.array([[0, -1j], [1j, 0]], dtype=complex)
PZ = np.array([[1, 0], [0, -1]], dtype=complex)
PAULI = {"I": I2, "X": PX, "Y": PY, "Z": PZ}
P1 = np.array([[0, 0], [0, 1]], dtype=complex)  # |1><1|
SQ2 = np.sqrt(2)


def kron(*ms):
    out = np.array([[1.0 + 0j]])
    for m in ms:
        out = np.kron(out, m)
    return out


def word(s):
    return kron(*[PAULI[c] for c in s])


FIXED = {
    # non_parametric_ops.py:98 / :366 / :579 / :807 / :1058 / :1198 / :1325 / :1458 / :1640 /
    # :1780 / :1936
    "Identity": I2,
    "Hadamard": np.array([[1, 1], [1, -1]], dtype=complex) / SQ2,
    "PauliX": PX, "PauliY": PY, "PauliZ": PZ,
    "S": np.diag([1, 1j]).astype(complex),
    "T": np.diag([1, np.exp(1j * np.pi / 4)]).astype(complex),
    "SX": 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]),
    "SWAP": np.eye(4, dtype=complex)[[0, 2, 1, 3]],
    "ECR": np.array([[0, 0, 1, 1j], [0, 0, 1j, 1], [1, -1j, 0, 0], [-1j, 1, 0, 0]]) / SQ2,
    "ISWAP": np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex),
    "SISWAP": np.array([[1, 0, 0, 0], [0, 1 / SQ2, 1j / SQ2, 0], [0, 1j / SQ2, 1 / SQ2, 0],
                        [0, 0, 0, 1]], dtype=complex),
}


def controlled(base, n_ctrl, control_values=None):
    """op_math/controlled.py: identity except on the block selected by the control values."""
    d = base.shape[-1]
    D = d << n_ctrl
    vals = [1] * n_ctrl if control_values is None else [int(bool(v)) for v in control_values]
    sel = 0
    for v in vals:
        sel = (sel << 1) | v
    out = np.eye(D, dtype=complex)
    out[sel * d:(sel + 1) * d, sel * d:(sel + 1) * d] = base
    return out


FIXED.update({
    "CNOT": controlled(PX, 1), "CZ": controlled(PZ, 1), "CY": controlled(PY, 1),
    "CH": controlled(FIXED["Hadamard"], 1), "CSWAP": controlled(FIXED["SWAP"], 1),
    "Toffoli": controlled(PX, 2), "CCZ": controlled(PZ, 2),
})

# one-parameter gates U(theta) = expm(i * theta * G): generator matrices from
# parametric_ops_single_qubit.py:111,322,540,826; parametric_ops_multi_qubit.py:133,451,1113,
# 1248,1392,1578; controlled_ops.py (CRX/CRY/CRZ/ControlledPhaseShift generators);
# qchem_ops.py:180,316,487,675.
GENERATORS = {
    "RX": lambda op: -0.5 * PX,
    "RY": lambda op: -0.5 * PY,
    "RZ": lambda op: -0.5 * PZ,
    "PhaseShift": lambda op: P1,
    "U1": lambda op: P1,
    "IsingXX": lambda op: -0.5 * word("XX"),
    "IsingYY": lambda op: -0.5 * word("YY"),
    "IsingZZ": lambda op: -0.5 * word("ZZ"),
    "IsingXY": lambda op: 0.25 * (word("XX") + word("YY")),
    "CRX": lambda op: -0.5 * kron(P1, PX),
    "CRY": lambda op: -0.5 * kron(P1, PY),
    "CRZ": lambda op: -0.5 * kron(P1, PZ),
    "ControlledPhaseShift": lambda op: kron(P1, P1),
    "MultiRZ": lambda op: -0.5 * word("Z" * len(op.wires)),
    "PauliRot": lambda op: -0.5 * word(op.hyperparameters["pauli_word"]),
    "SingleExcitation": lambda op: 0.25 * (word("XY") - word("YX")),
    "SingleExcitationMinus": lambda op: 0.25 * (-word("II") + word("XY") - word("YX") - word("ZZ")),
    "SingleExcitationPlus": lambda op: 0.25 * (word("II") + word("XY") - word("YX") + word("ZZ")),
    "DoubleExcitation": lambda op: 0.0625 * (
        word("XXXY") + word("XXYX") - word("XYXX") + word("XYYY") - word("YXXX") + word("YXYY")
        - word("YYXY") - word("YYYX")),
    "GlobalPhase": lambda op: -np.eye(2 ** max(1, len(op.wires)), dtype=complex),
}


def _scalar_matrix(op, theta):
    """Matrix of ``op`` for ONE (unbatched) parameter set."""
    name = op.name
    if name in FIXED:
        return FIXED[name]
    if name in GENERATORS:
        return expm(1j * float(theta[0]) * GENERATORS[name](op))
    if name == "Rot":  # parametric_ops_single_qubit.py:1000-1003: RZ(omega) RY(theta) RZ(phi)
        phi, th, om = (float(t) for t in theta)
        return expm(-0.5j * om * PZ) @ expm(-0.5j * th * PY) @ expm(-0.5j * phi * PZ)
    if name == "CRot":
        phi, th, om = (float(t) for t in theta)
        return controlled(expm(-0.5j * om * PZ) @ expm(-0.5j * th * PY) @ expm(-0.5j * phi * PZ), 1)
    if name == "U2":  # parametric_ops_single_qubit.py:1322-1326
        phi, de = (float(t) for t in theta)
        return np.array([[1, -np.exp(1j * de)], [np.exp(1j * phi), np.exp(1j * (phi + de))]]) / SQ2
    if name == "U3":  # parametric_ops_single_qubit.py:1460-1466
        th, phi, de = (float(t) for t in theta)
        c, s = np.cos(th / 2), np.sin(th / 2)
        return np.array([[c, -s * np.exp(1j * de)],
                         [s * np.exp(1j * phi), c * np.exp(1j * (phi + de))]])
    if name == "PSWAP":  # parametric_ops_multi_qubit.py:1791-1836
        e = np.exp(1j * float(theta[0]))
        return np.array([[1, 0, 0, 0], [0, 0, e, 0], [0, e, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if name == "MultiControlledX":  # controlled_ops.py:1284
        cv = op.hyperparameters.get("control_values", None)
        return controlled(PX, len(op.wires) - 1, cv)
    if name == "GroverOperator":  # templates/subroutines/grover.py: 2|s><s| - I
        dim = 2 ** len(op.wires)
        return np.full((dim, dim), 2.0 / dim, dtype=complex) - np.eye(dim)
    if name in ("QubitUnitary", "Hermitian"):
        return np.asarray(theta[0], dtype=complex)
    if name == "DiagonalQubitUnitary":
        return np.diag(np.asarray(theta[0], dtype=complex))
    if name == "Projector":
        st = np.asarray(theta[0])
        k = len(op.wires)
        if st.shape == (k,):
            idx = int("".join(str(int(b)) for b in st), 2)
            m = np.zeros((2**k, 2**k), dtype=complex)
            m[idx, idx] = 1
            return m
        return np.outer(st, np.conj(st)).astype(complex)
    raise NotImplementedError(f"oracle has no matrix for {name}")


def _ndim_params(op):
    nd = getattr(op, "ndim_params", None)
    if nd:
        return tuple(nd)
    return (0,) * len(op.data)


def matrix_of(op):
    """Matrix of an operator on ITS OWN wires (first wire most significant); a leading batch
    axis appears when a parameter is broadcast (apply_operation.py:191-197)."""
    name = op.name
    base = getattr(op, "base", None)
    if name.startswith("Adjoint(") and base is not None:  # op_math/adjoint.py:396-398
        m = matrix_of(base)
        return np.conj(np.swapaxes(m, -1, -2))
    if (name.startswith("C(") or name == "ControlledQubitUnitary") and base is not None:
        m = matrix_of(base)
        cv = getattr(op, "control_values", None)
        nc = len(op.control_wires)
        if m.ndim == 3:
            return np.stack([controlled(x, nc, cv) for x in m])
        return controlled(m, nc, cv)
    if name in ("Prod", "SProd", "Sum", "LinearCombination", "Hamiltonian"):
        return observable_matrix(op, list(op.wires))
    data = list(op.data)
    nds = _ndim_params(op)
    bs = None
    for p, nd in zip(data, nds):
        if np.ndim(p) == nd + 1:
            bs = np.shape(p)[0]
    if bs is None:
        return _scalar_matrix(op, data)
    mats = []
    for b in range(bs):
        theta = [np.asarray(p)[b] if np.ndim(p) == nd + 1 else p for p, nd in zip(data, nds)]
        mats.append(_scalar_matrix(op, theta))
    return np.stack(mats)


def expand(mat, wires, wire_order):
    """Embed a matrix on ``wires`` into ``wire_order`` (kron with identities + permutation)."""
    wires, wire_order = list(wires), list(wire_order)
    n = len(wire_order)
    extra = [w for w in wire_order if w not in wires]
    full = np.kron(mat, np.eye(2 ** len(extra), dtype=complex)) if extra else mat
    cur = wires + extra
    perm = [cur.index(w) for w in wire_order]
    t = full.reshape((2,) * (2 * n))
    t = t.transpose(perm + [n + p for p in perm])
    return t.reshape(2**n, 2**n)


def observable_matrix(obs, wire_order):
    """Dense matrix of an (arithmetic) observable on ``wire_order``."""
    name = obs.name
    if name == "Prod":
        m = np.eye(2 ** len(wire_order), dtype=complex)
        for o in obs.operands:
            m = m @ observable_matrix(o, wire_order)
        return m
    if name == "SProd":
        return obs.scalar * observable_matrix(obs.base, wire_order)
    if name in ("Sum", "LinearCombination", "Hamiltonian"):
        if hasattr(obs, "terms") and name != "Sum":
            cs, os_ = obs.terms()
            return sum(c * observable_matrix(o, wire_order) for c, o in zip(cs, os_))
        return sum(observable_matrix(o, wire_order) for o in obs.operands)
    if name == "Identity":
        return np.eye(2 ** len(wire_order), dtype=complex)
    return expand(matrix_of(obs), obs.wires, wire_order)


def generator_matrix(op):
    """Generator of a one-parameter gate on its own wires (pennylane/operation.py:59)."""
    name = op.name
    base = getattr(op, "base", None)
    if name.startswith("Adjoint(") and base is not None:
        return -generator_matrix(base)
    if name.startswith("C(") and base is not None:
        nc = len(op.control_wires)
        cv = [1] * nc if getattr(op, "control_values", None) is None else op.control_values
        proj = kron(*[np.diag([1 - int(bool(v)), int(bool(v))]).astype(complex) for v in cv])
        return np.kron(proj, generator_matrix(base))
    return GENERATORS[name](op)
