#!/usr/bin/env python
"""Generate tests/golden/reference_known_answers.json.

The reference (PennyLane v0.46.0-dev81, /root/reference) cannot be imported in the build
container (autograd, autoray, rustworkx, ... are absent and must not be stubbed), so fixtures
cannot be produced by running it.  Instead this script TRANSCRIBES the known-answer vectors that
the reference's own test-suite pins for the hot path: literal inputs are copied from the cited
test, expected outputs are evaluated from the closed-form expression the cited test asserts
(with numpy only — neither the oracle nor pennylane_b200 is imported here).

Every case records `source` = reference file:lines.  Gate specs are
{"name", "wires", "params", "hyper"}; measurement specs {"kind", "obs" | "wires"}; observable
specs are gate specs or {"name": "Hamiltonian"/"Sum"/"Prod", ...}.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def c(x):
    x = np.asarray(x, dtype=complex)
    return {"re": x.real.tolist(), "im": x.imag.tolist()}


def g(name, wires, params=(), **hyper):
    return {"name": name, "wires": list(wires), "params": [float(p) for p in params], "hyper": hyper}


cases = []

# ---- tests/devices/qubit/test_apply_operation.py:68-90 -------------------------------------------
mat = np.array([[0.39918205 + 0.3024376j, -0.86421077 + 0.04821758j],
                [0.73240679 + 0.46126509j, 0.49576832 - 0.07091251j]])
st = np.array([-0.30688912 - 0.4768824j, 0.8100052 - 0.14931113j])
cases.append({"id": "custom_operator_with_matrix", "type": "apply",
              "source": "tests/devices/qubit/test_apply_operation.py:68-90",
              "state": c(st), "op": {"name": "QubitUnitary", "wires": [0], "matrix": c(mat)},
              "expected": c(mat @ st), "atol": 1e-8})

# ---- tests/devices/qubit/test_apply_operation.py:255-445 (fixed two-qubit state) ------------------
s2 = np.array([[0.04624539 + 0.3895457j, 0.22399401 + 0.53870339j],
               [-0.483054 + 0.2468498j, -0.02772249 - 0.45901669j]])
for wire in (0, 1):
    def take(a, i):
        return np.take(a, i, axis=wire)

    def build(new0, new1):
        return np.stack([new0, new1], axis=wire)

    i0, i1 = take(s2, 0), take(s2, 1)
    r2 = 1 / np.sqrt(2)
    shift = np.exp(1j * -2.3)
    table = [
        ("paulix", g("PauliX", [wire]), build(i1, i0), ":255-276"),
        ("pauliz", g("PauliZ", [wire]), build(i0, -i1), ":278-298"),
        ("pauliy", g("PauliY", [wire]), build(-1j * i1, 1j * i0), ":300-320"),
        ("hadamard", g("Hadamard", [wire]), build(r2 * (i0 + i1), r2 * (i0 - i1)), ":322-345"),
        ("phaseshift", g("PhaseShift", [wire], [-2.3]), build(i0, shift * i1), ":347-368"),
        ("identity", g("Identity", [wire]), s2, ":419-431"),
        ("globalphase", g("GlobalPhase", [wire], [-2.3]), np.exp(2.3j) * s2, ":433-450"),
    ]
    for name, op, exp, lines in table:
        cases.append({"id": f"{name}_wire{wire}", "type": "apply",
                      "source": f"tests/devices/qubit/test_apply_operation.py{lines}",
                      "state": c(s2), "op": op, "expected": c(exp), "atol": 1e-8})
    # CNOT (:370-392): control = wire, target = 1 - wire; control-1 block has its target swapped
    control, target = wire, 1 - wire
    c0, c1 = np.take(s2, 0, axis=control), np.take(s2, 1, axis=control)
    exp = np.stack([c0, c1[::-1]], axis=control)
    cases.append({"id": f"cnot_control{wire}", "type": "apply",
                  "source": "tests/devices/qubit/test_apply_operation.py:370-392",
                  "state": c(s2), "op": g("CNOT", [control, target]), "expected": c(exp),
                  "atol": 1e-8})

# ---- tests/devices/qubit/test_measure.py:117-129 ------------------------------------------------------
sm = -0.5j * np.ones((2, 2))
cases.append({"id": "measure_state_no_obs", "type": "measure", "source":
              "tests/devices/qubit/test_measure.py:117-129", "state": c(sm),
              "measurement": {"kind": "state"}, "expected": c(-0.5j * np.ones(4)), "atol": 1e-8})
cases.append({"id": "measure_probs_wire0", "type": "measure", "source":
              "tests/devices/qubit/test_measure.py:117-129", "state": c(sm),
              "measurement": {"kind": "probs", "wires": [0]}, "expected": c([0.5, 0.5]),
              "atol": 1e-8})

# ---- tests/devices/qubit/test_measure.py:131-158 (RX(0.123)|0>) ----------------------------------------
srx = np.array([np.cos(0.123 / 2), -1j * np.sin(0.123 / 2)])
expH = 0.5 * np.sin(0.123) + 2 * np.cos(0.123)
cases.append({"id": "hamiltonian_expval", "type": "measure",
              "source": "tests/devices/qubit/test_measure.py:131-158", "state": c(srx),
              "measurement": {"kind": "expval", "obs": {
                  "name": "Hamiltonian", "coeffs": [-0.5, 2.0],
                  "ops": [g("PauliY", [0]), g("PauliZ", [0])]}},
              "expected": c(expH), "atol": 1e-8})
Hm = -0.5 * np.array([[0, -1j], [1j, 0]]) + 2 * np.diag([1.0, -1.0])
cases.append({"id": "hermitian_expval", "type": "measure",
              "source": "tests/devices/qubit/test_measure.py:131-158", "state": c(srx),
              "measurement": {"kind": "expval", "obs": {"name": "Hermitian", "wires": [0],
                                                        "matrix": c(Hm)}},
              "expected": c(expH), "atol": 1e-8})

# ---- tests/devices/qubit/test_measure.py:160-171 (8-wire Sum, tensor contraction) ----------------------
cases.append({"id": "sum_expval_tensor_contraction", "type": "simulate",
              "source": "tests/devices/qubit/test_measure.py:160-171",
              "ops": [g("RX", [i], [0.123]) for i in range(8)],
              "measurements": [{"kind": "expval", "obs": {"name": "Sum", "ops": [
                  {"name": "Prod", "ops": [g("PauliY", [i]), g("PauliZ", [i + 1])]}
                  for i in range(7)]}}],
              "expected": [c(7 * (-np.sin(0.123) * np.cos(0.123)))], "atol": 1e-8})
# ---- tests/devices/qubit/test_measure.py:173-190 --------------------------------------------------------
ops8 = [g("RX", [i], [i * np.pi / 2 + 0.123]) for i in range(8)]
cases.append({"id": "sum_expval_eigs_yz", "type": "simulate",
              "source": "tests/devices/qubit/test_measure.py:173-190", "ops": ops8,
              "measurements": [{"kind": "expval", "obs": {"name": "Sum", "ops": [
                  g("PauliY", [0]), g("PauliZ", [0])]}}],
              "expected": [c(-np.sin(0.123) + np.cos(0.123))], "atol": 1e-8})
cases.append({"id": "sum_expval_eigs_z8", "type": "simulate",
              "source": "tests/devices/qubit/test_measure.py:173-190", "ops": ops8,
              "measurements": [{"kind": "expval", "obs": {"name": "Sum", "ops": [
                  g("PauliZ", [i]) for i in range(8)]}}],
              "expected": [c(sum(np.sin(i * np.pi / 2 + 0.123) for i in range(8)))], "atol": 1e-8})

# ---- tests/devices/qubit/test_simulate.py:146-170 ---------------------------------------------------------
phi = 0.397
cases.append({"id": "simulate_basic_circuit", "type": "simulate",
              "source": "tests/devices/qubit/test_simulate.py:146-170",
              "ops": [g("RX", [0], [phi])],
              "measurements": [{"kind": "expval", "obs": g("PauliY", [0])},
                               {"kind": "expval", "obs": g("PauliZ", [0])},
                               {"kind": "state"}],
              "expected": [c(-np.sin(phi)), c(np.cos(phi)),
                           c([np.cos(phi / 2), -1j * np.sin(phi / 2)])], "atol": 1e-8})

# ---- tests/devices/qubit/test_sampling.py:137-142 (bit-exact, seed-pinned) ----------------------------------
cases.append({"id": "sample_state_seed_12345", "type": "sample",
              "source": "tests/devices/qubit/test_sampling.py:137-142",
              "state": c(np.array([[0, 1j], [-1, 0]]) / np.sqrt(2)), "shots": 4, "seed": 12345,
              "expected_samples": [[0, 1], [0, 1], [1, 0], [1, 0]]})

# ---- tests/devices/qubit/test_adjoint_jacobian.py (closed forms) --------------------------------------------
x = 0.654
cases.append({"id": "adjoint_jvp_single_param_multi_obs", "type": "jvp",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:369-383",
              "ops": [g("RY", [0], [x])], "trainable": [0], "tangents": [1.232],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [0])},
                               {"kind": "expval", "obs": g("PauliX", [0])}],
              "expected": (1.232 * np.array([-np.sin(x), np.cos(x)])).tolist(), "atol": 1e-8})
cases.append({"id": "adjoint_vjp_single_param_multi_obs", "type": "vjp",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:486-500",
              "ops": [g("RY", [0], [x])], "trainable": [0], "cotangents": [1.232, 2.963],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [0])},
                               {"kind": "expval", "obs": g("PauliX", [0])}],
              "expected": [float(np.dot([-np.sin(x), np.cos(x)], [1.232, 2.963]))], "atol": 1e-8})
cases.append({"id": "adjoint_jacobian_ry_z", "type": "jacobian",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:360-367, :475-484",
              "ops": [g("RY", [0], [x])], "trainable": [0],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [0])}],
              "expected": [[float(-np.sin(x))]], "atol": 1e-8})

# ---- tests/devices/qubit/test_simulate.py:172-261 (the same circuit at the angles of the four
#      interface tests; gradients of those tests are pinned by the adjoint cases below) --------------
for tag, phi_i, lines in (("autograd", -0.52, ":172-189"), ("jax", 0.678, ":192-214"),
                          ("torch", -0.526, ":216-237"), ("tf", 4.873, ":240-260")):
    cases.append({"id": f"simulate_rx_expval_yz_{tag}_angle", "type": "simulate",
                  "source": f"tests/devices/qubit/test_simulate.py{lines}",
                  "ops": [g("RX", [0], [phi_i])],
                  "measurements": [{"kind": "expval", "obs": g("PauliY", [0])},
                                   {"kind": "expval", "obs": g("PauliZ", [0])}],
                  "expected": [c(-np.sin(phi_i)), c(np.cos(phi_i))], "atol": 1e-8})
# ---- tests/devices/qubit/test_simulate.py:263-268 ---------------------------------------------------------
cases.append({"id": "simulate_rx_pi_expval_z", "type": "simulate",
              "source": "tests/devices/qubit/test_simulate.py:263-268",
              "ops": [g("RX", [0], [np.pi])],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [0])}],
              "expected": [c(-1.0)], "atol": 1e-8})
# ---- tests/devices/qubit/test_simulate.py:1003-1076 (quantum-information measurements of IsingXX(phi)|00>)
phi_q = -0.623
d_i = np.array([[np.cos(phi_q / 2) ** 2, 0], [0, np.sin(phi_q / 2) ** 2]])
d_both = np.array([[np.cos(phi_q / 2) ** 2, 0, 0, 0.0 + np.sin(phi_q) * 0.5j], [0, 0, 0, 0], [0, 0, 0, 0],
                   [0.0 - np.sin(phi_q) * 0.5j, 0, 0, np.sin(phi_q / 2) ** 2]])
root = np.sqrt(1 - 4 * np.cos(phi_q / 2) ** 2 * np.sin(phi_q / 2) ** 2)
eigs = [e for e in ((1 + root) / 2, (1 - root) / 2) if e > 0]
entropy = -np.sum(np.array(eigs) * np.log(eigs))
cases.append({"id": "simulate_qinfo_isingxx", "type": "simulate",
              "source": "tests/devices/qubit/test_simulate.py:1003-1076",
              "ops": [g("IsingXX", [0, 1], [phi_q])],
              "measurements": [{"kind": "density_matrix", "wires": [0]}, {"kind": "density_matrix", "wires": [1]},
                               {"kind": "density_matrix", "wires": [0, 1]}, {"kind": "vn_entropy", "wires": [0]},
                               {"kind": "vn_entropy", "wires": [1]},
                               {"kind": "mutual_info", "wires0": [0], "wires1": [1]}],
              "expected": [c(d_i), c(d_i), c(d_both), c(entropy), c(entropy), c(2 * entropy)], "atol": 1e-8})

# ---- tests/devices/qubit/test_adjoint_jacobian.py:133-149 (three RX, diagonal Jacobian) -----------------------
par3 = [np.pi, np.pi / 2, np.pi / 3]
cases.append({"id": "adjoint_jacobian_multiple_rx", "type": "jacobian",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:133-149",
              "ops": [g("RX", [i], [par3[i]]) for i in range(3)], "trainable": [0, 1, 2],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [i])} for i in range(3)],
              "expected": (-np.diag(np.sin(par3))).tolist(), "atol": 1e-8})
# ---- :230-268 (Hermitian observable on wires 0, 2) and :270-298 (X0 @ Y2): the same closed form ----------------
a_, b_, c_ = 0.5, 0.3, -0.7
ops3 = [g("RX", [0], [a_]), g("RX", [1], [b_]), g("RX", [2], [c_]), g("CNOT", [0, 1]), g("CNOT", [1, 2])]
exp3 = [np.cos(a_) * np.sin(b_) * np.sin(c_), np.cos(b_) * np.sin(a_) * np.sin(c_),
        np.cos(c_) * np.sin(b_) * np.sin(a_)]
mx = np.kron(np.array([[0, 1], [1, 0]]), np.array([[0, -1j], [1j, 0]]))
cases.append({"id": "adjoint_jacobian_hermitian_x0y2", "type": "jacobian",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:230-268",
              "ops": ops3, "trainable": [0, 1, 2],
              "measurements": [{"kind": "expval", "obs": {"name": "Hermitian", "wires": [0, 2], "matrix": c(mx)}}],
              "expected": [exp3], "atol": 1e-8})
cases.append({"id": "adjoint_jacobian_tensor_x0y2", "type": "jacobian",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:270-298",
              "ops": ops3, "trainable": [0, 1, 2],
              "measurements": [{"kind": "expval", "obs": {"name": "Prod", "ops": [g("PauliX", [0]), g("PauliY", [2])]}}],
              "expected": [exp3], "atol": 1e-8})
# ---- :392-432 (JVP, two parameters) and :503-545 (VJP, two parameters), RY(x) RZ(y) on one wire --------------
y = 1.221
jac3 = np.array([[-np.sin(x), 0], [np.cos(x) * np.cos(y), -np.sin(x) * np.sin(y)],
                 [np.cos(x) * np.sin(y), np.sin(x) * np.cos(y)]])
mps3 = [{"kind": "expval", "obs": g(nm, [0])} for nm in ("PauliZ", "PauliX", "PauliY")]
for tg in ((0.0, 0.653), (1.232, 2.963)):
    cases.append({"id": f"adjoint_jvp_multi_param_single_obs_{tg[1]}", "type": "jvp",
                  "source": "tests/devices/qubit/test_adjoint_jacobian.py:392-407",
                  "ops": [g("RY", [0], [x]), g("RZ", [0], [y])], "trainable": [0, 1], "tangents": list(tg),
                  "measurements": [mps3[2]], "expected": [float(jac3[2] @ np.array(tg))], "atol": 1e-8})
    cases.append({"id": f"adjoint_jvp_multi_param_multi_obs_{tg[1]}", "type": "jvp",
                  "source": "tests/devices/qubit/test_adjoint_jacobian.py:409-432",
                  "ops": [g("RY", [0], [x]), g("RZ", [0], [y])], "trainable": [0, 1], "tangents": list(tg),
                  "measurements": mps3, "expected": (jac3 @ np.array(tg)).tolist(), "atol": 1e-8})
# :434-455 uses wire labels [1, 0] / ["a", "b"]; transcribed on the standard labels it maps them to
cases.append({"id": "adjoint_jvp_two_wires_ry_rx", "type": "jvp",
              "source": "tests/devices/qubit/test_adjoint_jacobian.py:434-455",
              "ops": [g("RY", [0], [x]), g("RX", [1], [y])], "trainable": [0, 1], "tangents": [1.232, 2.963],
              "measurements": [{"kind": "expval", "obs": g("PauliZ", [0])}, {"kind": "expval", "obs": g("PauliY", [1])},
                               {"kind": "expval", "obs": g("PauliX", [0])}],
              "expected": (np.array([[-np.sin(x), 0], [0, -np.cos(y)], [np.cos(x), 0]]) @ np.array([1.232, 2.963])).tolist(),
              "atol": 1e-8})
for ct in ((0.0, 0.653, 0.0), (1.236, 0.0, 0.573), (1.232, 2.963, 1.942)):
    cases.append({"id": f"adjoint_vjp_multi_param_multi_obs_{ct[0]}", "type": "vjp",
                  "source": "tests/devices/qubit/test_adjoint_jacobian.py:519-545",
                  "ops": [g("RY", [0], [x]), g("RZ", [0], [y])], "trainable": [0, 1], "cotangents": list(ct),
                  "measurements": mps3, "expected": (np.array(ct) @ jac3).tolist(), "atol": 1e-8})

out = os.path.join(HERE, "reference_known_answers.json")
json.dump({"reference": "PennyLaneAI/pennylane v0.46.0-dev81 (tests/devices/qubit/*)",
           "generated_by": "tests/golden/make_golden.py", "cases": cases}, open(out, "w"), indent=1)
print(f"wrote {len(cases)} cases to {out}")
