"""GPU: every known-answer vector the reference's own tests hold for the hot path
(tests/golden/reference_known_answers.json, provenance per case in ``source``) through the CUDA
engine — the per-gate kernels (fusion 0), the fused path (fusion 1) and the adjoint sweeps — at the
tolerance the reference test uses (1e-8 absolute, written in the fixture)."""
import numpy as np
import pytest

import pennylane_b200 as qb
from golden_utils import build_mp, build_op, build_tape, cx, load_cases
from pennylane_b200 import adjoint
from pennylane_b200.simulate import measure
from pennylane_b200.statevector import StateVector

pytestmark = pytest.mark.gpu
CASES = load_cases()


def _of(kind):
    return [c for c in CASES if c["type"] == kind]


def _sv(state):
    state = np.asarray(state, dtype=np.complex128)
    n = int(np.log2(state.size))
    sv = StateVector(n)
    sv.set_state(state.reshape(-1))
    return sv


@pytest.mark.parametrize("case", _of("apply"), ids=lambda c: c["id"])
def test_apply(case):
    state = cx(case["state"])
    sv = _sv(state)
    sv.apply_operation(build_op(case["op"]))
    got = sv.to_numpy().reshape(state.shape)
    assert np.allclose(got, cx(case["expected"]), atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", _of("measure"), ids=lambda c: c["id"])
def test_measure(case):
    got = np.asarray(measure(build_mp(case["measurement"]), _sv(cx(case["state"]))))
    exp = cx(case["expected"])
    assert np.allclose(got, exp.real if np.isrealobj(got) else exp, atol=case["atol"], rtol=0)


@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("case", _of("simulate"), ids=lambda c: c["id"])
def test_simulate(case, fusion):
    tape = build_tape(case)
    res = qb.B200Qubit(wires=tape.num_wires, fusion=fusion).execute(tape)
    res = res if isinstance(res, tuple) else (res,)
    assert len(res) == len(case["expected"])
    for r, e in zip(res, case["expected"]):
        e = cx(e)
        r = np.asarray(r)
        assert np.allclose(r.reshape(e.shape), e.real if np.isrealobj(r) else e, atol=case["atol"], rtol=0)


@pytest.mark.parametrize("case", _of("sample"), ids=lambda c: c["id"])
def test_sample_bit_exact(case):
    got = _sv(cx(case["state"])).sample(case["shots"], np.random.default_rng(case["seed"]))
    assert np.asarray(got).tolist() == case["expected_samples"]


@pytest.mark.parametrize("fusion", [0, 1])
@pytest.mark.parametrize("case", _of("jacobian") + _of("jvp") + _of("vjp"), ids=lambda c: c["id"])
def test_adjoint(case, fusion):
    tape = build_tape(case)
    if case["type"] == "jacobian":
        got = np.atleast_2d(np.array(adjoint.adjoint_jacobian(tape, fusion=fusion), dtype=float))
    elif case["type"] == "jvp":
        got = np.array(adjoint.adjoint_jvp(tape, case["tangents"], fusion=fusion), dtype=float)
    else:
        got = np.array(adjoint.adjoint_vjp(tape, case["cotangents"], fusion=fusion), dtype=float)
    assert np.allclose(got.reshape(np.shape(case["expected"])), np.array(case["expected"]),
                       atol=case["atol"], rtol=0), case["id"]
