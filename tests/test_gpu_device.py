"""GPU: the device boundary (execute / derivatives / preprocessing / tracker / seeds) and the
shrunk twins of the BASELINE.json configurations, against the oracle.  Mirrors
tests/devices/default_qubit/test_default_qubit.py (result nesting :939-1010, TestRandomSeed
:1206-1290) and tests/devices/test_lightning_qubit.py:26-45 (SEL value + gradient parity)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle(tape, rng=None):
    from oracle import simulate as o_sim

    return o_sim.simulate(tape.map_to_standard_wires(), rng=rng)


def test_rx_known_answers():
    """tests/devices/qubit/test_simulate.py:146-170: RX(0.397): <Y> = -sin, <Z> = cos,
    state = [cos(phi/2), -i sin(phi/2)]."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    phi = 0.397
    dev = qb.B200Qubit()
    tape = qb.QuantumScript([q.RX(phi, wires=0)],
                            [qb.expval(q.PauliY(wires=0)), qb.expval(q.PauliZ(wires=0)), qb.state()])
    y, z, st = dev.execute(tape)
    assert abs(y + np.sin(phi)) < 1e-14 and abs(z - np.cos(phi)) < 1e-14
    assert np.allclose(st, [np.cos(phi / 2), -1j * np.sin(phi / 2)], atol=1e-15)
    assert isinstance(y, np.float64)


def test_batch_of_tapes_and_custom_wire_labels():
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    dev = qb.B200Qubit()
    t1 = qb.QuantumScript([q.Hadamard(wires="a"), q.CNOT(wires=["a", "b"])],
                          [qb.probs(wires=["b", "a"]), qb.expval(q.PauliZ(wires="c"))])
    t2 = qb.QuantumScript([q.RY(0.3, wires=2)], [qb.var(q.PauliX(wires=2))])
    res = dev.execute((t1, t2))
    assert isinstance(res, tuple) and len(res) == 2
    assert np.allclose(res[0][0], [0.5, 0, 0, 0.5]) and abs(res[0][1] - 1.0) < 1e-15
    ref = _oracle(t2)
    assert abs(res[1] - ref) < 1e-13


def test_broadcast_execution_matches_oracle():
    """Parameter broadcasting through the device (simulate.py:211,235; config 4's mechanism)."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    n, B = 6, 4
    rng = np.random.default_rng(4)
    ops_ = []
    for i in range(n):
        ops_ += [q.RY(rng.uniform(0, 6, B), wires=i), q.RZ(rng.uniform(0, 6, B), wires=i)]
    ops_ += [q.CNOT(wires=[i, (i + 1) % n]) for i in range(n)]
    ops_ += [q.RX(rng.uniform(0, 6, B), wires=i) for i in range(n)]
    heis = q.LinearCombination(
        [1.0] * (3 * (n - 1)),
        [P(wires=i) @ P(wires=i + 1) for i in range(n - 1) for P in (q.PauliX, q.PauliY, q.PauliZ)])
    tape = qb.QuantumScript(ops_, [qb.expval(heis), qb.probs(wires=[0, 3])])
    e, p = qb.B200Qubit().execute(tape)
    re, rp = _oracle(tape)
    assert e.shape == (B,) and p.shape == (B, 4)
    assert np.max(np.abs(e - re)) < 1e-12 and np.max(np.abs(p - rp)) < 1e-13


def _sel_tape(n, layers, seed):
    """StronglyEntanglingLayers after the adjoint pipeline's decomposition
    (templates/layers/strongly_entangling.py:224-244; Rot -> RZ RY RZ)."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    w = np.random.default_rng(seed).uniform(0, 2 * np.pi, (layers, n, 3))
    ops_ = []
    for l in range(layers):
        for i in range(n):
            ops_.append(q.Rot(w[l, i, 0], w[l, i, 1], w[l, i, 2], wires=i))
        r = (l % (n - 1)) + 1
        for i in range(n):
            ops_.append(q.CNOT(wires=[i, (i + r) % n]))
    return qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0))])


def test_config1_twin_sel_adjoint():
    """BASELINE config 1 shrunk (12 qubits, 4 layers): value and adjoint gradient."""
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim

    import pennylane_b200 as qb

    n = 12
    tape = _sel_tape(n, 4, 1)
    dev = qb.B200Qubit(wires=n)
    (ptape,), config = dev.preprocess(tape, qb.ExecutionConfig(gradient_method="adjoint"))
    assert all(len(op.data) <= 1 for op in ptape.operations)        # Rot decomposed
    assert len(ptape.trainable_params) == 4 * n * 3
    res, jac = dev.execute_and_compute_derivatives(ptape, config)
    st, _ = o_sim.get_final_state(ptape)
    ref = o_sim.measure_final_state(ptape, st, False)
    ref_jac = np.array(o_adj.adjoint_jacobian(ptape, st), dtype=float)
    assert abs(res - ref) < 1e-12
    assert np.max(np.abs(np.array(jac, dtype=float) - ref_jac)) < 1e-12


def test_config2_twin_qaoa_maxcut():
    """BASELINE config 2 shrunk (14 qubits, 3-regular graph, p=4): PauliRot(ZZ)/PauliRot(X)
    layers, expval of the Z-type cost Hamiltonian (qaoa/cost.py:253-321)."""
    import networkx as nx

    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    n, p = 14, 4
    g = nx.random_regular_graph(3, n, seed=2)
    edges = list(g.edges)
    par = np.random.default_rng(2).uniform(0, 2 * np.pi, (2, p))
    ops_ = [q.Hadamard(wires=i) for i in range(n)]
    for l in range(p):
        for a, b in edges:
            ops_.append(q.PauliRot(par[0, l], "ZZ", wires=[a, b]))
        for i in range(n):
            ops_.append(q.PauliRot(2 * par[1, l], "X", wires=[i]))
    cost = q.LinearCombination(
        [0.5] * len(edges) + [-0.5] * len(edges),
        [q.PauliZ(wires=a) @ q.PauliZ(wires=b) for a, b in edges] + [q.Identity(wires=a) for a, _ in edges])
    tape = qb.QuantumScript(ops_, [qb.expval(cost)])
    for dtype, tol in ((np.complex128, 1e-12), (np.complex64, 1e-5)):
        got = qb.B200Qubit(wires=n, c_dtype=dtype).execute(tape)
        assert abs(got - _oracle(tape)) < tol * max(1.0, abs(_oracle(tape)))


def test_config5_twin_random_circuit_sampling():
    """BASELINE config 5 shrunk (12 qubits, depth 6): random Rot + CNOT matching, 100k shots
    bit-identical to the oracle under the same seed."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    n, depth = 12, 6
    rng = np.random.default_rng(5)
    ops_ = []
    for _ in range(depth):
        for i in range(n):
            ops_.append(q.Rot(*rng.uniform(0, 2 * np.pi, 3), wires=i))
        perm = rng.permutation(n)
        for a, b in zip(perm[::2], perm[1::2]):
            ops_.append(q.CNOT(wires=[int(a), int(b)]))
    tape = qb.QuantumScript(ops_, [qb.sample(wires=range(n))], shots=100000)
    got = qb.B200Qubit(wires=n, seed=5).execute(tape)
    ref = _oracle(tape, rng=np.random.default_rng(5))
    assert got.shape == (100000, n) and np.array_equal(got, ref)


def test_seed_determinism():
    """test_default_qubit.py:1206-1290 (TestRandomSeed)."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    tape = qb.QuantumScript([q.Hadamard(wires=0), q.RY(0.7, wires=1)], [qb.sample(wires=[0, 1])], shots=200)
    a = qb.B200Qubit(seed=123).execute(tape)
    b = qb.B200Qubit(seed=123).execute(tape)
    c = qb.B200Qubit(seed=124).execute(tape)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    dev = qb.B200Qubit(seed=123)
    first, second = dev.execute(tape), dev.execute(tape)
    assert np.array_equal(first, a) and not np.array_equal(first, second)


def test_preprocess_errors_and_tracker():
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    dev = qb.B200Qubit(wires=2)
    with pytest.raises(qb.DeviceError, match="wires not found"):
        dev.preprocess(qb.QuantumScript([q.PauliX(wires=5)], [qb.expval(q.PauliZ(wires=5))]))
    with pytest.raises(qb.DeviceError, match="analytic"):
        dev.preprocess(qb.QuantumScript([q.PauliX(wires=0)], [qb.sample(wires=0)]))
    with pytest.raises(qb.DeviceError, match="device option"):
        dev.setup_execution_config(qb.ExecutionConfig(device_options={"bogus": 1}))
    with pytest.raises(qb.DeviceError, match="max_workers"):
        qb.B200Qubit(max_workers=2)
    cfg = dev.setup_execution_config(qb.ExecutionConfig(gradient_method="adjoint"))
    assert cfg.use_device_gradient and cfg.grad_on_execution and cfg.use_device_jacobian_product
    assert dev.supports_derivatives(qb.ExecutionConfig(gradient_method="adjoint"))
    assert not dev.supports_derivatives(qb.ExecutionConfig(gradient_method="backprop"))
    tape = qb.QuantumScript([q.RX(0.1, wires=0)], [qb.expval(q.PauliZ(wires=0))])
    with dev.tracker:
        dev.execute((tape, tape))
        dev.compute_derivatives(tape)
    assert dev.tracker.totals["batches"] == 1 and dev.tracker.totals["simulations"] == 2
    assert dev.tracker.totals["executions"] == 2 and dev.tracker.totals["derivative_batches"] == 1


def test_result_nesting_of_derivatives():
    """test_default_qubit.py:947-990: scalar / tuple / tuple-of-tuples by (n_obs, n_params)."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    dev = qb.B200Qubit()
    one = qb.QuantumScript([q.RX(0.3, wires=0)], [qb.expval(q.PauliZ(wires=0))])
    j = dev.compute_derivatives(one)
    assert isinstance(j, np.ndarray) and j.shape == ()
    two_p = qb.QuantumScript([q.RX(0.3, wires=0), q.RY(0.2, wires=0)], [qb.expval(q.PauliZ(wires=0))])
    j = dev.compute_derivatives(two_p)
    assert isinstance(j, tuple) and len(j) == 2
    two_o = qb.QuantumScript([q.RX(0.3, wires=0), q.RY(0.2, wires=0)],
                             [qb.expval(q.PauliZ(wires=0)), qb.expval(q.PauliX(wires=0))])
    j = dev.compute_derivatives(two_o)
    assert isinstance(j, tuple) and len(j) == 2 and len(j[0]) == 2
    batch = dev.compute_derivatives((one,))
    assert isinstance(batch, tuple) and len(batch) == 1
    res, jac = dev.execute_and_compute_derivatives((one, two_p))
    assert len(res) == 2 and len(jac) == 2
    assert abs(res[0] - np.cos(0.3)) < 1e-14 and abs(float(jac[0]) + np.sin(0.3)) < 1e-14


def test_return_torch_keeps_state_probs_and_jacobian_on_the_device(monkeypatch):
    """Device option ``return_torch``: ``state`` / ``probs`` results are
    torch CUDA tensors and no device-to-host copy happens while producing them; the adjoint
    Jacobian comes back as one CUDA tensor."""
    import torch

    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q

    n = 12
    rng = np.random.default_rng(4)
    ops_ = []
    for w in range(n):
        ops_ += [q.RY(rng.uniform(0, 6), wires=w), q.RZ(rng.uniform(0, 6), wires=w)]
    ops_ += [q.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    tape = qb.QuantumScript(ops_, [qb.state(), qb.probs(wires=[0, 3, 5])])
    dev = qb.B200Qubit(wires=n, return_torch=True)
    batch, cfg = dev.preprocess(tape)
    dev.execute(batch, cfg)                                   # warm-up: allocations, module loads
    calls = {"n": 0}
    real_cpu, real_item = torch.Tensor.cpu, torch.Tensor.item

    def counting_cpu(self, *a, **k):
        if self.is_cuda:
            calls["n"] += 1
        return real_cpu(self, *a, **k)

    def counting_item(self, *a, **k):
        if self.is_cuda:
            calls["n"] += 1
        return real_item(self, *a, **k)

    monkeypatch.setattr(torch.Tensor, "cpu", counting_cpu)
    monkeypatch.setattr(torch.Tensor, "item", counting_item)
    state, probs = dev.execute(batch, cfg)[0]
    monkeypatch.undo()
    assert calls["n"] == 0, "a device-to-host copy happened with return_torch"
    assert isinstance(state, torch.Tensor) and state.is_cuda and state.dtype == torch.complex128
    assert isinstance(probs, torch.Tensor) and probs.is_cuda and probs.shape == (8,)
    ref, _ = o_sim.get_final_state(tape)
    assert np.max(np.abs(state.cpu().numpy() - np.asarray(ref).reshape(-1))) < 1e-12
    ref_p = o_sim.measure_final_state(qb.QuantumScript(ops_, [qb.probs(wires=[0, 3, 5])]), ref, False)
    assert np.max(np.abs(probs.cpu().numpy() - ref_p)) < 1e-12
    # Jacobian
    tape_g = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0)), qb.expval(q.PauliX(wires=2))])
    jac_t = dev.compute_derivatives(tape_g, cfg)
    jac_n = qb.B200Qubit(wires=n).compute_derivatives(tape_g)
    assert isinstance(jac_t, torch.Tensor) and jac_t.is_cuda and jac_t.shape == (2, 2 * n)
    assert np.max(np.abs(jac_t.cpu().numpy() - np.array(jac_n, dtype=float))) < 1e-14
