"""ORACLE — CPU restatement of the reference's statevector hot path (test infrastructure only).

A plain-numpy restatement of ``pennylane/devices/qubit/*.py`` (PennyLane v0.46.0-dev81), calling
the same numpy primitives in the same order as the reference.  It exists to CHECK the CUDA
engine; it is never the thing measured or shipped.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it; ``pennylane_b200`` never does (tests/test_no_oracle_in_product.py enforces that).

Parity pinning: the reference itself cannot be imported in the build container (autograd,
autoray, rustworkx, ... are absent), so this restatement is pinned
against every known-answer vector the reference's own tests hold for the path
(tests/golden/reference_known_answers.json, transcribed with file:line provenance by
tests/golden/make_golden.py) and against numpy's own ``Generator.choice`` for sampling.
"""
