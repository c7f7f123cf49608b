"""One-shot execution of tapes with mid-circuit measurements — the caller side of the native
MCM path.  Mirror of pennylane/transforms/dynamic_one_shot.py (numpy interface, one
``MeasurementValue`` per measurement, ``postselect_mode`` None / "hw-like"):

* :func:`init_auxiliary_tape` (:214-243) — terminal measurements become per-shot samples, one
  ``sample(MeasurementValue([mcm]))`` is appended per ``MidMeasure``, shots become ``[1] * N``;
* :func:`parse_native_mid_circuit_measurements` (:271-325) with the ``gather_*`` rules
  (:413-575) — combine the per-shot result tuples the device returns
  (``simulate.py:354-381``), dropping the shots a postselection invalidates.

Nothing here touches amplitudes; the device (``simulate._simulate_native_mcm``) does.
"""
from __future__ import annotations

from collections import Counter

import numpy as np

from .mcm import MeasurementValue, is_mcm
from .measurements import SampleMP
from .tape import QuantumScript

_KINDS = ("counts", "expval", "probs", "sample", "var")


def branches(mv: MeasurementValue) -> dict:
    """measurement_value.py:100-107."""
    n = len(mv.measurements)
    return {tuple(int(b) for b in f"{i:0{n}b}"): mv.processing_fn(*(int(b) for b in f"{i:0{n}b}"))
            for i in range(2 ** n)}


def init_auxiliary_tape(circuit) -> QuantumScript:
    """dynamic_one_shot.py:214-243."""
    new_measurements = []
    for m in circuit.measurements:
        if getattr(m, "mv", None) is None:
            new_measurements.append(SampleMP(obs=m.obs) if m.kind == "var" else m)
    for op in circuit.operations:
        if is_mcm(op):
            new_measurements.append(SampleMP(obs=MeasurementValue([op])))
    return QuantumScript(circuit.operations, new_measurements,
                         shots=[1] * circuit.shots.total_shots,
                         trainable_params=circuit.trainable_params)


def _no_shots(m):
    """dynamic_one_shot.py:246-251."""
    return np.nan * np.ones(2 ** len(m.wires)) if m.kind == "probs" else np.nan


def _gather_non_mcm(m, samples, is_valid):
    """dynamic_one_shot.py:413-509."""
    if m.kind == "counts":
        tmp = Counter()
        if getattr(m, "all_outcomes", False) and getattr(m, "mv", None) is not None:
            tmp = Counter({float(v): 0 for v in branches(m.mv).values()})
        for i, d in enumerate(samples):
            tmp.update({k if isinstance(k, str) else float(k): v * int(is_valid[i])
                        for k, v in d.items()})
        if not getattr(m, "all_outcomes", False):
            tmp = Counter({k: v for k, v in tmp.items() if v > 0})
        return dict(sorted(tmp.items()))
    if m.kind == "sample":
        samples = np.concatenate([np.atleast_1d(s) for s in samples]) \
            if isinstance(samples, (list, tuple)) else samples
        return samples[is_valid]
    if m.kind == "expval":
        s = np.squeeze(np.stack([np.asarray(x) for x in samples]))
        return np.sum(s * is_valid) / np.sum(is_valid)
    if m.kind == "probs":
        s = np.stack([np.asarray(x) for x in samples], axis=0)
        return np.sum(s * np.reshape(is_valid, (-1, 1)), axis=0) / np.sum(is_valid)
    if m.kind == "var":
        s = np.squeeze(np.stack([np.asarray(x) for x in samples]))
        ev = np.sum(s * is_valid) / np.sum(is_valid)
        return np.sum((s - ev) ** 2 * is_valid) / np.sum(is_valid)
    raise TypeError(f"Native mid-circuit measurement mode does not support {m.kind} measurements.")


def _gather_mcm(m, mcm_samples: dict, is_valid):
    """dynamic_one_shot.py:512-575 (single ``MeasurementValue``)."""
    vals = np.asarray(m.mv.concretize(mcm_samples))
    if m.kind == "probs":
        vals = np.squeeze(vals, axis=-1) if vals.ndim > 1 else vals
        cnt = np.array([np.count_nonzero(np.logical_and(vals == v, is_valid))
                        for v in branches(m.mv).values()])
        return cnt / np.sum(cnt)
    samples = vals
    if m.kind == "counts":
        samples = [{float(np.asarray(s).item()): 1} for s in vals]
    res = _gather_non_mcm(m, samples, is_valid)
    return np.squeeze(res) if m.kind == "sample" else res


def parse_native_mid_circuit_measurements(circuit, results):
    """dynamic_one_shot.py:271-325.  ``results[i]`` = the per-shot values of aux measurement i."""
    all_mcms = [op for op in circuit.operations if is_mcm(op)]
    mcm_samples = np.hstack([np.reshape(np.vstack([np.asarray(r) for r in res]), (-1, 1))
                             for res in results[-len(all_mcms):]])
    has_ps = np.array([[op.postselect is not None for op in all_mcms]], dtype=mcm_samples.dtype)
    ps = np.array([[0 if op.postselect is None else op.postselect for op in all_mcms]],
                  dtype=mcm_samples.dtype)
    is_valid = np.all(mcm_samples * has_ps == ps, axis=1)
    has_valid = bool(np.any(is_valid))
    mcm_map = {mcm: mcm_samples[:, i:i + 1] for i, mcm in enumerate(all_mcms)}
    out, m_count = [], 0
    for m in circuit.measurements:
        if m.kind not in _KINDS:
            raise TypeError(
                f"Native mid-circuit measurement mode does not support {m.kind} measurements.")
        is_mv = getattr(m, "mv", None) is not None
        if not has_valid:
            out.append(_no_shots(m))
            m_count += int(not is_mv)
        elif is_mv:
            out.append(_gather_mcm(m, mcm_map, is_valid))
        else:
            out.append(_gather_non_mcm(m, results[m_count], is_valid))
            m_count += 1
    return tuple(out) if len(out) > 1 else out[0]


def dynamic_one_shot(tape):
    """dynamic_one_shot.py:85-185: ``(aux_tape, post_processing)``; the post-processing takes
    the device result of ``aux_tape`` (one result tuple per shot)."""
    if not any(is_mcm(o) for o in tape.operations):
        return tape, lambda res: res
    for m in tape.measurements:
        if m.kind not in _KINDS:
            raise TypeError(
                f"Native mid-circuit measurement mode does not support {m.kind} measurements.")
    if not tape.shots:
        raise ValueError("dynamic_one_shot is only supported with finite shots.")
    aux = init_auxiliary_tape(tape)
    num_mp = len(aux.measurements)

    def combine(res):
        if num_mp == 1:
            cols = [tuple(res)]
        else:
            cols = [tuple(r[i] for r in res) for i in range(num_mp)]
        return parse_native_mid_circuit_measurements(tape, cols)

    def processing_fn(results):
        if not tape.shots.has_partitioned_shots:
            return combine(results)
        out, start = [], 0                          # _add_shot_vector_support, :76-83
        for s in tape.shots:
            out.append(combine(results[start:start + s]))
            start += s
        return tuple(out)

    return aux, processing_fn


__all__ = ["dynamic_one_shot", "init_auxiliary_tape", "parse_native_mid_circuit_measurements"]
