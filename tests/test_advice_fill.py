"""Device-side witness fill (csrc/advice.cuh, p2g_fill_advice_device): the advice columns (wires >= 80) of a
trace are recomputed from its routed columns by one thread per row.  The expected values come from an independent restatement --
the witness generators of acir/p2acir.cpp (arithmetic_u32.rs:376, add_many_u32.rs:329, subtraction_u32.rs:298,
range_check_u32.rs:198, comparison.rs:439, plonky2's RandomAccessGenerator and PoseidonGenerator) -- on circuits translated from
opcodes.  CPU: the same function compiled for the host; GPU: the kernel, through the C ABI, followed by p2g_prove_device."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import acir_cases  # noqa: E402

ROUTED = 80


def _circuits(p2g):
    """(name, translated circuit, wires, public inputs) covering every gate that has advice wires."""
    A, EI = p2g.acir, p2g.ecdsa_inputs
    out = []
    for name, circuit, wit in acir_cases.u32_gadget_cases(A):
        out.append((name, circuit, wit))
    by_name = {c[0]: c for c in acir_cases.cases(A)}
    for name in ("memory_write", "memory_read_irregular_block", "and_32", "3x_plus_9y_equals_12"):   # RandomAccess, Poseidon (public inputs)
        _, circuit, wit, _ = by_name[name]
        out.append((name, circuit, wit))
    circuit, wit, _ = EI.circuit_and_witness(A, [EI.deterministic_case(21)], outputs=[1], assert_valid=True)
    out.append(("ecdsa", circuit, wit))
    res = []
    for name, circuit, wit in out:
        tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
        wires, pis = tr.generate_witness(wit)
        res.append((name, tr, wires, pis))
    return res


def test_host_twin_reproduces_the_generators_advice_wires(p2g):
    C = p2g.circuit
    seen = set()
    for name, tr, wires, _ in _circuits(p2g):
        stripped = wires.copy()
        stripped[ROUTED:] = 0
        assert name in ("3x_plus_9y_equals_12",) or wires[ROUTED:].any(), name        # there is something to recompute
        tr.fill_advice_host(stripped)
        bad = np.argwhere(stripped != wires)
        assert bad.size == 0, (name, bad[:5])
        seen |= {g.kind for g in tr.common.gates}
    assert {C.U32_ARITHMETIC, C.U32_ADD_MANY, C.U32_SUBTRACTION, C.U32_RANGE_CHECK, C.COMPARISON, C.RANDOM_ACCESS, C.POSEIDON} <= seen


@pytest.mark.parametrize("workload,bits", [("all_gates", 10), ("ecdsa", 12), ("sha256", 11)])
def test_host_twin_on_the_synthetic_circuits(p2g, workload, bits):
    """A third source of expected values: the circuit synthesiser (synth/synth.cpp) fills its rows with its own per-gate witness
    code and other gate shapes (e.g. ComparisonGate and RandomAccessGate of other sizes, Poseidon rows with swap = 1)."""
    import ctypes as C
    sc = p2g.synth.SyntheticCircuit(bits, workload, num_public_inputs=4, seed=77)
    com = sc.common
    table = np.array([[g.kind, *g.params[:4]] for g in com.gates], dtype=np.uint32).reshape(-1)
    # the synthesiser puts random filler on the wires a row's gate does not constrain, so only what the fill writes is compared:
    # the advice region starts as a sentinel (not a field element)
    SENTINEL = np.uint64(0xFFFFFFFFFFFFFFFF)
    stripped = sc.wires.copy()
    stripped[ROUTED:] = SENTINEL
    L = p2g.acir._acir_lib()
    L.p2a_fill_advice_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32]
    L.p2a_fill_advice_rows(table.ctypes.data_as(C.c_void_p), len(com.gates), sc.row_gate.ctypes.data_as(C.c_void_p),
                           stripped.ctypes.data_as(C.c_void_p), com.degree(), com.config.num_wires, com.config.num_routed_wires)
    written = stripped != SENTINEL
    written[:ROUTED] = False
    bad = np.argwhere(written & (stripped != sc.wires))
    assert bad.size == 0, [(int(c), int(r), com.gates[sc.row_gate[r]].id) for c, r in bad[:5]]
    # Poseidon rows (135 wires) always have advice wires; the small test shapes of the other gates may fit below wire 80
    C_ = p2g.circuit
    per_row = written.sum(axis=0)
    for gi, g in enumerate(com.gates):
        rows = np.nonzero(sc.row_gate == gi)[0]
        if len(rows) and g.kind == C_.POSEIDON:
            assert per_row[rows].min() == 135 - ROUTED, g.id
    assert written.any()


def test_routed_only_witness_generation(p2g):
    """generate_witness(routed_only=True): the generators skip the advice wires; the 80 columns equal those of the full witness,
    and the host twin of the device fill completes them to the full witness."""
    A, EI = p2g.acir, p2g.ecdsa_inputs
    circuit, wit, outs = EI.circuit_and_witness(A, [EI.deterministic_case(22)], outputs=[1], assert_valid=True)
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    full, pis = tr.generate_witness(wit)
    routed, pis2 = tr.generate_witness(wit, routed_only=True)
    assert routed.shape == (ROUTED, full.shape[1]) and pis == pis2 and np.array_equal(routed, full[:ROUTED])
    assert tr.read_witnesses(outs) == {outs[0]: 1}
    rebuilt = np.zeros_like(full)
    rebuilt[:ROUTED] = routed
    tr.fill_advice_host(rebuilt)
    assert np.array_equal(rebuilt, full)


@pytest.mark.gpu
def test_device_fill_reproduces_the_generators_advice_wires_and_the_proof(p2g):
    import torch
    for name, tr, wires, pis in _circuits(p2g):
        data, _ = tr.unpack()
        with data:
            # inside the upload pipeline, on a fresh handle (its staging matrix has never held this trace): only the 80 routed
            # columns cross the boundary (pageable, one array per column)
            routed = [np.array(wires[c]) for c in range(ROUTED)]
            got = data.prove_routed_columns(routed, pis).to_bytes()
            want = data.prove(wires, pis).to_bytes()
            assert got == want, name
            assert data.prove_routed_columns(routed, pis, compressed=True).to_bytes() == data.prove(wires, pis, compressed=True).to_bytes()
            # the stand-alone entry point on a device-resident trace
            full = torch.from_numpy(wires.view(np.int64)).cuda()
            stripped = full.clone()
            stripped[ROUTED:] = 0
            data.fill_advice(stripped)
            assert torch.equal(stripped, full), name
            assert data.prove(stripped, pis).to_bytes() == want, name
