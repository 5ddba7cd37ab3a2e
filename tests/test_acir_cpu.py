"""CPU: the ACIR translator + mini CircuitBuilder + witness generators on the reference's own
translator tests (circuit_translation/tests/test_assert_zero.rs, test_blackbox.rs, test_memory_operations.rs): the generated
trace satisfies every gate constraint and every copy constraint, the ORACLE prover proves it and the ORACLE verifier accepts --
the reference's `assert!(circuit_data.verify(proof).is_ok())` (tests/factories/utils.rs:16-27) -- and bad witnesses are refused
where the reference panics."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import acir_cases  # noqa: E402

P = 0xFFFFFFFF00000001


def _check_trace(p2g, corc, tr, wires, pis):
    from helpers import oracle_cd
    from oracle.pyref.hashing import PoseidonHash
    com = tr.common
    cd = oracle_cd(com)
    n = com.degree()
    # 1. gate constraints vanish on every row (gate_testing.rs property, through the oracle's evaluators)
    pi_hash = PoseidonHash.hash_no_pad_elems(pis) if pis else [0, 0, 0, 0]
    out = corc.eval_gate_constraints(cd, tr.constants_sigmas[:com.num_constants], wires, pi_hash)
    assert not out.any(), f"gate constraint {np.argwhere(out)[0]} does not vanish"
    # 2. copy constraints: the value on a routed wire equals the value on the wire sigma sends it to (sigma values are
    #    k_is[col'] * omega^row'; every one of them must name a routed wire position)
    w = pow(7277203076849721926, 1 << (32 - com.degree_bits()), P)
    sub, x = [], 1
    for r in range(n):
        sub.append(x)
        x = x * w % P
    ids = np.array([k * s % P for k in com.k_is for s in sub], dtype=np.uint64)          # position c * n + r
    order = np.argsort(ids, kind="stable")
    sorted_ids = ids[order]
    sig = np.ascontiguousarray(tr.constants_sigmas[com.num_constants:]).reshape(-1)
    at = np.searchsorted(sorted_ids, sig)
    assert (at < len(sorted_ids)).all() and (sorted_ids[np.minimum(at, len(sorted_ids) - 1)] == sig).all(), "sigma value is not a routed wire position"
    dest = order[at]
    routed = np.ascontiguousarray(wires[:80]).reshape(-1)
    assert (routed[dest] == routed).all(), "a copy constraint is violated"
    assert len(np.unique(dest)) == len(dest), "sigma is not a permutation"
    moved = int((dest != np.arange(len(dest))).sum())
    return cd, moved


NUM_CASES = 26


@pytest.mark.parametrize("case", range(NUM_CASES))
def test_reference_translator_cases(p2g, corc, case):
    all_cases = acir_cases.cases(p2g.acir)
    assert len(all_cases) == NUM_CASES
    name, circuit, witness, want_pis = all_cases[case]
    tr = p2g.acir.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    wires, pis = tr.generate_witness(witness)
    if want_pis is not None:
        assert pis == want_pis, name
    cd, moved = _check_trace(p2g, corc, tr, wires, pis)
    assert moved > 0, "the circuit has copy constraints"
    # the oracle proves and verifies (the reference test's acceptance criterion)
    from oracle.pyref import proof, verifier
    op = corc.OracleProver(cd, tr.constants_sigmas)
    pb = op.prove(wires, pis)
    cap, dg = op.cap_and_digest()
    pr = proof.parse_uncompressed(pb, cd)
    verifier.verify(pr, cd, cap, dg)
    assert [int(x) for x in pr.public_inputs] == pis


def test_bad_witnesses_are_refused_like_the_reference_panics(p2g):
    A = p2g.acir
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.AssertZero(A.Expression([], [(1, 0)], -4))], [0]))
    with pytest.raises(A.TranslationError):        # x = 5 where x - 4 = 0 is asserted
        tr.generate_witness({0: 5})
    for bits in (8, 16, 32):                       # test_blackbox.rs:17-25, 36-44, 55-63
        tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.Range(0, bits)], [0]))
        with pytest.raises(A.TranslationError):
            tr.generate_witness({0: 1 << bits})
    with pytest.raises(A.TranslationError, match="33 bits"):     # test_blackbox.rs:75-82
        A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.Range(0, 64)], [0]))
    # a provided output that disagrees with the computed one
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.And(0, 1, 8, 2)], [0, 1]))
    tr.generate_witness({0: 3, 1: 5, 2: 1})
    with pytest.raises(A.TranslationError):
        tr.generate_witness({0: 3, 1: 5, 2: 7})
    with pytest.raises(A.TranslationError):        # index beyond the real block length (3 cells padded to 4): memory_translator.rs:55-85
        tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.MemoryInit(0, [0, 1, 2]), A.MemoryRead(0, 3, 4)], [0, 1, 2, 3]))
        tr.generate_witness({0: 1, 1: 2, 2: 3, 3: 3})
    with pytest.raises(A.TranslationError):        # memory read past the block
        tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.MemoryInit(0, [0, 1]), A.MemoryRead(0, 2, 3)], [0, 1, 2]))
        tr.generate_witness({0: 1, 1: 2, 2: 5})


def test_assert_zero_chain_packs_rows_like_the_builder(p2g, corc):
    """A 300-opcode AssertZero chain: ArithmeticGate rows hold 20 operations with equal constants, constants live in
    ConstantGate rows, one PublicInputGate + one PoseidonGate row hash the public input; the trace is valid."""
    circuit, wit = acir_cases.chain(p2g.acir, 300)
    tr = p2g.acir.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    kinds = sorted(g.kind for g in tr.common.gates)
    C = p2g.circuit
    assert kinds == sorted([C.NOOP, C.CONSTANT, C.PUBLIC_INPUT, C.ARITHMETIC, C.POSEIDON])
    assert tr.common.degree_bits() == 8          # 6 gate operations per opcode (x * 1 and x + 0 need none) at 20 per row + 300 distinct constants at 2 per row
    wires, pis = tr.generate_witness(wit)
    assert pis == [1]
    _check_trace(p2g, corc, tr, wires, pis)


def test_sha256_compression_known_answer(p2g, corc):
    """Sha256Compression translated like sha256_translator.rs:60-273 (bit decompositions, rotations as re-wired fresh bool targets,
    xor / choose / majority / ripple-carry adders): SHA-256("abc") comes out of the witness generators (test_sha256_internal.rs:481-549
    pins the same digest), a wrong digest contradicts a copy constraint, and the trace satisfies the circuit."""
    A = p2g.acir
    circuit, wit, out_ids, digest = acir_cases.sha256_circuit(A, b"abc")
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    assert tr.common.degree_bits() == 15
    wires, pis = tr.generate_witness({**wit, **{out_ids[i]: digest[i] for i in range(8)}})
    with pytest.raises(A.TranslationError):
        tr.generate_witness({**wit, out_ids[0]: digest[0] ^ 1})
    assert pis[:16] == [wit[i] for i in range(16)]
    _check_trace(p2g, corc, tr, wires, pis)


NUM_U32_CASES = 6


@pytest.mark.parametrize("case", range(NUM_U32_CASES))
def test_u32_gadgets_and_their_generators(p2g, corc, case):
    """The reference's custom gates filled by restated generators (arithmetic_u32.rs:376, add_many_u32.rs:329, subtraction_u32.rs:298,
    range_check_u32.rs:198, comparison.rs:439): outputs match integer arithmetic, every gate constraint vanishes, the copy
    constraints hold, the oracle proves and verifies; a wrong expected output is refused."""
    all_cases = acir_cases.u32_gadget_cases(p2g.acir)
    assert len(all_cases) == NUM_U32_CASES
    name, circuit, witness = all_cases[case]
    A, C = p2g.acir, p2g.circuit
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    kinds = {g.kind for g in tr.common.gates}
    want = {"mul_add_u32": {C.U32_ARITHMETIC}, "add_many_u32": {C.U32_ARITHMETIC, C.U32_ADD_MANY}, "sub_u32": {C.U32_SUBTRACTION},
            "range_check_u32": {C.U32_RANGE_CHECK}, "cmp_le": {C.COMPARISON},
            "biguint_add_96": {C.U32_ADD_MANY, C.U32_RANGE_CHECK, C.COMPARISON, C.ARITHMETIC}}[name]
    assert want <= kinds, (name, kinds)
    wires, pis = tr.generate_witness(witness)
    cd, moved = _check_trace(p2g, corc, tr, wires, pis)
    assert moved > 0
    from oracle.pyref import proof, verifier
    op = corc.OracleProver(cd, tr.constants_sigmas)
    pb = op.prove(wires, pis)
    cap, dg = op.cap_and_digest()
    verifier.verify(proof.parse_uncompressed(pb, cd), cd, cap, dg)
    if name != "range_check_u32":
        last = max(witness)          # the last witness of every case is a computed output
        with pytest.raises(A.TranslationError):
            tr.generate_witness({**witness, last: witness[last] ^ 1})


def test_range_check_u32_refuses_a_33_bit_value(p2g, corc):
    A = p2g.acir
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.RangeCheckU32([0, 1])], [0, 1]))
    wires, pis = tr.generate_witness({0: 7, 1: 1 << 32})     # the generator truncates like the reference's `as u32`...
    with pytest.raises(AssertionError):                      # ...so the gate's constraints do not vanish on the trace
        _check_trace(p2g, corc, tr, wires, pis)


def test_brillig_calls_and_directives_are_ignored_like_the_reference_does(p2g):
    """mod.rs:97-104: unconstrained opcodes leave no trace in the circuit."""
    A = p2g.acir
    az = A.AssertZero(A.Expression([(1, 0, 1)], [(P - 1, 2)], 0))
    plain = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([az], [0], [1]))
    mixed = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(A.Circuit([A.BrilligCall(0), az, A.Directive("ToLeRadix")], [0], [1]))
    assert plain.common.degree_bits() == mixed.common.degree_bits() and np.array_equal(plain.constants_sigmas, mixed.constants_sigmas)
    w1, p1 = plain.generate_witness({0: 3, 1: 4, 2: 12})
    w2, p2 = mixed.generate_witness({0: 3, 1: 4, 2: 12})
    assert np.array_equal(w1, w2) and p1 == p2 == [3]
