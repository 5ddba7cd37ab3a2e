"""CPU: the reference's EcdsaSecp256k1 translator (circuit_translation/ecdsa_secp256k1_translator.rs) and the plonky2_ecdsa gadget
stack under it (biguint, non-native fields, affine curve arithmetic, GLV + windowed MSM), restated on the C++ builder with the
witness generators of those gadgets: BASELINE configs[3]'s circuit from the real opcode.  The reference's acceptance test is
test_precompiled.rs:8-44 (`ecdsa_secp256k1`: prove, then `circuit_data.verify(proof).is_ok()`); here the oracle plays prover and
verifier, every gate and copy constraint is checked on the trace, and the output bit is compared with an independent integer
model of the same operation chain (tests/ecdsa_model.py)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import ecdsa_model  # noqa: E402
from test_acir_cpu import _check_trace  # noqa: E402

# circuit_translation/tests/factories/noir_circuits_for_testing/ecdsa_secp256k1/Prover.toml
REF_MSG = [0xce, 0x7d, 0xf6, 0xb1, 0xb2, 0x85, 0x2c, 0x5c, 0x15, 0x6b, 0x68, 0x3a, 0x9f, 0x8d, 0x4a, 0x8d,
           0xae, 0xda, 0x2f, 0x35, 0xf0, 0x25, 0xcb, 0x0c, 0xf3, 0x49, 0x43, 0xdc, 0xac, 0x70, 0xd6, 0xa3]
REF_PKX = [0x7b, 0x83, 0xad, 0x6a, 0xfb, 0x12, 0x09, 0xf3, 0xc8, 0x2e, 0xbe, 0xb0, 0x8c, 0x0c, 0x5f, 0xa9,
           0xbf, 0x67, 0x24, 0x54, 0x85, 0x06, 0xf2, 0xfb, 0x4f, 0x99, 0x1e, 0x22, 0x87, 0xa7, 0x70, 0x90]
REF_PKY = [0x17, 0x73, 0x16, 0xca, 0x82, 0xb0, 0xbd, 0xf7, 0x0c, 0xd9, 0xde, 0xe1, 0x45, 0xc3, 0x00, 0x2c,
           0x0d, 0xa1, 0xd9, 0x26, 0x26, 0x44, 0x98, 0x75, 0x97, 0x2a, 0x27, 0x80, 0x7b, 0x73, 0xb4, 0x2e]
REF_SIG = [0x6f, 0x01, 0x56, 0x09, 0x1c, 0xbe, 0x91, 0x2f, 0x2d, 0x5d, 0x12, 0x15, 0xcc, 0x3c, 0xd8, 0x1c,
           0x09, 0x63, 0xc8, 0x83, 0x9b, 0x93, 0xaf, 0x60, 0xe0, 0x92, 0x1b, 0x61, 0xa1, 0x9c, 0x54, 0x30,
           0x0c, 0x71, 0x00, 0x6d, 0xd9, 0x3f, 0x35, 0x08, 0xc4, 0x32, 0xda, 0xca, 0x21, 0xdb, 0x00, 0x95,
           0xf4, 0xb1, 0x65, 0x42, 0x78, 0x2b, 0x79, 0x86, 0xf4, 0x8a, 0x5d, 0x0a, 0xe3, 0xc5, 0x83, 0xd4]


def test_the_reference_vector_is_a_valid_big_endian_signature(p2g):
    """Sanity of the helper curve code: Noir's vector verifies when its bytes are read big-endian, as Noir means them."""
    EI = p2g.ecdsa_inputs
    be = lambda b: int.from_bytes(bytes(b), "big")   # noqa: E731
    r = be(REF_SIG[:32])
    assert EI.recovered_x((be(REF_PKX), be(REF_PKY)), r, be(REF_SIG[32:]), be(REF_MSG)) % EI.N == r
    q, r, s, h = EI.deterministic_case(7)
    assert EI.recovered_x(q, r, s, h) % EI.N == r
    assert ecdsa_model.glv_mul(EI.G, 12345678901234567890123456789) == EI.point_mul(12345678901234567890123456789, EI.G)


def test_ecdsa_outputs_follow_the_reference_operation_chain(p2g):
    """Seven opcodes in one circuit: the reference's own vector (little-endian reads turn it into an off-curve input, and its
    `cmp_biguint` makes the output r <= x(R), not equality), two valid little-endian signatures, tampered messages, a tampered
    signature and a swapped key.  Every output witness the generators compute equals the integer model's, and a contradicting
    provided output is refused."""
    A, EI = p2g.acir, p2g.ecdsa_inputs
    enc = EI.encode
    cases, want = [(REF_PKX, REF_PKY, REF_SIG, REF_MSG)], []
    for seed in (1, 2):
        q, r, s, h = EI.deterministic_case(seed)
        cases.append((enc(q[0]), enc(q[1]), enc(r) + enc(s), enc(h)))
    q, r, s, h = EI.deterministic_case(3)
    cases.append((enc(q[0]), enc(q[1]), enc(r) + enc(s), enc((h + 1) % EI.N)))
    cases.append((enc(q[0]), enc(q[1]), enc(r) + enc((s * 3) % EI.N), enc(h)))
    q2 = EI.deterministic_case(4)[0]
    cases.append((enc(q2[0]), enc(q2[1]), enc(r) + enc(s), enc(h)))
    for seed in range(5, 60):                    # a tampered message whose recovered x falls below r: output 0
        q, r, s, h = EI.deterministic_case(seed)
        c = (enc(q[0]), enc(q[1]), enc(r) + enc(s), enc(h ^ 1))
        if ecdsa_model.circuit_output(*c)[0] == 0:
            cases.append(c)
            break
    want = [ecdsa_model.circuit_output(*c)[0] for c in cases]
    assert want[1] == want[2] == 1 and 0 in want, want          # valid signatures verify; the set exercises both outputs
    circuit, wit, outs = EI.circuit_and_witness(A, cases, range_checks=False)
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    tr.generate_witness(wit)
    got = tr.read_witnesses(outs)
    assert [got[o] for o in outs] == want
    flip = want.index(0)
    with pytest.raises(A.TranslationError):
        tr.generate_witness({**wit, outs[flip]: 1})


def test_ecdsa_circuit_from_the_opcode_proves_and_verifies(p2g, corc):
    """The Noir program of the reference's test (160 byte inputs with RANGE 8, one EcdsaSecp256k1, assert(valid)): 2^17 rows on the
    reference's 234-wire configuration, 17+ gate types (all five custom u32 / comparison gates, BaseSum<4>, RandomAccess(4));
    the trace satisfies every gate and copy constraint, the oracle proves it and the oracle verifier accepts."""
    A, EI, C = p2g.acir, p2g.ecdsa_inputs, p2g.circuit
    circuit, wit, outs = EI.circuit_and_witness(A, [EI.deterministic_case(1)], outputs=[1], assert_valid=True)
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    assert tr.common.degree_bits() == 17 and 90_000 < tr.rows_used() < 110_000
    kinds = {g.kind for g in tr.common.gates}
    assert {C.U32_ARITHMETIC, C.U32_ADD_MANY, C.U32_SUBTRACTION, C.U32_RANGE_CHECK, C.COMPARISON, C.BASE_SUM, C.RANDOM_ACCESS,
            C.ARITHMETIC} <= kinds
    wires, pis = tr.generate_witness(wit)
    assert pis == []
    cd, moved = _check_trace(p2g, corc, tr, wires, pis)
    assert moved > 1_000_000
    from oracle.pyref import proof, verifier
    op = corc.OracleProver(cd, tr.constants_sigmas)
    pb = op.prove(wires, pis)
    cap, dg = op.cap_and_digest()
    verifier.verify(proof.parse_uncompressed(pb, cd), cd, cap, dg)
    with pytest.raises(A.TranslationError):       # assert(valid) on a signature the circuit rejects
        for seed in range(5, 60):
            q, r, s, h = EI.deterministic_case(seed)
            if ecdsa_model.circuit_output(EI.encode(q[0]), EI.encode(q[1]), EI.encode(r) + EI.encode(s), EI.encode(h ^ 1))[0] == 0:
                break
        _, badwit, _ = EI.circuit_and_witness(A, [(q, r, s, h ^ 1)], outputs=[1], assert_valid=True)
        tr.generate_witness(badwit)


def test_witness_generation_is_the_same_on_one_thread_and_on_many(p2g):
    """The generators of the two scalar multiplications of every EcdsaSecp256k1 opcode (and of different opcodes) run on different
    threads, claiming copy classes with a compare-exchange: the wire matrix must not depend on the schedule."""
    import numpy as np
    A, EI = p2g.acir, p2g.ecdsa_inputs
    circuit, wit, outs = EI.circuit_and_witness(A, [EI.deterministic_case(i) for i in (11, 12)], range_checks=True)
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    try:
        A.set_threads(1)
        w1, _ = tr.generate_witness(wit)
        o1 = tr.read_witnesses(outs)
        A.set_threads(0)
        for _ in range(2):
            wn, _ = tr.generate_witness(wit)
            assert np.array_equal(w1, wn) and tr.read_witnesses(outs) == o1
    finally:
        A.set_threads(0)
    assert list(o1.values()) == [1, 1]
