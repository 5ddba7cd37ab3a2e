"""GPU: circuits translated from ACIR opcodes (not synthetic gate mixes) through the whole product path: translate -> witness ->
p2g_prove, compared byte for byte with the oracle prover on the same payload and accepted by the oracle verifier -- the
reference's translator tests (circuit_translation/tests/*.rs, `assert!(circuit_data.verify(proof).is_ok())`) on the CUDA prover."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import acir_cases  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [0, 4, 8, 10, 14, 16, 19, 21, 23, 24, 25])
def test_translated_circuits_prove_on_the_gpu(p2g, corc, case):
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    name, circuit, witness, want_pis = acir_cases.cases(p2g.acir)[case]
    tr = p2g.acir.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    wires, pis = tr.generate_witness(witness)
    cd = oracle_cd(tr.common)
    data, _ = tr.unpack()
    with data:
        pw = data.prove(wires, pis)
        pr = proof.parse_uncompressed(pw.to_bytes(), cd)
        verifier.verify(pr, cd, data.constants_sigmas_cap, data.circuit_digest)
        assert [int(x) for x in pr.public_inputs] == pis
        op = corc.OracleProver(cd, tr.constants_sigmas)
        assert pw.to_bytes() == op.prove(wires, pis), name


@pytest.mark.parametrize("case", range(6))
def test_u32_gadget_circuits_prove_on_the_gpu(p2g, corc, case):
    """The reference's custom u32 / comparison gates on rows filled by their restated witness generators (not synthetic rows):
    built with the gadgets, witnessed, proved by the CUDA prover, byte-compared with the oracle prover and verified."""
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    name, circuit, witness = acir_cases.u32_gadget_cases(p2g.acir)[case]
    tr = p2g.acir.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    wires, pis = tr.generate_witness(witness)
    cd = oracle_cd(tr.common)
    data, _ = tr.unpack()
    with data:
        got = data.prove(wires, pis).to_bytes()
        verifier.verify(proof.parse_uncompressed(got, cd), cd, data.constants_sigmas_cap, data.circuit_digest)
    assert got == corc.OracleProver(cd, tr.constants_sigmas).prove(wires, pis), name


def test_assert_zero_chain_from_real_opcodes(p2g, corc):
    """BASELINE configs[1] shape from real AssertZero opcodes (2^14 rows): translated, witnessed, proved, byte-compared."""
    from helpers import oracle_cd
    circuit, wit = acir_cases.chain(p2g.acir, 20000)
    tr = p2g.acir.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    assert tr.common.degree_bits() >= 14
    wires, pis = tr.generate_witness(wit)
    cd = oracle_cd(tr.common)
    data, _ = tr.unpack()
    with data:
        got = data.prove(wires, pis).to_bytes()
    assert got == corc.OracleProver(cd, tr.constants_sigmas).prove(wires, pis)


def test_config2_real_sha256_circuit_2_18(p2g, corc):
    """BASELINE configs[2] from real opcodes: SHA-256 of a 448-byte message = 8 chained Sha256Compression opcodes translated like
    sha256_translator.rs -> 2^18 rows (not the "SHA-256-shaped" synthetic mix).  The witness generators produce the digest hashlib
    computes; the CUDA prover's bytes equal the oracle prover's on the same payload, and the oracle verifier accepts them."""
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    A = p2g.acir
    message = bytes((7 * i + 3) & 0xFF for i in range(448))
    circuit, wit, out_ids, digest = acir_cases.sha256_circuit(A, message)
    assert len(circuit.opcodes) == 8
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    assert tr.common.degree_bits() == 18
    wires, pis = tr.generate_witness({**wit, **{out_ids[i]: digest[i] for i in range(8)}})
    cd = oracle_cd(tr.common)
    data, _ = tr.unpack()
    with data:
        got = data.prove(wires, pis).to_bytes()
        verifier.verify(proof.parse_uncompressed(got, cd), cd, data.constants_sigmas_cap, data.circuit_digest)
    assert got == corc.OracleProver(cd, tr.constants_sigmas).prove(wires, pis)


def _ecdsa_parity(p2g, corc, n_signatures, want_bits):
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    A, EI = p2g.acir, p2g.ecdsa_inputs
    cases = [EI.deterministic_case(100 + i) for i in range(n_signatures)]
    circuit, wit, outs = EI.circuit_and_witness(A, cases, outputs=[1] * n_signatures, assert_valid=True)
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
    assert tr.common.degree_bits() == want_bits
    wires, pis = tr.generate_witness(wit)
    cd = oracle_cd(tr.common)
    data, _ = tr.unpack()
    with data:
        got = data.prove(wires, pis).to_bytes()
        verifier.verify(proof.parse_uncompressed(got, cd), cd, data.constants_sigmas_cap, data.circuit_digest)
    assert got == corc.OracleProver(cd, tr.constants_sigmas).prove(wires, pis)


def test_config3_real_ecdsa_circuit_2_17(p2g, corc):
    """BASELINE configs[3] from the real opcode: the Noir program of the reference's `ecdsa_secp256k1` test (160 RANGE-checked byte
    inputs, one EcdsaSecp256k1 blackbox call, assert(valid)) translated like ecdsa_secp256k1_translator.rs over the plonky2_ecdsa
    gadgets -> 98.9 K rows = 2^17 on the 234-wire configuration, 18 gate types.  Witness from the restated generators; the CUDA
    prover's bytes equal the oracle prover's and the oracle verifier accepts them."""
    _ecdsa_parity(p2g, corc, 1, 17)


def test_config3_eight_real_ecdsa_signatures_2_20(p2g, corc):
    """The same program verifying eight signatures: 2^20 rows from real opcodes (BASELINE's size for configs[3])."""
    _ecdsa_parity(p2g, corc, 8, 20)
