"""End-to-end parity of the CUDA prover (through the C ABI) with the oracle: byte-identical proofs.
The reference's own tests only assert `verify(prove(..)).is_ok()` (e.g.
plonky2-backend/src/circuit_translation/tests/factories/utils.rs:26); here acceptance AND bytes are checked."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def product_common(p2g, cd):
    """CommonCircuitData of the product from the oracle's description of the same circuit."""
    C = p2g.circuit
    cfg = C.CircuitConfig(num_wires=cd.num_wires, num_routed_wires=cd.num_routed, hasher=cd.hasher,
                          proof_of_work_bits=cd.pow_bits, num_query_rounds=cd.num_queries, cap_height=cd.cap_height)
    gates = [C.Gate(g.kind, tuple(g.params)) for g in cd.gates]
    com = C.CommonCircuitData(cfg, cd.degree_bits, gates, cd.num_public_inputs)
    assert [g.kind for g in com.gates] == [g.kind for g in cd.gates]
    assert com.selector_indices == cd.selector_indices and com.groups == cd.groups
    return com


@pytest.mark.parametrize("name", ["basic_if", "basic_div"])
def test_golden_proofs_regenerated_on_gpu(p2g, corc, name):
    from oracle.pyref import golden, proof, verifier
    rec = golden.recover(name)
    cd = rec["cd"]
    cs = np.array(rec["trace"]["constants"] + rec["trace"]["sigmas"], dtype=np.uint64)
    w = np.array(rec["trace"]["wires"], dtype=np.uint64)
    cp = rec["cproof"]
    with p2g.circuit.CircuitData(product_common(p2g, cd), cs) as data:
        assert data.constants_sigmas_cap == rec["cs_cap"]
        assert data.circuit_digest == verifier.circuit_digest(cd, rec["cs_cap"])
        pw = data.prove(w, cp.public_inputs, forced_pow_witness=cp.pow_witness)
        pr = proof.parse_uncompressed(pw.to_bytes(), cd)
        ch = verifier.verify(pr, cd, rec["cs_cap"])
        assert proof.serialize_compressed(proof.compress_proof(pr, ch.indices, cd)) == rec["raw"]
        # and byte-for-byte against the oracle prover, including the deterministic proof-of-work search
        op = corc.OracleProver(cd, cs)
        assert pw.to_bytes() == op.prove(w, cp.public_inputs, forced_pow=cp.pow_witness)
        assert data.prove(w, cp.public_inputs).to_bytes() == op.prove(w, cp.public_inputs)
