"""Shared test helpers: bridge between the product's CommonCircuitData and the oracle's CommonData."""


def oracle_cd(common):
    from oracle.pyref.circuit import CommonData
    from oracle.pyref.gates import Gate as OGate
    cfg = common.config
    gates = [OGate(g.kind, *g.params) for g in common.gates]
    cd = CommonData(common.degree_bits(), gates, num_wires=cfg.num_wires, num_routed=cfg.num_routed_wires,
                    num_public_inputs=common.num_public_inputs, hasher=cfg.hasher, num_challenges=cfg.num_challenges,
                    rate_bits=cfg.rate_bits, cap_height=cfg.cap_height, pow_bits=cfg.proof_of_work_bits,
                    num_queries=cfg.num_query_rounds, qdf=cfg.max_quotient_degree_factor)
    # host-logic parity: gate order, selector groups, derived counts
    assert [(g.kind, tuple(g.params)) for g in cd.gates] == [(g.kind, tuple(g.params)) for g in common.gates]
    assert cd.selector_indices == common.selector_indices and [tuple(x) for x in cd.groups] == [tuple(x) for x in common.groups]
    assert cd.num_constants == common.num_constants and cd.num_gate_constraints == common.num_gate_constraints
    assert cd.num_partial_products == common.num_partial_products and cd.arity_bits == common.reduction_arity_bits
    assert cd.k_is == common.k_is
    return cd


def oracle_prove_and_verify(corc, sc, forced_pow=None):
    """Runs the oracle prover on a SyntheticCircuit and checks the oracle verifier accepts.  Returns (bytes, prover)."""
    from oracle.pyref import proof, verifier
    cd = oracle_cd(sc.common)
    op = corc.OracleProver(cd, sc.constants_sigmas)
    pb = op.prove(sc.wires, sc.public_inputs, forced_pow=forced_pow)
    cap, dg = op.cap_and_digest()
    verifier.verify(proof.parse_uncompressed(pb, cd), cd, cap, dg)
    return pb, op
