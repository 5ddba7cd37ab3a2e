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
        # the library's own compressed output IS the committed file (what the reference CLI wrote, prove_action.rs:75-78)
        assert data.prove(w, cp.public_inputs, forced_pow_witness=cp.pow_witness, compressed=True).to_bytes() == rec["raw"]
        # and byte-for-byte against the oracle prover, including the deterministic proof-of-work search
        op = corc.OracleProver(cd, cs)
        assert pw.to_bytes() == op.prove(w, cp.public_inputs, forced_pow=cp.pow_witness)
        assert data.prove(w, cp.public_inputs).to_bytes() == op.prove(w, cp.public_inputs)


# ---- synthetic circuits: every gate kind, FRI fold layers, both hashers -----------------------------------------------
CASES = [
    # degree_bits, workload, public inputs, hasher, wires
    (3, "assert_zero", 0, "keccak25", 234),     # BASELINE configs[0] shape ("fibonacci": constant-folded, 8 rows)
    (5, "all_gates", 2, "keccak25", 234),
    (6, "all_gates", 3, "poseidon", 234),
    (8, "ecdsa", 0, "keccak25", 234),
    (10, "all_gates", 2, "keccak25", 234),      # two FRI layers
    (11, "sha256", 4, "poseidon", 234),
    (12, "ecdsa", 1, "keccak25", 234),
    (13, "range", 0, "keccak25", 135),          # multi-pass NTT (n > 11), standard_recursion_config width
    (14, "assert_zero", 0, "keccak25", 234),    # three FRI layers
]


@pytest.mark.parametrize("degree_bits,workload,npi,hasher,wires", CASES)
def test_synthetic_proof_bytes_match_oracle(p2g, corc, degree_bits, workload, npi, hasher, wires):
    from helpers import oracle_prove_and_verify
    cfg = p2g.CircuitConfig(num_wires=wires, hasher=hasher)
    sc = p2g.synth.SyntheticCircuit(degree_bits, workload, config=cfg, num_public_inputs=npi, seed=1000 + degree_bits)
    ref_bytes, op = oracle_prove_and_verify(corc, sc)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        cap, dg = op.cap_and_digest()
        assert data.constants_sigmas_cap == cap and data.circuit_digest == dg
        pw = data.prove(sc.wires, sc.public_inputs)
        L = p2g.lib
        for what in (L.BUF_WIRES_CAP, L.BUF_ZS_PP_CAP, L.BUF_QUOTIENT_CAP, L.BUF_FRI_CAPS):
            assert np.array_equal(data.read(what, np.uint8), op.read(what, np.uint8)), what
        for what in (L.BUF_WIRES_COEFFS, L.BUF_ZS_PP_VALUES, L.BUF_QUOTIENT_CHUNKS, L.BUF_CHALLENGES, L.BUF_FINAL_POLY):
            assert np.array_equal(data.read(what), op.read(what)), what
        assert pw.to_bytes() == ref_bytes
        assert pw.timings["kernel_launches"] > 0 and pw.timings["total_ms"] > 0
        # compressed layout (CompressedProofWithPublicInputs::to_bytes) against the oracle's compressor
        from oracle.pyref import proof, verifier
        cd = op.cd if hasattr(op, "cd") else None
        if cd is not None:
            pr = proof.parse_uncompressed(ref_bytes, cd)
            ch = verifier.verify(pr, cd, cap, dg)
            want = proof.serialize_compressed(proof.compress_proof(pr, ch.indices, cd))
            assert data.prove(sc.wires, sc.public_inputs, compressed=True).to_bytes() == want
        # a second proof on the same handle (buffers reused) is identical: the prover is a pure function of its inputs
        assert data.prove(sc.wires, sc.public_inputs, timings=False).to_bytes() == ref_bytes


def test_prove_from_device_tensor_and_forced_pow(p2g, corc):
    import torch
    from helpers import oracle_prove_and_verify
    sc = p2g.synth.SyntheticCircuit(9, "all_gates", num_public_inputs=1, seed=77)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        host = data.prove(sc.wires, sc.public_inputs)
        d_w = torch.from_numpy(sc.wires.view(np.int64)).cuda()
        dev = data.prove(d_w, sc.public_inputs)
        assert dev.to_bytes() == host.to_bytes()
        pow_w = int(data.read(p2g.lib.BUF_CHALLENGES)[-(sc.config.num_query_rounds + 1)])
        assert data.prove(sc.wires, sc.public_inputs, forced_pow_witness=pow_w).to_bytes() == host.to_bytes()
        # an invalid forced witness is refused, like the verifier would (P2G_EUNSAT), not silently accepted
        bad = 0 if pow_w > 0 else None   # pow_w is the smallest valid witness, so 0 is invalid
        if bad is not None:
            with pytest.raises(p2g.P2GError) as ei:
                data.prove(sc.wires, sc.public_inputs, forced_pow_witness=bad)
            assert ei.value.code == p2g.lib.P2G_EUNSAT


def test_prove_columns_takes_the_witness_as_plonky2_holds_it(p2g, corc):
    """p2g_prove_columns: one pointer per MatrixWitness.wire_values[col] (separately allocated columns); same bytes as the flat
    matrix, uncompressed and compressed, and a null column is refused."""
    import ctypes as C
    sc = p2g.synth.SyntheticCircuit(13, "ecdsa", num_public_inputs=2, seed=78)   # 2^13 x 8 B columns: staged through the ring
    cols = [np.array(sc.wires[i], copy=True) for i in range(sc.wires.shape[0])]    # 234 separate heap allocations
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        flat = data.prove(sc.wires, sc.public_inputs)
        got = data.prove_columns(cols, sc.public_inputs)
        assert got.to_bytes() == flat.to_bytes()
        assert got.timings["h2d_bytes"] == sc.wires.nbytes
        assert (data.prove_columns(cols, sc.public_inputs, compressed=True).to_bytes()
                == data.prove(sc.wires, sc.public_inputs, compressed=True).to_bytes())
        # plonky2's in-memory GoldilocksField may hold any representative < 2^64: x + p for small x is the same field element
        P = 0xFFFFFFFF00000001
        raw = [c.copy() for c in cols]
        bumped = 0
        for c in raw[::7]:
            small = np.nonzero(c < np.uint64(0xFFFFFFFF))[0][:50]
            c[small] += np.uint64(P)
            bumped += len(small)
        assert bumped > 0
        assert data.prove_columns(raw, sc.public_inputs).to_bytes() == flat.to_bytes()
        with pytest.raises(ValueError):
            data.prove_columns(cols[:-1], sc.public_inputs)
        ptrs = (C.c_void_p * len(cols))(*[a.ctypes.data for a in cols])
        ptrs[5] = None
        out = C.create_string_buffer(p2g.lib.lib().p2g_proof_size_bound(data._h))
        ln = C.c_size_t(len(out))
        pis = np.array(sc.public_inputs, dtype=np.uint64)
        rc = p2g.lib.lib().p2g_prove_columns(data._h, ptrs, pis.ctypes.data_as(C.c_void_p), len(pis), None, 0, out, C.byref(ln), None)
        assert rc == p2g.lib.P2G_EBADARG
        # size query: out == NULL answers from the bound without proving
        ln = C.c_size_t(0)
        rc = p2g.lib.lib().p2g_prove(data._h, sc.wires.ctypes.data_as(C.c_void_p), pis.ctypes.data_as(C.c_void_p), len(pis), None, None,
                                     C.byref(ln), None)
        assert rc == p2g.lib.P2G_ESMALLBUF and ln.value == p2g.lib.lib().p2g_proof_size_bound(data._h)


@pytest.mark.parametrize("workload,hasher,npi", [("all_gates", "keccak25", 2), ("ecdsa", "poseidon", 0)])
def test_verifier_data_bytes_round_trip(p2g, workload, hasher, npi):
    """p2g_vk_bytes (write_vk_action.rs:76-79) == the oracle's writer, and the oracle's parser reads every field back."""
    from helpers import oracle_cd
    from oracle.pyref import vk
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
    sc = p2g.synth.SyntheticCircuit(7, workload, config=cfg, num_public_inputs=npi, seed=31)
    cd = oracle_cd(sc.common)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        raw = data.verifier_data_bytes()
        assert raw == vk.serialize_verifier_data(cd, data.constants_sigmas_cap, data.circuit_digest)
        v = vk.parse_verifier_data(raw, sc.common.hash_size)
        assert v["constants_sigmas_cap"] == data.constants_sigmas_cap and v["circuit_digest"] == data.circuit_digest
        assert v["gates"] == [(g.kind, tuple(list(g.params) + [0] * (4 - len(g.params)))[:4]) for g in sc.common.gates]
        assert v["selector_indices"] == sc.common.selector_indices and v["groups"] == [tuple(x) for x in sc.common.groups]
        assert v["k_is"] == sc.common.k_is and v["fri_params"]["degree_bits"] == 7
        assert v["config"]["num_wires"] == 234 and v["config"]["fri_config"]["reduction_strategy"] == ("ConstantArityBits", 4, 5)
        assert v["num_public_inputs"] == npi and v["num_partial_products"] == 9


def test_invalid_witness_is_rejected_by_the_verifier(p2g, corc):
    """The reference's negative tests panic in witness generation (before the seam); at the seam a bad trace still yields
    bytes, and those must NOT verify."""
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    sc = p2g.synth.SyntheticCircuit(6, "assert_zero", seed=3)
    w = sc.wires.copy()
    w[3, 1] = (int(w[3, 1]) + 1) % 0xFFFFFFFF00000001   # break one ArithmeticGate output
    cd = oracle_cd(sc.common)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        pb = data.prove(w, sc.public_inputs).to_bytes()
        with pytest.raises(verifier.VerifyError):
            verifier.verify(proof.parse_uncompressed(pb, cd), cd, data.constants_sigmas_cap, data.circuit_digest)


def test_error_codes(p2g):
    sc = p2g.synth.SyntheticCircuit(4, "assert_zero", seed=4)
    with pytest.raises(ValueError):
        p2g.CircuitData(sc.common, sc.constants_sigmas[:-1])
    bad = sc.constants_sigmas.copy()
    bad[0, 0] = 0xFFFFFFFF00000001  # non-canonical
    with pytest.raises(p2g.P2GError) as ei:
        p2g.CircuitData(sc.common, bad)
    assert ei.value.code == p2g.lib.P2G_EBADARG
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        with pytest.raises(p2g.P2GError) as ei:
            data.prove(sc.wires, [1, 2, 3])       # wrong public input count
        assert ei.value.code == p2g.lib.P2G_EBADARG
        with pytest.raises(ValueError):
            data.prove(sc.wires[:, :8])
        w = sc.wires.copy()
        w[7, 3] = 0xFFFFFFFF00000001 + 5      # non-canonical wire value: refused, not silently reduced
        with pytest.raises(p2g.P2GError) as ei:
            data.prove(w, sc.public_inputs)
        assert ei.value.code == p2g.lib.P2G_EBADARG
        assert data.prove(sc.wires, sc.public_inputs).to_bytes()       # the handle stays usable after an error
        # a device pointer handed to the host-memory entry point is refused, not dereferenced on the host
        import ctypes as C
        import torch
        d_w = torch.from_numpy(sc.wires.view(np.int64)).cuda()
        out = C.create_string_buffer(p2g.lib.lib().p2g_proof_size_bound(data._h))
        ln = C.c_size_t(len(out))
        pis = np.array(sc.public_inputs, dtype=np.uint64)
        rc = p2g.lib.lib().p2g_prove(data._h, C.c_void_p(d_w.data_ptr()), pis.ctypes.data_as(C.c_void_p), len(pis), None, out,
                                     C.byref(ln), None)
        assert rc == p2g.lib.P2G_EBADARG and b"p2g_prove_device" in p2g.lib.lib().p2g_last_error()


@pytest.mark.parametrize("workload,wires", [("all_gates", 234), ("ecdsa", 234), ("all_gates", 136)])
def test_gate_constraints_match_oracle_offdomain(p2g, corc, workload, wires):
    """gate_testing.rs-style: constraints are polynomial identities, so they are compared at RANDOM points (not valid rows),
    where every filter and every constraint is non-zero (plonky2_ecdsa/biguint/gates/gate_testing.rs:85-125)."""
    from helpers import oracle_cd
    cfg = p2g.CircuitConfig(num_wires=wires)
    sc = p2g.synth.SyntheticCircuit(5, workload, config=cfg, num_public_inputs=2, seed=9)
    com = sc.common
    rng = np.random.default_rng(11)
    P = 0xFFFFFFFF00000001
    npts = 300
    consts = rng.integers(0, P, size=(com.num_constants, npts), dtype=np.uint64)
    w = rng.integers(0, P, size=(wires, npts), dtype=np.uint64)
    pi = [int(x) for x in rng.integers(0, P, size=4, dtype=np.uint64)]
    got = p2g.circuit.eval_gate_constraints(com, consts, w, pi)
    ref = corc.eval_gate_constraints(oracle_cd(com), consts, w, pi)
    assert np.array_equal(got, ref)
    assert got.any()


def test_gate_constraints_vanish_on_valid_rows(p2g):
    """test_gate_constraint of each reference gate (e.g. arithmetic_u32.rs:476-569): valid wires => all constraints zero."""
    sc = p2g.synth.SyntheticCircuit(8, "all_gates", num_public_inputs=2, seed=10)
    com = sc.common
    consts = sc.constants_sigmas[:com.num_constants]
    from oracle.pyref.hashing import PoseidonHash
    pi_hash = PoseidonHash.hash_no_pad_elems(sc.public_inputs)
    out = p2g.circuit.eval_gate_constraints(com, consts, sc.wires, pi_hash)
    assert not out.any()


def test_sharded_full_size_2_20_matches_single_gpu(p2g):
    """BASELINE headline size through the coset-sharded path (2 ranks as threads on this GPU): same bytes as one GPU."""
    sc = p2g.synth.SyntheticCircuit(20, "ecdsa", num_public_inputs=4, seed=0xAC1D + 3)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        want = data.prove(sc.wires, sc.public_inputs, timings=False).to_bytes()
        cap = data.constants_sigmas_cap
    group = p2g.sharding.ThreadGroup(2)

    def rank_main(rank, member):
        with p2g.CircuitData(sc.common, sc.constants_sigmas, device=0, shard=member) as d:
            assert d.constants_sigmas_cap == cap
            return d.prove(sc.wires, sc.public_inputs, timings=False).to_bytes()
    for got in group.run(rank_main):
        assert got == want


def test_full_size_2_20_properties(p2g):
    """BASELINE headline size (2^20 rows, 234 wires, ECDSA-shaped mix).  The oracle does not finish this in seconds, so the
    size-independent checks are: the proof is a deterministic function of the inputs, its openings are consistent with
    the committed quotient (PLONK identity at zeta, checked by the oracle verifier's algebra on the opening set only),
    and the in-proof Merkle paths hash to the caps (full verifier)."""
    from helpers import oracle_cd
    from oracle.pyref import proof, verifier
    sc = p2g.synth.SyntheticCircuit(20, "ecdsa", num_public_inputs=4, seed=0xAC1D + 3)
    cd = oracle_cd(sc.common)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        a = data.prove(sc.wires, sc.public_inputs)
        b = data.prove(sc.wires, sc.public_inputs)
        assert a.to_bytes() == b.to_bytes()
        verifier.verify(proof.parse_uncompressed(a.to_bytes(), cd), cd, data.constants_sigmas_cap, data.circuit_digest)
