"""Coset-sharded proofs: `world` ranks -- here threads of one process with handles on cuda:0, exchanging through the
same p2g_allgather_fn callback the NCCL binding uses -- must produce the bytes of the single-GPU prover and of the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # degree_bits, workload, public inputs, hasher, world
    (5, "all_gates", 2, "keccak25", 2),
    (6, "all_gates", 3, "poseidon", 4),
    (10, "all_gates", 2, "keccak25", 8),
    (12, "ecdsa", 1, "keccak25", 4),      # column-sharded inverse NTT + coefficient all-gather (ncols >= 4 * world)
    (13, "range", 0, "poseidon", 2),
    (14, "assert_zero", 0, "keccak25", 8),
]


@pytest.mark.parametrize("degree_bits,workload,npi,hasher,world", CASES)
def test_sharded_proof_bytes_match_oracle(p2g, corc, degree_bits, workload, npi, hasher, world):
    from helpers import oracle_prove_and_verify
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
    sc = p2g.synth.SyntheticCircuit(degree_bits, workload, config=cfg, num_public_inputs=npi, seed=4000 + degree_bits)
    ref_bytes, op = oracle_prove_and_verify(corc, sc)
    cap, dg = op.cap_and_digest()
    group = p2g.sharding.ThreadGroup(world)

    def rank_main(rank, member):
        with p2g.CircuitData(sc.common, sc.constants_sigmas, device=0, shard=member) as data:
            assert data.constants_sigmas_cap == cap and data.circuit_digest == dg
            info = data.read(p2g.lib.BUF_SHARD_INFO)
            assert (int(info[0]), int(info[1])) == (rank, world) and int(info[3]) == 1    # peers mapped: fused exchange active
            a = data.prove(sc.wires, sc.public_inputs)
            b = data.prove(sc.wires, sc.public_inputs, timings=False)      # buffers reused
            caps = [bytes(data.read(w, np.uint8)) for w in (p2g.lib.BUF_WIRES_CAP, p2g.lib.BUF_ZS_PP_CAP, p2g.lib.BUF_QUOTIENT_CAP)]
            return a.to_bytes(), b.to_bytes(), caps

    outs = group.run(rank_main)
    want_caps = [bytes(op.read(w, np.uint8)) for w in (p2g.lib.BUF_WIRES_CAP, p2g.lib.BUF_ZS_PP_CAP, p2g.lib.BUF_QUOTIENT_CAP)]
    for a, b, caps in outs:
        assert caps == want_caps
        assert a == ref_bytes and b == ref_bytes


def test_sharded_create_rejects_bad_world(p2g):
    sc = p2g.synth.SyntheticCircuit(5, "assert_zero", num_public_inputs=0, seed=1)
    member = p2g.sharding.ThreadGroup(2).member(0)
    member.world = 3
    with pytest.raises(p2g.P2GError):
        p2g.CircuitData(sc.common, sc.constants_sigmas, shard=member)
    member.world = 16      # > 2^rate_bits
    with pytest.raises(p2g.P2GError):
        p2g.CircuitData(sc.common, sc.constants_sigmas, shard=member)


def test_sharded_without_peer_memory_uses_the_allgather(p2g, corc, monkeypatch):
    """P2G_NO_PEER=1: the inverse-NTT column blocks go through the allgather callback instead of peer stores; same bytes."""
    from helpers import oracle_prove_and_verify
    monkeypatch.setenv("P2G_NO_PEER", "1")
    sc = p2g.synth.SyntheticCircuit(12, "ecdsa", config=p2g.CircuitConfig.wide_ecc_config(), num_public_inputs=1, seed=4012)
    ref_bytes, _ = oracle_prove_and_verify(corc, sc)
    group = p2g.sharding.ThreadGroup(4)

    def rank_main(rank, member):
        with p2g.CircuitData(sc.common, sc.constants_sigmas, device=0, shard=member) as data:
            assert int(data.read(p2g.lib.BUF_SHARD_INFO)[3]) == 0
            return data.prove(sc.wires, sc.public_inputs).to_bytes()
    assert all(b == ref_bytes for b in group.run(rank_main))
