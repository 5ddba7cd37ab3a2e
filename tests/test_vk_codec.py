"""CPU: the oracle's verification-key codec (VerifierCircuitData::to_bytes with BackendGateSerializer, write_vk_action.rs:35-79)
is its own inverse on every gate the translators can emit, and refuses what the backend never writes."""
import pytest


def _cd(p2g, workload="all_gates", npi=2, hasher="keccak25"):
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import oracle_cd
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
    gates = [g for g, _ in p2g.synth.gate_mix(workload, cfg)] + [p2g.circuit.Gate.noop()]
    if npi:
        gates += [p2g.circuit.Gate.public_input(), p2g.circuit.Gate.poseidon()]
    com = p2g.circuit.CommonCircuitData(cfg, 9, gates, npi)
    return com, oracle_cd(com)


@pytest.mark.parametrize("workload,hasher", [("all_gates", "keccak25"), ("ecdsa", "poseidon"), ("range", "keccak25")])
def test_round_trip(p2g, workload, hasher):
    from oracle.pyref import vk
    com, cd = _cd(p2g, workload, 2, hasher)
    hs = com.hash_size
    cap = [bytes([i]) * hs for i in range(16)]
    digest = b"\x7f" * hs
    raw = vk.serialize_verifier_data(cd, cap, digest)
    v = vk.parse_verifier_data(raw, hs)
    assert v["constants_sigmas_cap"] == cap and v["circuit_digest"] == digest and v["cap_height"] == 4
    assert [k for k, _ in v["gates"]] == [g.kind for g in com.gates]
    assert v["k_is"] == com.k_is and v["groups"] == [tuple(x) for x in com.groups]
    assert v["fri_params"]["reduction_arity_bits"] == list(com.reduction_arity_bits) and not v["fri_params"]["hiding"]
    # serialising the parsed fields again gives the same bytes for the other reduction strategies too
    for st in (("Fixed",), ("MinSize", None), ("MinSize", 5)):
        r2 = vk.serialize_verifier_data(cd, cap, digest, {"reduction_strategy": st})
        assert vk.parse_verifier_data(r2, hs)["config"]["fri_config"]["reduction_strategy"][0] == st[0]
    with pytest.raises(ValueError):
        vk.parse_verifier_data(raw[:-1], hs)
    with pytest.raises(ValueError):
        vk.parse_verifier_data(raw + b"\0", hs)


def test_gate_tags_follow_the_reference_list(p2g):
    """write_vk_action.rs:37-61: 22 gate types, tag = position."""
    from oracle.pyref import vk
    assert len(vk.TAGS) == 22 and vk.TAGS[0] == "ArithmeticGate" and vk.TAGS[17:] == [
        "ComparisonGate", "U32AddManyGate", "U32ArithmeticGate", "U32RangeCheckGate", "U32SubtractionGate"]
    assert vk._gate_tag_and_payload(vk.BASE_SUM, (4, 16, 0, 0)) == (3, [16])
    assert vk._gate_tag_and_payload(vk.U32_ADD_MANY, (3, 9, 0, 0)) == (18, [3, 9])       # add_many_u32.rs:94-97
    assert vk._gate_tag_and_payload(vk.COMPARISON, (32, 16, 0, 0)) == (17, [32, 16])     # comparison.rs:104-108
    with pytest.raises(ValueError):
        vk._gate_tag_and_payload(vk.BASE_SUM, (3, 5, 0, 0))
