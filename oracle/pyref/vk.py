"""ORACLE = test infrastructure.  Verification-key codec: plonky2 0.2.2 `VerifierCircuitData::{to_bytes, from_bytes}` with the
reference's `BackendGateSerializer` (plonky2-backend/src/actions/write_vk_action.rs:35-62; read back by
noir_and_plonky2_serialization.rs:16-22 `VerifierCircuitData::from_bytes(vk, &BackendGateSerializer)`).

Layout restated from util/serialization.rs of the un-vendored plonky2 crate (write_verifier_only_circuit_data,
write_common_circuit_data, write_circuit_config, write_fri_config, write_fri_params, write_selectors_info, write_gate); the payload
of the five custom gates is read from the reference itself (add_many_u32.rs:94-97, arithmetic_u32.rs:93-95, comparison.rs:104-108,
range_check_u32.rs:59-61, subtraction_u32.rs `serialize`).  PARITY UNPINNED: the reference commits no VK file, so nothing here is
checked against bytes the Rust code produced -- only writer == parser^-1 and library == this writer.
"""
import struct

# gate tag = position in the impl_gate_serializer! list of write_vk_action.rs:37-61
TAGS = ["ArithmeticGate", "ArithmeticExtensionGate", "BaseSumGate<2>", "BaseSumGate<4>", "ConstantGate", "CosetInterpolationGate",
        "ExponentiationGate", "LookupGate", "LookupTableGate", "MulExtensionGate", "NoopGate", "PoseidonMdsGate", "PoseidonGate",
        "PublicInputGate", "RandomAccessGate", "ReducingExtensionGate", "ReducingGate", "ComparisonGate", "U32AddManyGate",
        "U32ArithmeticGate", "U32RangeCheckGate", "U32SubtractionGate"]
# gate kinds of include/p2g.h
(NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, BASE_SUM, POSEIDON, RANDOM_ACCESS, U32_ARITHMETIC, U32_ADD_MANY, U32_SUBTRACTION,
 U32_RANGE_CHECK, COMPARISON) = range(12)

DEFAULT_CONFIG = dict(config_num_constants=2, security_bits=100, max_quotient_degree_factor=8, use_base_arithmetic_gate=True,
                      zero_knowledge=False, reduction_strategy=("ConstantArityBits", 4, 5))


def _gate_tag_and_payload(kind, params):
    p = list(params)
    if kind == ARITHMETIC:
        return 0, [p[0]]
    if kind == BASE_SUM:
        if p[0] not in (2, 4):
            raise ValueError("BackendGateSerializer registers BaseSumGate<2> and BaseSumGate<4> only")
        return (2 if p[0] == 2 else 3), [p[1]]
    if kind == CONSTANT:
        return 4, [p[0]]
    if kind == NOOP:
        return 10, []
    if kind == POSEIDON:
        return 12, []
    if kind == PUBLIC_INPUT:
        return 13, []
    if kind == RANDOM_ACCESS:
        return 14, [p[0], p[1], p[2]]
    if kind == COMPARISON:
        return 17, [p[0], p[1]]
    if kind == U32_ADD_MANY:
        return 18, [p[0], p[1]]
    if kind == U32_ARITHMETIC:
        return 19, [p[0]]
    if kind == U32_RANGE_CHECK:
        return 20, [p[0]]
    if kind == U32_SUBTRACTION:
        return 21, [p[0]]
    raise ValueError(f"unknown gate kind {kind}")


_PAYLOAD_LEN = {0: 1, 2: 1, 3: 1, 4: 1, 10: 0, 12: 0, 13: 0, 14: 3, 17: 2, 18: 2, 19: 1, 20: 1, 21: 1}
_TAG_KIND = {0: ARITHMETIC, 2: BASE_SUM, 3: BASE_SUM, 4: CONSTANT, 10: NOOP, 12: POSEIDON, 13: PUBLIC_INPUT, 14: RANDOM_ACCESS,
             17: COMPARISON, 18: U32_ADD_MANY, 19: U32_ARITHMETIC, 20: U32_RANGE_CHECK, 21: U32_SUBTRACTION}


def serialize_verifier_data(cd, cap, digest, config=None):
    """cd: oracle.pyref.circuit.CommonData; cap: list of digests (bytes); digest: circuit_digest bytes."""
    k = dict(DEFAULT_CONFIG, **(config or {}))
    out = bytearray()
    usz = lambda x: out.extend(struct.pack("<Q", x))
    u8 = lambda x: out.append(1 if x is True else 0 if x is False else x)

    def fri_config():
        usz(cd.rate_bits)
        usz(cd.cap_height)
        usz(cd.num_queries)
        out.extend(struct.pack("<I", cd.pow_bits))
        st = k["reduction_strategy"]
        if st[0] == "Fixed":
            u8(0)
            usz(len(cd.arity_bits))
            for a in cd.arity_bits:
                usz(a)
        elif st[0] == "ConstantArityBits":
            u8(1)
            usz(st[1])
            usz(st[2])
        else:
            u8(2)
            u8(st[1] is not None)
            if st[1] is not None:
                usz(st[1])
    usz(cd.cap_height)
    for d in cap:
        out.extend(d)
    out.extend(digest)
    usz(cd.num_wires)
    usz(cd.num_routed)
    usz(k["config_num_constants"])
    usz(k["security_bits"])
    usz(cd.num_challenges)
    usz(k["max_quotient_degree_factor"])
    u8(bool(k["use_base_arithmetic_gate"]))
    u8(bool(k["zero_knowledge"]))
    fri_config()
    fri_config()
    usz(len(cd.arity_bits))
    for a in cd.arity_bits:
        usz(a)
    usz(cd.degree_bits)
    u8(bool(k["zero_knowledge"]))
    usz(len(cd.selector_indices))
    for s in cd.selector_indices:
        usz(s)
    usz(len(cd.groups))
    for lo, hi in cd.groups:
        usz(lo)
        usz(hi)
    usz(cd.qdf)
    usz(cd.num_gate_constraints)
    usz(cd.num_constants)
    usz(cd.num_public_inputs)
    usz(len(cd.k_is))
    for x in cd.k_is:
        usz(x)
    usz(cd.num_partial_products)
    usz(0)
    usz(0)
    usz(0)
    usz(len(cd.gates))
    for g in cd.gates:
        tag, payload = _gate_tag_and_payload(g.kind, g.params)
        out.extend(struct.pack("<I", tag))
        for x in payload:
            usz(x)
    return bytes(out)


def parse_verifier_data(raw, hash_size):
    """Inverse of serialize_verifier_data: a dict of every field (what VerifierCircuitData::from_bytes reads)."""
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(raw):
            raise ValueError("truncated verifier data")
        b = raw[pos:pos + n]
        pos += n
        return b
    usz = lambda: struct.unpack("<Q", take(8))[0]
    u8 = lambda: take(1)[0]

    def fri_config():
        c = dict(rate_bits=usz(), cap_height=usz(), num_query_rounds=usz(), proof_of_work_bits=struct.unpack("<I", take(4))[0])
        t = u8()
        if t == 0:
            c["reduction_strategy"] = ("Fixed", [usz() for _ in range(usz())])
        elif t == 1:
            c["reduction_strategy"] = ("ConstantArityBits", usz(), usz())
        elif t == 2:
            c["reduction_strategy"] = ("MinSize", usz() if u8() else None)
        else:
            raise ValueError("bad FriReductionStrategy tag")
        return c
    v = {}
    v["cap_height"] = usz()
    v["constants_sigmas_cap"] = [bytes(take(hash_size)) for _ in range(1 << v["cap_height"])]
    v["circuit_digest"] = bytes(take(hash_size))
    cfg = dict(num_wires=usz(), num_routed_wires=usz(), num_constants=usz(), security_bits=usz(), num_challenges=usz(),
               max_quotient_degree_factor=usz(), use_base_arithmetic_gate=bool(u8()), zero_knowledge=bool(u8()))
    cfg["fri_config"] = fri_config()
    v["config"] = cfg
    v["fri_params"] = dict(config=fri_config(), reduction_arity_bits=[usz() for _ in range(usz())], degree_bits=usz(), hiding=bool(u8()))
    v["selector_indices"] = [usz() for _ in range(usz())]
    v["groups"] = [(usz(), usz()) for _ in range(usz())]
    for key in ("quotient_degree_factor", "num_gate_constraints", "num_constants", "num_public_inputs"):
        v[key] = usz()
    v["k_is"] = [usz() for _ in range(usz())]
    v["num_partial_products"] = usz()
    v["num_lookup_polys"] = usz()
    v["num_lookup_selectors"] = usz()
    if usz() != 0:
        raise ValueError("lookup tables are not produced by this backend")
    gates = []
    for _ in range(usz()):
        tag = struct.unpack("<I", take(4))[0]
        if tag not in _PAYLOAD_LEN:
            raise ValueError(f"gate tag {tag} ({TAGS[tag] if tag < len(TAGS) else '?'}) is not reachable from the translators")
        payload = [usz() for _ in range(_PAYLOAD_LEN[tag])]
        kind = _TAG_KIND[tag]
        if kind == BASE_SUM:
            params = [2 if tag == 2 else 4, payload[0]]
        else:
            params = payload
        gates.append((kind, tuple(params + [0] * (4 - len(params)))))
    v["gates"] = gates
    if pos != len(raw):
        raise ValueError("trailing bytes after the verifier data")
    return v
