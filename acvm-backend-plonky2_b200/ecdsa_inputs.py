"""Inputs for the EcdsaSecp256k1 opcode: a tiny pure-Python secp256k1 (key generation, signing) and the ACIR circuit + witness
map a Noir program `std::ecdsa_secp256k1::verify_signature(pub_key_x, pub_key_y, signature, hashed_message)` compiles to
(160 byte witnesses, RANGE 8 on each, one BlackBoxFuncCall::EcdsaSecp256k1, `assert(valid)`), like the reference's
circuit_translation/tests/factories/noir_circuits_for_testing/ecdsa_secp256k1.

Byte order: the reference's translator (ecdsa_secp256k1_translator.rs:90-117) reads each 32-byte input as a LITTLE-endian
integer (byte 0 is the least significant byte of the least significant u32 limb).  `little_endian=True` encodes a valid
signature in that convention -- the only one under which the reference's circuit accepts a valid signature; Noir itself passes
big-endian bytes.  Host tooling: builds test / bench inputs, nothing here is on the proving path."""
import hashlib

P = 2 ** 256 - 2 ** 32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
     0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def point_add(p, q):
    """Complete affine addition on y^2 = x^3 + 7 (None = the point at infinity)."""
    if p is None:
        return q
    if q is None:
        return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % P == 0:
            return None
        lam = 3 * p[0] * p[0] * pow(2 * p[1], -1, P) % P
    else:
        lam = (q[1] - p[1]) * pow(q[0] - p[0], -1, P) % P
    x = (lam * lam - p[0] - q[0]) % P
    return x, (lam * (p[0] - x) - p[1]) % P


def point_mul(k, p):
    r = None
    while k:
        if k & 1:
            r = point_add(r, p)
        p = point_add(p, p)
        k >>= 1
    return r


def public_key(d):
    return point_mul(d % N, G)


def sign(d, h, k):
    """(r, s) for private key d, message hash h (an integer) and nonce k."""
    r = point_mul(k % N, G)[0] % N
    s = pow(k, -1, N) * (h + r * d) % N
    assert r and s
    return r, s


def recovered_x(q, r, s, h):
    """x of h/s G + r/s Q, or None at infinity: what the circuit compares r with."""
    s1 = pow(s, -1, N)
    pt = point_add(point_mul(h * s1 % N, G), point_mul(r * s1 % N, q))
    return None if pt is None else pt[0]


def deterministic_case(seed):
    """A key, a message hash and a valid signature derived from `seed`."""
    def H(tag):
        return int.from_bytes(hashlib.sha256(f"{tag}:{seed}".encode()).digest(), "big")
    d, k, h = H("key") % (N - 1) + 1, H("nonce") % (N - 1) + 1, H("message") % N
    q = public_key(d)
    r, s = sign(d, h, k)
    return q, r, s, h


def encode(v, little_endian=True):
    return list(v.to_bytes(32, "little" if little_endian else "big"))


def circuit_and_witness(acir, cases, outputs=None, range_checks=True, assert_valid=False, first_witness=0):
    """One EcdsaSecp256k1 opcode per case ((qx, qy), r, s, h as integers, or four 32/32/64/32 byte lists), all inputs private.
    Returns (Circuit, witness map without the outputs filled unless `outputs` gives them, list of output witness ids)."""
    ops, wit, out_ids, w = [], {}, [], first_witness
    for i, case in enumerate(cases):
        if isinstance(case[0], tuple):
            (qx, qy), r, s, h = case
            pkx, pky, sig, msg = encode(qx), encode(qy), encode(r) + encode(s), encode(h)
        else:
            pkx, pky, sig, msg = case
        ids = list(range(w, w + 160))
        for k, v in zip(ids, pkx + pky + sig + msg):
            wit[k] = v
        if range_checks:
            ops += [acir.Range(k, 8) for k in ids]
        out = w + 160
        ops.append(acir.EcdsaSecp256k1(ids[0:32], ids[32:64], ids[64:128], ids[128:160], out))
        if assert_valid:                       # assert(valid_signature): output - 1 = 0
            ops.append(acir.AssertZero(acir.Expression([], [(1, out)], 0xFFFFFFFF00000001 - 1)))
        if outputs is not None:
            wit[out] = outputs[i]
        out_ids.append(out)
        w += 161
    private = [k for k in range(first_witness, w) if k not in out_ids]
    return acir.Circuit(ops, [], private), wit, out_ids


class RealEcdsaCircuit:
    """A bench / test workload with SyntheticCircuit's attributes (common, constants_sigmas, wires, public_inputs, config,
    _wires_t when pinned) built from real opcodes: the Noir signature-check program above on 2^(degree_bits - 17) signatures
    (one EcdsaSecp256k1 call is 98.9 K rows = 2^17; eight fill 2^20).  Translation and witness generation run on the host."""

    ROWS_LOG2_PER_SIGNATURE = 17

    def __init__(self, degree_bits, acir, config=None, seed=0, pinned=False):
        import numpy as np
        if degree_bits < self.ROWS_LOG2_PER_SIGNATURE:
            raise ValueError("one EcdsaSecp256k1 circuit already has 2^17 rows")
        self.num_signatures = 1 << (degree_bits - self.ROWS_LOG2_PER_SIGNATURE)
        cases = [deterministic_case(f"{seed}/{i}") for i in range(self.num_signatures)]
        circuit, wit, _ = circuit_and_witness(acir, cases, outputs=[1] * self.num_signatures, assert_valid=True)
        tr = acir.CircuitBuilderFromAcirToPlonky2(config).translate_circuit(circuit)
        if tr.common.degree_bits() != degree_bits:
            raise RuntimeError(f"{self.num_signatures} signatures gave 2^{tr.common.degree_bits()} rows, not 2^{degree_bits}")
        self.config, self.common, self.constants_sigmas = tr.config, tr.common, tr.constants_sigmas
        self.workload = "ecdsa-real"
        self.rows_used = tr.rows_used()
        wires, self.public_inputs = tr.generate_witness(wit)
        if pinned:
            import torch
            self._wires_t = torch.empty(wires.shape, dtype=torch.int64).pin_memory()
            self.wires = self._wires_t.numpy().view(np.uint64)
            self.wires[:] = wires
        else:
            self.wires = wires
        tr.close()
