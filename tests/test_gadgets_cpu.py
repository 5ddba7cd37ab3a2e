"""CPU: the reference's gadget tests (plonky2_ecdsa/biguint/biguint.rs:374-520 test_biguint_{add,sub,mul,cmp,div_rem},
gadgets/nonnative.rs:739-909 test_nonnative_{add,many_adds,sub,mul,neg,inv}, gadgets/multiple_comparison.rs:93-150 test_list_le,
gadgets/arithmetic_u32.rs:362-394 test_add_many_u32s) on the C++ restatement of those gadgets and of their witness generators:
the generated outputs equal integer arithmetic, the trace satisfies every gate and copy constraint, the oracle proves and verifies
(`data.prove(pw)` / `data.verify(proof)` in the reference), and a wrong expected value is refused."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.dirname(__file__))
from test_acir_cpu import _check_trace  # noqa: E402

SECP_P = 2 ** 256 - 2 ** 32 - 977
SECP_N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
M32 = (1 << 32) - 1


def limbs(v, n):
    assert v < 1 << (32 * n)
    return [(v >> (32 * i)) & M32 for i in range(n)]


def nlimbs(v):
    return max(1, (v.bit_length() + 31) // 32)


class Program:
    """Witness-id allocator around a list of Gadget opcodes."""

    def __init__(self, acir):
        self.A, self.ops, self.wit, self.next = acir, [], {}, 0
        self.expect = {}

    def value(self, v, n):                       # an input big integer: n fresh witnesses holding its limbs
        ids = list(range(self.next, self.next + n))
        self.next += n
        self.wit.update(dict(zip(ids, limbs(v, n))))
        return ids

    def result(self, v, n):                      # an output: fresh witnesses whose expected limbs are remembered, not provided
        ids = list(range(self.next, self.next + n))
        self.next += n
        self.expect.update(dict(zip(ids, limbs(v, n))))
        return ids

    def add(self, name, lists, param=0):
        self.ops.append(self.A.Gadget(name, lists, param))

    def circuit(self):
        inputs = sorted(self.wit)
        return self.A.Circuit(self.ops, inputs[:1], inputs[1:])


def run(p2g, corc, prog, prove=True):
    A = p2g.acir
    tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(prog.circuit())
    wires, pis = tr.generate_witness(prog.wit)
    got = tr.read_witnesses(prog.expect)
    assert got == prog.expect
    cd, _ = _check_trace(p2g, corc, tr, wires, pis)
    if prove:
        from oracle.pyref import proof, verifier
        op = corc.OracleProver(cd, tr.constants_sigmas)
        pb = op.prove(wires, pis)
        cap, dg = op.cap_and_digest()
        verifier.verify(proof.parse_uncompressed(pb, cd), cd, cap, dg)
    wrong = dict(prog.wit)                       # providing an output that contradicts the generators is refused
    k = max(prog.expect)
    wrong[k] = prog.expect[k] ^ 1
    with pytest.raises(A.TranslationError):
        tr.generate_witness(wrong)
    return tr


def test_biguint_gadgets(p2g, corc):
    rng = random.Random(374)
    P = Program(p2g.acir)
    for _ in range(2):                           # test_biguint_add / _sub / _mul / _cmp / _div_rem on random 128-bit values
        x, y = rng.getrandbits(128), rng.getrandbits(128)
        if y > x:
            x, y = y, x
        P.add("add_biguint", [P.value(x, 4), P.value(y, 4), P.result(x + y, 5)])
        P.add("sub_biguint", [P.value(x, 4), P.value(y, 4), P.result(x - y, 4)])
        P.add("mul_biguint", [P.value(x, 4), P.value(y, 4), P.result(x * y, 9)])
        P.add("cmp_biguint", [P.value(x, 4), P.value(y, 4), P.result(int(x <= y), 1)])
        P.add("cmp_biguint", [P.value(y, 4), P.value(x, 4), P.result(1, 1)])
        y2 = rng.getrandbits(70) + 1
        P.add("div_rem_biguint", [P.value(x, 4), P.value(y2, 3), P.result(x // y2, 2), P.result(x % y2, 3)])
    # operands of different lengths, equality, carries through every limb, a zero remainder
    P.add("add_biguint", [P.value((1 << 160) - 1, 5), P.value(1, 1), P.result(1 << 160, 6)])
    P.add("cmp_biguint", [P.value(7, 1), P.value(7 + (1 << 64), 3), P.result(1, 1)])
    P.add("cmp_biguint", [P.value(12345 << 40, 3), P.value(12345 << 40, 3), P.result(1, 1)])
    P.add("mul_biguint", [P.value((1 << 96) - 1, 3), P.value(M32, 1), P.result(((1 << 96) - 1) * M32, 5)])
    P.add("div_rem_biguint", [P.value(35 << 64, 3), P.value(5 << 32, 2), P.result((35 << 64) // (5 << 32), 2), P.result(0, 2)])
    run(p2g, corc, P)


@pytest.mark.parametrize("field", [0, 1])
def test_nonnative_gadgets(p2g, corc, field):
    mod = SECP_N if field else SECP_P
    rng = random.Random(739 + field)
    P = Program(p2g.acir)
    x, y = rng.randrange(mod), rng.randrange(mod)
    P.add("add_nonnative", [P.value(x, 8), P.value(y, 8), P.result((x + y) % mod, 8)], field)           # test_nonnative_add
    P.add("sub_nonnative", [P.value(x, 8), P.value(y, 8), P.result((x - y) % mod, 8)], field)           # test_nonnative_sub
    P.add("sub_nonnative", [P.value(y, 8), P.value(x, 8), P.result((y - x) % mod, 8)], field)
    P.add("mul_nonnative", [P.value(x, 8), P.value(y, 8), P.result(x * y % mod, 8)], field)             # test_nonnative_mul
    P.add("neg_nonnative", [P.value(x, 8), P.result((mod - x) % mod, 8)], field)                        # test_nonnative_neg
    P.add("inv_nonnative", [P.value(x, 8), P.result(pow(x, -1, mod), 8)], field)                        # test_nonnative_inv
    summands = [rng.randrange(mod) for _ in range(8)]                                                   # test_nonnative_many_adds
    P.add("add_many_nonnative", [P.value(s, 8) for s in summands] + [P.result(sum(summands) % mod, 8)], field)
    # corners: a sum that lands exactly on the modulus stays unreduced (the generator's comparison is strict, the circuit's is <=);
    # x - x; -0; (p-1)^2; the inverse of 1 and of p-1
    P.add("add_nonnative", [P.value(mod - 5, 8), P.value(5, 8), P.result(mod, 8)], field)
    P.add("add_nonnative", [P.value(mod - 5, 8), P.value(6, 8), P.result(1, 8)], field)
    P.add("sub_nonnative", [P.value(x, 8), P.value(x, 8), P.result(0, 8)], field)
    P.add("neg_nonnative", [P.value(0, 8), P.result(0, 8)], field)
    P.add("mul_nonnative", [P.value(mod - 1, 8), P.value(mod - 1, 8), P.result(1, 8)], field)
    P.add("inv_nonnative", [P.value(1, 8), P.result(1, 8)], field)
    P.add("inv_nonnative", [P.value(mod - 1, 8), P.result(mod - 1, 8)], field)
    run(p2g, corc, P)


def test_list_le(p2g, corc):
    """multiple_comparison.rs test_list_le(size, num_bits): lists are little-endian digits of one big number."""
    rng = random.Random(93)
    P = Program(p2g.acir)
    for size, bits in ((1, 1), (3, 1), (1, 10), (4, 10), (8, 32), (5, 32)):
        a = [rng.getrandbits(bits) for _ in range(size)]
        b = [rng.getrandbits(bits) for _ in range(size)]
        val = lambda l: sum(d << (bits * i) for i, d in enumerate(l))   # noqa: E731
        for p, q in ((a, b), (b, a), (a, a)):
            ia = [P.value(d, 1)[0] for d in p]
            ib = [P.value(d, 1)[0] for d in q]
            P.add("list_le", [ia, ib, P.result(int(val(p) <= val(q)), 1)], bits)
    run(p2g, corc, P)


def test_curve_gadgets_on_curve_points(p2g, corc):
    """curve.rs / glv.rs (their own tests are commented out or need the curve crate): addition, doubling and the GLV
    multiplication of points ON the curve against textbook secp256k1 arithmetic."""
    EI = p2g.ecdsa_inputs
    P = Program(p2g.acir)
    p1, p2 = EI.point_mul(0xC0FFEE, EI.G), EI.point_mul(0xDECAF, EI.G)
    s = EI.point_add(p1, p2)
    P.add("curve_add", [P.value(p1[0], 8), P.value(p1[1], 8), P.value(p2[0], 8), P.value(p2[1], 8), P.result(s[0], 8), P.result(s[1], 8)])
    d = EI.point_add(p1, p1)
    P.add("curve_double", [P.value(p1[0], 8), P.value(p1[1], 8), P.result(d[0], 8), P.result(d[1], 8)])
    k = random.Random(120).randrange(SECP_N)
    m = EI.point_mul(k, p2)
    P.add("glv_mul", [P.value(p2[0], 8), P.value(p2[1], 8), P.value(k, 8), P.result(m[0], 8), P.result(m[1], 8)])
    run(p2g, corc, P, prove=False)     # 2^16 rows: constraints and outputs are checked, the proof is left to the ECDSA tests


@pytest.mark.parametrize("seed", range(6))
def test_random_gadget_programs(p2g, corc, seed):
    """Randomised programs over the biguint / non-native gadgets with operands of random limb counts (1..8): outputs against Python
    integers, every gate and copy constraint on the trace (no proof: the circuits differ only in size from the ones proved above)."""
    rng = random.Random(1000 + seed)
    P = Program(p2g.acir)

    def rnd(nl):   # a value that really uses nl limbs about half of the time
        v = rng.getrandbits(32 * nl)
        return v if rng.random() < 0.5 else v >> rng.randrange(0, 32 * nl)

    for _ in range(14):
        kind = rng.choice(["add", "sub", "mul", "cmp", "divrem", "nn_add", "nn_sub", "nn_mul", "nn_inv"])
        na, nb = rng.randrange(1, 9), rng.randrange(1, 9)
        x, y = rnd(na), rnd(nb)
        if kind == "add":
            P.add("add_biguint", [P.value(x, na), P.value(y, nb), P.result(x + y, max(na, nb) + 1)])
        elif kind == "sub":
            if x < y:
                x, y, na, nb = y, x, nb, na
            P.add("sub_biguint", [P.value(x, na), P.value(y, nb), P.result(x - y, max(na, nb))])
        elif kind == "mul":
            P.add("mul_biguint", [P.value(x, na), P.value(y, nb), P.result(x * y, na + nb + 1)])
        elif kind == "cmp":
            if rng.random() < 0.3:
                y, nb = x, na
            P.add("cmp_biguint", [P.value(x, na), P.value(y, nb), P.result(int(x <= y), 1)])
        elif kind == "divrem":
            y = y or 1
            if nb > na + 1:
                nb = na
                y = (y % (1 << (32 * nb))) or 1
            P.add("div_rem_biguint", [P.value(x, na), P.value(y, nb), P.result(x // y, na - nb + 1 if nb <= na + 1 else 0),
                                      P.result(x % y, nb)]) if x // y < 1 << (32 * max(0, na - nb + 1)) else None
        else:
            field = rng.randrange(2)
            mod = SECP_N if field else SECP_P
            x, y = rng.randrange(mod), rng.randrange(mod)
            if kind == "nn_add":
                P.add("add_nonnative", [P.value(x, 8), P.value(y, 8), P.result((x + y) % mod if x + y != mod else mod, 8)], field)
            elif kind == "nn_sub":
                P.add("sub_nonnative", [P.value(x, 8), P.value(y, 8), P.result((x - y) % mod, 8)], field)
            elif kind == "nn_mul":
                P.add("mul_nonnative", [P.value(x, 8), P.value(y, 8), P.result(x * y % mod, 8)], field)
            else:
                x = x or 1
                P.add("inv_nonnative", [P.value(x, 8), P.result(pow(x, -1, mod), 8)], field)
    run(p2g, corc, P, prove=False)
