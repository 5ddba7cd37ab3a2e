"""ACIR circuits of the reference's translator tests, built by hand like circuit_translation/tests/factories/circuit_factory.rs
(no nargo): (name, circuit, witness assignment, expected public inputs)."""
P = 0xFFFFFFFF00000001


def cases(acir):
    E, AZ, C = acir.Expression, acir.AssertZero, acir.Circuit
    out = []
    # test_assert_zero.rs:6-273
    out.append(("x_equals_0", C([AZ(E([], [(1, 0)], 0))], [0]), {0: 0}, [0]))
    out.append(("x_equals_4", C([AZ(E([], [(1, 0)], -4))], [0]), {0: 4}, [4]))
    out.append(("x_times_3_equals_12", C([AZ(E([], [(3, 0)], -12))], [0]), {0: 4}, [4]))
    out.append(("3x_plus_9y_equals_12", C([AZ(E([], [(3, 0), (9, 1)], -12))], [0, 1]), {0: 1, 1: 1}, [1, 1]))
    out.append(("multiple_linear_combinations", C([AZ(E([], [(3, w) for w in reversed(range(4))], -12))], [0, 1, 2, 3]),
                {0: 1, 1: 1, 2: 1, 3: 1}, [1, 1, 1, 1]))
    out.append(("two_x_x_equals_32", C([AZ(E([(2, 0, 0)], [], -32))], [0]), {0: 4}, [4]))
    out.append(("two_x_y_equals_40", C([AZ(E([(2, 0, 1)], [], -40))], [0, 1]), {0: 5, 1: 4}, [5, 4]))
    out.append(("multiple_cuadratic_terms", C([AZ(E([(2, 0, 0), (3, 0, 1), (4, 1, 1)], [], -(2 * 4 + 3 * 6 + 4 * 9)))], [0, 1]),
                {0: 2, 1: 3}, [2, 3]))
    out.append(("cuadratic_and_linear", C([AZ(E([(2, 0, 0), (3, 0, 1)], [(5, 0), (7, 1)], -(8 + 18 + 10 + 21)))], [0, 1]),
                {0: 2, 1: 3}, [2, 3]))
    out.append(("two_assert_zero_opcodes", C([AZ(E([], [(1, 0)], -4)), AZ(E([], [(3, 1)], -12))], [0, 1]), {0: 4, 1: 4}, [4, 4]))
    # private inputs, an intermediate witness shared by two opcodes (w2 = x * y ; w2 + x = 26)
    out.append(("private_and_intermediate", C([AZ(E([(1, 0, 1)], [(P - 1, 2)], 0)), AZ(E([], [(1, 2), (1, 0)], -25))], [0], [1]),
                {0: 5, 1: 4, 2: 20}, [5]))
    # test_blackbox.rs:8-82 (RANGE), :112-218 (AND / XOR; the output witness is computed by the generators, not provided)
    for bits, v in ((8, 255), (16, 65535), (32, (1 << 32) - 1), (33, (1 << 33) - 1)):
        out.append((f"range_u{bits}", C([acir.Range(0, bits)], [0]), {0: v}, [v]))
    for bits, a, b in ((8, 0b10101010, 0b11001100), (16, 0xBEEF, 0x0FF0), (32, 0xDEADBEEF, 0x12345678)):
        out.append((f"and_{bits}", C([acir.And(0, 1, bits, 2)], [0, 1]), {0: a, 1: b}, [a, b]))
        out.append((f"xor_{bits}", C([acir.Xor(0, 1, bits, 2)], [0, 1]), {0: a, 1: b}, [a, b]))
    # AND output fed to an AssertZero: out - expected = 0
    out.append(("and_output_is_constrained", C([acir.And(0, 1, 8, 2), AZ(E([], [(1, 2)], -(0b10101010 & 0b11001100)))], [0, 1]),
                {0: 0b10101010, 1: 0b11001100}, [0b10101010, 0b11001100]))
    # test_memory_operations.rs:10-37 (read), :82-121 (irregular block size)
    out.append(("memory_read", C([acir.MemoryInit(0, [0, 1, 2, 3]), acir.MemoryRead(0, 4, 5)], [0, 1, 2, 3, 4]),
                {0: 10, 1: 11, 2: 12, 3: 13, 4: 2}, [10, 11, 12, 13, 2]))
    out.append(("memory_read_irregular_block", C([acir.MemoryInit(0, [0, 1, 2]), acir.MemoryRead(0, 3, 4),
                                                  AZ(E([], [(1, 4)], -21))], [0, 1, 2, 3]),
                {0: 20, 1: 21, 2: 22, 3: 1}, [20, 21, 22, 1]))
    # test_memory_operations.rs:39-79: x[y] = v; assert(x[0] == 1); assert(x[1] == 11)
    out.append(("memory_write", C([acir.MemoryInit(0, [0, 1]), acir.MemoryWrite(0, 2, 3), acir.MemoryRead(0, 4, 6),
                                   acir.MemoryRead(0, 5, 7), AZ(E([], [(1, 6)], -1)), AZ(E([], [(1, 7)], -11))], [0, 1, 2, 3]),
                {0: 10, 1: 11, 2: 0, 3: 1, 4: 0, 5: 1, 6: 1, 7: 11}, [10, 11, 0, 1]))
    # a write into a block of irregular size, read back through a computed index
    out.append(("memory_write_irregular", C([acir.MemoryInit(1, [0, 1, 2]), acir.MemoryWrite(1, 3, 4), acir.MemoryRead(1, 3, 5),
                                             AZ(E([], [(1, 5), (P - 1, 4)], 0))], [0, 1, 2, 3, 4]),
                {0: 7, 1: 8, 2: 9, 3: 2, 4: 55}, [7, 8, 9, 2, 55]))
    return out


def chain(acir, n_ops, seed=1):
    """A longer AssertZero chain (BASELINE configs[1] shape from real opcodes): w_{i+1} = w_i * w_i + 3 w_i + i."""
    ops, wit, x = [], {0: seed}, seed
    for i in range(n_ops):
        y = (x * x + 3 * x + i) % P
        ops.append(acir.AssertZero(acir.Expression([(1, i, i)], [(3, i), (P - 1, i + 1)], i)))
        wit[i + 1] = y
        x = y
    return acir.Circuit(ops, [0]), wit


SHA256_IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]


def sha256_circuit(acir, message):
    """SHA-256 of `message` as chained Sha256Compression opcodes (one per 64-byte block of the padded message), the way the
    reference's sha256 tests drive the translator (circuit_translation/tests/test_sha256_internal.rs:481-549).  Public parameters:
    the message words and the IV; the final state words are the last 8 witnesses.  Returns (circuit, witness map without outputs,
    output witness ids, expected digest words)."""
    import hashlib
    import struct
    padded = message + b"\x80" + b"\0" * ((55 - len(message)) % 64) + struct.pack(">Q", 8 * len(message))
    nblocks = len(padded) // 64
    words = list(struct.unpack(f">{16 * nblocks}I", padded))
    wit = {i: w for i, w in enumerate(words)}
    iv_ids = list(range(16 * nblocks, 16 * nblocks + 8))
    wit.update({iv_ids[i]: SHA256_IV[i] for i in range(8)})
    nxt = iv_ids[-1] + 1
    ops, state = [], iv_ids
    for blk in range(nblocks):
        outs = list(range(nxt, nxt + 8))
        nxt += 8
        ops.append(acir.Sha256Compression(list(range(16 * blk, 16 * blk + 16)), state, outs))
        state = outs
    circuit = acir.Circuit(ops, list(range(16 * nblocks + 8)))
    return circuit, wit, state, list(struct.unpack(">8I", hashlib.sha256(message).digest()))


def u32_gadget_cases(acir):
    """Circuits built directly on the reference's u32 gadgets (plonky2_ecdsa/biguint/gadgets/arithmetic_u32.rs, range_check.rs,
    multiple_comparison.rs) — the operations its EcdsaSecp256k1 translator is made of; values like the gates' own tests
    (arithmetic_u32.rs:456-, subtraction_u32.rs:380-, comparison.rs:600-): (name, circuit, witness incl. expected outputs)."""
    import random
    E, AZ, C = acir.Expression, acir.AssertZero, acir.Circuit
    M = (1 << 32) - 1
    rng = random.Random(32)
    out = []
    # x*y+z: the all-ones corner is p-1 (high half = u32::MAX, low half must be 0: the gate's canonicity check)
    trip = [(M, M, M), (0, 0, 0), (M, M, 0), (1, M, 1)] + [tuple(rng.getrandbits(32) for _ in range(3)) for _ in range(5)]
    ops, wit = [], {}
    for k, (x, y, z) in enumerate(trip):          # 9 operations: more than the 6 one U32ArithmeticGate row holds
        b = 5 * k
        v = x * y + z
        ops.append(acir.MulAddU32(b, b + 1, b + 2, b + 3, b + 4))
        wit.update({b: x, b + 1: y, b + 2: z, b + 3: v & M, b + 4: v >> 32})
    out.append(("mul_add_u32", C(ops, [0, 1, 2], list(range(5, 5 * len(trip)))), wit))
    ops, wit, w = [], {}, 0
    for na in (2, 3, 4, 5, 16, 3, 3, 3, 3, 3, 3):   # 2 -> U32ArithmeticGate; 3 five times -> a second U32AddManyGate(3) row
        vals = [M] * na if na in (4, 16) else [rng.getrandbits(32) for _ in range(na)]
        s = sum(vals)
        ids = list(range(w, w + na))
        ops.append(acir.AddManyU32(ids, w + na, w + na + 1))
        wit.update({**dict(zip(ids, vals)), w + na: s & M, w + na + 1: s >> 32})
        w += na + 2
    out.append(("add_many_u32", C(ops, [0, 1], list(range(2, w))), wit))
    ops, wit = [], {}
    subs = [(5, 3, 0), (3, 5, 0), (0, 0, 1), (M, M, 1), (0, M, 1), (M, 0, 0), (7, 7, 0)] + \
           [(rng.getrandbits(32), rng.getrandbits(32), rng.getrandbits(1)) for _ in range(12)]   # 19 > 11 operations per row
    for k, (x, y, bw) in enumerate(subs):
        b = 5 * k
        d = x - y - bw
        ops.append(acir.SubU32(b, b + 1, b + 2, b + 3, b + 4))
        wit.update({b: x, b + 1: y, b + 2: bw, b + 3: d % (1 << 32), b + 4: 1 if d < 0 else 0})
    out.append(("sub_u32", C(ops, [0, 1, 2], list(range(5, 5 * len(subs)))), wit))
    vals = [0, 1, M, 1 << 31] + [rng.getrandbits(32) for _ in range(9)]
    out.append(("range_check_u32", C([acir.RangeCheckU32(list(range(13))), acir.RangeCheckU32([3])], [0], list(range(1, 13))),
                dict(enumerate(vals))))
    ops, wit = [], {}
    cmps = [(3, 5, 32), (5, 3, 32), (5, 5, 32), (0, M, 32), (M, 0, 32), (M, M, 32), (0, 0, 32), (513, 512, 10), (512, 513, 10),
            (1, 0, 1), (0, 1, 1), (0x12345678, 0x12355678, 32), (6, 5, 3)]
    for k, (a, b_, bits) in enumerate(cmps):
        b = 3 * k
        ops.append(acir.CmpLe(b, b + 1, bits, b + 2))
        wit.update({b: a, b + 1: b_, b + 2: 1 if a <= b_ else 0})
    out.append(("cmp_le", C(ops, [0, 1], list(range(3, 3 * len(cmps)))), wit))
    # a 96-bit add with carry chain then a comparison of the top limbs: gadgets feeding each other and an AssertZero
    a = [rng.getrandbits(32) for _ in range(3)]
    b_ = [rng.getrandbits(32) for _ in range(3)]
    ops, wit, carry = [], {i: a[i] for i in range(3)}, 0
    wit.update({3 + i: b_[i] for i in range(3)})
    wit[6] = 0                                   # carry in
    ops.append(AZ(E([], [(1, 6)], 0)))
    cid = 6
    for i in range(3):
        s = a[i] + b_[i] + carry
        ops.append(acir.AddManyU32([i, 3 + i, cid], 7 + 2 * i, 8 + 2 * i))
        wit.update({7 + 2 * i: s & M, 8 + 2 * i: s >> 32})
        carry, cid = s >> 32, 8 + 2 * i
    ops.append(acir.RangeCheckU32([7, 9, 11]))
    ops.append(acir.CmpLe(7, 11, 32, 13))
    wit[13] = 1 if wit[7] <= wit[11] else 0
    ops.append(AZ(E([], [(1, 12), (P - 1, 14)], 0)))   # w14 = final carry
    wit[14] = carry
    out.append(("biguint_add_96", C(ops, [0, 1, 2, 3, 4, 5], [6]), wit))
    return out
