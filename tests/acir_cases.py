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
