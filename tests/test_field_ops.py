"""csrc/gl.cuh against Python big integers on adversarial operands (the carry/borrow corner cases have probability ~2^-32 on
random inputs, so they are enumerated): host twins on CPU, the sm_100a PTX forms on the GPU."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001
EPS = 0xFFFFFFFF
M64 = (1 << 64) - 1

EDGE = sorted({x & M64 for x in [0, 1, 2, 3, 4, EPS - 1, EPS, EPS + 1, EPS + 2, 1 << 32, (1 << 33) - 1, 1 << 63, (1 << 63) - 1,
                                 P - 2, P - 1, P, P + 1, P + 2, P + EPS - 1, P + EPS, M64 - 1, M64, M64 - EPS, M64 - EPS - 1,
                                 M64 - EPS + 1, 0xFFFFFFFE00000000, 0xFFFFFFFE00000001, 0xFFFFFFFEFFFFFFFF, 0xFFFFFFFF00000000,
                                 0x00000001FFFFFFFF, 0x0000000100000000, 0x00000000FFFFFFFE, 0x8000000000000001,
                                 0x7FFFFFFF80000000, 0xFFFFFFFF7FFFFFFF, 0x00000001_00000001, 0xFFFFFFFD_FFFFFFFF]})


def operands(seed, canonical_b, nrand=20000):
    rng = np.random.default_rng(seed)
    a = [x for x in EDGE for _ in EDGE]
    b = [y for _ in EDGE for y in EDGE]
    ra = rng.integers(0, 1 << 64, nrand, dtype=np.uint64).tolist()
    rb = rng.integers(0, 1 << 64, nrand, dtype=np.uint64).tolist()
    # random values hugging the interesting boundaries
    for base in (0, EPS, 1 << 32, P, M64):
        d = rng.integers(0, 1 << 33, 2000, dtype=np.uint64).tolist()
        ra += [(base + x) & M64 for x in d[:1000]] + [(base - x) & M64 for x in d[1000:]]
        rb += rng.integers(0, 1 << 64, 2000, dtype=np.uint64).tolist()
    a += ra
    b += rb
    a, b = np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64)
    if canonical_b == "le_p":
        b = np.where(b > np.uint64(P), b - np.uint64(P), b)
    elif canonical_b == "lt_p":
        b = np.where(b >= np.uint64(P), b - np.uint64(P), b)
    elif canonical_b == "u32":
        b = b & np.uint64(EPS)
    pad = (-len(a)) % 8
    if pad:
        a, b = np.concatenate([a, a[:pad]]), np.concatenate([b, b[:pad]])
    return a, b


CASES = [("sub", "le_p", lambda a, b: (a - b) % P), ("add", "lt_p", lambda a, b: (a + b) % P),
         ("mul", None, lambda a, b: a * b % P), ("mulz", None, lambda a, b: a * b % P),
         ("mul_small", "u32", lambda a, b: a * b % P), ("canon", None, lambda a, b: a % P),
         ("reduce128", None, lambda a, b: ((a << 64) + b) % P),
         ("mulf", None, lambda a, b: a * b % P), ("canonf", None, lambda a, b: a % P)]


def run(p2g, device):
    for i, (op, bclass, ref) in enumerate(CASES):
        a, b = operands(100 + i, bclass)
        got = p2g.lib.field_ops(op, a, b, device=device)
        want = np.array([ref(int(x), int(y)) for x, y in zip(a.tolist(), b.tolist())], dtype=np.uint64)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (op, hex(int(a[bad[0]])), hex(int(b[bad[0]])), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
    # glf_add takes two canonical operands
    a, b = operands(55, "lt_p")
    a = np.where(a >= np.uint64(P), a - np.uint64(P), a)
    got = p2g.lib.field_ops("addf", a, b, device=device)
    want = np.array([(int(x) + int(y)) % P for x, y in zip(a.tolist(), b.tolist())], dtype=np.uint64)
    assert np.array_equal(got, want)
    # dot products with one final reduction: groups of 8 arbitrary u64 pairs, incl. all-ones (largest carries)
    a, b = operands(7, None)
    a[:64], b[:64] = np.uint64(M64), np.uint64(M64)
    got = p2g.lib.field_ops("dot8", a, b, device=device)
    al, bl = a.tolist(), b.tolist()
    want = np.array([sum(al[8 * i + k] * bl[8 * i + k] for k in range(8)) % P for i in range(len(al) // 8)], dtype=np.uint64)
    assert np.array_equal(got, want)
    b32 = b & np.uint64(EPS)
    b32[:64] = np.uint64(EPS)
    got = p2g.lib.field_ops("dot8_small", a, b32, device=device)
    bl = b32.tolist()
    want = np.array([sum(al[8 * i + k] * bl[8 * i + k] for k in range(8)) % P for i in range(len(al) // 8)], dtype=np.uint64)
    assert np.array_equal(got, want)


def test_field_ops_host_twins(p2g):
    run(p2g, -1)


@pytest.mark.gpu
def test_field_ops_device_ptx(p2g):
    run(p2g, 0)
