"""CPU: libp2acir.so exports what include/p2acir.h declares, the header compiles as C, and the big-integer arithmetic behind the
non-native witness generators (acir/bigint.h: schoolbook multiply, Knuth's algorithm D, modular exponentiation, the GLV
decomposition of glv.rs:46-91) agrees with Python integers -- including the divisions that need algorithm D's rare corrections."""
import ctypes as C
import os
import random
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


def test_header_symbols_are_exported_and_header_is_c(p2g, tmp_path):
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "p2acir.h")).read(), flags=re.S)
    names = sorted(set(re.findall(r"\b(p2a_[a-z0-9_]+)\s*\(", src)))
    assert len(names) == 14
    L = C.CDLL(p2g.acir.build())
    assert not [n for n in names if not hasattr(L, n)]
    c = tmp_path / "t.c"
    c.write_text('#include "p2acir.h"\nint main(void) { const char* (*f)(void) = p2a_last_error; return f ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o", str(tmp_path / "t.o")])


def _call(L, op, a, b=0, m=0):
    def arr(v):
        l = []
        while v:
            l.append(v & 0xFFFFFFFF)
            v >>= 32
        return (C.c_uint32 * max(1, len(l)))(*l), len(l)
    (pa, na), (pb, nb), (pm, nm) = arr(a), arr(b), arr(m)
    q, r = (C.c_uint32 * 40)(), (C.c_uint32 * 40)()
    nq, nr, fl = C.c_size_t(), C.c_size_t(), C.c_uint32()
    rc = L.p2a_bigint_selftest(op, pa, na, pb, nb, pm, nm, q, C.byref(nq), r, C.byref(nr), C.byref(fl))
    assert rc == 0, L.p2a_last_error()
    val = lambda x, n: sum(int(x[i]) << (32 * i) for i in range(n.value))   # noqa: E731
    return val(q, nq), val(r, nr), fl.value


def test_bigint_against_python_integers(p2g):
    L = p2g.acir._acir_lib()
    L.p2a_bigint_selftest.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rng = random.Random(41)
    pairs = [
        # Hacker's Delight divmnu corner cases, widened to 32-bit digits: qhat one too large / the add-back step
        (0x80000000_fffffffe_00000000, 0x80000000_ffffffff),
        (0x7fffffff_80000000_00000000_00000000, 0x80000000_00000000_00000001),
        (0x80000000_00000000_00000003, 0x20000000_00000000_00000001),
        (0x00008000_00007fff_00000000, 0x00008000_00000001),
        (0xffffffff_ffffffff_ffffffff_ffffffff, 0xffffffff_ffffffff),
        (0xffffffff_ffffffff_ffffffff_ffffffff, 0x1_00000000),
        ((1 << 512) - 1, N), (N * N, N), (N * N - 1, N), (5, 7), (0, 3), (1 << 255, 1 << 255), ((1 << 256) - 1, 3),
    ]
    for _ in range(600):
        na, nb = rng.randrange(1, 560), rng.randrange(1, 300)
        a, b = rng.getrandbits(na), rng.getrandbits(nb) | 1
        if rng.random() < 0.3:   # divisors with long runs of ones / zeros make qhat estimates fail more often
            b = ((1 << nb) - 1) ^ rng.getrandbits(max(1, nb // 3))
            b |= 1
        pairs.append((a, b))
    for a, b in pairs:
        assert _call(L, 0, a, b)[:2] == (a // b, a % b), (hex(a), hex(b))
        if a.bit_length() + b.bit_length() <= 1200:
            assert _call(L, 1, a, b)[0] == a * b
    for _ in range(20):
        x, e = rng.getrandbits(300), rng.getrandbits(256)
        assert _call(L, 2, x, e, N)[0] == pow(x, e, N)
    assert _call(L, 2, 12345, N - 2, N)[0] == pow(12345, -1, N)
    PF = 2 ** 256 - 2 ** 32 - 977
    for mod in (N, PF, 0xFFFFFFFF00000001, 97):      # binary extended Euclid against Python's modular inverse
        for x in [1, 2, mod - 1, mod - 2, mod + 5, (mod * 3 + 7)] + [rng.getrandbits(300) for _ in range(100)]:
            if x % mod:
                assert _call(L, 4, x, 0, mod)[0] == pow(x, -1, mod), (hex(x), hex(mod))
    # GLV: |k1|, |k2| < 2^128, and k1 + s k2 = k with the signs applied
    S = sum(v << (64 * i) for i, v in enumerate([16069571880186789234, 1310022930574435960, 11900229862571533402, 6008836872998760672]))
    for k in [0, 1, N - 1, N // 2, N // 2 + 1] + [rng.randrange(N) for _ in range(200)]:
        k1, k2, fl = _call(L, 3, k)
        assert k1 < 1 << 128 and k2 < 1 << 128
        s1 = -k1 if fl & 1 else k1
        s2 = -k2 if fl & 2 else k2
        assert (s1 + S * s2) % N == k
