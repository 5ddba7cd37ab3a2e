"""Goldilocks field and its quadratic extension, plonky2_field 0.2.2 conventions.

ORACLE = test infrastructure.  Follows plonky2_field/src/goldilocks_field.rs (ORDER,
MULTIPLICATIVE_GROUP_GENERATOR, POWER_OF_TWO_GENERATOR) and extension/quadratic.rs (W = 7); the two
generators are the ones verified against the golden proofs in SURVEY.md App. A.1 / F-table.
Reference call site of everything built on this: plonky2-backend/src/lib.rs:8-13 (F = GoldilocksField, D = 2).
"""

P = (1 << 64) - (1 << 32) + 1
MULTIPLICATIVE_GROUP_GENERATOR = 14293326489335486720
POWER_OF_TWO_GENERATOR = 7277203076849721926
TWO_ADICITY = 32
W = 7  # F_{p^2} = F_p[X]/(X^2 - 7)
COSET_SHIFT = MULTIPLICATIVE_GROUP_GENERATOR


def inv(a):
    return pow(a % P, P - 2, P)


def root_of_unity(bits):
    """F::primitive_root_of_unity(bits) = POWER_OF_TWO_GENERATOR^(2^(32-bits))."""
    assert 0 <= bits <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - bits), P)


def reverse_bits(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def reverse_index_bits(lst):
    n = len(lst)
    bits = n.bit_length() - 1
    assert 1 << bits == n
    return [lst[reverse_bits(i, bits)] for i in range(n)]


def log2_strict(n):
    b = n.bit_length() - 1
    assert 1 << b == n
    return b


class E2:
    """Element c0 + c1*X of F_p[X]/(X^2-7).  Accepts ints on either side of + - *."""
    __slots__ = ("c0", "c1")

    def __init__(self, c0=0, c1=0):
        self.c0 = c0 % P
        self.c1 = c1 % P

    @staticmethod
    def lift(x):
        return x if isinstance(x, E2) else E2(x, 0)

    def __add__(self, o):
        o = E2.lift(o)
        return E2(self.c0 + o.c0, self.c1 + o.c1)
    __radd__ = __add__

    def __sub__(self, o):
        o = E2.lift(o)
        return E2(self.c0 - o.c0, self.c1 - o.c1)

    def __rsub__(self, o):
        return E2.lift(o) - self

    def __neg__(self):
        return E2(-self.c0, -self.c1)

    def __mul__(self, o):
        if isinstance(o, int):
            return E2(self.c0 * o, self.c1 * o)
        return E2(self.c0 * o.c0 + W * self.c1 * o.c1, self.c0 * o.c1 + self.c1 * o.c0)
    __rmul__ = __mul__

    def inverse(self):
        # 1/(a+bX) = (a-bX)/(a^2 - 7 b^2)
        d = inv(self.c0 * self.c0 - W * self.c1 * self.c1)
        return E2(self.c0 * d, -self.c1 * d)

    def __truediv__(self, o):
        return self * E2.lift(o).inverse()

    def __pow__(self, e):
        r, b = E2(1, 0), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def __eq__(self, o):
        o = E2.lift(o)
        return self.c0 == o.c0 and self.c1 == o.c1

    def __hash__(self):
        return hash((self.c0, self.c1))

    def __repr__(self):
        return f"E2({self.c0}, {self.c1})"

    def is_zero(self):
        return self.c0 == 0 and self.c1 == 0


ZERO2 = E2(0, 0)
ONE2 = E2(1, 0)


def ntt_naive(vals, root):
    """out[i] = sum_k vals[k] root^(ik); O(n^2); small sizes only."""
    n = len(vals)
    out = []
    for i in range(n):
        w = pow(root, i, P)
        acc, x = 0, 1
        for k in range(n):
            acc = (acc + vals[k] * x) % P
            x = x * w % P
        out.append(acc)
    return out


def fft(coeffs):
    """Natural-order radix-2 NTT (values on <omega_n>, natural order). plonky2_field fft.rs semantics."""
    n = len(coeffs)
    if n == 1:
        return list(coeffs)
    bits = log2_strict(n)
    a = reverse_index_bits(list(coeffs))
    m = 1
    while m < n:
        w_m = root_of_unity(log2_strict(2 * m))
        for s in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                u, v = a[s + j], a[s + j + m] * w % P
                a[s + j] = (u + v) % P
                a[s + j + m] = (u - v) % P
                w = w * w_m % P
        m *= 2
    return a


def ifft(vals):
    n = len(vals)
    out = fft(vals)
    ninv = inv(n)
    res = [0] * n
    for i in range(n):
        res[i] = out[(n - i) % n] * ninv % P
    return res


def coset_fft(coeffs, shift):
    c, s = [], 1
    for x in coeffs:
        c.append(x * s % P)
        s = s * shift % P
    return fft(c)


def coset_ifft(vals, shift):
    c = ifft(vals)
    si = inv(shift)
    out, s = [], 1
    for x in c:
        out.append(x * s % P)
        s = s * si % P
    return out


def eval_poly_ext(coeffs, z):
    """Horner in F_{p^2} of a base-field coefficient list."""
    acc = ZERO2
    for c in reversed(coeffs):
        acc = acc * z + c
    return acc
