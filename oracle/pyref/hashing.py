"""Keccak-25 and Poseidon hashers + the Fiat-Shamir challenger (plonky2 0.2.2 hash/{keccak,poseidon,hashing}.rs,
iop/challenger.rs), restated from SURVEY.md App. A.5/A.6/D.  ORACLE = test infrastructure.

Reference call sites: plonky2-backend/src/lib.rs:13 (`C = KeccakGoldilocksConfig`, i.e. Hasher = KeccakHash<25>,
InnerHasher = PoseidonHash); PoseidonGoldilocksConfig at plonky2_ecdsa/biguint/gates/arithmetic_u32.rs:469-477.
"""
import struct

from .field import P
from .poseidon_constants import ALL_ROUND_CONSTANTS, MDS_CIRC, MDS_DIAG

M64 = (1 << 64) - 1

# ---------------------------------------------------------------- Keccak-f[1600] / Keccak-256 (original 0x01 padding)
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
    0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
    0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
    0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & M64 if n else x


def keccak_f(A):
    """A: list of 25 lanes, index x + 5*y."""
    for rnd in range(24):
        C = [A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20] for x in range(5)]
        D = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [A[i] ^ D[i % 5] for i in range(25)]
        B = [0] * 25
        for x in range(5):
            for y in range(5):
                B[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(A[x + 5 * y], _ROT[x][y])
        A = [B[i] ^ ((~B[(i % 5 + 1) % 5 + 5 * (i // 5)]) & B[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        A[0] ^= _RC[rnd]
    return A


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    padlen = rate - (len(msg) % rate)
    pad = bytearray(padlen)
    pad[0] |= 0x01
    pad[-1] |= 0x80
    msg += pad
    A = [0] * 25
    for off in range(0, len(msg), rate):
        blk = struct.unpack_from("<17Q", msg, off)
        for i in range(17):
            A[i] ^= blk[i]
        A = keccak_f(A)
    return struct.pack("<4Q", *A[:4])


# ---------------------------------------------------------------- Poseidon (width 12, x^7, 4 + 22 + 4 rounds)
def _mds(s):
    return [(sum(s[(i + r) % 12] * MDS_CIRC[i] for i in range(12)) + s[r] * MDS_DIAG[r]) % P for r in range(12)]


def poseidon_permute(state, trace=None):
    """Naive round form.  `trace`, if a list, receives the S-box inputs per round (for gate rows)."""
    s = [x % P for x in state]
    assert len(s) == 12
    for r in range(30):
        s = [(s[i] + ALL_ROUND_CONSTANTS[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            if trace is not None:
                trace.append(list(s))
            s = [pow(x, 7, P) for x in s]
        else:
            if trace is not None:
                trace.append([s[0]])
            s[0] = pow(s[0], 7, P)
        s = _mds(s)
    return s


# ---------------------------------------------------------------- Hasher objects
class KeccakHash25:
    """KeccakHash<25>: digest = 25 bytes; to_vec = 7-byte little-endian chunks (4 field elements)."""
    name = "keccak25"
    hash_size = 25

    @staticmethod
    def hash_no_pad(elems):
        return keccak256(b"".join(struct.pack("<Q", e) for e in elems))[:25]

    @staticmethod
    def two_to_one(l, r):
        return keccak256(l + r)[:25]

    @staticmethod
    def hash_to_elems(h):
        out = []
        for i in range(0, 25, 7):
            out.append(int.from_bytes(h[i:i + 7], "little"))
        return out

    @staticmethod
    def permute(state):
        """KeccakPermutation::permute: hash onion with rejection sampling."""
        buf = b"".join(struct.pack("<Q", e) for e in state)
        out = []
        while len(out) < 12:
            buf = keccak256(buf)
            for w in struct.unpack("<4Q", buf):
                if w < P and len(out) < 12:
                    out.append(w)
        return out


class PoseidonHash:
    name = "poseidon"
    hash_size = 32

    @staticmethod
    def hash_no_pad_elems(elems):
        st = [0] * 12
        for off in range(0, len(elems), 8):
            chunk = elems[off:off + 8]
            st[:len(chunk)] = chunk
            st = poseidon_permute(st)
        return st[:4]

    @staticmethod
    def hash_no_pad(elems):
        return b"".join(struct.pack("<Q", e) for e in PoseidonHash.hash_no_pad_elems(list(elems)))

    @staticmethod
    def two_to_one(l, r):
        le = list(struct.unpack("<4Q", l))
        re_ = list(struct.unpack("<4Q", r))
        return b"".join(struct.pack("<Q", e) for e in poseidon_permute(le + re_ + [0] * 4)[:4])

    @staticmethod
    def hash_to_elems(h):
        return list(struct.unpack("<4Q", h))

    @staticmethod
    def permute(state):
        return poseidon_permute(state)


HASHERS = {0: KeccakHash25, 1: PoseidonHash, "keccak25": KeccakHash25, "poseidon": PoseidonHash}


def hash_or_noop(H, elems):
    """Hasher::hash_or_noop: inputs that fit in a digest are copied, not hashed (plonky2 hash/hash_types.rs)."""
    if len(elems) * 8 <= H.hash_size:
        b = b"".join(struct.pack("<Q", e) for e in elems)
        return b + bytes(H.hash_size - len(b))
    return H.hash_no_pad(elems)


def hash_pad(H, elems):
    """Hasher::hash_pad: append 1, zero-fill so that len+1 is a multiple of RATE=8, append 1."""
    x = list(elems) + [1]
    while (len(x) + 1) % 8 != 0:
        x.append(0)
    x.append(1)
    return H.hash_no_pad(x)


class Challenger:
    """Overwrite-mode duplex sponge, width 12 / rate 8 (plonky2 iop/challenger.rs)."""

    def __init__(self, H):
        self.H = H
        self.state = [0] * 12
        self.inp = []
        self.out = []

    def observe(self, e):
        self.out = []
        self.inp.append(e % P)
        if len(self.inp) == 8:
            self._duplex()

    def observe_many(self, es):
        for e in es:
            self.observe(e)

    def observe_hash(self, h):
        self.observe_many(self.H.hash_to_elems(h))

    def observe_cap(self, cap):
        for h in cap:
            self.observe_hash(h)

    def observe_ext(self, e):
        self.observe(e.c0)
        self.observe(e.c1)

    def get_challenge(self):
        if self.inp or not self.out:
            self._duplex()
        return self.out.pop()

    def get_n(self, n):
        return [self.get_challenge() for _ in range(n)]

    def get_ext(self):
        from .field import E2
        c0 = self.get_challenge()
        c1 = self.get_challenge()
        return E2(c0, c1)

    def _duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = self.H.permute(self.state)
        self.out = list(self.state[:8])
