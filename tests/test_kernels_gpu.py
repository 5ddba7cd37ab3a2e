"""Parity of the stand-alone sm_100a kernels (through the C ABI) against the CPU oracle, bit-exact.
Mirrors the reference's use of PolynomialValues::ifft / PolynomialCoeffs::lde().fft() in
plonky2-backend/src/plonky2_ecdsa/biguint/gates/gate_testing.rs:78-83 and MerkleTree::new inside prove()."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


def rand_field(rng, shape):
    return rng.integers(0, P, size=shape, dtype=np.uint64)


@pytest.mark.parametrize("log_n,ncols", [(0, 3), (1, 2), (2, 5), (3, 135), (5, 7), (8, 3), (10, 4), (11, 3), (12, 5),
                                         (13, 2), (14, 3), (16, 2)])
def test_ifft_matches_oracle(p2g, corc, log_n, ncols):
    rng = np.random.default_rng(100 + log_n)
    v = rand_field(rng, (ncols, 1 << log_n))
    assert np.array_equal(p2g.lib.ifft(v), corc.ifft(v))


@pytest.mark.parametrize("log_n,ncols", [(0, 3), (1, 2), (3, 20), (4, 3), (7, 5), (10, 3), (11, 2), (12, 3), (13, 2),
                                         (14, 2), (16, 1)])
def test_lde_matches_oracle(p2g, corc, log_n, ncols):
    rng = np.random.default_rng(200 + log_n)
    c = rand_field(rng, (ncols, 1 << log_n))
    assert np.array_equal(p2g.lib.lde(c, 3), corc.lde(c, 3))


@pytest.mark.parametrize("log_n,ncols", [(1, 2), (3, 2), (6, 4), (10, 2), (11, 2), (12, 2), (13, 2), (15, 2), (17, 2), (19, 1)])
def test_coset_ifft_leaforder_matches_oracle(p2g, corc, log_n, ncols):
    rng = np.random.default_rng(300 + log_n)
    v = rand_field(rng, (ncols, 1 << log_n))
    assert np.array_equal(p2g.lib.coset_ifft_leaforder(v), corc.coset_ifft_leaforder(v))


def test_lde_roundtrip_full_size(p2g):
    """BASELINE size (2^20 rows): coefficients -> 8N-point LDE (leaf order) -> coset iNTT returns the zero-padded input,
    and ifft(values) composed with the first LDE coset is consistent: a size-independent property."""
    rng = np.random.default_rng(7)
    log_n = 20
    c = rand_field(rng, (2, 1 << log_n))
    l = p2g.lib.lde(c, 3)
    back = p2g.lib.coset_ifft_leaforder(l)
    assert np.array_equal(back[:, :1 << log_n], c)
    assert not back[:, 1 << log_n:].any()


def test_ifft_linearity_full_size(p2g):
    rng = np.random.default_rng(8)
    a = rand_field(rng, (1, 1 << 20))
    b = rand_field(rng, (1, 1 << 20))
    s = ((a.astype(object) + b.astype(object)) % P).astype(np.uint64)
    fa, fb, fs = p2g.lib.ifft(a), p2g.lib.ifft(b), p2g.lib.ifft(s)
    assert np.array_equal(((fa.astype(object) + fb.astype(object)) % P).astype(np.uint64), fs)


@pytest.mark.parametrize("hasher", ["keccak25", "poseidon"])
@pytest.mark.parametrize("log_leaves,ncols,cap_height", [(0, 5, 4), (2, 3, 4), (3, 4, 0), (4, 1, 4), (6, 135, 4), (6, 17, 4),
                                                          (7, 34, 4), (8, 16, 4), (9, 20, 2), (10, 83, 4), (11, 234, 4),
                                                          (12, 8, 4), (5, 32, 4), (6, 9, 6)])
def test_merkle_cap_matches_oracle(p2g, corc, hasher, log_leaves, ncols, cap_height):
    rng = np.random.default_rng(400 + log_leaves + ncols)
    leaves = rand_field(rng, (ncols, 1 << log_leaves))
    cap, dg = p2g.lib.merkle_cap(leaves, cap_height, hasher, want_digests=True)
    rcap, rdg = corc.merkle_cap(leaves, cap_height, hasher, want_digests=True)
    assert np.array_equal(dg, rdg)
    assert np.array_equal(cap, rcap)


def test_poseidon_permute_matches_oracle(p2g, corc):
    rng = np.random.default_rng(5)
    st = rand_field(rng, (1000, 12))
    st[0] = 0
    st[1] = 0
    st[1, 0] = 1
    out = p2g.lib.poseidon_permute(st)
    assert np.array_equal(out, corc.poseidon_permute(st))
    # SURVEY.md App. D known answers (plonky2's published vector and the golden-proof PoseidonGate row)
    assert int(out[0, 0]) == 0x3c18a9786cb0b359
    assert int(out[1, 0]) == 0xd074b8cee5dcf415


@pytest.mark.parametrize("msg_len", [0, 1, 8, 50, 96, 135, 136, 137, 200, 1872])
def test_keccak256_matches_oracle(p2g, corc, msg_len):
    rng = np.random.default_rng(6)
    m = rng.integers(0, 256, size=(300, msg_len), dtype=np.uint8)
    assert np.array_equal(p2g.lib.keccak256(m), corc.keccak256(m))
