"""Decode the two golden proofs the reference commits (plonky2-backend/example_programs/basic_{if,div}/proofs/*.proof,
copied verbatim to tests/golden/) and recover from them the complete FFI payload of their 8-row circuits
(SURVEY.md App. C steps 1-9).  ORACLE = test infrastructure.

Because N = 8 and 25 distinct rows of the 64-point LDE are opened, every committed polynomial (<= 8 coefficients) is
determined by the file: interpolating gives the wire / constant / sigma / Z / quotient polynomials in coefficient form;
evaluating the preprocessed ones on the whole coset rebuilds `constants_sigmas_cap` (the VK cap), after which the proof
verifies end to end.
"""
import os

from .field import P, MULTIPLICATIVE_GROUP_GENERATOR as G, inv, root_of_unity, reverse_bits, fft
from .gates import Gate
from .circuit import CommonData
from .hashing import HASHERS, hash_or_noop
from .proof import parse_compressed

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden")

GOLDEN = {
    # gate sets per SURVEY App. C step 7; both use the older 135-wire standard_recursion_config
    "basic_if": dict(gates=lambda: [Gate.noop(), Gate.constant(2), Gate.public_input(), Gate.base_sum(2, 63), Gate.arithmetic(80)],
                     num_public_inputs=0),
    "basic_div": dict(gates=lambda: [Gate.noop(), Gate.constant(2), Gate.public_input(), Gate.arithmetic(80), Gate.poseidon()],
                      num_public_inputs=1),
}


def golden_common_data(name):
    g = GOLDEN[name]
    return CommonData(3, g["gates"](), num_wires=135, num_routed=80, num_public_inputs=g["num_public_inputs"], hasher="keccak25")


def read_golden_bytes(name):
    with open(os.path.join(GOLDEN_DIR, name + ".proof")) as f:
        return bytes.fromhex(f.read().strip())


def _interpolate(xs, ys):
    """Coefficients of the unique poly of degree < len(xs) through (xs, ys)."""
    n = len(xs)
    coeffs = [0] * n
    for i in range(n):
        # basis_i(X) = prod_{j != i} (X - x_j) / (x_i - x_j)
        b = [1]
        den = 1
        for j in range(n):
            if j == i:
                continue
            nb = [0] * (len(b) + 1)
            for k, c in enumerate(b):
                nb[k + 1] = (nb[k + 1] + c) % P
                nb[k] = (nb[k] - c * xs[j]) % P
            b = nb
            den = den * (xs[i] - xs[j]) % P
        s = ys[i] * inv(den) % P
        for k in range(n):
            coeffs[k] = (coeffs[k] + b[k] * s) % P
    return coeffs


def _horner(c, x):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % P
    return acc


def merkle_cap_from_leaves(H, leaves, cap_height):
    layer = [hash_or_noop(H, l) for l in leaves]
    while len(layer) > (1 << cap_height):
        layer = [H.two_to_one(layer[2 * i], layer[2 * i + 1]) for i in range(len(layer) // 2)]
    return layer


def recover(name):
    """Returns dict(cd, cproof, raw, coeffs=[4 oracles][col][8], trace=dict(constants, sigmas, wires), cs_cap)."""
    cd = golden_common_data(name)
    raw = read_golden_bytes(name)
    cp = parse_compressed(raw, cd)
    H = HASHERS[cd.hasher]
    lde_bits = cd.lde_bits
    w = root_of_unity(lde_bits)
    idxs = sorted(cp.initial_by_index)
    xs = [G * pow(w, reverse_bits(i, lde_bits), P) % P for i in idxs]
    coeffs = []
    for oi, width in enumerate(cd.oracle_widths()):
        cols = []
        for c in range(width):
            ys = [cp.initial_by_index[i][oi][0][c] for i in idxs]
            co = _interpolate(xs[:8], ys[:8])
            for x, y in zip(xs[8:], ys[8:]):
                if _horner(co, x) != y:
                    raise ValueError(f"{name}: oracle {oi} column {c} is not a degree<8 polynomial on the opened rows")
            cols.append(co)
        coeffs.append(cols)
    # trace values on <omega_8>
    vals = [[fft(co) for co in cols] for cols in coeffs]
    trace = dict(constants=vals[0][:cd.num_constants], sigmas=vals[0][cd.num_constants:], wires=vals[1],
                 zs_pp=vals[2])
    # rebuild the preprocessed commitment: LDE on the 64-point coset, bit-reversed leaves, Keccak Merkle, cap
    lde = 1 << lde_bits
    pts = [G * pow(w, reverse_bits(j, lde_bits), P) % P for j in range(lde)]
    leaves = [[_horner(co, x) for co in coeffs[0]] for x in pts]
    cs_cap = merkle_cap_from_leaves(H, leaves, cd.cap_height)
    return dict(cd=cd, cproof=cp, raw=raw, coeffs=coeffs, trace=trace, cs_cap=cs_cap)
