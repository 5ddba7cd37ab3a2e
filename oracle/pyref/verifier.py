"""plonky2 0.2.2 verifier, restated (plonk/{verifier,get_challenges,vanishing_poly}.rs, fri/verifier.rs; SURVEY.md App.
A.3, A.8, A.10, A.11).  This is the ACCEPTANCE ORACLE of every reference test (`assert!(verify(proof).is_ok())`,
e.g. plonky2-backend/src/circuit_translation/tests/test_assert_zero.rs:24, tests/test_precompiled.rs:43) and of the
CLI's `verify` action (plonky2-backend/src/actions/verify_action.rs:11-17).  ORACLE = test infrastructure.
"""
from .field import (E2, P, MULTIPLICATIVE_GROUP_GENERATOR, inv, root_of_unity, reverse_bits, log2_strict)
from .hashing import HASHERS, Challenger, PoseidonHash, hash_or_noop, hash_pad
from .proof import (Proof, CompressedProof, decompress_proof, compress_proof, flatten_ext)
from .circuit import UNUSED_SELECTOR


class VerifyError(Exception):
    pass


def circuit_digest(cd, constants_sigmas_cap):
    """H.hash_no_pad(cap.flatten() || H.hash_pad([]).to_vec() || [degree_bits])  (SURVEY A.3)."""
    H = HASHERS[cd.hasher]
    parts = []
    for h in constants_sigmas_cap:
        parts += H.hash_to_elems(h)
    parts += H.hash_to_elems(hash_pad(H, []))
    parts.append(cd.degree_bits)
    return H.hash_no_pad(parts)


class Challenges:
    pass


def get_challenges(pr, cd, digest):
    H = HASHERS[cd.hasher]
    ch = Challenges()
    c = Challenger(H)
    ch.pi_hash = PoseidonHash.hash_no_pad_elems(list(pr.public_inputs)) if pr.public_inputs else [0, 0, 0, 0]
    c.observe_hash(digest)
    # public_inputs_hash is an InnerHasher (Poseidon) HashOut: observed as its 4 elements
    c.observe_many(ch.pi_hash)
    c.observe_cap(pr.wires_cap)
    ch.betas = c.get_n(cd.num_challenges)
    ch.gammas = c.get_n(cd.num_challenges)
    c.observe_cap(pr.zs_pp_cap)
    ch.alphas = c.get_n(cd.num_challenges)
    c.observe_cap(pr.quotient_cap)
    ch.zeta = c.get_ext()
    for e in pr.openings.zeta_batch() + pr.openings.zeta_next_batch():
        c.observe_ext(e)
    ch.fri_alpha = c.get_ext()
    ch.fri_betas = []
    for cap in pr.fri_caps:
        c.observe_cap(cap)
        ch.fri_betas.append(c.get_ext())
    for e in pr.final_poly:
        c.observe_ext(e)
    c.observe(pr.pow_witness)
    ch.pow_response = c.get_challenge()
    lde = 1 << cd.lde_bits
    ch.indices = [c.get_challenge() % lde for _ in range(cd.num_queries)]
    return ch


def compute_filter(row, group, s, many):
    lo, hi = group
    r = E2(1)
    for i in range(lo, hi):
        if i != row:
            r = r * (E2(i) - s)
    if many:
        r = r * (E2(UNUSED_SELECTOR) - s)
    return r


def evaluate_gate_constraints(cd, constants, wires, pi_hash):
    acc = [E2(0)] * cd.num_gate_constraints
    for i, g in enumerate(cd.gates):
        sel = cd.selector_indices[i]
        f = compute_filter(i, cd.groups[sel], constants[sel], cd.num_selectors > 1)
        res = g.eval_unfiltered(constants[cd.num_selectors:], wires, pi_hash)
        for k, v in enumerate(res):
            acc[k] = acc[k] + v * f
    return acc


def eval_vanishing_poly(cd, x, constants, sigmas, wires, zs, zs_next, pps, pi_hash, betas, gammas, alphas):
    """x in E2.  Returns [vanish_c for c in challenges]  (SURVEY A.8)."""
    n = cd.n
    constraint_terms = evaluate_gate_constraints(cd, constants, wires, pi_hash)
    zh = x ** n - 1
    l0 = zh / ((x - 1) * n)
    z1_terms = []
    pp_terms = []
    npp = cd.num_partial_products
    for c in range(cd.num_challenges):
        z1_terms.append(l0 * (zs[c] - 1))
        nums = [wires[j] + x * (betas[c] * cd.k_is[j] % P) + gammas[c] for j in range(cd.num_routed)]
        dens = [wires[j] + sigmas[j] * betas[c] + gammas[c] for j in range(cd.num_routed)]
        accs = [zs[c]] + list(pps[c * npp:(c + 1) * npp]) + [zs_next[c]]
        for m in range(npp + 1):
            pn, pd = E2(1), E2(1)
            for j in range(m * cd.qdf, min((m + 1) * cd.qdf, cd.num_routed)):
                pn = pn * nums[j]
                pd = pd * dens[j]
            pp_terms.append(accs[m] * pn - accs[m + 1] * pd)
    terms = z1_terms + pp_terms + constraint_terms
    res = []
    for a in alphas:
        acc = E2(0)
        for t in reversed(terms):
            acc = acc * a + t
        res.append(acc)
    return res


def verify_merkle_proof_to_cap(H, leaf, index, cap, path):
    cur = hash_or_noop(H, leaf)
    for sib in path:
        cur = H.two_to_one(cur, sib) if index & 1 == 0 else H.two_to_one(sib, cur)
        index >>= 1
    if cur != cap[index]:
        raise VerifyError("Invalid Merkle proof")


def _reduce(vals, alpha):
    acc = E2(0)
    for v in reversed(vals):
        acc = acc * alpha + v
    return acc


def batch_polys(cd):
    """FRI instance: batch 0 = every polynomial of oracles 0..3 at zeta; batch 1 = Z polys at g*zeta."""
    b0 = []
    for oi, w in enumerate(cd.oracle_widths()):
        b0 += [(oi, j) for j in range(w)]
    b1 = [(2, j) for j in range(cd.num_challenges)]
    return [b0, b1]


def fri_combine_initial(cd, initial_values, alpha, subgroup_x, reduced_openings, points):
    x = E2(subgroup_x)
    s = E2(0)
    for polys, ro, pt in zip(batch_polys(cd), reduced_openings, points):
        evals = [initial_values[oi][pj] for oi, pj in polys]
        red = _reduce(evals, alpha)
        s = s * (alpha ** len(evals))
        s = s + (red - ro) / (x - pt)
    return s


def compute_evaluation(x, within, arity_bits, evals, beta):
    """fri/verifier? no: fri/mod `compute_evaluation`: interpolate the 16 coset values, evaluate at beta."""
    arity = 1 << arity_bits
    g = root_of_unity(arity_bits)
    ev = [evals[reverse_bits(i, arity_bits)] for i in range(arity)]
    rev = reverse_bits(within, arity_bits)
    start = x * pow(g, arity - rev, P) % P
    pts = [start * pow(g, i, P) % P for i in range(arity)]
    # Lagrange interpolation at beta
    res = E2(0)
    for i in range(arity):
        num, den = E2(1), 1
        for j in range(arity):
            if j != i:
                num = num * (beta - pts[j])
                den = den * (pts[i] - pts[j]) % P
        res = res + ev[i] * num * inv(den)
    return res


def _fri_points(cd, zeta):
    g = root_of_unity(cd.degree_bits)
    return [zeta, zeta * g]


def _precomputed(pr, ch):
    return [_reduce(pr.openings.zeta_batch(), ch.fri_alpha), _reduce(pr.openings.zeta_next_batch(), ch.fri_alpha)]


def decompress(cp, cd, digest):
    ch = get_challenges(cp, cd, digest)
    if ch.indices != cp.indices:
        raise VerifyError("query indices do not match the transcript")
    ro = _precomputed(cp, ch)
    pts = _fri_points(cd, ch.zeta)
    cache = {}

    def inferred(qn, index, ini_vals, layer_evals, layer):
        # replay the fold chain of this query up to `layer`
        lde_bits = cd.lde_bits
        x = MULTIPLICATIVE_GROUP_GENERATOR * pow(root_of_unity(lde_bits), reverse_bits(index, lde_bits), P) % P
        old = fri_combine_initial(cd, ini_vals, ch.fri_alpha, x, ro, pts)
        xi = index
        for j in range(layer):
            ab = cd.arity_bits[j]
            within = xi & ((1 << ab) - 1)
            old = compute_evaluation(x, within, ab, layer_evals[j], ch.fri_betas[j])
            x = pow(x, 1 << ab, P)
            xi >>= ab
        return old
    return decompress_proof(cp, cd, inferred), ch


def verify(pr, cd, constants_sigmas_cap, digest=None):
    """verify_with_challenges + verify_fri_proof.  `pr` uncompressed Proof.  Raises VerifyError."""
    H = HASHERS[cd.hasher]
    if digest is None:
        digest = circuit_digest(cd, constants_sigmas_cap)
    ch = get_challenges(pr, cd, digest)
    os_ = pr.openings
    zeta = ch.zeta
    van = eval_vanishing_poly(cd, zeta, os_.constants, os_.sigmas, os_.wires, os_.zs, os_.zs_next,
                              os_.partial_products, ch.pi_hash, ch.betas, ch.gammas, ch.alphas)
    zh = zeta ** cd.n - 1
    zeta_n = zeta ** cd.n
    for c in range(cd.num_challenges):
        chunk = os_.quotient[c * cd.qdf:(c + 1) * cd.qdf]
        if not (van[c] == zh * _reduce(chunk, zeta_n)):
            raise VerifyError(f"PLONK identity fails for challenge {c}")
    # ---- FRI
    if len(pr.final_poly) != cd.final_poly_len:
        raise VerifyError("final poly length")
    lz = 64 - ch.pow_response.bit_length()
    if lz < cd.pow_bits:
        raise VerifyError("Invalid proof of work witness")
    ro = _precomputed(pr, ch)
    pts = _fri_points(cd, zeta)
    caps = [constants_sigmas_cap, pr.wires_cap, pr.zs_pp_cap, pr.quotient_cap]
    lde_bits = cd.lde_bits
    if len(pr.query_rounds) != cd.num_queries:
        raise VerifyError("number of query rounds")
    for x_index, qr in zip(ch.indices, pr.query_rounds):
        for (vals, path), cap in zip(qr.initial, caps):
            if len(path) != lde_bits - cd.cap_height:
                raise VerifyError("path length")
            verify_merkle_proof_to_cap(H, vals, x_index, cap, path)
        x = MULTIPLICATIVE_GROUP_GENERATOR * pow(root_of_unity(lde_bits), reverse_bits(x_index, lde_bits), P) % P
        old = fri_combine_initial(cd, [v for v, _ in qr.initial], ch.fri_alpha, x, ro, pts)
        xi = x_index
        for j, ab in enumerate(cd.arity_bits):
            evals, path = qr.steps[j]
            coset_index = xi >> ab
            within = xi & ((1 << ab) - 1)
            if not (evals[within] == old):
                raise VerifyError(f"FRI fold consistency fails at layer {j}")
            old = compute_evaluation(x, within, ab, evals, ch.fri_betas[j])
            verify_merkle_proof_to_cap(H, flatten_ext(evals), coset_index, pr.fri_caps[j], path)
            x = pow(x, 1 << ab, P)
            xi = coset_index
        acc = E2(0)
        for cf in reversed(pr.final_poly):
            acc = acc * x + cf
        if not (acc == old):
            raise VerifyError("FRI final polynomial check fails")
    return ch


def verify_compressed(cp, cd, constants_sigmas_cap, digest=None):
    if digest is None:
        digest = circuit_digest(cd, constants_sigmas_cap)
    pr, _ = decompress(cp, cd, digest)
    return verify(pr, cd, constants_sigmas_cap, digest)
