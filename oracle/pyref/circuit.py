"""CommonCircuitData restated (plonky2 0.2.2 plonk/circuit_data.rs, gates/selectors.rs, fri/reduction_strategies.rs;
SURVEY.md App. A.2).  ORACLE = test infrastructure.

Reference parameters in force: plonky2-backend/src/circuit_translation/mod.rs:69 (wide_ecc_config: 234 wires, 80 routed),
plonky2-backend/src/lib.rs:11-13 (D = 2, KeccakGoldilocksConfig).
"""
from .field import P, MULTIPLICATIVE_GROUP_GENERATOR
from .gates import Gate

UNUSED_SELECTOR = (1 << 32) - 1


def reduction_arity_bits(degree_bits, rate_bits=3, cap_height=4, arity_bits=4, final_poly_bits=5):
    """FriReductionStrategy::ConstantArityBits(4, 5)."""
    out = []
    d = degree_bits
    while d > final_poly_bits and d + rate_bits - arity_bits >= cap_height:
        out.append(arity_bits)
        d -= arity_bits
    return out


def selector_groups(gates, max_degree=9):
    """gates sorted by (degree, id).  Returns (selector_indices, groups) as gates/selectors.rs::selector_polynomials."""
    num_gates = len(gates)
    max_gate_degree = gates[-1].degree
    if max_gate_degree + num_gates - 1 <= max_degree:
        return [0] * num_gates, [(0, num_gates)]
    assert max_gate_degree < max_degree
    groups = []
    start = 0
    while start < num_gates:
        size = 0
        while start + size < num_gates and size + gates[start + size].degree < max_degree:
            size += 1
        groups.append((start, start + size))
        start += size
    sel = []
    for i in range(num_gates):
        for gi, (lo, hi) in enumerate(groups):
            if lo <= i < hi:
                sel.append(gi)
    return sel, groups


class CommonData:
    def __init__(self, degree_bits, gates, num_wires=234, num_routed=80, num_public_inputs=0, hasher="keccak25",
                 num_challenges=2, rate_bits=3, cap_height=4, pow_bits=16, num_queries=28, qdf=8,
                 config_num_constants=2):
        self.degree_bits = degree_bits
        self.n = 1 << degree_bits
        self.gates = sorted(gates, key=lambda g: g.sort_key())
        self.num_wires = num_wires
        self.num_routed = num_routed
        self.num_public_inputs = num_public_inputs
        self.hasher = hasher
        self.num_challenges = num_challenges
        self.rate_bits = rate_bits
        self.cap_height = cap_height
        self.pow_bits = pow_bits
        self.num_queries = num_queries
        self.qdf = qdf
        self.selector_indices, self.groups = selector_groups(self.gates, qdf + 1)
        self.num_selectors = len(self.groups)
        self.num_gate_constants = max(g.num_constants for g in self.gates)
        self.num_constants = self.num_selectors + self.num_gate_constants  # constant columns incl. selectors
        self.num_partial_products = (num_routed + qdf - 1) // qdf - 1
        self.num_gate_constraints = max(g.num_constraints for g in self.gates)
        self.k_is = [pow(MULTIPLICATIVE_GROUP_GENERATOR, i, P) for i in range(num_routed)]
        self.arity_bits = reduction_arity_bits(degree_bits, rate_bits, cap_height)
        self.lde_bits = degree_bits + rate_bits

    @property
    def num_preprocessed(self):
        return self.num_constants + self.num_routed

    @property
    def num_zs_pp(self):
        return self.num_challenges * (1 + self.num_partial_products)

    @property
    def num_quotient(self):
        return self.num_challenges * self.qdf

    @property
    def final_poly_len(self):
        return 1 << (self.degree_bits - sum(self.arity_bits))

    def oracle_widths(self):
        return [self.num_preprocessed, self.num_wires, self.num_zs_pp, self.num_quotient]

    def gate_index(self, gate):
        for i, g in enumerate(self.gates):
            if g.kind == gate.kind and g.params == gate.params:
                return i
        raise KeyError(gate)
