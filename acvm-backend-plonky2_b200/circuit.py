"""Host-side mirror of the reference interface for the hot path.

The reference reaches the prover through plonky2's `CircuitData` (`builder.build::<C>()` at
plonky2-backend/src/circuit_translation/mod.rs:81, `circuit_data.prove(witnesses)` at
plonky2-backend/src/actions/prove_action.rs:96).  This module keeps those names: `CircuitConfig`, `CommonCircuitData`,
`CircuitData.prove` -> `ProofWithPublicInputs`, with plonky2's argument meaning and its error behaviour (the reference
`.unwrap()`s, so failures raise).  It computes nothing itself: it fills the `p2g_circuit_desc` of include/p2g.h and calls
libp2g.so.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib

P = 0xFFFFFFFF00000001
MULTIPLICATIVE_GROUP_GENERATOR = 14293326489335486720
UNUSED_SELECTOR = (1 << 32) - 1

# gate kinds = enum p2g_gate_kind (include/p2g.h)
(NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, BASE_SUM, POSEIDON, RANDOM_ACCESS, U32_ARITHMETIC, U32_ADD_MANY,
 U32_SUBTRACTION, U32_RANGE_CHECK, COMPARISON) = range(12)
_GF = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"


def _cdiv(a, b):
    return -(-a // b)


# kind -> (name, degree(p), num_constraints(p), num_constants(p), id-format)   p = params list
_GATE_TABLE = {
    NOOP: ("NoopGate", lambda p: 0, lambda p: 0, lambda p: 0, lambda p: "NoopGate"),
    CONSTANT: ("ConstantGate", lambda p: 1, lambda p: p[0], lambda p: p[0],
               lambda p: f"ConstantGate {{ num_consts: {p[0]} }}"),
    PUBLIC_INPUT: ("PublicInputGate", lambda p: 1, lambda p: 4, lambda p: 0, lambda p: "PublicInputGate"),
    ARITHMETIC: ("ArithmeticGate", lambda p: 3, lambda p: p[0], lambda p: 2,
                 lambda p: f"ArithmeticGate {{ num_ops: {p[0]} }}"),
    BASE_SUM: ("BaseSumGate", lambda p: p[0], lambda p: 1 + p[1], lambda p: 0,
               lambda p: f"BaseSumGate {{ num_limbs: {p[1]} }} + Base: {p[0]}"),
    POSEIDON: ("PoseidonGate", lambda p: 7, lambda p: 123, lambda p: 0, lambda p: f"PoseidonGate({_GF})<WIDTH=12>"),
    RANDOM_ACCESS: ("RandomAccessGate", lambda p: p[0] + 1, lambda p: (p[0] + 2) * p[1] + p[2], lambda p: p[2],
                    lambda p: f"RandomAccessGate {{ bits: {p[0]}, num_copies: {p[1]}, num_extra_constants: {p[2]}, "
                              f"_phantom: {_GF} }}<D=2>"),
    U32_ARITHMETIC: ("U32ArithmeticGate", lambda p: 4, lambda p: 36 * p[0], lambda p: 0,
                     lambda p: f"U32ArithmeticGate {{ num_ops: {p[0]}, _phantom: {_GF} }}"),
    U32_ADD_MANY: ("U32AddManyGate", lambda p: 4, lambda p: 21 * p[1], lambda p: 0,
                   lambda p: f"U32AddManyGate {{ num_addends: {p[0]}, num_ops: {p[1]}, _phantom: {_GF} }}"),
    U32_SUBTRACTION: ("U32SubtractionGate", lambda p: 4, lambda p: 19 * p[0], lambda p: 0,
                      lambda p: f"U32SubtractionGate {{ num_ops: {p[0]}, _phantom: {_GF} }}"),
    U32_RANGE_CHECK: ("U32RangeCheckGate", lambda p: 4, lambda p: 17 * p[0], lambda p: 0,
                      lambda p: f"U32RangeCheckGate {{ num_input_limbs: {p[0]}, _phantom: {_GF} }}"),
    COMPARISON: ("ComparisonGate", lambda p: 1 << _cdiv(p[0], p[1]), lambda p: 6 + 5 * p[1] + _cdiv(p[0], p[1]),
                 lambda p: 0,
                 lambda p: f"ComparisonGate {{ num_bits: {p[0]}, num_chunks: {p[1]}, _phantom: {_GF} }}<D=2>"),
}


@dataclass(frozen=True)
class CircuitConfig:
    """plonky2 `CircuitConfig` + `FriConfig`, the fields the prover reads."""
    num_wires: int = 135
    num_routed_wires: int = 80
    num_constants: int = 2
    num_challenges: int = 2
    max_quotient_degree_factor: int = 8
    rate_bits: int = 3
    cap_height: int = 4
    proof_of_work_bits: int = 16
    num_query_rounds: int = 28
    fri_arity_bits: int = 4          # FriReductionStrategy::ConstantArityBits(4, 5)
    fri_final_poly_bits: int = 5
    hasher: str = "keccak25"         # plonky2-backend/src/lib.rs:13: C = KeccakGoldilocksConfig

    @staticmethod
    def standard_recursion_config(**kw):
        return CircuitConfig(**kw)

    @staticmethod
    def standard_ecc_config(**kw):
        return CircuitConfig(num_wires=136, **kw)

    @staticmethod
    def wide_ecc_config(**kw):
        """The config the backend builds every circuit with (circuit_translation/mod.rs:69)."""
        return CircuitConfig(num_wires=234, **kw)


@dataclass(frozen=True)
class Gate:
    kind: int
    params: tuple = (0, 0, 0, 0)

    def __post_init__(self):
        p = tuple(self.params) + (0,) * (4 - len(self.params))
        object.__setattr__(self, "params", p)

    name = property(lambda s: _GATE_TABLE[s.kind][0])
    degree = property(lambda s: _GATE_TABLE[s.kind][1](s.params))
    num_constraints = property(lambda s: _GATE_TABLE[s.kind][2](s.params))
    num_constants = property(lambda s: _GATE_TABLE[s.kind][3](s.params))
    id = property(lambda s: _GATE_TABLE[s.kind][4](s.params))

    # constructors with the reference's `new_from_config` arithmetic
    @staticmethod
    def noop():
        return Gate(NOOP)

    @staticmethod
    def constant(cfg):
        return Gate(CONSTANT, (cfg.num_constants,))

    @staticmethod
    def public_input():
        return Gate(PUBLIC_INPUT)

    @staticmethod
    def arithmetic(cfg):
        return Gate(ARITHMETIC, (cfg.num_routed_wires // 4,))

    @staticmethod
    def base_sum(base, num_limbs):
        return Gate(BASE_SUM, (base, num_limbs))

    @staticmethod
    def poseidon():
        return Gate(POSEIDON)

    @staticmethod
    def random_access(cfg, bits):
        vec = 1 << bits
        copies = min(cfg.num_routed_wires // (2 + vec), cfg.num_wires // (2 + vec + bits))
        extra = min(cfg.num_routed_wires - (2 + vec) * copies, cfg.num_constants)
        return Gate(RANDOM_ACCESS, (bits, copies, extra))

    @staticmethod
    def u32_arithmetic(cfg):   # plonky2_ecdsa/biguint/gates/arithmetic_u32.rs:40-43
        return Gate(U32_ARITHMETIC, (min(cfg.num_wires // 38, cfg.num_routed_wires // 6),))

    @staticmethod
    def u32_add_many(cfg, num_addends):   # add_many_u32.rs:43-48
        return Gate(U32_ADD_MANY, (num_addends, min(cfg.num_wires // (num_addends + 21),
                                                    cfg.num_routed_wires // (num_addends + 3))))

    @staticmethod
    def u32_subtraction(cfg):   # subtraction_u32.rs:38-42
        return Gate(U32_SUBTRACTION, (min(cfg.num_wires // 21, cfg.num_routed_wires // 5),))

    @staticmethod
    def u32_range_check(num_input_limbs):
        return Gate(U32_RANGE_CHECK, (num_input_limbs,))

    @staticmethod
    def comparison(num_bits=32, num_chunks=16):
        return Gate(COMPARISON, (num_bits, num_chunks))


def _selector_groups(gates, max_degree):
    """plonky2 gates/selectors.rs selector_polynomials: gates are sorted by (degree, id)."""
    n = len(gates)
    if gates[-1].degree + n - 1 <= max_degree:
        return [0] * n, [(0, n)]
    groups, start = [], 0
    while start < n:
        size = 0
        while start + size < n and size + gates[start + size].degree < max_degree:
            size += 1
        if size == 0:
            raise ValueError(f"gate {gates[start].id} does not fit any selector group")
        groups.append((start, start + size))
        start += size
    index = [gi for gi, (lo, hi) in enumerate(groups) for _ in range(lo, hi)]
    return index, groups


class CommonCircuitData:
    """plonky2 `CommonCircuitData`: everything about the circuit that is independent of the witness."""

    def __init__(self, config, degree_bits, gates, num_public_inputs=0):
        self.config = config
        self.degree_bits_ = degree_bits
        self.gates = sorted(set(gates), key=lambda g: (g.degree, g.id))
        self.num_public_inputs = num_public_inputs
        self.quotient_degree_factor = config.max_quotient_degree_factor
        self.selector_indices, self.groups = _selector_groups(self.gates, self.quotient_degree_factor + 1)
        self.num_selectors = len(self.groups)
        self.num_constants = self.num_selectors + max(g.num_constants for g in self.gates)
        self.num_gate_constraints = max(g.num_constraints for g in self.gates)
        r, q = config.num_routed_wires, self.quotient_degree_factor
        self.num_partial_products = _cdiv(r, q) - 1
        self.k_is = [pow(MULTIPLICATIVE_GROUP_GENERATOR, i, P) for i in range(r)]
        bits, d = [], degree_bits
        while d > config.fri_final_poly_bits and d + config.rate_bits - config.fri_arity_bits >= config.cap_height:
            bits.append(config.fri_arity_bits)
            d -= config.fri_arity_bits
        self.reduction_arity_bits = bits

    def degree_bits(self):
        return self.degree_bits_

    def degree(self):
        return 1 << self.degree_bits_

    def lde_size(self):
        return 1 << (self.degree_bits_ + self.config.rate_bits)

    @property
    def num_preprocessed(self):
        return self.num_constants + self.config.num_routed_wires

    @property
    def hash_size(self):
        return 25 if _lib.HASHER_ID[self.config.hasher] == 0 else 32

    def gate_index(self, gate):
        return self.gates.index(gate)

    def selector_value(self, gate, group):
        """Value of selector column `group` on a row occupied by `gate`."""
        i = self.gate_index(gate)
        return i if self.selector_indices[i] == group else UNUSED_SELECTOR

    def fill_desc(self, constants_sigmas, circuit_digest=None):
        cfg = self.config
        cs = np.ascontiguousarray(constants_sigmas, dtype=np.uint64)
        if cs.shape != (self.num_preprocessed, self.degree()):
            raise ValueError(f"constants_sigmas must be {(self.num_preprocessed, self.degree())}, got {cs.shape}")
        gates = (_lib.GateS * len(self.gates))()
        for i, g in enumerate(self.gates):
            gates[i].kind = g.kind
            for k in range(4):
                gates[i].params[k] = g.params[k]
            gates[i].selector_index = self.selector_indices[i]
            gates[i].group_lo, gates[i].group_hi = self.groups[self.selector_indices[i]]
            gates[i].num_constraints = g.num_constraints
        k_is = np.array(self.k_is, dtype=np.uint64)
        d = _lib.DescS()
        d.struct_size = C.sizeof(_lib.DescS)
        d.degree_bits = self.degree_bits_
        d.num_wires = cfg.num_wires
        d.num_routed_wires = cfg.num_routed_wires
        d.num_constants = self.num_constants
        d.num_selectors = self.num_selectors
        d.num_challenges = cfg.num_challenges
        d.rate_bits = cfg.rate_bits
        d.cap_height = cfg.cap_height
        d.pow_bits = cfg.proof_of_work_bits
        d.num_query_rounds = cfg.num_query_rounds
        d.quotient_degree_factor = self.quotient_degree_factor
        d.num_partial_products = self.num_partial_products
        d.num_gate_constraints = self.num_gate_constraints
        d.num_public_inputs = self.num_public_inputs
        d.hasher = _lib.HASHER_ID[cfg.hasher]
        d.num_fri_layers = len(self.reduction_arity_bits)
        for i, a in enumerate(self.reduction_arity_bits):
            d.reduction_arity_bits[i] = a
        d.num_gates = len(self.gates)
        d.gates = C.cast(gates, C.POINTER(_lib.GateS))
        d.constants_sigmas = cs.ctypes.data
        d.k_is = k_is.ctypes.data
        dg = None
        if circuit_digest is not None:
            dg = C.create_string_buffer(bytes(circuit_digest), len(circuit_digest))
            d.circuit_digest = C.cast(dg, C.c_void_p).value
        return d, (gates, cs, k_is, dg)


@dataclass
class ProofWithPublicInputs:
    """Bytes in plonky2's uncompressed `ProofWithPublicInputs::to_bytes` layout, plus the device stage timings."""
    proof_bytes: bytes
    public_inputs: list
    timings: dict = field(default_factory=dict)

    def to_bytes(self):
        return self.proof_bytes


class CircuitData:
    """plonky2 `CircuitData`: `prover_only` (preprocessed polynomials, resident on the GPU) + `common`."""

    def __init__(self, common, constants_sigmas, circuit_digest=None, device=0, shard=None):
        """`shard`: a sharding.TorchDistGroup / ThreadGroup member -> this handle is one rank of a coset-sharded prover
        (every rank is created with the same arguments and must call prove() with the same witness)."""
        self.common = common
        self.device = device
        self.shard = shard
        self._h = C.c_void_p()
        desc, keep = common.fill_desc(constants_sigmas, circuit_digest)
        if shard is None or shard.world == 1:
            _lib.check(_lib.lib().p2g_circuit_create(C.byref(desc), device, C.byref(self._h)))
        elif getattr(shard, "in_library", False):   # sharding.NcclGroup: the library owns the NCCL communicator
            idb = C.create_string_buffer(shard.unique_id, len(shard.unique_id))
            _lib.check(_lib.lib().p2g_circuit_create_sharded_nccl(C.byref(desc), device, shard.rank, shard.world, idb,
                                                                  C.byref(self._h)))
        else:
            _lib.check(_lib.lib().p2g_circuit_create_sharded(C.byref(desc), device, shard.rank, shard.world,
                                                             shard.callback(), None, C.byref(self._h)))
        del keep
        hs, ncap = common.hash_size, 1 << min(common.config.cap_height, common.degree_bits_ + common.config.rate_bits)
        cap = C.create_string_buffer(ncap * hs)
        dg = C.create_string_buffer(hs)
        _lib.check(_lib.lib().p2g_circuit_cap(self._h, cap, len(cap), dg, len(dg)))
        self.constants_sigmas_cap = [cap.raw[i * hs:(i + 1) * hs] for i in range(ncap)]   # verifier_only
        self.circuit_digest = dg.raw
        self._out = None

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().p2g_circuit_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _run(self, fn, wires_ptr, public_inputs, forced_pow_witness, timings, compressed=None):
        pis = np.array([int(x) for x in public_inputs], dtype=np.uint64)
        if self._out is None:
            self._out = C.create_string_buffer(_lib.lib().p2g_proof_size_bound(self._h))
        ln = C.c_size_t(len(self._out))
        fp = C.byref(C.c_uint64(forced_pow_witness)) if forced_pow_witness is not None else None
        tm = _lib.TimingsS() if timings else None
        if compressed is not None:   # compressed = 0 / 1: wires on host / device
            rc = _lib.lib().p2g_prove_compressed(self._h, wires_ptr, compressed, pis.ctypes.data_as(C.c_void_p), len(pis), fp,
                                                 self._out, C.byref(ln), C.byref(tm) if timings else None)
        else:
            rc = fn(self._h, wires_ptr, pis.ctypes.data_as(C.c_void_p), len(pis), fp, self._out, C.byref(ln),
                    C.byref(tm) if timings else None)
        _lib.check(rc)
        return ProofWithPublicInputs(self._out.raw[:ln.value], [int(x) for x in pis], tm.as_dict() if timings else {})

    def prove(self, wires, public_inputs=(), forced_pow_witness=None, timings=True, compressed=False):
        """`wires`: the full witness matrix (`MatrixWitness.wire_values`), shape [num_wires, N], canonical u64; a numpy
        array (host; may be pinned) or a CUDA torch tensor on this circuit's device.  compressed=True returns plonky2's
        compressed layout -- what `proof.compress(..).to_bytes()` gives at prove_action.rs:77-78, the CLI's proof file."""
        shape = (self.common.config.num_wires, self.common.degree())
        if hasattr(wires, "is_cuda"):
            if tuple(wires.shape) != shape or not wires.is_contiguous() or wires.element_size() != 8:
                raise ValueError(f"wires must be a contiguous 64-bit tensor of shape {shape}")
            fn = _lib.lib().p2g_prove_device if wires.is_cuda else _lib.lib().p2g_prove
            if wires.is_cuda and wires.device.index != self.device:
                raise ValueError("wires live on another device")
            return self._run(fn, C.c_void_p(wires.data_ptr()), public_inputs, forced_pow_witness, timings,
                             (1 if wires.is_cuda else 0) if compressed else None)
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        if w.shape != shape:
            raise ValueError(f"wires must have shape {shape}, got {w.shape}")
        return self._run(_lib.lib().p2g_prove, w.ctypes.data_as(C.c_void_p), public_inputs, forced_pow_witness, timings,
                         0 if compressed else None)

    def fill_advice(self, wires):
        """Device-side witness fill (p2g_fill_advice_device): `wires` is a CUDA tensor [num_wires, N] on this circuit's device
        whose routed columns (the first num_routed_wires) hold the witness; the advice columns are computed in place."""
        shape = (self.common.config.num_wires, self.common.degree())
        if not getattr(wires, "is_cuda", False) or tuple(wires.shape) != shape or not wires.is_contiguous() or wires.element_size() != 8:
            raise ValueError(f"wires must be a contiguous 64-bit CUDA tensor of shape {shape}")
        if wires.device.index != self.device:
            raise ValueError("wires live on another device")
        _lib.check(_lib.lib().p2g_fill_advice_device(self._h, C.c_void_p(wires.data_ptr())))
        return wires

    def verifier_data_bytes(self):
        """`verifier_data().to_bytes(&BackendGateSerializer)`: the file `write_vk` writes (write_vk_action.rs:76-79), from this
        handle (the CircuitConfig fields the prover does not read default to wide_ecc_config)."""
        ln = C.c_size_t(0)
        rc = _lib.lib().p2g_vk_bytes(self._h, None, None, C.byref(ln))
        if rc != _lib.P2G_ESMALLBUF:
            _lib.check(rc)
        out = C.create_string_buffer(ln.value)
        _lib.check(_lib.lib().p2g_vk_bytes(self._h, None, out, C.byref(ln)))
        return out.raw[:ln.value]

    def prove_columns(self, wire_columns, public_inputs=(), forced_pow_witness=None, timings=True, compressed=False):
        """`wire_columns`: num_wires separate 1-D uint64 arrays of N values each -- `MatrixWitness.wire_values` as plonky2 holds
        it (`Vec<Vec<F>>`, one allocation per column), handed over as column pointers without building a flat copy."""
        W, n = self.common.config.num_wires, self.common.degree()
        if len(wire_columns) != W:
            raise ValueError(f"expected {W} wire columns, got {len(wire_columns)}")
        cols = []
        for c in wire_columns:
            a = np.ascontiguousarray(c, dtype=np.uint64)
            if a.shape != (n,):
                raise ValueError(f"every wire column must have shape ({n},), got {a.shape}")
            cols.append(a)
        ptrs = (C.c_void_p * W)(*[a.ctypes.data for a in cols])
        fn = lambda h, _w, pis, npi, fp, out, ln, tm: _lib.lib().p2g_prove_columns(h, ptrs, pis, npi, fp, 1 if compressed else 0,
                                                                                   out, ln, tm)
        return self._run(fn, None, public_inputs, forced_pow_witness, timings, None)

    def prove_routed_columns(self, routed_columns, public_inputs=(), forced_pow_witness=None, timings=True, compressed=False):
        """prove_columns with only the ROUTED wire columns (the first num_routed_wires of `MatrixWitness.wire_values`): the advice
        columns -- two thirds of the trace -- are computed on the device (p2g_prove_routed_columns)."""
        R, n = self.common.config.num_routed_wires, self.common.degree()
        if len(routed_columns) != R:
            raise ValueError(f"expected {R} routed wire columns, got {len(routed_columns)}")
        cols = []
        for c in routed_columns:
            a = np.ascontiguousarray(c, dtype=np.uint64)
            if a.shape != (n,):
                raise ValueError(f"every wire column must have shape ({n},), got {a.shape}")
            cols.append(a)
        ptrs = (C.c_void_p * R)(*[a.ctypes.data for a in cols])
        fn = lambda h, _w, pis, npi, fp, out, ln, tm: _lib.lib().p2g_prove_routed_columns(h, ptrs, pis, npi, fp, 1 if compressed else 0,
                                                                                          out, ln, tm)
        return self._run(fn, None, public_inputs, forced_pow_witness, timings, None)

    def read(self, what, dtype=np.uint64):
        """Intermediates of the last proof (enum p2g_buffer), for parity tests."""
        ln = C.c_size_t(0)
        _lib.lib().p2g_circuit_read(self._h, what, None, C.byref(ln))
        buf = np.empty(max(ln.value, 1), dtype=np.uint8)
        _lib.check(_lib.lib().p2g_circuit_read(self._h, what, buf.ctypes.data_as(C.c_void_p), C.byref(ln)))
        buf = buf[:ln.value]
        return buf.view(dtype) if dtype != np.uint8 else buf


def eval_gate_constraints(common, constants, wires, pi_hash=(0, 0, 0, 0), device=0):
    """`evaluate_gate_constraints_base_batch` at arbitrary points: constants [num_constants, n], wires [num_wires, n]."""
    cs = np.ascontiguousarray(constants, dtype=np.uint64)
    w = np.ascontiguousarray(wires, dtype=np.uint64)
    npts = w.shape[1]
    d, keep = common.fill_desc(np.zeros((common.num_preprocessed, common.degree()), dtype=np.uint64))
    pi = np.array([int(x) for x in pi_hash], dtype=np.uint64)
    out = np.empty((common.num_gate_constraints, npts), dtype=np.uint64)
    _lib.check(_lib.lib().p2g_eval_gate_constraints(C.byref(d), cs.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                                                    pi.ctypes.data_as(C.c_void_p), npts, out.ctypes.data_as(C.c_void_p),
                                                    device))
    del keep
    return out
