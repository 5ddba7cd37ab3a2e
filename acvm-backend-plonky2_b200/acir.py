"""ACIR circuits -> the payload of include/p2g.h, without the Rust layers.

Mirror of the reference's `CircuitBuilderFromAcirToPlonky2` (plonky2-backend/src/circuit_translation/mod.rs:55-190): the same
names -- `translate_circuit`, `unpack`, `witness_target_map` semantics -- over libp2acir.so (C++: the translator, the slice of
plonky2's CircuitBuilder it drives, and the witness generators).  ACIR values are the ones the reference's test factories build by
hand (circuit_translation/tests/factories/circuit_factory.rs): `Expression(mul_terms, linear_combinations, q_c)`, `AssertZero`,
`BlackBoxFuncCall::{RANGE, AND, XOR, Sha256Compression}`, `MemoryInit`, `MemoryOp` (read and write).  Field elements are Goldilocks (mod.rs:43-45).
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

from . import lib as _lib
from .circuit import CircuitConfig, CircuitData, CommonCircuitData, Gate, P

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libp2acir.so")
    srcs = [os.path.join(_HERE, "acir", f) for f in os.listdir(os.path.join(_HERE, "acir")) if f.endswith((".cpp", ".h"))]
    stale = lambda: force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs)   # noqa: E731
    if stale():
        with _lib.build_lock():          # ranks of one box may get here together: one builds, the others wait and re-check
            if stale():
                subprocess.check_call(["make", "-C", os.path.join(_HERE, "acir"), "-s"] + (["-B"] if force else []))
    return so


def _acir_lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.p2a_last_error.restype = C.c_char_p
        L.p2a_translate.restype = C.c_void_p
        L.p2a_translate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.p2a_destroy.argtypes = [C.c_void_p]
        L.p2a_shape.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.p2a_gate_types.argtypes = [C.c_void_p, C.c_void_p]
        L.p2a_constants_sigmas.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.p2a_witness.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.p2a_witness_routed.argtypes = L.p2a_witness.argtypes
        L.p2a_read_witnesses.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.p2a_set_threads.argtypes = [C.c_int]
        L.p2a_fill_advice.argtypes = [C.c_void_p, C.c_void_p]
        L.p2a_rows_used.argtypes = [C.c_void_p]
        L.p2a_rows_used.restype = C.c_uint32
        _LIB = L
    return _LIB


def set_threads(n):
    """Worker threads of witness generation (generator groups of heavy opcodes run side by side); 0 = all cores."""
    _acir_lib().p2a_set_threads(int(n))


class TranslationError(Exception):
    """Where the reference panics (unsupported opcode, RANGE over 33 bits, unsatisfiable witness)."""


# ---- ACIR values (acir::native_types / acir::circuit, the subset the translators accept) -----------------------------------
@dataclass
class Expression:
    mul_terms: list = field(default_factory=list)            # [(coefficient, witness, witness)]
    linear_combinations: list = field(default_factory=list)  # [(coefficient, witness)]
    q_c: int = 0


@dataclass
class AssertZero:
    expr: Expression


@dataclass
class Range:          # BlackBoxFuncCall::RANGE { input: FunctionInput { witness, num_bits } }
    witness: int
    num_bits: int


@dataclass
class And:            # BlackBoxFuncCall::AND { lhs, rhs, output }
    lhs: int
    rhs: int
    num_bits: int
    output: int


@dataclass
class Xor:
    lhs: int
    rhs: int
    num_bits: int
    output: int


@dataclass
class MemoryInit:
    block_id: int
    init: list


@dataclass
class MemoryRead:     # MemoryOp { operation = 0, index, value }
    block_id: int
    index: int
    value: int


@dataclass
class MemoryWrite:    # MemoryOp { operation = 1, index, value }
    block_id: int
    index: int
    value: int


@dataclass
class Sha256Compression:   # BlackBoxFuncCall::Sha256Compression { inputs: [_; 16], hash_values: [_; 8], outputs: [_; 8] }
    inputs: list
    hash_values: list
    outputs: list


@dataclass
class EcdsaSecp256k1:   # BlackBoxFuncCall::EcdsaSecp256k1 { public_key_x: [_; 32], public_key_y: [_; 32], signature: [_; 64], hashed_message: [_; 32], output }
    public_key_x: list
    public_key_y: list
    signature: list
    hashed_message: list
    output: int


# ---- gadget-level operations (not ACIR opcodes): the reference's u32 gadgets, plonky2_ecdsa/biguint/gadgets/*.rs, over its custom
# gates.  The reference reaches them only through its EcdsaSecp256k1 translator; here a circuit can be built on them directly, the
# way the reference's gadget tests do.  Witness indices name the targets; outputs are computed by the gates' generators.
@dataclass
class MulAddU32:      # x * y + z = low + 2^32 high           (U32ArithmeticGate)
    x: int
    y: int
    z: int
    low: int
    high: int


@dataclass
class AddManyU32:     # sum(addends) = result + 2^32 carry     (U32AddManyGate; 2 addends -> U32ArithmeticGate)
    addends: list
    result: int
    carry: int


@dataclass
class SubU32:         # x - y - borrow = result - 2^32 borrow_out   (U32SubtractionGate)
    x: int
    y: int
    borrow: int
    result: int
    borrow_out: int


@dataclass
class RangeCheckU32:  # every value < 2^32                      (U32RangeCheckGate)
    values: list


@dataclass
class CmpLe:          # result = (a <= b) on num_bits-bit values  (ComparisonGate, 2-bit chunks)
    a: int
    b: int
    num_bits: int
    result: int


@dataclass
class BrilligCall:    # Opcode::BrilligCall { .. }: unconstrained code, "ignored since it has no impact in the circuit" (mod.rs:97-103)
    id: int = 0


@dataclass
class Directive:      # Opcode::Directive(_): ignored as well (mod.rs:104)
    name: str = ""


GADGET_KINDS = {"add_biguint": 1, "sub_biguint": 2, "mul_biguint": 3, "cmp_biguint": 4, "div_rem_biguint": 5, "add_nonnative": 6,
                "sub_nonnative": 7, "mul_nonnative": 8, "neg_nonnative": 9, "inv_nonnative": 10, "add_many_nonnative": 11, "list_le": 12,
                "glv_mul": 13, "curve_add": 14, "curve_double": 15}
BASE_FIELD, SCALAR_FIELD = 0, 1


@dataclass
class Gadget:         # a plonky2_ecdsa gadget on u32-limb witnesses (least significant limb first); see p2acir.cpp opcode 110
    name: str         # one of GADGET_KINDS
    lists: list       # operand and result witness-id lists, in the gadget's order
    param: int = 0    # non-native gadgets: BASE_FIELD / SCALAR_FIELD; list_le: bits per element


@dataclass
class Circuit:
    opcodes: list
    public_parameters: list = field(default_factory=list)
    private_parameters: list = field(default_factory=list)


def _encode(circuit):
    words = []
    for op in circuit.opcodes:
        if isinstance(op, AssertZero):
            e = op.expr
            words += [1, len(e.mul_terms), len(e.linear_combinations), e.q_c % P]
            for c, a, b in e.mul_terms:
                words += [c % P, a, b]
            for c, w in e.linear_combinations:
                words += [c % P, w]
        elif isinstance(op, Range):
            words += [2, op.witness, op.num_bits]
        elif isinstance(op, And):
            words += [3, op.lhs, op.rhs, op.num_bits, op.output]
        elif isinstance(op, Xor):
            words += [4, op.lhs, op.rhs, op.num_bits, op.output]
        elif isinstance(op, MemoryInit):
            words += [5, op.block_id, len(op.init)] + list(op.init)
        elif isinstance(op, MemoryRead):
            words += [6, op.block_id, op.index, op.value]
        elif isinstance(op, MemoryWrite):
            words += [8, op.block_id, op.index, op.value]
        elif isinstance(op, (BrilligCall, Directive)):
            continue
        elif isinstance(op, Gadget):
            words += [110, GADGET_KINDS[op.name], op.param, len(op.lists)]
            for l in op.lists:
                words += [len(l)] + list(l)
        elif isinstance(op, EcdsaSecp256k1):
            if (len(op.public_key_x), len(op.public_key_y), len(op.signature), len(op.hashed_message)) != (32, 32, 64, 32):
                raise TranslationError("EcdsaSecp256k1 takes 32 + 32 + 64 + 32 byte witnesses")
            words += [9] + list(op.public_key_x) + list(op.public_key_y) + list(op.signature) + list(op.hashed_message) + [op.output]
        elif isinstance(op, MulAddU32):
            words += [101, op.x, op.y, op.z, op.low, op.high]
        elif isinstance(op, AddManyU32):
            words += [102, len(op.addends)] + list(op.addends) + [op.result, op.carry]
        elif isinstance(op, SubU32):
            words += [103, op.x, op.y, op.borrow, op.result, op.borrow_out]
        elif isinstance(op, RangeCheckU32):
            words += [104, len(op.values)] + list(op.values)
        elif isinstance(op, CmpLe):
            words += [105, op.a, op.b, op.num_bits, op.result]
        elif isinstance(op, Sha256Compression):
            if (len(op.inputs), len(op.hash_values), len(op.outputs)) != (16, 8, 8):
                raise TranslationError("Sha256Compression takes 16 inputs, 8 hash values and 8 outputs")
            words += [7] + list(op.inputs) + list(op.hash_values) + list(op.outputs)
        else:
            raise TranslationError(f"Opcode not supported yet: {op!r}")
    return np.array(words, dtype=np.uint64)


_ARITY = {0: 0, 1: 1, 2: 0, 3: 1, 4: 2, 5: 0, 6: 3}


class CircuitBuilderFromAcirToPlonky2:
    """translate_circuit(circuit) then unpack() -> (CircuitData inputs, witness generator), like mod.rs:72-83."""

    def __init__(self, config=None):
        self.config = config or CircuitConfig.wide_ecc_config()
        if (self.config.num_wires, self.config.num_routed_wires, self.config.num_constants) != (234, 80, 2):
            raise ValueError("the translator builds with CircuitConfig::wide_ecc_config() (circuit_translation/mod.rs:69)")
        self._h = None

    def translate_circuit(self, circuit):
        L = _acir_lib()
        pub = np.array(sorted(circuit.public_parameters), dtype=np.uint64)       # BTreeSet order (mod.rs:291-296)
        priv = np.array(sorted(circuit.private_parameters), dtype=np.uint64)
        ops = _encode(circuit)
        h = L.p2a_translate(pub.ctypes.data_as(C.c_void_p), len(pub), priv.ctypes.data_as(C.c_void_p), len(priv),
                            ops.ctypes.data_as(C.c_void_p), len(ops))
        if not h:
            raise TranslationError(L.p2a_last_error().decode())
        self._h = C.c_void_p(h)
        db, ng, npub = C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.p2a_shape(self._h, C.byref(db), C.byref(ng), C.byref(npub))
        raw = np.zeros(5 * ng.value, dtype=np.uint32)
        L.p2a_gate_types(self._h, raw.ctypes.data_as(C.c_void_p))
        types = [Gate(int(raw[5 * i]), tuple(int(x) for x in raw[5 * i + 1:5 * i + 5])) for i in range(ng.value)]
        self.common = CommonCircuitData(self.config, db.value, types, npub.value)
        com = self.common
        # preprocessed polynomials for the sorted gate table
        table = (_lib.GateS * len(com.gates))()
        for i, g in enumerate(com.gates):
            table[i].kind = g.kind
            for k in range(4):
                table[i].params[k] = g.params[k]
            table[i].selector_index = com.selector_indices[i]
            table[i].group_lo, table[i].group_hi = com.groups[com.selector_indices[i]]
            table[i].num_constraints = g.num_constraints
        t2g = np.array([com.gate_index(t) for t in types], dtype=np.uint32)
        k_is = np.array(com.k_is, dtype=np.uint64)
        self.constants_sigmas = np.zeros((com.num_preprocessed, com.degree()), dtype=np.uint64)
        rc = L.p2a_constants_sigmas(self._h, table, len(com.gates), t2g.ctypes.data_as(C.c_void_p), com.num_selectors, com.num_constants,
                                    k_is.ctypes.data_as(C.c_void_p), self.constants_sigmas.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise TranslationError(L.p2a_last_error().decode())
        return self

    def generate_witness(self, witness_map, routed_only=False):
        """plonky2 `generate_partial_witness(..).full_witness()` for an ACIR witness map {witness index: value} (what
        prove_action.rs:99-117 feeds in): returns (wires [234, N], public_inputs).  routed_only=True returns just the routed
        columns [80, N] -- the generators skip the advice wires, which CircuitData.prove_routed_columns computes on the device."""
        L = _acir_lib()
        ids = np.array(list(witness_map.keys()), dtype=np.uint64)
        vals = np.array([int(v) % P for v in witness_map.values()], dtype=np.uint64)
        ncols = self.config.num_routed_wires if routed_only else self.config.num_wires
        wires = np.zeros((ncols, self.common.degree()), dtype=np.uint64)
        pis = np.zeros(max(1, self.common.num_public_inputs), dtype=np.uint64)
        fn = L.p2a_witness_routed if routed_only else L.p2a_witness
        rc = fn(self._h, ids.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p), len(ids),
                wires.ctypes.data_as(C.c_void_p), pis.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise TranslationError(L.p2a_last_error().decode())
        return wires, [int(x) for x in pis[:self.common.num_public_inputs]]

    def read_witnesses(self, ids):
        """{witness index: value} as the last generate_witness() left them -- computed outputs included (the reference reads them
        through its witness_target_map); witnesses the circuit never mentions are omitted."""
        L = _acir_lib()
        a = np.array(list(ids), dtype=np.uint64)
        vals, known = np.zeros(len(a), dtype=np.uint64), np.zeros(len(a), dtype=np.uint8)
        L.p2a_read_witnesses(self._h, a.ctypes.data_as(C.c_void_p), len(a), vals.ctypes.data_as(C.c_void_p), known.ctypes.data_as(C.c_void_p))
        return {int(k): int(v) for k, v, ok in zip(a, vals, known) if ok}

    def fill_advice_host(self, wires):
        """The host twin of CircuitData.fill_advice (csrc/advice.cuh compiled for the CPU): recomputes the advice columns of a
        [234, N] uint64 matrix in place from its routed columns.  Test infrastructure."""
        assert wires.dtype == np.uint64 and wires.flags.c_contiguous and wires.shape == (self.config.num_wires, self.common.degree())
        _acir_lib().p2a_fill_advice(self._h, wires.ctypes.data_as(C.c_void_p))
        return wires

    def rows_used(self):
        """Rows in use before the power-of-two padding."""
        return int(_acir_lib().p2a_rows_used(self._h))

    def unpack(self, device=0):
        """(CircuitData on the GPU, self): the analogue of `translator.unpack()` -> (circuit_data, witness_target_map)."""
        return CircuitData(self.common, self.constants_sigmas, device=device), self

    def close(self):
        if self._h:
            _acir_lib().p2a_destroy(self._h)
            self._h = None

    __del__ = close
