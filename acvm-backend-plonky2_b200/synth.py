"""Synthetic ACIR-shaped circuits: the FFI payload (constants, sigmas, gate table, witness matrix) for the BASELINE configs.

Stand-in for the reference's Rust layers that produce this payload (circuit_translation/*.rs + plonky2's witness
generators), which cannot run in this environment.  The gate mixes follow the shapes the translators emit
(SURVEY.md section 8d / App. E): AssertZero -> ArithmeticGate rows, RANGE -> BaseSum<2>, SHA-256 -> ~97% Arithmetic + BaseSum,
EcdsaSecp256k1 -> the u32 / comparison / random-access gates of plonky2_ecdsa.  Everything is seeded and reproducible.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import lib as _lib
from .circuit import CircuitConfig, CommonCircuitData, Gate

_HERE = os.path.dirname(os.path.abspath(__file__))
_SYNTH = None


class _SpecS(C.Structure):
    _fields_ = [("degree_bits", C.c_uint32), ("num_wires", C.c_uint32), ("num_routed_wires", C.c_uint32),
                ("num_constants", C.c_uint32), ("num_selectors", C.c_uint32), ("num_gates", C.c_uint32),
                ("num_public_inputs", C.c_uint32), ("tie_permille", C.c_uint32), ("seed", C.c_uint64),
                ("gates", C.POINTER(_lib.GateS)), ("row_gate", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libp2synth.so")
    srcs = [os.path.join(_HERE, "synth", f) for f in os.listdir(os.path.join(_HERE, "synth")) if f.endswith((".cpp", ".h"))]
    stale = lambda: force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in srcs)   # noqa: E731
    if stale():
        with _lib.build_lock():          # ranks of one box may get here together: one builds, the others wait and re-check
            if stale():
                subprocess.check_call(["make", "-C", os.path.join(_HERE, "synth"), "-s"] + (["-B"] if force else []))
    return so


def _synth_lib():
    global _SYNTH
    if _SYNTH is None:
        _SYNTH = C.CDLL(build())
    return _SYNTH


def set_threads(n):
    """OpenMP threads of the generator (torchrun pins OMP_NUM_THREADS=1 for its workers)."""
    _synth_lib().p2s_set_threads(int(n))


def gate_mix(name, cfg):
    """[(Gate, weight)] for a named workload; weights are row fractions of the non-padding rows."""
    A = Gate.arithmetic(cfg)
    K = Gate.constant(cfg)
    if name == "assert_zero":          # BASELINE configs[1]: synthetic AssertZero chain
        return [(A, 0.97), (K, 0.03)]
    if name == "sha256":               # configs[2]: bit operations -> ArithmeticGate, split_le/le_sum -> BaseSum<2>
        return [(A, 0.955), (Gate.base_sum(2, 63), 0.03), (K, 0.015)]
    if name == "ecdsa":                # configs[3]: EcdsaSecp256k1 (nonnative + GLV + comparison), the headline workload
        mix = [(A, 0.30), (Gate.base_sum(2, 63), 0.05), (Gate.base_sum(4, 16), 0.05), (Gate.random_access(cfg, 4), 0.05),
               (Gate.u32_arithmetic(cfg), 0.25), (Gate.u32_subtraction(cfg), 0.05), (Gate.u32_range_check(8), 0.04),
               (Gate.u32_range_check(1), 0.01), (Gate.comparison(32, 16), 0.05)]
        for na, wgt in [(2, 0.05), (3, 0.03), (4, 0.02), (8, 0.03), (16, 0.02)]:
            mix.append((Gate.u32_add_many(cfg, na), wgt))
        return mix
    if name == "range":                # configs[4]: AssertZero + RANGE
        return [(A, 0.79), (Gate.base_sum(2, 63), 0.20), (K, 0.01)]
    if name == "all_gates":            # every gate kind, small shapes (tests)
        return [(A, 0.2), (K, 0.05), (Gate.base_sum(2, 63), 0.05), (Gate.base_sum(4, 16), 0.05),
                (Gate.random_access(cfg, 4), 0.07), (Gate.random_access(cfg, 2), 0.05), (Gate.u32_arithmetic(cfg), 0.1),
                (Gate.u32_add_many(cfg, 2), 0.07), (Gate.u32_add_many(cfg, 5), 0.05), (Gate.u32_subtraction(cfg), 0.08),
                (Gate.u32_range_check(8), 0.05), (Gate.u32_range_check(2), 0.03), (Gate.comparison(32, 16), 0.08),
                (Gate.poseidon(), 0.07)]
    raise KeyError(name)


class SyntheticCircuit:
    """common (CommonCircuitData) + constants_sigmas [P, N] + wires [W, N] + public_inputs."""

    def __init__(self, degree_bits, workload="assert_zero", config=None, num_public_inputs=0, seed=0xAC1D, tie_permille=150,
                 fill=0.98, mix=None, pinned=False):
        cfg = config or CircuitConfig.wide_ecc_config()
        self.config = cfg
        self.workload = workload
        n = 1 << degree_bits
        mix = mix if mix is not None else gate_mix(workload, cfg)
        gates = [g for g, _ in mix] + [Gate.noop()]
        if num_public_inputs:
            gates += [Gate.public_input(), Gate.poseidon()]
        self.common = CommonCircuitData(cfg, degree_bits, gates, num_public_inputs)
        com = self.common
        # row assignment: PublicInput row + its Poseidon rows first, then the mix in blocks (like a builder would emit
        # gate instances), Noop padding up to 2^degree_bits
        rows = []
        if num_public_inputs:
            rows += [com.gate_index(Gate.public_input())] + [com.gate_index(Gate.poseidon())] * ((num_public_inputs + 7) // 8)
        budget = max(0, int(n * fill) - len(rows))
        total_w = sum(w for _, w in mix)
        rng = np.random.default_rng(seed)
        body = []
        for g, w in mix:
            body += [com.gate_index(g)] * int(round(budget * w / total_w))
        body = np.array(body[:max(0, n - len(rows))], dtype=np.uint8)
        # interleave in runs (the builder emits runs of the same gate), keep it deterministic
        if len(body):
            run = 16
            pad = (-len(body)) % run
            b = np.concatenate([body, np.full(pad, 255, dtype=np.uint8)]).reshape(-1, run)
            b = b[rng.permutation(b.shape[0])].reshape(-1)
            body = b[b != 255]
        row_gate = np.full(n, com.gate_index(Gate.noop()), dtype=np.uint8)
        k = len(rows)
        row_gate[:k] = rows
        row_gate[k:k + len(body)] = body[:n - k]
        self.row_gate = row_gate
        P, W = com.num_preprocessed, cfg.num_wires
        if pinned:
            import torch
            self._wires_t = torch.empty((W, n), dtype=torch.int64).pin_memory()
            self.wires = self._wires_t.numpy().view(np.uint64)
        else:
            self.wires = np.empty((W, n), dtype=np.uint64)
        self.constants_sigmas = np.empty((P, n), dtype=np.uint64)
        pis = np.zeros(max(1, num_public_inputs), dtype=np.uint64)
        gates_c = (_lib.GateS * len(com.gates))()
        for i, g in enumerate(com.gates):
            gates_c[i].kind = g.kind
            for j in range(4):
                gates_c[i].params[j] = g.params[j]
            gates_c[i].selector_index = com.selector_indices[i]
            gates_c[i].group_lo, gates_c[i].group_hi = com.groups[com.selector_indices[i]]
            gates_c[i].num_constraints = g.num_constraints
        spec = _SpecS(degree_bits, W, cfg.num_routed_wires, com.num_constants, com.num_selectors, len(com.gates),
                      num_public_inputs, tie_permille, seed, C.cast(gates_c, C.POINTER(_lib.GateS)), row_gate.ctypes.data)
        rc = _synth_lib().p2s_synthesize(C.byref(spec), self.constants_sigmas.ctypes.data_as(C.c_void_p),
                                         self.wires.ctypes.data_as(C.c_void_p), pis.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"p2s_synthesize failed: {rc}")
        self.public_inputs = [int(x) for x in pis[:num_public_inputs]]

    def gate_histogram(self):
        idx, cnt = np.unique(self.row_gate, return_counts=True)
        return {self.common.gates[i].id.split(" ")[0].split("(")[0] + str(list(self.common.gates[i].params[:2])): int(c)
                for i, c in zip(idx, cnt)}


# BASELINE.json configs -> (degree_bits, workload, public inputs)
BASELINE_CONFIGS = {
    "fibonacci": dict(degree_bits=3, workload="assert_zero", num_public_inputs=0),
    "assert_zero_2^16": dict(degree_bits=16, workload="assert_zero", num_public_inputs=0),
    "sha256_2^18": dict(degree_bits=18, workload="sha256", num_public_inputs=4),
    "ecdsa_2^20": dict(degree_bits=20, workload="ecdsa", num_public_inputs=4),
    "range_2^22": dict(degree_bits=22, workload="range", num_public_inputs=0),
}
