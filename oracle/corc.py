"""ctypes binding of oracle/liborc.so (the C restatement of the plonky2 prover).  ORACLE = test infrastructure:
importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

(BUF_WIRES_CAP, BUF_ZS_PP_CAP, BUF_QUOTIENT_CAP, BUF_CS_CAP, BUF_ZS_PP_VALUES, BUF_QUOTIENT_CHUNKS, BUF_WIRES_COEFFS,
 BUF_CHALLENGES, BUF_FINAL_POLY, BUF_FRI_CAPS, BUF_WIRES_LDE) = range(11)


class GateS(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("params", C.c_uint32 * 4), ("selector_index", C.c_uint32),
                ("group_lo", C.c_uint32), ("group_hi", C.c_uint32), ("num_constraints", C.c_uint32)]


class DescS(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("degree_bits", C.c_uint32), ("num_wires", C.c_uint32),
                ("num_routed_wires", C.c_uint32), ("num_constants", C.c_uint32), ("num_selectors", C.c_uint32),
                ("num_challenges", C.c_uint32), ("rate_bits", C.c_uint32), ("cap_height", C.c_uint32),
                ("pow_bits", C.c_uint32), ("num_query_rounds", C.c_uint32), ("quotient_degree_factor", C.c_uint32),
                ("num_partial_products", C.c_uint32), ("num_gate_constraints", C.c_uint32),
                ("num_public_inputs", C.c_uint32), ("hasher", C.c_uint32), ("num_fri_layers", C.c_uint32),
                ("reduction_arity_bits", C.c_uint32 * 8), ("num_gates", C.c_uint32), ("gates", C.POINTER(GateS)),
                ("constants_sigmas", C.c_void_p), ("k_is", C.c_void_p), ("circuit_digest", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c"))]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


_NATIVE_NOTE = "generic x86-64 build (-O3)"


def build_native():
    """CPU-baseline build: the same sources with -march=native, compiled ON THE MACHINE THAT RUNS IT (the .so is keyed by the
    host's CPU flags, so a library built in the development container is never executed on a different CPU).  Returns the
    path, or None when no compiler is available (the portable build is used instead and the bench line says so)."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    key = hashlib.sha256(flags.encode()).hexdigest()[:12]
    out_dir = os.path.join(_HERE, "_native")
    so = os.path.join(out_dir, f"liborc_native_{key}.so")
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c"))]
    if os.path.exists(so) and all(os.path.getmtime(s) <= os.path.getmtime(so) for s in srcs):
        return so
    try:
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["make", "-C", _HERE, "-s", "native", f"NATIVE_OUT={so}"])
        return so
    except Exception:
        return None


def lib(native=False):
    """native=True (bench.py's CPU arms only, before any other use in the process): load the -march=native build."""
    global _LIB, _NATIVE_NOTE
    if _LIB is None:
        path = None
        if native:
            path = build_native()
            if path:
                _NATIVE_NOTE = "built on this host with gcc -O3 -march=native -funroll-loops, OpenMP"
        _LIB = C.CDLL(path or build())
        _LIB.orc_last_error.restype = C.c_char_p
        _LIB.orc_last_error.argtypes = [C.c_void_p]
        _LIB.orc_destroy.argtypes = [C.c_void_p]
        _LIB.orc_destroy.restype = None
    return _LIB


HASHER_ID = {"keccak25": 0, "poseidon": 1, 0: 0, 1: 1}


def make_desc(cd, constants_sigmas, digest=None):
    """cd: oracle.pyref.circuit.CommonData; constants_sigmas: uint64 [num_constants+num_routed, N] values."""
    cs = np.ascontiguousarray(constants_sigmas, dtype=np.uint64)
    assert cs.shape == (cd.num_preprocessed, cd.n), (cs.shape, cd.num_preprocessed, cd.n)
    gates = (GateS * len(cd.gates))()
    for i, g in enumerate(cd.gates):
        gates[i].kind = g.kind
        for k in range(4):
            gates[i].params[k] = g.params[k]
        gates[i].selector_index = cd.selector_indices[i]
        gates[i].group_lo, gates[i].group_hi = cd.groups[cd.selector_indices[i]]
        gates[i].num_constraints = g.num_constraints
    k_is = np.array(cd.k_is, dtype=np.uint64)
    d = DescS()
    d.struct_size = C.sizeof(DescS)
    d.degree_bits = cd.degree_bits
    d.num_wires = cd.num_wires
    d.num_routed_wires = cd.num_routed
    d.num_constants = cd.num_constants
    d.num_selectors = cd.num_selectors
    d.num_challenges = cd.num_challenges
    d.rate_bits = cd.rate_bits
    d.cap_height = cd.cap_height
    d.pow_bits = cd.pow_bits
    d.num_query_rounds = cd.num_queries
    d.quotient_degree_factor = cd.qdf
    d.num_partial_products = cd.num_partial_products
    d.num_gate_constraints = cd.num_gate_constraints
    d.num_public_inputs = cd.num_public_inputs
    d.hasher = HASHER_ID[cd.hasher]
    d.num_fri_layers = len(cd.arity_bits)
    for i, a in enumerate(cd.arity_bits):
        d.reduction_arity_bits[i] = a
    d.num_gates = len(cd.gates)
    d.gates = C.cast(gates, C.POINTER(GateS))
    d.constants_sigmas = cs.ctypes.data
    d.k_is = k_is.ctypes.data
    dg = None
    if digest is not None:
        dg = C.create_string_buffer(bytes(digest), len(digest))
        d.circuit_digest = C.cast(dg, C.c_void_p).value
    keep = (gates, cs, k_is, dg)
    return d, keep


def _u64p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleProver:
    def __init__(self, cd, constants_sigmas, digest=None):
        self.cd = cd
        self.L = lib()
        d, keep = make_desc(cd, constants_sigmas, digest)
        self._keep = keep
        self.h = C.c_void_p()
        rc = self.L.orc_create(C.byref(d), C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"orc_create failed: {rc}")
        self.hs = 25 if HASHER_ID[cd.hasher] == 0 else 32

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cap_and_digest(self):
        cap = C.create_string_buffer((1 << self.cd.cap_height) * self.hs)
        dg = C.create_string_buffer(self.hs)
        self.L.orc_cap(self.h, cap, dg)
        return [cap.raw[i * self.hs:(i + 1) * self.hs] for i in range(1 << self.cd.cap_height)], dg.raw

    def prove(self, wires, public_inputs=(), forced_pow=None):
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        assert w.shape == (self.cd.num_wires, self.cd.n)
        pis = np.array(list(public_inputs), dtype=np.uint64)
        cap = 1 << 22
        while True:
            out = C.create_string_buffer(cap)
            ln = C.c_size_t(cap)
            fp = C.byref(C.c_uint64(forced_pow)) if forced_pow is not None else None
            rc = self.L.orc_prove(self.h, _u64p(w), _u64p(pis), C.c_size_t(len(pis)), fp, out, C.byref(ln))
            if rc == -6:
                cap = ln.value
                continue
            if rc != 0:
                raise RuntimeError(f"orc_prove failed: {rc}: {self.L.orc_last_error(self.h).decode()}")
            return out.raw[:ln.value]

    def read(self, what, dtype=np.uint64):
        ln = C.c_size_t(0)
        self.L.orc_read(self.h, what, None, C.byref(ln))
        buf = np.empty(ln.value, dtype=np.uint8)
        rc = self.L.orc_read(self.h, what, buf.ctypes.data_as(C.c_void_p), C.byref(ln))
        if rc != 0:
            raise RuntimeError(f"orc_read({what}) failed: {rc}")
        return buf.view(dtype) if dtype != np.uint8 else buf


def ifft(values):
    v = np.ascontiguousarray(values, dtype=np.uint64)
    ncols, n = v.shape
    out = np.empty_like(v)
    lib().orc_ifft(_u64p(v), _u64p(out), n.bit_length() - 1, ncols)
    return out


def lde(coeffs, rate_bits=3):
    v = np.ascontiguousarray(coeffs, dtype=np.uint64)
    ncols, n = v.shape
    out = np.empty((ncols, n << rate_bits), dtype=np.uint64)
    lib().orc_lde(_u64p(v), _u64p(out), n.bit_length() - 1, rate_bits, ncols)
    return out


def coset_ifft_leaforder(values):
    v = np.ascontiguousarray(values, dtype=np.uint64)
    ncols, n = v.shape
    out = np.empty_like(v)
    lib().orc_coset_ifft_leaforder(_u64p(v), _u64p(out), n.bit_length() - 1, ncols)
    return out


def merkle_cap(leaves_colmajor, cap_height, hasher, want_digests=False):
    v = np.ascontiguousarray(leaves_colmajor, dtype=np.uint64)
    ncols, nl = v.shape
    hs = 25 if HASHER_ID[hasher] == 0 else 32
    ncap = 1 << min(cap_height, nl.bit_length() - 1)
    cap = np.empty(ncap * hs, dtype=np.uint8)
    dg = np.empty(nl * hs, dtype=np.uint8) if want_digests else None
    lib().orc_merkle_cap(_u64p(v), nl.bit_length() - 1, ncols, cap_height, HASHER_ID[hasher], cap.ctypes.data_as(C.c_void_p),
                         dg.ctypes.data_as(C.c_void_p) if want_digests else None)
    return (cap, dg) if want_digests else cap


def poseidon_permute(states):
    s = np.ascontiguousarray(states, dtype=np.uint64)
    out = np.empty_like(s)
    lib().orc_poseidon_permute(_u64p(s), _u64p(out), C.c_size_t(s.shape[0]))
    return out


def keccak256(msgs):
    m = np.ascontiguousarray(msgs, dtype=np.uint8)
    n, ln = m.shape
    out = np.empty((n, 32), dtype=np.uint8)
    lib().orc_keccak256(m.ctypes.data_as(C.c_void_p), C.c_size_t(ln), C.c_size_t(n), out.ctypes.data_as(C.c_void_p))
    return out


def eval_gate_constraints(cd, constants, wires, pi_hash):
    cs = np.ascontiguousarray(constants, dtype=np.uint64)
    w = np.ascontiguousarray(wires, dtype=np.uint64)
    npts = w.shape[1]
    d, keep = make_desc(cd, np.zeros((cd.num_preprocessed, cd.n), dtype=np.uint64))
    pi = np.array(list(pi_hash), dtype=np.uint64)
    out = np.empty((cd.num_gate_constraints, npts), dtype=np.uint64)
    lib().orc_eval_gate_constraints(C.byref(d), _u64p(cs), _u64p(w), _u64p(pi), C.c_size_t(npts), _u64p(out))
    return out


def num_threads():
    return lib().orc_num_threads()


def build_note():
    return _NATIVE_NOTE
