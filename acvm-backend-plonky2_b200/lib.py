"""ctypes binding of libp2g.so, the C-ABI library declared in include/p2g.h.

The library is the product: every entry point here runs hand-written sm_100a kernels.  There is no CPU fallback -- if the
shared object is missing or no CUDA device is visible the calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libp2g.so")
_LIB = None

P2G_OK, P2G_EBADARG, P2G_ENOMEM, P2G_ECUDA, P2G_ENCCL, P2G_EUNSAT, P2G_ESMALLBUF = 0, -1, -2, -3, -4, -5, -6
HASH_KECCAK25, HASH_POSEIDON = 0, 1
HASHER_ID = {"keccak25": 0, "poseidon": 1, 0: 0, 1: 1}
(BUF_WIRES_CAP, BUF_ZS_PP_CAP, BUF_QUOTIENT_CAP, BUF_CS_CAP, BUF_ZS_PP_VALUES, BUF_QUOTIENT_CHUNKS, BUF_WIRES_COEFFS,
 BUF_CHALLENGES, BUF_FINAL_POLY, BUF_FRI_CAPS, BUF_WIRES_LDE, BUF_SHARD_INFO) = range(12)


class P2GError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libp2g error {code}: {msg}")
        self.code = code


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


class TimingsS(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("wires_commit_ms", C.c_float), ("zs_pp_ms", C.c_float),
                ("quotient_ms", C.c_float), ("openings_ms", C.c_float), ("fri_ms", C.c_float),
                ("total_ms", C.c_float), ("ntt_ms", C.c_float), ("merkle_ms", C.c_float),
                ("quotient_kernel_ms", C.c_float), ("ntt_bytes", C.c_double), ("merkle_bytes", C.c_double),
                ("kernel_launches", C.c_uint32), ("leaf_hash_launches", C.c_uint32),
                ("leaf_hash_ms", C.c_float), ("lde_ms", C.c_float), ("leaf_hash_bytes", C.c_double),
                ("lde_bytes", C.c_double), ("d2h_ms", C.c_float), ("lde_launches", C.c_uint32),
                ("h2d_bytes", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int)

# every symbol include/p2g.h declares (tests/test_abi.py checks the header against this list and the built library)
EXPORTS = ["p2g_version", "p2g_device_count", "p2g_last_error", "p2g_host_alloc", "p2g_host_free", "p2g_circuit_create", "p2g_circuit_destroy",
           "p2g_circuit_cap", "p2g_prove", "p2g_prove_device", "p2g_fill_advice_device", "p2g_prove_compressed", "p2g_prove_columns", "p2g_prove_routed_columns", "p2g_proof_size_bound", "p2g_circuit_create_sharded",
           "p2g_nccl_unique_id", "p2g_circuit_create_sharded_nccl", "p2g_vk_bytes",
           "p2g_circuit_read", "p2g_ifft", "p2g_lde", "p2g_coset_ifft_leaforder", "p2g_merkle_cap",
           "p2g_poseidon_permute", "p2g_keccak256", "p2g_eval_gate_constraints", "p2g_test_field_ops"]


class build_lock:
    """Serialises `make` across the processes of one box (torchrun ranks importing the package at the same moment): an
    exclusive flock on a file next to the libraries; the ranks that waited find the library up to date."""

    def __enter__(self):
        import fcntl
        self._f = open(os.path.join(_HERE, ".build.lock"), "w")
        fcntl.flock(self._f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self._f, fcntl.LOCK_UN)
        self._f.close()
        return False


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into libp2g.so (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if not f.startswith("build"))
        hdr = os.path.join(_HERE, "..", "include", "p2g.h")
        newest = max(newest, os.path.getmtime(hdr))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    args = ["make", "-C", csrc, "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(args)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise P2GError(P2G_ECUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`"
                                      " (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        missing = [s for s in EXPORTS if not hasattr(L, s)]
        if missing:
            raise P2GError(P2G_ECUDA, f"{LIB_PATH} lacks symbols {missing}: rebuild it")
        L.p2g_last_error.restype = C.c_char_p
        L.p2g_proof_size_bound.restype = C.c_size_t
        L.p2g_proof_size_bound.argtypes = [C.c_void_p]
        L.p2g_circuit_destroy.argtypes = [C.c_void_p]
        L.p2g_circuit_destroy.restype = None
        L.p2g_circuit_create.argtypes = [C.POINTER(DescS), C.c_int, C.POINTER(C.c_void_p)]
        L.p2g_circuit_cap.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.p2g_prove.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                C.POINTER(C.c_size_t), C.POINTER(TimingsS)]
        L.p2g_prove_device.argtypes = L.p2g_prove.argtypes
        L.p2g_fill_advice_device.argtypes = [C.c_void_p, C.c_void_p]
        L.p2g_prove_compressed.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_size_t), C.POINTER(TimingsS)]
        L.p2g_prove_routed_columns.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                               C.POINTER(C.c_size_t), C.c_void_p]
        L.p2g_prove_columns.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p,
                                        C.POINTER(C.c_size_t), C.c_void_p]
        L.p2g_circuit_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_size_t)]
        L.p2g_circuit_create_sharded.argtypes = [C.POINTER(DescS), C.c_int, C.c_int, C.c_int, ALLGATHER_FN, C.c_void_p,
                                                 C.POINTER(C.c_void_p)]
        L.p2g_nccl_unique_id.argtypes = [C.c_void_p]
        L.p2g_circuit_create_sharded_nccl.argtypes = [C.POINTER(DescS), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.p2g_vk_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
        L.p2g_ifft.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.p2g_lde.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.p2g_coset_ifft_leaforder.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
        L.p2g_merkle_cap.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                     C.c_int]
        L.p2g_poseidon_permute.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.p2g_keccak256.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int]
        L.p2g_eval_gate_constraints.argtypes = [C.POINTER(DescS), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                                C.c_void_p, C.c_int]
        L.p2g_host_alloc.restype = C.c_void_p
        L.p2g_host_alloc.argtypes = [C.c_size_t]
        L.p2g_host_free.argtypes = [C.c_void_p]
        L.p2g_host_free.restype = None
        L.p2g_test_field_ops.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise P2GError(rc, lib().p2g_last_error().decode(errors="replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a


# ---- stand-alone kernels (host buffers in/out) -------------------------------------------------------------------
def ifft(values, device=0):
    v = _u64(values)
    ncols, n = v.shape
    out = np.empty_like(v)
    check(lib().p2g_ifft(_p(v), _p(out), n.bit_length() - 1, ncols, device))
    return out


def lde(coeffs, rate_bits=3, device=0):
    v = _u64(coeffs)
    ncols, n = v.shape
    out = np.empty((ncols, n << rate_bits), dtype=np.uint64)
    check(lib().p2g_lde(_p(v), _p(out), n.bit_length() - 1, rate_bits, ncols, device))
    return out


def coset_ifft_leaforder(values, device=0):
    v = _u64(values)
    ncols, n = v.shape
    out = np.empty_like(v)
    check(lib().p2g_coset_ifft_leaforder(_p(v), _p(out), n.bit_length() - 1, ncols, device))
    return out


def merkle_cap(leaves_colmajor, cap_height, hasher, want_digests=False, device=0):
    v = _u64(leaves_colmajor)
    ncols, nl = v.shape
    h = HASHER_ID[hasher]
    hs = 25 if h == 0 else 32
    ncap = 1 << min(cap_height, nl.bit_length() - 1)
    cap = np.empty(ncap * hs, dtype=np.uint8)
    dg = np.empty(nl * hs, dtype=np.uint8) if want_digests else None
    check(lib().p2g_merkle_cap(_p(v), nl.bit_length() - 1, ncols, cap_height, h, _p(cap),
                               _p(dg) if want_digests else None, device))
    return (cap, dg) if want_digests else cap


def poseidon_permute(states, device=0):
    s = _u64(states)
    out = np.empty_like(s)
    check(lib().p2g_poseidon_permute(_p(s), _p(out), s.shape[0], device))
    return out


def keccak256(msgs, device=0):
    m = np.ascontiguousarray(msgs, dtype=np.uint8)
    n, ln = m.shape
    out = np.empty((n, 32), dtype=np.uint8)
    check(lib().p2g_keccak256(_p(m), ln, n, _p(out), device))
    return out


FIELD_OPS = {"sub": 0, "add": 1, "mul": 2, "mulz": 3, "mul_small": 4, "canon": 5, "reduce128": 6, "dot8": 7, "dot8_small": 8, "addf": 9, "mulf": 10, "canonf": 11}


def field_ops(op, a, b, device=0):
    """Self-test of csrc/gl.cuh: device >= 0 runs the sm_100a PTX forms, device < 0 their host twins."""
    a, b = _u64(a), _u64(b)
    out = np.empty(a.size // 8 if op.startswith("dot8") else a.size, dtype=np.uint64)
    check(lib().p2g_test_field_ops(FIELD_OPS[op], _p(a), _p(b), _p(out), a.size, device))
    return out
