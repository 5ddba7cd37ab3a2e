"""CPU-only: libp2g.so loads and exports every function include/p2g.h declares; the ctypes structs match the header."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "p2g.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(p2g_[a-z0-9_]+)\s*\(", src)) - {"p2g_allgather_fn"})


def test_header_symbols_are_exported(p2g):
    names = header_functions()
    assert len(names) >= 18
    L = C.CDLL(p2g.lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(p2g.lib.EXPORTS) == names          # the binding knows exactly the header's surface


def test_version_and_error_slot(p2g):
    L = p2g.lib.lib()
    assert L.p2g_version() == 2
    assert isinstance(L.p2g_last_error(), bytes)


def test_struct_layout_matches_header(p2g):
    # p2g_gate: 9 x u32; p2g_circuit_desc: 17 u32 + 8 u32 + u32 (+pad) + 4 pointers
    assert C.sizeof(p2g.lib.GateS) == 36
    assert C.sizeof(p2g.lib.DescS) == 4 * 26 + 4 * 8 + (0 if (4 * 26) % 8 == 0 else 4)
    # a descriptor with the wrong struct_size is refused before any CUDA call
    d = p2g.lib.DescS()
    d.struct_size = 4
    h = C.c_void_p()
    rc = p2g.lib.lib().p2g_circuit_create(C.byref(d), 0, C.byref(h))
    assert rc == p2g.lib.P2G_EBADARG and b"struct_size" in p2g.lib.lib().p2g_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "acvm-backend-plonky2_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle|liborc", txt, flags=re.M), f


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: the header must compile as C (no C++-isms, no torch / CUDA types in the signatures)."""
    import subprocess
    src = tmp_path / "use_p2g.c"
    src.write_text('#include "p2g.h"\n'
                   'int main(void) { p2g_circuit_desc d; d.struct_size = (uint32_t)sizeof d; (void)d;\n'
                   '  return p2g_version() == P2G_VERSION && p2g_last_error() != 0 ? 0 : 1; }\n')
    exe = tmp_path / "use_p2g"
    libdir = os.path.join(ROOT, "acvm-backend-plonky2_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libp2g.so", "-Wl,-rpath," + libdir])
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "p2g.h")).read(), flags=re.S)   # declarations only
    assert "torch" not in hdr and "cudaStream" not in hdr and "std::" not in hdr
    assert subprocess.run([str(exe)]).returncode == 0        # p2g_version / p2g_last_error need no GPU


def _desc(p2g, workload="all_gates", bits=6):
    sc = p2g.synth.SyntheticCircuit(bits, workload, num_public_inputs=2, seed=5)
    d, keep = sc.common.fill_desc(sc.constants_sigmas)
    return sc, d, keep


def test_malformed_descriptors_are_refused_before_any_cuda_call(p2g):
    """include/p2g.h promises P2G_EBADARG for malformed descriptors: gate parameters are checked against the circuit
    shape, so a bad table can never drive the kernels out of bounds.  No GPU needed: validation precedes device work."""
    L = p2g.lib.lib()
    h = C.c_void_p()

    def refused(mutate, needle):
        sc, d, keep = _desc(p2g)
        mutate(d, sc)
        rc = L.p2g_circuit_create(C.byref(d), 0, C.byref(h))
        assert rc == p2g.lib.P2G_EBADARG, (needle, rc, L.p2g_last_error())
        assert needle.encode() in L.p2g_last_error(), L.p2g_last_error()
        del keep

    def gate_of(d, kind):
        return next(i for i in range(d.num_gates) if d.gates[i].kind == kind)
    K = p2g.circuit
    refused(lambda d, sc: setattr(d, "num_wires", 100), "more wires than num_wires")                    # U32 gates need > 200 wires
    refused(lambda d, sc: setattr(d.gates[gate_of(d, K.ARITHMETIC)], "num_constraints", 19), "num_constraints does not match")
    refused(lambda d, sc: setattr(d, "num_selectors", d.num_constants + 1), "num_selectors > num_constants")
    refused(lambda d, sc: setattr(d, "pow_bits", 65), "pow_bits")
    refused(lambda d, sc: setattr(d, "cap_height", 12), "cap_height")
    refused(lambda d, sc: d.gates[gate_of(d, K.COMPARISON)].params.__setitem__(1, 0), "ComparisonGate")  # num_chunks = 0 would divide by zero
    refused(lambda d, sc: d.gates[gate_of(d, K.RANDOM_ACCESS)].params.__setitem__(0, 7), "RandomAccessGate bits")
    refused(lambda d, sc: setattr(d, "num_constants", d.num_selectors + 1), "more constants than")       # ArithmeticGate reads 2 constants

    def deep_fri(d, sc):
        d.num_fri_layers = 1
        d.reduction_arity_bits[0] = 6          # 2^(6 + 3 - 6) = 8 leaves < 2^cap_height
    refused(deep_fri, "fewer than 2^cap_height leaves")
    # and the untouched descriptor passes validation: the only failure left on a machine without a GPU is the device itself
    sc, d, keep = _desc(p2g)
    rc = L.p2g_circuit_create(C.byref(d), 0, C.byref(h))
    if rc == 0:
        L.p2g_circuit_destroy(h)
    else:
        assert rc == p2g.lib.P2G_ECUDA, L.p2g_last_error()
