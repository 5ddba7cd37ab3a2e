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
