"""include/p2g.hpp, the C++ host-side mirror of plonky2's CircuitData interface (the reference's host code is compiled Rust):
its derivations against the Python mirror on CPU, and a proof through it on the GPU against the Python path's bytes."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory, p2g):
    out = tmp_path_factory.mktemp("cpp") / "host_mirror"
    libdir = os.path.join(ROOT, "acvm-backend-plonky2_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp"), "-o", str(out), "-L", libdir, "-l:libp2g.so",
                           "-Wl,-rpath," + libdir])
    return str(out)


def spec_text(common, hasher_id, pis=None):
    lines = [f"{common.degree_bits_} {common.num_public_inputs} {hasher_id} {common.config.num_wires}"]
    for g in reversed(common.gates):     # any order: the mirror sorts
        lines.append(f"{g.kind} " + " ".join(str(x) for x in g.params))
    if pis is not None:
        lines.append("pis " + " ".join(str(int(x)) for x in pis))
    return "\n".join(lines) + "\n"


def test_cpp_mirror_derives_what_the_python_mirror_derives(p2g, exe):
    for hasher in ("keccak25", "poseidon"):
        for wl in ("assert_zero", "sha256", "ecdsa", "range", "all_gates"):
            for bits in (3, 12, 20, 22):
                cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
                gates = [g for g, _ in p2g.synth.gate_mix(wl, cfg)] + [p2g.Gate.noop(), p2g.Gate.public_input(), p2g.Gate.poseidon()]
                com = p2g.CommonCircuitData(cfg, bits, gates, num_public_inputs=4)
                r = subprocess.run([exe, "dump"], input=spec_text(com, p2g.lib.HASHER_ID[hasher]), capture_output=True, text=True)
                assert r.returncode == 0, r.stderr
                d = json.loads(r.stdout)
                assert d["gates"] == [[g.kind, *g.params] for g in com.gates]
                assert d["ids"] == [g.id for g in com.gates]
                assert d["selector_indices"] == com.selector_indices and d["groups"] == [list(x) for x in com.groups]
                assert (d["num_constants"], d["num_gate_constraints"], d["num_partial_products"], d["num_selectors"]) == \
                    (com.num_constants, com.num_gate_constraints, com.num_partial_products, com.num_selectors)
                assert d["reduction_arity_bits"] == com.reduction_arity_bits and d["k_is"] == com.k_is


@pytest.mark.gpu
@pytest.mark.parametrize("bits,workload,hasher", [(10, "all_gates", "keccak25"), (12, "ecdsa", "poseidon")])
def test_cpp_circuit_data_proves_the_same_bytes(p2g, exe, tmp_path, bits, workload, hasher):
    cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
    sc = p2g.synth.SyntheticCircuit(bits, workload, config=cfg, num_public_inputs=2, seed=77)
    with p2g.CircuitData(sc.common, sc.constants_sigmas) as data:
        want = data.prove(sc.wires, sc.public_inputs).to_bytes()
        want_c = data.prove(sc.wires, sc.public_inputs, compressed=True).to_bytes()
        cap = b"".join(data.constants_sigmas_cap)
        vk = data.verifier_data_bytes()
    (tmp_path / "spec").write_text(spec_text(sc.common, p2g.lib.HASHER_ID[hasher], sc.public_inputs))
    np.ascontiguousarray(sc.constants_sigmas).tofile(tmp_path / "cs.bin")
    np.ascontiguousarray(sc.wires).tofile(tmp_path / "wires.bin")
    r = subprocess.run([exe, "prove", str(tmp_path / "spec"), str(tmp_path / "cs.bin"), str(tmp_path / "wires.bin"),
                        str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert (tmp_path / "out.proof").read_bytes() == want
    assert (tmp_path / "out.cproof").read_bytes() == want_c
    assert (tmp_path / "out.cap").read_bytes() == cap
    assert (tmp_path / "out.vk").read_bytes() == vk
