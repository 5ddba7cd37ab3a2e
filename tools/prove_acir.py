"""What the reference's `prove` command does end to end (prove_action.rs:24-96: translate the ACIR program, build the circuit,
generate the witness, prove, compress), on this repository's stack, for programs made of real opcodes:
    python tools/prove_acir.py ecdsa [signatures]      the Noir signature-check program (one call = 2^17 rows)
    python tools/prove_acir.py sha256 [blocks]         SHA-256 of a (64 blocks - 64)-byte message: `blocks` Sha256Compression opcodes
Prints one JSON line with the time of every step (host steps on the CPU, circuit build and proof on cuda:0)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_product  # noqa: E402

p2g = load_product()
A = p2g.acir
program = sys.argv[1] if len(sys.argv) > 1 else "ecdsa"
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if program == "ecdsa":
    EI = p2g.ecdsa_inputs
    circuit, wit, _ = EI.circuit_and_witness(A, [EI.deterministic_case(i) for i in range(count)], outputs=[1] * count, assert_valid=True)
    what = f"{count} EcdsaSecp256k1 call(s) + 160 RANGE opcodes each + assert(valid)"
else:
    import acir_cases
    message = bytes((7 * i + 3) & 0xFF for i in range(64 * count - 64))
    circuit, wit, out_ids, digest = acir_cases.sha256_circuit(A, message)
    wit = {**wit, **{out_ids[i]: digest[i] for i in range(8)}}
    what = f"SHA-256 of {len(message)} bytes: {len(circuit.opcodes)} Sha256Compression opcodes"
t0 = time.perf_counter()
tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
t1 = time.perf_counter()
data, _ = tr.unpack()                      # p2g_circuit_create: the preprocessed commitment (plonky2's build())
t2 = time.perf_counter()
data2, _ = tr.unpack()                     # again, with the CUDA context, kernels and twiddle tables already in place
t2b = time.perf_counter()
data2.close()
wires, pis = tr.generate_witness(wit)      # generate_partial_witness + full_witness
t3 = time.perf_counter()
first = data.prove(wires, pis, compressed=True)   # the CLI writes the compressed proof (prove_action.rs:75-78)
t4 = time.perf_counter()
ts = []
for _ in range(5):
    t = time.perf_counter()
    pw = data.prove(wires, pis)
    ts.append(time.perf_counter() - t)
print(json.dumps({"program": what, "rows_log2": tr.common.degree_bits(), "rows_used": tr.rows_used(), "gate_types": len(tr.common.gates),
                  "translate_s": round(t1 - t0, 3), "circuit_build_s": round(t2 - t1, 3), "circuit_build_warm_s": round(t2b - t2, 3),
                  "witness_generation_s": round(t3 - t2b, 3),
                  "first_prove_s": round(t4 - t3, 3), "prove_ms_e2e_pageable": round(1e3 * min(ts), 2),
                  "device_total_ms": round(pw.timings["total_ms"], 2),
                  "stages_ms": {k: round(pw.timings[k], 2) for k in ("wires_commit_ms", "zs_pp_ms", "quotient_ms", "openings_ms", "fri_ms")},
                  "proof_bytes": len(pw.to_bytes()),
                  "cli_equivalent_s": round((t2 - t0) + (t3 - t2b) + min(ts), 3)}))
