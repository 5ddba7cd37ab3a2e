"""configs[2] from REAL opcodes: SHA-256 of a 448-byte message (8 Sha256Compression opcodes translated like the reference's
sha256_translator.rs) -> 2^18 rows; translate, generate the witness, prove on cuda:0 and print the timings as one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from __graft_entry__ import load_product  # noqa: E402
import acir_cases  # noqa: E402

p2g = load_product()
A = p2g.acir
message = bytes((7 * i + 3) & 0xFF for i in range(448))
circuit, wit, out_ids, digest = acir_cases.sha256_circuit(A, message)
t0 = time.perf_counter()
tr = A.CircuitBuilderFromAcirToPlonky2().translate_circuit(circuit)
t1 = time.perf_counter()
wires, pis = tr.generate_witness({**wit, **{out_ids[i]: digest[i] for i in range(8)}})
t2 = time.perf_counter()
data, _ = tr.unpack()
t3 = time.perf_counter()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
data.prove(wires, pis)
ts = []
for _ in range(reps):
    t = time.perf_counter()
    pw = data.prove(wires, pis)
    ts.append(time.perf_counter() - t)
print(json.dumps({"workload": "real SHA-256 circuit, 8 Sha256Compression opcodes", "rows_log2": tr.common.degree_bits(),
                  "gates": [g.id.split("(")[0] for g in tr.common.gates], "translate_s": round(t1 - t0, 3),
                  "witness_generation_s": round(t2 - t1, 3), "circuit_create_s": round(t3 - t2, 3),
                  "prove_ms_e2e_pageable": round(1e3 * min(ts), 2), "device_total_ms": round(pw.timings["total_ms"], 2),
                  "stages_ms": {k: round(pw.timings[k], 2) for k in ("wires_commit_ms", "zs_pp_ms", "quotient_ms", "openings_ms", "fri_ms")},
                  "proof_bytes": len(pw.to_bytes())}))
