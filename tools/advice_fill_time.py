"""Time p2g_fill_advice_device on the real 2^20-row ECDSA circuit and check it against the generators' witness (B200)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_product  # noqa: E402

p2g = load_product()
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sc = p2g.ecdsa_inputs.RealEcdsaCircuit(bits, p2g.acir, seed=3, pinned=True)
data = p2g.CircuitData(sc.common, sc.constants_sigmas)
full = sc._wires_t.cuda()
dev = torch.empty_like(full)
routed_host = sc._wires_t[:80]
ts, up = [], []
for _ in range(4):
    dev.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev[:80].copy_(routed_host, non_blocking=False)       # all the host still has to ship: 80 of 234 columns
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    data.fill_advice(dev)
    t2 = time.perf_counter()
    up.append(t1 - t0)
    ts.append(t2 - t1)
same = bool(torch.equal(dev, full))
t0 = time.perf_counter()
full2 = sc._wires_t.cuda()
torch.cuda.synchronize()
t_full = time.perf_counter() - t0
cols = [np.array(sc.wires[c]) for c in range(sc.wires.shape[0])]       # pageable, one allocation per column
def best(fn, k=4):
    fn()
    out = []
    for _ in range(k):
        t = time.perf_counter()
        r = fn()
        out.append(time.perf_counter() - t)
    return 1e3 * min(out), r
ms_all, p_all = best(lambda: data.prove_columns(cols, sc.public_inputs))
ms_routed, p_routed = best(lambda: data.prove_routed_columns(cols[:80], sc.public_inputs))
pw = data.prove(dev, sc.public_inputs)
print(json.dumps({"rows_log2": bits, "fill_advice_ms": round(1e3 * min(ts), 3), "upload_routed_columns_ms": round(1e3 * min(up), 2),
                  "upload_all_columns_ms": round(1e3 * t_full, 2), "bytes_routed": int(routed_host.numel() * 8),
                  "bytes_all": int(sc._wires_t.numel() * 8), "matches_generators": same,
                  "prove_columns_pageable_ms": round(ms_all, 2), "prove_routed_columns_pageable_ms": round(ms_routed, 2),
                  "routed_columns_proof_equal": p_all.to_bytes() == p_routed.to_bytes(),
                  "proof_equals_full_witness_proof": pw.to_bytes() == data.prove(full, sc.public_inputs).to_bytes()}))
