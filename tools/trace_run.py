"""Host-side stage trace (P2G_TRACE=1) of one device-resident and one host-resident proof: python tools/trace_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_product  # noqa: E402

p2g = load_product()
sc = p2g.synth.SyntheticCircuit(20, "ecdsa", num_public_inputs=4, seed=1, pinned=True)
data = p2g.CircuitData(sc.common, sc.constants_sigmas)
wd = sc._wires_t.cuda()
for _ in range(2):
    data.prove(wd, sc.public_inputs)
os.environ["P2G_TRACE"] = "1"
print("--- device path", flush=True)
print(data.prove(wd, sc.public_inputs).timings, flush=True)
print("--- host path", flush=True)
print(data.prove(sc._wires_t, sc.public_inputs).timings, flush=True)
