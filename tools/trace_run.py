import sys, os, time
sys.path.insert(0,'.')
from __graft_entry__ import load_product
p2g = load_product()
import numpy as np, torch
sc = p2g.synth.SyntheticCircuit(20, "ecdsa", num_public_inputs=4, seed=1, pinned=True)
data = p2g.CircuitData(sc.common, sc.constants_sigmas)
wd = sc._wires_t.cuda()
for i in range(2):
    data.prove(wd, sc.public_inputs)
os.environ["P2G_TRACE"]="1"
print("--- device path", flush=True)
r=data.prove(wd, sc.public_inputs); print(r.timings, flush=True)
print("--- host path", flush=True)
r=data.prove(sc._wires_t, sc.public_inputs); print(r.timings, flush=True)
