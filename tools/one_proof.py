"""One proof on cuda:0 (profiling driver: ncu -k regex:<kernel> python tools/one_proof.py [bits] [workload] [n] [hasher]).
workload: a synthetic gate mix of p2g.synth, or ecdsa-real = the EcdsaSecp256k1 program on 2^(bits-17) signatures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_product
p2g = load_product()
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = sys.argv[2] if len(sys.argv) > 2 else "ecdsa"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
hasher = sys.argv[4] if len(sys.argv) > 4 else "keccak25"
cfg = p2g.CircuitConfig.wide_ecc_config(hasher=hasher)
if wl == "ecdsa-real":
    sc = p2g.ecdsa_inputs.RealEcdsaCircuit(bits, p2g.acir, config=cfg, seed=1)
else:
    sc = p2g.synth.SyntheticCircuit(bits, wl, config=cfg, num_public_inputs=4, seed=1)
data = p2g.CircuitData(sc.common, sc.constants_sigmas)
for _ in range(reps):
    r = data.prove(sc.wires, sc.public_inputs)
print({k: round(v, 3) if isinstance(v, float) else v for k, v in r.timings.items()})
