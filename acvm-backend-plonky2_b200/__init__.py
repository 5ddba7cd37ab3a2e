"""p2g: a B200-native (sm_100a) Plonky2 prover behind the reference's `circuit_data.prove(witnesses)` call
(/root/reference/plonky2-backend/src/actions/prove_action.rs:96).  See DESIGN.md and include/p2g.h.

Host-side Python mirror of the reference interface for this path; all compute is in libp2g.so (hand-written CUDA)."""
from . import lib  # noqa: F401
from .lib import P2GError, build  # noqa: F401
from . import circuit  # noqa: F401,E402
from . import synth  # noqa: F401,E402
from .circuit import CircuitConfig, CircuitData, CommonCircuitData, Gate, ProofWithPublicInputs  # noqa: F401,E402
from . import sharding  # noqa: F401,E402
from . import acir  # noqa: F401,E402
from . import ecdsa_inputs  # noqa: F401,E402
