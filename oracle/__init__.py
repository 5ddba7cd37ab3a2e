"""ORACLE = TEST INFRASTRUCTURE.  CPU restatement of the plonky2 0.2.2 prover/verifier that the reference
(eryxcoop/acvm-backend-plonky2) calls at plonky2-backend/src/actions/prove_action.rs:96.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  The product (acvm-backend-plonky2_b200/) never does.

Parity pinning: the verifier in oracle/pyref accepts the two golden proofs committed by the reference
(tests/golden/basic_{if,div}.proof) and the C prover in oracle/c regenerates both byte-for-byte from the
traces recovered from them.  FRI fold layers and the U32*/Comparison/RandomAccess gates are NOT exercised by
those proofs: for them parity is "restated verifier accepts restated prover" (see DESIGN.md).
"""
