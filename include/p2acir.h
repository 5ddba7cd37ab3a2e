/* p2acir.h -- C ABI of libp2acir.so: ACIR program -> Plonky2 circuit payload and witness, outside Rust.
 *
 * HOST TOOLING, not part of the drop-in boundary (that is p2g.h): it restates the two reference layers that sit ABOVE
 * `circuit_data.prove(witnesses)` so that realistic circuits and witnesses exist without cargo/nargo:
 *   - the translator  plonky2-backend/src/circuit_translation/mod.rs:72-190  (CircuitBuilderFromAcirToPlonky2::translate_circuit)
 *     with assert_zero_translator.rs, memory_translator.rs, sha256_translator.rs, ecdsa_secp256k1_translator.rs,
 *     binary_digits_target.rs and the plonky2_ecdsa gadgets (biguint/, curve/);
 *   - plonky2's generate_partial_witness + full_witness (prove_action.rs:96 runs them inside prove) for the generators those
 *     circuits contain, the reference's custom ones included (plonky2_ecdsa/biguint/gates/{run_once}, gadgets/nonnative.rs,
 *     biguint.rs, curve/gadgets/glv.rs).
 * In a deployment this payload comes from the Rust side through rust_shim/; nothing here runs on the GPU.
 *
 * Opcode stream (`ops`, u64 words), witness ids are ACIR witness indices:
 *   1 AssertZero: n_mul, n_lin, q_c, n_mul x (coef, w1, w2), n_lin x (coef, w)
 *   2 RANGE: w, num_bits (<= 33)          3 AND / 4 XOR: lhs, rhs, num_bits, output
 *   5 MemoryInit: block, n, n witnesses   6 MemoryRead / 8 MemoryWrite: block, index witness, value witness
 *   7 Sha256Compression: 16 inputs, 8 hash values, 8 outputs
 *   9 EcdsaSecp256k1: 32 public_key_x, 32 public_key_y, 64 signature, 32 hashed_message byte witnesses, output
 *   101..105, 110: gadget-level operations for tests (u32 gadgets; biguint / non-native / curve gadgets), see p2acir.cpp.
 * Field elements are canonical Goldilocks values.  Functions returning int give 0 or -1 (message in p2a_last_error()).
 */
#ifndef P2ACIR_H
#define P2ACIR_H
#include <stddef.h>
#include <stdint.h>

#include "p2g.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* p2a_last_error(void);

/* translate_circuit: public / private parameter witness ids (each sorted), the opcode stream -> handle, or NULL.
 * The circuit uses CircuitConfig::wide_ecc_config() (234 wires, 80 routed, 2 constants per gate; mod.rs:69). */
void* p2a_translate(const uint64_t* public_params, size_t n_public, const uint64_t* private_params, size_t n_private,
                    const uint64_t* ops, size_t n_words);
void p2a_destroy(void* circuit);

/* degree_bits, number of distinct gate types, number of public inputs */
void p2a_shape(void* circuit, uint32_t* degree_bits, uint32_t* num_gate_types, uint32_t* num_public_inputs);
/* the gate types in creation order: kind (P2G_GATE_*) + params[4] each; the caller sorts them like plonky2 and derives selectors */
void p2a_gate_types(void* circuit, uint32_t* kind_and_params);
/* rows in use before the power-of-two padding with NoopGate */
uint32_t p2a_rows_used(void* circuit);

/* constants_sigmas [num_constants + 80][2^degree_bits] for the SORTED gate table `gates` (selector data filled in);
 * type_to_gate[i] = index in `gates` of creation-order gate type i; k_is = the 80 coset shifts */
int p2a_constants_sigmas(void* circuit, const p2g_gate* gates, uint32_t num_gates, const uint32_t* type_to_gate, uint32_t num_selectors,
                         uint32_t num_constants, const uint64_t* k_is, uint64_t* out);

/* generate_partial_witness + full_witness: ACIR witness map (ids, values) -> wires [234][2^degree_bits] (unset wires = 0) and the
 * public inputs in registration order.  -1 when a copy constraint is contradicted (the reference panics there). */
int p2a_witness(void* circuit, const uint64_t* ids, const uint64_t* values, size_t n, uint64_t* wires, uint64_t* public_inputs);
/* the same, producing only the routed columns [80][2^degree_bits]: the advice wires are not generated (the device computes them,
 * p2g_prove_routed_columns / p2g_fill_advice_device) */
int p2a_witness_routed(void* circuit, const uint64_t* ids, const uint64_t* values, size_t n, uint64_t* routed_wires,
                       uint64_t* public_inputs);
/* the values the last p2a_witness left on ACIR witnesses (outputs computed by generators included); known[i] = 0 if unset */
void p2a_read_witnesses(void* circuit, const uint64_t* ids, size_t n, uint64_t* values, uint8_t* known);
/* the host twin of p2g_fill_advice_device (csrc/advice.cuh): recomputes, in place, the advice columns (>= 80) of a wire matrix of
 * this circuit from its routed columns; used by the tests to check that function against the witness generators */
void p2a_fill_advice(void* circuit, uint64_t* wires);
/* the same for any trace: the gate table as (kind, params[4]) per gate, the gate index of every row, wires [num_wires][n] */
void p2a_fill_advice_rows(const uint32_t* kind_and_params, uint32_t num_gates, const uint8_t* row_gate, uint64_t* wires, size_t n,
                          uint32_t num_wires, uint32_t num_routed);
/* worker threads of p2a_witness / p2a_constants_sigmas (0 = all cores) */
void p2a_set_threads(int n);

/* self-test hook for the big-integer arithmetic behind the non-native generators (acir/bigint.h): limbs are u32, least significant
 * first.  op 0: q = a / b, r = a % b;  op 1: q = a * b;  op 2: q = a^b mod m;  op 3: GLV decomposition of a (mod the secp256k1
 * group order): q = |k1|, r = |k2|, flags bit 0 = k1 < 0, bit 1 = k2 < 0;  op 4: q = a^-1 mod m (m odd, gcd(a, m) = 1).  Outputs hold up to 40 limbs; returns 0 or -1. */
int p2a_bigint_selftest(int op, const uint32_t* a, size_t na, const uint32_t* b, size_t nb, const uint32_t* m, size_t nm,
                        uint32_t* q, size_t* nq, uint32_t* r, size_t* nr, uint32_t* flags);

#ifdef __cplusplus
}
#endif
#endif
