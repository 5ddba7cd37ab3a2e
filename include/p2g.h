/*
 * p2g.h -- C ABI of the B200-native Plonky2 prover ("libp2g").
 *
 * This is the drop-in boundary for ONE expression of the reference:
 *
 *     circuit_data.prove(witnesses).unwrap()
 *         /root/reference/plonky2-backend/src/actions/prove_action.rs:91-97           (CLI `prove`)
 *         /root/reference/plonky2-backend/src/circuit_translation/tests/factories/utils.rs:16-27  (every translator test)
 *
 * i.e. plonky2 0.2.2 `prove_with_partition_witness(&ProverOnlyCircuitData, &CommonCircuitData, PartitionWitness, ..)`.
 * The reference has no FFI of its own; INTEGRATION.md shows the Rust `extern "C"` block and the shim a
 * maintainer adds at prove_action.rs:96.  Everything crossing this boundary is plain pointers + sizes:
 * caller-owned host buffers in, caller-owned host buffers out, device memory owned by the handle.
 *
 * All field elements are canonical Goldilocks u64 (< p = 2^64 - 2^32 + 1), little-endian.
 */
#ifndef P2G_H
#define P2G_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2G_VERSION 2

/* ---- error codes (replace the reference's unwrap()/expect() panics; prove_action.rs:77,96) ---- */
#define P2G_OK 0
#define P2G_EBADARG (-1)   /* malformed descriptor / non-canonical input / null pointer                                */
#define P2G_ENOMEM (-2)    /* host or device allocation failed                                                         */
#define P2G_ECUDA (-3)     /* CUDA runtime error (message in p2g_last_error)                                           */
#define P2G_ENCCL (-4)     /* NCCL error                                                                               */
#define P2G_EUNSAT (-5)    /* zeta fell in the subgroup / forced pow_witness invalid / quotient check failed          */
#define P2G_ESMALLBUF (-6) /* output buffer too small; *out_len holds the required size                                */

/* ---- hashers: plonky2 `GenericConfig::Hasher` (plonky2-backend/src/lib.rs:13 uses KeccakGoldilocksConfig) ---- */
enum p2g_hasher {
    P2G_HASH_KECCAK25 = 0, /* KeccakHash<25>: Merkle + challenger of the CLI                                          */
    P2G_HASH_POSEIDON = 1  /* PoseidonHash (PoseidonGoldilocksConfig; plonky2_ecdsa gadget tests)                     */
};

/* ---- gate universe: the 13 gates reachable from the translators, of the 22 that
 *      /root/reference/plonky2-backend/src/actions/write_vk_action.rs:37-61 registers.
 *      params[] = the integers BackendGateSerializer writes for the gate, in its order. ---- */
enum p2g_gate_kind {
    P2G_GATE_NOOP = 0,
    P2G_GATE_CONSTANT = 1,        /* params: num_consts                                                                */
    P2G_GATE_PUBLIC_INPUT = 2,
    P2G_GATE_ARITHMETIC = 3,      /* params: num_ops                                                                   */
    P2G_GATE_BASE_SUM = 4,        /* params: B, num_limbs                                                              */
    P2G_GATE_POSEIDON = 5,
    P2G_GATE_RANDOM_ACCESS = 6,   /* params: bits, num_copies, num_extra_constants                                     */
    P2G_GATE_U32_ARITHMETIC = 7,  /* params: num_ops                      plonky2_ecdsa/biguint/gates/arithmetic_u32.rs  */
    P2G_GATE_U32_ADD_MANY = 8,    /* params: num_addends, num_ops         .../add_many_u32.rs                           */
    P2G_GATE_U32_SUBTRACTION = 9, /* params: num_ops                      .../subtraction_u32.rs                        */
    P2G_GATE_U32_RANGE_CHECK = 10,/* params: num_input_limbs              .../range_check_u32.rs                        */
    P2G_GATE_COMPARISON = 11,     /* params: num_bits, num_chunks         .../comparison.rs                             */
    P2G_GATE_KIND_COUNT = 12
};

/* One entry of `common.gates` (sorted by (degree, id)), with its selector data from `common.selectors_info`. */
typedef struct p2g_gate {
    uint32_t kind;           /* enum p2g_gate_kind                                                                     */
    uint32_t params[4];
    uint32_t selector_index; /* selectors_info.selector_indices[i]                                                     */
    uint32_t group_lo;       /* selectors_info.groups[selector_index].start                                            */
    uint32_t group_hi;       /* ... .end (exclusive)                                                                   */
    uint32_t num_constraints;
} p2g_gate;

#define P2G_MAX_FRI_LAYERS 8

/* What `CommonCircuitData` + `ProverOnlyCircuitData` hold that the prover reads. */
typedef struct p2g_circuit_desc {
    uint32_t struct_size;            /* = sizeof(p2g_circuit_desc), ABI guard                                          */
    uint32_t degree_bits;            /* common.degree_bits(); N = 2^degree_bits rows                                   */
    uint32_t num_wires;              /* config.num_wires (234 for wide_ecc_config, circuit_translation/mod.rs:69)      */
    uint32_t num_routed_wires;       /* config.num_routed_wires (80)                                                   */
    uint32_t num_constants;          /* common.num_constants: selector columns + gate constant columns                 */
    uint32_t num_selectors;          /* selectors_info.num_selectors()                                                 */
    uint32_t num_challenges;         /* config.num_challenges (2)                                                      */
    uint32_t rate_bits;              /* fri_config.rate_bits (3)                                                       */
    uint32_t cap_height;             /* fri_config.cap_height (4)                                                      */
    uint32_t pow_bits;               /* fri_config.proof_of_work_bits (16)                                             */
    uint32_t num_query_rounds;       /* fri_config.num_query_rounds (28)                                               */
    uint32_t quotient_degree_factor; /* common.quotient_degree_factor (8)                                              */
    uint32_t num_partial_products;   /* common.num_partial_products (9)                                                */
    uint32_t num_gate_constraints;   /* common.num_gate_constraints                                                    */
    uint32_t num_public_inputs;      /* common.num_public_inputs                                                       */
    uint32_t hasher;                 /* enum p2g_hasher                                                                */
    uint32_t num_fri_layers;         /* fri_params.reduction_arity_bits.len()                                          */
    uint32_t reduction_arity_bits[P2G_MAX_FRI_LAYERS];
    uint32_t num_gates;
    const p2g_gate* gates;           /* [num_gates]                                                                    */
    /* Preprocessed polynomials as VALUES on <omega_N> in natural row order, column-major:
     * column c (0 <= c < num_constants + num_routed_wires) at constants_sigmas[c*N .. (c+1)*N).
     * Columns: selectors, gate constants, then the num_routed sigma columns (prover_only.sigmas transposed). */
    const uint64_t* constants_sigmas;
    const uint64_t* k_is;            /* [num_routed_wires] common.k_is                                                 */
    const uint8_t* circuit_digest;   /* prover_only.circuit_digest bytes (25 or 32), or NULL to derive it              */
} p2g_circuit_desc;

/* Per-stage device timings of the last p2g_prove (CUDA events), plonky2 TimingTree stage names. */
typedef struct p2g_timings {
    float h2d_ms;              /* trace upload                                                                         */
    float wires_commit_ms;     /* "compute wires commitment"                                                           */
    float zs_pp_ms;            /* "compute partial products" (incl. commitment)                                        */
    float quotient_ms;         /* "compute quotient polys" + "commit to quotient polys"                                */
    float openings_ms;         /* "construct the opening set"                                                          */
    float fri_ms;              /* "compute opening proofs" (combine, commit phase, PoW, queries)                       */
    float total_ms;            /* whole call on the device timeline                                                    */
    float ntt_ms;              /* sum over NTT/LDE kernels                                                             */
    float merkle_ms;           /* sum over leaf + node hashing kernels                                                 */
    float quotient_kernel_ms;  /* constraint-evaluation kernel alone                                                   */
    double ntt_bytes;          /* algorithmic bytes moved by the NTT/LDE kernels                 */
    double merkle_bytes;       /* algorithmic bytes hashed                                                             */
    uint32_t kernel_launches;  /* launches of libp2g kernels inside the call                                           */
    uint32_t leaf_hash_launches; /* launches of the Merkle leaf-hashing kernel (the dominant kernel)                   */
    float leaf_hash_ms;        /* device time of those launches, CUDA events on the library's stream                   */
    float lde_ms;              /* device time of the coset-LDE passes alone                                            */
    double leaf_hash_bytes;    /* algorithmic bytes read by the leaf-hashing launches (8 * leaves * columns)           */
    double lde_bytes;          /* algorithmic bytes of the LDE passes (8N read + 64N written per column)               */
    float d2h_ms;              /* query-opening gather + proof assembly                                                */
    uint32_t lde_launches;
    double h2d_bytes;          /* bytes of the trace this call uploaded (a sharded rank uploads only the columns it reads)  */
} p2g_timings;

typedef struct p2g_circuit p2g_circuit;

/* ---- lifecycle ---------------------------------------------------------------------------------------------------- */
int p2g_version(void);
int p2g_device_count(void);
const char* p2g_last_error(void); /* thread-local, never NULL */

/* Pinned (page-locked) host memory for the witness matrix handed to p2g_prove; pageable memory works too, slower. */
void* p2g_host_alloc(size_t bytes);
void p2g_host_free(void* p);

/* Per circuit, once: uploads the preprocessed polynomials, builds their LDE + Merkle tree on the device (what
 * CircuitBuilder::build() does at circuit_translation/mod.rs:81) and keeps them resident across proofs.
 * device = CUDA ordinal. */
int p2g_circuit_create(const p2g_circuit_desc* desc, int device, p2g_circuit** out);
void p2g_circuit_destroy(p2g_circuit* c);

/* constants_sigmas_cap (2^cap_height digests) and circuit_digest, so the caller can assert equality with
 * verifier_only.constants_sigmas_cap / prover_only.circuit_digest. */
int p2g_circuit_cap(const p2g_circuit* c, uint8_t* cap_out, size_t cap_len, uint8_t* digest_out, size_t digest_len);

/* ---- the hot path -------------------------------------------------------------------------------------------------
 * wires:  num_wires x N, column-major (MatrixWitness.wire_values), canonical.
 * public_inputs: values of the registered public-input targets.
 * forced_pow_witness: NULL = smallest valid witness (deterministic; the reference's Rayon find_any is not).
 * out/out_len: plonky2 *uncompressed* ProofWithPublicInputs::to_bytes layout; in: capacity, out: size.
 */
int p2g_prove(p2g_circuit* c, const uint64_t* wires, const uint64_t* public_inputs, size_t num_public_inputs,
              const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len, p2g_timings* timings);

/* Same, with the trace already resident on the device (d_wires is a device pointer on the handle's device). */
int p2g_prove_device(p2g_circuit* c, const uint64_t* d_wires, const uint64_t* public_inputs, size_t num_public_inputs,
                     const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len, p2g_timings* timings);

/* Device-side witness fill, first step (csrc/advice.cuh): completes, in place, the ADVICE columns -- wires at or
 * above num_routed_wires, which no copy constraint reaches -- of a device-resident trace d_wires [num_wires][N] from its routed
 * columns: the 2-bit limbs of the u32 gates (arithmetic_u32.rs:376-426, add_many_u32.rs:329-375, subtraction_u32.rs:298-343,
 * range_check_u32.rs:198-220), the tails of ComparisonGate (comparison.rs:439-537) and RandomAccessGate, the internal state of
 * PoseidonGate.  What plonky2's generator queue would have written there; the routed columns (all that a host-side witness
 * generator still has to produce and upload: 80 of 234) are read, never written.  Follow with p2g_prove_device. */
int p2g_fill_advice_device(p2g_circuit* c, uint64_t* d_wires);

/* p2g_prove_columns with two thirds of the trace left at home: routed_columns holds only the num_routed_wires ROUTED columns of
 * MatrixWitness.wire_values (N words each, any representative < 2^64); they are uploaded through the same chunked pipeline and
 * the advice columns are computed on the device as soon as they are in place (what p2g_fill_advice_device does, inside the
 * pipeline).  Single-GPU handles; same bytes as p2g_prove_columns on the full witness. */
int p2g_prove_routed_columns(p2g_circuit* c, const uint64_t* const* routed_columns, const uint64_t* public_inputs,
                             size_t num_public_inputs, const uint64_t* forced_pow_witness, int compressed, uint8_t* out,
                             size_t* out_len, p2g_timings* timings);

/* Same proof in plonky2's *compressed* layout, CompressedProofWithPublicInputs::to_bytes -- byte for byte what the reference
 * CLI writes to the proof file (prove_action.rs:75-78 `proof.compress(..)`, `to_bytes()`; the committed golden proofs under
 * example_programs/ are in this format).  Covers SURVEY 8(f) row f3.  wires_on_device: 0 = host pointer, 1 = device pointer. */
int p2g_prove_compressed(p2g_circuit* c, const uint64_t* wires, int wires_on_device, const uint64_t* public_inputs,
                         size_t num_public_inputs, const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len,
                         p2g_timings* timings);

/* Same, taking the witness the way plonky2 holds it: `MatrixWitness.wire_values` is a `Vec<Vec<F>>` with one allocation per
 * wire column (plonky2 iop/witness.rs; reached from prove_action.rs:96 through `PartitionWitness::full_witness()`, SURVEY 8a row
 * a3).  wire_columns[i] points at the N u64 of column i (GoldilocksField is a transparent u64), so the Rust shim passes
 * num_wires pointers and never builds a flat copy.  Unlike p2g_prove, the words may be ANY representative < 2^64 of the field
 * element (plonky2's add/sub leave values in [p, 2^64)); the library reduces them on the device.  compressed: 0 = ProofWithPublicInputs::to_bytes, 1 = the CLI's compressed
 * file format.  Pageable columns are staged through a pinned ring by several host threads; page-locked ones are copied directly. */
int p2g_prove_columns(p2g_circuit* c, const uint64_t* const* wire_columns, const uint64_t* public_inputs,
                      size_t num_public_inputs, const uint64_t* forced_pow_witness, int compressed, uint8_t* out,
                      size_t* out_len, p2g_timings* timings);

/* ---- verification key -------------------------------------------------------------------------------
 * `verifier_data().to_bytes(&BackendGateSerializer)` -- the file `write_vk` writes (write_vk_action.rs:76-79) -- from a circuit
 * handle, so the CLI does not have to rebuild the circuit a second time for it: VerifierOnlyCircuitData (cap height, the
 * constants_sigmas cap, circuit_digest) followed by CommonCircuitData (config, FRI params, selectors, counts, k_is, gate list
 * with the tags of the 22-entry BackendGateSerializer list, write_vk_action.rs:37-61, and each gate's own payload, e.g.
 * add_many_u32.rs:94-97).  The CircuitConfig / FriConfig fields the prover never reads travel in p2g_vk_config (NULL = the
 * values of CircuitConfig::wide_ecc_config(), circuit_translation/mod.rs:69).  out == NULL or too small: P2G_ESMALLBUF with
 * the size in *out_len.  plonky2's layout is restated from util/serialization.rs of the un-vendored crate; no VK file is committed
 * in the reference, so this layout is NOT pinned by a golden vector. */
typedef struct p2g_vk_config {
    uint32_t struct_size;                /* = sizeof(p2g_vk_config)                                                          */
    uint32_t config_num_constants;       /* config.num_constants (2)                                                         */
    uint32_t security_bits;              /* config.security_bits (100)                                                       */
    uint32_t max_quotient_degree_factor; /* config.max_quotient_degree_factor (8)                                            */
    uint32_t use_base_arithmetic_gate;   /* config.use_base_arithmetic_gate (1)                                              */
    uint32_t zero_knowledge;             /* config.zero_knowledge (0) = fri_params.hiding                                    */
    uint32_t reduction_strategy;         /* FriReductionStrategy: 0 Fixed(reduction_arity_bits), 1 ConstantArityBits, 2 MinSize */
    uint32_t strategy_params[2];         /* ConstantArityBits(4, 5); MinSize: {is_some, max}                                 */
} p2g_vk_config;
int p2g_vk_bytes(const p2g_circuit* c, const p2g_vk_config* cfg, uint8_t* out, size_t* out_len);

/* Upper bound of the proof size; also what p2g_prove* return in *out_len (with P2G_ESMALLBUF) when out == NULL. */
size_t p2g_proof_size_bound(const p2g_circuit* c);

/* ---- multi-GPU (one process per GPU): coset sharding, SURVEY 8(e) ----------------------------------------------
 * One proof across `world` GPUs (a power of two <= 2^min(rate_bits, cap_height)).  Rank r owns the leaves
 * [r * 8N/world, (r+1) * 8N/world) of every oracle = whole LDE cosets = whole Merkle-cap subtrees: their LDE, leaf and
 * subtree hashing, their share of the quotient evaluation and of the query openings.  `allgather` is called by the
 * library (same sequence on every rank) whenever ranks exchange data -- inverse-NTT column blocks, Merkle subtree caps,
 * quotient values, opened rows; the host binds it to NCCL (torch.distributed / ncclAllGather).  When every rank can map every
 * other rank's coefficient buffer (CUDA IPC across processes, plain pointers inside one process) the largest exchange -- the
 * inverse-NTT column blocks of the trace -- does not go through the callback at all: the final inverse-NTT pass stores its
 * results straight into the peers over NVLink (P2G_BUF_SHARD_INFO reports it; P2G_NO_PEER=1 disables it).  It must gather `bytes`
 * from every rank into recv (world * bytes, rank order); send may alias recv + rank * bytes (in-place).  is_device:
 * both buffers are device memory on the handle's device, and all prior work on them has completed.  Return 0 on success.
 * Every rank passes the same desc, the same wires and public inputs to p2g_prove; all ranks return the same bytes. */
typedef int (*p2g_allgather_fn)(void* user, const void* send, void* recv, size_t bytes, int is_device);
int p2g_circuit_create_sharded(const p2g_circuit_desc* desc, int device, int rank, int world, p2g_allgather_fn allgather,
                               void* user, p2g_circuit** out);

/* The same with the communicator inside the library (SURVEY 8b "Threading"): no callback, the exchanges are ncclAllGather calls
 * enqueued on the handle's own stream, with no host round trip for device data.  Rank 0 calls p2g_nccl_unique_id and hands
 * the 128 bytes to the other ranks by any means (MPI, torch.distributed, a file); then every rank -- one process or one thread per
 * GPU -- calls p2g_circuit_create_sharded_nccl (ncclCommInitRank; collective: returns when all ranks have joined).  libnccl.so.2
 * is resolved at run time (the copy already loaded in the process, e.g. PyTorch's, else the system one; P2G_NCCL_LIB overrides).
 * NCCL failures and a peer that never arrives (P2G_NCCL_TIMEOUT_S, default 600 s) return P2G_ENCCL and poison the handle. */
#define P2G_NCCL_UNIQUE_ID_BYTES 128
int p2g_nccl_unique_id(uint8_t* id_out /* [P2G_NCCL_UNIQUE_ID_BYTES] */);
int p2g_circuit_create_sharded_nccl(const p2g_circuit_desc* desc, int device, int rank, int world,
                                    const uint8_t* nccl_id /* [P2G_NCCL_UNIQUE_ID_BYTES] */, p2g_circuit** out);

/* ---- intermediates of the last prove, for parity tests ---------------------------------------------------------- */
enum p2g_buffer {
    P2G_BUF_WIRES_CAP = 0,      /* 2^cap_height digests                                                               */
    P2G_BUF_ZS_PP_CAP = 1,
    P2G_BUF_QUOTIENT_CAP = 2,
    P2G_BUF_CS_CAP = 3,
    P2G_BUF_ZS_PP_VALUES = 4,   /* (num_challenges*(1+num_partial_products)) x N u64, col-major, natural row order     */
    P2G_BUF_QUOTIENT_CHUNKS = 5,/* (num_challenges*qdf) x N u64 coefficients                                           */
    P2G_BUF_WIRES_COEFFS = 6,   /* num_wires x N u64 coefficients                                                      */
    P2G_BUF_CHALLENGES = 7,     /* u64: betas[nc], gammas[nc], alphas[nc], zeta[2], fri_alpha[2], fri_betas[2*L], pow_witness, indices[q] */
    P2G_BUF_FINAL_POLY = 8,     /* ext coefficients (2 u64 each)                                                       */
    P2G_BUF_FRI_CAPS = 9,       /* num_fri_layers x 2^cap_height digests                                               */
    P2G_BUF_WIRES_LDE = 10,     /* num_wires x 8N/world u64, col-major, leaf (bit-reversed) order: this rank's leaves    */
    P2G_BUF_SHARD_INFO = 11     /* u64: rank, world, mapped peers, 1 if the inverse-NTT column exchange runs over peer memory,
                                   1 if the communicator is the library's own NCCL one, exchanges since create */
};
int p2g_circuit_read(p2g_circuit* c, int what, void* out, size_t* len /* in: capacity, out: bytes */);

/* ---- stand-alone kernels, host buffers in/out (parity tests + micro-benchmarks) ------------------------------- */
/* values on <omega_N> (natural) -> coefficients (natural); ncols columns, col-major.  plonky2_field fft.rs ifft */
int p2g_ifft(const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t ncols, int device);
/* coefficients -> values on the coset shift*<omega_{N<<rate_bits}>, output col-major [col][N<<rate_bits] in LEAF order
 * (index j holds the point shift*omega^{bitrev(j)}): PolynomialBatch::lde_values + reverse_index_bits_in_place */
int p2g_lde(const uint64_t* coeffs, uint64_t* lde, uint32_t log_n, uint32_t rate_bits, uint32_t ncols, int device);
/* inverse of p2g_lde with rate_bits = 0 and arbitrary size: leaf-order coset values -> coefficients (coset_ifft) */
int p2g_coset_ifft_leaforder(const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t ncols, int device);
/* MerkleTree::new over `nleaves` leaves of `ncols` u64 (col-major input [col][nleaves]); writes the 2^cap_height cap
 * and, if digests_out != NULL, all leaf digests (nleaves * hash_size bytes). */
int p2g_merkle_cap(const uint64_t* leaves_colmajor, uint32_t log_leaves, uint32_t ncols, uint32_t cap_height,
                   uint32_t hasher, uint8_t* cap_out, uint8_t* digests_out, int device);
/* n independent Poseidon permutations of 12-element states */
int p2g_poseidon_permute(const uint64_t* in, uint64_t* out, size_t n, int device);
/* n independent Keccak-256 of fixed-length messages (msg_len bytes each), 32-byte digests */
int p2g_keccak256(const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* out, int device);
/* filtered gate constraints at npoints points: constants [num_constants][npoints], wires [num_wires][npoints],
 * out [num_gate_constraints][npoints] = sum_g filter_g * constraint_{g,k}  (evaluate_gate_constraints_base_batch) */
int p2g_eval_gate_constraints(const p2g_circuit_desc* desc, const uint64_t* constants, const uint64_t* wires,
                              const uint64_t* pi_hash, size_t npoints, uint64_t* out, int device);

/* field-arithmetic self-test of the device (device >= 0) or host (device < 0) forms of csrc/gl.cuh; see ctx.cu for `op` */
int p2g_test_field_ops(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, int device);

#ifdef __cplusplus
}
#endif
#endif /* P2G_H */
