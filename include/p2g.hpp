// p2g.hpp -- C++ host-side mirror of the reference interface for the hot path, on top of the C ABI in p2g.h.
//
// The reference reaches the prover through plonky2's `CircuitData`:
//     builder.build::<C>()                    /root/reference/plonky2-backend/src/circuit_translation/mod.rs:81
//     circuit_data.prove(witnesses).unwrap()  /root/reference/plonky2-backend/src/actions/prove_action.rs:96
//     proof.compress(..), to_bytes()          /root/reference/plonky2-backend/src/actions/prove_action.rs:75-78
// This header keeps those names -- CircuitConfig (wide_ecc_config, mod.rs:69), Gate, CommonCircuitData, CircuitData::prove,
// ProofWithPublicInputs::to_bytes -- with plonky2's argument meaning and the reference's error behaviour (it `.unwrap()`s, so
// failures throw).  It derives what plonky2's builder derives (gate ordering, selector groups, FRI schedule, k_is) and calls
// libp2g; it computes nothing of the proof itself.  Header-only, C++17; the Python mirror (circuit.py) is checked against it.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "p2g.h"

namespace p2g {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("libp2g error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != P2G_OK) throw Error(rc, p2g_last_error());
}

constexpr uint64_t GOLDILOCKS_P = 0xFFFFFFFF00000001ULL;
constexpr uint64_t MULTIPLICATIVE_GROUP_GENERATOR = 14293326489335486720ULL;
inline uint64_t mul_mod_p(uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % GOLDILOCKS_P); }

// plonky2 `CircuitConfig` + `FriConfig`: the fields the prover reads
struct CircuitConfig {
    uint32_t num_wires = 135, num_routed_wires = 80, num_constants = 2, num_challenges = 2, max_quotient_degree_factor = 8;
    uint32_t rate_bits = 3, cap_height = 4, proof_of_work_bits = 16, num_query_rounds = 28;
    uint32_t fri_arity_bits = 4, fri_final_poly_bits = 5;   // FriReductionStrategy::ConstantArityBits(4, 5)
    uint32_t hasher = P2G_HASH_KECCAK25;                     // plonky2-backend/src/lib.rs:13: C = KeccakGoldilocksConfig
    static CircuitConfig standard_recursion_config() { return CircuitConfig(); }
    static CircuitConfig standard_ecc_config() {
        CircuitConfig c;
        c.num_wires = 136;
        return c;
    }
    static CircuitConfig wide_ecc_config() {   // what the backend builds every circuit with (circuit_translation/mod.rs:69)
        CircuitConfig c;
        c.num_wires = 234;
        return c;
    }
};

// One of the 12 gate kinds of p2g.h with its parameters (the integers BackendGateSerializer writes, write_vk_action.rs:35-62)
struct Gate {
    uint32_t kind = P2G_GATE_NOOP;
    uint32_t params[4] = {0, 0, 0, 0};
    Gate() {}
    Gate(uint32_t k, uint32_t a = 0, uint32_t b = 0, uint32_t c = 0, uint32_t d = 0) : kind(k), params{a, b, c, d} {}

    static uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
    uint32_t degree() const {
        const uint32_t* p = params;
        switch (kind) {
        case P2G_GATE_NOOP: return 0;
        case P2G_GATE_CONSTANT: case P2G_GATE_PUBLIC_INPUT: return 1;
        case P2G_GATE_ARITHMETIC: return 3;
        case P2G_GATE_BASE_SUM: return p[0];
        case P2G_GATE_POSEIDON: return 7;
        case P2G_GATE_RANDOM_ACCESS: return p[0] + 1;
        case P2G_GATE_U32_ARITHMETIC: case P2G_GATE_U32_ADD_MANY: case P2G_GATE_U32_SUBTRACTION: case P2G_GATE_U32_RANGE_CHECK: return 4;
        case P2G_GATE_COMPARISON: return 1u << cdiv(p[0], p[1]);
        default: throw std::invalid_argument("unknown gate kind");
        }
    }
    uint32_t num_constraints() const {
        const uint32_t* p = params;
        switch (kind) {
        case P2G_GATE_NOOP: return 0;
        case P2G_GATE_CONSTANT: return p[0];
        case P2G_GATE_PUBLIC_INPUT: return 4;
        case P2G_GATE_ARITHMETIC: return p[0];
        case P2G_GATE_BASE_SUM: return 1 + p[1];
        case P2G_GATE_POSEIDON: return 123;
        case P2G_GATE_RANDOM_ACCESS: return (p[0] + 2) * p[1] + p[2];
        case P2G_GATE_U32_ARITHMETIC: return 36 * p[0];      // plonky2_ecdsa/biguint/gates/arithmetic_u32.rs
        case P2G_GATE_U32_ADD_MANY: return 21 * p[1];        // add_many_u32.rs
        case P2G_GATE_U32_SUBTRACTION: return 19 * p[0];     // subtraction_u32.rs
        case P2G_GATE_U32_RANGE_CHECK: return 17 * p[0];     // range_check_u32.rs
        case P2G_GATE_COMPARISON: return 6 + 5 * p[1] + cdiv(p[0], p[1]);   // comparison.rs
        default: throw std::invalid_argument("unknown gate kind");
        }
    }
    uint32_t num_constants() const {
        switch (kind) {
        case P2G_GATE_CONSTANT: return params[0];
        case P2G_GATE_ARITHMETIC: return 2;
        case P2G_GATE_RANDOM_ACCESS: return params[2];
        default: return 0;
        }
    }
    // plonky2 `Gate::id()` (the Debug rendering): gates are ordered by (degree, id)
    std::string id() const {
        const std::string GF = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";
        auto s = [](uint32_t x) { return std::to_string(x); };
        const uint32_t* p = params;
        switch (kind) {
        case P2G_GATE_NOOP: return "NoopGate";
        case P2G_GATE_CONSTANT: return "ConstantGate { num_consts: " + s(p[0]) + " }";
        case P2G_GATE_PUBLIC_INPUT: return "PublicInputGate";
        case P2G_GATE_ARITHMETIC: return "ArithmeticGate { num_ops: " + s(p[0]) + " }";
        case P2G_GATE_BASE_SUM: return "BaseSumGate { num_limbs: " + s(p[1]) + " } + Base: " + s(p[0]);
        case P2G_GATE_POSEIDON: return "PoseidonGate(" + GF + ")<WIDTH=12>";
        case P2G_GATE_RANDOM_ACCESS:
            return "RandomAccessGate { bits: " + s(p[0]) + ", num_copies: " + s(p[1]) + ", num_extra_constants: " + s(p[2]) +
                   ", _phantom: " + GF + " }<D=2>";
        case P2G_GATE_U32_ARITHMETIC: return "U32ArithmeticGate { num_ops: " + s(p[0]) + ", _phantom: " + GF + " }";
        case P2G_GATE_U32_ADD_MANY:
            return "U32AddManyGate { num_addends: " + s(p[0]) + ", num_ops: " + s(p[1]) + ", _phantom: " + GF + " }";
        case P2G_GATE_U32_SUBTRACTION: return "U32SubtractionGate { num_ops: " + s(p[0]) + ", _phantom: " + GF + " }";
        case P2G_GATE_U32_RANGE_CHECK: return "U32RangeCheckGate { num_input_limbs: " + s(p[0]) + ", _phantom: " + GF + " }";
        case P2G_GATE_COMPARISON:
            return "ComparisonGate { num_bits: " + s(p[0]) + ", num_chunks: " + s(p[1]) + ", _phantom: " + GF + " }<D=2>";
        default: throw std::invalid_argument("unknown gate kind");
        }
    }
    bool operator==(const Gate& o) const { return kind == o.kind && std::equal(params, params + 4, o.params); }

    // constructors with the reference's `new_from_config` arithmetic
    static Gate noop() { return Gate(P2G_GATE_NOOP); }
    static Gate constant(const CircuitConfig& c) { return Gate(P2G_GATE_CONSTANT, c.num_constants); }
    static Gate public_input() { return Gate(P2G_GATE_PUBLIC_INPUT); }
    static Gate arithmetic(const CircuitConfig& c) { return Gate(P2G_GATE_ARITHMETIC, c.num_routed_wires / 4); }
    static Gate base_sum(uint32_t base, uint32_t num_limbs) { return Gate(P2G_GATE_BASE_SUM, base, num_limbs); }
    static Gate poseidon() { return Gate(P2G_GATE_POSEIDON); }
    static Gate random_access(const CircuitConfig& c, uint32_t bits) {
        uint32_t vec = 1u << bits;
        uint32_t copies = std::min(c.num_routed_wires / (2 + vec), c.num_wires / (2 + vec + bits));
        uint32_t extra = std::min(c.num_routed_wires - (2 + vec) * copies, c.num_constants);
        return Gate(P2G_GATE_RANDOM_ACCESS, bits, copies, extra);
    }
    static Gate u32_arithmetic(const CircuitConfig& c) {   // arithmetic_u32.rs:40-43
        return Gate(P2G_GATE_U32_ARITHMETIC, std::min(c.num_wires / 38, c.num_routed_wires / 6));
    }
    static Gate u32_add_many(const CircuitConfig& c, uint32_t num_addends) {   // add_many_u32.rs:43-48
        return Gate(P2G_GATE_U32_ADD_MANY, num_addends, std::min(c.num_wires / (num_addends + 21), c.num_routed_wires / (num_addends + 3)));
    }
    static Gate u32_subtraction(const CircuitConfig& c) {   // subtraction_u32.rs:38-42
        return Gate(P2G_GATE_U32_SUBTRACTION, std::min(c.num_wires / 21, c.num_routed_wires / 5));
    }
    static Gate u32_range_check(uint32_t num_input_limbs) { return Gate(P2G_GATE_U32_RANGE_CHECK, num_input_limbs); }
    static Gate comparison(uint32_t num_bits = 32, uint32_t num_chunks = 16) { return Gate(P2G_GATE_COMPARISON, num_bits, num_chunks); }
};

// plonky2 `CommonCircuitData`: everything about the circuit that is independent of the witness
class CommonCircuitData {
  public:
    CircuitConfig config;
    uint32_t degree_bits = 0, num_public_inputs = 0, quotient_degree_factor = 8;
    std::vector<Gate> gates;                                  // sorted by (degree, id), deduplicated
    std::vector<uint32_t> selector_indices;                   // selectors_info.selector_indices
    std::vector<std::pair<uint32_t, uint32_t>> groups;        // selectors_info.groups [start, end)
    uint32_t num_selectors = 0, num_constants = 0, num_gate_constraints = 0, num_partial_products = 0;
    std::vector<uint64_t> k_is;
    std::vector<uint32_t> reduction_arity_bits;

    CommonCircuitData(const CircuitConfig& cfg, uint32_t degree_bits_, std::vector<Gate> gate_set, uint32_t num_public_inputs_ = 0)
        : config(cfg), degree_bits(degree_bits_), num_public_inputs(num_public_inputs_), quotient_degree_factor(cfg.max_quotient_degree_factor) {
        if (gate_set.empty()) throw std::invalid_argument("a circuit has at least one gate");
        std::vector<std::pair<std::pair<uint32_t, std::string>, Gate>> keyed;
        for (const Gate& g : gate_set) keyed.push_back({{g.degree(), g.id()}, g});
        std::sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (const auto& kg : keyed)
            if (gates.empty() || !(gates.back() == kg.second)) gates.push_back(kg.second);
        // plonky2 gates/selectors.rs selector_polynomials
        const uint32_t max_degree = quotient_degree_factor + 1, n = (uint32_t)gates.size();
        if (gates.back().degree() + n - 1 <= max_degree) {
            selector_indices.assign(n, 0);
            groups.push_back({0, n});
        } else {
            uint32_t start = 0;
            while (start < n) {
                uint32_t size = 0;
                while (start + size < n && size + gates[start + size].degree() < max_degree) size++;
                if (size == 0) throw std::invalid_argument("gate " + gates[start].id() + " does not fit any selector group");
                groups.push_back({start, start + size});
                start += size;
            }
            for (uint32_t gi = 0; gi < groups.size(); gi++)
                for (uint32_t i = groups[gi].first; i < groups[gi].second; i++) selector_indices.push_back(gi);
        }
        num_selectors = (uint32_t)groups.size();
        uint32_t maxc = 0;
        for (const Gate& g : gates) {
            maxc = std::max(maxc, g.num_constants());
            num_gate_constraints = std::max(num_gate_constraints, g.num_constraints());
        }
        num_constants = num_selectors + maxc;
        num_partial_products = Gate::cdiv(cfg.num_routed_wires, quotient_degree_factor) - 1;
        uint64_t k = 1;
        for (uint32_t i = 0; i < cfg.num_routed_wires; i++) {
            k_is.push_back(k);
            k = mul_mod_p(k, MULTIPLICATIVE_GROUP_GENERATOR);
        }
        uint32_t d = degree_bits;
        while (d > cfg.fri_final_poly_bits && d + cfg.rate_bits >= cfg.cap_height + cfg.fri_arity_bits) {
            reduction_arity_bits.push_back(cfg.fri_arity_bits);
            d -= cfg.fri_arity_bits;
        }
    }
    size_t degree() const { return (size_t)1 << degree_bits; }
    size_t num_preprocessed() const { return num_constants + config.num_routed_wires; }
    size_t hash_size() const { return config.hasher == P2G_HASH_KECCAK25 ? 25 : 32; }
    size_t gate_index(const Gate& g) const {
        auto it = std::find(gates.begin(), gates.end(), g);
        if (it == gates.end()) throw std::invalid_argument("gate not in circuit");
        return (size_t)(it - gates.begin());
    }
    std::vector<p2g_gate> gate_table() const {
        std::vector<p2g_gate> t(gates.size());
        for (size_t i = 0; i < gates.size(); i++) {
            t[i].kind = gates[i].kind;
            std::copy(gates[i].params, gates[i].params + 4, t[i].params);
            t[i].selector_index = selector_indices[i];
            t[i].group_lo = groups[selector_indices[i]].first;
            t[i].group_hi = groups[selector_indices[i]].second;
            t[i].num_constraints = gates[i].num_constraints();
        }
        return t;
    }
};

// Proof bytes in plonky2's wire format (uncompressed `ProofWithPublicInputs::to_bytes` or the compressed layout)
struct ProofWithPublicInputs {
    std::vector<uint8_t> proof_bytes;
    std::vector<uint64_t> public_inputs;
    bool compressed = false;
    p2g_timings timings{};
    const std::vector<uint8_t>& to_bytes() const { return proof_bytes; }
};

// plonky2 `CircuitData`: `prover_only` (preprocessed polynomials, resident on the GPU) + `common`
class CircuitData {
  public:
    CommonCircuitData common;
    std::vector<std::vector<uint8_t>> constants_sigmas_cap;   // verifier_only.constants_sigmas_cap
    std::vector<uint8_t> circuit_digest;                       // verifier_only.circuit_digest

    // constants_sigmas: values on the subgroup, column-major [num_preprocessed][N]: selectors, gate constants, sigmas
    // rank / world / nccl_id: this handle is one rank of a coset-sharded prover whose NCCL communicator lives inside the library
    // (p2g_circuit_create_sharded_nccl; nccl_id = the P2G_NCCL_UNIQUE_ID_BYTES from p2g_nccl_unique_id() on rank 0); world = 1: one GPU.
    CircuitData(const CommonCircuitData& c, const uint64_t* constants_sigmas, int device = 0, int rank = 0, int world = 1,
                const uint8_t* nccl_id = nullptr)
        : common(c) {
        std::vector<p2g_gate> table = common.gate_table();
        p2g_circuit_desc d{};
        d.struct_size = (uint32_t)sizeof d;
        d.degree_bits = common.degree_bits;
        d.num_wires = common.config.num_wires;
        d.num_routed_wires = common.config.num_routed_wires;
        d.num_constants = common.num_constants;
        d.num_selectors = common.num_selectors;
        d.num_challenges = common.config.num_challenges;
        d.rate_bits = common.config.rate_bits;
        d.cap_height = common.config.cap_height;
        d.pow_bits = common.config.proof_of_work_bits;
        d.num_query_rounds = common.config.num_query_rounds;
        d.quotient_degree_factor = common.quotient_degree_factor;
        d.num_partial_products = common.num_partial_products;
        d.num_gate_constraints = common.num_gate_constraints;
        d.num_public_inputs = common.num_public_inputs;
        d.hasher = common.config.hasher;
        d.num_fri_layers = (uint32_t)common.reduction_arity_bits.size();
        for (size_t i = 0; i < common.reduction_arity_bits.size(); i++) d.reduction_arity_bits[i] = common.reduction_arity_bits[i];
        d.num_gates = (uint32_t)table.size();
        d.gates = table.data();
        d.constants_sigmas = constants_sigmas;
        d.k_is = common.k_is.data();
        d.circuit_digest = nullptr;
        check(world > 1 ? p2g_circuit_create_sharded_nccl(&d, device, rank, world, nccl_id, &h_) : p2g_circuit_create(&d, device, &h_));
        const size_t hs = common.hash_size();
        const size_t ncap = (size_t)1 << std::min<uint32_t>(common.config.cap_height, common.degree_bits + common.config.rate_bits);
        std::vector<uint8_t> cap(ncap * hs);
        circuit_digest.resize(hs);
        check(p2g_circuit_cap(h_, cap.data(), cap.size(), circuit_digest.data(), circuit_digest.size()));
        for (size_t i = 0; i < ncap; i++) constants_sigmas_cap.emplace_back(cap.begin() + i * hs, cap.begin() + (i + 1) * hs);
    }
    CircuitData(const CircuitData&) = delete;
    CircuitData& operator=(const CircuitData&) = delete;
    ~CircuitData() { p2g_circuit_destroy(h_); }

    // `wires`: MatrixWitness.wire_values, column-major [num_wires][N], canonical; host memory (pageable or p2g_host_alloc'd).
    // forced_pow_witness: nullptr = smallest valid witness.  compressed: the CLI's final bytes (prove_action.rs:75-78).
    ProofWithPublicInputs prove(const uint64_t* wires, const std::vector<uint64_t>& public_inputs, const uint64_t* forced_pow_witness = nullptr,
                                bool compressed = false) {
        ProofWithPublicInputs pw;
        pw.public_inputs = public_inputs;
        pw.compressed = compressed;
        pw.proof_bytes.resize(p2g_proof_size_bound(h_));
        size_t len = pw.proof_bytes.size();
        int rc = compressed ? p2g_prove_compressed(h_, wires, 0, public_inputs.data(), public_inputs.size(), forced_pow_witness,
                                                   pw.proof_bytes.data(), &len, &pw.timings)
                            : p2g_prove(h_, wires, public_inputs.data(), public_inputs.size(), forced_pow_witness, pw.proof_bytes.data(),
                                        &len, &pw.timings);
        check(rc);
        pw.proof_bytes.resize(len);
        return pw;
    }
    // The witness as plonky2 holds it: one pointer per MatrixWitness.wire_values[col] (N words each, any representative < 2^64).
    ProofWithPublicInputs prove_columns(const std::vector<const uint64_t*>& wire_columns, const std::vector<uint64_t>& public_inputs,
                                        const uint64_t* forced_pow_witness = nullptr, bool compressed = false) {
        if (wire_columns.size() != common.config.num_wires) throw std::invalid_argument("one column pointer per wire expected");
        ProofWithPublicInputs pw;
        pw.public_inputs = public_inputs;
        pw.compressed = compressed;
        pw.proof_bytes.resize(p2g_proof_size_bound(h_));
        size_t len = pw.proof_bytes.size();
        check(p2g_prove_columns(h_, wire_columns.data(), public_inputs.data(), public_inputs.size(), forced_pow_witness, compressed ? 1 : 0,
                                pw.proof_bytes.data(), &len, &pw.timings));
        pw.proof_bytes.resize(len);
        return pw;
    }
    // prove_columns with only the ROUTED columns (the first num_routed_wires of MatrixWitness.wire_values): the advice columns are
    // computed on the device inside the upload pipeline (p2g_prove_routed_columns); same bytes as prove_columns on the full witness
    ProofWithPublicInputs prove_routed_columns(const std::vector<const uint64_t*>& routed_columns, const std::vector<uint64_t>& public_inputs,
                                               const uint64_t* forced_pow_witness = nullptr, bool compressed = false) {
        if (routed_columns.size() != common.config.num_routed_wires) throw std::invalid_argument("one column pointer per routed wire expected");
        ProofWithPublicInputs pw;
        pw.public_inputs = public_inputs;
        pw.compressed = compressed;
        pw.proof_bytes.resize(p2g_proof_size_bound(h_));
        size_t len = pw.proof_bytes.size();
        check(p2g_prove_routed_columns(h_, routed_columns.data(), public_inputs.data(), public_inputs.size(), forced_pow_witness,
                                       compressed ? 1 : 0, pw.proof_bytes.data(), &len, &pw.timings));
        pw.proof_bytes.resize(len);
        return pw;
    }
    // Device-side witness fill: d_wires [num_wires][N] on this circuit's device holds the routed columns; the advice columns
    // (>= num_routed_wires) are computed in place (p2g_fill_advice_device).  Follow with p2g_prove_device.
    void fill_advice_device(uint64_t* d_wires) { check(p2g_fill_advice_device(h_, d_wires)); }
    // verifier_data().to_bytes(&BackendGateSerializer): the file `write_vk` writes (write_vk_action.rs:76-79)
    std::vector<uint8_t> verifier_data_bytes(const p2g_vk_config* cfg = nullptr) const {
        size_t len = 0;
        int rc = p2g_vk_bytes(h_, cfg, nullptr, &len);
        if (rc != P2G_ESMALLBUF) check(rc);
        std::vector<uint8_t> out(len);
        check(p2g_vk_bytes(h_, cfg, out.data(), &len));
        out.resize(len);
        return out;
    }
    p2g_circuit* handle() const { return h_; }

  private:
    p2g_circuit* h_ = nullptr;
};

}  // namespace p2g
