// ACIR -> gate table + witness, outside Rust.
//
// Host-side C++ mirror of the two reference layers that sit above the prover:
//   * the translator  plonky2-backend/src/circuit_translation/mod.rs:72-190 (CircuitBuilderFromAcirToPlonky2::translate_circuit)
//     with assert_zero_translator.rs:30-116 (AssertZero), mod.rs:131-137 (RANGE -> builder.range_check), mod.rs:213-232 +
//     binary_digits_target.rs:120-186 (AND / XOR on bit decompositions), memory_translator.rs:145-171 (MemoryInit) and :125-137
//     (memory read -> RandomAccessGate), :87-113 (memory write), sha256_translator.rs:60-273 (Sha256Compression) and, in ecdsa.h,
//     ecdsa_secp256k1_translator.rs with the whole plonky2_ecdsa gadget stack under it (u32 / biguint / non-native / curve / GLV);
//   * the part of plonky2's CircuitBuilder those translators drive (arithmetic / add / mul / mul_const / sub, constant, connect,
//     assert_zero, split_le, le_sum, random_access, register_public_input, build()) and its witness generators
//     (generate_partial_witness: ArithmeticBaseGenerator, BaseSplitGenerator / BaseSumGenerator, WireSplitGenerator,
//     RandomAccessGenerator, PoseidonGenerator, EqualityGenerator), plus the generators of the reference's five custom gates and
//     of its big-integer gadgets.  The C ABI is declared in include/p2acir.h.
// It emits exactly what crosses the C ABI of include/p2g.h: the gate list, the per-row gate assignment + gate constants, the
// sigma permutation of the copy constraints, and -- from the ACIR witness map -- the full wire matrix and the public inputs.
//
// What it is NOT: plonky2's builder bit for bit.  Gate packing (which op lands in which row) follows the same rules
// (operations with equal constants share a row, 20 ops per ArithmeticGate, ConstantGate rows of 2, BaseSumGate<2> of 63 limbs
// for split_le, public-input hash through PoseidonGate rows connected to a PublicInputGate) but is not guaranteed to coincide
// with the fork's, so its circuit_digest differs from the Rust one; every circuit it builds is a valid plonky2 circuit in the
// reference's configuration, and proofs of it verify.  Host tooling, not on the hot path.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../csrc/hash.cuh"
#include "../../include/p2g.h"
#include "../csrc/advice.cuh"
#include "bigint.h"
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

typedef int32_t Target;   // >= 0: index into the target table (virtual targets and wires alike)

struct GateType {
    u32 kind;
    u32 params[4];
    bool operator<(const GateType& o) const { return std::tie(kind, params[0], params[1], params[2], params[3]) < std::tie(o.kind, o.params[0], o.params[1], o.params[2], o.params[3]); }
    bool operator==(const GateType& o) const { return !(*this < o) && !(o < *this); }
};

struct Row {
    int gate;                 // index into Builder::gate_types
    std::vector<u64> consts;  // gate-local constants
};

enum GenKind { GEN_ARITH, GEN_CONST, GEN_SPLIT, GEN_BASE_SPLIT, GEN_BASE_SUM, GEN_RANDOM_ACCESS, GEN_POSEIDON, GEN_EQUAL,
               GEN_U32_ARITH, GEN_U32_ADD_MANY, GEN_U32_SUB, GEN_U32_RANGE, GEN_COMPARISON, GEN_BIG };
struct Gen {
    GenKind kind;
    int row, i;         // gate row, op / copy index
    u64 c0, c1;         // arithmetic constants / constant value
    Target t;           // GEN_SPLIT: the integer; GEN_EQUAL: x
    Target t2, t3, t4;  // GEN_EQUAL: y, equal, inv
    std::vector<int> rows;   // GEN_SPLIT: the BaseSum rows
    int n;              // limbs / bits
    int group = 0;      // generators of one group run in creation order on one thread
    bool done = false;
};

struct Error {
    std::string msg;
};

// generators over big integers (plonky2_ecdsa: BigUintDivRemGenerator, NonNative{Addition,Subtraction,Multiplication,Inverse}Generator,
// GLVDecompositionGenerator): kept in a side table, Gen::i indexes it
typedef std::vector<Target> BigT;   // BigUintTarget: u32 limbs, least significant first
enum BigGenKind { BG_DIVREM, BG_NN_ADD, BG_NN_ADD_MANY, BG_NN_SUB, BG_NN_MUL, BG_NN_INV, BG_GLV };
struct BigGen {
    BigGenKind kind;
    int field;        // 0: secp256k1 base field, 1: scalar field
    BigT a, b;        // inputs
    BigT o1, o2;      // outputs (sum / diff / prod / inv / div / k1, overflow / div / rem / k2)
    Target t1, t2;    // bool outputs (overflow; k1_neg, k2_neg)
};

const int NUM_WIRES = 234, NUM_ROUTED = 80, ARITH_OPS = 20, CONSTS_PER_GATE = 2, SPLIT_LIMBS = 63;

struct Builder {
    // targets: routed wires become targets lazily, (row, col) -> id; the advice wires (col >= 80) can never be copy-constrained, so
    // they are not targets at all: wire() names them by a negative code and witness generation keeps their values in a dense table
    std::vector<Target> routed_target;               // [row * 80 + col] -> target or -1
    std::vector<Target> parent;                      // union-find over targets (copy constraints)
    std::vector<Row> rows;
    std::vector<GateType> gate_types;
    std::map<u64, Target> constants;                 // value -> target
    std::vector<u64> constant_order;
    std::map<std::tuple<u64, u64>, std::pair<int, int>> free_arith;   // (c0, c1) -> (row, next op)
    struct ArithKey {
        u64 c0, c1;
        Target x, y, z;
        bool operator==(const ArithKey& o) const { return c0 == o.c0 && c1 == o.c1 && x == o.x && y == o.y && z == o.z; }
    };
    struct ArithKeyHash {
        size_t operator()(const ArithKey& k) const {   // multiply-xorshift rounds over the five fields
            u64 h = ((u64)(u32)k.x << 32 | (u32)k.y) * 0x9E3779B97F4A7C15ULL;
            h ^= h >> 32;
            h = (h + (u32)k.z) * 0xD6E8FEB86659FD93ULL;
            h ^= h >> 32;
            h = (h + k.c0 + (k.c1 << 1 | k.c1 >> 63)) * 0xC2B2AE3D27D4EB4FULL;
            return (size_t)(h ^ (h >> 29));
        }
    };
    struct ArithCache {   // open addressing, linear probing: one cache line per lookup instead of a node chase (millions of entries)
        struct Slot {
            ArithKey key;
            Target out;   // -1 = empty
        };
        std::vector<Slot> slots;
        size_t count = 0;
        Slot* locate(const ArithKey& k) {   // the slot holding k, or the empty slot where it belongs
            if (slots.empty()) slots.assign(1 << 12, Slot{{0, 0, 0, 0, 0}, -1});
            const size_t mask = slots.size() - 1;
            size_t i = ArithKeyHash()(k) & mask;
            while (slots[i].out >= 0 && !(slots[i].key == k)) i = (i + 1) & mask;
            return &slots[i];
        }
        void insert(Slot* at, const ArithKey& k, Target out) {
            at->key = k;
            at->out = out;
            if (++count * 5 < slots.size() * 3) return;
            std::vector<Slot> old;
            old.swap(slots);
            slots.assign(old.size() * 2, Slot{{0, 0, 0, 0, 0}, -1});
            for (const Slot& s : old)
                if (s.out >= 0) *locate(s.key) = s;
        }
    } arith_cache;
    std::map<int, std::pair<int, int>> free_ra;      // bits -> (row, next copy)
    std::pair<int, int> free_u32_arith = {0, 0}, free_u32_sub = {0, 0};   // (row, next op)
    std::map<int, std::pair<int, int>> free_add_many;                     // num_addends -> (row, next op)
    std::vector<Target> public_inputs;
    std::vector<Gen> gens;
    int cur_group = 0, num_groups = 1;   // witness-generation groups: 0 = light opcodes and build(); heavy opcodes open their own
    std::vector<std::vector<int>> group_deps = {{}};   // groups (earlier ones) whose generators must have run first
    int new_group(std::vector<int> deps = {}) {
        group_deps.push_back(std::move(deps));
        return num_groups++;
    }
    void push_gen(Gen g) {
        g.group = cur_group;
        gens.push_back(std::move(g));
    }
    std::vector<BigGen> big_gens;
    std::unordered_map<Target, u64> target_consts;   // plonky2's targets_to_constants (target_as_constant)
    bool built = false;
    int degree_bits = 0;
    Target pi_hash[4];

    Target new_target(int row = -1, int col = -1) {
        parent.push_back((Target)parent.size());
        return (Target)parent.size() - 1;
    }
    Target add_virtual_target() { return new_target(); }
    static Target advice_code(int row, int col) { return -2 - (Target)(row * NUM_WIRES + col); }
    Target wire(int row, int col) {
        if (col >= NUM_ROUTED) return advice_code(row, col);
        Target& slot = routed_target[(size_t)row * NUM_ROUTED + col];
        if (slot < 0) slot = new_target(row, col);
        return slot;
    }
    Target routed_or_none(size_t row, int col) const { return row < rows.size() ? routed_target[row * NUM_ROUTED + col] : -1; }
    Target find(Target t) {
        while (parent[t] != t) {
            parent[t] = parent[parent[t]];
            t = parent[t];
        }
        return t;
    }
    void connect(Target a, Target b) {
        if (a < 0 || b < 0) throw Error{"connect: wire is not routable"};
        a = find(a);
        b = find(b);
        if (a != b) parent[b] = a;
    }
    int gate_type(u32 kind, u32 p0 = 0, u32 p1 = 0, u32 p2 = 0) {
        GateType g = {kind, {p0, p1, p2, 0}};
        for (size_t i = 0; i < gate_types.size(); i++)
            if (gate_types[i] == g) return (int)i;
        gate_types.push_back(g);
        return (int)gate_types.size() - 1;
    }
    int add_gate(int gt, std::vector<u64> consts = {}) {
        rows.push_back({gt, std::move(consts)});
        if (rows.size() > (size_t)1 << 22) throw Error{"circuit has more than 2^22 rows"};
        if (routed_target.size() < rows.size() * NUM_ROUTED) routed_target.resize((rows.size() + 4095) * NUM_ROUTED, -1);
        return (int)rows.size() - 1;
    }

    // ---- constants (plonky2 CircuitBuilder::constant: one virtual target per distinct value, placed in ConstantGate rows by build())
    Target constant(u64 c) {
        auto it = constants.find(c);
        if (it != constants.end()) return it->second;
        Target t = add_virtual_target();
        constants[c] = t;
        constant_order.push_back(c);
        target_consts[t] = c;
        return t;
    }
    Target zero() { return constant(0); }
    Target one() { return constant(1); }
    // ---- arithmetic: c0 * x * y + c1 * z   (gadgets/arithmetic.rs arithmetic / add_base_arithmetic_operation)
    // plonky2 gadgets/arithmetic.rs arithmetic_special_cases: results that need no ArithmeticGate operation
    bool arithmetic_special_case(u64 c0, u64 c1, Target x, Target y, Target z, Target* out) {
        const Target zero_t = zero();
        u64 xc = 0, yc = 0, zc = 0;
        const bool x_const = as_constant(x, &xc), y_const = as_constant(y, &yc), z_const = as_constant(z, &zc);
        const bool first_zero = c0 == 0 || x == zero_t || y == zero_t, second_zero = c1 == 0 || z == zero_t;
        // both terms constant: their (constant) sum
        bool first_known = first_zero, second_known = second_zero;
        u64 first = 0, second = 0;
        if (!first_zero && x_const && y_const) {
            first_known = true;
            first = gl_mul(gl_mul(xc, yc), c0);
        }
        if (!second_zero && z_const) {
            second_known = true;
            second = gl_mul(zc, c1);
        }
        if (first_known && second_known) {
            *out = constant(gl_add(first, second));
            return true;
        }
        if (first_zero && c1 == 1) {
            *out = z;
            return true;
        }
        if (second_zero) {
            if (x_const && gl_mul(xc, c0) == 1) {
                *out = y;
                return true;
            }
            if (y_const && gl_mul(yc, c0) == 1) {
                *out = x;
                return true;
            }
        }
        return false;
    }
    Target arithmetic(u64 c0, u64 c1, Target x, Target y, Target z) {
        Target special;
        if (arithmetic_special_case(c0, c1, x, y, z, &special)) return special;
        // plonky2's base_arithmetic_results: the same operation (same constants, same operands in the same order) is computed once
        const ArithKey key = {c0, c1, find(x), find(y), find(z)};
        if (ArithCache::Slot* hit = arith_cache.locate(key); hit->out >= 0) return hit->out;
        auto& slot = free_arith[std::make_tuple(c0, c1)];
        if (slot.second == 0 || slot.second >= ARITH_OPS) {   // no open row for these constants
            slot.first = add_gate(gate_type(P2G_GATE_ARITHMETIC, ARITH_OPS), {c0, c1});
            slot.second = 0;
        }
        const int row = slot.first, i = slot.second++;
        connect(x, wire(row, 4 * i));
        connect(y, wire(row, 4 * i + 1));
        connect(z, wire(row, 4 * i + 2));
        Target out = wire(row, 4 * i + 3);
        Gen g = {};
        g.kind = GEN_ARITH;
        g.row = row;
        g.i = i;
        g.c0 = c0;
        g.c1 = c1;
        push_gen(g);
        arith_cache.insert(arith_cache.locate(key), key, out);
        return out;
    }
    Target mul(Target x, Target y) { return arithmetic(1, 0, x, y, x); }
    Target add(Target x, Target y) { return arithmetic(1, 1, x, one(), y); }
    Target sub(Target x, Target y) { return arithmetic(1, GL_P - 1, x, one(), y); }
    Target mul_const(u64 c, Target x) { return mul(constant(c), x); }
    Target mul_const_add(u64 c, Target x, Target y) { return arithmetic(c, 1, x, one(), y); }
    void assert_zero(Target x) { connect(x, zero()); }
    // bool gadgets (gadgets/arithmetic.rs and / or / not)
    Target b_and(Target a, Target b) { return mul(a, b); }
    Target b_or(Target a, Target b) { return add(arithmetic(GL_P - 1, 1, a, b, a), b); }
    Target b_not(Target a) { return sub(one(), a); }
    Target b_xor(Target a, Target b) {   // binary_digits_target.rs:169-175: (a or b) and not (a and b)
        return b_and(b_or(a, b), b_not(b_and(a, b)));
    }
    Target mul_sub(Target x, Target y, Target z) { return arithmetic(1, GL_P - 1, x, y, z); }
    Target select(Target b, Target x, Target y) {   // gadgets/select.rs: b x - (b y - y)
        return mul_sub(b, x, mul_sub(b, y, y));
    }
    // is_equal (gadgets/arithmetic.rs): equal * (x - y) = 0 and 1 - equal = (x - y) * inv, with EqualityGenerator filling equal / inv
    Target is_equal(Target x, Target y) {
        Target equal = add_virtual_target(), not_equal = b_not(equal), inv = add_virtual_target();
        Gen g = {};
        g.kind = GEN_EQUAL;
        g.t = x;
        g.t2 = y;
        g.t3 = equal;
        g.t4 = inv;
        push_gen(g);
        Target diff = sub(x, y);
        Target not_equal_check = mul(equal, diff);
        Target diff_normalized = mul(diff, inv);
        connect(not_equal, diff_normalized);
        connect(not_equal_check, zero());
        return equal;
    }
    Target add_virtual_bool_target_safe() {   // a fresh target constrained to {0, 1}: b b - b = 0
        Target t = add_virtual_target();
        connect(mul_sub(t, t, t), zero());
        return t;
    }

    // ---- split_le (gadgets/split_base.rs / split_join.rs): BaseSumGate<2> rows of 63 limbs, unused limbs tied to zero
    std::vector<Target> split_le(Target x, int nbits) {
        std::vector<Target> bits;
        if (nbits == 0) return bits;
        const int k = (nbits + SPLIT_LIMBS - 1) / SPLIT_LIMBS;
        std::vector<int> grows;
        for (int g = 0; g < k; g++) grows.push_back(add_gate(gate_type(P2G_GATE_BASE_SUM, 2, SPLIT_LIMBS)));
        for (int g : grows)
            for (int l = 0; l < SPLIT_LIMBS; l++) bits.push_back(wire(g, 1 + l));
        for (size_t b = nbits; b < bits.size(); b++) assert_zero(bits[b]);
        bits.resize(nbits);
        u64 base = gl_pow(2, SPLIT_LIMBS);
        Target acc = zero();
        for (int g = k - 1; g >= 0; g--) acc = mul_const_add(base, acc, wire(grows[g], 0));
        connect(acc, x);
        Gen sg = {};
        sg.kind = GEN_SPLIT;
        sg.t = x;
        sg.rows = grows;
        sg.n = SPLIT_LIMBS;
        push_gen(sg);
        for (int g : grows) {
            Gen bg = {};
            bg.kind = GEN_BASE_SPLIT;
            bg.row = g;
            bg.n = SPLIT_LIMBS;
            push_gen(bg);
        }
        return bits;
    }
    void range_check(Target x, int nbits) { split_le(x, nbits); }
    // split_le_base::<4>(x, num_limbs) (gadgets/split_base.rs): one BaseSumGate<4>, the sum wire tied to x
    std::vector<Target> split_le_base4(Target x, int num_limbs) {
        const int row = add_gate(gate_type(P2G_GATE_BASE_SUM, 4, (u32)num_limbs));
        connect(x, wire(row, 0));
        Gen g = {};
        g.kind = GEN_BASE_SPLIT;
        g.row = row;
        g.n = num_limbs;
        g.c0 = 4;
        push_gen(g);
        std::vector<Target> limbs;
        for (int l = 0; l < num_limbs; l++) limbs.push_back(wire(row, 1 + l));
        return limbs;
    }
    // le_sum (gadgets/split_base.rs): one BaseSumGate<2> with exactly bits.size() limbs
    Target le_sum(const std::vector<Target>& bits) {
        if (bits.empty()) return zero();
        if ((int)bits.size() > SPLIT_LIMBS) throw Error{"le_sum: more than 63 bits"};
        const int row = add_gate(gate_type(P2G_GATE_BASE_SUM, 2, (u32)bits.size()));
        for (size_t i = 0; i < bits.size(); i++) connect(bits[i], wire(row, 1 + (int)i));
        Gen g = {};
        g.kind = GEN_BASE_SUM;
        g.row = row;
        g.n = (int)bits.size();
        push_gen(g);
        return wire(row, 0);
    }
    // random_access (gadgets/random_access.rs): RandomAccessGate::new_from_config(bits): copies = min(routed / (2 + 2^bits), ...)
    static int ra_copies(int bits) {
        const int vec = 1 << bits;
        int by_routed = NUM_ROUTED / (2 + vec), by_wires = NUM_WIRES / (2 + vec + bits);
        return std::max(1, std::min(by_routed, by_wires));
    }
    Target random_access(Target index, const std::vector<Target>& v) {
        int bits = 0;
        while ((1 << bits) < (int)v.size()) bits++;
        if ((size_t)(1 << bits) != v.size()) throw Error{"random_access: length must be a power of two"};
        if (bits == 0) return v[0];
        if (bits > 6 || 2 + (1 << bits) > NUM_ROUTED) throw Error{"random_access: vector too long for one gate"};
        const int copies = ra_copies(bits), vec = 1 << bits;
        auto& slot = free_ra[bits];
        if (slot.second == 0 || slot.second >= copies) {
            slot.first = add_gate(gate_type(P2G_GATE_RANDOM_ACCESS, bits, copies, 0));
            slot.second = 0;
        }
        const int row = slot.first, cp = slot.second++;
        const int base = (2 + vec) * cp;
        connect(index, wire(row, base));
        for (int i = 0; i < vec; i++) connect(v[i], wire(row, base + 2 + i));
        Gen g = {};
        g.kind = GEN_RANDOM_ACCESS;
        g.row = row;
        g.i = cp;
        g.n = bits;
        push_gen(g);
        return wire(row, base + 1);
    }
    // ---- the reference's u32 gadgets (plonky2-backend/src/plonky2_ecdsa/biguint/gadgets/{arithmetic_u32,range_check,multiple_comparison}.rs)
    // over its custom gates; gate shapes from `new_from_config` with wide_ecc_config (234 wires, 80 routed)
    static int u32_arith_ops() { return std::min(NUM_WIRES / 38, NUM_ROUTED / 6); }                       // arithmetic_u32.rs:40-43
    static int u32_sub_ops() { return std::min(NUM_WIRES / 21, NUM_ROUTED / 5); }                         // subtraction_u32.rs:38-42
    static int add_many_ops(int na) { return std::min(NUM_WIRES / (na + 21), NUM_ROUTED / (na + 3)); }    // add_many_u32.rs:43-48
    // x * y + z = low + 2^32 high
    bool as_constant(Target t, u64* c) const {
        auto it = target_consts.find(t);
        if (it == target_consts.end()) return false;
        *c = it->second;
        return true;
    }
    std::pair<Target, Target> mul_add_u32(Target x, Target y, Target z) {
        u64 cx, cy, cz;
        if (as_constant(x, &cx) && as_constant(y, &cy) && as_constant(z, &cz)) {   // arithmetic_u32_special_cases (gadget :112-137)
            const u64 sum = gl_add(gl_mul(cx, cy), cz);
            return {constant(sum & 0xFFFFFFFFULL), constant(sum >> 32)};
        }
        const int ops = u32_arith_ops();
        auto& slot = free_u32_arith;
        if (slot.second == 0 || slot.second >= ops) {
            slot.first = add_gate(gate_type(P2G_GATE_U32_ARITHMETIC, ops));
            slot.second = 0;
        }
        const int row = slot.first, i = slot.second++;
        connect(x, wire(row, 6 * i));
        connect(y, wire(row, 6 * i + 1));
        connect(z, wire(row, 6 * i + 2));
        Gen g = {};
        g.kind = GEN_U32_ARITH;
        g.row = row;
        g.i = i;
        g.n = ops;
        push_gen(g);
        return {wire(row, 6 * i + 3), wire(row, 6 * i + 4)};
    }
    std::pair<Target, Target> add_u32(Target a, Target b) { return mul_add_u32(a, one(), b); }
    // sum of the addends = result + 2^32 carry  (arithmetic_u32.rs gadget add_many_u32: 0 / 1 / 2 addends are special-cased)
    std::pair<Target, Target> mul_u32(Target a, Target b) { return mul_add_u32(a, b, zero()); }
    std::pair<Target, Target> add_many_u32(const std::vector<Target>& v) {
        if (v.empty()) return {zero(), zero()};
        if (v.size() == 1) return {v[0], zero()};
        if (v.size() == 2) return add_u32(v[0], v[1]);
        return add_many_gate(v, zero());
    }
    std::pair<Target, Target> add_u32s_with_carry(const std::vector<Target>& v, Target carry) {   // gadget :198-224
        if (v.size() == 1) return add_u32(v[0], carry);
        return add_many_gate(v, carry);
    }
    std::pair<Target, Target> add_many_gate(const std::vector<Target>& v, Target carry_in) {
        const int na = (int)v.size();
        if (na < 1) throw Error{"add_many_u32: no addends"};
        if (na > 16) throw Error{"add_many_u32: more than 16 addends"};
        const int ops = add_many_ops(na);
        auto& slot = free_add_many[na];
        if (slot.second == 0 || slot.second >= ops) {
            slot.first = add_gate(gate_type(P2G_GATE_U32_ADD_MANY, na, ops));
            slot.second = 0;
        }
        const int row = slot.first, i = slot.second++, q = (na + 3) * i;
        for (int j = 0; j < na; j++) connect(v[j], wire(row, q + j));
        connect(carry_in, wire(row, q + na));
        Gen g = {};
        g.kind = GEN_U32_ADD_MANY;
        g.row = row;
        g.i = i;
        g.n = na;
        g.c0 = ops;
        push_gen(g);
        return {wire(row, q + na + 1), wire(row, q + na + 2)};
    }
    // x - y - borrow = result - 2^32 borrow_out
    std::pair<Target, Target> sub_u32(Target x, Target y, Target borrow) {
        const int ops = u32_sub_ops();
        auto& slot = free_u32_sub;
        if (slot.second == 0 || slot.second >= ops) {
            slot.first = add_gate(gate_type(P2G_GATE_U32_SUBTRACTION, ops));
            slot.second = 0;
        }
        const int row = slot.first, i = slot.second++;
        connect(x, wire(row, 5 * i));
        connect(y, wire(row, 5 * i + 1));
        connect(borrow, wire(row, 5 * i + 2));
        Gen g = {};
        g.kind = GEN_U32_SUB;
        g.row = row;
        g.i = i;
        g.n = ops;
        push_gen(g);
        return {wire(row, 5 * i + 3), wire(row, 5 * i + 4)};
    }
    void range_check_u32(const std::vector<Target>& vals) {   // range_check.rs: one U32RangeCheckGate for the whole vector
        if (vals.empty()) return;
        if (vals.size() > 13) throw Error{"range_check_u32: more than 13 values in one gate"};
        const int n = (int)vals.size(), row = add_gate(gate_type(P2G_GATE_U32_RANGE_CHECK, n));
        for (int i = 0; i < n; i++) connect(vals[i], wire(row, i));
        Gen g = {};
        g.kind = GEN_U32_RANGE;
        g.row = row;
        g.n = n;
        push_gen(g);
    }
    Target cmp_le(Target a, Target b, int num_bits) {   // multiple_comparison.rs list_le_circuit for one pair: ComparisonGate, 2-bit chunks
        const int nc = (num_bits + 1) / 2, row = add_gate(gate_type(P2G_GATE_COMPARISON, num_bits, nc));
        connect(a, wire(row, 0));
        connect(b, wire(row, 1));
        Gen g = {};
        g.kind = GEN_COMPARISON;
        g.row = row;
        g.n = num_bits;
        g.i = nc;
        push_gen(g);
        return wire(row, 2);
    }
    void register_public_input(Target t) { public_inputs.push_back(t); }

    // ---- build(): public-input hash, constants, unused gate slots, padding (plonk/circuit_builder.rs build)
    void build() {
        if (built) return;
        // hash_n_to_hash_no_pad::<PoseidonHash>(public_inputs): sponge of rate 8, overwrite mode, one PoseidonGate row per permutation
        Target state[12];
        for (auto& s : state) s = zero();
        for (size_t off = 0; off < public_inputs.size(); off += 8) {
            const size_t len = std::min<size_t>(8, public_inputs.size() - off);
            for (size_t i = 0; i < len; i++) state[i] = public_inputs[off + i];
            const int row = add_gate(gate_type(P2G_GATE_POSEIDON));
            for (int i = 0; i < 12; i++) connect(state[i], wire(row, i));
            connect(zero(), wire(row, 24));   // swap = 0
            Gen g = {};
            g.kind = GEN_POSEIDON;
            g.row = row;
            push_gen(g);
            for (int i = 0; i < 12; i++) state[i] = wire(row, 12 + i);
        }
        const int pi_row = add_gate(gate_type(P2G_GATE_PUBLIC_INPUT));
        for (int i = 0; i < 4; i++) {
            pi_hash[i] = state[i];
            connect(state[i], wire(pi_row, i));
        }
        // unused arithmetic / random-access slots compute on zeros (their output wires are simply 0 * 0 * c0 + 0 * c1 = 0)
        // constants -> ConstantGate rows
        for (size_t off = 0; off < constant_order.size(); off += CONSTS_PER_GATE) {
            std::vector<u64> cs;
            for (size_t i = off; i < std::min(constant_order.size(), off + CONSTS_PER_GATE); i++) cs.push_back(constant_order[i]);
            while ((int)cs.size() < CONSTS_PER_GATE) cs.push_back(0);
            const int row = add_gate(gate_type(P2G_GATE_CONSTANT, CONSTS_PER_GATE), cs);
            for (size_t i = off; i < std::min(constant_order.size(), off + CONSTS_PER_GATE); i++) {
                connect(constants[constant_order[i]], wire(row, (int)(i - off)));
                Gen g = {};
                g.kind = GEN_CONST;
                g.row = row;
                g.i = (int)(i - off);
                g.c0 = constant_order[i];
                push_gen(g);
            }
        }
        // pad with NoopGate to a power of two (at least 2^2 rows so that the LDE has 2^cap_height leaves)
        size_t n = 4;
        while (n < rows.size()) n <<= 1;
        const int noop = gate_type(P2G_GATE_NOOP);
        while (rows.size() < n) add_gate(noop);
        degree_bits = 0;
        while (((size_t)1 << degree_bits) < n) degree_bits++;
        cur_group = 0;
        finalize();
        built = true;
    }

    // ---- witness generation: plonky2 iop/generator.rs generate_partial_witness restricted to the generators above.
    // Values are kept per copy-constraint class (PartitionWitness); a class set twice with different values is an unsatisfied
    // copy constraint -- plonky2 panics there, this returns an error.  Wires that are no target (the advice columns, and routed
    // wires nothing was ever connected to) live in a dense per-cell table.  Nothing below creates targets or touches the
    // union-find, so generators of different groups can run on different threads: a class is claimed with a compare-exchange on
    // its flag (0 empty, 2 being written, 1 ready).
    std::vector<Target> root;          // flattened union-find (finalize())
    std::vector<Target> routed_root;   // [row * 80 + col] -> class root or -1
    template <typename E>
    struct ZeroBuf {   // calloc'ed: the pages are zero-filled by the kernel when first touched (by whichever thread gets there)
        E* p = nullptr;
        size_t n = 0;
        ZeroBuf() {}
        ZeroBuf(const ZeroBuf&) = delete;
        ZeroBuf& operator=(const ZeroBuf&) = delete;
        ~ZeroBuf() { free(p); }
        void reset(size_t count) {
            free(p);
            n = count;
            p = (E*)calloc(count ? count : 1, sizeof(E));
            if (!p) throw Error{"witness generation: out of memory"};
        }
        bool empty() const { return n == 0; }
        E& operator[](size_t i) { return p[i]; }
        const E& operator[](size_t i) const { return p[i]; }
    };
    ZeroBuf<u64> val, cell_val;    // per class root; per wire [row * 234 + col]
    ZeroBuf<unsigned char> has, cell_has;
    std::vector<std::vector<u32>> group_gens;   // generator indices per group, creation order; group 0 = everything light
    std::vector<int> group_order;               // groups sorted by dependency level: the order the worker threads take them in
    // Field inversions whose result no other generator reads (the u32 arithmetic gate's canonicity inverse, the comparison gate's
    // equality dummies): queued per group while the generators run and done at the end with one batched inversion (Montgomery's
    // trick: three multiplications each) -- they were a third of the run.
    struct DeferredInv {
        int row, col;
        u64 value;   // non-zero
    };
    std::vector<std::vector<DeferredInv>> deferred_inv;   // [group]
    void defer_inverse(const Gen& g, int col, u64 value) { deferred_inv[g.group].push_back({g.row, col, value}); }
    void flush_deferred_inverses() {
        std::vector<DeferredInv> all;
        for (auto& v : deferred_inv) {
            all.insert(all.end(), v.begin(), v.end());
            v.clear();
        }
        const long n = (long)all.size(), CH = 4096;
        std::string err;
#pragma omp parallel for schedule(dynamic, 4)
        for (long lo = 0; lo < n; lo += CH) {
            const long hi = std::min(n, lo + CH);
            std::vector<u64> prefix(hi - lo);
            u64 acc = 1;
            for (long i = lo; i < hi; i++) {
                prefix[i - lo] = acc;
                acc = gl_mul(acc, all[i].value);
            }
            u64 inv = gl_inv(acc);
            try {
                for (long i = hi; i-- > lo;) {
                    setw(all[i].row, all[i].col, gl_mul(inv, prefix[i - lo]));
                    inv = gl_mul(inv, all[i].value);
                }
            } catch (const Error& e) {
#pragma omp critical
                err = e.msg;
            }
        }
        if (!err.empty()) throw Error{err};
    }
    std::vector<u32> const_gens;
    void finalize() {
        routed_target.resize(rows.size() * NUM_ROUTED, -1);   // it grows in blocks of rows
        root.resize(parent.size());
        for (Target t = 0; t < (Target)parent.size(); t++) root[t] = find(t);
        routed_root.assign(routed_target.size(), -1);
        for (size_t k = 0; k < routed_target.size(); k++)
            if (routed_target[k] >= 0) routed_root[k] = root[routed_target[k]];
        group_gens.assign(num_groups, {});
        std::vector<int> level(num_groups, 0);
        for (int g = 0; g < num_groups; g++)
            for (int d : group_deps[g]) level[g] = std::max(level[g], level[d] + 1);   // deps point to earlier groups
        group_order.resize(num_groups);
        for (int g = 0; g < num_groups; g++) group_order[g] = g;
        std::stable_sort(group_order.begin(), group_order.end(), [&](int x, int y) { return level[x] < level[y]; });
        const_gens.clear();
        for (size_t i = 0; i < gens.size(); i++) {
            if (gens[i].kind == GEN_CONST) const_gens.push_back((u32)i);
            else group_gens[gens[i].group].push_back((u32)i);
        }
    }
    void reset_witness() {
        val.reset(parent.size());
        has.reset(parent.size());
        cell_val.reset(rows.size() * NUM_WIRES);
        cell_has.reset(rows.size() * NUM_WIRES);
        deferred_inv.assign(num_groups, {});
        for (auto& g : gens) g.done = false;
    }
    void set_class(Target r, u64 v) {
        unsigned char expected = 0;
        if (__atomic_compare_exchange_n(&has.p[r], &expected, (unsigned char)2, false, __ATOMIC_ACQUIRE, __ATOMIC_ACQUIRE)) {
            val[r] = v;
            __atomic_store_n(&has.p[r], (unsigned char)1, __ATOMIC_RELEASE);
            return;
        }
        while (__atomic_load_n(&has.p[r], __ATOMIC_ACQUIRE) != 1) {
        }
        if (val[r] != v) throw Error{"witness generation: a copy-constrained target was set twice with different values (unsatisfiable witness)"};
    }
    bool get_class(Target r, u64* v) const {
        if (__atomic_load_n(&has.p[r], __ATOMIC_ACQUIRE) != 1) return false;
        *v = val[r];
        return true;
    }
    void set(Target t, u64 v) {
        if (t < 0) {   // advice wire named through wire()
            const size_t k = (size_t)(-2 - t);
            cell_val[k] = v;
            cell_has[k] = 1;
            return;
        }
        set_class(root[t], v);
    }
    bool get(Target t, u64* v) const {
        if (t < 0) {
            const size_t k = (size_t)(-2 - t);
            *v = cell_val[k];
            return cell_has[k] != 0;
        }
        return get_class(root[t], v);
    }
    // gate-local wire access for the generators
    bool skip_advice = false;   // p2a_witness_routed: the advice columns are left to the device (csrc/advice.cuh)
    void setw(int row, int col, u64 v) {
        if (col >= NUM_ROUTED && skip_advice) return;
        const Target r = col < NUM_ROUTED ? routed_root[(size_t)row * NUM_ROUTED + col] : -1;
        if (r >= 0) {
            set_class(r, v);
            return;
        }
        const size_t k = (size_t)row * NUM_WIRES + col;
        cell_val[k] = v;
        cell_has[k] = 1;
    }
    bool getw(int row, int col, u64* v) const {
        const Target r = col < NUM_ROUTED ? routed_root[(size_t)row * NUM_ROUTED + col] : -1;
        if (r >= 0) return get_class(r, v);
        const size_t k = (size_t)row * NUM_WIRES + col;
        *v = cell_val[k];
        return cell_has[k] != 0;
    }
    u64 wire_value(size_t row, int col) const {   // full_witness(): unset wires read as zero
        const Target r = col < NUM_ROUTED ? routed_root[row * NUM_ROUTED + col] : -1;
        if (r >= 0) return has[r] == 1 ? val[r] : 0;
        return cell_val[row * NUM_WIRES + col];
    }
    bool run_big(BigGen& g);
    bool run(Gen& g) {
        switch (g.kind) {
        case GEN_BIG:
            return run_big(big_gens[g.i]);
        case GEN_CONST:
            setw(g.row, g.i, g.c0);
            return true;
        case GEN_ARITH: {
            u64 x, y, z;
            if (!getw(g.row, 4 * g.i, &x) || !getw(g.row, 4 * g.i + 1, &y) || !getw(g.row, 4 * g.i + 2, &z)) return false;
            setw(g.row, 4 * g.i + 3, gl_add(gl_mul(gl_mul(x, y), g.c0), gl_mul(z, g.c1)));
            return true;
        }
        case GEN_SPLIT: {   // WireSplitGenerator: integer -> the sum wire of each BaseSum row, 63 bits at a time
            u64 x;
            if (!get(g.t, &x)) return false;
            for (size_t k = 0; k < g.rows.size(); k++) {
                u64 chunk = (k * g.n >= 64) ? 0 : (x >> (k * g.n)) & (g.n >= 64 ? ~0ULL : (((u64)1 << g.n) - 1));
                setw(g.rows[k], 0, chunk);
            }
            return true;
        }
        case GEN_BASE_SPLIT: {   // BaseSplitGenerator: sum -> limbs
            u64 s;
            if (!getw(g.row, 0, &s)) return false;
            if (g.c0 == 4) {   // split_le_base::<4>
                for (int l = 0; l < g.n; l++) setw(g.row, 1 + l, l < 32 ? (s >> (2 * l)) & 3 : 0);
                return true;
            }
            for (int l = 0; l < g.n; l++) setw(g.row, 1 + l, l < 64 ? (s >> l) & 1 : 0);
            return true;
        }
        case GEN_BASE_SUM: {   // BaseSumGenerator: limbs -> sum
            u64 s = 0, pw = 1;
            for (int l = 0; l < g.n; l++) {
                u64 b;
                if (!getw(g.row, 1 + l, &b)) return false;
                s = gl_add(s, gl_mul(b, pw));
                pw = gl_add(pw, pw);
            }
            setw(g.row, 0, s);
            return true;
        }
        case GEN_RANDOM_ACCESS: {
            const GateType& gt = gate_types[rows[g.row].gate];
            const int bits = gt.params[0], copies = gt.params[1], vec = 1 << bits, base = (2 + vec) * g.i;
            u64 idx;
            if (!getw(g.row, base, &idx)) return false;
            if (idx >= (u64)vec) throw Error{"random_access: index out of range"};
            u64 item;
            if (!getw(g.row, base + 2 + (int)idx, &item)) return false;
            setw(g.row, base + 1, item);
            const int routed_used = (2 + vec) * copies + (int)gt.params[2];
            for (int b = 0; b < bits; b++) setw(g.row, routed_used + g.i * bits + b, (idx >> b) & 1);
            return true;
        }
        case GEN_EQUAL: {
            u64 x, y;
            if (!get(g.t, &x) || !get(g.t2, &y)) return false;
            set(g.t3, x == y ? 1 : 0);
            set(g.t4, x == y ? 0 : gl_inv(gl_sub(x, y)));
            return true;
        }
        case GEN_U32_ARITH: {   // arithmetic_u32.rs:376-426
            u64 x, y, z;
            if (!getw(g.row, 6 * g.i, &x) || !getw(g.row, 6 * g.i + 1, &y) || !getw(g.row, 6 * g.i + 2, &z)) return false;
            u64 out = gl_add(gl_mul(x, y), z), hi = out >> 32, lo = out & 0xFFFFFFFFULL;
            setw(g.row, 6 * g.i + 3, lo);
            setw(g.row, 6 * g.i + 4, hi);
            const u64 diff = 0xFFFFFFFFULL - hi;
            if (diff) defer_inverse(g, 6 * g.i + 5, diff);
            else setw(g.row, 6 * g.i + 5, 0);
            for (int j = 0; j < 32; j++) setw(g.row, 6 * g.n + 32 * g.i + j, (out >> (2 * j)) & 3);
            return true;
        }
        case GEN_U32_ADD_MANY: {   // add_many_u32.rs:329-375
            const int na = g.n, ops = (int)g.c0, q = (na + 3) * g.i;
            u64 sum = 0, v;
            for (int j = 0; j <= na; j++) {
                if (!getw(g.row, q + j, &v)) return false;
                sum = gl_add(sum, v);
            }
            const u64 carry = sum >> 32, res = sum & 0xFFFFFFFFULL;
            setw(g.row, q + na + 1, res);
            setw(g.row, q + na + 2, carry);
            const int lw = (na + 3) * ops + 18 * g.i;
            for (int j = 0; j < 16; j++) setw(g.row, lw + j, (res >> (2 * j)) & 3);
            for (int j = 0; j < 2; j++) setw(g.row, lw + 16 + j, (carry >> (2 * j)) & 3);
            return true;
        }
        case GEN_U32_SUB: {   // subtraction_u32.rs:298-343
            u64 x, y, b;
            if (!getw(g.row, 5 * g.i, &x) || !getw(g.row, 5 * g.i + 1, &y) || !getw(g.row, 5 * g.i + 2, &b)) return false;
            const u64 init = gl_sub(gl_sub(x, y), b);
            const u64 bout = init > (1ULL << 32) ? 1 : 0;
            const u64 res = gl_add(init, bout ? (1ULL << 32) : 0);
            setw(g.row, 5 * g.i + 3, res);
            setw(g.row, 5 * g.i + 4, bout);
            for (int j = 0; j < 16; j++) setw(g.row, 5 * g.n + 16 * g.i + j, (res >> (2 * j)) & 3);
            return true;
        }
        case GEN_U32_RANGE: {   // range_check_u32.rs:198-220
            for (int i = 0; i < g.n; i++) {
                u64 v;
                if (!getw(g.row, i, &v)) return false;
                const u32 v32 = (u32)v;
                for (int j = 0; j < 16; j++) setw(g.row, g.n + 16 * i + j, (v32 >> (2 * j)) & 3);
            }
            return true;
        }
        case GEN_COMPARISON: {   // comparison.rs:439-537
            const int nc = g.i, cb = (g.n + nc - 1) / nc;
            u64 a, b;
            if (!getw(g.row, 0, &a) || !getw(g.row, 1, &b)) return false;
            setw(g.row, 2, a <= b ? 1 : 0);
            const u64 cs = 1ULL << cb;
            u64 msd = 0;
            for (int i = 0; i < nc; i++) {
                const u64 fa = (a >> (cb * i)) & (cs - 1), fb = (b >> (cb * i)) & (cs - 1);
                setw(g.row, 4 + i, fa);
                setw(g.row, 4 + nc + i, fb);
                if (fa == fb) setw(g.row, 4 + 2 * nc + i, 1);   // equality dummy
                else defer_inverse(g, 4 + 2 * nc + i, gl_sub(fb, fa));
                setw(g.row, 4 + 3 * nc + i, fa == fb ? 1 : 0);                        // chunks equal
                if (fa != fb) {
                    msd = gl_sub(fb, fa);
                    setw(g.row, 4 + 4 * nc + i, 0);
                } else {
                    setw(g.row, 4 + 4 * nc + i, msd);
                }
            }
            setw(g.row, 3, msd);
            const u64 t = gl_add(cs, msd);
            for (int i = 0; i <= cb; i++) setw(g.row, 4 + 5 * nc + i, (t >> i) & 1);
            return true;
        }
        case GEN_POSEIDON: {
            u64 st[12];
            for (int i = 0; i < 12; i++)
                if (!getw(g.row, i, &st[i])) return false;
            u64 swap;
            if (!getw(g.row, 24, &swap)) return false;
            for (int i = 0; i < 4; i++) {
                u64 delta = gl_mul(swap, gl_sub(st[i + 4], st[i]));
                setw(g.row, 25 + i, delta);
                st[i] = gl_add(st[i], delta);
                st[i + 4] = gl_sub(st[i + 4], delta);
            }
            int rnd = 0;
            for (int r = 0; r < 4; r++, rnd++) {
                for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
                if (r != 0)
                    for (int i = 0; i < 12; i++) setw(g.row, 29 + 12 * (r - 1) + i, st[i]);
                for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
                poseidon_mds(st);
            }
            for (int r = 0; r < 22; r++, rnd++) {
                for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
                setw(g.row, 65 + r, st[0]);
                st[0] = poseidon_sbox(st[0]);
                poseidon_mds(st);
            }
            for (int r = 0; r < 4; r++, rnd++) {
                for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
                for (int i = 0; i < 12; i++) setw(g.row, 87 + 12 * r + i, st[i]);
                for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
                poseidon_mds(st);
            }
            for (int i = 0; i < 12; i++) setw(g.row, 12 + i, gl_canon(st[i]));
            return true;
        }
        }
        return false;
    }
};

// ---- the translator (circuit_translation/mod.rs) -----------------------------------------------------------------------------
struct Translator {
    Builder b;
    std::map<u32, Target> witness_target_map;
    std::map<u32, std::pair<std::vector<Target>, size_t>> memory_blocks;
    Target target_for_witness(u32 w) {
        auto it = witness_target_map.find(w);
        if (it != witness_target_map.end()) return it->second;
        Target t = b.add_virtual_target();
        witness_target_map[w] = t;
        return t;
    }
    std::vector<Target> binary_number_target_for_witness(u32 w, int digits) {   // most significant bit first (mod.rs:262-274)
        std::vector<Target> bits = b.split_le(target_for_witness(w), digits);
        std::reverse(bits.begin(), bits.end());
        return bits;
    }
    Target convert_binary_number_to_number(std::vector<Target> bits) {
        std::reverse(bits.begin(), bits.end());
        return b.le_sum(bits);
    }
};

#include "ecdsa.h"

// ---- BinaryDigitsTarget (plonky2-backend/src/binary_digits_target.rs): bit vectors, most significant bit first ---------------
typedef std::vector<Target> Bits;
Bits rotate_right(Builder& b, const Bits& t, size_t times) {   // :21-41
    Bits out;
    const size_t n = t.size();
    for (size_t i = n - times; i < n; i++) {
        Target nb = b.add_virtual_bool_target_safe();
        b.connect(t[i], nb);
        out.push_back(nb);
    }
    for (size_t i = 0; i < n - times; i++) {
        Target nb = b.add_virtual_bool_target_safe();
        b.connect(t[i], nb);
        out.push_back(nb);
    }
    return out;
}
Bits shift_right(Builder& b, const Bits& t, size_t times) {   // :43-63
    Bits out;
    for (size_t i = 0; i < times; i++) out.push_back(b.constant(0));
    for (size_t i = 0; i < t.size() - times; i++) {
        Target nb = b.add_virtual_bool_target_safe();
        b.connect(t[i], nb);
        out.push_back(nb);
    }
    return out;
}
Bits bits_xor(Builder& b, const Bits& x, const Bits& y) {
    Bits o(x.size());
    for (size_t i = 0; i < x.size(); i++) o[i] = b.b_xor(x[i], y[i]);
    return o;
}
Bits choose(Builder& b, const Bits& c, const Bits& t, const Bits& f) {   // :65-82
    Bits o(c.size());
    for (size_t i = 0; i < c.size(); i++) o[i] = b.select(c[i], t[i], f[i]);
    return o;
}
Bits majority(Builder& b, const Bits& x, const Bits& y, const Bits& z) {   // :84-106: select(z, x or y, x and y)
    Bits o(x.size());
    for (size_t i = 0; i < x.size(); i++) o[i] = b.select(z[i], b.b_or(x[i], y[i]), b.b_and(x[i], y[i]));
    return o;
}
Bits add_module_32_bits(Builder& b, const Bits& x, const Bits& y) {   // :188-221: ripple-carry adder, carry out dropped
    const size_t n = x.size();
    Bits ps(n), pc(n), sum;
    for (size_t i = 0; i < n; i++) ps[i] = b.b_xor(x[i], y[i]);
    for (size_t i = 0; i < n; i++) pc[i] = b.b_and(x[i], y[i]);
    Target carry = b.constant(0);
    for (size_t k = n; k-- > 0;) {
        Target s = b.b_xor(ps[k], carry);
        Target pair = b.b_and(carry, ps[k]);
        carry = b.b_or(pc[k], pair);
        sum.push_back(s);
    }
    std::reverse(sum.begin(), sum.end());
    return sum;
}
Bits bits_for_constant(Builder& b, u32 c, int digits) {   // mod.rs:243-253
    Bits o;
    for (int pos = digits - 1; pos >= 0; pos--) o.push_back(b.constant((c >> pos) & 1));
    return o;
}

// Sha256Compression (circuit_translation/sha256_translator.rs:60-273): 16 message words, 8 state words -> 8 state words
void sha256_compression(Translator& T, const u32* in_w, const u32* hv_w, const u32* out_w) {
    static const u32 K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    Builder& b = T.b;
    auto x3 = [&](const Bits& t, int r1, int r2, int r3, bool shift3) {   // sigma / big sigma: rot ^ rot ^ (rot | shift)
        Bits a = rotate_right(b, t, r1), c = rotate_right(b, t, r2), d = shift3 ? shift_right(b, t, r3) : rotate_right(b, t, r3);
        return bits_xor(b, bits_xor(b, a, c), d);
    };
    std::vector<Bits> w;
    for (int i = 0; i < 16; i++) w.push_back(T.binary_number_target_for_witness(in_w[i], 32));
    for (int t = 16; t < 64; t++) {   // calculate_w_t
        Bits s1 = x3(w[t - 2], 17, 19, 10, true);
        Bits a1 = add_module_32_bits(b, s1, w[t - 7]);
        Bits s0 = x3(w[t - 15], 7, 18, 3, true);
        Bits a2 = add_module_32_bits(b, s0, w[t - 16]);
        w.push_back(add_module_32_bits(b, a1, a2));
    }
    std::vector<Bits> k;
    for (int t = 0; t < 64; t++) k.push_back(bits_for_constant(b, K[t], 32));
    std::vector<Bits> h0;
    for (int i = 0; i < 8; i++) h0.push_back(T.binary_number_target_for_witness(hv_w[i], 32));
    std::vector<Bits> st = h0;   // a b c d e f g h
    for (int t = 0; t < 64; t++) {   // compression_function_iteration
        Bits S1 = x3(st[4], 6, 11, 25, false);
        Bits ch = choose(b, st[4], st[5], st[6]);
        Bits s0 = add_module_32_bits(b, k[t], w[t]);
        Bits s1 = add_module_32_bits(b, st[7], S1);
        Bits s2 = add_module_32_bits(b, ch, s0);
        Bits t1 = add_module_32_bits(b, s1, s2);
        Bits S0 = x3(st[0], 2, 13, 22, false);
        Bits mj = majority(b, st[0], st[1], st[2]);
        Bits t2 = add_module_32_bits(b, S0, mj);
        std::vector<Bits> nx(8);
        nx[0] = add_module_32_bits(b, t1, t2);
        nx[1] = st[0];
        nx[2] = st[1];
        nx[3] = st[2];
        nx[4] = add_module_32_bits(b, st[3], t1);
        nx[5] = st[4];
        nx[6] = st[5];
        nx[7] = st[6];
        st = nx;
    }
    for (int i = 0; i < 8; i++) {
        Bits fin = add_module_32_bits(b, h0[i], st[i]);
        T.witness_target_map[out_w[i]] = T.convert_binary_number_to_number(fin);
    }
}

// memory_translator.rs:55-85: index <= max_allowed_value, bit by bit from the most significant one
void assert_less_or_equal(Builder& b, size_t max_allowed, Target index) {
    int nbits = 1;
    while ((max_allowed >> nbits) != 0) nbits++;
    Bits bits = b.split_le(index, nbits);
    std::reverse(bits.begin(), bits.end());
    Target acc = b.one();
    for (int i = 0; i < nbits; i++) {
        const int bit = (int)((max_allowed >> (nbits - 1 - i)) & 1);
        if (bit == 0) b.assert_zero(b.mul(bits[i], acc));
        else acc = b.mul(acc, bits[i]);
    }
}

// flat opcode stream (what acir.py writes): u64 words
//   1 AssertZero: n_mul, n_lin, q_c, then n_mul x (coef, w1, w2), n_lin x (coef, w)          assert_zero_translator.rs:30-116
//   2 RANGE: witness, num_bits                                                                  mod.rs:131-137
//   3 AND / 4 XOR: lhs, rhs, num_bits, output                                                   mod.rs:139-155, 213-232
//   5 MemoryInit: block_id, n, then n witnesses                                                 memory_translator.rs:145-156
//   6 MemoryRead: block_id, index witness, value witness                                        memory_translator.rs:125-137
//   7 Sha256Compression: 16 input witnesses, 8 hash-value witnesses, 8 output witnesses         sha256_translator.rs:60-111
//   8 MemoryWrite: block_id, index witness, value witness                                       memory_translator.rs:87-113
//   9 EcdsaSecp256k1: 32 public_key_x, 32 public_key_y, 64 signature, 32 hashed_message byte witnesses, output   ecdsa_secp256k1_translator.rs:38-60
// Gadget-level operations (NOT ACIR opcodes: the reference reaches them only through its EcdsaSecp256k1 translator; here they let
// a circuit be built directly on the reference's u32 gadgets, like its gadget tests do).  Witness ids name the targets.
// 101 MulAddU32: x, y, z, low, high      102 AddManyU32: n, n addends, result, carry      103 SubU32: x, y, borrow, result, borrow_out
// 104 RangeCheckU32: n, n values         105 CmpLe: a, b, num_bits, result
// 110 Gadget: kind, param, number of lists, then (length, witness ids...) per list -- the biguint / non-native / comparison gadgets of
//     plonky2_ecdsa on u32-limb witnesses (least significant limb first), as its gadget tests drive them.  kind: 1 add_biguint [a, b, out]
//     2 sub_biguint [a, b, out]  3 mul_biguint [a, b, out]  4 cmp_biguint [a, b, [result]]  5 div_rem_biguint [a, b, div, rem]
//     6 add_nonnative  7 sub_nonnative  8 mul_nonnative [a, b, out]  9 neg_nonnative  10 inv_nonnative [a, out]
//     11 add_many_nonnative [summand..., out]  12 list_le [a, b, [result]] (param = bits per element)
//     13 glv_mul [px, py, k, outx, outy]  14 curve_add [p1x, p1y, p2x, p2y, outx, outy]  15 curve_double [px, py, outx, outy]
//     param for 6..11: 0 = secp256k1 base field, 1 = scalar field
enum { OP_ASSERT_ZERO = 1, OP_RANGE = 2, OP_AND = 3, OP_XOR = 4, OP_MEM_INIT = 5, OP_MEM_READ = 6, OP_SHA256_COMPRESSION = 7, OP_MEM_WRITE = 8, OP_ECDSA_SECP256K1 = 9,
       OP_MUL_ADD_U32 = 101, OP_ADD_MANY_U32 = 102, OP_SUB_U32 = 103, OP_RANGE_CHECK_U32 = 104, OP_CMP_LE = 105, OP_GADGET = 110 };

// gadget-level operations on limb witnesses (opcode 110): what the reference's gadget tests exercise
void gadget(Translator& T, int kind, int param, const std::vector<std::vector<u32>>& lists) {
    Ecc e(T.b);
    auto in = [&](size_t k) {
        if (k >= lists.size()) throw Error{"Gadget: missing operand list"};
        BigT t;
        for (u32 w : lists[k]) t.push_back(T.target_for_witness(w));
        return t;
    };
    auto out = [&](size_t k, const BigT& v) {   // bind the result limbs to witness ids; missing high limbs must be zero
        if (k >= lists.size()) throw Error{"Gadget: missing result list"};
        if (lists[k].size() > v.size()) throw Error{"Gadget: result list longer than the result (" + std::to_string(v.size()) + " limbs)"};
        for (size_t i = 0; i < v.size(); i++) {
            if (i < lists[k].size()) T.witness_target_map[lists[k][i]] = v[i];
            else T.b.assert_zero(v[i]);
        }
    };
    const int f = param ? FIELD_SCALAR : FIELD_BASE;
    switch (kind) {
    case 1: out(2, e.add_biguint(in(0), in(1))); break;
    case 2: out(2, e.sub_biguint(in(0), in(1))); break;
    case 3: out(2, e.mul_biguint(in(0), in(1))); break;
    case 4: out(2, BigT{e.cmp_biguint(in(0), in(1))}); break;
    case 5: {
        auto qr = e.div_rem_biguint(in(0), in(1));
        out(2, qr.first);
        out(3, qr.second);
        break;
    }
    case 6: out(2, e.add_nonnative(f, in(0), in(1))); break;
    case 7: out(2, e.sub_nonnative(f, in(0), in(1))); break;
    case 8: out(2, e.mul_nonnative(f, in(0), in(1))); break;
    case 9: out(1, e.neg_nonnative(f, in(0))); break;
    case 10: out(1, e.inv_nonnative(f, in(0))); break;
    case 11: {
        std::vector<BigT> summands;
        for (size_t k = 0; k + 1 < lists.size(); k++) summands.push_back(in(k));
        if (summands.empty()) throw Error{"Gadget: add_many_nonnative without summands"};
        out(lists.size() - 1, e.add_many_nonnative(f, summands));
        break;
    }
    case 12: out(2, BigT{e.list_le(in(0), in(1), param)}); break;
    case 13: {
        AffinePoint r = e.glv_mul({in(0), in(1)}, in(2));
        out(3, r.x);
        out(4, r.y);
        break;
    }
    case 14: {
        AffinePoint r = e.curve_add({in(0), in(1)}, {in(2), in(3)});
        out(4, r.x);
        out(5, r.y);
        break;
    }
    case 15: {
        AffinePoint r = e.curve_double({in(0), in(1)});
        out(2, r.x);
        out(3, r.y);
        break;
    }
    default: throw Error{"Gadget: unknown kind " + std::to_string(kind)};
    }
}

void translate(Translator& T, const u64* pub, size_t npub, const u64* priv, size_t npriv, const u64* ops, size_t nwords) {
    Builder& b = T.b;
    for (size_t i = 0; i < npub; i++) {   // _register_witnesses_from_acir_circuit (mod.rs:289-316)
        Target t = b.add_virtual_target();
        b.register_public_input(t);
        T.witness_target_map[(u32)pub[i]] = t;
    }
    for (size_t i = 0; i < npriv; i++) T.target_for_witness((u32)priv[i]);
    size_t p = 0;
    auto next = [&]() {
        if (p >= nwords) throw Error{"opcode stream truncated"};
        return ops[p++];
    };
    while (p < nwords) {
        const u64 op = next();
        switch (op) {
        case OP_ASSERT_ZERO: {
            const u64 n_mul = next(), n_lin = next(), q_c = next();
            if (q_c >= GL_P) throw Error{"AssertZero: non-canonical constant"};
            std::vector<std::tuple<u64, u32, u32>> muls;
            std::vector<std::pair<u64, u32>> lins;
            for (u64 i = 0; i < n_mul; i++) {
                u64 c = next(), w1 = next(), w2 = next();
                muls.emplace_back(c, (u32)w1, (u32)w2);
            }
            for (u64 i = 0; i < n_lin; i++) {
                u64 c = next(), w = next();
                lins.emplace_back(c, (u32)w);
            }
            for (auto& m : muls) {   // _register_intermediate_witnesses_for_assert_zero
                T.target_for_witness(std::get<1>(m));
                T.target_for_witness(std::get<2>(m));
            }
            for (auto& l : lins) T.target_for_witness(l.second);
            Target acc = b.constant(q_c);
            for (auto& l : lins) acc = b.add(b.mul_const(l.first, T.target_for_witness(l.second)), acc);
            for (auto& m : muls) {
                Target q = b.mul(T.target_for_witness(std::get<1>(m)), T.target_for_witness(std::get<2>(m)));
                acc = b.add(b.mul_const(std::get<0>(m), q), acc);
            }
            b.assert_zero(acc);
            break;
        }
        case OP_RANGE: {
            const u64 w = next(), bits = next();
            if (bits > 33) throw Error{"Range checks with more than 33 bits are not allowed yet while using Plonky2 prover"};
            b.range_check(T.target_for_witness((u32)w), (int)bits);
            break;
        }
        case OP_AND:
        case OP_XOR: {
            const u64 lhs = next(), rhs = next(), bits = next(), out = next();
            std::vector<Target> l = T.binary_number_target_for_witness((u32)lhs, (int)bits);
            std::vector<Target> r = T.binary_number_target_for_witness((u32)rhs, (int)bits);
            std::vector<Target> o(l.size());
            for (size_t i = 0; i < l.size(); i++) o[i] = op == OP_AND ? b.b_and(l[i], r[i]) : b.b_xor(l[i], r[i]);
            T.witness_target_map[(u32)out] = T.convert_binary_number_to_number(o);
            break;
        }
        case OP_MEM_INIT: {
            const u64 id = next(), n = next();
            std::vector<Target> v;
            for (u64 i = 0; i < n; i++) v.push_back(T.target_for_witness((u32)next()));
            size_t real = v.size(), len = 1;
            while (len < real) len <<= 1;
            while (v.size() < len) v.push_back(b.zero());
            T.memory_blocks[(u32)id] = {v, real};
            break;
        }
        case OP_MEM_READ: {
            const u64 id = next(), iw = next(), vw = next();
            auto it = T.memory_blocks.find((u32)id);
            if (it == T.memory_blocks.end()) throw Error{"MemoryOp on an uninitialised block"};
            assert_less_or_equal(b, it->second.second - 1, T.target_for_witness((u32)iw));
            Target res = b.random_access(T.target_for_witness((u32)iw), it->second.first);
            T.witness_target_map[(u32)vw] = res;
            break;
        }
        case OP_MEM_WRITE: {   // memory_translator.rs:87-113: every cell becomes  index == position ? new value : old value
            const u64 id = next(), iw = next(), vw = next();
            auto it = T.memory_blocks.find((u32)id);
            if (it == T.memory_blocks.end()) throw Error{"MemoryOp on an uninitialised block"};
            Target idx = T.target_for_witness((u32)iw), val = T.target_for_witness((u32)vw);
            assert_less_or_equal(b, it->second.second - 1, idx);
            std::vector<Target>& cells = it->second.first;
            for (size_t pos = 0; pos < cells.size(); pos++) {
                Target eq = b.is_equal(idx, b.constant(pos));
                cells[pos] = b.select(eq, val, cells[pos]);
            }
            break;
        }
        case OP_MUL_ADD_U32: {
            const u64 x = next(), y = next(), z = next(), lo = next(), hi = next();
            auto r = b.mul_add_u32(T.target_for_witness((u32)x), T.target_for_witness((u32)y), T.target_for_witness((u32)z));
            T.witness_target_map[(u32)lo] = r.first;
            T.witness_target_map[(u32)hi] = r.second;
            break;
        }
        case OP_ADD_MANY_U32: {
            const u64 n = next();
            std::vector<Target> v;
            for (u64 i = 0; i < n; i++) v.push_back(T.target_for_witness((u32)next()));
            const u64 res = next(), carry = next();
            auto r = b.add_many_u32(v);
            T.witness_target_map[(u32)res] = r.first;
            T.witness_target_map[(u32)carry] = r.second;
            break;
        }
        case OP_SUB_U32: {
            const u64 x = next(), y = next(), bw = next(), res = next(), bo = next();
            auto r = b.sub_u32(T.target_for_witness((u32)x), T.target_for_witness((u32)y), T.target_for_witness((u32)bw));
            T.witness_target_map[(u32)res] = r.first;
            T.witness_target_map[(u32)bo] = r.second;
            break;
        }
        case OP_RANGE_CHECK_U32: {
            const u64 n = next();
            std::vector<Target> v;
            for (u64 i = 0; i < n; i++) v.push_back(T.target_for_witness((u32)next()));
            b.range_check_u32(v);
            break;
        }
        case OP_CMP_LE: {
            const u64 x = next(), y = next(), bits = next(), res = next();
            if (bits == 0 || bits > 32) throw Error{"CmpLe: 1..32 bits"};
            T.witness_target_map[(u32)res] = b.cmp_le(T.target_for_witness((u32)x), T.target_for_witness((u32)y), (int)bits);
            break;
        }
        case OP_GADGET: {
            const u64 kind = next(), param = next(), nlists = next();
            std::vector<std::vector<u32>> lists(nlists);
            for (auto& l : lists) {
                const u64 len = next();
                for (u64 i = 0; i < len; i++) l.push_back((u32)next());
            }
            gadget(T, (int)kind, (int)param, lists);
            break;
        }
        case OP_ECDSA_SECP256K1: {
            u32 ws[161];
            for (int i = 0; i < 161; i++) ws[i] = (u32)next();
            ecdsa_secp256k1(T, ws, ws + 32, ws + 64, ws + 128, ws[160]);
            break;
        }
        case OP_SHA256_COMPRESSION: {
            u32 ws[32];
            for (int i = 0; i < 32; i++) ws[i] = (u32)next();
            b.cur_group = b.new_group();   // heavy opcode: its generators form a group of their own (p2a_witness)
            sha256_compression(T, ws, ws + 16, ws + 24);
            b.cur_group = 0;
            break;
        }
        default: throw Error{"Opcode not supported yet: " + std::to_string(op)};
        }
    }
    b.build();
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* p2a_last_error(void) { return g_err.c_str(); }

// translate an ACIR circuit; returns a handle or NULL
void* p2a_translate(const u64* pub, size_t npub, const u64* priv, size_t npriv, const u64* ops, size_t nwords) {
    Translator* T = new Translator();
    try {
        translate(*T, pub, npub, priv, npriv, ops, nwords);
        return T;
    } catch (const Error& e) {
        g_err = e.msg;
        delete T;
        return nullptr;
    }
}
void p2a_destroy(void* h) { delete (Translator*)h; }

// degree_bits, number of distinct gate types, number of public inputs, rows used before padding
void p2a_shape(void* h, u32* degree_bits, u32* ngates, u32* npub) {
    Translator* T = (Translator*)h;
    *degree_bits = T->b.degree_bits;
    *ngates = (u32)T->b.gate_types.size();
    *npub = (u32)T->b.public_inputs.size();
}
// distinct gate types in creation order: kind + params[4] each
void p2a_gate_types(void* h, u32* out) {
    Translator* T = (Translator*)h;
    for (size_t i = 0; i < T->b.gate_types.size(); i++) {
        out[5 * i] = T->b.gate_types[i].kind;
        for (int k = 0; k < 4; k++) out[5 * i + 1 + k] = T->b.gate_types[i].params[k];
    }
}

// The preprocessed polynomials for the SORTED gate table `gates` (common.gates order with selector data, from the host-side
// CommonCircuitData): constants_sigmas [num_constants + 80][N] values; `type_to_gate[i]` = position of creation-order type i in it.
int p2a_constants_sigmas(void* h, const p2g_gate* gates, u32 ngates, const u32* type_to_gate, u32 num_selectors, u32 num_constants,
                         const u64* k_is, u64* out) {
    Translator* T = (Translator*)h;
    Builder& b = T->b;
    try {
        const size_t n = (size_t)1 << b.degree_bits;
        const u64 UNUSED = 0xFFFFFFFFULL;
        // selectors + gate constants
        for (size_t r = 0; r < n; r++) {
            const u32 g = type_to_gate[b.rows[r].gate];
            if (g >= ngates) throw Error{"bad gate mapping"};
            for (u32 s = 0; s < num_selectors; s++) out[(size_t)s * n + r] = (s == gates[g].selector_index) ? g : UNUSED;
            for (u32 c = num_selectors; c < num_constants; c++) {
                const size_t k = c - num_selectors;
                out[(size_t)c * n + r] = k < b.rows[r].consts.size() ? b.rows[r].consts[k] : 0;
            }
        }
        // sigmas: plonky2 plonk/permutation_argument.rs WirePartition::get_sigma_polys -- the routed wires of a copy class, in
        // row-major order, form one cycle; sigma(row, col) = k_is[col'] * omega^row' of the next wire of the cycle
        const u64 w = gl_root_of_unity(b.degree_bits);
        std::vector<u64> subgroup(n);
        u64 x = 1;
        for (size_t r = 0; r < n; r++) {
            subgroup[r] = x;
            x = gl_mul(x, w);
        }
        u64* sig = out + (size_t)num_constants * n;
#pragma omp parallel for schedule(static)
        for (int c = 0; c < NUM_ROUTED; c++)
            for (size_t r = 0; r < n; r++) sig[(size_t)c * n + r] = gl_mul(k_is[c], subgroup[r]);
        // walk the routed wires in row-major order, chaining each to the previous wire of its class; close every cycle at the end
        const u32 NONE = 0xFFFFFFFFu;
        std::vector<u32> first(b.parent.size(), NONE), last(b.parent.size(), NONE);
        auto link = [&](u32 from, u32 to) { sig[(size_t)(from % NUM_ROUTED) * n + from / NUM_ROUTED] = gl_mul(k_is[to % NUM_ROUTED], subgroup[to / NUM_ROUTED]); };
        for (size_t pos = 0; pos < b.routed_root.size(); pos++) {
            const Target t = b.routed_root[pos];
            if (t < 0) continue;   // never connected: fixed point
            if (first[t] == NONE) first[t] = (u32)pos;
            else link(last[t], (u32)pos);
            last[t] = (u32)pos;
        }
        for (size_t t = 0; t < first.size(); t++)
            if (first[t] != NONE) link(last[t], first[t]);
        return 0;
    } catch (const Error& e) {
        g_err = e.msg;
        return -1;
    }
}

// Witness generation: ACIR witness map (ids + canonical values) -> wires [234][N] (unset wires are zero, as
// in plonky2's full_witness) and the public inputs in registration order.  Returns 0, or -1 (message in p2a_last_error) when a
// copy constraint is contradicted -- where the reference panics inside witness generation, before the prover.
static int witness_impl(void* h, const u64* ids, const u64* values, size_t nw, u64* wires, u64* public_inputs, bool routed_only) {
    Translator* T = (Translator*)h;
    Builder& b = T->b;
    try {
        b.skip_advice = routed_only;
        const int out_cols = routed_only ? NUM_ROUTED : NUM_WIRES;
        const size_t n = (size_t)1 << b.degree_bits;
        const bool trace = getenv("P2A_TRACE") != nullptr;
        auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; };
        const double t0 = now();
        b.reset_witness();
        for (size_t i = 0; i < nw; i++) {
            if (values[i] >= GL_P) throw Error{"witness value is not canonical"};
            auto it = T->witness_target_map.find((u32)ids[i]);
            if (it == T->witness_target_map.end()) continue;   // a witness the circuit never mentions (Brillig intermediates)
            b.set(it->second, values[i]);
        }
        for (u32 i : b.const_gens) b.gens[i].done = b.run(b.gens[i]);
        // Rounds over the groups: one thread walks a group's generators in creation order (the order plonky2's queue would
        // reach them in), different groups run side by side; a generator whose inputs another group has not produced yet stays
        // pending for the next round.  Whatever is still pending after the rounds is swept sequentially to a fixed point.
        std::string first_error;
        size_t pending = 0;
        int rounds = 0;
        for (; rounds < 12; rounds++) {
            long progress = 0;
            pending = 0;
            // work queue in dependency-level order; a thread that takes a group first waits for the groups it depends on (they
            // were handed out earlier, so some other thread is running them)
            std::vector<unsigned char> group_done(b.num_groups, 0);
            long next_slot = 0;
#pragma omp parallel reduction(+ : progress, pending)
            for (;;) {
                const long slot = __atomic_fetch_add(&next_slot, 1, __ATOMIC_RELAXED);
                if (slot >= (long)b.group_order.size()) break;
                const int gi = b.group_order[slot];
                for (int d : b.group_deps[gi])
                    while (!__atomic_load_n(&group_done[d], __ATOMIC_ACQUIRE)) {
                    }
                try {
                    for (u32 i : b.group_gens[gi]) {
                        Gen& g = b.gens[i];
                        if (g.done) continue;
                        if (b.run(g)) {
                            g.done = true;
                            progress++;
                        } else {
                            pending++;
                        }
                    }
                } catch (const Error& e) {
#pragma omp critical
                    if (first_error.empty()) first_error = e.msg;
                }
                __atomic_store_n(&group_done[gi], (unsigned char)1, __ATOMIC_RELEASE);
            }
            if (!first_error.empty()) throw Error{first_error};
            if (!pending || !progress) break;
        }
        for (int sweep = 0; pending && sweep < 64; sweep++) {
            bool progress = false;
            pending = 0;
            for (auto& g : b.gens) {
                if (g.done) continue;
                if (b.run(g)) {
                    g.done = true;
                    progress = true;
                } else {
                    pending++;
                }
            }
            if (!progress) break;
        }
        b.flush_deferred_inverses();
        const double t1 = now();
        // the wire matrix is column-major [wire][row]: transpose the row-major tables in blocks of rows (every row < n has a gate)
        if (b.rows.size() != n) throw Error{"witness generation: the circuit is not padded"};
        const long BLK = 64, nblk = (long)((n + BLK - 1) / BLK);
#pragma omp parallel for schedule(static)
        for (long k = 0; k < nblk; k++) {
            const size_t r0 = (size_t)k * BLK, r1 = std::min(n, r0 + BLK);
            for (int c = 0; c < out_cols; c++)
                for (size_t r = r0; r < r1; r++) wires[(size_t)c * n + r] = b.wire_value(r, c);
        }
        for (size_t i = 0; i < b.public_inputs.size(); i++) {
            u64 v;
            if (!b.get(b.public_inputs[i], &v)) throw Error{"public input has no value"};
            public_inputs[i] = v;
        }
        if (trace)
            fprintf(stderr, "p2a_witness: %zu generators in %d groups, %d rounds, %.3f s (%zu left unset); wire matrix %.3f s\n", b.gens.size(),
                    b.num_groups, rounds + 1, t1 - t0, pending, now() - t1);
        return 0;
    } catch (const Error& e) {
        g_err = e.msg;
        return -1;
    }
}

int p2a_witness(void* h, const u64* ids, const u64* values, size_t nw, u64* wires, u64* public_inputs) {
    return witness_impl(h, ids, values, nw, wires, public_inputs, false);
}
// The same, producing only the routed columns [80][2^degree_bits]: the generators do not write the advice wires (the limbs of the u32
// gates, ...), which p2g_prove_routed_columns / p2g_fill_advice_device compute on the device
int p2a_witness_routed(void* h, const u64* ids, const u64* values, size_t nw, u64* routed_wires, u64* public_inputs) {
    return witness_impl(h, ids, values, nw, routed_wires, public_inputs, true);
}

// Values the last p2a_witness run assigned to ACIR witnesses (outputs computed by the generators included): what the reference reads
// back through its witness_target_map.  known[i] = 0 for a witness the circuit never mentions or that stayed unset.
void p2a_read_witnesses(void* h, const u64* ids, size_t n, u64* values, uint8_t* known) {
    Translator* T = (Translator*)h;
    for (size_t i = 0; i < n; i++) {
        auto it = T->witness_target_map.find((u32)ids[i]);
        u64 v = 0;
        known[i] = it != T->witness_target_map.end() && !T->b.has.empty() && T->b.get(it->second, &v);
        values[i] = v;
    }
}
// worker threads for witness generation and the preprocessed-polynomial fill (0 = OpenMP's default)
void p2a_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
#else
    (void)n;
#endif
}
// csrc/advice.cuh on the host: recompute, in place, the advice columns (>= 80) of a wire matrix [234][2^degree_bits] of this circuit
// from its routed columns -- the CPU twin of p2g_fill_advice_device, checked in tests/ against the generators above
void p2a_fill_advice(void* h, u64* wires) {
    Translator* T = (Translator*)h;
    const Builder& b = T->b;
    const size_t n = (size_t)1 << b.degree_bits;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) {
        const GateType& gt = b.gate_types[b.rows[r].gate];
        auto get = [&](u32 col) -> u64 { return col < (u32)NUM_WIRES ? wires[(size_t)col * n + r] : 0; };
        auto put = [&](u32 col, u64 v) {
            if (col >= (u32)NUM_ROUTED && col < (u32)NUM_WIRES) wires[(size_t)col * n + r] = v;
        };
        fill_advice_row(gt.kind, gt.params, get, put);
    }
}
// the same for any trace: gate table (kind + params[4] per gate), the gate index of every row, wires [num_wires][n]
void p2a_fill_advice_rows(const u32* kind_and_params, u32 num_gates, const uint8_t* row_gate, u64* wires, size_t n, u32 num_wires,
                          u32 num_routed) {
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) {
        const u32 g = row_gate[r];
        if (g >= num_gates) continue;
        auto get = [&](u32 col) -> u64 { return col < num_wires ? wires[(size_t)col * n + r] : 0; };
        auto put = [&](u32 col, u64 v) {
            if (col >= num_routed && col < num_wires) wires[(size_t)col * n + r] = v;
        };
        fill_advice_row(kind_and_params[5 * g], kind_and_params + 5 * g + 1, get, put);
    }
}
// self-test hook for acir/bigint.h (include/p2acir.h)
int p2a_bigint_selftest(int op, const u32* a, size_t na, const u32* b, size_t nb, const u32* m, size_t nm, u32* q, size_t* nq, u32* r,
                        size_t* nr, u32* flags) {
    try {
        auto load = [](const u32* p, size_t n) {
            Big v;
            v.d.assign(p, p + n);
            v.trim();
            return v;
        };
        auto store = [](const Big& v, u32* p, size_t* n) {
            if (v.d.size() > 40) throw Error{"bigint selftest: result too long"};
            std::copy(v.d.begin(), v.d.end(), p);
            *n = v.d.size();
        };
        const Big A = load(a, na), B = load(b, nb), M = load(m, nm);
        Big Q, R;
        *flags = 0;
        switch (op) {
        case 0:
            if (B.is_zero()) throw Error{"bigint selftest: division by zero"};
            big_divrem(A, B, &Q, &R);
            break;
        case 1: Q = big_mul(A, B); break;
        case 2:
            if (M.is_zero()) throw Error{"bigint selftest: zero modulus"};
            Q = big_powmod(A, B, M);
            break;
        case 4:
            if (M.is_zero() || !(M.d[0] & 1)) throw Error{"bigint selftest: odd modulus expected"};
            Q = big_invmod_odd(A, M);
            break;
        case 3: {
            bool n1, n2;
            glv_decompose(big_mod(A, field_order(FIELD_SCALAR)), &Q, &R, &n1, &n2);
            *flags = (n1 ? 1 : 0) | (n2 ? 2 : 0);
            break;
        }
        default: throw Error{"bigint selftest: unknown op"};
        }
        store(Q, q, nq);
        store(R, r, nr);
        return 0;
    } catch (const Error& e) {
        g_err = e.msg;
        return -1;
    }
}
// rows in use before the power-of-two padding
u32 p2a_rows_used(void* h) {
    Translator* T = (Translator*)h;
    const int noop = T->b.gate_type(P2G_GATE_NOOP);
    size_t n = T->b.rows.size();
    while (n > 0 && T->b.rows[n - 1].gate == noop) n--;
    return (u32)n;
}

}  // extern "C"
