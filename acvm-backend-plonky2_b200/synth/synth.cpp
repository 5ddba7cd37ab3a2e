// Synthetic circuit-table generator: emits exactly what crosses the FFI of include/p2g.h -- constants + sigmas
// (p2g_circuit_desc.constants_sigmas) and a VALID witness matrix (p2g_prove `wires`) -- for a requested mix of gate rows.
//
// It stands in for the two reference layers that stay in Rust and cannot run here (no cargo / nargo):
// the ACIR -> CircuitBuilder translation (plonky2-backend/src/circuit_translation/*.rs) and plonky2's witness generators
// (`run_once` of the gates in plonky2-backend/src/plonky2_ecdsa/biguint/gates/*.rs, e.g. arithmetic_u32.rs:376-426,
// add_many_u32.rs:329-375, subtraction_u32.rs:298-343, range_check_u32.rs:198-220, comparison.rs:439-537).
// Rows are filled gate by gate with random canonical inputs and the outputs those generators would compute; copy
// constraints are random sigma-cycles between cells forced equal.  Host-side tooling (OpenMP), not part of the hot path.
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../csrc/hash.cuh"
#include "../../include/p2g.h"

extern "C" {
typedef struct p2s_spec {
    uint32_t degree_bits, num_wires, num_routed_wires, num_constants, num_selectors, num_gates, num_public_inputs;
    uint32_t tie_permille;    // fraction of tie-able input cells that copy another cell (copy constraints)
    uint64_t seed;
    const p2g_gate* gates;    // sorted gate table (same as the circuit descriptor)
    const uint8_t* row_gate;  // [N] gate index occupying each row
} p2s_spec;
int p2s_synthesize(const p2s_spec* spec, uint64_t* constants_sigmas, uint64_t* wires, uint64_t* public_inputs);
}

namespace {

inline u64 mix64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
inline u64 rnd(u64 seed, u64 a, u64 b, u64 salt) { return mix64(mix64(seed ^ (salt * 0xD6E8FEB86659FD93ULL)) ^ mix64(a * 0x100000001B3ULL + b)); }
inline u64 rnd_field(u64 seed, u64 a, u64 b) {
    u64 v = rnd(seed, a, b, 1);
    return v >= GL_P ? v - GL_P : v;
}

enum CellClass { CL_NONE = 0, CL_FIELD = 1, CL_U32 = 2 };

// free (input) cells of a gate: column + value class; returns count
int free_cells(const p2g_gate& g, int* cols, int* cls) {
    const uint32_t* p = g.params;
    int n = 0;
    switch (g.kind) {
    case P2G_GATE_ARITHMETIC:
        for (uint32_t i = 0; i < p[0]; i++)
            for (int k = 0; k < 3; k++) { cols[n] = 4 * i + k; cls[n++] = CL_FIELD; }
        break;
    case P2G_GATE_POSEIDON:
        for (int i = 0; i < 12; i++) { cols[n] = i; cls[n++] = CL_FIELD; }
        break;
    case P2G_GATE_RANDOM_ACCESS: {
        uint32_t vec = 1u << p[0];
        for (uint32_t c = 0; c < p[1]; c++)
            for (uint32_t i = 0; i < vec; i++) { cols[n] = (2 + vec) * c + 2 + i; cls[n++] = CL_FIELD; }
        break;
    }
    case P2G_GATE_U32_ARITHMETIC:
        for (uint32_t i = 0; i < p[0]; i++)
            for (int k = 0; k < 3; k++) { cols[n] = 6 * i + k; cls[n++] = CL_U32; }
        break;
    case P2G_GATE_U32_ADD_MANY:
        for (uint32_t i = 0; i < p[1]; i++)
            for (uint32_t k = 0; k < p[0]; k++) { cols[n] = (p[0] + 3) * i + k; cls[n++] = CL_U32; }
        break;
    case P2G_GATE_U32_SUBTRACTION:
        for (uint32_t i = 0; i < p[0]; i++)
            for (int k = 0; k < 2; k++) { cols[n] = 5 * i + k; cls[n++] = CL_U32; }
        break;
    case P2G_GATE_U32_RANGE_CHECK:
        for (uint32_t i = 0; i < p[0]; i++) { cols[n] = i; cls[n++] = CL_U32; }
        break;
    case P2G_GATE_COMPARISON:
        if (p[0] == 32) {
            cols[n] = 0; cls[n++] = CL_U32;
            cols[n] = 1; cls[n++] = CL_U32;
        }
        break;
    default:
        break;
    }
    return n;
}

struct Ctx {
    const p2s_spec* s;
    size_t n;
    int W, R, C, S;
    u64* cs;     // [C + R][N]
    u64* w;      // [W][N]
    u64 pi_hash[4];
    std::vector<std::vector<int>> fcols, fcls;  // per gate
};

inline bool is_dest(const Ctx& c, size_t row, int col) { return rnd(c.s->seed, row, col, 7) % 1000 < c.s->tie_permille; }

// source cell of a tied destination, or false
bool pick_source(const Ctx& c, size_t row, int col, int cls, size_t* srow, int* scol) {
    for (int attempt = 0; attempt < 8; attempt++) {
        u64 h = rnd(c.s->seed, row * 131 + attempt, col, 11);
        size_t r2 = h % c.n;
        const std::vector<int>& fc = c.fcols[c.s->row_gate[r2]];
        const std::vector<int>& fl = c.fcls[c.s->row_gate[r2]];
        if (fc.empty()) continue;
        size_t k = (h >> 32) % fc.size();
        if (fl[k] != cls || fc[k] >= c.R) continue;
        if (r2 == row && fc[k] == col) continue;
        if (is_dest(c, r2, fc[k])) continue;
        *srow = r2;
        *scol = fc[k];
        return true;
    }
    return false;
}

inline u64& WIRE(Ctx& c, int col, size_t row) { return c.w[(size_t)col * c.n + row]; }
inline u64& CONST(Ctx& c, int k, size_t row) { return c.cs[(size_t)(c.S + k) * c.n + row]; }

void fill_inputs(Ctx& c, size_t row) {
    const p2g_gate& g = c.s->gates[c.s->row_gate[row]];
    const uint32_t* p = g.params;
    const u64 seed = c.s->seed;
    const std::vector<int>& fc = c.fcols[c.s->row_gate[row]];
    const std::vector<int>& fl = c.fcls[c.s->row_gate[row]];
    for (size_t k = 0; k < fc.size(); k++)
        WIRE(c, fc[k], row) = fl[k] == CL_U32 ? (rnd(seed, row, fc[k], 2) & 0xFFFFFFFFULL) : rnd_field(seed, row, fc[k]);
    switch (g.kind) {
    case P2G_GATE_CONSTANT:
        for (uint32_t i = 0; i < p[0]; i++) CONST(c, i, row) = rnd_field(seed, row, 1000 + i);
        break;
    case P2G_GATE_ARITHMETIC: {
        // the builder's usual constant pairs (mul: 1,0; add: 1,1; mul_const/arbitrary) all occur; draw one per row
        u64 sel = rnd(seed, row, 2000, 3) % 4;
        CONST(c, 0, row) = sel == 3 ? rnd_field(seed, row, 2001) : 1;
        CONST(c, 1, row) = sel == 0 ? 0 : (sel == 3 ? rnd_field(seed, row, 2002) : 1);
        break;
    }
    case P2G_GATE_RANDOM_ACCESS:
        for (uint32_t i = 0; i < p[2]; i++) CONST(c, i, row) = rnd_field(seed, row, 1000 + i);
        break;
    default:
        break;
    }
}

void base4_limbs(Ctx& c, size_t row, int first_col, u64 v, int count) {
    for (int j = 0; j < count; j++) WIRE(c, first_col + j, row) = (v >> (2 * j)) & 3;
}

void fill_outputs(Ctx& c, size_t row) {
    const p2g_gate& g = c.s->gates[c.s->row_gate[row]];
    const uint32_t* p = g.params;
    const u64 seed = c.s->seed;
    switch (g.kind) {
    case P2G_GATE_CONSTANT:
        for (uint32_t i = 0; i < p[0]; i++) WIRE(c, i, row) = CONST(c, i, row);
        break;
    case P2G_GATE_PUBLIC_INPUT:
        for (int i = 0; i < 4; i++) WIRE(c, i, row) = c.pi_hash[i];
        for (int i = 4; i < c.W; i++) WIRE(c, i, row) = rnd_field(seed, row, i);  // randomize_unused_pi_wires
        break;
    case P2G_GATE_ARITHMETIC: {
        u64 c0 = CONST(c, 0, row), c1 = CONST(c, 1, row);
        for (uint32_t i = 0; i < p[0]; i++)
            WIRE(c, 4 * i + 3, row) = gl_add(gl_mul(gl_mul(WIRE(c, 4 * i, row), WIRE(c, 4 * i + 1, row)), c0), gl_mul(WIRE(c, 4 * i + 2, row), c1));
        break;
    }
    case P2G_GATE_BASE_SUM: {
        uint32_t B = p[0], L = p[1];
        u64 v = 0, pw = 1;
        for (uint32_t k = 0; k < L; k++) {
            u64 limb = rnd(seed, row, 1 + k, 4) % B;
            WIRE(c, 1 + k, row) = limb;
            v = gl_add(v, gl_mul(limb, pw));
            pw = gl_mul(pw, B);
        }
        WIRE(c, 0, row) = v;
        break;
    }
    case P2G_GATE_POSEIDON: {
        u64 st[12];
        for (int i = 0; i < 12; i++) st[i] = WIRE(c, i, row);
        WIRE(c, 24, row) = 0;
        for (int i = 0; i < 4; i++) WIRE(c, 25 + i, row) = 0;
        int rd = 0;
        for (int r = 0; r < 4; r++, rd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], P2G_POSEIDON_RC[12 * rd + i]);
            if (r != 0)
                for (int i = 0; i < 12; i++) WIRE(c, 29 + 12 * (r - 1) + i, row) = st[i];
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int r = 0; r < 22; r++, rd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], P2G_POSEIDON_RC[12 * rd + i]);
            WIRE(c, 65 + r, row) = st[0];
            st[0] = poseidon_sbox(st[0]);
            poseidon_mds(st);
        }
        for (int r = 0; r < 4; r++, rd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], P2G_POSEIDON_RC[12 * rd + i]);
            for (int i = 0; i < 12; i++) WIRE(c, 87 + 12 * r + i, row) = st[i];
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int i = 0; i < 12; i++) WIRE(c, 12 + i, row) = st[i];
        break;
    }
    case P2G_GATE_RANDOM_ACCESS: {
        uint32_t bits = p[0], copies = p[1], extra = p[2], vec = 1u << bits;
        uint32_t routed_used = (2 + vec) * copies + extra;
        for (uint32_t cp = 0; cp < copies; cp++) {
            uint32_t base = (2 + vec) * cp;
            u64 idx = rnd(seed, row, base, 5) % vec;
            WIRE(c, base, row) = idx;
            WIRE(c, base + 1, row) = WIRE(c, base + 2 + (int)idx, row);
            for (uint32_t b = 0; b < bits; b++) WIRE(c, routed_used + cp * bits + b, row) = (idx >> b) & 1;
        }
        for (uint32_t i = 0; i < extra; i++) WIRE(c, (2 + vec) * copies + i, row) = CONST(c, i, row);
        break;
    }
    case P2G_GATE_U32_ARITHMETIC: {
        uint32_t ops = p[0];
        for (uint32_t i = 0; i < ops; i++) {
            u64 out = WIRE(c, 6 * i, row) * WIRE(c, 6 * i + 1, row) + WIRE(c, 6 * i + 2, row);
            u64 lo = out & 0xFFFFFFFFULL, hi = out >> 32;
            WIRE(c, 6 * i + 3, row) = lo;
            WIRE(c, 6 * i + 4, row) = hi;
            WIRE(c, 6 * i + 5, row) = hi == 0xFFFFFFFFULL ? 0 : gl_inv(0xFFFFFFFFULL - hi);
            base4_limbs(c, row, 6 * ops + 32 * i, out, 32);
        }
        break;
    }
    case P2G_GATE_U32_ADD_MANY: {
        uint32_t na = p[0], ops = p[1];
        for (uint32_t i = 0; i < ops; i++) {
            uint32_t q = (na + 3) * i;
            u64 carry_in = rnd(seed, row, q + na, 6) % 16;
            WIRE(c, q + na, row) = carry_in;
            u64 out = carry_in;
            for (uint32_t k = 0; k < na; k++) out += WIRE(c, q + k, row);
            u64 res = out & 0xFFFFFFFFULL, carry = out >> 32;
            WIRE(c, q + na + 1, row) = res;
            WIRE(c, q + na + 2, row) = carry;
            base4_limbs(c, row, (na + 3) * ops + 18 * i, res, 16);
            base4_limbs(c, row, (na + 3) * ops + 18 * i + 16, carry, 2);
        }
        break;
    }
    case P2G_GATE_U32_SUBTRACTION: {
        uint32_t ops = p[0];
        for (uint32_t i = 0; i < ops; i++) {
            u64 x = WIRE(c, 5 * i, row), y = WIRE(c, 5 * i + 1, row);
            u64 bin = rnd(seed, row, 5 * i + 2, 6) & 1;
            WIRE(c, 5 * i + 2, row) = bin;
            u64 bout = x < y + bin ? 1 : 0;
            u64 res = x + (bout << 32) - y - bin;
            WIRE(c, 5 * i + 3, row) = res;
            WIRE(c, 5 * i + 4, row) = bout;
            base4_limbs(c, row, 5 * ops + 16 * i, res, 16);
        }
        break;
    }
    case P2G_GATE_U32_RANGE_CHECK: {
        uint32_t nl = p[0];
        for (uint32_t i = 0; i < nl; i++) base4_limbs(c, row, nl + 16 * i, WIRE(c, i, row), 16);
        break;
    }
    case P2G_GATE_COMPARISON: {
        uint32_t nb = p[0], nc = p[1], cb = (nb + nc - 1) / nc;
        u64 a, b;
        if (nb == 32) {
            a = WIRE(c, 0, row);
            b = WIRE(c, 1, row);
        } else {
            u64 mask = nb >= 64 ? ~0ULL : ((1ULL << nb) - 1);
            a = rnd(seed, row, 0, 8) & mask;
            b = rnd(seed, row, 1, 8) & mask;
            WIRE(c, 0, row) = a;
            WIRE(c, 1, row) = b;
        }
        u64 cmask = (1ULL << cb) - 1;
        u64 msd = 0;
        for (uint32_t i = 0; i < nc; i++) {
            u64 fa = (a >> (cb * i)) & cmask, fb = (b >> (cb * i)) & cmask;
            WIRE(c, 4 + i, row) = fa;
            WIRE(c, 4 + nc + i, row) = fb;
            u64 diff = gl_sub(fb, fa);
            bool eq = fa == fb;
            WIRE(c, 4 + 2 * nc + i, row) = eq ? 1 : gl_inv(diff);
            WIRE(c, 4 + 3 * nc + i, row) = eq ? 1 : 0;
            WIRE(c, 4 + 4 * nc + i, row) = eq ? msd : 0;
            if (!eq) msd = diff;
        }
        WIRE(c, 3, row) = msd;
        u64 shifted = gl_add(msd, 1ULL << cb);
        for (uint32_t i = 0; i <= cb; i++) WIRE(c, 4 + 5 * nc + i, row) = (shifted >> i) & 1;
        WIRE(c, 2, row) = (shifted >> cb) & 1;
        break;
    }
    default:
        break;
    }
}

}  // namespace

// torchrun pins OMP_NUM_THREADS=1 for its workers; the generator may use its share of the host cores anyway
extern "C" void p2s_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

extern "C" int p2s_synthesize(const p2s_spec* s, uint64_t* constants_sigmas, uint64_t* wires, uint64_t* public_inputs) {
    if (!s || !constants_sigmas || !wires || !s->gates || !s->row_gate) return -1;
    Ctx c;
    c.s = s;
    c.n = (size_t)1 << s->degree_bits;
    c.W = s->num_wires;
    c.R = s->num_routed_wires;
    c.C = s->num_constants;
    c.S = s->num_selectors;
    c.cs = constants_sigmas;
    c.w = wires;
    const size_t n = c.n;
    c.fcols.resize(s->num_gates);
    c.fcls.resize(s->num_gates);
    for (uint32_t g = 0; g < s->num_gates; g++) {
        std::vector<int> cols(512), cls(512);
        int k = free_cells(s->gates[g], cols.data(), cls.data());
        c.fcols[g].assign(cols.begin(), cols.begin() + k);
        c.fcls[g].assign(cls.begin(), cls.begin() + k);
    }
    for (size_t r = 0; r < n; r++)
        if (s->row_gate[r] >= s->num_gates) return -2;
    // public inputs + their Poseidon hash (the prover recomputes it; PublicInputGate rows must carry it)
    for (uint32_t i = 0; i < s->num_public_inputs; i++) public_inputs[i] = rnd_field(s->seed, 0xABCDEF, i);
    memset(c.pi_hash, 0, sizeof c.pi_hash);
    if (s->num_public_inputs) {
        digest_t h = hash_no_pad(P2G_H_POSEIDON, public_inputs, s->num_public_inputs);
        memcpy(c.pi_hash, h.w, 32);
    }
    // selectors, zeroed constants / wires, identity permutation (as cell ids col*N + row)
    u64* sig = c.cs + (size_t)c.C * n;
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) {
        const p2g_gate& g = s->gates[s->row_gate[r]];
        for (int k = 0; k < c.C; k++) c.cs[(size_t)k * n + r] = 0;
        for (int k = 0; k < c.S; k++) c.cs[(size_t)k * n + r] = (uint32_t)k == g.selector_index ? s->row_gate[r] : 0xFFFFFFFFULL;
        for (int k = 0; k < c.W; k++) c.w[(size_t)k * n + r] = 0;
        for (int k = 0; k < c.R; k++) sig[(size_t)k * n + r] = (u64)k * n + r;
    }
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) fill_inputs(c, r);
    // copy constraints: tied destinations take their source's value ...
    if (s->tie_permille) {
#pragma omp parallel for schedule(static)
        for (long r = 0; r < (long)n; r++) {
            const std::vector<int>& fc = c.fcols[s->row_gate[r]];
            const std::vector<int>& fl = c.fcls[s->row_gate[r]];
            for (size_t k = 0; k < fc.size(); k++) {
                if (fc[k] >= c.R || !is_dest(c, r, fc[k])) continue;
                size_t sr;
                int sc;
                if (pick_source(c, r, fc[k], fl[k], &sr, &sc)) WIRE(c, fc[k], r) = WIRE(c, sc, sr);  // sources are never destinations
            }
        }
        // ... and join its sigma-cycle (sequential splice: next[dest] = next[src]; next[src] = dest)
        for (size_t r = 0; r < n; r++) {
            const std::vector<int>& fc = c.fcols[s->row_gate[r]];
            const std::vector<int>& fl = c.fcls[s->row_gate[r]];
            for (size_t k = 0; k < fc.size(); k++) {
                if (fc[k] >= c.R || !is_dest(c, r, fc[k])) continue;
                size_t sr;
                int sc;
                if (!pick_source(c, r, fc[k], fl[k], &sr, &sc)) continue;
                u64& nd = sig[(size_t)fc[k] * n + r];
                u64& ns = sig[(size_t)sc * n + sr];
                nd = ns;
                ns = (u64)fc[k] * n + r;
            }
        }
    }
#pragma omp parallel for schedule(static)
    for (long r = 0; r < (long)n; r++) fill_outputs(c, r);
    // sigma cell ids -> k_col * omega^row
    std::vector<u64> kis(c.R), wpow(n);
    kis[0] = 1;
    for (int i = 1; i < c.R; i++) kis[i] = gl_mul(kis[i - 1], GL_GEN);
    u64 wn = gl_root_of_unity(s->degree_bits);
    wpow[0] = 1;
    for (size_t i = 1; i < n; i++) wpow[i] = gl_mul(wpow[i - 1], wn);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)((size_t)c.R * n); i++) {
        u64 id = sig[i];
        sig[i] = gl_mul(kis[id / n], wpow[id % n]);
    }
    return 0;
}
