// Device-side constraint evaluators for the 12 gate kinds of include/p2g.h (the 13 gates reachable from the reference's
// translators; BaseSum<2> and BaseSum<4> share a kind), evaluated in the base field at one LDE point per thread.
//
// Replaces `Gate::eval_unfiltered_base_batch` / `eval_unfiltered_base_packed`:
//   * custom gates, from the reference itself -- plonky2-backend/src/plonky2_ecdsa/biguint/gates/
//       arithmetic_u32.rs:289-348, add_many_u32.rs:151-192, subtraction_u32.rs:234-271, range_check_u32.rs:95-117,
//       comparison.rs:337-415 (wire layouts at arithmetic_u32.rs:50-87, add_many_u32.rs:50-87, subtraction_u32.rs:45-79,
//       range_check_u32.rs:43-51, comparison.rs:52-96);
//   * plonky2 0.2.2 built-ins (noop, constant, public_input, arithmetic_base, base_sum, poseidon, random_access) per
//     SURVEY.md App. B.
// Constraints are emitted in the reference's order into a Sink (the quotient kernel folds them into the alpha-weighted sum
// on the fly; the stand-alone entry point stores them), so no per-point constraint vector is ever materialised.
#pragma once
#include "hash.cuh"
#include "../../include/p2g.h"

struct GateDev {
    u32 kind;
    u32 params[4];
    u32 selector_index, group_lo, group_hi, num_constraints;
};

// l (l-1)(l-2)(l-3): the base-4 limb range check shared by the u32 gates
__device__ __forceinline__ u64 limb4_check(u64 l) {
    u64 t = glz_mul(l, gl_sub(l, 3));   // l(l-3) = l^2 - 3l ;  (l-1)(l-2) = l^2 - 3l + 2        (l: C; result N)
    return glz_mul(t, gl_add(t, 2));
}
// prod_{v < base} (l - v)
__device__ __forceinline__ u64 limb_range_product(u64 l, u32 base) {
    u64 r = l;
    for (u32 v = 1; v < base; v++) r = glz_mul(r, gl_sub(l, v));   // l: C; result N
    return r;
}

// Number of independently evaluable "ops" of a gate (the unit the quotient kernel splits work by); 1 = indivisible.
__host__ __device__ inline u32 gate_num_ops(u32 kind, const u32* p) {
    switch (kind) {
    case P2G_GATE_ARITHMETIC: return p[0];
    case P2G_GATE_BASE_SUM: return p[1];          // one op per limb (op 0 also carries the sum constraint)
    case P2G_GATE_RANDOM_ACCESS: return p[1];     // copies (the last one also carries the extra-constant constraints)
    case P2G_GATE_U32_ARITHMETIC: return p[0];
    case P2G_GATE_U32_ADD_MANY: return p[1];
    case P2G_GATE_U32_SUBTRACTION: return p[0];
    case P2G_GATE_U32_RANGE_CHECK: return p[0];
    default: return 1;
    }
}

// W: callable u64(int wire); K: callable u64(int gate_local_constant); S: sink with seek(int constraint_index), emit(u64).
// Wires and constants are canonical (class C of gl.cuh); emitted constraint values may be unreduced (class N): the sinks
// only ever multiply them.  Second operands of gl_sub / gl_add are always C.
// Evaluates ops [op_lo, op_hi) of the gate (all of it for indivisible gates); constraints keep the reference's numbering.
// KIND is a compile-time constant, so each instantiation contains the code of exactly one gate.
template <int KIND, class W, class K, class S>
__device__ __forceinline__ void eval_gate_kind(const GateDev& g, u32 op_lo, u32 op_hi, const W& w, const K& c, const u64* pi_hash,
                                               S& sink) {
    const u32* p = g.params;
    switch (KIND) {
    case P2G_GATE_NOOP:
        break;
    case P2G_GATE_CONSTANT:
        for (u32 i = 0; i < p[0]; i++) sink.emit(gl_sub(c(i), w(i)));
        break;
    case P2G_GATE_PUBLIC_INPUT:
        for (int i = 0; i < 4; i++) sink.emit(gl_sub(w(i), pi_hash[i]));
        break;
    case P2G_GATE_ARITHMETIC: {
        const u64 c0 = c(0), c1 = c(1);
        sink.seek(op_lo);
        for (u32 i0 = op_lo; i0 < op_hi; i0 += 4) {   // 16 independent loads per batch (ncu: the one-op-at-a-time loop waited on memory)
            u64 x[4], y[4], z[4], o[4];
#pragma unroll
            for (u32 k = 0; k < 4; k++) {
                const bool on = i0 + k < op_hi;
                x[k] = on ? w(4 * (i0 + k)) : 0;
                y[k] = on ? w(4 * (i0 + k) + 1) : 0;
                z[k] = on ? w(4 * (i0 + k) + 2) : 0;
                o[k] = on ? w(4 * (i0 + k) + 3) : 0;
            }
#pragma unroll
            for (u32 k = 0; k < 4; k++) {
                if (i0 + k < op_hi) {
                    u64 prod = gl_mul(glz_mul(x[k], y[k]), c0);
                    sink.emit(gl_sub(o[k], gl_add(prod, gl_mul(z[k], c1))));
                }
            }
        }
        break;
    }
    case P2G_GATE_BASE_SUM: {
        // limbs are loaded once, in batches of 8 independent loads; sum_k limb_k base^k accumulates unreduced
        const u32 base = p[0], nl = p[1];
        gl_acc sum;
        sum.clear();
        u64 pw = 1;   // base^k, canonical
        sink.seek(1 + op_lo);
        for (u32 k0 = 0; k0 < nl; k0 += 8) {
            u64 L[8];
#pragma unroll
            for (int i = 0; i < 8; i++) L[i] = (k0 + i < nl) ? w(1 + k0 + i) : 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const u32 k = k0 + i;
                if (k < nl) {
                    if (op_lo == 0) {
                        sum.mac(L[i], pw);
                        pw = gl_mul_small(pw, base);
                    }
                    if (k >= op_lo && k < op_hi) sink.emit(limb_range_product(L[i], base));
                }
            }
        }
        if (op_lo == 0) {
            sink.seek(0);
            sink.emit(gl_sub(sum.reduce(), w(0)));
        }
        break;
    }
    case P2G_GATE_POSEIDON: {
        // wires: in 0..12, out 12..24, swap 24, delta 25..29, full-round-0 sbox inputs 29..65 (rounds 1-3),
        // partial sbox inputs 65..87, full-round-1 sbox inputs 87..135
        const u64 swap = w(24);
        sink.emit(glz_mul(swap, gl_sub(swap, 1)));
        u64 st[12];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u64 lhs = w(i), rhs = w(i + 4), delta = w(25 + i);
            sink.emit(gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), delta));
            st[i] = gl_add(lhs, delta);
            st[i + 4] = gl_sub(rhs, delta);
        }
#pragma unroll
        for (int i = 8; i < 12; i++) st[i] = w(i);
        int rnd = 0;
#pragma unroll 1
        for (int r = 0; r < 4; r++, rnd++) {
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
            if (r != 0) {
#pragma unroll
                for (int i = 0; i < 12; i++) {
                    u64 sb = w(29 + 12 * (r - 1) + i);
                    sink.emit(gl_sub(st[i], sb));
                    st[i] = sb;
                }
            }
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
#pragma unroll 1
        for (int r = 0; r < 22; r++, rnd++) {
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
            u64 sb = w(65 + r);
            sink.emit(gl_sub(st[0], sb));
            st[0] = poseidon_sbox(sb);
            poseidon_mds(st);
        }
#pragma unroll 1
        for (int r = 0; r < 4; r++, rnd++) {
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], POSEIDON_RC(12 * rnd + i));
#pragma unroll
            for (int i = 0; i < 12; i++) {
                u64 sb = w(87 + 12 * r + i);
                sink.emit(gl_sub(st[i], sb));
                st[i] = sb;
            }
#pragma unroll
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) sink.emit(gl_sub(st[i], w(12 + i)));
        break;
    }
    case P2G_GATE_RANDOM_ACCESS: {
        const u32 bits = p[0], copies = p[1], extra = p[2];
        const u32 vec = 1u << bits;
        const u32 routed_used = (2 + vec) * copies + extra;
        sink.seek(op_lo * (bits + 2));
        for (u32 cp = op_lo; cp < op_hi; cp++) {
            const u32 base = (2 + vec) * cp;
            const u32 bw = routed_used + cp * bits;
            u64 rec = 0;
            for (u32 i = 0; i < bits; i++) {
                u64 b = w(bw + i);
                sink.emit(glz_mul(b, gl_sub(b, 1)));
            }
            for (int i = (int)bits - 1; i >= 0; i--) rec = gl_add(gl_dbl(rec), w(bw + i));
            sink.emit(gl_sub(rec, w(base)));
            // fold the list by the index bits, LSB first: x + b (y - x); vec <= 64 items in local registers/stack
            u64 items[64];
            for (u32 i = 0; i < vec; i++) items[i] = w(base + 2 + i);
            u32 len = vec;
            for (u32 b = 0; b < bits; b++) {
                u64 bit = w(bw + b);
                len >>= 1;
                for (u32 i = 0; i < len; i++)
                    items[i] = gl_add(items[2 * i], gl_mul(bit, gl_sub(items[2 * i + 1], items[2 * i])));
            }
            sink.emit(gl_sub(items[0], w(base + 1)));
        }
        if (op_hi == copies)
            for (u32 i = 0; i < extra; i++) sink.emit(gl_sub(c(i), w((2 + vec) * copies + i)));
        break;
    }
    case P2G_GATE_U32_ARITHMETIC: {
        const u32 ops = p[0];
        sink.seek(op_lo * 36);
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = 6 * i;
            u64 computed = gl_add(gl_mul(w(q), w(q + 1)), w(q + 2));
            u64 lo = w(q + 3), hi = w(q + 4), inv = w(q + 5);
            u64 hi_not_max = gl_sub(gl_mul(inv, gl_sub(0xFFFFFFFFULL, hi)), 1);
            sink.emit(glz_mul(hi_not_max, lo));
            sink.emit(gl_sub(gl_add(gl_mul(hi, 1ULL << 32), lo), computed));
            gl_acc32 comb_lo, comb_hi;   // sum_j limb_j 4^j, unreduced
            comb_lo.clear();
            comb_hi.clear();
            const u32 lw = 6 * ops + 32 * i;
#pragma unroll 1
            for (int b = 3; b >= 0; b--) {   // four batches of 8 independent loads
                u64 L[8];
#pragma unroll
                for (int j = 0; j < 8; j++) L[j] = w(lw + 8 * b + j);
                const u32 sh = (b & 1) ? 16 : 0;
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    sink.emit(limb4_check(L[j]));
                    if (b >= 2) comb_hi.mac(L[j], 1u << (sh + 2 * j));
                    else comb_lo.mac(L[j], 1u << (sh + 2 * j));
                }
            }
            sink.emit(gl_sub(comb_lo.reduce(), lo));
            sink.emit(gl_sub(comb_hi.reduce(), hi));
        }
        break;
    }
    case P2G_GATE_U32_ADD_MANY: {
        const u32 na = p[0], ops = p[1];
        sink.seek(op_lo * 21);
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = (na + 3) * i;
            u64 computed = 0;
            for (u32 j = 0; j <= na; j++) computed = gl_add(computed, w(q + j));  // addends then carry-in
            u64 res = w(q + na + 1), carry = w(q + na + 2);
            sink.emit(gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), computed));
            gl_acc32 comb_res, comb_carry;
            comb_res.clear();
            comb_carry.clear();
            const u32 lw = (na + 3) * ops + 18 * i;
#pragma unroll 1
            for (int b = 2; b >= 0; b--) {   // three batches of 6 independent loads
                u64 L[6];
#pragma unroll
                for (int j = 0; j < 6; j++) L[j] = w(lw + 6 * b + j);
#pragma unroll
                for (int jj = 5; jj >= 0; jj--) {
                    const int j = 6 * b + jj;
                    sink.emit(limb4_check(L[jj]));
                    if (j < 16) comb_res.mac(L[jj], 1u << (2 * j));
                    else comb_carry.mac(L[jj], 1u << (2 * (j - 16)));
                }
            }
            sink.emit(gl_sub(comb_res.reduce(), res));
            sink.emit(gl_sub(comb_carry.reduce(), carry));
        }
        break;
    }
    case P2G_GATE_U32_SUBTRACTION: {
        const u32 ops = p[0];
        sink.seek(op_lo * 19);
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = 5 * i;
            u64 initial = gl_sub(gl_sub(w(q), w(q + 1)), w(q + 2));
            u64 res = w(q + 3), bout = w(q + 4);
            sink.emit(gl_sub(res, gl_add(initial, gl_mul(bout, 1ULL << 32))));
            u64 comb = 0;
            const u32 lw = 5 * ops + 16 * i;
#pragma unroll 1
            for (int b = 1; b >= 0; b--) {   // two batches of 8 independent loads
                u64 L[8];
#pragma unroll
                for (int j = 0; j < 8; j++) L[j] = w(lw + 8 * b + j);
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    sink.emit(limb4_check(L[j]));
                    comb = gl_add(glz_mul_small(comb, 4), L[j]);   // N + C -> N  (fewer live registers than gl_acc32 here)
                }
            }
            sink.emit(gl_sub(comb, res));
            sink.emit(glz_mul(bout, gl_sub(1, bout)));
        }
        break;
    }
    case P2G_GATE_U32_RANGE_CHECK: {
        const u32 nl = p[0];
        sink.seek(op_lo * 17);
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 aw = nl + 16 * i;
            gl_acc32 acc;
            acc.clear();
            sink.seek(i * 17 + 1);
#pragma unroll 1
            for (int b = 0; b < 2; b++) {   // two batches of 8 independent loads
                u64 L[8];
#pragma unroll
                for (int j = 0; j < 8; j++) L[j] = w(aw + 8 * b + j);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    acc.mac(L[j], 1u << (16 * b + 2 * j));
                    sink.emit(limb4_check(L[j]));
                }
            }
            sink.seek(i * 17);
            sink.emit(gl_sub(acc.reduce(), w(i)));
            sink.seek(i * 17 + 17);
        }
        break;
    }
    case P2G_GATE_COMPARISON: {
        const u32 nb = p[0], nc = p[1];
        const u32 cb = (nb + nc - 1) / nc, cs = 1u << cb;
        const u32 W_FC = 4, W_SC = 4 + nc, W_EQD = 4 + 2 * nc, W_CHEQ = 4 + 3 * nc, W_INT = 4 + 4 * nc, W_BITS = 4 + 5 * nc;
        u64 fcomb = 0, scomb = 0;
        for (int i = (int)nc - 1; i >= 0; i--) {
            fcomb = gl_add(glz_mul_small(fcomb, cs), w(W_FC + i));
            scomb = gl_add(glz_mul_small(scomb, cs), w(W_SC + i));
        }
        sink.emit(gl_sub(fcomb, w(0)));
        sink.emit(gl_sub(scomb, w(1)));
        u64 msd = 0;
        for (u32 i = 0; i < nc; i++) {
            u64 f = w(W_FC + i), s = w(W_SC + i);
            sink.emit(limb_range_product(f, cs));
            sink.emit(limb_range_product(s, cs));
            u64 diff = gl_sub(s, f);
            u64 eqd = w(W_EQD + i), cheq = w(W_CHEQ + i), inter = w(W_INT + i);
            sink.emit(gl_sub(gl_mul(diff, eqd), gl_sub(1, cheq)));
            sink.emit(glz_mul(cheq, diff));
            sink.emit(gl_sub(inter, gl_mul(cheq, msd)));
            msd = gl_add(inter, gl_mul(gl_sub(1, cheq), diff));
        }
        sink.emit(gl_sub(w(3), msd));
        u64 bcomb = 0;
        for (u32 i = 0; i <= cb; i++) {
            u64 b = w(W_BITS + i);
            sink.emit(glz_mul(b, gl_sub(1, b)));
        }
        for (int i = (int)cb; i >= 0; i--) bcomb = gl_add(gl_dbl(bcomb), w(W_BITS + i));
        sink.emit(gl_sub(gl_add(w(3), cs), bcomb));
        sink.emit(gl_sub(w(2), w(W_BITS + cb)));
        break;
    }
    default:
        break;
    }
}

// sum_j limb_j 4^j by Horner with shifts, exact in 96 bits (up to 16 limbs of 64 bits): acc = 4 acc + l is two funnel
// shifts, one shift and a 96-bit add -- no multiply -- and one reduction at the end.
struct gl_h4 {
    u32 w0, w1, w2;
    __device__ __forceinline__ void clear() { w0 = w1 = w2 = 0; }
    __device__ __forceinline__ void push(u64 l) {   // l: any u64
        w2 = __funnelshift_l(w1, w2, 2);
        w1 = __funnelshift_l(w0, w1, 2);
        w0 <<= 2;
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;"
            : "+r"(w0), "+r"(w1), "+r"(w2) : "r"((u32)l), "r"((u32)(l >> 32)));
    }
    __device__ __forceinline__ u64 reduce() const { return gl_reduce128((u64)w2, ((u64)w1 << 32) | w0); }   // -> C
};

// true for the gates whose 2-bit limb range checks l(l-1)(l-2)(l-3) are evaluated by the limb sweep of quotient.cu
__host__ __device__ inline bool gate_has_limb4_sweep(u32 kind, const u32* p) {
    switch (kind) {
    case P2G_GATE_U32_ARITHMETIC: case P2G_GATE_U32_ADD_MANY: case P2G_GATE_U32_SUBTRACTION: case P2G_GATE_U32_RANGE_CHECK:
        return gate_num_ops(kind, p) > 0;
    case P2G_GATE_BASE_SUM: return p[0] == 4 && p[1] >= 1 && p[1] <= 16;
    default: return false;
    }
}

// The same gates WITHOUT their limb range checks (those come from the sweep): every other constraint, with the reference's
// numbering (arithmetic_u32.rs:289-348, add_many_u32.rs:151-192, subtraction_u32.rs:234-271, range_check_u32.rs:95-117; plonky2
// gates/base_sum.rs).  Limb recombinations use gl_h4.
// Evaluates ops [op_lo, op_hi) of the gate (BaseSum: a single op).
template <int KIND, class W, class S>
__device__ __forceinline__ void eval_gate_nonlimb(const GateDev& g, u32 op_lo, u32 op_hi, const W& w, S& sink) {
    const u32* p = g.params;
    switch (KIND) {
    case P2G_GATE_BASE_SUM: {
        const u32 nl = p[1];
        if (op_hi <= op_lo) break;
        gl_h4 sum;
        sum.clear();
        for (int k = (int)nl - 1; k >= 0; k--) sum.push(w(1 + k));
        sink.seek(0);
        sink.emit(gl_sub(sum.reduce(), w(0)));
        break;
    }
    case P2G_GATE_U32_ARITHMETIC: {
        const u32 ops = p[0];
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = 6 * i;
            u64 computed = gl_add(gl_mul(w(q), w(q + 1)), w(q + 2));
            u64 lo = w(q + 3), hi = w(q + 4), inv = w(q + 5);
            u64 hi_not_max = gl_sub(gl_mul(inv, gl_sub(0xFFFFFFFFULL, hi)), 1);
            sink.seek(36 * i);
            sink.emit(glz_mul(hi_not_max, lo));
            sink.emit(gl_sub(gl_add(gl_mul(hi, 1ULL << 32), lo), computed));
            const u32 lw = 6 * ops + 32 * i;
            gl_h4 clo, chi;
            clo.clear();
            chi.clear();
#pragma unroll 1
            for (int b = 1; b >= 0; b--) {   // batches of 8 + 8 independent loads
                u64 A[8], B[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    A[j] = w(lw + 8 * b + j);
                    B[j] = w(lw + 16 + 8 * b + j);
                }
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    clo.push(A[j]);
                    chi.push(B[j]);
                }
            }
            sink.seek(36 * i + 34);
            sink.emit(gl_sub(clo.reduce(), lo));
            sink.emit(gl_sub(chi.reduce(), hi));
        }
        break;
    }
    case P2G_GATE_U32_ADD_MANY: {
        const u32 na = p[0], ops = p[1];
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = (na + 3) * i;
            u64 computed = 0;
            for (u32 j = 0; j <= na; j++) computed = gl_add(computed, w(q + j));  // addends then carry-in
            u64 res = w(q + na + 1), carry = w(q + na + 2);
            sink.seek(21 * i);
            sink.emit(gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), computed));
            const u32 lw = (na + 3) * ops + 18 * i;
            gl_h4 cres;
            cres.clear();
#pragma unroll 1
            for (int b = 1; b >= 0; b--) {
                u64 A[8];
#pragma unroll
                for (int j = 0; j < 8; j++) A[j] = w(lw + 8 * b + j);
#pragma unroll
                for (int j = 7; j >= 0; j--) cres.push(A[j]);
            }
            u64 l16 = w(lw + 16), l17 = w(lw + 17);
            sink.seek(21 * i + 19);
            sink.emit(gl_sub(cres.reduce(), res));
            sink.emit(gl_sub(gl_add(glz_mul_small(l17, 4), l16), carry));
        }
        break;
    }
    case P2G_GATE_U32_SUBTRACTION: {
        const u32 ops = p[0];
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 q = 5 * i;
            u64 initial = gl_sub(gl_sub(w(q), w(q + 1)), w(q + 2));
            u64 res = w(q + 3), bout = w(q + 4);
            sink.seek(19 * i);
            sink.emit(gl_sub(res, gl_add(initial, gl_mul(bout, 1ULL << 32))));
            const u32 lw = 5 * ops + 16 * i;
            gl_h4 comb;
            comb.clear();
#pragma unroll 1
            for (int b = 1; b >= 0; b--) {
                u64 A[8];
#pragma unroll
                for (int j = 0; j < 8; j++) A[j] = w(lw + 8 * b + j);
#pragma unroll
                for (int j = 7; j >= 0; j--) comb.push(A[j]);
            }
            sink.seek(19 * i + 17);
            sink.emit(gl_sub(comb.reduce(), res));
            sink.emit(glz_mul(bout, gl_sub(1, bout)));
        }
        break;
    }
    case P2G_GATE_U32_RANGE_CHECK: {
        const u32 nl = p[0];
        for (u32 i = op_lo; i < op_hi; i++) {
            const u32 aw = nl + 16 * i;
            gl_h4 comb;
            comb.clear();
#pragma unroll 1
            for (int b = 1; b >= 0; b--) {
                u64 A[8];
#pragma unroll
                for (int j = 0; j < 8; j++) A[j] = w(aw + 8 * b + j);
#pragma unroll
                for (int j = 7; j >= 0; j--) comb.push(A[j]);
            }
            sink.seek(17 * i);
            sink.emit(gl_sub(comb.reduce(), w(i)));
        }
        break;
    }
    default:
        break;
    }
}

// runtime-kind dispatch (stand-alone gate evaluation)
template <class W, class K, class S>
__device__ void eval_gate_unfiltered(const GateDev& g, u32 op_lo, u32 op_hi, const W& w, const K& c, const u64* pi_hash, S& sink) {
    switch (g.kind) {
#define P2G_GATE_CASE(KIND) case KIND: eval_gate_kind<KIND>(g, op_lo, op_hi, w, c, pi_hash, sink); break;
        P2G_GATE_CASE(P2G_GATE_CONSTANT) P2G_GATE_CASE(P2G_GATE_PUBLIC_INPUT) P2G_GATE_CASE(P2G_GATE_ARITHMETIC)
        P2G_GATE_CASE(P2G_GATE_BASE_SUM) P2G_GATE_CASE(P2G_GATE_POSEIDON) P2G_GATE_CASE(P2G_GATE_RANDOM_ACCESS)
        P2G_GATE_CASE(P2G_GATE_U32_ARITHMETIC) P2G_GATE_CASE(P2G_GATE_U32_ADD_MANY) P2G_GATE_CASE(P2G_GATE_U32_SUBTRACTION)
        P2G_GATE_CASE(P2G_GATE_U32_RANGE_CHECK) P2G_GATE_CASE(P2G_GATE_COMPARISON)
#undef P2G_GATE_CASE
    default: break;
    }
}

// filter_g(s) = prod_{i in group, i != row} (i - s) * [several selectors: (UNUSED_SELECTOR - s)]   (plonky2 gates/gate.rs)
__device__ __forceinline__ u64 gate_filter(const GateDev& g, u32 row, u64 s, bool many_selectors) {
    u64 r = 1;
    for (u32 i = g.group_lo; i < g.group_hi; i++)
        if (i != row) r = glz_mul(r, gl_sub((u64)i, s));
    if (many_selectors) r = glz_mul(r, gl_sub(0xFFFFFFFFULL, s));   // result N
    return r;
}
