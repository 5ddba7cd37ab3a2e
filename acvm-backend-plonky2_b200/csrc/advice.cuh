// Device-side witness fill, first step: the ADVICE wires of a trace row -- the columns at or above
// num_routed_wires, which no copy constraint can reach -- are functions of the row's routed wires alone, so they need neither
// plonky2's generator queue nor the host: one thread per row recomputes them from the routed columns.  For the reference's
// 234-wire configuration that is 154 of 234 columns (the 2-bit limbs of the u32 gates, the tails of the comparison and
// random-access gates, the Poseidon gate's internal state), i.e. two thirds of the witness never cross PCIe.
//
// Restates, per gate, the part of these generators that writes advice wires (values as plonky2 would set them):
//   U32ArithmeticGate   plonky2-backend/src/plonky2_ecdsa/biguint/gates/arithmetic_u32.rs:376-426   (limbs of low + 2^32 high)
//   U32AddManyGate      .../gates/add_many_u32.rs:329-375        (16 limbs of the result, 2 of the carry)
//   U32SubtractionGate  .../gates/subtraction_u32.rs:298-343     (16 limbs of the result)
//   U32RangeCheckGate   .../gates/range_check_u32.rs:198-220     (16 limbs per input)
//   ComparisonGate      .../gates/comparison.rs:439-537          (chunks, equality dummies, intermediate values, MSD bits)
//   RandomAccessGate    plonky2 gates/random_access.rs RandomAccessGenerator (index bits)
//   PoseidonGate        plonky2 gates/poseidon.rs PoseidonGenerator (swap deltas, S-box inputs of every round)
// The function is host + device: the CUDA kernel k_fill_advice (prover.cu) and the CPU check in tests/ run the same code; the
// independent restatement it is compared with is the generator set of acir/p2acir.cpp.
#pragma once
#include "../../include/p2g.h"
#include "hash.cuh"

// get(col) -> canonical wire value of this row; put(col, v) stores a wire (the caller drops columns it does not own)
template <class G, class S>
GL_HD void fill_advice_row(u32 kind, const u32* p, const G& get, const S& put) {
    switch (kind) {
    case P2G_GATE_U32_ARITHMETIC: {
        const u32 ops = p[0];
        for (u32 i = 0; i < ops; i++) {
            const u64 v = (get(6 * i + 4) << 32) + get(6 * i + 3);   // output = low + 2^32 high
            for (u32 j = 0; j < 32; j++) put(6 * ops + 32 * i + j, (v >> (2 * j)) & 3);
        }
        break;
    }
    case P2G_GATE_U32_ADD_MANY: {
        const u32 na = p[0], ops = p[1];
        for (u32 i = 0; i < ops; i++) {
            const u32 q = (na + 3) * i, lw = (na + 3) * ops + 18 * i;
            const u64 res = get(q + na + 1), carry = get(q + na + 2);
            for (u32 j = 0; j < 16; j++) put(lw + j, (res >> (2 * j)) & 3);
            for (u32 j = 0; j < 2; j++) put(lw + 16 + j, (carry >> (2 * j)) & 3);
        }
        break;
    }
    case P2G_GATE_U32_SUBTRACTION: {
        const u32 ops = p[0];
        for (u32 i = 0; i < ops; i++) {
            const u64 res = get(5 * i + 3);
            for (u32 j = 0; j < 16; j++) put(5 * ops + 16 * i + j, (res >> (2 * j)) & 3);
        }
        break;
    }
    case P2G_GATE_U32_RANGE_CHECK: {
        const u32 n = p[0];
        for (u32 i = 0; i < n; i++) {
            const u32 v = (u32)get(i);   // the generator casts to u32
            for (u32 j = 0; j < 16; j++) put(n + 16 * i + j, (v >> (2 * j)) & 3);
        }
        break;
    }
    case P2G_GATE_COMPARISON: {
        const u32 nb = p[0], nc = p[1], cb = (nb + nc - 1) / nc;
        const u64 a = get(0), b = get(1), cs = (u64)1 << cb;
        u64 msd = 0;
        for (u32 i = 0; i < nc; i++) {
            const u64 fa = cb * i < 64 ? (a >> (cb * i)) & (cs - 1) : 0, fb = cb * i < 64 ? (b >> (cb * i)) & (cs - 1) : 0;
            put(4 + i, fa);
            put(4 + nc + i, fb);
            put(4 + 2 * nc + i, fa == fb ? 1 : gl_inv(gl_sub(fb, fa)));
            put(4 + 3 * nc + i, fa == fb ? 1 : 0);
            if (fa != fb) {
                msd = gl_sub(fb, fa);
                put(4 + 4 * nc + i, 0);
            } else {
                put(4 + 4 * nc + i, msd);
            }
        }
        const u64 t = gl_add(cs, msd);
        for (u32 i = 0; i <= cb; i++) put(4 + 5 * nc + i, (t >> i) & 1);
        break;
    }
    case P2G_GATE_RANDOM_ACCESS: {
        const u32 bits = p[0], copies = p[1], extra = p[2], vec = 1u << bits, routed_used = (2 + vec) * copies + extra;
        for (u32 cp = 0; cp < copies; cp++) {
            const u64 idx = get((2 + vec) * cp);
            for (u32 b = 0; b < bits; b++) put(routed_used + cp * bits + b, (idx >> b) & 1);
        }
        break;
    }
    case P2G_GATE_POSEIDON: {
        u64 st[12];
        for (int i = 0; i < 12; i++) st[i] = get(i);
        const u64 swap = get(24);
        for (int i = 0; i < 4; i++) {
            const u64 delta = gl_mul(swap, gl_sub(st[i + 4], st[i]));
            put(25 + i, delta);
            st[i] = gl_add(st[i], delta);
            st[i + 4] = gl_sub(st[i + 4], delta);
        }
        int rnd = 0;
        for (int r = 0; r < 4; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_canon(gl_add(gl_canon(st[i]), POSEIDON_RC(12 * rnd + i)));
            if (r != 0)
                for (int i = 0; i < 12; i++) put(29 + 12 * (r - 1) + i, st[i]);
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int r = 0; r < 22; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_canon(gl_add(gl_canon(st[i]), POSEIDON_RC(12 * rnd + i)));
            put(65 + r, st[0]);
            st[0] = poseidon_sbox(st[0]);
            poseidon_mds(st);
        }
        for (int r = 0; r < 4; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_canon(gl_add(gl_canon(st[i]), POSEIDON_RC(12 * rnd + i)));
            for (int i = 0; i < 12; i++) put(87 + 12 * r + i, st[i]);
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int i = 0; i < 12; i++) put(12 + i, gl_canon(st[i]));
        break;
    }
    default:
        break;   // Noop, Constant, PublicInput, Arithmetic, BaseSum: every wire is routed
    }
}
