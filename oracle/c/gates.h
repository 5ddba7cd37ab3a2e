/* ORACLE = test infrastructure.  Base-field constraint evaluators of the 13 reachable gates.
 * Custom gates follow /root/reference/plonky2-backend/src/plonky2_ecdsa/biguint/gates/:
 *   arithmetic_u32.rs:289-348, add_many_u32.rs:151-192, subtraction_u32.rs:234-271, range_check_u32.rs:95-117,
 *   comparison.rs:337-415.  Built-ins restate plonky2 0.2.2 gates/{noop,constant,...}.rs (not vendored in the reference). */
#ifndef ORC_GATES_H
#define ORC_GATES_H
#include "../../include/p2g.h"
#include "hash.h"

static inline u64 limb4_product(u64 l) { /* l (l-1)(l-2)(l-3) */
    u64 r = gl_mul(l, gl_sub(l, 1));
    r = gl_mul(r, gl_sub(l, 2));
    return gl_mul(r, gl_sub(l, 3));
}

/* c: gate-local constants (selector prefix removed); w: wires; pi: public_inputs_hash.  Returns #constraints written. */
static int eval_gate_unfiltered(const p2g_gate* g, const u64* c, const u64* w, const u64 pi[4], u64* out) {
    const u32* p = g->params;
    int n = 0;
    switch (g->kind) {
    case P2G_GATE_NOOP:
        break;
    case P2G_GATE_CONSTANT:
        for (u32 i = 0; i < p[0]; i++) out[n++] = gl_sub(c[i], w[i]);
        break;
    case P2G_GATE_PUBLIC_INPUT:
        for (int i = 0; i < 4; i++) out[n++] = gl_sub(w[i], pi[i]);
        break;
    case P2G_GATE_ARITHMETIC:
        for (u32 i = 0; i < p[0]; i++) {
            u64 m = gl_mul(gl_mul(w[4 * i], w[4 * i + 1]), c[0]);
            u64 a = gl_mul(w[4 * i + 2], c[1]);
            out[n++] = gl_sub(w[4 * i + 3], gl_add(m, a));
        }
        break;
    case P2G_GATE_BASE_SUM: {
        u32 B = p[0], nl = p[1];
        u64 acc = 0;
        for (int k = (int)nl - 1; k >= 0; k--) acc = gl_add(gl_mul(acc, B), w[1 + k]);
        out[n++] = gl_sub(acc, w[0]);
        for (u32 k = 0; k < nl; k++) {
            u64 pr = 1;
            for (u32 v = 0; v < B; v++) pr = gl_mul(pr, gl_sub(w[1 + k], v));
            out[n++] = pr;
        }
        break;
    }
    case P2G_GATE_POSEIDON: {
        u64 swap = w[24];
        out[n++] = gl_mul(swap, gl_sub(swap, 1));
        for (int i = 0; i < 4; i++) out[n++] = gl_sub(gl_mul(swap, gl_sub(w[i + 4], w[i])), w[25 + i]);
        u64 st[12];
        for (int i = 0; i < 4; i++) {
            st[i] = gl_add(w[i], w[25 + i]);
            st[i + 4] = gl_sub(w[i + 4], w[25 + i]);
        }
        for (int i = 8; i < 12; i++) st[i] = w[i];
        int rnd = 0;
        for (int r = 0; r < 4; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], ORC_POSEIDON_RC[12 * rnd + i]);
            if (r != 0)
                for (int i = 0; i < 12; i++) {
                    u64 sb = w[29 + 12 * (r - 1) + i];
                    out[n++] = gl_sub(st[i], sb);
                    st[i] = sb;
                }
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int r = 0; r < 22; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], ORC_POSEIDON_RC[12 * rnd + i]);
            u64 sb = w[65 + r];
            out[n++] = gl_sub(st[0], sb);
            st[0] = poseidon_sbox(sb);
            poseidon_mds(st);
        }
        for (int r = 0; r < 4; r++, rnd++) {
            for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], ORC_POSEIDON_RC[12 * rnd + i]);
            for (int i = 0; i < 12; i++) {
                u64 sb = w[87 + 12 * r + i];
                out[n++] = gl_sub(st[i], sb);
                st[i] = sb;
            }
            for (int i = 0; i < 12; i++) st[i] = poseidon_sbox(st[i]);
            poseidon_mds(st);
        }
        for (int i = 0; i < 12; i++) out[n++] = gl_sub(st[i], w[12 + i]);
        break;
    }
    case P2G_GATE_RANDOM_ACCESS: {
        u32 bits = p[0], copies = p[1], extra = p[2];
        u32 vec = 1u << bits;
        u32 routed_used = (2 + vec) * copies + extra;
        u64 items[64];
        for (u32 cp = 0; cp < copies; cp++) {
            u32 base = (2 + vec) * cp;
            u64 idx = w[base], claimed = w[base + 1];
            const u64* bs = w + routed_used + cp * bits;
            for (u32 i = 0; i < bits; i++) out[n++] = gl_mul(bs[i], gl_sub(bs[i], 1));
            u64 rec = 0;
            for (int i = (int)bits - 1; i >= 0; i--) rec = gl_add(gl_add(rec, rec), bs[i]);
            out[n++] = gl_sub(rec, idx);
            for (u32 i = 0; i < vec; i++) items[i] = w[base + 2 + i];
            u32 len = vec;
            for (u32 b = 0; b < bits; b++) {
                for (u32 j = 0; j < len / 2; j++)
                    items[j] = gl_add(items[2 * j], gl_mul(bs[b], gl_sub(items[2 * j + 1], items[2 * j])));
                len /= 2;
            }
            out[n++] = gl_sub(items[0], claimed);
        }
        for (u32 i = 0; i < extra; i++) out[n++] = gl_sub(c[i], w[(2 + vec) * copies + i]);
        break;
    }
    case P2G_GATE_U32_ARITHMETIC: {
        u32 ops = p[0];
        for (u32 i = 0; i < ops; i++) {
            const u64* q = w + 6 * i;
            u64 computed = gl_add(gl_mul(q[0], q[1]), q[2]);
            u64 lo = q[3], hi = q[4], inv = q[5];
            u64 diff = gl_sub(0xFFFFFFFFULL, hi);
            u64 hi_not_max = gl_sub(gl_mul(inv, diff), 1);
            out[n++] = gl_mul(hi_not_max, lo);
            out[n++] = gl_sub(gl_add(gl_mul(hi, 1ULL << 32), lo), computed);
            u64 cl = 0, chh = 0;
            for (int j = 31; j >= 0; j--) {
                u64 limb = w[6 * ops + 32 * i + j];
                out[n++] = limb4_product(limb);
                if (j < 16) cl = gl_add(gl_mul(cl, 4), limb);
                else chh = gl_add(gl_mul(chh, 4), limb);
            }
            out[n++] = gl_sub(cl, lo);
            out[n++] = gl_sub(chh, hi);
        }
        break;
    }
    case P2G_GATE_U32_ADD_MANY: {
        u32 na = p[0], ops = p[1];
        for (u32 i = 0; i < ops; i++) {
            const u64* q = w + (na + 3) * i;
            u64 computed = 0;
            for (u32 j = 0; j < na; j++) computed = gl_add(computed, q[j]);
            computed = gl_add(computed, q[na]);
            u64 res = q[na + 1], carry = q[na + 2];
            out[n++] = gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), computed);
            u64 cr = 0, cc = 0;
            for (int j = 17; j >= 0; j--) {
                u64 limb = w[(na + 3) * ops + 18 * i + j];
                out[n++] = limb4_product(limb);
                if (j < 16) cr = gl_add(gl_mul(cr, 4), limb);
                else cc = gl_add(gl_mul(cc, 4), limb);
            }
            out[n++] = gl_sub(cr, res);
            out[n++] = gl_sub(cc, carry);
        }
        break;
    }
    case P2G_GATE_U32_SUBTRACTION: {
        u32 ops = p[0];
        for (u32 i = 0; i < ops; i++) {
            const u64* q = w + 5 * i;
            u64 init = gl_sub(gl_sub(q[0], q[1]), q[2]);
            u64 res = q[3], bout = q[4];
            out[n++] = gl_sub(res, gl_add(init, gl_mul(bout, 1ULL << 32)));
            u64 comb = 0;
            for (int j = 15; j >= 0; j--) {
                u64 limb = w[5 * ops + 16 * i + j];
                out[n++] = limb4_product(limb);
                comb = gl_add(gl_mul(comb, 4), limb);
            }
            out[n++] = gl_sub(comb, res);
            out[n++] = gl_mul(bout, gl_sub(1, bout));
        }
        break;
    }
    case P2G_GATE_U32_RANGE_CHECK: {
        u32 nl = p[0];
        for (u32 i = 0; i < nl; i++) {
            const u64* aux = w + nl + 16 * i;
            u64 acc = 0;
            for (int j = 15; j >= 0; j--) acc = gl_add(gl_mul(acc, 4), aux[j]);
            out[n++] = gl_sub(acc, w[i]);
            for (int j = 0; j < 16; j++) out[n++] = limb4_product(aux[j]);
        }
        break;
    }
    case P2G_GATE_COMPARISON: {
        u32 nb = p[0], nc = p[1];
        u32 cb = (nb + nc - 1) / nc, cs = 1u << cb;
        const u64 *fc = w + 4, *sc = w + 4 + nc;
        u64 fcomb = 0, scomb = 0;
        for (int i = (int)nc - 1; i >= 0; i--) {
            fcomb = gl_add(gl_mul(fcomb, cs), fc[i]);
            scomb = gl_add(gl_mul(scomb, cs), sc[i]);
        }
        out[n++] = gl_sub(fcomb, w[0]);
        out[n++] = gl_sub(scomb, w[1]);
        u64 msd = 0;
        for (u32 i = 0; i < nc; i++) {
            u64 fp = 1, sp = 1;
            for (u32 x = 0; x < cs; x++) {
                fp = gl_mul(fp, gl_sub(fc[i], x));
                sp = gl_mul(sp, gl_sub(sc[i], x));
            }
            out[n++] = fp;
            out[n++] = sp;
            u64 diff = gl_sub(sc[i], fc[i]);
            u64 eqd = w[4 + 2 * nc + i], cheq = w[4 + 3 * nc + i], inter = w[4 + 4 * nc + i];
            out[n++] = gl_sub(gl_mul(diff, eqd), gl_sub(1, cheq));
            out[n++] = gl_mul(cheq, diff);
            out[n++] = gl_sub(inter, gl_mul(cheq, msd));
            msd = gl_add(inter, gl_mul(gl_sub(1, cheq), diff));
        }
        out[n++] = gl_sub(w[3], msd);
        const u64* bits = w + 4 + 5 * nc;
        u64 bcomb = 0;
        for (u32 i = 0; i <= cb; i++) out[n++] = gl_mul(bits[i], gl_sub(1, bits[i]));
        for (int i = (int)cb; i >= 0; i--) bcomb = gl_add(gl_add(bcomb, bcomb), bits[i]);
        out[n++] = gl_sub(gl_add(w[3], cs), bcomb);
        out[n++] = gl_sub(w[2], bits[cb]);
        break;
    }
    default:
        return -1;
    }
    return n;
}
#endif
