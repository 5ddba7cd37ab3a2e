/* ORACLE = test infrastructure.  Goldilocks field + quadratic extension, CPU restatement of
 * plonky2_field 0.2.2 goldilocks_field.rs / extension/quadratic.rs.
 * Used by the reference through plonky2-backend/src/lib.rs:8-13 (F = GoldilocksField, D = 2). */
#ifndef ORC_GL_H
#define ORC_GL_H
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;
typedef unsigned __int128 u128;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL /* 2^64 mod p */
#define GL_GEN 14293326489335486720ULL          /* MULTIPLICATIVE_GROUP_GENERATOR = coset shift */
#define GL_POW2_GEN 7277203076849721926ULL      /* POWER_OF_TWO_GENERATOR, order 2^32 */

static inline u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= GL_P) s -= GL_P;
    return s;
}
static inline u64 gl_sub(u64 a, u64 b) { return a >= b ? a - b : a - b + GL_P; }
static inline u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
static inline u64 gl_reduce128(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS; /* borrow: add p (mod 2^64) */
    u64 t1 = hi_lo * GL_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS; /* carry: 2^64 = eps (mod p) */
    if (r >= GL_P) r -= GL_P;
    return r;
}
static inline u64 gl_mul(u64 a, u64 b) { return gl_reduce128((u128)a * b); }
static inline u64 gl_sqr(u64 a) { return gl_mul(a, a); }
static inline u64 gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
static inline u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
static inline u64 gl_root_of_unity(int bits) { /* primitive 2^bits-th root */
    u64 r = GL_POW2_GEN;
    for (int i = bits; i < 32; i++) r = gl_sqr(r);
    return r;
}
static inline u64 gl_from_u64(u64 x) { return x >= GL_P ? x - GL_P : x; }

/* F_p[X]/(X^2 - 7) */
typedef struct { u64 c0, c1; } e2;
static inline e2 e2_make(u64 a, u64 b) { e2 r = {a, b}; return r; }
static inline e2 e2_add(e2 a, e2 b) { return e2_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
static inline e2 e2_sub(e2 a, e2 b) { return e2_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
static inline e2 e2_mul(e2 a, e2 b) {
    u64 c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul(7, gl_mul(a.c1, b.c1)));
    u64 c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return e2_make(c0, c1);
}
static inline e2 e2_mul_base(e2 a, u64 b) { return e2_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
static inline e2 e2_add_base(e2 a, u64 b) { return e2_make(gl_add(a.c0, b), a.c1); }
static inline e2 e2_inv(e2 a) {
    u64 d = gl_inv(gl_sub(gl_sqr(a.c0), gl_mul(7, gl_sqr(a.c1))));
    return e2_make(gl_mul(a.c0, d), gl_neg(gl_mul(a.c1, d)));
}
static inline e2 e2_pow(e2 b, u64 e) {
    e2 r = {1, 0};
    while (e) {
        if (e & 1) r = e2_mul(r, b);
        b = e2_mul(b, b);
        e >>= 1;
    }
    return r;
}
static inline int e2_eq(e2 a, e2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }

static inline u64 bitrev64(u64 x, int bits) {
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
#endif
