// Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[X]/(X^2 - 7), for host and sm_100a device.
//
// Replaces plonky2_field 0.2.2 goldilocks_field.rs / extension/quadratic.rs as used by the reference through
// plonky2-backend/src/lib.rs:8-13 (F = GoldilocksField, D = 2).  All values are canonical (< p) at function boundaries,
// so results are bit-identical to the reference's canonical serialisation regardless of evaluation order.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GL_HD __host__ __device__ __forceinline__
#else
#define GL_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef uint8_t u8;

#define GL_P 0xFFFFFFFF00000001ULL
#define GL_EPS 0xFFFFFFFFULL                 // 2^64 mod p
#define GL_GEN 14293326489335486720ULL       // MULTIPLICATIVE_GROUP_GENERATOR (coset shift, base of k_i)
#define GL_POW2_GEN 7277203076849721926ULL   // POWER_OF_TWO_GENERATOR, order 2^32

GL_HD u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a) s += GL_EPS;          // wrapped: 2^64 = eps (mod p); a,b < p keeps the sum below p
    else if (s >= GL_P) s -= GL_P;
    return s;
}
GL_HD u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) d -= GL_EPS;          // borrow: -2^64 = -eps (mod p)
    return d;
}
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }
GL_HD u64 gl_dbl(u64 a) { return gl_add(a, a); }

GL_HD u64 gl_reduce128(u64 hi, u64 lo) {
    // hi*2^64 + lo  with 2^64 = 2^32 - 1, 2^96 = -1 (mod p)
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    u64 t1 = hi_lo * GL_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
GL_HD u64 gl_mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return gl_reduce128(__umul64hi(a, b), a * b);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    return gl_reduce128((u64)(x >> 64), (u64)x);
#endif
}
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }
// a * b + c, canonical
GL_HD u64 gl_mad(u64 a, u64 b, u64 c) { return gl_add(gl_mul(a, b), c); }
// a * small (small < 2^32): one 64x32 product, cheaper reduction
GL_HD u64 gl_mul_small(u64 a, u32 s) {
#if defined(__CUDA_ARCH__)
    u64 lo = a * (u64)s;
    u64 hi = __umul64hi(a, (u64)s); // < 2^32
#else
    unsigned __int128 x = (unsigned __int128)a * s;
    u64 lo = (u64)x, hi = (u64)(x >> 64);
#endif
    u64 t1 = hi * GL_EPS;
    u64 r = lo + t1;
    if (r < t1) r += GL_EPS;
    if (r >= GL_P) r -= GL_P;
    return r;
}
GL_HD u64 gl_pow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = gl_mul(r, b);
        b = gl_sqr(b);
        e >>= 1;
    }
    return r;
}
GL_HD u64 gl_inv(u64 a) { return gl_pow(a, GL_P - 2); }
GL_HD u64 gl_root_of_unity(int bits) {  // primitive 2^bits-th root of unity
    u64 r = GL_POW2_GEN;
    for (int i = bits; i < 32; i++) r = gl_sqr(r);
    return r;
}

struct e2 {
    u64 c0, c1;
};
GL_HD e2 e2_make(u64 a, u64 b) {
    e2 r;
    r.c0 = a;
    r.c1 = b;
    return r;
}
GL_HD e2 e2_add(e2 a, e2 b) { return e2_make(gl_add(a.c0, b.c0), gl_add(a.c1, b.c1)); }
GL_HD e2 e2_sub(e2 a, e2 b) { return e2_make(gl_sub(a.c0, b.c0), gl_sub(a.c1, b.c1)); }
GL_HD e2 e2_mul(e2 a, e2 b) {
    u64 c0 = gl_add(gl_mul(a.c0, b.c0), gl_mul_small(gl_mul(a.c1, b.c1), 7));
    u64 c1 = gl_add(gl_mul(a.c0, b.c1), gl_mul(a.c1, b.c0));
    return e2_make(c0, c1);
}
GL_HD e2 e2_mul_base(e2 a, u64 b) { return e2_make(gl_mul(a.c0, b), gl_mul(a.c1, b)); }
GL_HD e2 e2_add_base(e2 a, u64 b) { return e2_make(gl_add(a.c0, b), a.c1); }
GL_HD e2 e2_pow(e2 b, u64 e) {
    e2 r = e2_make(1, 0);
    while (e) {
        if (e & 1) r = e2_mul(r, b);
        b = e2_mul(b, b);
        e >>= 1;
    }
    return r;
}
GL_HD bool e2_eq(e2 a, e2 b) { return a.c0 == b.c0 && a.c1 == b.c1; }

GL_HD u32 bitrev32(u32 x, int bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    u32 r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
#endif
}
