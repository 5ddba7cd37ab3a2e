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

// Value classes used in the comments below:  C = canonical (< p),  N = any u64 congruent to the value (not reduced).
// Every function states what it accepts and returns; whatever is stored to memory, hashed or serialised is always C.
// On sm_100a the bodies are 64-bit carry chains written in PTX (IADD3 / IADD3.X, IMAD.WIDE.U32 with carry in/out) instead
// of compare + select; the host bodies are the plain forms and agree with them on every input of the stated class.

// (a - b) mod p.  a: N, b: <= p.  Result N; C when a is C.   [device: 5 instructions]
//   borrow  =>  true value = d - 2^64 = d - eps (mod p); d >= 2^64 - p = eps there, so the correction cannot borrow again.
GL_HD u64 gl_sub(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u64 d;
    asm("{\n\t.reg .u32 m;\n\t.reg .u64 mm;\n\t"
        "sub.cc.u64 %0, %1, %2;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // borrow ? 0xffffffff : 0   (= eps or 0)
        "cvt.u64.u32 mm, m;\n\t"
        "sub.u64 %0, %0, mm;\n\t}"
        : "=l"(d) : "l"(a), "l"(b));
    return d;
#else
    u64 d = a - b;
    if (a < b) d -= GL_EPS;
    return d;
#endif
}
// (a + b) mod p.  a: N, b: C.  Result N; C when a is C  (a + b = a - (p - b), and p - b <= p).   [7 instructions, 5 for constant b]
GL_HD u64 gl_add(u64 a, u64 b) { return gl_sub(a, GL_P - b); }
// x: N -> C
GL_HD u64 gl_canon(u64 x) { return x >= GL_P ? x - GL_P : x; }
GL_HD u64 gl_neg(u64 a) { return a ? GL_P - a : 0; }   // C -> C
GL_HD u64 gl_dbl(u64 a) { return gl_add(a, a); }       // C -> C

// hi 2^64 + lo -> N, with 2^64 = eps, 2^96 = -1 (mod p); hi = (r3:r2), lo = (r1:r0).
//   t = lo - r3 (borrow => -eps; t > 2^64 - 2^32 there, no second borrow);  u = r2 * eps = (r2 << 32) - r2 <= 2^64 - 2^33 + 1;
//   x = t + u (carry => +eps;  t + u - 2^64 <= 2^64 - 2^33, so the correction cannot carry again).
GL_HD u64 glz_reduce128(u64 hi, u64 lo) {
#if defined(__CUDA_ARCH__)
    u32 x0, x1;
    asm("{\n\t.reg .u32 b, u0, u1, m;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"
        "subc.cc.u32 %1, %3, 0;\n\t"
        "subc.u32 b, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, b;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "sub.cc.u32 u0, 0, %4;\n\t"
        "subc.u32 u1, %4, 0;\n\t"
        "add.cc.u32 %0, %0, u0;\n\t"
        "addc.cc.u32 %1, %1, u1;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.u32 m, 0, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(x0), "=&r"(x1) : "r"((u32)lo), "r"((u32)(lo >> 32)), "r"((u32)hi), "r"((u32)(hi >> 32)));
    return ((u64)x1 << 32) | x0;
#else
    u64 hi_hi = hi >> 32, hi_lo = hi & GL_EPS;
    u64 t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= GL_EPS;
    u64 t1 = hi_lo * GL_EPS;
    u64 r = t0 + t1;
    if (r < t1) r += GL_EPS;
    return r;
#endif
}
GL_HD u64 gl_reduce128(u64 hi, u64 lo) { return gl_canon(glz_reduce128(hi, lo)); }   // -> C

// a, b: N -> N  (no final conditional subtraction)
GL_HD u64 glz_mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %7;\n\t"
        "mul.hi.u32 %3, %5, %7;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return glz_reduce128(((u64)r3 << 32) | r2, ((u64)r1 << 32) | r0);
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    return glz_reduce128((u64)(x >> 64), (u64)x);
#endif
}
GL_HD u64 gl_mul(u64 a, u64 b) { return gl_canon(glz_mul(a, b)); }   // a, b: N -> C
GL_HD u64 gl_sqr(u64 a) { return gl_mul(a, a); }
GL_HD u64 glz_sqr(u64 a) { return glz_mul(a, a); }
GL_HD u64 gl_mad(u64 a, u64 b, u64 c) { return gl_add(gl_mul(a, b), c); }   // c: C
// a: N, s < 2^32 -> N : 96-bit product, lo + h * eps with one carry correction
GL_HD u64 glz_mul_small(u64 a, u32 s) {
    u64 lo = (u64)(u32)a * s;
    u64 hi = (a >> 32) * (u64)s;          // value = lo + hi 2^32
    u64 t = lo + (hi << 32);              // may carry once
    u64 h = (hi >> 32) + (t < lo);        // multiples of 2^64, h < 2^32
    u64 u = (h << 32) - h;                // h * eps
    u64 r = t + u;
    if (r < u) r += GL_EPS;               // t + u - 2^64 + eps < 2^64
    return r;
}
GL_HD u64 gl_mul_small(u64 a, u32 s) { return gl_canon(glz_mul_small(a, s)); }

// ---- "f" forms: the same functions with the carry corrections issued on the FMA pipe ------------------------------------
// The integer ALU pipe (IADD3 / LOP3 / SEL, 64 lanes/clk/SM) is what the NTT butterflies saturate (ncu: 76% ALU, 25% FMA);
// x + m * eps with m in {0, 1} is one IMAD.WIDE.U32 on the other pipe instead of a negate + two carry adds.  eps comes from
// constant memory so that ptxas keeps the multiply form instead of strength-reducing it back to adds.
#if defined(__CUDACC__)
static __constant__ u32 d_gl_eps = 0xFFFFFFFFu;
#endif
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ u64 glf_fix(u64 x, u32 m) {   // x + m * eps (mod 2^64), m in {0, 1}
    u64 r;
    asm("{\n\t.reg .u32 x0, x1;\n\t"
        "mov.b64 {x0, x1}, %3;\n\t"
        "mad.lo.cc.u32 x0, %1, %2, x0;\n\t"
        "madc.hi.u32 x1, %1, %2, x1;\n\t"
        "mov.b64 %0, {x0, x1};\n\t}"
        : "=l"(r) : "r"(m), "r"(d_gl_eps), "l"(x));
    return r;
}
#endif
// x: N -> C        (x >= p  <=>  x + eps carries;  x - p = x + eps mod 2^64)
GL_HD u64 glf_canon(u64 x) {
#if defined(__CUDA_ARCH__)
    u32 m;
    asm("{\n\t.reg .u64 y;\n\tadd.cc.u64 y, %1, 0xffffffff;\n\taddc.u32 %0, 0, 0;\n\t}" : "=r"(m) : "l"(x));
    return glf_fix(x, m);
#else
    return gl_canon(x);
#endif
}
// a, b: C -> C     (s = a + b carries, or s >= p: never both; either way the result is s + eps mod 2^64)
GL_HD u64 glf_add(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u64 s;
    u32 c;
    asm("{\n\t.reg .u64 y;\n\t"
        "add.cc.u64 %0, %2, %3;\n\t"
        "addc.u32 %1, 0, 0;\n\t"
        "add.cc.u64 y, %0, 0xffffffff;\n\t"
        "addc.u32 %1, %1, 0;\n\t}"
        : "=&l"(s), "=&r"(c) : "l"(a), "l"(b));
    return glf_fix(s, c);
#else
    return gl_add(a, b);
#endif
}
// a, b: N -> C
GL_HD u64 glf_mul(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %7;\n\t"
        "mul.hi.u32 %3, %5, %7;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    // t = lo - r3 (borrow => - eps), x = t + r2 * eps (one IMAD.WIDE with carry out), + carry * eps, canonicalise
    u64 t = gl_sub(((u64)r1 << 32) | r0, (u64)r3), x;
    u32 m;
    asm("{\n\t.reg .u32 x0, x1, t0, t1;\n\t"
        "mov.b64 {t0, t1}, %2;\n\t"
        "mad.lo.cc.u32 x0, %3, %4, t0;\n\t"
        "madc.hi.cc.u32 x1, %3, %4, t1;\n\t"
        "addc.u32 %1, 0, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\t}"
        : "=l"(x), "=r"(m) : "l"(t), "r"(r2), "r"(d_gl_eps));
    return glf_canon(glf_fix(x, m));
#else
    return gl_mul(a, b);
#endif
}

// x * 2^S for a compile-time 0 <= S < 96, x: N -> C.  2 is a 192nd root of unity in this field (2^96 = -1) and plonky2's
// roots of unity of order <= 64 are powers of two (omega_64 = 2^3, omega_8 = 2^24, omega_4 = 2^48), so the twiddles of the last
// six stages of every transform are shifts: two 64-bit shifts and one 128-bit reduction instead of four wide multiplies.
template <int S>
GL_HD u64 gl_shl(u64 x) {
    static_assert(S >= 0 && S < 96, "shift out of range");
    if (S == 0) return gl_canon(x);
    if (S < 64) return gl_reduce128(x >> (64 - (S ? S : 1)), x << S);
    return gl_shl<(S >= 64 ? S - 48 : 0)>(gl_shl<48>(x));
}

// Sum of products with ONE reduction at the end: acc += a * b for a, b: N.  The 128-bit products are accumulated exactly --
// even limbs (a0 b0 + a1 b1 2^64) in `e`, cross terms (a0 b1 + a1 b0) in `o` -- so on the device each term costs 4
// IMAD.WIDE.U32 and 3 carry adds and no modular reduction.  Holds up to 2^31 terms.
struct gl_acc {
    u32 e0, e1, e2, e3, e4, o0, o1, o2;
    GL_HD void clear() { e0 = e1 = e2 = e3 = e4 = o0 = o1 = o2 = 0; }
    GL_HD void mac(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
        u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
        asm("{\n\t"
            "mad.lo.cc.u32 %0, %8, %10, %0;\n\t"
            "madc.hi.cc.u32 %1, %8, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %9, %11, %2;\n\t"
            "madc.hi.cc.u32 %3, %9, %11, %3;\n\t"
            "addc.u32 %4, %4, 0;\n\t"
            "mad.lo.cc.u32 %5, %8, %11, %5;\n\t"
            "madc.hi.cc.u32 %6, %8, %11, %6;\n\t"
            "addc.u32 %7, %7, 0;\n\t"
            "mad.lo.cc.u32 %5, %9, %10, %5;\n\t"
            "madc.hi.cc.u32 %6, %9, %10, %6;\n\t"
            "addc.u32 %7, %7, 0;\n\t"
            "}"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4), "+r"(o0), "+r"(o1), "+r"(o2)
            : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
#else
        u64 a0 = (u32)a, a1 = a >> 32, b0 = (u32)b, b1 = b >> 32;
        unsigned __int128 E = ((unsigned __int128)(((u64)e3 << 32) | e2) << 64) | (((u64)e1 << 32) | e0);
        unsigned __int128 add = ((unsigned __int128)(a1 * b1) << 64) | (a0 * b0);
        unsigned __int128 E2 = E + add;
        if (E2 < E) e4++;
        e0 = (u32)E2; e1 = (u32)(E2 >> 32); e2 = (u32)(E2 >> 64); e3 = (u32)(E2 >> 96);
        unsigned __int128 O = ((unsigned __int128)o2 << 64) | (((u64)o1 << 32) | o0);
        O += a0 * b1;
        O += a1 * b0;
        o0 = (u32)O; o1 = (u32)(O >> 32); o2 = (u32)(O >> 64);
#endif
    }
    // total = E + O 2^32,  E = e0 + e1 2^32 + e2 2^64 + e3 2^96 + e4 2^128,  2^128 = -2^32 (mod p)  ->  C
    GL_HD u64 reduce() const {
        u64 r = gl_reduce128(((u64)e3 << 32) | e2, ((u64)e1 << 32) | e0);
        u64 x = gl_reduce128(((u64)o2 << 32) | o1, (u64)o0 << 32);
        r = gl_add(r, x);
        return gl_sub(r, gl_canon((u64)e4 << 32));
    }
};

// Same idea for a 32-bit multiplier: acc += a * c, a: N, c < 2^32 (limb recombination sum_j limb_j B^j): 2 IMAD.WIDE.U32 and 2
// carry adds per term.  Low halves a0 * c accumulate in `e`, high halves a1 * c in `o`; total = e + o 2^32.  Up to 2^31 terms.
struct gl_acc32 {
    u32 e0, e1, e2, o0, o1, o2;
    GL_HD void clear() { e0 = e1 = e2 = o0 = o1 = o2 = 0; }
    GL_HD void mac(u64 a, u32 c) {
#if defined(__CUDA_ARCH__)
        u32 a0 = (u32)a, a1 = (u32)(a >> 32);
        asm("{\n\t"
            "mad.lo.cc.u32 %0, %6, %8, %0;\n\t"
            "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"
            "addc.u32 %2, %2, 0;\n\t"
            "mad.lo.cc.u32 %3, %7, %8, %3;\n\t"
            "madc.hi.cc.u32 %4, %7, %8, %4;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "}"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(o0), "+r"(o1), "+r"(o2)
            : "r"(a0), "r"(a1), "r"(c));
#else
        unsigned __int128 E = ((unsigned __int128)e2 << 64) | (((u64)e1 << 32) | e0);
        E += (u64)(u32)a * c;
        e0 = (u32)E; e1 = (u32)(E >> 32); e2 = (u32)(E >> 64);
        unsigned __int128 O = ((unsigned __int128)o2 << 64) | (((u64)o1 << 32) | o0);
        O += (a >> 32) * (u64)c;
        o0 = (u32)O; o1 = (u32)(O >> 32); o2 = (u32)(O >> 64);
#endif
    }
    // e0 + e1 2^32 + e2 2^64  +  (o0 + o1 2^32 + o2 2^64) 2^32   ->  C
    GL_HD u64 reduce() const {
        u64 r = gl_reduce128((u64)e2, ((u64)e1 << 32) | e0);
        u64 x = gl_reduce128(((u64)o2 << 32) | o1, (u64)o0 << 32);
        return gl_add(r, x);
    }
};

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
