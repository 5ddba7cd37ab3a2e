// Keccak-f[1600] / Keccak-256 (pre-SHA3 padding), Poseidon-Goldilocks (width 12) and the two plonky2 hashers built on them,
// plus the Fiat-Shamir challenger -- shared by the sm_100a kernels and the host-side transcript.
//
// Replaces plonky2 0.2.2 hash/{keccak,poseidon,hashing,hash_types}.rs and iop/challenger.rs, selected by the reference at
// plonky2-backend/src/lib.rs:13 (C = KeccakGoldilocksConfig: Hasher = KeccakHash<25>, InnerHasher = PoseidonHash) and by
// the plonky2_ecdsa tests (PoseidonGoldilocksConfig, e.g. plonky2_ecdsa/gadgets/nonnative.rs:743).
//
// Digest layout on the device: one 32-byte slot (4 x u64) per digest.  KeccakHash<25> keeps bytes 0..24 (word 3 masked to
// its low byte); PoseidonHash keeps 4 field elements.
#pragma once
#include "gl.cuh"
#include "poseidon_constants.h"

#define P2G_H_KECCAK25 0
#define P2G_H_POSEIDON 1

struct digest_t {
    u64 w[4];
};

GL_HD int hasher_bytes(int h) { return h == P2G_H_KECCAK25 ? 25 : 32; }

// ---------------------------------------------------------------------------------------------------------------------
// Keccak
// ---------------------------------------------------------------------------------------------------------------------
#define P2G_KECCAK_RC_LIST                                                                                              \
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,  \
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,  \
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,  \
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,  \
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL
static const u64 h_keccak_rc[24] = {P2G_KECCAK_RC_LIST};
#if defined(__CUDACC__)
static __constant__ u64 d_keccak_rc[24] = {P2G_KECCAK_RC_LIST};
#endif

GL_HD u64 rotl64(u64 x, int n) {
#if defined(__CUDA_ARCH__)
    // funnel shifts on the two 32-bit halves (SHF.L.W): 2 instructions per rotation
    u32 lo = (u32)x, hi = (u32)(x >> 32);
    if (n == 0) return x;
    if (n == 32) return ((u64)lo << 32) | hi;
    if (n < 32) {
        u32 nlo = __funnelshift_l(hi, lo, n);
        u32 nhi = __funnelshift_l(lo, hi, n);
        return ((u64)nhi << 32) | nlo;
    } else {
        u32 nlo = __funnelshift_l(lo, hi, n - 32);
        u32 nhi = __funnelshift_l(hi, lo, n - 32);
        return ((u64)nhi << 32) | nlo;
    }
#else
    return n ? (x << n) | (x >> (64 - n)) : x;
#endif
}

// a ^ (~b & c): one LOP3 per 32-bit half on the device
GL_HD u64 chi_op(u64 a, u64 b, u64 c) { return a ^ ((~b) & c); }

#define KECCAK_RHO_PI(B, A)                                                                                             \
    B[0] = A[0];                  B[10] = rotl64(A[1], 1);     B[20] = rotl64(A[2], 62);                                \
    B[5] = rotl64(A[3], 28);      B[15] = rotl64(A[4], 27);    B[16] = rotl64(A[5], 36);                                \
    B[1] = rotl64(A[6], 44);      B[11] = rotl64(A[7], 6);     B[21] = rotl64(A[8], 55);                                \
    B[6] = rotl64(A[9], 20);      B[7] = rotl64(A[10], 3);     B[17] = rotl64(A[11], 10);                               \
    B[2] = rotl64(A[12], 43);     B[12] = rotl64(A[13], 25);   B[22] = rotl64(A[14], 39);                               \
    B[23] = rotl64(A[15], 41);    B[8] = rotl64(A[16], 45);    B[18] = rotl64(A[17], 15);                               \
    B[3] = rotl64(A[18], 21);     B[13] = rotl64(A[19], 8);    B[14] = rotl64(A[20], 18);                               \
    B[24] = rotl64(A[21], 2);     B[9] = rotl64(A[22], 61);    B[19] = rotl64(A[23], 56);                               \
    B[4] = rotl64(A[24], 14);

GL_HD void keccak_f1600(u64 (&A)[25]) {
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 C0 = A[0] ^ A[5] ^ A[10] ^ A[15] ^ A[20];
        u64 C1 = A[1] ^ A[6] ^ A[11] ^ A[16] ^ A[21];
        u64 C2 = A[2] ^ A[7] ^ A[12] ^ A[17] ^ A[22];
        u64 C3 = A[3] ^ A[8] ^ A[13] ^ A[18] ^ A[23];
        u64 C4 = A[4] ^ A[9] ^ A[14] ^ A[19] ^ A[24];
        u64 D0 = C4 ^ rotl64(C1, 1), D1 = C0 ^ rotl64(C2, 1), D2 = C1 ^ rotl64(C3, 1), D3 = C2 ^ rotl64(C4, 1),
            D4 = C3 ^ rotl64(C0, 1);
#pragma unroll
        for (int y = 0; y < 25; y += 5) {
            A[y] ^= D0;
            A[y + 1] ^= D1;
            A[y + 2] ^= D2;
            A[y + 3] ^= D3;
            A[y + 4] ^= D4;
        }
        u64 B[25];
        KECCAK_RHO_PI(B, A)
#pragma unroll
        for (int y = 0; y < 25; y += 5) {
            A[y] = chi_op(B[y], B[y + 1], B[y + 2]);
            A[y + 1] = chi_op(B[y + 1], B[y + 2], B[y + 3]);
            A[y + 2] = chi_op(B[y + 2], B[y + 3], B[y + 4]);
            A[y + 3] = chi_op(B[y + 3], B[y + 4], B[y]);
            A[y + 4] = chi_op(B[y + 4], B[y], B[y + 1]);
        }
#if defined(__CUDA_ARCH__)
        A[0] ^= d_keccak_rc[rnd];
#else
        A[0] ^= h_keccak_rc[rnd];
#endif
    }
}

// Keccak-256 of `nwords` little-endian u64 words plus `tail_bytes` (< 8) extra bytes held in the low bytes of `tail`.
// Sequential absorber: call absorb_word for every full word, then finish().
struct keccak_sponge {
    u64 A[25];
    int pos;  // lane index inside the 17-lane rate
    GL_HD void init() {
#pragma unroll
        for (int i = 0; i < 25; i++) A[i] = 0;
        pos = 0;
    }
    GL_HD void absorb_word(u64 wd) {
        // dynamic lane index -> switch keeps A[] in registers on the device
#pragma unroll
        for (int i = 0; i < 17; i++)
            if (i == pos) A[i] ^= wd;
        if (++pos == 17) {
            keccak_f1600(A);
            pos = 0;
        }
    }
    // tail: remaining bytes (nbytes < 8) in the low bytes of `tail`; applies 0x01 .. 0x80 padding and permutes
    GL_HD void finish(u64 tail, int nbytes) {
        u64 wd = tail | (0x01ULL << (8 * nbytes));
#pragma unroll
        for (int i = 0; i < 17; i++)
            if (i == pos) A[i] ^= wd;
        A[16] ^= 0x8000000000000000ULL;
        keccak_f1600(A);
    }
};

// KeccakHash<25>::two_to_one: keccak256(l[0..25] || r[0..25])[0..25]
GL_HD digest_t keccak25_two_to_one(const digest_t& l, const digest_t& r) {
    u64 A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0;
    A[0] = l.w[0];
    A[1] = l.w[1];
    A[2] = l.w[2];
    A[3] = (l.w[3] & 0xff) | (r.w[0] << 8);
    A[4] = (r.w[0] >> 56) | (r.w[1] << 8);
    A[5] = (r.w[1] >> 56) | (r.w[2] << 8);
    A[6] = (r.w[2] >> 56) | ((r.w[3] & 0xff) << 8) | (0x01ULL << 16);
    A[16] = 0x8000000000000000ULL;
    keccak_f1600(A);
    digest_t o;
    o.w[0] = A[0];
    o.w[1] = A[1];
    o.w[2] = A[2];
    o.w[3] = A[3] & 0xff;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------------
// Poseidon (x^7, 4 full + 22 partial + 4 full rounds; MDS = circulant [17,15,41,16,2,28,13,13,39,18,34,20] + diag [8,0,..])
// ---------------------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define POSEIDON_RC(i) d_poseidon_rc[i]
#else
#define POSEIDON_RC(i) P2G_POSEIDON_RC[i]
#endif

// x: N -> x^7: N (the MDS layer that always follows splits its inputs into 32-bit halves and accepts any u64)
GL_HD u64 poseidon_sbox(u64 x) {
    u64 x2 = glz_sqr(x), x4 = glz_sqr(x2), x3 = glz_mul(x2, x);
    return glz_mul(x3, x4);
}

// MDS row sums over split 32-bit halves: every partial sum stays below 2^42, so no carries until the final recombination
GL_HD void poseidon_mds(u64 (&s)[12]) {
    const u32 CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u64 r[12];
#if defined(__CUDA_ARCH__)
    // sm_100a: 32 x 32 -> 64 multiplies (IMAD.WIDE) issue at a quarter of the 32-bit IMAD rate, so the state is cut into
    // 22 + 21 + 21-bit limbs: every row sum of limb x coefficient stays below 2^22 * 284 < 2^31 and is a chain of plain 32-bit
    // multiply-adds (3 per term instead of 2 wide ones); the three limb sums are recombined exactly and reduced once.
    u32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        l0[i] = (u32)s[i] & 0x3FFFFFu;
        l1[i] = (u32)(s[i] >> 22) & 0x1FFFFFu;
        l2[i] = (u32)(s[i] >> 43);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u32 s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            const int k = (i + j) % 12;
            s0 += l0[k] * CIRC[j];
            s1 += l1[k] * CIRC[j];
            s2 += l2[k] * CIRC[j];
        }
        if (i == 0) {
            s0 += l0[0] * 8;
            s1 += l1[0] * 8;
            s2 += l2[0] * 8;
        }
        // value = s0 + s1 2^22 + s2 2^43  (< 2^74)
        u64 t = (u64)s0 + ((u64)s1 << 22);   // < 2^54
        u64 u = (u64)s2 << 43;               // low 64 bits of s2 2^43
        u64 lo = t + u;
        u32 hi = (s2 >> 21) + (lo < u);      // multiples of 2^64, < 2^11
        u64 he = ((u64)hi << 32) - hi;       // hi * eps < 2^43
        u64 v = lo + he;
        if (v < he) v += GL_EPS;
        r[i] = gl_canon(v);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = r[i];
    return;
#endif
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u64 lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 12; j++) {
            u64 v = s[(i + j) % 12];
            lo += (u64)(u32)v * CIRC[j];
            hi += (v >> 32) * CIRC[j];
        }
        if (i == 0) {
            lo += (u64)(u32)s[0] * 8;
            hi += (s[0] >> 32) * 8;
        }
        // value = lo + 2^32 * hi,  hi = hh * 2^32 + hl  ->  lo + (hl << 32) + eps * hh
        u64 hh = hi >> 32, hl = hi & GL_EPS;
        u64 t = (hl << 32) + hh * GL_EPS;  // (hl<<32) <= 2^64 - 2^32, hh*eps < 2^42: may wrap
        if (t < (hl << 32)) t += GL_EPS;
        u64 v = lo + t;
        if (v < t) v += GL_EPS;
        if (v >= GL_P) v -= GL_P;
        r[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = r[i];
}

GL_HD void poseidon_permute(u64 (&s)[12]) {
#if defined(__CUDA_ARCH__)
    // One loop body for all 30 rounds (a uniform branch skips the 11 extra S-boxes of the partial rounds): with three bodies -- full,
    // partial, full, 31 + 16 + 31 KB of SASS -- ncu showed the leaf kernel stalled on instruction fetch (no_instruction 8.3 cycles per
    // issue, issue slots 47 % busy): warps at different rounds thrash the instruction cache.  Same arithmetic, a third of the code.
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC(12 * r + i));
        s[0] = poseidon_sbox(s[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) s[i] = poseidon_sbox(s[i]);
        }
        poseidon_mds(s);
    }
#else
    for (int r = 0; r < 4; r++) {
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], POSEIDON_RC(12 * r + i)));
        poseidon_mds(s);
    }
    for (int r = 4; r < 26; r++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], POSEIDON_RC(12 * r + i));
        s[0] = poseidon_sbox(s[0]);
        poseidon_mds(s);
    }
    for (int r = 26; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add(s[i], POSEIDON_RC(12 * r + i)));
        poseidon_mds(s);
    }
#endif
}

GL_HD digest_t poseidon_two_to_one(const digest_t& l, const digest_t& r) {
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        s[i] = l.w[i];
        s[4 + i] = r.w[i];
        s[8 + i] = 0;
    }
    poseidon_permute(s);
    digest_t o;
#pragma unroll
    for (int i = 0; i < 4; i++) o.w[i] = s[i];
    return o;
}

GL_HD digest_t two_to_one(int h, const digest_t& l, const digest_t& r) {
    return h == P2G_H_KECCAK25 ? keccak25_two_to_one(l, r) : poseidon_two_to_one(l, r);
}

// Hasher::hash_no_pad over a contiguous element array (host transcript, small device cases)
GL_HD digest_t hash_no_pad(int h, const u64* in, size_t n) {
    digest_t o;
    if (h == P2G_H_KECCAK25) {
        keccak_sponge sp;
        sp.init();
        for (size_t i = 0; i < n; i++) sp.absorb_word(in[i]);
        sp.finish(0, 0);
        o.w[0] = sp.A[0];
        o.w[1] = sp.A[1];
        o.w[2] = sp.A[2];
        o.w[3] = sp.A[3] & 0xff;
    } else {
        u64 s[12];
        for (int i = 0; i < 12; i++) s[i] = 0;
        for (size_t off = 0; off < n; off += 8) {
            size_t k = n - off < 8 ? n - off : 8;
            for (size_t i = 0; i < k; i++) s[i] = in[off + i];
            poseidon_permute(s);
        }
        for (int i = 0; i < 4; i++) o.w[i] = s[i];
    }
    return o;
}

// GenericHashOut::to_vec: KeccakHash<25> bytes -> 7-byte little-endian chunks (7,7,7,4); PoseidonHash -> its 4 elements
GL_HD void digest_to_elems(int h, const digest_t& d, u64 out[4]) {
    if (h == P2G_H_KECCAK25) {
        const u64 M56 = 0x00FFFFFFFFFFFFFFULL;
        out[0] = d.w[0] & M56;
        out[1] = ((d.w[0] >> 56) | (d.w[1] << 8)) & M56;
        out[2] = ((d.w[1] >> 48) | (d.w[2] << 16)) & M56;
        out[3] = ((d.w[2] >> 40) | ((d.w[3] & 0xff) << 24)) & 0xFFFFFFFFULL;
    } else {
        for (int i = 0; i < 4; i++) out[i] = d.w[i];
    }
}

// KeccakPermutation / PoseidonPermutation acting on the challenger's 12-element state
GL_HD void hasher_permute(int h, u64 (&st)[12]) {
    if (h == P2G_H_POSEIDON) {
        poseidon_permute(st);
        return;
    }
    // hash onion keccak256(state bytes), keccak256(previous digest), ...; words >= p are rejected
    u64 A[25];
    for (int i = 0; i < 25; i++) A[i] = 0;
    for (int i = 0; i < 12; i++) A[i] = st[i];
    A[12] ^= 0x01;
    A[16] ^= 0x8000000000000000ULL;
    int got = 0;
    while (true) {
        keccak_f1600(A);
        u64 d0 = A[0], d1 = A[1], d2 = A[2], d3 = A[3];
        u64 d[4] = {d0, d1, d2, d3};
        for (int i = 0; i < 4 && got < 12; i++)
            if (d[i] < GL_P) st[got++] = d[i];
        if (got >= 12) break;
        for (int i = 0; i < 25; i++) A[i] = 0;
        A[0] = d0;
        A[1] = d1;
        A[2] = d2;
        A[3] = d3;
        A[4] = 0x01;
        A[16] = 0x8000000000000000ULL;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Challenger<F, H>: overwrite-mode duplex sponge, width 12, rate 8 (plonky2 iop/challenger.rs)
// ---------------------------------------------------------------------------------------------------------------------
struct challenger_t {
    int h;
    int nin, nout;
    u64 state[12];
    u64 inbuf[8];
    u64 outbuf[8];

    GL_HD void init(int hasher) {
        h = hasher;
        nin = nout = 0;
        for (int i = 0; i < 12; i++) state[i] = 0;
    }
    GL_HD void duplex() {
        for (int i = 0; i < nin; i++) state[i] = inbuf[i];
        nin = 0;
        hasher_permute(h, state);
        for (int i = 0; i < 8; i++) outbuf[i] = state[i];
        nout = 8;
    }
    GL_HD void observe(u64 e) {
        nout = 0;
        inbuf[nin++] = e;
        if (nin == 8) duplex();
    }
    GL_HD void observe_many(const u64* e, size_t n) {
        for (size_t i = 0; i < n; i++) observe(e[i]);
    }
    GL_HD void observe_digest(const digest_t& d) {
        u64 e[4];
        digest_to_elems(h, d, e);
        observe_many(e, 4);
    }
    GL_HD void observe_e2(e2 x) {
        observe(x.c0);
        observe(x.c1);
    }
    GL_HD u64 get() {
        if (nin > 0 || nout == 0) duplex();
        return outbuf[--nout];
    }
    GL_HD e2 get_e2() {
        e2 r;
        r.c0 = get();
        r.c1 = get();
        return r;
    }
};
