/* ORACLE = test infrastructure.  Keccak-256 (original padding, tiny-keccak 2.0.2 semantics), Poseidon-Goldilocks,
 * the two plonky2 hashers built on them and the Fiat-Shamir challenger.
 * Restates plonky2 0.2.2 hash/{keccak,poseidon,hashing,hash_types}.rs and iop/challenger.rs per SURVEY.md App. A.5, A.6, D.
 * Reference: plonky2-backend/src/lib.rs:13 (KeccakGoldilocksConfig), Cargo.lock:1296-1297 (tiny-keccak). */
#ifndef ORC_HASH_H
#define ORC_HASH_H
#include <string.h>
#include "gl.h"
#include "poseidon_constants.h"

/* ---------------- Keccak-f[1600] ---------------- */
static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_RHO[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PI[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline u64 rol64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

static inline void keccak_f1600(u64 a[25]) {
    for (int rnd = 0; rnd < 24; rnd++) {
        u64 c[5];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            u64 d = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
            for (int y = 0; y < 25; y += 5) a[y + x] ^= d;
        }
        u64 t = a[1];
        for (int i = 0; i < 24; i++) {
            int j = KECCAK_PI[i];
            u64 b = a[j];
            a[j] = rol64(t, KECCAK_RHO[i]);
            t = b;
        }
        for (int y = 0; y < 25; y += 5) {
            u64 r[5];
            for (int x = 0; x < 5; x++) r[x] = a[y + x];
            for (int x = 0; x < 5; x++) a[y + x] = r[x] ^ ((~r[(x + 1) % 5]) & r[(x + 2) % 5]);
        }
        a[0] ^= KECCAK_RC[rnd];
    }
}

/* Keccak-256 with the pre-SHA3 0x01 domain byte; little-endian host assumed (x86-64). */
static inline void keccak256(const u8* data, size_t len, u8 out[32]) {
    u64 st[25];
    memset(st, 0, sizeof st);
    const size_t rate = 136;
    while (len >= rate) {
        for (int i = 0; i < 17; i++) {
            u64 w;
            memcpy(&w, data + 8 * i, 8);
            st[i] ^= w;
        }
        keccak_f1600(st);
        data += rate;
        len -= rate;
    }
    u8 blk[136];
    memset(blk, 0, rate);
    memcpy(blk, data, len);
    blk[len] ^= 0x01;
    blk[rate - 1] ^= 0x80;
    for (int i = 0; i < 17; i++) {
        u64 w;
        memcpy(&w, blk + 8 * i, 8);
        st[i] ^= w;
    }
    keccak_f1600(st);
    memcpy(out, st, 32);
}

/* ---------------- Poseidon (width 12, x^7, 4 + 22 + 4 rounds, naive round form) ---------------- */
static const u64 POSEIDON_CIRC[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};

static inline void poseidon_mds(u64 s[12]) {
    u64 r[12];
    for (int i = 0; i < 12; i++) {
        u128 acc = 0;
        for (int j = 0; j < 12; j++) acc += (u128)s[(i + j) % 12] * POSEIDON_CIRC[j];
        if (i == 0) acc += (u128)s[0] * 8; /* MDS_MATRIX_DIAG = [8, 0, ...] */
        r[i] = gl_reduce128(acc);
    }
    memcpy(s, r, sizeof r);
}
static inline u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_sqr(x), x4 = gl_sqr(x2);
    return gl_mul(gl_mul(x4, x2), x);
}
static inline void poseidon_permute(u64 s[12]) {
    for (int r = 0; r < 30; r++) {
        for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], ORC_POSEIDON_RC[12 * r + i]);
        if (r < 4 || r >= 26) {
            for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(s[i]);
        } else {
            s[0] = poseidon_sbox(s[0]);
        }
        poseidon_mds(s);
    }
}

/* ---------------- plonky2 Hasher<F> ---------------- */
#define ORC_KECCAK25 0
#define ORC_POSEIDON 1
#define ORC_MAX_HS 32

static inline int hasher_size(int h) { return h == ORC_KECCAK25 ? 25 : 32; }

static inline void poseidon_hash_no_pad(const u64* in, size_t n, u64 out[4]) {
    u64 st[12];
    memset(st, 0, sizeof st);
    for (size_t off = 0; off < n; off += 8) {
        size_t k = n - off < 8 ? n - off : 8;
        for (size_t i = 0; i < k; i++) st[i] = in[off + i];
        poseidon_permute(st);
    }
    memcpy(out, st, 32);
}

static inline void hash_no_pad(int h, const u64* in, size_t n, u8* out) {
    if (h == ORC_KECCAK25) {
        u8 d[32];
        keccak256((const u8*)in, 8 * n, d);
        memcpy(out, d, 25);
    } else {
        u64 o[4];
        poseidon_hash_no_pad(in, n, o);
        memcpy(out, o, 32);
    }
}
static inline void hash_or_noop(int h, const u64* in, size_t n, u8* out) {
    int hs = hasher_size(h);
    if (8 * n <= (size_t)hs) {
        memset(out, 0, hs);
        memcpy(out, in, 8 * n);
    } else {
        hash_no_pad(h, in, n, out);
    }
}
static inline void two_to_one(int h, const u8* l, const u8* r, u8* out) {
    if (h == ORC_KECCAK25) {
        u8 buf[50], d[32];
        memcpy(buf, l, 25);
        memcpy(buf + 25, r, 25);
        keccak256(buf, 50, d);
        memcpy(out, d, 25);
    } else {
        u64 st[12];
        memcpy(st, l, 32);
        memcpy(st + 4, r, 32);
        memset(st + 8, 0, 32);
        poseidon_permute(st);
        memcpy(out, st, 32);
    }
}
/* GenericHashOut::to_vec: Keccak bytes -> 7-byte LE chunks; Poseidon -> its 4 elements */
static inline void hash_to_elems(int h, const u8* d, u64 out[4]) {
    if (h == ORC_KECCAK25) {
        for (int i = 0; i < 4; i++) {
            u64 w = 0;
            int len = (i < 3) ? 7 : 4;
            memcpy(&w, d + 7 * i, len);
            out[i] = w;
        }
    } else {
        memcpy(out, d, 32);
    }
}
static inline void hasher_permute(int h, u64 st[12]) {
    if (h == ORC_POSEIDON) {
        poseidon_permute(st);
        return;
    }
    /* KeccakPermutation: hash onion + rejection sampling */
    u8 buf[96];
    memcpy(buf, st, 96);
    size_t len = 96;
    int got = 0;
    while (got < 12) {
        u8 d[32];
        keccak256(buf, len, d);
        memcpy(buf, d, 32);
        len = 32;
        for (int i = 0; i < 4 && got < 12; i++) {
            u64 w;
            memcpy(&w, d + 8 * i, 8);
            if (w < GL_P) st[got++] = w;
        }
    }
}

/* ---------------- Challenger ---------------- */
typedef struct {
    int h;
    u64 state[12];
    u64 inbuf[8];
    int nin;
    u64 outbuf[8];
    int nout;
} challenger;

static inline void ch_init(challenger* c, int h) { memset(c, 0, sizeof *c); c->h = h; }
static inline void ch_duplex(challenger* c) {
    for (int i = 0; i < c->nin; i++) c->state[i] = c->inbuf[i];
    c->nin = 0;
    hasher_permute(c->h, c->state);
    memcpy(c->outbuf, c->state, 64);
    c->nout = 8;
}
static inline void ch_observe(challenger* c, u64 e) {
    c->nout = 0;
    c->inbuf[c->nin++] = e;
    if (c->nin == 8) ch_duplex(c);
}
static inline void ch_observe_many(challenger* c, const u64* e, size_t n) { for (size_t i = 0; i < n; i++) ch_observe(c, e[i]); }
static inline void ch_observe_hash(challenger* c, const u8* d) {
    u64 e[4];
    hash_to_elems(c->h, d, e);
    ch_observe_many(c, e, 4);
}
static inline void ch_observe_cap(challenger* c, const u8* cap, int ncap) {
    int hs = hasher_size(c->h);
    for (int i = 0; i < ncap; i++) ch_observe_hash(c, cap + (size_t)i * hs);
}
static inline void ch_observe_e2(challenger* c, e2 x) { ch_observe(c, x.c0); ch_observe(c, x.c1); }
static inline u64 ch_get(challenger* c) {
    if (c->nin > 0 || c->nout == 0) ch_duplex(c);
    return c->outbuf[--c->nout];
}
static inline e2 ch_get_e2(challenger* c) {
    e2 r;
    r.c0 = ch_get(c);
    r.c1 = ch_get(c);
    return r;
}
#endif
