// Unsigned big integers on 32-bit limbs for the witness generators of the reference's non-native gadgets (num::BigUint's role in
// plonky2_ecdsa/biguint/biguint.rs and gadgets/nonnative.rs).  Host tooling; sizes are at most ~20 limbs, schoolbook everywhere.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

struct Big {
    std::vector<uint32_t> d;   // least significant limb first, no trailing zero limbs (`to_u32_digits`)

    Big() {}
    explicit Big(uint64_t v) {
        while (v) {
            d.push_back((uint32_t)v);
            v >>= 32;
        }
    }
    static Big from_u64_limbs(const uint64_t* w, int n) {
        Big r;
        for (int i = 0; i < n; i++) {
            r.d.push_back((uint32_t)w[i]);
            r.d.push_back((uint32_t)(w[i] >> 32));
        }
        r.trim();
        return r;
    }
    static Big from_be_bytes(const uint8_t* p, int n) {
        Big r;
        r.d.assign((n + 3) / 4, 0);
        for (int i = 0; i < n; i++) r.d[i / 4] |= (uint32_t)p[n - 1 - i] << (8 * (i % 4));
        r.trim();
        return r;
    }
    void trim() {
        while (!d.empty() && d.back() == 0) d.pop_back();
    }
    bool is_zero() const { return d.empty(); }
    uint32_t limb(size_t i) const { return i < d.size() ? d[i] : 0; }
    bool bit(size_t i) const { return (limb(i / 32) >> (i % 32)) & 1; }
    size_t bits() const {
        if (d.empty()) return 0;
        size_t b = 32 * (d.size() - 1);
        for (uint32_t t = d.back(); t; t >>= 1) b++;
        return b;
    }
};

inline int big_cmp(const Big& a, const Big& b) {
    if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
    for (size_t i = a.d.size(); i-- > 0;)
        if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
    return 0;
}
inline Big big_add(const Big& a, const Big& b) {
    Big r;
    const size_t n = std::max(a.d.size(), b.d.size());
    uint64_t c = 0;
    for (size_t i = 0; i < n; i++) {
        c += (uint64_t)a.limb(i) + b.limb(i);
        r.d.push_back((uint32_t)c);
        c >>= 32;
    }
    if (c) r.d.push_back((uint32_t)c);
    return r;
}
inline Big big_sub(const Big& a, const Big& b) {   // a >= b
    Big r;
    int64_t bw = 0;
    for (size_t i = 0; i < a.d.size(); i++) {
        int64_t t = (int64_t)a.d[i] - b.limb(i) - bw;
        bw = t < 0;
        r.d.push_back((uint32_t)t);
    }
    r.trim();
    return r;
}
inline Big big_mul(const Big& a, const Big& b) {
    Big r;
    if (a.is_zero() || b.is_zero()) return r;
    r.d.assign(a.d.size() + b.d.size(), 0);
    for (size_t i = 0; i < a.d.size(); i++) {
        uint64_t c = 0;
        for (size_t j = 0; j < b.d.size(); j++) {
            c += (uint64_t)a.d[i] * b.d[j] + r.d[i + j];
            r.d[i + j] = (uint32_t)c;
            c >>= 32;
        }
        r.d[i + b.d.size()] = (uint32_t)c;
    }
    r.trim();
    return r;
}
inline Big big_shl(const Big& a, size_t s) {   // a * 2^s
    Big r;
    if (a.is_zero()) return r;
    const size_t w = s / 32, k = s % 32;
    r.d.assign(a.d.size() + w + 1, 0);
    for (size_t i = 0; i < a.d.size(); i++) {
        const uint64_t t = (uint64_t)a.d[i] << k;
        r.d[i + w] |= (uint32_t)t;
        r.d[i + w + 1] |= (uint32_t)(t >> 32);
    }
    r.trim();
    return r;
}
// Knuth's algorithm D.  b != 0.
inline void big_divrem(const Big& a, const Big& b, Big* q, Big* r) {
    q->d.clear();
    r->d.clear();
    if (big_cmp(a, b) < 0) {
        *r = a;
        return;
    }
    const size_t n = b.d.size(), m = a.d.size() - n;
    if (n == 1) {
        uint64_t rem = 0;
        q->d.assign(a.d.size(), 0);
        for (size_t i = a.d.size(); i-- > 0;) {
            const uint64_t cur = (rem << 32) | a.d[i];
            q->d[i] = (uint32_t)(cur / b.d[0]);
            rem = cur % b.d[0];
        }
        q->trim();
        *r = Big(rem);
        return;
    }
    int s = 0;
    while (!((b.d[n - 1] << s) & 0x80000000u)) s++;
    std::vector<uint32_t> v(n), u(a.d.size() + 1);
    for (size_t i = n; i-- > 0;) v[i] = (b.d[i] << s) | (s && i ? (uint32_t)((uint64_t)b.d[i - 1] >> (32 - s)) : 0);
    u[a.d.size()] = s ? (uint32_t)((uint64_t)a.d.back() >> (32 - s)) : 0;
    for (size_t i = a.d.size(); i-- > 0;) u[i] = (a.d[i] << s) | (s && i ? (uint32_t)((uint64_t)a.d[i - 1] >> (32 - s)) : 0);
    q->d.assign(m + 1, 0);
    for (size_t j = m + 1; j-- > 0;) {
        const uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
        uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
        while (qhat >> 32 || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            qhat--;
            rhat += v[n - 1];
            if (rhat >> 32) break;
        }
        int64_t bw = 0;
        uint64_t carry = 0;
        for (size_t i = 0; i < n; i++) {
            const uint64_t p = qhat * v[i] + carry;
            carry = p >> 32;
            const int64_t t = (int64_t)u[i + j] - bw - (int64_t)(p & 0xFFFFFFFFu);
            u[i + j] = (uint32_t)t;
            bw = t < 0;
        }
        const int64_t t = (int64_t)u[j + n] - bw - (int64_t)carry;
        u[j + n] = (uint32_t)t;
        if (t < 0) {   // qhat was one too large: add the divisor back
            qhat--;
            uint64_t c = 0;
            for (size_t i = 0; i < n; i++) {
                c += (uint64_t)u[i + j] + v[i];
                u[i + j] = (uint32_t)c;
                c >>= 32;
            }
            u[j + n] += (uint32_t)c;
        }
        q->d[j] = (uint32_t)qhat;
    }
    q->trim();
    r->d.assign(n, 0);
    for (size_t i = 0; i < n; i++) r->d[i] = (u[i] >> s) | (s ? (uint32_t)((uint64_t)u[i + 1] << (32 - s)) : 0);
    r->trim();
}
inline Big big_mod(const Big& a, const Big& m) {
    Big q, r;
    big_divrem(a, m, &q, &r);
    return r;
}
inline Big big_mulmod(const Big& a, const Big& b, const Big& m) { return big_mod(big_mul(a, b), m); }
inline Big big_submod(const Big& a, const Big& b, const Big& m) {   // a, b < m
    return big_cmp(a, b) >= 0 ? big_sub(a, b) : big_sub(big_add(a, m), b);
}
inline Big big_powmod(const Big& x, const Big& e, const Big& m) {
    Big r(1), base = big_mod(x, m);
    for (size_t i = e.bits(); i-- > 0;) {
        r = big_mulmod(r, r, m);
        if (e.bit(i)) r = big_mulmod(r, base, m);
    }
    return r;
}

// x^-1 mod m for an odd modulus m of at most 16 limbs and gcd(x, m) = 1: binary extended Euclid on fixed-size limbs (shifts and
// subtractions only; ~15 x faster than x^(m-2) with a division per multiplication).  x is reduced first.
inline Big big_invmod_odd(const Big& x_in, const Big& m) {
    const int L = 17;   // one spare limb for (u + m) before the halving
    struct Fix {
        uint32_t w[L];
    };
    auto load = [&](const Big& v) {
        Fix f = {};
        for (size_t i = 0; i < v.d.size() && i < (size_t)L; i++) f.w[i] = v.d[i];
        return f;
    };
    auto is_zero = [&](const Fix& f) {
        for (int i = 0; i < L; i++)
            if (f.w[i]) return false;
        return true;
    };
    auto even = [](const Fix& f) { return (f.w[0] & 1) == 0; };
    auto shr1 = [&](Fix& f) {
        for (int i = 0; i < L - 1; i++) f.w[i] = (f.w[i] >> 1) | (f.w[i + 1] << 31);
        f.w[L - 1] >>= 1;
    };
    auto add = [&](Fix& f, const Fix& g) {
        uint64_t c = 0;
        for (int i = 0; i < L; i++) {
            c += (uint64_t)f.w[i] + g.w[i];
            f.w[i] = (uint32_t)c;
            c >>= 32;
        }
    };
    auto sub = [&](Fix& f, const Fix& g) {   // f >= g
        int64_t bw = 0;
        for (int i = 0; i < L; i++) {
            const int64_t t = (int64_t)f.w[i] - g.w[i] - bw;
            f.w[i] = (uint32_t)t;
            bw = t < 0;
        }
    };
    auto ge = [&](const Fix& f, const Fix& g) {
        for (int i = L; i-- > 0;)
            if (f.w[i] != g.w[i]) return f.w[i] > g.w[i];
        return true;
    };
    const Fix M = load(m);
    Fix a = load(big_mod(x_in, m)), b = M, u = load(Big(1)), v = {};
    auto halve_mod = [&](Fix& t) {   // t / 2 mod m
        if (!even(t)) add(t, M);
        shr1(t);
    };
    auto sub_mod = [&](Fix& t, const Fix& s) {   // t - s mod m, both < m
        if (!ge(t, s)) add(t, M);
        sub(t, s);
    };
    while (!is_zero(a)) {
        while (even(a)) {
            shr1(a);
            halve_mod(u);
        }
        while (even(b)) {
            shr1(b);
            halve_mod(v);
        }
        if (ge(a, b)) {
            sub(a, b);
            sub_mod(u, v);
        } else {
            sub(b, a);
            sub_mod(v, u);
        }
    }
    Big r;   // b = gcd = 1, v = x^-1
    r.d.assign(v.w, v.w + L);
    r.trim();
    return r;
}
