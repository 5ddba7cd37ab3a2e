// The reference's EcdsaSecp256k1 translator and the gadget stack under it, restated on the mini CircuitBuilder of p2acir.cpp
// (included from it, inside its namespace, after Builder and Translator):
//   plonky2-backend/src/circuit_translation/ecdsa_secp256k1_translator.rs:38-117   translate / _calculate_r / _32_bytes_to_field_element
//   plonky2-backend/src/plonky2_ecdsa/biguint/biguint.rs:82-262                   BigUintTarget arithmetic, div_rem, cmp
//   plonky2-backend/src/plonky2_ecdsa/biguint/gadgets/multiple_comparison.rs:15-64  list_le_circuit
//   plonky2-backend/src/plonky2_ecdsa/biguint/gadgets/nonnative.rs:135-447        non-native field arithmetic (+ its generators :470-680)
//   plonky2-backend/src/plonky2_ecdsa/biguint/gadgets/split_nonnative.rs:50-60    2-bit limbs
//   plonky2-backend/src/plonky2_ecdsa/curve/gadgets/curve.rs:98-215               affine secp256k1 points, incomplete formulas
//   plonky2-backend/src/plonky2_ecdsa/curve/gadgets/glv.rs:48-254, 339-383        GLV decomposition, 2-bit windowed double MSM
// The operation order follows the reference's so that gates are created in the same sequence.  Host tooling, not on the hot path.

const uint64_t SECP256K1_P64[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL};
const uint64_t SECP256K1_N64[4] = {0xBFD25E8CD0364141ULL, 0xBAAEDCE6AF48A03BULL, 0xFFFFFFFFFFFFFFFEULL, 0xFFFFFFFFFFFFFFFFULL};
const uint64_t SECP256K1_GX64[4] = {0x59F2815B16F81798ULL, 0x029BFCDB2DCE28D9ULL, 0x55A06295CE870B07ULL, 0x79BE667EF9DCBBACULL};   // curve.rs:17-22
const uint64_t SECP256K1_GY64[4] = {0x9C47D08FFB10D4B8ULL, 0xFD17B448A6855419ULL, 0x5DA4FBFC0E1108A8ULL, 0x483ADA7726A3C465ULL};   // curve.rs:25-30
const uint64_t GLV_BETA64[4] = {13923278643952681454ULL, 11308619431505398165ULL, 7954561588662645993ULL, 8856726876819556112ULL};   // glv.rs:23-28
const uint64_t GLV_S64[4] = {16069571880186789234ULL, 1310022930574435960ULL, 11900229862571533402ULL, 6008836872998760672ULL};    // glv.rs:30-35
const uint64_t GLV_A1_64[2] = {16747920425669159701ULL, 3496713202691238861ULL};                                                    // glv.rs:37
const uint64_t GLV_MINUS_B1_64[2] = {8022177200260244675ULL, 16448129721693014056ULL};                                              // glv.rs:39-40
const uint64_t GLV_A2_64[3] = {6323353552219852760ULL, 1498098850674701302ULL, 1ULL};                                               // glv.rs:42
const uint64_t GLV_B2_64[2] = {16747920425669159701ULL, 3496713202691238861ULL};                                                    // glv.rs:44

enum { FIELD_BASE = 0, FIELD_SCALAR = 1 };
inline Big field_order(int f) { return Big::from_u64_limbs(f == FIELD_BASE ? SECP256K1_P64 : SECP256K1_N64, 4); }

// glv.rs:46-91: Algorithm 15.41 of the Handbook of Elliptic and Hyperelliptic Curve Cryptography; (|k1|, |k2|, k1 < 0, k2 < 0)
inline void glv_decompose(const Big& k, Big* k1, Big* k2, bool* k1_neg, bool* k2_neg) {
    const Big n = field_order(FIELD_SCALAR), a1 = Big::from_u64_limbs(GLV_A1_64, 2), mb1 = Big::from_u64_limbs(GLV_MINUS_B1_64, 2),
              a2 = Big::from_u64_limbs(GLV_A2_64, 3), b2 = Big::from_u64_limbs(GLV_B2_64, 2);
    auto round_div = [&](const Big& num) {   // Ratio::round: half away from zero = floor((2 num + n) / (2 n)) for num >= 0
        Big q, r;
        big_divrem(big_add(big_shl(num, 1), n), big_shl(n, 1), &q, &r);
        return q;
    };
    const Big c1 = big_mod(round_div(big_mul(b2, k)), n), c2 = big_mod(round_div(big_mul(mb1, k)), n);
    const Big k1_raw = big_submod(big_submod(big_mod(k, n), big_mulmod(c1, a1, n), n), big_mulmod(c2, a2, n), n);
    const Big k2_raw = big_submod(big_mulmod(c1, mb1, n), big_mulmod(c2, b2, n), n);
    Big half, rem;
    big_divrem(n, Big(2), &half, &rem);
    *k1_neg = big_cmp(k1_raw, half) > 0;
    *k1 = *k1_neg ? big_mod(big_sub(n, k1_raw), n) : k1_raw;
    *k2_neg = big_cmp(k2_raw, half) > 0;
    *k2 = *k2_neg ? big_mod(big_sub(n, k2_raw), n) : k2_raw;
}

// ---- generators over big integers -------------------------------------------------------------------------------------------------
inline bool Builder::run_big(BigGen& g) {
    auto get_big = [&](const BigT& t, Big* out) {   // WitnessBigUint::get_biguint_target: fold (acc << 32) + limb, most significant limb first
        Big acc;
        for (size_t i = t.size(); i-- > 0;) {
            u64 v;
            if (!get(t[i], &v)) return false;
            acc = big_add(big_shl(acc, 32), Big(v));
        }
        *out = acc;
        return true;
    };
    auto set_big = [&](const BigT& t, const Big& v) {   // GeneratedValuesBigUint::set_biguint_target
        if (t.size() < v.d.size()) throw Error{"witness generation: big integer does not fit its target"};
        for (size_t i = 0; i < t.size(); i++) set(t[i], v.limb(i));
    };
    const Big m = field_order(g.field);
    if (g.kind == BG_NN_ADD_MANY) {   // nonnative.rs:527-552; g.b holds the end offset of each summand inside g.a
        Big sum;
        size_t lo = 0;
        for (Target end : g.b) {
            Big v;
            if (!get_big(BigT(g.a.begin() + lo, g.a.begin() + end), &v)) return false;
            sum = big_add(sum, big_mod(v, m));
            lo = (size_t)end;
        }
        Big q, r;
        big_divrem(sum, m, &q, &r);
        set_big(g.o1, r);
        set(g.t1, q.limb(0));
        return true;
    }
    Big a, b;
    if (!get_big(g.a, &a) || !get_big(g.b, &b)) return false;
    switch (g.kind) {
    case BG_DIVREM: {   // biguint.rs:311-318
        if (b.is_zero()) throw Error{"witness generation: division by zero"};
        Big q, r;
        big_divrem(a, b, &q, &r);
        set_big(g.o1, q);
        set_big(g.o2, r);
        return true;
    }
    case BG_NN_ADD: {   // nonnative.rs:469-484 (the comparison is strict: a sum equal to the modulus is left as it is)
        Big sum = big_add(big_mod(a, m), big_mod(b, m));
        const bool overflow = big_cmp(sum, m) > 0;
        if (overflow) sum = big_sub(sum, m);
        set_big(g.o1, sum);
        set(g.t1, overflow);
        return true;
    }
    case BG_NN_SUB: {   // nonnative.rs:585-600
        const Big ar = big_mod(a, m), br = big_mod(b, m);
        const bool overflow = big_cmp(ar, br) < 0;
        set_big(g.o1, overflow ? big_sub(big_add(m, ar), br) : big_sub(ar, br));
        set(g.t1, overflow);
        return true;
    }
    case BG_NN_MUL: {   // nonnative.rs:641-654
        Big q, r;
        big_divrem(big_mul(big_mod(a, m), big_mod(b, m)), m, &q, &r);
        set_big(g.o1, r);
        set_big(g.o2, q);
        return true;
    }
    case BG_NN_INV: {   // nonnative.rs:689-703
        const Big x = big_mod(a, m);
        if (x.is_zero()) throw Error{"witness generation: inverse of zero (the reference's curve formulas are incomplete)"};
        const Big inv = big_invmod_odd(x, m);
        Big q, r;
        big_divrem(big_mul(x, inv), m, &q, &r);
        set_big(g.o2, q);
        set_big(g.o1, inv);
        return true;
    }
    case BG_NN_ADD_MANY: return false;   // handled above
    case BG_GLV: {   // glv.rs:272-285
        Big k1, k2;
        bool n1, n2;
        glv_decompose(big_mod(a, m), &k1, &k2, &n1, &n2);
        set_big(g.o1, k1);
        set_big(g.o2, k2);
        set(g.t1, n1);
        set(g.t2, n2);
        return true;
    }
    }
    return false;
}

struct AffinePoint {
    BigT x, y;   // NonNativeTarget<Secp256K1Base> each
};

struct Ecc {
    Builder& b;
    explicit Ecc(Builder& builder) : b(builder) {}

    void add_big_gen(const BigGen& g) {
        b.big_gens.push_back(g);
        Gen e = {};
        e.kind = GEN_BIG;
        e.i = (int)b.big_gens.size() - 1;
        b.push_gen(e);
    }
    BigT virtual_biguint(size_t n) {
        BigT t;
        for (size_t i = 0; i < n; i++) t.push_back(b.add_virtual_target());
        return t;
    }
    void assert_bool(Target t) { b.connect(b.mul_sub(t, t, t), b.zero()); }

    // ---- biguint.rs ------------------------------------------------------------------------------------------------------------
    BigT constant_biguint(const Big& v) {   // :82-87
        BigT t;
        for (u32 limb : v.d) t.push_back(b.constant(limb));
        return t;
    }
    void connect_biguint(const BigT& l, const BigT& r) {   // :93-105
        const size_t mn = std::min(l.size(), r.size());
        for (size_t i = 0; i < mn; i++) b.connect(l[i], r[i]);
        for (size_t i = mn; i < l.size(); i++) b.assert_zero(l[i]);
        for (size_t i = mn; i < r.size(); i++) b.assert_zero(r[i]);
    }
    Target list_le(const BigT& x, const BigT& y, int num_bits) {   // multiple_comparison.rs:15-64
        if (x.size() != y.size()) throw Error{"Comparison must be between same number of inputs and outputs"};
        const Target one = b.one();
        Target result = one;
        for (size_t i = 0; i < x.size(); i++) {
            const Target a_le_b = b.cmp_le(x[i], y[i], num_bits);
            const Target b_le_a = b.cmp_le(y[i], x[i], num_bits);
            const Target these_limbs_equal = b.mul(a_le_b, b_le_a);
            const Target these_limbs_less_than = b.sub(one, b_le_a);
            result = b.arithmetic(1, 1, these_limbs_equal, result, these_limbs_less_than);
        }
        return result;
    }
    Target cmp_biguint(BigT x, BigT y) {   // :131-135 with pad_biguints :107-129
        while (x.size() < y.size()) x.push_back(b.zero());
        while (y.size() < x.size()) y.push_back(b.zero());
        return list_le(x, y, 32);
    }
    BigT add_biguint(const BigT& x, const BigT& y) {   // :143-165
        const size_t n = std::max(x.size(), y.size());
        BigT out;
        Target carry = b.zero();
        for (size_t i = 0; i < n; i++) {
            const Target xl = i < x.size() ? x[i] : b.zero(), yl = i < y.size() ? y[i] : b.zero();
            auto r = b.add_many_u32({carry, xl, yl});
            carry = r.second;
            out.push_back(r.first);
        }
        out.push_back(carry);
        return out;
    }
    BigT sub_biguint(BigT x, BigT y) {   // :167-184: the first is assumed larger, the last borrow is dropped
        while (x.size() < y.size()) x.push_back(b.zero());
        while (y.size() < x.size()) y.push_back(b.zero());
        BigT out;
        Target borrow = b.zero();
        for (size_t i = 0; i < x.size(); i++) {
            auto r = b.sub_u32(x[i], y[i], borrow);
            out.push_back(r.first);
            borrow = r.second;
        }
        return out;
    }
    BigT mul_biguint(const BigT& x, const BigT& y) {   // :186-210
        const size_t total = x.size() + y.size();
        std::vector<std::vector<Target>> to_add(total);
        for (size_t i = 0; i < x.size(); i++)
            for (size_t j = 0; j < y.size(); j++) {
                auto r = b.mul_u32(x[i], y[j]);
                to_add[i + j].push_back(r.first);
                to_add[i + j + 1].push_back(r.second);
            }
        BigT out;
        Target carry = b.zero();
        for (auto& summands : to_add) {
            if (summands.empty()) throw Error{"mul_biguint: an operand has no limbs"};
            auto r = b.add_u32s_with_carry(summands, carry);
            out.push_back(r.first);
            carry = r.second;
        }
        out.push_back(carry);
        return out;
    }
    BigT mul_biguint_by_bool(const BigT& x, Target bit) {   // :212-223
        BigT out;
        for (Target l : x) out.push_back(b.mul(l, bit));
        return out;
    }
    std::pair<BigT, BigT> div_rem_biguint(const BigT& x, const BigT& y) {   // :236-262
        const size_t div_limbs = y.size() > x.size() + 1 ? 0 : x.size() - y.size() + 1;
        BigT div = virtual_biguint(div_limbs), rem = virtual_biguint(y.size());
        BigGen g = {};
        g.kind = BG_DIVREM;
        g.a = x;
        g.b = y;
        g.o1 = div;
        g.o2 = rem;
        add_big_gen(g);
        BigT div_b = mul_biguint(div, y);
        BigT div_b_plus_rem = add_biguint(div_b, rem);
        connect_biguint(x, div_b_plus_rem);
        b.connect(cmp_biguint(rem, y), b.one());
        return {div, rem};
    }

    // ---- nonnative.rs ----------------------------------------------------------------------------------------------------------
    BigT constant_nonnative(int f, const Big& v) { return constant_biguint(big_mod(v, field_order(f))); }   // :150-153
    BigT add_nonnative(int f, const BigT& x, const BigT& y) {   // :189-217
        BigT sum = virtual_biguint(8);
        const Target overflow = b.add_virtual_target();
        BigGen g = {};
        g.kind = BG_NN_ADD;
        g.field = f;
        g.a = x;
        g.b = y;
        g.o1 = sum;
        g.t1 = overflow;
        add_big_gen(g);
        BigT sum_expected = add_biguint(x, y);
        BigT modulus = constant_biguint(field_order(f));
        BigT mod_times_overflow = mul_biguint_by_bool(modulus, overflow);
        BigT sum_actual = add_biguint(sum, mod_times_overflow);
        connect_biguint(sum_expected, sum_actual);
        b.connect(cmp_biguint(sum, modulus), b.one());
        return sum;
    }
    BigT add_many_nonnative(int f, const std::vector<BigT>& to_add) {   // :250-284 (not used by the ECDSA translator)
        if (to_add.size() == 1) return to_add[0];
        BigT sum = virtual_biguint(8);
        const Target overflow = b.add_virtual_target();
        BigGen g = {};
        g.kind = BG_NN_ADD_MANY;
        g.field = f;
        for (const BigT& t : to_add) {   // the summands, concatenated; b records where each ends
            g.a.insert(g.a.end(), t.begin(), t.end());
            g.b.push_back((Target)g.a.size());
        }
        g.o1 = sum;
        g.t1 = overflow;
        add_big_gen(g);
        b.range_check_u32(sum);
        b.range_check_u32({overflow});
        BigT sum_expected = constant_biguint(Big());
        for (const BigT& t : to_add) sum_expected = add_biguint(sum_expected, t);
        BigT modulus = constant_biguint(field_order(f));
        BigT mod_times_overflow = mul_biguint(modulus, BigT{overflow});
        BigT sum_actual = add_biguint(sum, mod_times_overflow);
        connect_biguint(sum_expected, sum_actual);
        b.connect(cmp_biguint(sum, modulus), b.one());
        return sum;
    }
    BigT sub_nonnative(int f, const BigT& x, const BigT& y) {   // :286-312
        BigT diff = virtual_biguint(8);
        const Target overflow = b.add_virtual_target();
        BigGen g = {};
        g.kind = BG_NN_SUB;
        g.field = f;
        g.a = x;
        g.b = y;
        g.o1 = diff;
        g.t1 = overflow;
        add_big_gen(g);
        b.range_check_u32(diff);
        assert_bool(overflow);
        BigT diff_plus_b = add_biguint(diff, y);
        BigT modulus = constant_biguint(field_order(f));
        BigT mod_times_overflow = mul_biguint_by_bool(modulus, overflow);
        BigT reduced = sub_biguint(diff_plus_b, mod_times_overflow);
        connect_biguint(x, reduced);
        return diff;
    }
    BigT mul_nonnative(int f, const BigT& x, const BigT& y) {   // :314-344
        BigT prod = virtual_biguint(8);
        BigT modulus = constant_biguint(field_order(f));
        if (x.size() + y.size() < modulus.size()) throw Error{"mul_nonnative: operands shorter than the modulus"};
        BigT overflow = virtual_biguint(x.size() + y.size() - modulus.size());
        BigGen g = {};
        g.kind = BG_NN_MUL;
        g.field = f;
        g.a = x;
        g.b = y;
        g.o1 = prod;
        g.o2 = overflow;
        add_big_gen(g);
        b.range_check_u32(prod);
        b.range_check_u32(overflow);
        BigT prod_expected = mul_biguint(x, y);
        BigT mod_times_overflow = mul_biguint(modulus, overflow);
        BigT prod_actual = add_biguint(prod, mod_times_overflow);
        connect_biguint(prod_expected, prod_actual);
        return prod;
    }
    BigT neg_nonnative(int f, const BigT& x) { return sub_nonnative(f, constant_biguint(Big()), x); }   // :361-366
    BigT inv_nonnative(int f, const BigT& x) {   // :368-393
        BigT inv = virtual_biguint(x.size()), div = virtual_biguint(x.size());
        BigGen g = {};
        g.kind = BG_NN_INV;
        g.field = f;
        g.a = x;
        g.o1 = inv;
        g.o2 = div;
        add_big_gen(g);
        BigT product = mul_biguint(x, inv);
        BigT modulus = constant_biguint(field_order(f));
        BigT mod_times_div = mul_biguint(modulus, div);
        BigT one = constant_biguint(Big(1));
        BigT expected_product = add_biguint(mod_times_div, one);
        connect_biguint(product, expected_product);
        return inv;
    }
    BigT reduce(int f, const BigT& x) { return div_rem_biguint(x, constant_biguint(field_order(f))).second; }   // :395-405
    BigT nonnative_conditional_neg(int f, const BigT& x, Target bit) {   // :436-447
        const Target not_b = b.b_not(bit);
        BigT neg = neg_nonnative(f, x);
        BigT x_if_true = mul_biguint_by_bool(neg, bit);
        BigT x_if_false = mul_biguint_by_bool(x, not_b);
        return add_nonnative(f, x_if_true, x_if_false);
    }
    std::vector<Target> split_nonnative_to_2_bit_limbs(const BigT& x) {   // split_nonnative.rs:50-60
        std::vector<Target> out;
        for (Target l : x) {
            std::vector<Target> limbs = b.split_le_base4(l, 16);
            out.insert(out.end(), limbs.begin(), limbs.end());
        }
        return out;
    }

    // ---- curve.rs --------------------------------------------------------------------------------------------------------------
    AffinePoint constant_affine_point(const Big& x, const Big& y) { return {constant_nonnative(FIELD_BASE, x), constant_nonnative(FIELD_BASE, y)}; }
    AffinePoint curve_conditional_neg(const AffinePoint& p, Target bit) { return {p.x, nonnative_conditional_neg(FIELD_BASE, p.y, bit)}; }   // :149-154
    AffinePoint curve_double(const AffinePoint& p) {   // :156-179
        const int F = FIELD_BASE;
        BigT double_y = add_nonnative(F, p.y, p.y);
        BigT inv_double_y = inv_nonnative(F, double_y);
        BigT x_squared = mul_nonnative(F, p.x, p.x);
        BigT double_x_squared = add_nonnative(F, x_squared, x_squared);
        BigT triple_x_squared = add_nonnative(F, double_x_squared, x_squared);
        BigT a = constant_nonnative(F, Big());   // SECP256K1_A = 0
        BigT triple_xx_a = add_nonnative(F, triple_x_squared, a);
        BigT lambda = mul_nonnative(F, triple_xx_a, inv_double_y);
        BigT lambda_squared = mul_nonnative(F, lambda, lambda);
        BigT x_double = add_nonnative(F, p.x, p.x);
        BigT x3 = sub_nonnative(F, lambda_squared, x_double);
        BigT x_diff = sub_nonnative(F, p.x, x3);
        BigT lambda_x_diff = mul_nonnative(F, lambda, x_diff);
        BigT y3 = sub_nonnative(F, lambda_x_diff, p.y);
        return {x3, y3};
    }
    AffinePoint curve_repeated_double(AffinePoint p, int n) {   // :181-189
        for (int i = 0; i < n; i++) p = curve_double(p);
        return p;
    }
    AffinePoint curve_add(const AffinePoint& p1, const AffinePoint& p2) {   // :191-207: the points are assumed different
        const int F = FIELD_BASE;
        BigT u = sub_nonnative(F, p2.y, p1.y);
        BigT v = sub_nonnative(F, p2.x, p1.x);
        BigT v_inv = inv_nonnative(F, v);
        BigT s = mul_nonnative(F, u, v_inv);
        BigT s_squared = mul_nonnative(F, s, s);
        BigT x_sum = add_nonnative(F, p2.x, p1.x);
        BigT x3 = sub_nonnative(F, s_squared, x_sum);
        BigT x_diff = sub_nonnative(F, p1.x, x3);
        BigT prod = mul_nonnative(F, s, x_diff);
        BigT y3 = sub_nonnative(F, prod, p1.y);
        return {x3, y3};
    }
    AffinePoint curve_conditional_add(const AffinePoint& p1, const AffinePoint& p2, Target bit) {   // :209-227
        const int F = FIELD_BASE;
        const Target not_b = b.b_not(bit);
        AffinePoint sum = curve_add(p1, p2);
        BigT x_if_true = mul_biguint_by_bool(sum.x, bit);
        BigT y_if_true = mul_biguint_by_bool(sum.y, bit);
        BigT x_if_false = mul_biguint_by_bool(p1.x, not_b);
        BigT y_if_false = mul_biguint_by_bool(p1.y, not_b);
        BigT x = add_nonnative(F, x_if_true, x_if_false);
        BigT y = add_nonnative(F, y_if_true, y_if_false);
        return {x, y};
    }

    // ---- glv.rs ----------------------------------------------------------------------------------------------------------------
    AffinePoint random_access_curve_points(Target index, const std::vector<AffinePoint>& v) {   // :339-383
        const Target zero = b.zero();
        AffinePoint out;
        for (int coord = 0; coord < 2; coord++)
            for (size_t i = 0; i < 8; i++) {
                std::vector<Target> limbs;
                for (const AffinePoint& p : v) {
                    const BigT& c = coord == 0 ? p.x : p.y;
                    limbs.push_back(i < c.size() ? c[i] : zero);
                }
                (coord == 0 ? out.x : out.y).push_back(b.random_access(index, limbs));
            }
        return out;
    }
    AffinePoint curve_msm(const AffinePoint& p, const AffinePoint& q, const BigT& n, const BigT& m) {   // :168-254: n p + m q, 2-bit windows
        static const uint8_t RANDO_X[32] = {168, 108, 112, 254, 40, 235, 44, 180, 232, 129, 170, 129, 151, 26, 229, 18,
                                            19, 137, 245, 62, 139, 130, 119, 30, 84, 53, 9, 156, 170, 172, 160, 15};
        static const uint8_t RANDO_Y[32] = {60, 32, 167, 79, 44, 197, 157, 125, 248, 190, 148, 181, 142, 227, 95, 8,
                                            136, 133, 192, 43, 110, 22, 130, 29, 171, 221, 92, 43, 9, 1, 185, 27};
        static const uint8_t NEG_RANDO_Y[32] = {195, 223, 88, 176, 211, 58, 98, 130, 7, 65, 107, 74, 113, 28, 160, 247,
                                                119, 122, 63, 212, 145, 233, 125, 226, 84, 34, 163, 211, 246, 254, 67, 20};
        static const uint8_t TO_ADD_X[32] = {4, 240, 116, 128, 2, 142, 26, 67, 121, 228, 15, 172, 125, 56, 178, 55,
                                             220, 178, 31, 194, 90, 168, 40, 127, 59, 193, 0, 121, 236, 178, 130, 29};
        static const uint8_t TO_ADD_Y[32] = {195, 20, 74, 65, 215, 167, 153, 201, 235, 110, 231, 40, 207, 121, 30, 55,
                                             18, 16, 205, 138, 169, 66, 20, 253, 49, 54, 35, 152, 247, 117, 246, 155};
        std::vector<Target> limbs_n = split_nonnative_to_2_bit_limbs(n), limbs_m = split_nonnative_to_2_bit_limbs(m);
        if (limbs_n.size() != limbs_m.size()) throw Error{"curve_msm: scalars of different sizes"};
        const AffinePoint rando = constant_affine_point(Big::from_be_bytes(RANDO_X, 32), Big::from_be_bytes(RANDO_Y, 32));
        const AffinePoint neg_rando = constant_affine_point(Big::from_be_bytes(RANDO_X, 32), Big::from_be_bytes(NEG_RANDO_Y, 32));
        // precomputation[i + 4 j] = i p + j q
        std::vector<AffinePoint> pre(16, p);
        AffinePoint cur_p = rando, cur_q = rando;
        for (int i = 0; i < 4; i++) {
            pre[i] = cur_p;
            pre[4 * i] = cur_q;
            cur_p = curve_add(cur_p, p);
            cur_q = curve_add(cur_q, q);
        }
        for (int i = 1; i < 4; i++) {
            pre[i] = curve_add(pre[i], neg_rando);
            pre[4 * i] = curve_add(pre[4 * i], neg_rando);
        }
        for (int i = 1; i < 4; i++)
            for (int j = 1; j < 4; j++) pre[i + 4 * j] = curve_add(pre[i], pre[4 * j]);
        const Target four = b.constant(4), zero = b.zero();
        AffinePoint result = rando;
        for (size_t k = limbs_n.size(); k-- > 0;) {
            result = curve_repeated_double(result, 2);
            const Target index = b.arithmetic(1, 1, four, limbs_m[k], limbs_n[k]);
            AffinePoint r = random_access_curve_points(index, pre);
            const Target is_zero = b.is_equal(index, zero);
            const Target should_add = b.b_not(is_zero);
            result = curve_conditional_add(result, r, should_add);
        }
        const AffinePoint to_add = constant_affine_point(Big::from_be_bytes(TO_ADD_X, 32), Big::from_be_bytes(TO_ADD_Y, 32));
        return curve_add(result, to_add);
    }
    AffinePoint glv_mul(const AffinePoint& p, const BigT& k) {   // :120-165
        const int S = FIELD_SCALAR;
        // decompose_secp256k1_scalar
        BigT k1 = virtual_biguint(4), k2 = virtual_biguint(4);
        const Target k1_neg = b.add_virtual_target(), k2_neg = b.add_virtual_target();
        BigGen g = {};
        g.kind = BG_GLV;
        g.field = S;
        g.a = k;
        g.o1 = k1;
        g.o2 = k2;
        g.t1 = k1_neg;
        g.t2 = k2_neg;
        add_big_gen(g);
        BigT k1_raw = nonnative_conditional_neg(S, k1, k1_neg);
        BigT k2_raw = nonnative_conditional_neg(S, k2, k2_neg);
        BigT s = constant_nonnative(S, Big::from_u64_limbs(GLV_S64, 4));
        BigT should_be_k = mul_nonnative(S, s, k2_raw);
        should_be_k = add_nonnative(S, should_be_k, k1_raw);
        connect_biguint(should_be_k, k);
        // glv_mul
        BigT beta = constant_nonnative(FIELD_BASE, Big::from_u64_limbs(GLV_BETA64, 4));
        BigT beta_px = mul_nonnative(FIELD_BASE, beta, p.x);
        AffinePoint sp = {beta_px, p.y};
        AffinePoint p_neg = curve_conditional_neg(p, k1_neg);
        AffinePoint sp_neg = curve_conditional_neg(sp, k2_neg);
        return curve_msm(p_neg, sp_neg, k1, k2);
    }
};

// ecdsa_secp256k1_translator.rs:90-117: 32 byte witnesses -> 8 u32 limbs (byte 0 is the least significant byte of limb 0, which is
// the least significant limb: the reference reads the 32 bytes as a little-endian integer) -> a non-native field element
inline BigT bytes32_to_field_element(Translator& T, const u32* byte_witnesses) {
    Builder& b = T.b;
    std::vector<Target> bytes;
    for (int i = 0; i < 32; i++) bytes.push_back(T.target_for_witness(byte_witnesses[i]));
    BigT limbs;
    for (int l = 0; l < 8; l++) {
        const Target t0 = bytes[4 * l];
        const Target t1 = b.mul_const(1ULL << 8, bytes[4 * l + 1]);
        const Target t2 = b.mul_const(1ULL << 16, bytes[4 * l + 2]);
        const Target t3 = b.mul_const(1ULL << 24, bytes[4 * l + 3]);
        Target acc = b.zero();   // add_many: fold from zero
        for (Target t : {t0, t1, t2, t3}) acc = b.add(acc, t);
        limbs.push_back(acc);
    }
    return limbs;
}

// ecdsa_secp256k1_translator.rs:38-88: output = (r <= x(h / s * G + r / s * PK)) -- the reference compares with cmp_biguint
inline void ecdsa_secp256k1(Translator& T, const u32* public_key_x, const u32* public_key_y, const u32* signature, const u32* hashed_msg,
                            u32 output) {
    Ecc e(T.b);
    // witness-generation groups (p2a_witness): the two scalar multiplications are independent of each other
    const int g_pre = T.b.new_group(), g_mul1 = T.b.new_group({g_pre}), g_mul2 = T.b.new_group({g_pre}),
              g_post = T.b.new_group({g_mul1, g_mul2});
    T.b.cur_group = g_pre;
    AffinePoint public_key;
    public_key.x = bytes32_to_field_element(T, public_key_x);
    public_key.y = bytes32_to_field_element(T, public_key_y);
    BigT r = bytes32_to_field_element(T, signature);
    BigT s = bytes32_to_field_element(T, signature + 32);
    BigT h = bytes32_to_field_element(T, hashed_msg);
    // _calculate_r
    BigT s1 = e.inv_nonnative(FIELD_SCALAR, s);
    BigT u1 = e.mul_nonnative(FIELD_SCALAR, h, s1);
    BigT u2 = e.mul_nonnative(FIELD_SCALAR, r, s1);
    AffinePoint generator = e.constant_affine_point(Big::from_u64_limbs(SECP256K1_GX64, 4), Big::from_u64_limbs(SECP256K1_GY64, 4));
    T.b.cur_group = g_mul1;
    AffinePoint r_factor_1 = e.glv_mul(generator, u1);
    T.b.cur_group = g_mul2;
    AffinePoint r_factor_2 = e.glv_mul(public_key, u2);
    T.b.cur_group = g_post;
    AffinePoint r_point = e.curve_add(r_factor_1, r_factor_2);
    const Target does_signature_verify = e.cmp_biguint(r, r_point.x);
    T.b.connect(does_signature_verify, T.target_for_witness(output));
    T.b.cur_group = 0;
}
