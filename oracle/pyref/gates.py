"""Constraint evaluators of the 13 gates reachable from the reference's translators, evaluated
over F_{p^2} (E2).  Base-field evaluation = lift to E2 and read c0.  ORACLE = test infrastructure.

Sources followed:
  * custom gates, read from /root/reference/plonky2-backend/src/plonky2_ecdsa/biguint/gates/:
      arithmetic_u32.rs:289-348, add_many_u32.rs:151-192, subtraction_u32.rs:234-271, range_check_u32.rs:95-117,
      comparison.rs:337-415
  * plonky2 0.2.2 built-ins (gates/{noop,constant,public_input,arithmetic_base,base_sum,poseidon,random_access}.rs),
    restated from SURVEY.md App. B (the crate is not vendored in the reference).
Gate universe: plonky2-backend/src/actions/write_vk_action.rs:37-61.
"""
from .field import E2, P
from .poseidon_constants import ALL_ROUND_CONSTANTS, MDS_CIRC, MDS_DIAG

NOOP, CONSTANT, PUBLIC_INPUT, ARITHMETIC, BASE_SUM, POSEIDON, RANDOM_ACCESS, U32_ARITHMETIC, U32_ADD_MANY, \
    U32_SUBTRACTION, U32_RANGE_CHECK, COMPARISON = range(12)

KIND_NAMES = ["NoopGate", "ConstantGate", "PublicInputGate", "ArithmeticGate", "BaseSumGate", "PoseidonGate",
              "RandomAccessGate", "U32ArithmeticGate", "U32AddManyGate", "U32SubtractionGate", "U32RangeCheckGate",
              "ComparisonGate"]


def ceil_div(a, b):
    return (a + b - 1) // b


class Gate:
    """kind + up to 4 integer parameters (the payload BackendGateSerializer writes for the gate)."""

    def __init__(self, kind, *params):
        self.kind = kind
        self.params = list(params) + [0] * (4 - len(params))

    # ---- constructors mirroring the reference / plonky2 `new_from_config`
    @staticmethod
    def noop():
        return Gate(NOOP)

    @staticmethod
    def constant(num_consts=2):
        return Gate(CONSTANT, num_consts)

    @staticmethod
    def public_input():
        return Gate(PUBLIC_INPUT)

    @staticmethod
    def arithmetic(num_routed):
        return Gate(ARITHMETIC, num_routed // 4)

    @staticmethod
    def base_sum(base, num_limbs):
        return Gate(BASE_SUM, base, num_limbs)

    @staticmethod
    def poseidon():
        return Gate(POSEIDON)

    @staticmethod
    def random_access(bits, num_wires, num_routed, num_constants=2):
        vec = 1 << bits
        copies = min(num_routed // (2 + vec), num_wires // (2 + vec + bits))
        extra = min(num_routed - (2 + vec) * copies, num_constants)
        return Gate(RANDOM_ACCESS, bits, copies, extra)

    @staticmethod
    def u32_arithmetic(num_wires, num_routed):
        # arithmetic_u32.rs:40-43
        return Gate(U32_ARITHMETIC, min(num_wires // 38, num_routed // 6))

    @staticmethod
    def u32_add_many(num_addends, num_wires, num_routed):
        # add_many_u32.rs:43-48
        return Gate(U32_ADD_MANY, num_addends, min(num_wires // (num_addends + 3 + 18), num_routed // (num_addends + 3)))

    @staticmethod
    def u32_subtraction(num_wires, num_routed):
        # subtraction_u32.rs:38-42
        return Gate(U32_SUBTRACTION, min(num_wires // 21, num_routed // 5))

    @staticmethod
    def u32_range_check(num_input_limbs):
        return Gate(U32_RANGE_CHECK, num_input_limbs)

    @staticmethod
    def comparison(num_bits=32, num_chunks=16):
        return Gate(COMPARISON, num_bits, num_chunks)

    # ---- static properties
    @property
    def degree(self):
        k, p = self.kind, self.params
        if k == NOOP:
            return 0
        if k in (CONSTANT, PUBLIC_INPUT):
            return 1
        if k == ARITHMETIC:
            return 3
        if k == BASE_SUM:
            return p[0]
        if k == POSEIDON:
            return 7
        if k == RANDOM_ACCESS:
            return p[0] + 1
        if k in (U32_ARITHMETIC, U32_ADD_MANY, U32_SUBTRACTION, U32_RANGE_CHECK):
            return 4
        if k == COMPARISON:
            return 1 << ceil_div(p[0], p[1])
        raise ValueError(k)

    @property
    def num_constraints(self):
        k, p = self.kind, self.params
        if k == NOOP:
            return 0
        if k == CONSTANT:
            return p[0]
        if k == PUBLIC_INPUT:
            return 4
        if k == ARITHMETIC:
            return p[0]
        if k == BASE_SUM:
            return 1 + p[1]
        if k == POSEIDON:
            return 12 * 7 + 22 + 12 + 1 + 4
        if k == RANDOM_ACCESS:
            return (p[0] + 2) * p[1] + p[2]
        if k == U32_ARITHMETIC:
            return p[0] * 36
        if k == U32_ADD_MANY:
            return p[1] * 21
        if k == U32_SUBTRACTION:
            return p[0] * 19
        if k == U32_RANGE_CHECK:
            return p[0] * 17
        if k == COMPARISON:
            return 6 + 5 * p[1] + ceil_div(p[0], p[1])
        raise ValueError(k)

    @property
    def num_constants(self):
        k, p = self.kind, self.params
        if k == CONSTANT:
            return p[0]
        if k == ARITHMETIC:
            return 2
        if k == RANDOM_ACCESS:
            return p[2]
        return 0

    @property
    def id(self):
        """Approximation of plonky2's `Gate::id()` Debug string; only its ORDER within equal degree matters."""
        k, p = self.kind, self.params
        n = KIND_NAMES[k]
        if k == CONSTANT:
            return f"{n} {{ num_consts: {p[0]} }}"
        if k == ARITHMETIC:
            return f"{n} {{ num_ops: {p[0]} }}"
        if k == BASE_SUM:
            return f"{n} {{ num_limbs: {p[1]} }} + Base: {p[0]}"
        if k == POSEIDON:
            return "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>"
        if k == RANDOM_ACCESS:
            return f"{n} {{ bits: {p[0]}, num_copies: {p[1]}, num_extra_constants: {p[2]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}<D=2>"
        if k == U32_ARITHMETIC:
            return f"{n} {{ num_ops: {p[0]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}"
        if k == U32_ADD_MANY:
            return f"{n} {{ num_addends: {p[0]}, num_ops: {p[1]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}"
        if k == U32_SUBTRACTION:
            return f"{n} {{ num_ops: {p[0]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}"
        if k == U32_RANGE_CHECK:
            return f"{n} {{ num_input_limbs: {p[0]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}"
        if k == COMPARISON:
            return f"{n} {{ num_bits: {p[0]}, num_chunks: {p[1]}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}<D=2>"
        return n

    def sort_key(self):
        return (self.degree, self.id)

    def __repr__(self):
        return f"Gate({KIND_NAMES[self.kind]}, {self.params})"

    # ---- constraint evaluation
    def eval_unfiltered(self, c, w, pi_hash):
        """c: gate-local constants (selector prefix removed), w: wires, pi_hash: 4 values.  Elements are E2 (or int).
        Returns the list of num_constraints E2 values, in plonky2's yield order."""
        k, p = self.kind, self.params
        L = E2.lift
        c = [L(x) for x in c]
        w = [L(x) for x in w]
        out = []
        if k == NOOP:
            pass
        elif k == CONSTANT:
            for i in range(p[0]):
                out.append(c[i] - w[i])
        elif k == PUBLIC_INPUT:
            for i in range(4):
                out.append(w[i] - L(pi_hash[i]))
        elif k == ARITHMETIC:
            for i in range(p[0]):
                m0, m1, ad, o = w[4 * i:4 * i + 4]
                out.append(o - (m0 * m1 * c[0] + ad * c[1]))
        elif k == BASE_SUM:
            B, nl = p[0], p[1]
            limbs = w[1:1 + nl]
            acc = E2(0)
            for l in reversed(limbs):
                acc = acc * B + l
            out.append(acc - w[0])
            for l in limbs:
                pr = E2(1)
                for v in range(B):
                    pr = pr * (l - v)
                out.append(pr)
        elif k == POSEIDON:
            out = _poseidon_gate(w)
        elif k == RANDOM_ACCESS:
            bits, copies, extra = p[0], p[1], p[2]
            vec = 1 << bits
            routed_used = (2 + vec) * copies + extra
            for cp in range(copies):
                base = (2 + vec) * cp
                idx, claimed = w[base], w[base + 1]
                items = w[base + 2:base + 2 + vec]
                bs = [w[routed_used + cp * bits + i] for i in range(bits)]
                for b in bs:
                    out.append(b * (b - 1))
                rec = E2(0)
                for b in reversed(bs):
                    rec = rec + rec + b
                out.append(rec - idx)
                for b in bs:
                    items = [items[2 * j] + b * (items[2 * j + 1] - items[2 * j]) for j in range(len(items) // 2)]
                out.append(items[0] - claimed)
            for i in range(extra):
                out.append(c[i] - w[(2 + vec) * copies + i])
        elif k == U32_ARITHMETIC:
            ops = p[0]
            for i in range(ops):
                m0, m1, ad, lo, hi, inv_ = w[6 * i:6 * i + 6]
                computed = m0 * m1 + ad
                diff = E2((1 << 32) - 1) - hi
                hi_not_max = inv_ * diff - 1
                out.append(hi_not_max * lo)
                out.append(hi * (1 << 32) + lo - computed)
                cl, ch = E2(0), E2(0)
                for j in reversed(range(32)):
                    limb = w[6 * ops + 32 * i + j]
                    out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                    if j < 16:
                        cl = cl * 4 + limb
                    else:
                        ch = ch * 4 + limb
                out.append(cl - lo)
                out.append(ch - hi)
        elif k == U32_ADD_MANY:
            na, ops = p[0], p[1]
            for i in range(ops):
                b0 = (na + 3) * i
                computed = E2(0)
                for j in range(na):
                    computed = computed + w[b0 + j]
                computed = computed + w[b0 + na]
                res, carry = w[b0 + na + 1], w[b0 + na + 2]
                out.append(carry * (1 << 32) + res - computed)
                cr, cc = E2(0), E2(0)
                for j in reversed(range(18)):
                    limb = w[(na + 3) * ops + 18 * i + j]
                    out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                    if j < 16:
                        cr = cr * 4 + limb
                    else:
                        cc = cc * 4 + limb
                out.append(cr - res)
                out.append(cc - carry)
        elif k == U32_SUBTRACTION:
            ops = p[0]
            for i in range(ops):
                x, y, bor, res, bout = w[5 * i:5 * i + 5]
                init = x - y - bor
                out.append(res - (init + bout * (1 << 32)))
                comb = E2(0)
                for j in reversed(range(16)):
                    limb = w[5 * ops + 16 * i + j]
                    out.append(limb * (limb - 1) * (limb - 2) * (limb - 3))
                    comb = comb * 4 + limb
                out.append(comb - res)
                out.append(bout * (1 - bout))
        elif k == U32_RANGE_CHECK:
            n = p[0]
            for i in range(n):
                aux = [w[n + 16 * i + j] for j in range(16)]
                acc = E2(0)
                for a in reversed(aux):
                    acc = acc * 4 + a
                out.append(acc - w[i])
                for a in aux:
                    out.append(a * (a - 1) * (a - 2) * (a - 3))
        elif k == COMPARISON:
            nb, nc = p[0], p[1]
            cb = ceil_div(nb, nc)
            cs = 1 << cb
            first, second = w[0], w[1]
            fc = [w[4 + i] for i in range(nc)]
            sc = [w[4 + nc + i] for i in range(nc)]

            def rwp(xs, b):
                acc = E2(0)
                for x in reversed(xs):
                    acc = acc * b + x
                return acc
            out.append(rwp(fc, cs) - first)
            out.append(rwp(sc, cs) - second)
            msd = E2(0)
            for i in range(nc):
                fp_, sp_ = E2(1), E2(1)
                for x in range(cs):
                    fp_ = fp_ * (fc[i] - x)
                    sp_ = sp_ * (sc[i] - x)
                out.append(fp_)
                out.append(sp_)
                diff = sc[i] - fc[i]
                eq_dummy = w[4 + 2 * nc + i]
                ch_eq = w[4 + 3 * nc + i]
                out.append(diff * eq_dummy - (1 - ch_eq))
                out.append(ch_eq * diff)
                inter = w[4 + 4 * nc + i]
                out.append(inter - ch_eq * msd)
                msd = inter + (1 - ch_eq) * diff
            msd_w = w[3]
            out.append(msd_w - msd)
            bits = [w[4 + 5 * nc + i] for i in range(cb + 1)]
            for b in bits:
                out.append(b * (1 - b))
            out.append(msd_w + cs - rwp(bits, 2))
            out.append(w[2] - bits[cb])
        else:
            raise ValueError(k)
        assert len(out) == self.num_constraints, (self, len(out))
        return out


def _mds_e2(s):
    return [sum((s[(i + r) % 12] * MDS_CIRC[i] for i in range(12)), E2(0)) + s[r] * MDS_DIAG[r] for r in range(12)]


def _sbox(x):
    x2 = x * x
    x4 = x2 * x2
    return x4 * x2 * x


def _poseidon_gate(w):
    """PoseidonGate constraints in the naive round form (polynomial-identical to plonky2's fast partial rounds;
    SURVEY App. B).  Wire layout: in 0..12, out 12..24, swap 24, delta 25..29, full0 29..65, partial 65..87, full1 87..135."""
    out = []
    swap = w[24]
    out.append(swap * (swap - 1))
    for i in range(4):
        out.append(swap * (w[i + 4] - w[i]) - w[25 + i])
    st = [None] * 12
    for i in range(4):
        st[i] = w[i] + w[25 + i]
        st[i + 4] = w[i + 4] - w[25 + i]
    for i in range(8, 12):
        st[i] = w[i]
    rnd = 0
    for r in range(4):
        st = [st[i] + ALL_ROUND_CONSTANTS[12 * rnd + i] for i in range(12)]
        if r != 0:
            for i in range(12):
                sb = w[29 + 12 * (r - 1) + i]
                out.append(st[i] - sb)
                st[i] = sb
        st = _mds_e2([_sbox(x) for x in st])
        rnd += 1
    for r in range(22):
        st = [st[i] + ALL_ROUND_CONSTANTS[12 * rnd + i] for i in range(12)]
        sb = w[65 + r]
        out.append(st[0] - sb)
        st[0] = _sbox(sb)
        st = _mds_e2(st)
        rnd += 1
    for r in range(4):
        st = [st[i] + ALL_ROUND_CONSTANTS[12 * rnd + i] for i in range(12)]
        for i in range(12):
            sb = w[87 + 12 * r + i]
            out.append(st[i] - sb)
            st[i] = sb
        st = _mds_e2([_sbox(x) for x in st])
        rnd += 1
    for i in range(12):
        out.append(st[i] - w[12 + i])
    return out
