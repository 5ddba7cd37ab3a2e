"""Proof data model, byte codecs (uncompressed + compressed) and Merkle-path (de)compression.

Restates plonky2 0.2.2 util/serialization.rs (write_proof / write_compressed_proof), plonk/proof.rs, fri/proof.rs
(FriProof::compress / decompress) and hash/path_compression.rs, per SURVEY.md App. A.12.  ORACLE = test infrastructure.
Reference call sites: plonky2-backend/src/actions/prove_action.rs:75-78 (compress + to_bytes),
plonky2-backend/src/noir_and_plonky2_serialization.rs:24-33 (CompressedProofWithPublicInputs::from_bytes).
"""
import struct

from .field import E2, P
from .hashing import HASHERS, hash_or_noop


class OpeningSet:
    FIELDS = ["constants", "sigmas", "wires", "zs", "zs_next", "partial_products", "quotient"]

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw.get(f, []))

    def zeta_batch(self):
        # to_fri_openings(): constants, sigmas, wires, zs, partial_products, quotient
        return self.constants + self.sigmas + self.wires + self.zs + self.partial_products + self.quotient

    def zeta_next_batch(self):
        return list(self.zs_next)


class QueryRound:
    def __init__(self, initial, steps):
        self.initial = initial  # list over 4 oracles of (values [int], path [bytes])
        self.steps = steps      # list over layers of (evals [E2], path [bytes])


class Proof:
    def __init__(self):
        self.wires_cap = []
        self.zs_pp_cap = []
        self.quotient_cap = []
        self.openings = None
        self.fri_caps = []
        self.query_rounds = []
        self.final_poly = []
        self.pow_witness = 0
        self.public_inputs = []


class CompressedProof(Proof):
    def __init__(self):
        super().__init__()
        self.indices = []
        self.initial_by_index = {}   # index -> list over oracles of (values, path)
        self.steps_by_index = []     # per layer: index -> (evals minus inferred, path)


class _Reader:
    def __init__(self, data):
        self.d = data
        self.o = 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise ValueError("proof truncated")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b

    def u8(self):
        return self.take(1)[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def u64(self):
        v = struct.unpack("<Q", self.take(8))[0]
        if v >= P:
            raise ValueError("non-canonical field element")
        return v

    def ext(self):
        return E2(self.u64(), self.u64())

    def remaining(self):
        return len(self.d) - self.o


def _w64(x):
    return struct.pack("<Q", x)


def _wext(e):
    return struct.pack("<QQ", e.c0, e.c1)


def _hs(cd):
    return HASHERS[cd.hasher].hash_size


def _read_cap(r, cd):
    return [r.take(_hs(cd)) for _ in range(1 << cd.cap_height)]


def _read_path(r, cd):
    n = r.u8()
    return [r.take(_hs(cd)) for _ in range(n)]


def _write_path(path):
    return bytes([len(path)]) + b"".join(path)


def _read_openings(r, cd):
    def ev(n):
        return [r.ext() for _ in range(n)]
    os_ = OpeningSet()
    os_.constants = ev(cd.num_constants)
    os_.sigmas = ev(cd.num_routed)
    os_.wires = ev(cd.num_wires)
    os_.zs = ev(cd.num_challenges)
    os_.zs_next = ev(cd.num_challenges)
    # lookup_zs / next_lookup_zs: empty (no lookups reachable from the translators)
    os_.partial_products = ev(cd.num_challenges * cd.num_partial_products)
    os_.quotient = ev(cd.num_quotient)
    return os_


def _write_openings(os_):
    out = b""
    for f in ["constants", "sigmas", "wires", "zs", "zs_next", "partial_products", "quotient"]:
        out += b"".join(_wext(e) for e in getattr(os_, f))
    return out


def _read_head(r, cd, pr):
    pr.wires_cap = _read_cap(r, cd)
    pr.zs_pp_cap = _read_cap(r, cd)
    pr.quotient_cap = _read_cap(r, cd)
    pr.openings = _read_openings(r, cd)
    pr.fri_caps = [_read_cap(r, cd) for _ in cd.arity_bits]


def _write_head(pr):
    out = b"".join(pr.wires_cap) + b"".join(pr.zs_pp_cap) + b"".join(pr.quotient_cap)
    out += _write_openings(pr.openings)
    for cap in pr.fri_caps:
        out += b"".join(cap)
    return out


def _read_initial(r, cd):
    res = []
    for width in cd.oracle_widths():
        vals = [r.u64() for _ in range(width)]
        res.append((vals, _read_path(r, cd)))
    return res


def _write_initial(ini):
    out = b""
    for vals, path in ini:
        out += b"".join(_w64(v) for v in vals) + _write_path(path)
    return out


def _read_tail(r, cd, pr):
    pr.final_poly = [r.ext() for _ in range(cd.final_poly_len)]
    pr.pow_witness = r.u64()
    pr.public_inputs = [r.u64() for _ in range(cd.num_public_inputs)]
    if r.remaining() != 0:
        raise ValueError(f"{r.remaining()} trailing bytes")


def _write_tail(pr):
    return b"".join(_wext(e) for e in pr.final_poly) + _w64(pr.pow_witness) + b"".join(_w64(v) for v in pr.public_inputs)


def parse_uncompressed(data, cd):
    r = _Reader(data)
    pr = Proof()
    _read_head(r, cd, pr)
    for _ in range(cd.num_queries):
        ini = _read_initial(r, cd)
        steps = []
        for ab in cd.arity_bits:
            evals = [r.ext() for _ in range(1 << ab)]
            steps.append((evals, _read_path(r, cd)))
        pr.query_rounds.append(QueryRound(ini, steps))
    _read_tail(r, cd, pr)
    return pr


def serialize_uncompressed(pr):
    out = _write_head(pr)
    for qr in pr.query_rounds:
        out += _write_initial(qr.initial)
        for evals, path in qr.steps:
            out += b"".join(_wext(e) for e in evals) + _write_path(path)
    return out + _write_tail(pr)


def parse_compressed(data, cd):
    r = _Reader(data)
    pr = CompressedProof()
    _read_head(r, cd, pr)
    pr.indices = [r.u32() for _ in range(cd.num_queries)]
    for idx in sorted(set(pr.indices)):
        pr.initial_by_index[idx] = _read_initial(r, cd)
    cur = list(pr.indices)
    for ab in cd.arity_bits:
        cur = [i >> ab for i in cur]
        d = {}
        for idx in sorted(set(cur)):
            evals = [r.ext() for _ in range((1 << ab) - 1)]
            d[idx] = (evals, _read_path(r, cd))
        pr.steps_by_index.append(d)
    _read_tail(r, cd, pr)
    return pr


def serialize_compressed(pr):
    out = _write_head(pr)
    out += b"".join(struct.pack("<I", i) for i in pr.indices)
    for idx in sorted(pr.initial_by_index):
        out += _write_initial(pr.initial_by_index[idx])
    for d in pr.steps_by_index:
        for idx in sorted(d):
            evals, path = d[idx]
            out += b"".join(_wext(e) for e in evals) + _write_path(path)
    return out + _write_tail(pr)


# ------------------------------------------------------------------ Merkle path compression (hash/path_compression.rs)
def compress_merkle_proofs(cap_height, indices, proofs):
    height = cap_height + len(proofs[0])
    num_leaves = 1 << height
    known = [False] * (2 * num_leaves)
    for i in indices:
        for j in range(height - cap_height):
            known[(i + num_leaves) >> j] = True
    out = []
    for i, p in zip(indices, proofs):
        cp = []
        index = i + num_leaves
        for sib in p:
            si = index ^ 1
            if not known[si]:
                cp.append(sib)
                known[si] = True
            index >>= 1
        out.append(cp)
    return out


def decompress_merkle_proofs(H, leaves_data, indices, cproofs, height, cap_height):
    num_leaves = 1 << height
    seen = {}
    for i, v in zip(indices, leaves_data):
        seen[i + num_leaves] = hash_or_noop(H, v)
    iters = [iter(p) for p in cproofs]
    for layer in range(height - cap_height):
        for i, it in zip(indices, iters):
            index = (i + num_leaves) >> layer
            cur = seen[index]
            si = index ^ 1
            if si not in seen:
                seen[si] = next(it)
            sib = seen[si]
            seen[index >> 1] = H.two_to_one(cur, sib) if index % 2 == 0 else H.two_to_one(sib, cur)
    out = []
    for i in indices:
        index = i + num_leaves
        p = []
        for _ in range(height - cap_height):
            p.append(seen[index ^ 1])
            index >>= 1
        out.append(p)
    return out


def flatten_ext(evals):
    out = []
    for e in evals:
        out += [e.c0, e.c1]
    return out


def compress_proof(pr, indices, cd):
    """ProofWithPublicInputs::compress given the FS query indices (fri/proof.rs FriProof::compress)."""
    cp = CompressedProof()
    for f in ["wires_cap", "zs_pp_cap", "quotient_cap", "openings", "fri_caps", "final_poly", "pow_witness", "public_inputs"]:
        setattr(cp, f, getattr(pr, f))
    cp.indices = list(indices)
    n_or = len(pr.query_rounds[0].initial)
    nl = len(cd.arity_bits)
    ini_idx = [[] for _ in range(n_or)]
    ini_proofs = [[] for _ in range(n_or)]
    st_idx = [[] for _ in range(nl)]
    st_evals = [[] for _ in range(nl)]
    st_proofs = [[] for _ in range(nl)]
    for index, qr in zip(indices, pr.query_rounds):
        for i, (vals, path) in enumerate(qr.initial):
            ini_idx[i].append(index)
            ini_proofs[i].append(path)
        for i, (evals, path) in enumerate(qr.steps):
            within = index & ((1 << cd.arity_bits[i]) - 1)
            index >>= cd.arity_bits[i]
            st_idx[i].append(index)
            ev = list(evals)
            del ev[within]
            st_evals[i].append(ev)
            st_proofs[i].append(path)
    ini_c = [compress_merkle_proofs(cd.cap_height, a, b) for a, b in zip(ini_idx, ini_proofs)]
    st_c = [compress_merkle_proofs(cd.cap_height, a, b) for a, b in zip(st_idx, st_proofs)]
    cp.steps_by_index = [dict() for _ in range(nl)]
    for i, index in enumerate(indices):
        ini = [(pr.query_rounds[i].initial[j][0], ini_c[j][i]) for j in range(n_or)]
        cp.initial_by_index.setdefault(index, ini)
        for j in range(nl):
            index >>= cd.arity_bits[j]
            cp.steps_by_index[j].setdefault(index, (st_evals[j][i], st_c[j][i]))
    return cp


def decompress_proof(cp, cd, inferred_fn):
    """CompressedFriProof::decompress.  inferred_fn(query_no, x_index, initial_values, layer_evals_so_far) is supplied by
    the verifier (it owns the challenges): returns, per layer, the element to re-insert."""
    H = HASHERS[cd.hasher]
    pr = Proof()
    for f in ["wires_cap", "zs_pp_cap", "quotient_cap", "openings", "fri_caps", "final_poly", "pow_witness", "public_inputs"]:
        setattr(pr, f, getattr(cp, f))
    n_or = 4
    nl = len(cd.arity_bits)
    ini_idx = [[] for _ in range(n_or)]
    ini_leaves = [[] for _ in range(n_or)]
    ini_proofs = [[] for _ in range(n_or)]
    st_idx = [[] for _ in range(nl)]
    st_evals = [[] for _ in range(nl)]
    st_proofs = [[] for _ in range(nl)]
    for qn, index in enumerate(cp.indices):
        ini = cp.initial_by_index[index]
        for j, (vals, path) in enumerate(ini):
            ini_idx[j].append(index)
            ini_leaves[j].append(vals)
            ini_proofs[j].append(path)
        x = index
        layer_evals = []
        for j in range(nl):
            within = x & ((1 << cd.arity_bits[j]) - 1)
            x >>= cd.arity_bits[j]
            ev, path = cp.steps_by_index[j][x]
            ev = list(ev)
            ev.insert(within, inferred_fn(qn, index, [v for v, _ in ini], layer_evals, j))
            layer_evals.append(ev)
            st_idx[j].append(x)
            st_evals[j].append(ev)
            st_proofs[j].append(path)
    height = cd.lde_bits
    ini_d = [decompress_merkle_proofs(H, ini_leaves[j], ini_idx[j], ini_proofs[j], height, cd.cap_height) for j in range(n_or)]
    st_d = []
    h = height
    for j in range(nl):
        h -= cd.arity_bits[j]
        st_d.append(decompress_merkle_proofs(H, [flatten_ext(e) for e in st_evals[j]], st_idx[j], st_proofs[j], h, cd.cap_height))
    for qn in range(len(cp.indices)):
        ini = [(ini_leaves[j][qn], ini_d[j][qn]) for j in range(n_or)]
        steps = [(st_evals[j][qn], st_d[j][qn]) for j in range(nl)]
        pr.query_rounds.append(QueryRound(ini, steps))
    return pr
