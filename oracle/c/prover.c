/* ORACLE = TEST INFRASTRUCTURE (also the CPU baseline, kind "port").  Not linked by the product.
 *
 * CPU restatement (C + OpenMP) of plonky2 0.2.2 `prove_with_partition_witness`, the function behind the reference's
 *     circuit_data.prove(witnesses).unwrap()      /root/reference/plonky2-backend/src/actions/prove_action.rs:96
 * following SURVEY.md App. A.3-A.10 step by step (plonk/prover.rs, fri/oracle.rs, fri/prover.rs, hash/merkle_tree.rs,
 * plonk/vanishing_poly.rs, util/serialization.rs of the un-vendored crate).  Parallelism mirrors the reference's Rayon
 * use: per column (FFTs), per leaf (hashing), per LDE point (quotient), per polynomial (openings).
 *
 * Pinned by: regenerating both golden proofs of the reference byte-for-byte (tests/test_oracle_golden.py).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "gates.h"

#define ORC_EXPORT __attribute__((visibility("default")))

typedef struct {
    int log_leaves, hs, nlevels; /* levels[0] = leaf digests ... levels[nlevels-1] = cap */
    u8** levels;
    int ncols;
    u64* leaves; /* row-major [leaf][col] */
} mtree;

typedef struct {
    int ncols, log_n, rate_bits;
    u64* coeffs; /* col-major [col][N] */
    mtree tree;  /* leaves = LDE rows in bit-reversed order */
} batch;

typedef struct orc_ctx {
    p2g_circuit_desc d;
    p2g_gate* gates;
    u64* k_is;
    int n, lde, hs, h;
    u64* cs_values; /* preprocessed values [P][N] */
    batch cs, wires, zs_pp, quot;
    u8 digest[ORC_MAX_HS];
    u64* roots[33];
    /* dumps of the last prove */
    u64* zs_pp_values;
    u64 challenges[64 + 64];
    int nchallenges;
    e2* final_poly;
    int final_len;
    u8* fri_caps;
    char err[256];
} orc_ctx;

static u64* get_roots(orc_ctx* c, int logn) { /* table of omega^k, k < n/2 */
    if (c->roots[logn]) return c->roots[logn];
    size_t half = logn ? ((size_t)1 << (logn - 1)) : 1;
    u64* t = (u64*)malloc(half * 8);
    u64 w = gl_root_of_unity(logn), x = 1;
    for (size_t i = 0; i < half; i++) { t[i] = x; x = gl_mul(x, w); }
    c->roots[logn] = t;
    return t;
}

/* natural-order in, natural-order out radix-2 DIT (plonky2_field fft.rs fft_classic semantics) */
static void fft_inplace(orc_ctx* c, u64* a, int logn) {
    size_t n = (size_t)1 << logn;
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev64(i, logn);
        if (i < j) { u64 t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    const u64* roots = get_roots(c, logn);
    for (int s = 0; s < logn; s++) {
        size_t m = (size_t)1 << s, stride = n >> (s + 1);
        for (size_t k = 0; k < n; k += 2 * m)
            for (size_t j = 0; j < m; j++) {
                u64 u = a[k + j], v = gl_mul(a[k + j + m], roots[j * stride]);
                a[k + j] = gl_add(u, v);
                a[k + j + m] = gl_sub(u, v);
            }
    }
}
static void ifft_inplace(orc_ctx* c, u64* a, int logn) {
    size_t n = (size_t)1 << logn;
    fft_inplace(c, a, logn);
    u64 ninv = gl_inv(n % GL_P);
    a[0] = gl_mul(a[0], ninv);
    if (n > 1) a[n / 2] = gl_mul(a[n / 2], ninv);
    for (size_t i = 1; i < n / 2; i++) {
        u64 x = gl_mul(a[i], ninv), y = gl_mul(a[n - i], ninv);
        a[i] = y;
        a[n - i] = x;
    }
}
static void coset_scale(u64* a, size_t n, u64 shift) {
    u64 s = 1;
    for (size_t i = 0; i < n; i++) { a[i] = gl_mul(a[i], s); s = gl_mul(s, shift); }
}

/* ---------------- Merkle ---------------- */
static void mtree_free(mtree* t) {
    if (t->levels) {
        for (int i = 0; i < t->nlevels; i++) free(t->levels[i]);
        free(t->levels);
    }
    free(t->leaves);
    memset(t, 0, sizeof *t);
}
/* takes ownership of leaves (row-major) */
static void mtree_build(mtree* t, int h, u64* leaves, int log_leaves, int ncols, int cap_height) {
    t->log_leaves = log_leaves;
    t->hs = hasher_size(h);
    t->ncols = ncols;
    t->leaves = leaves;
    if (cap_height > log_leaves) cap_height = log_leaves;
    t->nlevels = log_leaves - cap_height + 1;
    t->levels = (u8**)calloc(t->nlevels, sizeof(u8*));
    size_t nl = (size_t)1 << log_leaves;
    int hs = t->hs;
    t->levels[0] = (u8*)malloc(nl * hs);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nl; i++) hash_or_noop(h, leaves + (size_t)i * ncols, ncols, t->levels[0] + (size_t)i * hs);
    for (int k = 1; k < t->nlevels; k++) {
        size_t cnt = nl >> k;
        t->levels[k] = (u8*)malloc(cnt * hs);
        const u8* prev = t->levels[k - 1];
        u8* cur = t->levels[k];
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)cnt; i++) two_to_one(h, prev + (size_t)(2 * i) * hs, prev + (size_t)(2 * i + 1) * hs, cur + (size_t)i * hs);
    }
}
static const u8* mtree_cap(const mtree* t) { return t->levels[t->nlevels - 1]; }
static int mtree_ncap(const mtree* t) { return 1 << (t->log_leaves - (t->nlevels - 1)); }

/* ---------------- polynomial batch: PolynomialBatch::from_coeffs (fri/oracle.rs) ---------------- */
static void batch_free(batch* b) {
    free(b->coeffs);
    mtree_free(&b->tree);
    memset(b, 0, sizeof *b);
}
/* takes ownership of coeffs [ncols][N] */
static void batch_from_coeffs(orc_ctx* c, batch* b, u64* coeffs, int ncols, int log_n) {
    int rate_bits = c->d.rate_bits;
    int log_lde = log_n + rate_bits;
    size_t n = (size_t)1 << log_n, lde = (size_t)1 << log_lde;
    b->ncols = ncols;
    b->log_n = log_n;
    b->rate_bits = rate_bits;
    b->coeffs = coeffs;
    u64* leaves = (u64*)malloc(lde * ncols * 8);
    get_roots(c, log_lde); /* build the table outside the parallel region */
#pragma omp parallel
    {
        u64* tmp = (u64*)malloc(lde * 8);
#pragma omp for schedule(dynamic, 1)
        for (int col = 0; col < ncols; col++) {
            memcpy(tmp, coeffs + (size_t)col * n, n * 8);
            memset(tmp + n, 0, (lde - n) * 8);
            coset_scale(tmp, n, GL_GEN);
            fft_inplace(c, tmp, log_lde);
            /* transpose + reverse_index_bits_in_place: leaf j = natural index bitrev(j) */
            for (size_t i = 0; i < lde; i++) leaves[bitrev64(i, log_lde) * ncols + col] = tmp[i];
        }
        free(tmp);
    }
    mtree_build(&b->tree, c->h, leaves, log_lde, ncols, c->d.cap_height);
}
/* takes ownership of values [ncols][N] (converted in place to coefficients) */
static void batch_from_values(orc_ctx* c, batch* b, u64* values, int ncols, int log_n) {
    size_t n = (size_t)1 << log_n;
    get_roots(c, log_n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int col = 0; col < ncols; col++) ifft_inplace(c, values + (size_t)col * n, log_n);
    batch_from_coeffs(c, b, values, ncols, log_n);
}

/* ---------------- API ---------------- */
ORC_EXPORT const char* orc_last_error(orc_ctx* c) { return c ? c->err : "null ctx"; }

ORC_EXPORT void orc_destroy(orc_ctx* c) {
    if (!c) return;
    batch_free(&c->cs);
    batch_free(&c->wires);
    batch_free(&c->zs_pp);
    batch_free(&c->quot);
    free(c->cs_values);
    free(c->gates);
    free(c->k_is);
    free(c->zs_pp_values);
    free(c->final_poly);
    free(c->fri_caps);
    for (int i = 0; i < 33; i++) free(c->roots[i]);
    free(c);
}

static void compute_circuit_digest(orc_ctx* c) {
    /* H.hash_no_pad(cap.flatten() || H.hash_pad([]).to_vec() || [degree_bits])    */
    int ncap = mtree_ncap(&c->cs.tree);
    u64* parts = (u64*)malloc((4 * ncap + 5) * 8);
    int k = 0;
    for (int i = 0; i < ncap; i++) { hash_to_elems(c->h, mtree_cap(&c->cs.tree) + (size_t)i * c->hs, parts + k); k += 4; }
    u64 pad[8] = {1, 0, 0, 0, 0, 0, 0, 1};
    u8 dsep[ORC_MAX_HS];
    hash_no_pad(c->h, pad, 8, dsep);
    hash_to_elems(c->h, dsep, parts + k);
    k += 4;
    parts[k++] = c->d.degree_bits;
    hash_no_pad(c->h, parts, k, c->digest);
    free(parts);
}

ORC_EXPORT int orc_create(const p2g_circuit_desc* desc, orc_ctx** out) {
    if (!desc || !out || desc->struct_size != sizeof(p2g_circuit_desc)) return P2G_EBADARG;
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof *c);
    c->d = *desc;
    c->n = 1 << desc->degree_bits;
    c->lde = c->n << desc->rate_bits;
    c->h = desc->hasher;
    c->hs = hasher_size(c->h);
    c->gates = (p2g_gate*)malloc(sizeof(p2g_gate) * desc->num_gates);
    memcpy(c->gates, desc->gates, sizeof(p2g_gate) * desc->num_gates);
    c->k_is = (u64*)malloc(8 * desc->num_routed_wires);
    memcpy(c->k_is, desc->k_is, 8 * desc->num_routed_wires);
    int P = desc->num_constants + desc->num_routed_wires;
    size_t sz = (size_t)P * c->n * 8;
    c->cs_values = (u64*)malloc(sz);
    memcpy(c->cs_values, desc->constants_sigmas, sz);
    u64* tmp = (u64*)malloc(sz);
    memcpy(tmp, desc->constants_sigmas, sz);
    batch_from_values(c, &c->cs, tmp, P, desc->degree_bits);
    if (desc->circuit_digest) memcpy(c->digest, desc->circuit_digest, c->hs);
    else compute_circuit_digest(c);
    c->d.gates = c->gates;
    c->d.k_is = c->k_is;
    c->d.constants_sigmas = c->cs_values;
    c->d.circuit_digest = NULL;
    *out = c;
    return P2G_OK;
}

ORC_EXPORT int orc_cap(orc_ctx* c, u8* cap_out, u8* digest_out) {
    if (cap_out) memcpy(cap_out, mtree_cap(&c->cs.tree), (size_t)mtree_ncap(&c->cs.tree) * c->hs);
    if (digest_out) memcpy(digest_out, c->digest, c->hs);
    return P2G_OK;
}

/* filter_g(x) = prod_{j in group, j != g} (j - s) * [many selectors: (UNUSED - s)]   (gates/gate.rs compute_filter) */
static u64 compute_filter(const p2g_gate* g, u32 row, u64 s, int many) {
    u64 r = 1;
    for (u32 i = g->group_lo; i < g->group_hi; i++)
        if (i != row) r = gl_mul(r, gl_sub(i, s));
    if (many) r = gl_mul(r, gl_sub(0xFFFFFFFFULL, s));
    return r;
}

/* evaluate_gate_constraints_base_batch for one point */
static void eval_all_gates(const orc_ctx* c, const u64* consts, const u64* wires, const u64 pi[4], u64* acc, u64* tmp) {
    int ngc = c->d.num_gate_constraints;
    memset(acc, 0, 8 * ngc);
    for (u32 gi = 0; gi < c->d.num_gates; gi++) {
        const p2g_gate* g = &c->gates[gi];
        u64 f = compute_filter(g, gi, consts[g->selector_index], c->d.num_selectors > 1);
        int k = eval_gate_unfiltered(g, consts + c->d.num_selectors, wires, pi, tmp);
        for (int i = 0; i < k; i++) acc[i] = gl_add(acc[i], gl_mul(tmp[i], f));
    }
}

ORC_EXPORT int orc_eval_gate_constraints(const p2g_circuit_desc* desc, const u64* constants, const u64* wires, const u64* pi_hash,
                                         size_t npoints, u64* out) {
    orc_ctx c;
    memset(&c, 0, sizeof c);
    c.d = *desc;
    c.gates = (p2g_gate*)desc->gates;
    int ngc = desc->num_gate_constraints, nw = desc->num_wires, nc = desc->num_constants;
#pragma omp parallel
    {
        u64* acc = (u64*)malloc(8 * (ngc + 1));
        u64* tmp = (u64*)malloc(8 * (ngc + 1));
        u64* cw = (u64*)malloc(8 * (nc + nw));
#pragma omp for
        for (long i = 0; i < (long)npoints; i++) {
            for (int j = 0; j < nc; j++) cw[j] = constants[(size_t)j * npoints + i];
            for (int j = 0; j < nw; j++) cw[nc + j] = wires[(size_t)j * npoints + i];
            eval_all_gates(&c, cw, cw + nc, pi_hash, acc, tmp);
            for (int k = 0; k < ngc; k++) out[(size_t)k * npoints + i] = acc[k];
        }
        free(acc); free(tmp); free(cw);
    }
    return P2G_OK;
}

/* ---------------- byte writer ---------------- */
typedef struct { u8* p; size_t len, cap; } wbuf;
static void wb_put(wbuf* w, const void* src, size_t n) {
    if (w->len + n <= w->cap) memcpy(w->p + w->len, src, n);
    w->len += n;
}
static void wb_u64(wbuf* w, u64 x) { wb_put(w, &x, 8); }
static void wb_e2(wbuf* w, e2 x) { wb_u64(w, x.c0); wb_u64(w, x.c1); }
static void wb_u8(wbuf* w, u8 x) { wb_put(w, &x, 1); }
static void wb_merkle_path(wbuf* w, const mtree* t, size_t leaf) {
    int len = t->nlevels - 1;
    wb_u8(w, (u8)len);
    for (int k = 0; k < len; k++) wb_put(w, t->levels[k] + ((leaf >> k) ^ 1) * t->hs, t->hs);
}

static e2 eval_base_poly_e2(const u64* coeffs, size_t n, e2 z) {
    e2 acc = {0, 0};
    for (size_t i = n; i-- > 0;) acc = e2_add_base(e2_mul(acc, z), coeffs[i]);
    return acc;
}

/* ---------------- the prover ---------------- */
/* ORC_TRACE=1: wall-clock seconds per stage on stderr (used to see where the CPU baseline spends its time) */
static double orc_now(void) { return omp_get_wtime(); }
#define ORC_MARK(what)                                                                    \
    do {                                                                                  \
        if (orc_trace) {                                                                  \
            double t__ = orc_now();                                                       \
            fprintf(stderr, "[orc] %-28s %8.3f s\n", what, t__ - orc_last);                 \
            orc_last = t__;                                                               \
        }                                                                                 \
    } while (0)

ORC_EXPORT int orc_prove(orc_ctx* c, const u64* wires_in, const u64* public_inputs, size_t n_pi, const u64* forced_pow,
                         u8* out, size_t* out_len) {
    const int orc_trace = getenv("ORC_TRACE") != NULL;
    double orc_last = orc_now();

    const p2g_circuit_desc* d = &c->d;
    if (n_pi != d->num_public_inputs) { snprintf(c->err, sizeof c->err, "public input count"); return P2G_EBADARG; }
    const int n = c->n, lde = c->lde, logn = d->degree_bits, loglde = logn + d->rate_bits;
    const int W = d->num_wires, R = d->num_routed_wires, C = d->num_constants, NC = d->num_challenges;
    const int NPP = d->num_partial_products, QDF = d->quotient_degree_factor, NGC = d->num_gate_constraints;
    const int h = c->h, hs = c->hs;
    if (NC > 4 || QDF != (1 << d->rate_bits)) { snprintf(c->err, sizeof c->err, "unsupported config"); return P2G_EBADARG; }
    batch_free(&c->wires); batch_free(&c->zs_pp); batch_free(&c->quot);
    int nch = 0;

    /* 1. public inputs hash (InnerHasher = Poseidon, always) */
    u64 pi_hash[4] = {0, 0, 0, 0};
    if (n_pi) poseidon_hash_no_pad(public_inputs, n_pi, pi_hash);

    ORC_MARK("pi hash");
    /* 2. wires commitment */
    u64* wv = (u64*)malloc((size_t)W * n * 8);
    memcpy(wv, wires_in, (size_t)W * n * 8);
    batch_from_values(c, &c->wires, wv, W, logn);

    ORC_MARK("wires commit");
    /* 3-4. challenger */
    challenger ch;
    ch_init(&ch, h);
    ch_observe_hash(&ch, c->digest);
    ch_observe_many(&ch, pi_hash, 4);
    ch_observe_cap(&ch, mtree_cap(&c->wires.tree), mtree_ncap(&c->wires.tree));
    u64 betas[4], gammas[4], alphas[4];
    for (int i = 0; i < NC; i++) betas[i] = ch_get(&ch);
    for (int i = 0; i < NC; i++) gammas[i] = ch_get(&ch);

    ORC_MARK("challenger");
    /* 5. Z and partial products (A.7) */
    int nzp = NC * (1 + NPP);
    u64* zp = (u64*)malloc((size_t)nzp * n * 8); /* [Z_0..Z_{NC-1}, PP_0[0..NPP), PP_1[0..NPP)] x N */
    {
        u64 wn = gl_root_of_unity(logn);
        u64* xs = (u64*)malloc(8 * (size_t)n);
        xs[0] = 1;
        for (int i = 1; i < n; i++) xs[i] = gl_mul(xs[i - 1], wn);
        int nchunk = NPP + 1;
        u64* chunkprod = (u64*)malloc((size_t)n * nchunk * 8);
        for (int cc = 0; cc < NC; cc++) {
            u64 beta = betas[cc], gamma = gammas[cc];
#pragma omp parallel
            {
                u64* num = (u64*)malloc(8 * R);
                u64* den = (u64*)malloc(8 * R);
                u64* pre = (u64*)malloc(8 * R);
#pragma omp for schedule(static)
                for (int i = 0; i < n; i++) {
                    for (int j = 0; j < R; j++) {
                        u64 wj = wires_in[(size_t)j * n + i];
                        num[j] = gl_add(gl_add(wj, gl_mul(beta, gl_mul(c->k_is[j], xs[i]))), gamma);
                        den[j] = gl_add(gl_add(wj, gl_mul(beta, c->cs_values[(size_t)(C + j) * n + i])), gamma);
                    }
                    /* batch inverse of den */
                    u64 acc = 1;
                    for (int j = 0; j < R; j++) { pre[j] = acc; acc = gl_mul(acc, den[j]); }
                    u64 ainv = gl_inv(acc);
                    for (int j = R - 1; j >= 0; j--) { u64 dinv = gl_mul(ainv, pre[j]); ainv = gl_mul(ainv, den[j]); den[j] = dinv; }
                    for (int m = 0; m < nchunk; m++) {
                        u64 pr = 1;
                        for (int j = m * QDF; j < (m + 1) * QDF && j < R; j++) pr = gl_mul(pr, gl_mul(num[j], den[j]));
                        chunkprod[(size_t)i * nchunk + m] = pr;
                    }
                }
                free(num); free(den); free(pre);
            }
            u64 z = 1;
            for (int i = 0; i < n; i++) {
                zp[(size_t)cc * n + i] = z;
                u64 acc = z;
                for (int m = 0; m < nchunk; m++) {
                    acc = gl_mul(acc, chunkprod[(size_t)i * nchunk + m]);
                    if (m < NPP) zp[(size_t)(NC + cc * NPP + m) * n + i] = acc;
                }
                z = acc;
            }
        }
        free(xs); free(chunkprod);
    }
    free(c->zs_pp_values);
    c->zs_pp_values = (u64*)malloc((size_t)nzp * n * 8);
    memcpy(c->zs_pp_values, zp, (size_t)nzp * n * 8);
    batch_from_values(c, &c->zs_pp, zp, nzp, logn);
    ch_observe_cap(&ch, mtree_cap(&c->zs_pp.tree), mtree_ncap(&c->zs_pp.tree));

    ORC_MARK("z/pp + commit");
    /* 6. alphas */
    for (int i = 0; i < NC; i++) alphas[i] = ch_get(&ch);

    ORC_MARK("alphas");
    /* 7. quotient (A.8) */
    int nq = NC * QDF;
    u64* qv = (u64*)malloc((size_t)NC * lde * 8); /* natural order values per challenge */
    {
        u64 wl = gl_root_of_unity(loglde);
        u64 zh_inv[64], zh[64];
        u64 gn = gl_pow(GL_GEN, n);
        u64 wr = gl_root_of_unity(d->rate_bits);
        for (int i = 0; i < QDF; i++) { zh[i] = gl_sub(gl_mul(gn, gl_pow(wr, i)), 1); zh_inv[i] = gl_inv(zh[i]); }
        const u64* L_cs = c->cs.tree.leaves;
        const u64* L_w = c->wires.tree.leaves;
        const u64* L_z = c->zs_pp.tree.leaves;
        int Pn = C + R;
        int nterms = NC + NC * (NPP + 1) + NGC;
#pragma omp parallel
        {
            u64* terms = (u64*)malloc(8 * (nterms + 1));
            u64* tmp = (u64*)malloc(8 * (NGC + 1));
#pragma omp for schedule(static)
            for (int i = 0; i < lde; i++) {
                size_t row = bitrev64(i, loglde), row_next = bitrev64((i + QDF) % lde, loglde);
                const u64* cs = L_cs + row * Pn;
                const u64* w = L_w + row * W;
                const u64* zrow = L_z + row * nzp;
                const u64* znext = L_z + row_next * nzp;
                u64 x = gl_mul(GL_GEN, gl_pow(wl, i));
                u64 zhx = zh[i % QDF];
                u64 l0 = gl_mul(zhx, gl_inv(gl_mul(n, gl_sub(x, 1))));
                int t = 0;
                for (int cc = 0; cc < NC; cc++) terms[t++] = gl_mul(l0, gl_sub(zrow[cc], 1));
                for (int cc = 0; cc < NC; cc++) {
                    u64 prev = zrow[cc];
                    for (int m = 0; m <= NPP; m++) {
                        u64 pn = 1, pd = 1;
                        for (int j = m * QDF; j < (m + 1) * QDF && j < R; j++) {
                            u64 num = gl_add(gl_add(w[j], gl_mul(betas[cc], gl_mul(c->k_is[j], x))), gammas[cc]);
                            u64 den = gl_add(gl_add(w[j], gl_mul(betas[cc], cs[C + j])), gammas[cc]);
                            pn = gl_mul(pn, num);
                            pd = gl_mul(pd, den);
                        }
                        u64 next = (m < NPP) ? zrow[NC + cc * NPP + m] : znext[cc];
                        terms[t++] = gl_sub(gl_mul(prev, pn), gl_mul(next, pd));
                        prev = next;
                    }
                }
                eval_all_gates(c, cs, w, pi_hash, terms + t, tmp);
                for (int cc = 0; cc < NC; cc++) {
                    u64 acc = 0;
                    for (int k = nterms - 1; k >= 0; k--) acc = gl_add(gl_mul(acc, alphas[cc]), terms[k]);
                    qv[(size_t)cc * lde + i] = gl_mul(acc, zh_inv[i % QDF]);
                }
            }
            free(terms); free(tmp);
        }
    }
    u64* qc = (u64*)malloc((size_t)nq * n * 8);
    {
        u64 ginv = gl_inv(GL_GEN);
        get_roots(c, loglde);
        for (int cc = 0; cc < NC; cc++) {
            u64* v = qv + (size_t)cc * lde;
            ifft_inplace(c, v, loglde);
            coset_scale(v, lde, ginv);
            /* chunks of N: quotient_c(X) = sum_i chunk_i(X) X^{iN} */
            memcpy(qc + (size_t)cc * QDF * n, v, (size_t)lde * 8);
        }
    }
    free(qv);
    batch_from_coeffs(c, &c->quot, qc, nq, logn);
    ch_observe_cap(&ch, mtree_cap(&c->quot.tree), mtree_ncap(&c->quot.tree));

    ORC_MARK("quotient + commit");
    /* 8. zeta */
    e2 zeta = ch_get_e2(&ch);
    if (e2_eq(e2_pow(zeta, n), e2_make(1, 0))) { snprintf(c->err, sizeof c->err, "Opening point is in the subgroup."); return P2G_EUNSAT; }
    u64 g = gl_root_of_unity(logn);
    e2 zeta_next = e2_mul_base(zeta, g);

    ORC_MARK("zeta");
    /* 9. openings (A.9) */
    batch* oracles[4] = {&c->cs, &c->wires, &c->zs_pp, &c->quot};
    int widths[4] = {C + R, W, nzp, nq};
    int total = widths[0] + widths[1] + widths[2] + widths[3];
    e2* op = (e2*)malloc(sizeof(e2) * total); /* oracle order */
    {
        int off = 0;
        for (int o = 0; o < 4; o++) {
            batch* b = oracles[o];
#pragma omp parallel for schedule(dynamic, 1)
            for (int j = 0; j < widths[o]; j++) op[off + j] = eval_base_poly_e2(b->coeffs + (size_t)j * n, n, zeta);
            off += widths[o];
        }
    }
    e2 zs_next[4];
    for (int cc = 0; cc < NC; cc++) zs_next[cc] = eval_base_poly_e2(c->zs_pp.coeffs + (size_t)cc * n, n, zeta_next);
    /* observe: constants, sigmas, wires, zs, partial_products, quotient (= oracle order), then zs_next */
    for (int i = 0; i < total; i++) ch_observe_e2(&ch, op[i]);
    for (int cc = 0; cc < NC; cc++) ch_observe_e2(&ch, zs_next[cc]);

    ORC_MARK("openings");
    /* 10. FRI (A.10) */
    e2 fri_alpha = ch_get_e2(&ch);
    e2* fin = (e2*)calloc((size_t)lde, sizeof(e2)); /* final_poly coefficients, padded to lde */
    {
        /* batch 0: all polys at zeta */
        e2* comp = (e2*)calloc(n, sizeof(e2));
        /* composition = sum_j alpha^j poly_j  -> per coefficient Horner over polys in reverse */
#pragma omp parallel for schedule(static)
        for (int k = 0; k < n; k++) {
            e2 acc = {0, 0};
            for (int o = 3; o >= 0; o--)
                for (int j = widths[o] - 1; j >= 0; j--) acc = e2_add_base(e2_mul(acc, fri_alpha), oracles[o]->coeffs[(size_t)j * n + k]);
            comp[k] = acc;
        }
        /* divide_by_linear(zeta): synthetic division, quotient has n-1 coeffs, then push 0 */
        e2* q0 = (e2*)calloc(n, sizeof(e2));
        {
            e2 carry = {0, 0};
            for (int k = n - 1; k >= 1; k--) { carry = e2_add(comp[k], e2_mul(carry, zeta)); q0[k - 1] = carry; }
        }
        /* batch 1: Z polys at g*zeta */
#pragma omp parallel for schedule(static)
        for (int k = 0; k < n; k++) {
            e2 acc = {0, 0};
            for (int j = NC - 1; j >= 0; j--) acc = e2_add_base(e2_mul(acc, fri_alpha), c->zs_pp.coeffs[(size_t)j * n + k]);
            comp[k] = acc;
        }
        e2* q1 = (e2*)calloc(n, sizeof(e2));
        {
            e2 carry = {0, 0};
            for (int k = n - 1; k >= 1; k--) { carry = e2_add(comp[k], e2_mul(carry, zeta_next)); q1[k - 1] = carry; }
        }
        e2 a_pow = e2_pow(fri_alpha, NC);
        for (int k = 0; k < n; k++) fin[k] = e2_add(e2_mul(q0[k], a_pow), q1[k]);
        free(comp); free(q0); free(q1);
    }
    /* values = coset_fft(final.lde(rate_bits)) componentwise */
    int nl = d->num_fri_layers;
    e2* coeffs = fin;
    size_t cur = lde;
    int logcur = loglde;
    u64* re = (u64*)malloc(8 * (size_t)lde);
    u64* im = (u64*)malloc(8 * (size_t)lde);
    e2* values = (e2*)malloc(sizeof(e2) * (size_t)lde);
    u64 shift = GL_GEN;
#define COSET_FFT_E2()                                                              \
    do {                                                                            \
        for (size_t i_ = 0; i_ < cur; i_++) { re[i_] = coeffs[i_].c0; im[i_] = coeffs[i_].c1; } \
        coset_scale(re, cur, shift);                                                \
        coset_scale(im, cur, shift);                                                \
        get_roots(c, logcur);                                                       \
        fft_inplace(c, re, logcur);                                                 \
        fft_inplace(c, im, logcur);                                                 \
        for (size_t i_ = 0; i_ < cur; i_++) values[i_] = e2_make(re[i_], im[i_]);   \
    } while (0)
    COSET_FFT_E2();
    mtree* ftrees = (mtree*)calloc(nl ? nl : 1, sizeof(mtree));
    e2 fri_betas[P2G_MAX_FRI_LAYERS];
    free(c->fri_caps);
    c->fri_caps = (u8*)malloc((size_t)(nl ? nl : 1) * (1 << d->cap_height) * hs);
    for (int l = 0; l < nl; l++) {
        int ab = d->reduction_arity_bits[l], arity = 1 << ab;
        /* reverse_index_bits_in_place(values); leaves = chunks of arity, flattened */
        size_t nleaves = cur >> ab;
        u64* leaves = (u64*)malloc(cur * 16);
        for (size_t j = 0; j < cur; j++) {
            e2 v = values[bitrev64(j, logcur)];
            leaves[2 * j] = v.c0;
            leaves[2 * j + 1] = v.c1;
        }
        mtree_build(&ftrees[l], h, leaves, logcur - ab, 2 * arity, d->cap_height);
        ch_observe_cap(&ch, mtree_cap(&ftrees[l]), mtree_ncap(&ftrees[l]));
        memcpy(c->fri_caps + (size_t)l * (1 << d->cap_height) * hs, mtree_cap(&ftrees[l]), (size_t)mtree_ncap(&ftrees[l]) * hs);
        e2 beta = ch_get_e2(&ch);
        fri_betas[l] = beta;
        /* coeffs'[k] = sum_i coeffs[arity k + i] beta^i */
        for (size_t k = 0; k < nleaves; k++) {
            e2 acc = {0, 0};
            for (int i = arity - 1; i >= 0; i--) acc = e2_add(e2_mul(acc, beta), coeffs[(size_t)arity * k + i]);
            coeffs[k] = acc;
        }
        cur = nleaves;
        logcur -= ab;
        shift = gl_pow(shift, arity);
        COSET_FFT_E2();
    }
    size_t final_len = cur >> d->rate_bits;
    free(c->final_poly);
    c->final_poly = (e2*)malloc(sizeof(e2) * final_len);
    memcpy(c->final_poly, coeffs, sizeof(e2) * final_len);
    c->final_len = (int)final_len;
    for (size_t i = 0; i < final_len; i++) ch_observe_e2(&ch, coeffs[i]);

    /* PoW: smallest witness whose response has >= pow_bits leading zeros (reference: Rayon find_any) */
    u64 pow_witness = 0;
    {
        challenger base = ch;
        if (forced_pow) {
            pow_witness = *forced_pow;
        } else {
            int found = 0;
            u64 start = 0;
            const u64 chunk = 1 << 14;
            while (!found) {
                u64 best = ~0ULL;
#pragma omp parallel for schedule(static) reduction(min : best)
                for (long k = 0; k < (long)chunk; k++) {
                    challenger t = base;
                    ch_observe(&t, start + k);
                    u64 r = ch_get(&t);
                    if ((d->pow_bits == 0 || (r >> (64 - d->pow_bits)) == 0) && start + k < best) best = start + k;
                }
                if (best != ~0ULL) { pow_witness = best; found = 1; }
                start += chunk;
            }
        }
        ch_observe(&ch, pow_witness);
        u64 resp = ch_get(&ch);
        if (d->pow_bits && (resp >> (64 - d->pow_bits)) != 0) {
            snprintf(c->err, sizeof c->err, "forced pow_witness is invalid");
            return P2G_EUNSAT;
        }
    }
    int NQ = d->num_query_rounds;
    u64* indices = (u64*)malloc(8 * NQ);
    for (int q = 0; q < NQ; q++) indices[q] = ch_get(&ch) % (u64)lde;

    /* dump challenges */
    for (int i = 0; i < NC; i++) c->challenges[nch++] = betas[i];
    for (int i = 0; i < NC; i++) c->challenges[nch++] = gammas[i];
    for (int i = 0; i < NC; i++) c->challenges[nch++] = alphas[i];
    c->challenges[nch++] = zeta.c0; c->challenges[nch++] = zeta.c1;
    c->challenges[nch++] = fri_alpha.c0; c->challenges[nch++] = fri_alpha.c1;
    for (int l = 0; l < nl; l++) { c->challenges[nch++] = fri_betas[l].c0; c->challenges[nch++] = fri_betas[l].c1; }
    c->challenges[nch++] = pow_witness;
    for (int q = 0; q < NQ; q++) c->challenges[nch++] = indices[q];
    c->nchallenges = nch;

    /* serialise: Proof || public_inputs  (A.12, uncompressed) */
    wbuf wb = {out, 0, out ? *out_len : 0};
    int ncap = 1 << d->cap_height;
    wb_put(&wb, mtree_cap(&c->wires.tree), (size_t)ncap * hs);
    wb_put(&wb, mtree_cap(&c->zs_pp.tree), (size_t)ncap * hs);
    wb_put(&wb, mtree_cap(&c->quot.tree), (size_t)ncap * hs);
    {
        /* OpeningSet: constants | sigmas | wires | zs | zs_next | partial_products | quotient */
        int o_w = widths[0], o_z = o_w + W, o_q = o_z + nzp;
        for (int i = 0; i < o_w; i++) wb_e2(&wb, op[i]);
        for (int i = 0; i < W; i++) wb_e2(&wb, op[o_w + i]);
        for (int i = 0; i < NC; i++) wb_e2(&wb, op[o_z + i]);
        for (int i = 0; i < NC; i++) wb_e2(&wb, zs_next[i]);
        for (int i = NC; i < nzp; i++) wb_e2(&wb, op[o_z + i]);
        for (int i = 0; i < nq; i++) wb_e2(&wb, op[o_q + i]);
    }
    for (int l = 0; l < nl; l++) wb_put(&wb, mtree_cap(&ftrees[l]), (size_t)ncap * hs);
    for (int q = 0; q < NQ; q++) {
        size_t x = indices[q];
        for (int o = 0; o < 4; o++) {
            const mtree* t = &oracles[o]->tree;
            wb_put(&wb, t->leaves + x * t->ncols, (size_t)t->ncols * 8);
            wb_merkle_path(&wb, t, x);
        }
        for (int l = 0; l < nl; l++) {
            int ab = d->reduction_arity_bits[l];
            const mtree* t = &ftrees[l];
            size_t leaf = x >> ab;
            wb_put(&wb, t->leaves + leaf * t->ncols, (size_t)t->ncols * 8);
            wb_merkle_path(&wb, t, leaf);
            x = leaf;
        }
    }
    for (size_t i = 0; i < final_len; i++) wb_e2(&wb, c->final_poly[i]);
    wb_u64(&wb, pow_witness);
    for (size_t i = 0; i < n_pi; i++) wb_u64(&wb, public_inputs[i]);

    for (int l = 0; l < nl; l++) mtree_free(&ftrees[l]);
    free(ftrees); free(values); free(re); free(im); free(fin); free(op); free(indices);
    int rc = P2G_OK;
    if (!out || wb.len > wb.cap) rc = P2G_ESMALLBUF;
    *out_len = wb.len;
    return rc;
}

/* intermediates of the last prove (same numbering as enum p2g_buffer) */
ORC_EXPORT int orc_read(orc_ctx* c, int what, void* out, size_t* len) {
    const void* src = NULL;
    size_t sz = 0;
    size_t ncap = (size_t)1 << c->d.cap_height;
    size_t n = c->n;
    switch (what) {
    case P2G_BUF_WIRES_CAP: src = c->wires.tree.levels ? mtree_cap(&c->wires.tree) : NULL; sz = ncap * c->hs; break;
    case P2G_BUF_ZS_PP_CAP: src = c->zs_pp.tree.levels ? mtree_cap(&c->zs_pp.tree) : NULL; sz = ncap * c->hs; break;
    case P2G_BUF_QUOTIENT_CAP: src = c->quot.tree.levels ? mtree_cap(&c->quot.tree) : NULL; sz = ncap * c->hs; break;
    case P2G_BUF_CS_CAP: src = mtree_cap(&c->cs.tree); sz = ncap * c->hs; break;
    case P2G_BUF_ZS_PP_VALUES: src = c->zs_pp_values; sz = (size_t)c->d.num_challenges * (1 + c->d.num_partial_products) * n * 8; break;
    case P2G_BUF_QUOTIENT_CHUNKS: src = c->quot.coeffs; sz = (size_t)c->d.num_challenges * c->d.quotient_degree_factor * n * 8; break;
    case P2G_BUF_WIRES_COEFFS: src = c->wires.coeffs; sz = (size_t)c->d.num_wires * n * 8; break;
    case P2G_BUF_CHALLENGES: src = c->challenges; sz = (size_t)c->nchallenges * 8; break;
    case P2G_BUF_FINAL_POLY: src = c->final_poly; sz = (size_t)c->final_len * 16; break;
    case P2G_BUF_FRI_CAPS: src = c->fri_caps; sz = (size_t)c->d.num_fri_layers * ncap * c->hs; break;
    default: return P2G_EBADARG;
    }
    if (!src) return P2G_EBADARG;
    if (what == P2G_BUF_WIRES_LDE) return P2G_EBADARG;
    if (!out || *len < sz) { *len = sz; return P2G_ESMALLBUF; }
    memcpy(out, src, sz);
    *len = sz;
    return P2G_OK;
}

/* ---------------- stand-alone pieces for per-kernel parity ---------------- */
ORC_EXPORT int orc_ifft(const u64* values, u64* coeffs, u32 log_n, u32 ncols) {
    orc_ctx c;
    memset(&c, 0, sizeof c);
    size_t n = (size_t)1 << log_n;
    get_roots(&c, log_n);
    memcpy(coeffs, values, n * ncols * 8);
#pragma omp parallel for schedule(dynamic, 1)
    for (int col = 0; col < (int)ncols; col++) ifft_inplace(&c, coeffs + (size_t)col * n, log_n);
    free(c.roots[log_n]);
    return P2G_OK;
}
ORC_EXPORT int orc_lde(const u64* coeffs, u64* lde_out, u32 log_n, u32 rate_bits, u32 ncols) {
    orc_ctx c;
    memset(&c, 0, sizeof c);
    int log_lde = log_n + rate_bits;
    size_t n = (size_t)1 << log_n, lde = (size_t)1 << log_lde;
    get_roots(&c, log_lde);
#pragma omp parallel
    {
        u64* tmp = (u64*)malloc(lde * 8);
#pragma omp for schedule(dynamic, 1)
        for (int col = 0; col < (int)ncols; col++) {
            memcpy(tmp, coeffs + (size_t)col * n, n * 8);
            memset(tmp + n, 0, (lde - n) * 8);
            coset_scale(tmp, n, GL_GEN);
            fft_inplace(&c, tmp, log_lde);
            for (size_t i = 0; i < lde; i++) lde_out[(size_t)col * lde + bitrev64(i, log_lde)] = tmp[i];
        }
        free(tmp);
    }
    free(c.roots[log_lde]);
    return P2G_OK;
}
ORC_EXPORT int orc_coset_ifft_leaforder(const u64* values, u64* coeffs, u32 log_n, u32 ncols) {
    orc_ctx c;
    memset(&c, 0, sizeof c);
    size_t n = (size_t)1 << log_n;
    get_roots(&c, log_n);
    u64 ginv = gl_inv(GL_GEN);
#pragma omp parallel for schedule(dynamic, 1)
    for (int col = 0; col < (int)ncols; col++) {
        u64* o = coeffs + (size_t)col * n;
        for (size_t i = 0; i < n; i++) o[i] = values[(size_t)col * n + bitrev64(i, log_n)];
        ifft_inplace(&c, o, log_n);
        coset_scale(o, n, ginv);
    }
    free(c.roots[log_n]);
    return P2G_OK;
}
ORC_EXPORT int orc_merkle_cap(const u64* leaves_colmajor, u32 log_leaves, u32 ncols, u32 cap_height, u32 hasher, u8* cap_out,
                              u8* digests_out) {
    size_t nl = (size_t)1 << log_leaves;
    u64* leaves = (u64*)malloc(nl * ncols * 8);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nl; i++)
        for (u32 cidx = 0; cidx < ncols; cidx++) leaves[(size_t)i * ncols + cidx] = leaves_colmajor[(size_t)cidx * nl + i];
    mtree t;
    memset(&t, 0, sizeof t);
    mtree_build(&t, hasher, leaves, log_leaves, ncols, cap_height);
    memcpy(cap_out, mtree_cap(&t), (size_t)mtree_ncap(&t) * t.hs);
    if (digests_out) memcpy(digests_out, t.levels[0], nl * t.hs);
    mtree_free(&t);
    return P2G_OK;
}
ORC_EXPORT int orc_poseidon_permute(const u64* in, u64* out, size_t n) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        u64 s[12];
        memcpy(s, in + 12 * i, 96);
        poseidon_permute(s);
        memcpy(out + 12 * i, s, 96);
    }
    return P2G_OK;
}
ORC_EXPORT int orc_keccak256(const u8* msgs, size_t msg_len, size_t n, u8* out) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) keccak256(msgs + (size_t)i * msg_len, msg_len, out + 32 * (size_t)i);
    return P2G_OK;
}
ORC_EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
