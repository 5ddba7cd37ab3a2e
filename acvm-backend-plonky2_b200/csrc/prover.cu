// The prover: p2g_circuit_create / p2g_prove and the kernels of the stages that are neither NTT nor Merkle nor gate
// evaluation -- permutation argument (Z + partial products, prefix-product scan), openings, FRI batch combination with
// synthetic division by (X - z) as a suffix scan, FRI folding, proof-of-work search, query gathers.
//
// Replaces plonky2 0.2.2 plonk/prover.rs prove_with_partition_witness, the function behind
//     circuit_data.prove(witnesses).unwrap()       /root/reference/plonky2-backend/src/actions/prove_action.rs:96
// step for step: the Fiat-Shamir transcript runs on the host (challenger_t in hash.cuh), every
// data-parallel stage on the device.  Output: plonky2's uncompressed ProofWithPublicInputs::to_bytes layout (App. A.12).
#include <string.h>

#include "internal.h"
#include "nccl_dyn.h"
#include "quotient.h"
#include "advice.cuh"

#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include <thread>
namespace {

// P2G_TRACE=1: host wall-clock trace of the prove stages on stderr (debugging aid)
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0, last;
    Trace() : on(getenv("P2G_TRACE") != nullptr), t0(std::chrono::steady_clock::now()), last(t0) {}
    void mark(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[p2g] %-28s +%8.3f ms  (t=%9.3f)\n", what, std::chrono::duration<double, std::milli>(now - last).count(),
                std::chrono::duration<double, std::milli>(now - t0).count());
        last = now;
    }
};

__device__ __forceinline__ u64 tab_pow_d(const u64* tab, int split, u32 i) {
    return gl_mul(__ldg(tab + (i & ((1u << split) - 1))), __ldg(tab + (1u << split) + (i >> split)));
}

// ---------------------------------------------------------------------------------------------------------------------
// Z and partial products (plonk/prover.rs all_wires_permutation_partial_products)
// ---------------------------------------------------------------------------------------------------------------------
struct ZppParams {
    int logn, num_routed, chunk, nchunk, num_challenges;
    u64 betas[2], gammas[2];
    u64 beta_k[2][P2G_MAX_ROUTED];
};

#define ZPP_MAX_CHUNKS 16

// q[c][m][i] = prod_{r in chunk m} (w_r + beta k_r x + gamma) / (w_r + beta sigma_r + gamma) at row i
__global__ void __launch_bounds__(128) k_zpp_chunks(const __grid_constant__ ZppParams P, const u64* __restrict__ wires, size_t wires_cs,
                                                    const u64* __restrict__ sigmas, const u64* __restrict__ xtab, int xsplit,
                                                    u64* __restrict__ q, size_t i0, size_t i1) {
    const size_t n = (size_t)1 << P.logn;
    size_t i = i0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // rows [i0, i1): a coset-sharded rank computes its row block
    if (i >= i1) return;
    const int c = blockIdx.y;
    const u64 x = tab_pow_d(xtab, xsplit, (u32)i);
    const u64 beta = P.betas[c], gamma = P.gammas[c];
    u64 nums[ZPP_MAX_CHUNKS], dens[ZPP_MAX_CHUNKS];
    for (int m = 0; m < P.nchunk; m++) {
        u64 pn = 1, pd = 1;
        int hi = min((m + 1) * P.chunk, P.num_routed);
        for (int r = m * P.chunk; r < hi; r++) {
            u64 base = gl_add(wires[(size_t)r * wires_cs + i], gamma);
            pn = gl_mul(pn, gl_add(base, gl_mul(P.beta_k[c][r], x)));
            pd = gl_mul(pd, gl_add(base, gl_mul(beta, sigmas[(size_t)r * n + i])));
        }
        nums[m] = pn;
        dens[m] = pd;
    }
    // one inversion for the whole row (Montgomery's trick over the chunk denominators)
    u64 pre[ZPP_MAX_CHUNKS];
    u64 acc = 1;
    for (int m = 0; m < P.nchunk; m++) {
        pre[m] = acc;
        acc = gl_mul(acc, dens[m]);
    }
    u64 inv = gl_inv(acc);
    for (int m = P.nchunk - 1; m >= 0; m--) {
        u64 dinv = gl_mul(inv, pre[m]);
        inv = gl_mul(inv, dens[m]);
        q[((size_t)c * P.nchunk + m) * n + i] = gl_mul(nums[m], dinv);
    }
}

#define SCAN_CH 16  // rows per thread in the scans

// phase 1: totals[c][t] = prod over rows of chunk t of prod_m q[c][m][row]
// chunks [t0, t0 + nt) of 16 rows (a rank's row block); totals is indexed by the local chunk number
__global__ void k_zscan_totals(const __grid_constant__ ZppParams P, const u64* __restrict__ q, u64* __restrict__ totals, size_t nt,
                               size_t t0) {
    const size_t n = (size_t)1 << P.logn;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int c = blockIdx.y;
    u64 acc = 1;
    size_t hi = min(n, (t0 + t + 1) * SCAN_CH);
    for (size_t i = (t0 + t) * SCAN_CH; i < hi; i++)
        for (int m = 0; m < P.nchunk; m++) acc = gl_mul(acc, q[((size_t)c * P.nchunk + m) * n + i]);
    totals[(size_t)c * nt + t] = acc;
}
// phase 2: exclusive prefix product of totals, one block per challenge
// block_total (optional): the product of all nt entries of block b goes to block_total[b]
__global__ void __launch_bounds__(1024) k_scan_mul_excl(u64* totals, size_t nt, u64* block_total) {
    __shared__ u64 sh[1024];
    u64* tt = totals + (size_t)blockIdx.x * nt;
    const int T = blockDim.x;
    size_t per = (nt + T - 1) / T;
    size_t lo = min(nt, threadIdx.x * per), hi = min(nt, lo + per);
    u64 acc = 1;
    for (size_t i = lo; i < hi; i++) acc = gl_mul(acc, tt[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 1; d < T; d <<= 1) {  // Hillis-Steele inclusive scan over the per-thread products
        u64 v = threadIdx.x >= d ? sh[threadIdx.x - d] : 1;
        __syncthreads();
        sh[threadIdx.x] = gl_mul(sh[threadIdx.x], v);
        __syncthreads();
    }
    if (block_total && threadIdx.x == T - 1) block_total[blockIdx.x] = sh[T - 1];
    u64 carry = threadIdx.x ? sh[threadIdx.x - 1] : 1;
    for (size_t i = lo; i < hi; i++) {
        u64 v = tt[i];
        tt[i] = carry;
        carry = gl_mul(carry, v);
    }
}
// phase 3: write Z (column c) and the partial products (column NC + c*NPP + m)
// rank_totals (sharded): [world][NC] products of every rank's row block; rows of this rank start at Z = prod of the ranks before it
__global__ void k_zscan_apply(const __grid_constant__ ZppParams P, const u64* __restrict__ q, const u64* __restrict__ totals, size_t nt,
                              u64* __restrict__ out, size_t t0, const u64* __restrict__ rank_totals, int rank) {
    const size_t n = (size_t)1 << P.logn;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const int c = blockIdx.y, NC = P.num_challenges, NPP = P.nchunk - 1;
    u64 z = totals[(size_t)c * nt + t];
    for (int r = 0; r < rank; r++) z = gl_mul(z, rank_totals[r * NC + c]);
    size_t hi = min(n, (t0 + t + 1) * SCAN_CH);
    for (size_t i = (t0 + t) * SCAN_CH; i < hi; i++) {
        out[(size_t)c * n + i] = z;
        u64 acc = z;
        for (int m = 0; m < P.nchunk; m++) {
            acc = gl_mul(acc, q[((size_t)c * P.nchunk + m) * n + i]);
            if (m < NPP) out[(size_t)(NC + c * NPP + m) * n + i] = acc;
        }
        z = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// openings: sum_i c_i z^i in F_{p^2} against a table of powers (plonk/proof.rs OpeningSet::new; App. A.9)
// ---------------------------------------------------------------------------------------------------------------------
// tab[0..n) = re(z^i), tab[n..2n) = im(z^i)
// one square-and-multiply per 16 consecutive powers (a thread per power spent 40 extension multiplies on each)
#define POW_RUN 16
__global__ void k_e2_powers(u64* tab, size_t n, e2 z) {
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * POW_RUN;
    if (i >= n) return;
    e2 p = e2_pow(z, i);
    const size_t hi = min(n, i + POW_RUN);
    for (; i < hi; i++) {
        tab[i] = p.c0;
        tab[n + i] = p.c1;
        p = e2_mul(p, z);
    }
}

#define EVAL_SPLIT 16384  // coefficients per block
// partial[(col * nsplit + s)] = sum over the s-th slice
__global__ void __launch_bounds__(256) k_eval_polys(const u64* __restrict__ coeffs, size_t cs, size_t n, const u64* __restrict__ ptab,
                                                    e2* __restrict__ partial, int nsplit) {
    __shared__ u64 s0[256], s1[256];
    const int col = blockIdx.x, sp = blockIdx.y;
    const u64* cf = coeffs + (size_t)col * cs;
    size_t lo = (size_t)sp * EVAL_SPLIT, hi = min(n, lo + EVAL_SPLIT);
    u64 a0 = 0, a1 = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        u64 cv = cf[i];
        a0 = gl_add(a0, gl_mul(cv, ptab[i]));
        a1 = gl_add(a1, gl_mul(cv, ptab[n + i]));
    }
    s0[threadIdx.x] = a0;
    s1[threadIdx.x] = a1;
    __syncthreads();
    for (int d = blockDim.x / 2; d > 0; d >>= 1) {
        if (threadIdx.x < d) {
            s0[threadIdx.x] = gl_add(s0[threadIdx.x], s0[threadIdx.x + d]);
            s1[threadIdx.x] = gl_add(s1[threadIdx.x], s1[threadIdx.x + d]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)col * nsplit + sp] = e2_make(s0[0], s1[0]);
}

// ---------------------------------------------------------------------------------------------------------------------
// FRI batch combination + division by (X - z)   (fri/oracle.rs prove_openings; App. A.10)
// ---------------------------------------------------------------------------------------------------------------------
struct CombineArgs {
    const u64* coeffs[4];
    int ncols[4];
    size_t cs[4];
    int num_challenges;
    int logn;
};
// u[b][0/1][k] = (sum_j alpha^j poly_{b,j}[k]) * z_b^k ;  batch 0 = every polynomial, batch 1 = the Z polynomials
__global__ void __launch_bounds__(128) k_fri_combine(CombineArgs a, const e2* __restrict__ apow, const u64* __restrict__ ztab0,
                                                     const u64* __restrict__ ztab1, u64* __restrict__ u, size_t k0, size_t k1) {
    const size_t n = (size_t)1 << a.logn;
    size_t k = k0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // coefficients [k0, k1): a coset-sharded rank combines its block
    if (k >= k1) return;
    u64 r0 = 0, r1 = 0, z0 = 0, z1 = 0;
    int j = 0;
    for (int o = 0; o < 4; o++) {
        const u64* base = a.coeffs[o] + k;
        for (int col = 0; col < a.ncols[o]; col++, j++) {
            u64 cv = base[(size_t)col * a.cs[o]];
            e2 ap = apow[j];
            r0 = gl_add(r0, gl_mul(cv, ap.c0));
            r1 = gl_add(r1, gl_mul(cv, ap.c1));
            if (o == 2 && col < a.num_challenges) {
                e2 aq = apow[col];
                z0 = gl_add(z0, gl_mul(cv, aq.c0));
                z1 = gl_add(z1, gl_mul(cv, aq.c1));
            }
        }
    }
    e2 p0 = e2_mul(e2_make(r0, r1), e2_make(ztab0[k], ztab0[n + k]));
    e2 p1 = e2_mul(e2_make(z0, z1), e2_make(ztab1[k], ztab1[n + k]));
    u[k] = p0.c0;
    u[n + k] = p0.c1;
    u[2 * n + k] = p1.c0;
    u[3 * n + k] = p1.c1;
}
// additive suffix scan over the 4 component arrays u[comp][k] (independent field sums), chunked like the Z scan
__global__ void k_sscan_totals(const u64* __restrict__ u, u64* __restrict__ totals, size_t n, size_t nt) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    const u64* a = u + (size_t)blockIdx.y * n;
    u64 acc = 0;
    size_t hi = min(n, (t + 1) * SCAN_CH);
    for (size_t i = t * SCAN_CH; i < hi; i++) acc = gl_add(acc, a[i]);
    totals[(size_t)blockIdx.y * nt + t] = acc;
}
// exclusive suffix sums of the totals (totals[t] <- sum of totals[t'] for t' > t), one block per component
__global__ void __launch_bounds__(1024) k_scan_add_suffix_excl(u64* totals, size_t nt) {
    __shared__ u64 sh[1024];
    u64* tt = totals + (size_t)blockIdx.x * nt;
    const int T = blockDim.x;
    size_t per = (nt + T - 1) / T;
    size_t lo = min(nt, threadIdx.x * per), hi = min(nt, lo + per);
    u64 acc = 0;
    for (size_t i = lo; i < hi; i++) acc = gl_add(acc, tt[i]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int d = 1; d < T; d <<= 1) {  // inclusive suffix scan across threads
        u64 v = threadIdx.x + d < T ? sh[threadIdx.x + d] : 0;
        __syncthreads();
        sh[threadIdx.x] = gl_add(sh[threadIdx.x], v);
        __syncthreads();
    }
    u64 carry = threadIdx.x + 1 < T ? sh[threadIdx.x + 1] : 0;
    for (size_t i = hi; i-- > lo;) {
        u64 v = tt[i];
        tt[i] = carry;
        carry = gl_add(carry, v);
    }
}
// T_k = sum_{i >= k} u_i ;  q_b[k-1] = T_k * z_b^{-k} ;  final[k-1] = q_0[k-1] * alpha^NC + q_1[k-1] ;  final[n-1] = 0
__global__ void k_sscan_apply(const u64* __restrict__ u, const u64* __restrict__ totals, size_t n, size_t nt,
                              const u64* __restrict__ zinv0, const u64* __restrict__ zinv1, e2 alpha_nc, u64* __restrict__ fin) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nt) return;
    u64 c[4];
    for (int k = 0; k < 4; k++) c[k] = totals[(size_t)k * nt + t];
    size_t lo = t * SCAN_CH, hi = min(n, (t + 1) * SCAN_CH);
    for (size_t i = hi; i-- > lo;) {
        for (int k = 0; k < 4; k++) c[k] = gl_add(c[k], u[(size_t)k * n + i]);
        if (i >= 1) {
            e2 q0 = e2_mul(e2_make(c[0], c[1]), e2_make(zinv0[i], zinv0[n + i]));
            e2 q1 = e2_mul(e2_make(c[2], c[3]), e2_make(zinv1[i], zinv1[n + i]));
            e2 f = e2_add(e2_mul(q0, alpha_nc), q1);
            fin[i - 1] = f.c0;
            fin[n + i - 1] = f.c1;
        }
    }
    if (t == nt - 1) {
        fin[n - 1] = 0;
        fin[2 * n - 1] = 0;
    }
}

// coeffs'[k] = sum_{i < arity} coeffs[arity k + i] beta^i   (fri/prover.rs fri_committed_trees)
__global__ void k_fri_fold(const u64* __restrict__ in, size_t n_in, u64* __restrict__ out, size_t n_out, int arity, e2 beta) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out) return;
    e2 acc = e2_make(0, 0);
    for (int i = arity - 1; i >= 0; i--) {
        size_t idx = (size_t)arity * k + i;
        acc = e2_add(e2_mul(acc, beta), e2_make(in[idx], in[n_in + idx]));
    }
    out[k] = acc.c0;
    out[n_out + k] = acc.c1;
}

// proof of work: smallest witness w in [start, start + count) with >= pow_bits leading zero bits in the response
__global__ void k_pow_search(challenger_t base, u64 start, u64 count, int pow_bits, unsigned long long* best) {
    u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    challenger_t ch = base;
    ch.observe(start + t);
    u64 r = ch.get();
    if (pow_bits == 0 || (r >> (64 - pow_bits)) == 0) atomicMin(best, (unsigned long long)(start + t));
}

// the ABI requires canonical field elements (the reference hands over F::to_canonical_u64); a value >= p would silently
// change the committed polynomial, so the trace is checked on the device (one streaming read, ~0.4 ms at 2^20 rows)
__global__ void k_check_canonical(const u64* __restrict__ v, size_t count, int* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < count; i += stride) bad |= v[i] >= GL_P;
    if (bad) *flag = 1;
}

// p2g_prove_columns hands over plonky2's in-memory GoldilocksField words, which may be any representative < 2^64 (its add / sub
// leave values in [p, 2^64) with probability ~2^-32): reduce them in the library's staging buffer instead of refusing them
__global__ void k_canonicalize(u64* __restrict__ v, size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        const u64 x = v[i];
        if (x >= GL_P) v[i] = x - GL_P;
    }
}

// the same two over rows [0, rows) of `ncols` columns of pitch `pitch` (a sharded rank's share of the routed columns)
__global__ void k_check_canonical_2d(const u64* __restrict__ v, size_t pitch, size_t rows, int ncols, int* __restrict__ flag) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x, count = rows * ncols;
    bool bad = false;
    for (; i < count; i += stride) bad |= v[(i / rows) * pitch + i % rows] >= GL_P;
    if (bad) *flag = 1;
}
__global__ void k_canonicalize_2d(u64* __restrict__ v, size_t pitch, size_t rows, int ncols) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x, count = rows * ncols;
    for (; i < count; i += stride) {
        u64* p = v + (i / rows) * pitch + i % rows;
        const u64 x = *p;
        if (x >= GL_P) *p = x - GL_P;
    }
}

// query phase gathers
__global__ void k_gather_rows(const u64* __restrict__ lde, size_t cs, int ncols, const u32* __restrict__ idx, int nq,
                              u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * ncols) return;
    int q = t / ncols, c = t % ncols;
    out[t] = lde[(size_t)c * cs + idx[q]];
}
__global__ void k_gather_fri_rows(const u64* __restrict__ vals, size_t cs, int arity, const u32* __restrict__ idx, int nq, int shift,
                                  u64* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int w = 2 * arity;
    if (t >= nq * w) return;
    int q = t / w, k = t % w;
    size_t leaf = idx[q] >> shift;
    out[t] = vals[(size_t)(k & 1) * cs + leaf * arity + (k >> 1)];
}
__global__ void k_gather_paths(const digest_t* const* __restrict__ levels, int nlev, const u32* __restrict__ idx, int nq, int shift,
                               digest_t* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * nlev) return;
    int q = t / nlev, k = t % nlev;
    size_t leaf = idx[q] >> shift;
    out[t] = levels[k][(leaf >> k) ^ 1];
}

template <class T>
T* ensure(dbuf<T>& b, size_t count) {  // grow-only scratch: no cudaMalloc / cudaFree on the steady-state prove path
    if (b.n < count) b.alloc(count);
    return b.p;
}

struct PolyBatch {
    int ncols = 0, logn = 0;
    dbuf<u64> coeffs;  // [ncols][N]
    dbuf<u64> lde;     // [ncols][N << rate_bits], leaf order
    MerkleTree tree;
    dbuf<const digest_t*> d_levels;  // device array of level pointers (query gathers)
    std::vector<const digest_t*> h_levels;   // its host image (kept alive for the asynchronous upload)
};

}  // namespace

struct p2g_circuit {
    p2g_circuit_desc d;
    std::vector<p2g_gate> gates;
    std::vector<u64> k_is;
    DevCtx* ctx = nullptr;
    int logn = 0, loglde = 0, hs = 0, h = 0;
    size_t n = 0, lde = 0;
    dbuf<u64> sigma_values;  // [R][N], natural row order (Z computation)
    dbuf<uint8_t> row_gate;  // [N] index into `gates` of the row's gate, from the selector columns (p2g_fill_advice_device)
    dbuf<u64> wires_values;  // [W][N] staging for p2g_prove (host trace upload)
    PolyBatch cs, wires, zpp, quot;
    dbuf<u64> xs, l0s;
    u64 zh[1 << P2G_MAX_RATE], zh_inv[1 << P2G_MAX_RATE];
    digest_t digest;
    std::vector<digest_t> cs_cap;
    std::mutex mu;
    // dumps of the last prove
    dbuf<u64> zpp_values;
    std::vector<u64> challenges;
    std::vector<u64> final_poly;  // c0,c1 interleaved
    std::vector<digest_t> fri_caps, last_caps[3];   // wires, Z/PP, quotient caps of the last proof
    bool proved = false;
    // scratch that survives between proofs
    struct FriLayer {
        dbuf<u64> values;  // [2][cur] leaf order
        PolyBatch tree;
        size_t cur = 0;
        int logcur = 0, ab = 0;
    };
    struct {
        dbuf<u64> q, totals, rank_totals, ztab0, ztab1, u, fin, coeffs_a, coeffs_b, rows;
        dbuf<e2> partial, apow;
        dbuf<unsigned long long> best;
        dbuf<u32> idx;
        dbuf<int> flag, bar;
        dbuf<digest_t> paths;
        std::vector<FriLayer> layers;
    } ws;
    // coset sharding (one process per GPU).  world == 1: single device.  A rank owns the `nzl` cosets
    // [z0, z0 + nzl) = leaves [j0, j0 + lde_l) of every oracle: their LDE columns, their Merkle subtrees down to the
    // cap entries [cap0, cap0 + ncap_l), their share of the quotient evaluation and of the query openings.
    int rank = 0, world = 1, logworld = 0;
    // rows [row0, row1) of the trace are this rank's share of the row-parallel stages (Z / partial products, the FRI batch
    // combination): all of them on a single GPU, N / world consecutive rows when sharded and N / world >= 1024
    bool row_shard = false;
    size_t row0 = 0, row1 = 0;
    int z0 = 0, nzl = 0, ncap_l = 0, cap0 = 0;
    size_t lde_l = 0, j0 = 0;
    int loglde_l = 0;
    p2g_allgather_fn allgather = nullptr;
    void* allgather_user = nullptr;
    // in-library NCCL (p2g_circuit_create_sharded_nccl): collectives are enqueued on the handle's stream, no host round trip
    ncclComm_t nccl = nullptr;
    bool nccl_dead = false;        // the communicator was aborted: every later call returns P2G_ENCCL
    dbuf<uint8_t> nccl_stage;      // device staging of the small host-side gathers (caps, opened rows)
    unsigned collectives = 0;      // exchanges since create (P2G_BUF_SHARD_INFO)
    // peer views of every other rank's wires.coeffs buffer (same process: the pointer itself; another process: CUDA IPC
    // mapping).  When all ranks could map all peers, the inverse NTT of the trace stores its column block straight into the
    // peers over NVLink and the coefficient all-gather disappears; otherwise the NCCL all-gather is used.
    int npeer = 0;
    u64* peer_coeffs[P2G_MAX_PEERS] = {};
    std::vector<void*> ipc_opened;
    // host-trace upload, pipelined with the inverse NTT: column chunks go up on `copy` while earlier chunks are transformed
    struct Upload {
        cudaStream_t copy = nullptr;
        std::vector<cudaEvent_t> pool;                    // reusable events
        std::vector<std::tuple<int, int, cudaEvent_t>> chunks;   // [col0, col1) ready when the event fires
        cudaEvent_t start = nullptr, routed = nullptr;    // routed: columns [0, R) needed by the Z computation (sharded only)
        cudaEvent_t last = nullptr;
        bool active = false;
        bool canon = false;   // reduce the uploaded words mod p before anything reads them (p2g_prove_columns)
        int fill_from = -1;   // >= 0: only the columns below it are uploaded; the chunks from it on (null events) are computed on
                              // the device by k_fill_advice once the uploaded ones are in place (p2g_prove_routed_columns)
        double bytes = 0;
        // pageable caller memory: chunks are first copied (multi-threaded) into a ring of pinned staging buffers, so the DMA
        // runs at full PCIe rate instead of the driver's single-threaded bounce copy
        static const int NSTAGE = 3;
        u64* stage[NSTAGE] = {};
        size_t stage_words = 0;
        cudaEvent_t stage_free[NSTAGE] = {};
    } up;
    ~p2g_circuit() {
        for (cudaEvent_t e : up.pool) cudaEventDestroy(e);
        if (up.copy) cudaStreamDestroy(up.copy);
        for (int i = 0; i < Upload::NSTAGE; i++) {
            if (up.stage[i]) cudaFreeHost(up.stage[i]);
            if (up.stage_free[i]) cudaEventDestroy(up.stage_free[i]);
        }
        for (void* q : ipc_opened) cudaIpcCloseMemHandle(q);
        if (nccl) {
            if (nccl_dead) nccl_api().CommAbort(nccl);
            else nccl_api().CommDestroy(nccl);
        }
        free_child_ctx(ctx);
    }
};

namespace {

void upload_level_ptrs(DevCtx* c, PolyBatch& b) {
    int nlev = (int)b.tree.levels.size() - 1;
    if (nlev <= 0) return;
    std::vector<const digest_t*> p(nlev);
    for (int k = 0; k < nlev; k++) p[k] = b.tree.levels[k].p;
    if (b.d_levels.n == (size_t)nlev && p == b.h_levels) return;   // trees keep their buffers between proofs
    if (b.d_levels.n != (size_t)nlev) b.d_levels.alloc(nlev);
    // stream order is enough for the kernels that read the table (no host synchronisation); only reached when a tree was
    // (re)allocated: circuit build and the first proof.  The source is pageable memory, which the runtime stages before returning.
    b.h_levels = p;
    CUDA_CHECK(cudaMemcpyAsync(b.d_levels.p, b.h_levels.data(), sizeof(void*) * nlev, cudaMemcpyHostToDevice, c->stream));
}

[[noreturn]] void nccl_fail(p2g_circuit* C, const std::string& what) {
    // a failed or timed-out collective leaves the communicator unusable: abort it so nothing blocks in it again
    C->nccl_dead = true;
    throw p2g_error(P2G_ENCCL, what);
}
void nccl_check(p2g_circuit* C, ncclResult_t r, const char* what) {
    if (r == ncclSuccess || r == ncclInProgress) return;
    nccl_fail(C, std::string(what) + ": " + nccl_api().GetErrorString(r));
}
// Host wait for the handle's stream.  With an NCCL communicator the wait polls, watches the communicator's asynchronous error
// state and gives up after P2G_NCCL_TIMEOUT_S seconds (default 600): a rank that failed on its own never leaves its peers
// blocked for ever inside a collective -- they return P2G_ENCCL.
void stream_sync(p2g_circuit* C) {
    cudaStream_t st = C->ctx->stream;
    if (!C->nccl) {
        CUDA_CHECK(cudaStreamSynchronize(st));
        return;
    }
    static const double limit = getenv("P2G_NCCL_TIMEOUT_S") ? atof(getenv("P2G_NCCL_TIMEOUT_S")) : 600.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (unsigned spin = 0;; spin++) {
        cudaError_t q = cudaStreamQuery(st);
        if (q == cudaSuccess) return;
        if (q != cudaErrorNotReady) CUDA_CHECK(q);
        if ((spin & 1023) == 1023) {
            ncclResult_t ae = ncclSuccess;
            nccl_check(C, nccl_api().CommGetAsyncError(C->nccl, &ae), "ncclCommGetAsyncError");
            if (ae != ncclSuccess && ae != ncclInProgress) nccl_fail(C, std::string("NCCL asynchronous error: ") + nccl_api().GetErrorString(ae));
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit)
                nccl_fail(C, "collective timed out (a peer rank failed or never arrived)");
        }
    }
}

// All-gather `bytes` from every rank into recv (rank order).  Device buffers: enqueued on the handle's stream (NCCL) -- the
// caller keeps launching; host buffers (caps, opened rows, a few KB): staged through the device, returns when recv is filled.
void shard_allgather(p2g_circuit* C, const void* send, void* recv, size_t bytes, bool is_device) {
    C->collectives++;
    if (C->nccl) {
        if (C->nccl_dead) throw p2g_error(P2G_ENCCL, "the NCCL communicator of this handle was aborted by an earlier failure");
        cudaStream_t st = C->ctx->stream;
        if (is_device) {
            nccl_check(C, nccl_api().AllGather(send, recv, bytes, ncclUint8, C->nccl, st), "ncclAllGather");
            return;
        }
        uint8_t* stage = ensure(C->nccl_stage, bytes * C->world);
        CUDA_CHECK(cudaMemcpyAsync(stage + (size_t)C->rank * bytes, send, bytes, cudaMemcpyHostToDevice, st));
        nccl_check(C, nccl_api().AllGather(stage + (size_t)C->rank * bytes, stage, bytes, ncclUint8, C->nccl, st), "ncclAllGather");
        CUDA_CHECK(cudaMemcpyAsync(recv, stage, bytes * C->world, cudaMemcpyDeviceToHost, st));
        stream_sync(C);
        return;
    }
    if (is_device) CUDA_CHECK(cudaStreamSynchronize(C->ctx->stream));   // the host binding runs on its own stream
    int rc = C->allgather(C->allgather_user, send, recv, bytes, is_device ? 1 : 0);
    if (rc != 0) throw p2g_error(P2G_ENCCL, "allgather callback failed (" + std::to_string(rc) + ")");
}

// Row blocks of `ncols` columns: rank r holds rows [r * per, (r + 1) * per) of every column (column c at base + c * cs); afterwards
// every rank holds every row.  NCCL: the per-column in-place all-gathers form one group (one launch); callback: one call per column.
void shard_allgather_cols(p2g_circuit* C, u64* base, int ncols, size_t cs, size_t per) {
    if (C->nccl) {
        if (C->nccl_dead) throw p2g_error(P2G_ENCCL, "the NCCL communicator of this handle was aborted by an earlier failure");
        C->collectives++;
        nccl_check(C, nccl_api().GroupStart(), "ncclGroupStart");
        for (int col = 0; col < ncols; col++) {
            u64* colp = base + (size_t)col * cs;
            nccl_check(C, nccl_api().AllGather(colp + (size_t)C->rank * per, colp, per * 8, ncclUint8, C->nccl, C->ctx->stream), "ncclAllGather");
        }
        nccl_check(C, nccl_api().GroupEnd(), "ncclGroupEnd");
        return;
    }
    for (int col = 0; col < ncols; col++) {
        u64* colp = base + (size_t)col * cs;
        shard_allgather(C, colp + (size_t)C->rank * per, colp, per * 8, true);
    }
}

// coefficients already in b.coeffs: LDE of this rank's cosets + their Merkle subtrees
void commit_from_coeffs(p2g_circuit* C, PolyBatch& b) {
    DevCtx* c = C->ctx;
    size_t want = (size_t)b.ncols * C->lde_l;
    if (b.lde.n != want) b.lde.alloc(want);
    ntt_lde(c, b.coeffs.p, C->n, b.lde.p, C->lde_l, C->logn, C->d.rate_bits, b.ncols, GL_GEN, C->z0, C->nzl);
    merkle_build(c, &b.tree, b.lde.p, C->lde_l, C->loglde_l, b.ncols, (int)C->d.cap_height - C->logworld, C->h);
    upload_level_ptrs(c, b);
}
// values -> coefficients.  Sharded: each rank transforms a contiguous block of columns and the blocks are all-gathered
// (the coefficient buffer is padded to world * ceil(ncols / world) columns so the blocks are equal).
// the block of columns whose inverse NTT this rank computes; false = every rank transforms all columns (no exchange)
bool column_block(const p2g_circuit* C, int ncols, int* c0, int* c1) {
    if (C->world == 1 || ncols < 4 * C->world) {
        *c0 = 0;
        *c1 = ncols;
        return false;
    }
    const int per = (ncols + C->world - 1) / C->world;
    *c0 = std::min(ncols, C->rank * per);
    *c1 = std::min(ncols, *c0 + per);
    return true;
}
// uploaded chunk [a, e) of the staging buffer: reduce mod p in place when the caller handed over raw GoldilocksField words
// ---- device-side witness fill (advice.cuh) ---------------------------------------------------------------------------------------
// the gate of every row: the selector column of the gate's group holds the gate's index, the others hold UNUSED (plonky2
// plonk/circuit_builder.rs selector_polynomials)
__global__ void k_row_gate(const u64* __restrict__ consts, size_t n, int num_selectors, uint8_t* __restrict__ row_gate) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    u32 g = 255;
    for (int s = 0; s < num_selectors; s++) {
        const u64 v = consts[(size_t)s * n + r];
        if (v != 0xFFFFFFFFULL) g = (u32)v;
    }
    row_gate[r] = (uint8_t)g;
}
struct AdviceGates {
    int num_gates, num_routed, num_wires, pad_;
    u32 kind[P2G_MAX_GATES];
    u32 params[P2G_MAX_GATES][4];
};
__global__ void __launch_bounds__(128) k_fill_advice(const __grid_constant__ AdviceGates T, const uint8_t* __restrict__ row_gate,
                                                     u64* __restrict__ wires, size_t n) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const u32 g = row_gate[r];
    if (g >= (u32)T.num_gates) return;
    auto get = [&](u32 col) -> u64 { return col < (u32)T.num_wires ? wires[(size_t)col * n + r] : 0; };
    auto put = [&](u32 col, u64 v) {
        if (col >= (u32)T.num_routed && col < (u32)T.num_wires) wires[(size_t)col * n + r] = v;
    };
    fill_advice_row(T.kind[g], T.params[g], get, put);
}

static void launch_fill_advice(p2g_circuit* C, u64* d_wires) {   // on the handle's stream; the caller orders it after the routed columns
    const p2g_circuit_desc& d = C->d;
    AdviceGates T = {};
    T.num_gates = (int)d.num_gates;
    T.num_routed = (int)d.num_routed_wires;
    T.num_wires = (int)d.num_wires;
    for (u32 g = 0; g < d.num_gates; g++) {
        T.kind[g] = C->gates[g].kind;
        for (int k = 0; k < 4; k++) T.params[g][k] = C->gates[g].params[k];
    }
    k_fill_advice<<<(unsigned)((C->n + 127) / 128), 128, 0, C->ctx->stream>>>(T, C->row_gate.p, d_wires, C->n);
    CUDA_CHECK(cudaGetLastError());
    count_launch(C->ctx);
}

void canon_chunk(p2g_circuit* C, const p2g_circuit::Upload* up, const u64* d_values, size_t values_cs, int a, int e) {
    if (!up || !up->canon || e <= a) return;
    const size_t cnt = (size_t)(e - a) * values_cs;
    const unsigned blocks = (unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 16);
    k_canonicalize<<<blocks, 256, 0, C->ctx->stream>>>(const_cast<u64*>(d_values) + (size_t)a * values_cs, cnt);
    count_launch(C->ctx);
}
void ifft_columns(p2g_circuit* C, PolyBatch& b, const u64* d_values, size_t values_cs, const p2g_circuit::Upload* up = nullptr) {
    const int per = (b.ncols + C->world - 1) / C->world;
    size_t want = (size_t)per * C->world * C->n;
    if (b.coeffs.n != want) b.coeffs.alloc(want);
    int c0, c1;
    const bool sharded = column_block(C, b.ncols, &c0, &c1);
    // fused exchange: only for the trace (the buffer the peers mapped at create time)
    const bool fused = sharded && C->npeer == C->world - 1 && &b == &C->wires;
    u64* peers[P2G_MAX_PEERS];
    auto peer_ptrs = [&](int col) {
        for (int p = 0; p < C->npeer; p++) peers[p] = C->peer_coeffs[p] + (size_t)col * C->n;
        return fused ? C->npeer : 0;
    };
    if (up) {   // chunks of [c0, c1) arrive on the copy stream; transform each as soon as it is there
        for (auto& ch : up->chunks) {
            int a = std::get<0>(ch), e = std::get<1>(ch);
            CUDA_CHECK(cudaStreamWaitEvent(C->ctx->stream, std::get<2>(ch), 0));
            canon_chunk(C, up, d_values, values_cs, a, e);
            int np = peer_ptrs(a);
            ntt_ifft(C->ctx, d_values + (size_t)a * values_cs, values_cs, b.coeffs.p + (size_t)a * C->n, C->n, C->logn, e - a, np, peers);
        }
    } else {
        int np = peer_ptrs(c0);
        ntt_ifft(C->ctx, d_values + (size_t)c0 * values_cs, values_cs, b.coeffs.p + (size_t)c0 * C->n, C->n, C->logn, c1 - c0, np, peers);
    }
    if (fused && C->nccl) {
        // every rank's block is in every buffer once all ranks' kernels have completed: a stream-ordered barrier (a 4-byte
        // all-gather enqueued behind the stores), no host round trip
        int* bar = ensure(C->ws.bar, (size_t)C->world);
        shard_allgather(C, bar + C->rank, bar, sizeof(int), true);
    } else if (fused) {
        CUDA_CHECK(cudaStreamSynchronize(C->ctx->stream));   // drain, then a (tiny) barrier through the callback
        int one = 1;
        std::vector<int> all(C->world);
        shard_allgather(C, &one, all.data(), sizeof(int), false);
    } else if (sharded) {
        shard_allgather(C, b.coeffs.p + (size_t)C->rank * per * C->n, b.coeffs.p, (size_t)per * C->n * 8, true);
    }
}

// Map every peer's wires.coeffs buffer (called once at create, after the buffer exists).  All ranks agree on the outcome.
void setup_peers(p2g_circuit* C) {
    if (C->world == 1 || getenv("P2G_NO_PEER")) return;
    struct Info {
        int pid, device;
        unsigned long long ptr;
        cudaIpcMemHandle_t h;
    };
    Info mine = {};
    mine.pid = (int)getpid();
    mine.device = C->ctx->device;
    mine.ptr = (unsigned long long)C->wires.coeffs.p;
    bool ok = cudaIpcGetMemHandle(&mine.h, C->wires.coeffs.p) == cudaSuccess;
    if (!ok) cudaGetLastError();
    std::vector<Info> all(C->world);
    shard_allgather(C, &mine, all.data(), sizeof(Info), false);
    int np = 0;
    for (int r = 0; r < C->world && ok; r++) {
        if (r == C->rank) continue;
        u64* q = nullptr;
        if (all[r].pid == mine.pid) {   // another handle of this process (thread-group ranks)
            if (all[r].device != mine.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(all[r].device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
            q = (u64*)all[r].ptr;
        } else {
            void* m = nullptr;
            if (cudaIpcOpenMemHandle(&m, all[r].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                C->ipc_opened.push_back(m);
                q = (u64*)m;
            } else {
                cudaGetLastError();
                ok = false;
            }
        }
        if (ok) C->peer_coeffs[np++] = q;
    }
    int flag = ok ? 1 : 0;
    std::vector<int> flags(C->world);
    shard_allgather(C, &flag, flags.data(), sizeof(int), false);
    for (int f : flags) ok = ok && f;
    C->npeer = ok ? np : 0;
}
void commit_from_values(p2g_circuit* C, PolyBatch& b, const u64* d_values, size_t values_cs, const p2g_circuit::Upload* up = nullptr) {
    if (up && C->world == 1) {
        // single GPU, trace arriving from the host: inverse NTT + LDE per column chunk, so both hide behind the upload
        DevCtx* c = C->ctx;
        size_t want = (size_t)b.ncols * C->n;
        if (b.coeffs.n < want) b.coeffs.alloc(want);
        if (b.lde.n != (size_t)b.ncols * C->lde_l) b.lde.alloc((size_t)b.ncols * C->lde_l);
        for (auto& ch : up->chunks) {
            int a = std::get<0>(ch), e = std::get<1>(ch);
            const bool filled = up->fill_from >= 0 && a >= up->fill_from;
            // the routed chunks before this one have been waited for and reduced on this stream: their advice columns can be computed
            if (filled && a == up->fill_from) {
                // wires no generator sets read as zero (full_witness): clear the advice region of the staging matrix, which still
                // holds the previous proof's columns, then let every row's gate write its own
                CUDA_CHECK(cudaMemsetAsync(const_cast<u64*>(d_values) + (size_t)up->fill_from * values_cs, 0,
                                           (size_t)(b.ncols - up->fill_from) * values_cs * 8, c->stream));
                launch_fill_advice(C, const_cast<u64*>(d_values));
            }
            if (std::get<2>(ch)) CUDA_CHECK(cudaStreamWaitEvent(c->stream, std::get<2>(ch), 0));
            if (!filled) canon_chunk(C, up, d_values, values_cs, a, e);
            ntt_ifft(c, d_values + (size_t)a * values_cs, values_cs, b.coeffs.p + (size_t)a * C->n, C->n, C->logn, e - a);
            ntt_lde(c, b.coeffs.p + (size_t)a * C->n, C->n, b.lde.p + (size_t)a * C->lde_l, C->lde_l, C->logn, C->d.rate_bits, e - a,
                    GL_GEN, C->z0, C->nzl);
        }
        merkle_build(c, &b.tree, b.lde.p, C->lde_l, C->loglde_l, b.ncols, (int)C->d.cap_height - C->logworld, C->h);
        upload_level_ptrs(c, b);
        return;
    }
    ifft_columns(C, b, d_values, values_cs, up);
    commit_from_coeffs(C, b);
}
// the 2^cap_height cap of an oracle: this rank's subtree roots, all-gathered when sharded (the one collective of the
// commitment)
// `status` (optional, sharded proofs): a rank-local failure flag that travels with the cap.  On return it holds the OR over all
// ranks, so every rank takes the same decision before the next collective (a rank that threw on its own would leave the
// others blocked in it).
std::vector<digest_t> read_cap(p2g_circuit* C, const MerkleTree& t, bool sharded = true, int* status = nullptr) {
    std::vector<digest_t> mine(t.ncap() + 1);
    CUDA_CHECK(cudaMemcpyAsync(mine.data(), t.cap(), sizeof(digest_t) * t.ncap(), cudaMemcpyDeviceToHost, C->ctx->stream));
    stream_sync(C);
    if (C->world == 1 || !sharded) {
        mine.pop_back();
        return mine;
    }
    memset(&mine.back(), 0, sizeof(digest_t));
    mine.back().w[0] = status ? (u64)*status : 0;
    const size_t per = mine.size();
    std::vector<digest_t> all(per * C->world), cap;
    shard_allgather(C, mine.data(), all.data(), sizeof(digest_t) * per, false);
    int any = 0;
    for (int r = 0; r < C->world; r++) {
        cap.insert(cap.end(), all.begin() + r * per, all.begin() + r * per + (per - 1));
        any |= (int)all[r * per + per - 1].w[0];
    }
    if (status) *status = any;
    return cap;
}

struct Writer {
    uint8_t* p;
    size_t len = 0, cap;
    Writer(uint8_t* out, size_t capacity) : p(out), cap(out ? capacity : 0) {}
    void put(const void* src, size_t nbytes) {
        if (len + nbytes <= cap) memcpy(p + len, src, nbytes);
        len += nbytes;
    }
    void u64v(u64 x) { put(&x, 8); }
    void e2v(e2 x) {
        u64v(x.c0);
        u64v(x.c1);
    }
    void u8v(uint8_t x) { put(&x, 1); }
    void digest(const digest_t& d, int hs) { put(&d, hs); }
};

void validate_desc(const p2g_circuit_desc* d) {
    auto bad = [](const char* m) { throw p2g_error(P2G_EBADARG, std::string("p2g_circuit_create: ") + m); };
    if (!d) bad("null descriptor");
    if (d->struct_size != sizeof(p2g_circuit_desc)) bad("struct_size mismatch (ABI)");
    if (d->degree_bits > 24) bad("degree_bits > 24");
    if (d->num_challenges < 1 || d->num_challenges > 2) bad("num_challenges must be 1 or 2");
    if (d->rate_bits < 1 || d->rate_bits > P2G_MAX_RATE) bad("rate_bits out of range");
    if (d->quotient_degree_factor != (1u << d->rate_bits)) bad("quotient_degree_factor must equal 2^rate_bits");
    if (d->num_routed_wires == 0 || d->num_routed_wires > P2G_MAX_ROUTED || d->num_routed_wires > d->num_wires) bad("num_routed_wires");
    if (d->num_gates == 0 || d->num_gates > P2G_MAX_GATES || !d->gates) bad("gate table");
    if (d->hasher > 1) bad("hasher");
    if (d->num_fri_layers > P2G_MAX_FRI_LAYERS) bad("num_fri_layers");
    if (!d->constants_sigmas || !d->k_is) bad("null preprocessed data");
    u32 nchunk = (d->num_routed_wires + d->quotient_degree_factor - 1) / d->quotient_degree_factor;
    if (d->num_partial_products + 1 != nchunk || nchunk > ZPP_MAX_CHUNKS) bad("num_partial_products");
    u32 nterms = d->num_challenges * (2 + d->num_partial_products) + d->num_gate_constraints;
    if (nterms > P2G_MAX_TERMS) bad("too many constraint terms");
    u32 red = 0;
    for (u32 i = 0; i < d->num_fri_layers; i++) {
        if (d->reduction_arity_bits[i] < 1 || d->reduction_arity_bits[i] > 6) bad("reduction_arity_bits");
        red += d->reduction_arity_bits[i];
    }
    if (red > d->degree_bits) bad("FRI schedule longer than the polynomial");
    if (d->num_selectors > d->num_constants) bad("num_selectors > num_constants");
    if (d->pow_bits > 64) bad("pow_bits > 64");
    if (d->cap_height > d->degree_bits + d->rate_bits) bad("cap_height exceeds the LDE tree height");
    {   // every FRI layer's tree must have at least 2^cap_height leaves (plonky2's MerkleTree::new asserts the same)
        u32 bits = d->degree_bits + d->rate_bits;
        for (u32 i = 0; i < d->num_fri_layers; i++) {
            bits -= d->reduction_arity_bits[i];
            if (bits < d->cap_height) bad("FRI layer has fewer than 2^cap_height leaves");
        }
    }
    for (u32 g = 0; g < d->num_gates; g++) {
        const p2g_gate& gt = d->gates[g];
        if (gt.kind >= P2G_GATE_KIND_COUNT) bad("unknown gate kind");
        if (gt.selector_index >= d->num_selectors || gt.group_lo > g || gt.group_hi <= g || gt.group_hi > d->num_gates) bad("selector data");
        // wires / gate-local constants / constraints implied by (kind, params): the same formulas as circuit.py and p2g.hpp
        const u32* p = gt.params;
        uint64_t rw = 0, rc = 0, nk = 0;
        switch (gt.kind) {
        case P2G_GATE_NOOP: break;
        case P2G_GATE_CONSTANT: rw = rc = nk = p[0]; break;
        case P2G_GATE_PUBLIC_INPUT: rw = nk = 4; break;
        case P2G_GATE_ARITHMETIC: rw = 4ull * p[0]; rc = 2; nk = p[0]; break;
        case P2G_GATE_BASE_SUM:
            if (p[0] < 2 || p[0] > 256) bad("BaseSumGate base out of range");
            rw = 1ull + p[1]; nk = 1ull + p[1];
            break;
        case P2G_GATE_POSEIDON: rw = 135; nk = 123; break;
        case P2G_GATE_RANDOM_ACCESS:
            if (p[0] > 6) bad("RandomAccessGate bits > 6");
            rw = (2ull + (1ull << p[0])) * p[1] + p[2] + (uint64_t)p[1] * p[0];
            rc = p[2];
            nk = (p[0] + 2ull) * p[1] + p[2];
            break;
        case P2G_GATE_U32_ARITHMETIC: rw = 38ull * p[0]; nk = 36ull * p[0]; break;
        case P2G_GATE_U32_ADD_MANY: rw = (p[0] + 3ull) * p[1] + 18ull * p[1]; nk = 21ull * p[1]; break;
        case P2G_GATE_U32_SUBTRACTION: rw = 21ull * p[0]; nk = 19ull * p[0]; break;
        case P2G_GATE_U32_RANGE_CHECK: rw = 17ull * p[0]; nk = 17ull * p[0]; break;
        case P2G_GATE_COMPARISON: {
            if (p[1] == 0 || p[0] == 0) bad("ComparisonGate num_bits / num_chunks must be positive");
            const uint64_t cb = (p[0] + p[1] - 1) / p[1];
            if (cb > 8) bad("ComparisonGate chunk bits > 8");
            rw = 4 + 5ull * p[1] + cb + 1;
            nk = 6 + 5ull * p[1] + cb;
            break;
        }
        default: break;
        }
        if (rw > d->num_wires) bad("gate needs more wires than num_wires");
        if (rc + d->num_selectors > d->num_constants) bad("gate needs more constants than num_constants - num_selectors");
        if (nk != gt.num_constraints) bad("gate num_constraints does not match (kind, params)");
        if (gt.num_constraints > d->num_gate_constraints) bad("gate num_constraints > num_gate_constraints");
    }
}

void fill_gates(QuotientParams& qp, const p2g_circuit* C) {
    qp.num_gates = (int)C->gates.size();
    for (size_t g = 0; g < C->gates.size(); g++) {
        GateDev& o = qp.gates[g];
        const p2g_gate& s = C->gates[g];
        o.kind = s.kind;
        for (int k = 0; k < 4; k++) o.params[k] = s.params[k];
        o.selector_index = s.selector_index;
        o.group_lo = s.group_lo;
        o.group_hi = s.group_hi;
        o.num_constraints = s.num_constraints;
    }
}

}  // namespace

// =====================================================================================================================
// circuit handle
// =====================================================================================================================
static int circuit_create_impl(const p2g_circuit_desc* desc, int device, int rank, int world, p2g_allgather_fn allgather,
                               void* user, p2g_circuit** out, const uint8_t* nccl_id = nullptr) {
    if (out) *out = nullptr;
    p2g_circuit* C = nullptr;
    int rc = guard([&] {
        if (!out) throw p2g_error(P2G_EBADARG, "p2g_circuit_create: null out");
        validate_desc(desc);
        int logworld = 0;
        while ((1 << logworld) < world) logworld++;
        if (world < 1 || (1 << logworld) != world || rank < 0 || rank >= world || (world > 1 && !allgather && !nccl_id) ||
            logworld > (int)desc->rate_bits || logworld > (int)desc->cap_height)
            throw p2g_error(P2G_EBADARG, "p2g_circuit_create_sharded: world must be a power of two <= 2^min(rate_bits, cap_height), "
                                         "0 <= rank < world, and an allgather callback is required");
        C = new p2g_circuit();
        DevCtx* c = new_child_ctx(device);   // own stream: proofs on different handles overlap on the device
        C->d = *desc;
        C->ctx = c;
        C->rank = rank;
        C->world = world;
        C->logworld = logworld;
        C->allgather = allgather;
        C->allgather_user = user;
        if (nccl_id && world > 1) {
            const NcclApi& api = nccl_api();
            if (!api.ok) throw p2g_error(P2G_ENCCL, "libnccl.so.2 could not be loaded (set P2G_NCCL_LIB)");
            ncclUniqueId id;
            static_assert(sizeof(ncclUniqueId) == P2G_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
            memcpy(&id, nccl_id, sizeof id);
            CUDA_CHECK(cudaSetDevice(device));
            ncclResult_t r = api.CommInitRank(&C->nccl, world, id, rank);
            if (r != ncclSuccess) {
                C->nccl = nullptr;
                throw p2g_error(P2G_ENCCL, std::string("ncclCommInitRank: ") + api.GetErrorString(r));
            }
        }
        C->gates.assign(desc->gates, desc->gates + desc->num_gates);
        C->k_is.assign(desc->k_is, desc->k_is + desc->num_routed_wires);
        C->d.gates = C->gates.data();
        C->d.k_is = C->k_is.data();
        C->d.constants_sigmas = nullptr;
        C->d.circuit_digest = nullptr;
        C->logn = desc->degree_bits;
        C->loglde = desc->degree_bits + desc->rate_bits;
        C->n = (size_t)1 << C->logn;
        C->lde = (size_t)1 << C->loglde;
        C->loglde_l = C->loglde - logworld;
        C->lde_l = C->lde >> logworld;
        C->j0 = (size_t)rank * C->lde_l;
        C->row_shard = world > 1 && (C->n >> logworld) >= 1024 && !getenv("P2G_NO_ROW_SHARD");
        C->row0 = C->row_shard ? (size_t)rank * (C->n >> logworld) : 0;
        C->row1 = C->row_shard ? C->row0 + (C->n >> logworld) : C->n;
        C->nzl = (1 << desc->rate_bits) >> logworld;
        C->z0 = rank * C->nzl;
        C->h = desc->hasher;
        C->hs = hasher_bytes(C->h);
        const int Cc = desc->num_constants, R = desc->num_routed_wires, P = Cc + R;
        const size_t n = C->n;
        for (size_t i = 0; i < (size_t)P * n; i++)
            if (desc->constants_sigmas[i] >= GL_P) throw p2g_error(P2G_EBADARG, "p2g_circuit_create: non-canonical preprocessed value");
        // preprocessed commitment (what CircuitBuilder::build() does, circuit_translation/mod.rs:81)
        dbuf<u64> vals((size_t)P * n);
        CUDA_CHECK(cudaMemcpyAsync(vals.p, desc->constants_sigmas, (size_t)P * n * 8, cudaMemcpyHostToDevice, c->stream));
        C->sigma_values.alloc((size_t)R * n);
        CUDA_CHECK(cudaMemcpyAsync(C->sigma_values.p, vals.p + (size_t)Cc * n, (size_t)R * n * 8, cudaMemcpyDeviceToDevice, c->stream));
        C->row_gate.alloc(n);
        k_row_gate<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(vals.p, n, (int)desc->num_selectors, C->row_gate.p);
        CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        C->cs.ncols = P;
        commit_from_values(C, C->cs, vals.p, n);
        C->cs_cap = read_cap(C, C->cs.tree);
        // Z_H on the 2^rate_bits cosets: (shift * omega_lde^r)^N - 1 = shift^N * omega_{2^rate}^r - 1
        u64 gn = gl_pow(GL_GEN, (u64)n), wr = gl_root_of_unity(desc->rate_bits);
        for (u32 r = 0; r < (1u << desc->rate_bits); r++) {
            C->zh[r] = gl_sub(gl_mul(gn, gl_pow(wr, r)), 1);
            C->zh_inv[r] = gl_inv(C->zh[r]);
        }
        C->xs.alloc(C->lde);
        C->l0s.alloc(C->lde);
        quotient_points(c, C->xs.p, C->l0s.p, C->logn, desc->rate_bits, C->zh);
        if (desc->circuit_digest) {
            memset(&C->digest, 0, sizeof(digest_t));
            memcpy(&C->digest, desc->circuit_digest, C->hs);
        } else {
            // H.hash_no_pad(cap.flatten() || H.hash_pad([]).to_vec() || [degree_bits])   
            std::vector<u64> parts;
            u64 e[4];
            for (const digest_t& dg : C->cs_cap) {
                digest_to_elems(C->h, dg, e);
                parts.insert(parts.end(), e, e + 4);
            }
            const u64 pad[8] = {1, 0, 0, 0, 0, 0, 0, 1};
            digest_to_elems(C->h, hash_no_pad(C->h, pad, 8), e);
            parts.insert(parts.end(), e, e + 4);
            parts.push_back(desc->degree_bits);
            C->digest = hash_no_pad(C->h, parts.data(), parts.size());
        }
        // persistent buffers of the per-proof commitments
        const int W = desc->num_wires, NC = desc->num_challenges, nzp = NC * (1 + desc->num_partial_products),
                  nq = NC * desc->quotient_degree_factor;
        auto padded = [&](int cols) { return (size_t)((cols + world - 1) / world) * world; };
        C->wires.ncols = W;
        C->wires.coeffs.alloc(padded(W) * n);
        C->wires.lde.alloc((size_t)W * C->lde_l);
        C->zpp.ncols = nzp;
        C->zpp.coeffs.alloc(padded(nzp) * n);
        C->zpp.lde.alloc((size_t)nzp * C->lde_l);
        C->zpp_values.alloc((size_t)nzp * n);
        C->quot.ncols = nq;
        C->quot.coeffs.alloc(padded(nq) * n);
        C->quot.lde.alloc((size_t)nq * C->lde_l);
        setup_peers(C);
        *out = C;
    });
    if (rc != P2G_OK && C) delete C;
    return rc;
}

extern "C" int p2g_circuit_create(const p2g_circuit_desc* desc, int device, p2g_circuit** out) {
    return circuit_create_impl(desc, device, 0, 1, nullptr, nullptr, out);
}
extern "C" int p2g_circuit_create_sharded(const p2g_circuit_desc* desc, int device, int rank, int world, p2g_allgather_fn allgather,
                                          void* user, p2g_circuit** out) {
    return circuit_create_impl(desc, device, rank, world, allgather, user, out);
}

// One process (or thread) per GPU, the library owns the communicator: rank 0 obtains an id with p2g_nccl_unique_id, hands it to the
// other ranks by any means, and every rank calls this with the same descriptor.  Exchanges are ncclAllGather on the handle's stream.
extern "C" int p2g_nccl_unique_id(uint8_t* id_out) {
    return guard([&] {
        if (!id_out) throw p2g_error(P2G_EBADARG, "p2g_nccl_unique_id: null argument");
        const NcclApi& api = nccl_api();
        if (!api.ok) throw p2g_error(P2G_ENCCL, "libnccl.so.2 could not be loaded (set P2G_NCCL_LIB)");
        ncclUniqueId id;
        ncclResult_t r = api.GetUniqueId(&id);
        if (r != ncclSuccess) throw p2g_error(P2G_ENCCL, std::string("ncclGetUniqueId: ") + api.GetErrorString(r));
        memcpy(id_out, &id, sizeof id);
    });
}
extern "C" int p2g_circuit_create_sharded_nccl(const p2g_circuit_desc* desc, int device, int rank, int world, const uint8_t* nccl_id,
                                               p2g_circuit** out) {
    if (world > 1 && !nccl_id) {
        set_last_error("p2g_circuit_create_sharded_nccl: null id");
        if (out) *out = nullptr;
        return P2G_EBADARG;
    }
    return circuit_create_impl(desc, device, rank, world, nullptr, nullptr, out, nccl_id);
}

extern "C" void p2g_circuit_destroy(p2g_circuit* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    delete c;
}

extern "C" int p2g_circuit_cap(const p2g_circuit* c, uint8_t* cap_out, size_t cap_len, uint8_t* digest_out, size_t digest_len) {
    return guard([&] {
        if (!c) throw p2g_error(P2G_EBADARG, "p2g_circuit_cap: null handle");
        size_t need = c->cs_cap.size() * c->hs;
        if (cap_out) {
            if (cap_len < need) throw p2g_error(P2G_ESMALLBUF, "p2g_circuit_cap: cap buffer too small");
            pack_digests(c->h, c->cs_cap.data(), c->cs_cap.size(), cap_out);
        }
        if (digest_out) {
            if (digest_len < (size_t)c->hs) throw p2g_error(P2G_ESMALLBUF, "p2g_circuit_cap: digest buffer too small");
            memcpy(digest_out, &c->digest, c->hs);
        }
    });
}

// VerifierCircuitData::to_bytes(&BackendGateSerializer): plonky2 0.2.2 util/serialization.rs write_verifier_circuit_data =
// write_verifier_only_circuit_data + write_common_circuit_data; usize as u64 LE, bool as u8, field elements as canonical u64 LE.
extern "C" int p2g_vk_bytes(const p2g_circuit* c, const p2g_vk_config* cfg, uint8_t* out, size_t* out_len) {
    return guard([&] {
        if (!c || !out_len) throw p2g_error(P2G_EBADARG, "p2g_vk_bytes: null argument");
        p2g_vk_config k = {};
        k.struct_size = sizeof k;
        k.config_num_constants = 2;
        k.security_bits = 100;
        k.max_quotient_degree_factor = 8;
        k.use_base_arithmetic_gate = 1;
        k.zero_knowledge = 0;
        k.reduction_strategy = 1;
        k.strategy_params[0] = 4;
        k.strategy_params[1] = 5;
        if (cfg) {
            if (cfg->struct_size != sizeof k) throw p2g_error(P2G_EBADARG, "p2g_vk_bytes: p2g_vk_config.struct_size mismatch (ABI)");
            k = *cfg;
        }
        const p2g_circuit_desc& d = c->d;
        Writer w(out, out ? *out_len : 0);
        auto usz = [&](u64 x) { w.u64v(x); };
        auto u32v = [&](u32 x) { w.put(&x, 4); };
        auto fri_config = [&] {
            usz(d.rate_bits);
            usz(d.cap_height);
            usz(d.num_query_rounds);
            u32v(d.pow_bits);
            w.u8v((uint8_t)k.reduction_strategy);
            if (k.reduction_strategy == 0) {
                usz(d.num_fri_layers);
                for (u32 i = 0; i < d.num_fri_layers; i++) usz(d.reduction_arity_bits[i]);
            } else if (k.reduction_strategy == 1) {
                usz(k.strategy_params[0]);
                usz(k.strategy_params[1]);
            } else {
                w.u8v(k.strategy_params[0] ? 1 : 0);
                if (k.strategy_params[0]) usz(k.strategy_params[1]);
            }
        };
        // VerifierOnlyCircuitData
        usz(d.cap_height);
        for (const digest_t& dg : c->cs_cap) w.digest(dg, c->hs);
        w.digest(c->digest, c->hs);
        // CommonCircuitData: config
        usz(d.num_wires);
        usz(d.num_routed_wires);
        usz(k.config_num_constants);
        usz(k.security_bits);
        usz(d.num_challenges);
        usz(k.max_quotient_degree_factor);
        w.u8v(k.use_base_arithmetic_gate ? 1 : 0);
        w.u8v(k.zero_knowledge ? 1 : 0);
        fri_config();
        // fri_params
        fri_config();
        usz(d.num_fri_layers);
        for (u32 i = 0; i < d.num_fri_layers; i++) usz(d.reduction_arity_bits[i]);
        usz(d.degree_bits);
        w.u8v(k.zero_knowledge ? 1 : 0);
        // selectors_info
        usz(d.num_gates);
        for (u32 g = 0; g < d.num_gates; g++) usz(c->gates[g].selector_index);
        usz(d.num_selectors);
        {
            std::vector<std::pair<u32, u32>> groups(d.num_selectors, {0, 0});
            for (u32 g = 0; g < d.num_gates; g++) groups[c->gates[g].selector_index] = {c->gates[g].group_lo, c->gates[g].group_hi};
            for (auto& gr : groups) {
                usz(gr.first);
                usz(gr.second);
            }
        }
        usz(d.quotient_degree_factor);
        usz(d.num_gate_constraints);
        usz(d.num_constants);
        usz(d.num_public_inputs);
        usz(d.num_routed_wires);   // k_is.len()
        for (u32 i = 0; i < d.num_routed_wires; i++) usz(c->k_is[i]);
        usz(d.num_partial_products);
        usz(0);   // num_lookup_polys
        usz(0);   // num_lookup_selectors
        usz(0);   // luts.len()
        // gates: u32 tag = position in BackendGateSerializer's impl_gate_serializer! list (write_vk_action.rs:37-61), then the
        // gate's own serialize payload
        usz(d.num_gates);
        for (u32 g = 0; g < d.num_gates; g++) {
            const p2g_gate& gt = c->gates[g];
            const u32* p = gt.params;
            switch (gt.kind) {
            case P2G_GATE_ARITHMETIC: u32v(0); usz(p[0]); break;
            case P2G_GATE_BASE_SUM:
                if (p[0] != 2 && p[0] != 4) throw p2g_error(P2G_EBADARG, "p2g_vk_bytes: BackendGateSerializer only registers BaseSumGate<2> and <4>");
                u32v(p[0] == 2 ? 2 : 3);
                usz(p[1]);
                break;
            case P2G_GATE_CONSTANT: u32v(4); usz(p[0]); break;
            case P2G_GATE_NOOP: u32v(10); break;
            case P2G_GATE_POSEIDON: u32v(12); break;
            case P2G_GATE_PUBLIC_INPUT: u32v(13); break;
            case P2G_GATE_RANDOM_ACCESS: u32v(14); usz(p[0]); usz(p[1]); usz(p[2]); break;
            case P2G_GATE_COMPARISON: u32v(17); usz(p[0]); usz(p[1]); break;
            case P2G_GATE_U32_ADD_MANY: u32v(18); usz(p[0]); usz(p[1]); break;
            case P2G_GATE_U32_ARITHMETIC: u32v(19); usz(p[0]); break;
            case P2G_GATE_U32_RANGE_CHECK: u32v(20); usz(p[0]); break;
            case P2G_GATE_U32_SUBTRACTION: u32v(21); usz(p[0]); break;
            default: throw p2g_error(P2G_EBADARG, "p2g_vk_bytes: unknown gate kind");
            }
        }
        const size_t need = w.len;
        const bool small = !out || need > w.cap;
        *out_len = need;
        if (small) throw p2g_error(P2G_ESMALLBUF, "p2g_vk_bytes: output buffer too small");
    });
}

extern "C" size_t p2g_proof_size_bound(const p2g_circuit* c) {
    if (!c) return 0;
    const p2g_circuit_desc& d = c->d;
    size_t hs = c->hs, ncap = (size_t)1 << d.cap_height;
    size_t P = d.num_constants + d.num_routed_wires, W = d.num_wires, nzp = d.num_challenges * (1 + d.num_partial_products),
           nq = d.num_challenges * d.quotient_degree_factor;
    size_t sz = 3 * ncap * hs + 16 * (P + W + nzp + d.num_challenges + nq) + d.num_fri_layers * ncap * hs;
    size_t path = 1 + hs * (size_t)c->loglde;
    size_t per_q = 8 * (P + W + nzp + nq) + 4 * path;
    for (u32 l = 0; l < d.num_fri_layers; l++) per_q += 16 * ((size_t)1 << d.reduction_arity_bits[l]) + path;
    sz += d.num_query_rounds * per_q;
    sz += 16 * c->n + 8 + 8 * d.num_public_inputs;
    return sz;
}

// =====================================================================================================================
// prove
// =====================================================================================================================
static void prove_impl(p2g_circuit* C, const u64* d_wires, const u64* public_inputs, size_t n_pi, const u64* forced_pow,
                       uint8_t* out, size_t* out_len, p2g_timings* tm, cudaEvent_t ev_start, bool compressed,
                       bool canonicalize = false) {
    DevCtx* c = C->ctx;
    const p2g_circuit_desc& d = C->d;
    const size_t n = C->n, lde = C->lde;
    const int logn = C->logn, W = d.num_wires, R = d.num_routed_wires, Cc = d.num_constants, NC = d.num_challenges;
    const int NPP = d.num_partial_products, QDF = d.quotient_degree_factor, nchunk = NPP + 1;
    const int nzp = NC * (1 + NPP), nq = NC * QDF, P = Cc + R;
    const int h = C->h, hs = C->hs;
    cudaStream_t st = c->stream;
    cudaEvent_t ev[7];
    for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
    struct EvGuard {
        cudaEvent_t* e;
        ~EvGuard() { for (int i = 0; i < 7; i++) cudaEventDestroy(e[i]); }
    } evg{ev};
    CUDA_CHECK(cudaEventRecord(ev[0], st));
    Trace tr;

    // 1. public inputs hash (InnerHasher = Poseidon for both configs)
    u64 pi_hash[4] = {0, 0, 0, 0};
    if (n_pi) {
        digest_t ph = hash_no_pad(P2G_H_POSEIDON, public_inputs, n_pi);
        memcpy(pi_hash, ph.w, 32);
    }

    // 2. wires commitment
    commit_from_values(C, C->wires, d_wires, n, C->up.active ? &C->up : nullptr);
    int bad_wire = 0;
    {
        // every column this rank reads: its inverse-NTT block and the routed columns (all of them on a single GPU)
        int* flag = ensure(C->ws.flag, 1);
        CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int), st));
        if (C->up.active && C->up.routed) CUDA_CHECK(cudaStreamWaitEvent(st, C->up.routed, 0));
        int c0, c1;
        column_block(C, W, &c0, &c1);
        const int ranges[3][2] = {{c0, c1}, {0, std::min(c0, R)}, {std::max(c1, 0), std::max(c1, R)}};
        for (int ri = 0; ri < 3; ri++) {
            const int* rg = ranges[ri];
            if (rg[1] <= rg[0]) continue;
            // raw GoldilocksField words (p2g_prove_columns): the block was reduced chunk by chunk before its inverse NTT; the routed
            // columns outside the block are reduced here, before the Z computation reads them
            if (ri > 0 && C->up.active) {
                // only rows [row0, row1) of these columns were uploaded (all rows when the rank is not row-sharded)
                const size_t rows = C->row1 - C->row0, cnt2 = rows * (size_t)(rg[1] - rg[0]);
                const unsigned blocks2 = (unsigned)std::min<size_t>((cnt2 + 255) / 256, 148 * 16);
                u64* base = const_cast<u64*>(d_wires) + (size_t)rg[0] * n + C->row0;
                if (canonicalize) k_canonicalize_2d<<<blocks2, 256, 0, st>>>(base, n, rows, rg[1] - rg[0]);
                k_check_canonical_2d<<<blocks2, 256, 0, st>>>(base, n, rows, rg[1] - rg[0], flag);
                count_launch(c, canonicalize ? 2 : 1);
                continue;
            }
            size_t cnt = (size_t)(rg[1] - rg[0]) * n;
            unsigned blocks = (unsigned)std::min<size_t>((cnt + 255) / 256, 148 * 16);
            k_check_canonical<<<blocks, 256, 0, st>>>(d_wires + (size_t)rg[0] * n, cnt, flag);
            count_launch(c);
        }
        CUDA_CHECK(cudaMemcpyAsync(&bad_wire, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    std::vector<digest_t> wires_cap = read_cap(C, C->wires.tree, true, &bad_wire);   // synchronises the stream; flag OR-ed over ranks
    if (bad_wire) throw p2g_error(P2G_EBADARG, "p2g_prove: non-canonical wire value (>= p)");
    CUDA_CHECK(cudaEventRecord(ev[1], st));
    tr.mark("wires commit");

    // 3-4. transcript
    challenger_t ch;
    ch.init(h);
    ch.observe_digest(C->digest);
    ch.observe_many(pi_hash, 4);
    for (auto& dg : wires_cap) ch.observe_digest(dg);
    u64 betas[2] = {0, 0}, gammas[2] = {0, 0}, alphas[2] = {0, 0};
    for (int i = 0; i < NC; i++) betas[i] = ch.get();
    for (int i = 0; i < NC; i++) gammas[i] = ch.get();

    // 5. Z and partial products
    {
        ZppParams zp = {};
        zp.logn = logn;
        zp.num_routed = R;
        zp.chunk = QDF;
        zp.nchunk = nchunk;
        zp.num_challenges = NC;
        for (int cc = 0; cc < NC; cc++) {
            zp.betas[cc] = betas[cc];
            zp.gammas[cc] = gammas[cc];
            for (int r = 0; r < R; r++) zp.beta_k[cc][r] = gl_mul(betas[cc], C->k_is[r]);
        }
        int xsplit;
        const u64* xtab = c->get_powtab(logn, gl_root_of_unity(logn), 1, &xsplit);
        // rows [row0, row1) only: a sharded rank computes its row block; the running product enters each block through the
        // all-gathered per-rank block products, and the 20 value columns are completed by a row all-gather
        const size_t rows = C->row1 - C->row0;
        const size_t nt = (rows + SCAN_CH - 1) / SCAN_CH, t0 = C->row0 / SCAN_CH;
        u64* q = ensure(C->ws.q, (size_t)NC * nchunk * n);
        u64* totals = ensure(C->ws.totals, 4 * ((n + SCAN_CH - 1) / SCAN_CH));
        u64* rank_totals = ensure(C->ws.rank_totals, (size_t)C->world * NC);
        dim3 g1((unsigned)((rows + 127) / 128), NC);
        if (C->up.active && C->up.routed) CUDA_CHECK(cudaStreamWaitEvent(st, C->up.routed, 0));
        k_zpp_chunks<<<g1, 128, 0, st>>>(zp, d_wires, n, C->sigma_values.p, xtab, xsplit, q, C->row0, C->row1);
        dim3 g2((unsigned)((nt + 127) / 128), NC);
        k_zscan_totals<<<g2, 128, 0, st>>>(zp, q, totals, nt, t0);
        k_scan_mul_excl<<<NC, 1024, 0, st>>>(totals, nt, rank_totals + (size_t)C->rank * NC);
        if (C->row_shard) shard_allgather(C, rank_totals + (size_t)C->rank * NC, rank_totals, (size_t)NC * 8, true);
        k_zscan_apply<<<g2, 128, 0, st>>>(zp, q, totals, nt, C->zpp_values.p, t0, rank_totals, C->row_shard ? C->rank : 0);
        count_launch(c, 4);
        CUDA_CHECK(cudaGetLastError());
        if (C->row_shard) shard_allgather_cols(C, C->zpp_values.p, nzp, n, rows);
        tr.mark("zpp launches");
        commit_from_values(C, C->zpp, C->zpp_values.p, n);
        tr.mark("zpp commit launched");
    }
    std::vector<digest_t> zpp_cap = read_cap(C, C->zpp.tree);
    for (auto& dg : zpp_cap) ch.observe_digest(dg);
    CUDA_CHECK(cudaEventRecord(ev[2], st));
    tr.mark("zpp cap + observe");

    // 6. alphas
    for (int i = 0; i < NC; i++) alphas[i] = ch.get();

    // 7. quotient
    {
        static thread_local QuotientParams qp;
        memset(&qp, 0, sizeof(qp));
        qp.logn = logn;
        qp.rate_bits = d.rate_bits;
        qp.num_challenges = NC;
        qp.num_partial_products = NPP;
        qp.num_routed = R;
        qp.num_constants = Cc;
        qp.num_selectors = d.num_selectors;
        qp.qdf = QDF;
        for (int cc = 0; cc < NC; cc++) {
            qp.betas[cc] = betas[cc];
            qp.gammas[cc] = gammas[cc];
            for (int r = 0; r < R; r++) qp.beta_k[cc][r] = gl_mul(betas[cc], C->k_is[r]);
            // alpha^k for every constraint term, and at least up to num_wires (the limb sweep weights wire w by alpha^w)
            int nterms = std::min<int>(P2G_MAX_TERMS, std::max<int>(NC * (2 + NPP) + d.num_gate_constraints, W));
            u64 a = 1;
            for (int k = 0; k < nterms; k++) {
                qp.apow[cc][k] = a;
                a = gl_mul(a, alphas[cc]);
            }
        }
        for (u32 r = 0; r < (1u << d.rate_bits); r++) qp.zh_inv[r] = C->zh_inv[r];
        memcpy(qp.pi_hash, pi_hash, 32);
        fill_gates(qp, C);
        // quotient values land in quot.lde's first NC columns region?  No: they are transformed in place into the
        // chunk coefficients, so evaluate straight into quot.coeffs viewed as [NC][8N]
        u64* qv = C->quot.coeffs.p;
        {
            StageTimer tq(c, &c->quot_ms);
            static thread_local LimbPlan lp;
            const bool sweep = quotient_limb_plan(qp, W, alphas, &lp);
            quotient_eval(c, qp, sweep ? &lp : nullptr, C->cs.lde.p, C->wires.lde.p, C->zpp.lde.p, C->xs.p, C->l0s.p, qv, C->lde_l, C->j0, lde);
        }
        if (C->world > 1)   // every rank needs all 8N quotient values for the size-8N inverse transform
            for (int cc = 0; cc < NC; cc++)
                shard_allgather(C, qv + (size_t)cc * lde + C->j0, qv + (size_t)cc * lde, C->lde_l * 8, true);
        ntt_coset_ifft_leaforder(c, qv, lde, C->loglde, NC, GL_GEN);
        // [NC][8N] coefficients == [NC*QDF][N] chunk polynomials
        commit_from_coeffs(C, C->quot);
    }
    std::vector<digest_t> quot_cap = read_cap(C, C->quot.tree);
    for (auto& dg : quot_cap) ch.observe_digest(dg);
    CUDA_CHECK(cudaEventRecord(ev[3], st));
    tr.mark("quotient");

    // 8. zeta
    e2 zeta = ch.get_e2();
    if (e2_eq(e2_pow(zeta, (u64)n), e2_make(1, 0))) throw p2g_error(P2G_EUNSAT, "Opening point is in the subgroup.");
    if (zeta.c0 == 0 && zeta.c1 == 0) throw p2g_error(P2G_EUNSAT, "Opening point is zero.");
    const u64 g = gl_root_of_unity(logn);
    e2 zeta_next = e2_mul_base(zeta, g);

    // 9. openings
    PolyBatch* oracles[4] = {&C->cs, &C->wires, &C->zpp, &C->quot};
    const int widths[4] = {P, W, nzp, nq};
    const int total = P + W + nzp + nq;
    u64* ztab0 = ensure(C->ws.ztab0, 2 * n);
    u64* ztab1 = ensure(C->ws.ztab1, 2 * n);
    k_e2_powers<<<(unsigned)((n / POW_RUN + 128) / 128), 128, 0, st>>>(ztab0, n, zeta);
    k_e2_powers<<<(unsigned)((n / POW_RUN + 128) / 128), 128, 0, st>>>(ztab1, n, zeta_next);
    count_launch(c, 2);
    const int nsplit = (int)((n + EVAL_SPLIT - 1) / EVAL_SPLIT);
    std::vector<e2> op(total), zs_next(NC);
    {
        // evaluation g of the opening set: polynomial g of the four oracles at zeta (g < total), Z_{g - total} at g zeta.  A sharded
        // rank evaluates the slice [g0, g1) -- every rank holds every coefficient -- and the values are all-gathered.
        const int nev = total + NC;
        const bool split_evals = C->row_shard;
        const int per = split_evals ? (nev + C->world - 1) / C->world : nev;
        const int g0 = split_evals ? std::min(nev, C->rank * per) : 0, g1 = std::min(nev, g0 + per);
        e2* partial = ensure(C->ws.partial, (size_t)nev * nsplit);
        int off = 0;
        for (int o = 0; o < 5; o++) {
            const int width = o < 4 ? widths[o] : NC;
            const int lo = std::max(g0, off), hi = std::min(g1, off + width);
            if (hi > lo) {
                const u64* cf = (o < 4 ? oracles[o]->coeffs.p : C->zpp.coeffs.p) + (size_t)(lo - off) * n;
                dim3 grid(hi - lo, nsplit);
                k_eval_polys<<<grid, 256, 0, st>>>(cf, n, n, o < 4 ? ztab0 : ztab1, partial + (size_t)lo * nsplit, nsplit);
                count_launch(c);
            }
            off += width;
        }
        CUDA_CHECK(cudaGetLastError());
        std::vector<e2> hp((size_t)std::max(per, 1) * nsplit), vals((size_t)per * (split_evals ? C->world : 1));
        if (g1 > g0) CUDA_CHECK(cudaMemcpyAsync(hp.data(), partial + (size_t)g0 * nsplit, (size_t)(g1 - g0) * nsplit * sizeof(e2), cudaMemcpyDeviceToHost, st));
        stream_sync(C);
        for (int i = 0; i < g1 - g0; i++) {
            e2 sacc = e2_make(0, 0);
            for (int k = 0; k < nsplit; k++) sacc = e2_add(sacc, hp[(size_t)i * nsplit + k]);
            vals[(size_t)(split_evals ? C->rank * per : 0) + i] = sacc;
        }
        if (split_evals) shard_allgather(C, vals.data() + (size_t)C->rank * per, vals.data(), (size_t)per * sizeof(e2), false);
        for (int i = 0; i < nev; i++) {
            if (i < total) op[i] = vals[i];
            else zs_next[i - total] = vals[i];
        }
    }
    for (int i = 0; i < total; i++) ch.observe_e2(op[i]);
    for (int i = 0; i < NC; i++) ch.observe_e2(zs_next[i]);
    CUDA_CHECK(cudaEventRecord(ev[4], st));
    tr.mark("openings");

    // 10. FRI
    e2 fri_alpha = ch.get_e2();
    const int nl = d.num_fri_layers;
    u64* fin = ensure(C->ws.fin, 2 * n);  // final polynomial, re | im
    {
        std::vector<e2> apow(total);
        e2 a = e2_make(1, 0);
        for (int j = 0; j < total; j++) {
            apow[j] = a;
            a = e2_mul(a, fri_alpha);
        }
        e2* d_apow = ensure(C->ws.apow, total);
        CUDA_CHECK(cudaMemcpyAsync(d_apow, apow.data(), sizeof(e2) * total, cudaMemcpyHostToDevice, st));
        CombineArgs ca = {};
        for (int o = 0; o < 4; o++) {
            ca.coeffs[o] = oracles[o]->coeffs.p;
            ca.ncols[o] = widths[o];
            ca.cs[o] = n;
        }
        ca.num_challenges = NC;
        ca.logn = logn;
        u64* u = ensure(C->ws.u, 4 * n);
        k_fri_combine<<<(unsigned)((C->row1 - C->row0 + 127) / 128), 128, 0, st>>>(ca, d_apow, ztab0, ztab1, u, C->row0, C->row1);
        if (C->row_shard) shard_allgather_cols(C, u, 4, n, C->row1 - C->row0);
        // tables of inverse powers (reuse the forward tables' storage after the combine)
        e2 zi0, zi1;  // inverse in F_{p^2}: z^-1 = conj(z) / norm(z)
        {
            auto inv2 = [](e2 z) {
                u64 nrm = gl_sub(gl_sqr(z.c0), gl_mul_small(gl_sqr(z.c1), 7));
                u64 ni = gl_inv(nrm);
                return e2_make(gl_mul(z.c0, ni), gl_neg(gl_mul(z.c1, ni)));
            };
            zi0 = inv2(zeta);
            zi1 = inv2(zeta_next);
        }
        k_e2_powers<<<(unsigned)((n / POW_RUN + 128) / 128), 128, 0, st>>>(ztab0, n, zi0);
        k_e2_powers<<<(unsigned)((n / POW_RUN + 128) / 128), 128, 0, st>>>(ztab1, n, zi1);
        size_t nt = (n + SCAN_CH - 1) / SCAN_CH;
        u64* totals = ensure(C->ws.totals, 4 * nt);
        dim3 g2((unsigned)((nt + 127) / 128), 4);
        k_sscan_totals<<<g2, 128, 0, st>>>(u, totals, n, nt);
        k_scan_add_suffix_excl<<<4, 1024, 0, st>>>(totals, nt);
        e2 alpha_nc = e2_pow(fri_alpha, (u64)NC);
        k_sscan_apply<<<(unsigned)((nt + 127) / 128), 128, 0, st>>>(u, totals, n, nt, ztab0, ztab1, alpha_nc, fin);
        count_launch(c, 6);
        CUDA_CHECK(cudaGetLastError());
        stream_sync(C);
    }
    tr.mark("fri combine");
    // commit phase
    typedef p2g_circuit::FriLayer Layer;
    std::vector<Layer>& layers = C->ws.layers;
    if ((int)layers.size() != nl) layers.resize(nl);
    std::vector<e2> fri_betas(nl);
    C->fri_caps.clear();
    u64* coeffs_a = ensure(C->ws.coeffs_a, 2 * n);
    u64* coeffs_b = nullptr;
    CUDA_CHECK(cudaMemcpyAsync(coeffs_a, fin, 16 * n, cudaMemcpyDeviceToDevice, st));
    u64* cur_coeffs = coeffs_a;
    size_t m = n;  // number of (possibly) non-zero coefficients
    int logm = logn;
    u64 shift = GL_GEN;
    if (nl) coeffs_b = ensure(C->ws.coeffs_b, 2 * (n >> d.reduction_arity_bits[0]) + 2);
    for (int l = 0; l < nl; l++) {
        Layer& L = layers[l];
        L.ab = d.reduction_arity_bits[l];
        L.logcur = logm + d.rate_bits;
        L.cur = (size_t)1 << L.logcur;
        ensure(L.values, 2 * L.cur);
        // values = coset_fft(coeffs.lde(rate_bits), shift), both components, leaf order
        ntt_lde(c, cur_coeffs, m, L.values.p, L.cur, logm, d.rate_bits, 2, shift);
        const int arity = 1 << L.ab;
        merkle_build(c, &L.tree.tree, L.values.p, L.cur, L.logcur - L.ab, 2 * arity, d.cap_height, h, true);
        upload_level_ptrs(c, L.tree);
        std::vector<digest_t> cap = read_cap(C, L.tree.tree, false);   // FRI layers are replicated on every rank
        for (auto& dg : cap) ch.observe_digest(dg);
        C->fri_caps.insert(C->fri_caps.end(), cap.begin(), cap.end());
        e2 beta = ch.get_e2();
        fri_betas[l] = beta;
        size_t m2 = m >> L.ab;
        u64* nxt = (cur_coeffs == coeffs_a) ? coeffs_b : coeffs_a;
        k_fri_fold<<<(unsigned)((m2 + 127) / 128), 128, 0, st>>>(cur_coeffs, m, nxt, m2, arity, beta);
        count_launch(c);
        CUDA_CHECK(cudaGetLastError());
        cur_coeffs = nxt;
        m = m2;
        logm -= L.ab;
        shift = gl_pow(shift, (u64)arity);
    }
    // final polynomial
    C->final_poly.assign(2 * m, 0);
    {
        std::vector<u64> tmp(2 * m);
        CUDA_CHECK(cudaMemcpyAsync(tmp.data(), cur_coeffs, 8 * m, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaMemcpyAsync(tmp.data() + m, cur_coeffs + m, 8 * m, cudaMemcpyDeviceToHost, st));
        stream_sync(C);
        for (size_t i = 0; i < m; i++) {
            C->final_poly[2 * i] = tmp[i];
            C->final_poly[2 * i + 1] = tmp[m + i];
        }
    }
    for (size_t i = 0; i < m; i++) ch.observe_e2(e2_make(C->final_poly[2 * i], C->final_poly[2 * i + 1]));
    tr.mark("fri commit phase");

    // proof of work
    u64 pow_witness = 0;
    if (forced_pow) {
        pow_witness = *forced_pow;
    } else {
        unsigned long long* best = ensure(C->ws.best, 1);
        // 2^(pow_bits + 2) candidates per launch: the smallest witness is in the first batch with probability 1 - e^-4 (98 %);
        // a 2^20 batch hashed 16 x more candidates than needed at pow_bits = 16 (0.8 ms per proof)
        const u64 batch = (u64)1 << std::min<u32>(20, std::max<u32>(12, d.pow_bits + 2));
        for (u64 start = 0;; start += batch) {
            CUDA_CHECK(cudaMemsetAsync(best, 0xff, 8, st));
            k_pow_search<<<(unsigned)(batch / 128), 128, 0, st>>>(ch, start, batch, d.pow_bits, best);
            count_launch(c);
            unsigned long long hb = 0;
            CUDA_CHECK(cudaMemcpyAsync(&hb, best, 8, cudaMemcpyDeviceToHost, st));
            stream_sync(C);
            if (hb != ~0ULL) {
                pow_witness = hb;
                break;
            }
            if (start > ((u64)1 << 40)) throw p2g_error(P2G_EUNSAT, "proof-of-work search exhausted");
        }
    }
    ch.observe(pow_witness);
    {
        u64 resp = ch.get();
        if (d.pow_bits && (resp >> (64 - d.pow_bits)) != 0) throw p2g_error(P2G_EUNSAT, "forced pow_witness is invalid");
    }
    const int NQ = d.num_query_rounds;
    std::vector<u32> indices(NQ);
    for (int q = 0; q < NQ; q++) indices[q] = (u32)(ch.get() % (u64)lde);
    CUDA_CHECK(cudaEventRecord(ev[5], st));
    tr.mark("pow + indices");

    // challenges dump (P2G_BUF_CHALLENGES)
    C->challenges.clear();
    for (int i = 0; i < NC; i++) C->challenges.push_back(betas[i]);
    for (int i = 0; i < NC; i++) C->challenges.push_back(gammas[i]);
    for (int i = 0; i < NC; i++) C->challenges.push_back(alphas[i]);
    C->challenges.push_back(zeta.c0);
    C->challenges.push_back(zeta.c1);
    C->challenges.push_back(fri_alpha.c0);
    C->challenges.push_back(fri_alpha.c1);
    for (int l = 0; l < nl; l++) {
        C->challenges.push_back(fri_betas[l].c0);
        C->challenges.push_back(fri_betas[l].c1);
    }
    C->challenges.push_back(pow_witness);
    for (int q = 0; q < NQ; q++) C->challenges.push_back(indices[q]);

    // query rounds: gather opened rows and Merkle paths on the device, one D2H
    // d_idx: global leaf indices (FRI layers, replicated); d_lidx: indices into this rank's leaves (0 for queries it does not own)
    u32* d_idx = ensure(C->ws.idx, 2 * NQ);
    u32* d_lidx = d_idx + NQ;
    std::vector<u32> idx2(2 * NQ);
    for (int q = 0; q < NQ; q++) {
        idx2[q] = indices[q];
        bool mine = indices[q] >= C->j0 && indices[q] < C->j0 + C->lde_l;
        idx2[NQ + q] = mine ? (u32)(indices[q] - C->j0) : 0;
    }
    CUDA_CHECK(cudaMemcpyAsync(d_idx, idx2.data(), 8 * NQ, cudaMemcpyHostToDevice, st));
    size_t row_words = 0, path_digests = 0;
    size_t row_off[4 + P2G_MAX_FRI_LAYERS], path_off[4 + P2G_MAX_FRI_LAYERS];
    int path_len[4 + P2G_MAX_FRI_LAYERS];
    for (int o = 0; o < 4; o++) {
        row_off[o] = row_words;
        row_words += (size_t)NQ * widths[o];
        path_off[o] = path_digests;
        path_len[o] = (int)oracles[o]->tree.levels.size() - 1;
        path_digests += (size_t)NQ * path_len[o];
    }
    for (int l = 0; l < nl; l++) {
        row_off[4 + l] = row_words;
        row_words += (size_t)NQ * 2 * (1 << layers[l].ab);
        path_off[4 + l] = path_digests;
        path_len[4 + l] = (int)layers[l].tree.tree.levels.size() - 1;
        path_digests += (size_t)NQ * path_len[4 + l];
    }
    u64* d_rows = ensure(C->ws.rows, row_words);
    digest_t* d_paths = ensure(C->ws.paths, path_digests + 1);
    for (int o = 0; o < 4; o++) {
        int cnt = NQ * widths[o];
        k_gather_rows<<<(cnt + 127) / 128, 128, 0, st>>>(oracles[o]->lde.p, C->lde_l, widths[o], d_lidx, NQ, d_rows + row_off[o]);
        if (path_len[o] > 0) {
            int pc = NQ * path_len[o];
            k_gather_paths<<<(pc + 127) / 128, 128, 0, st>>>(oracles[o]->d_levels.p, path_len[o], d_lidx, NQ, 0, d_paths + path_off[o]);
        }
        count_launch(c, 2);
    }
    {
        int sh = 0;
        for (int l = 0; l < nl; l++) {
            sh += layers[l].ab;
            int arity = 1 << layers[l].ab;
            int cnt = NQ * 2 * arity;
            k_gather_fri_rows<<<(cnt + 127) / 128, 128, 0, st>>>(layers[l].values.p, layers[l].cur, arity, d_idx, NQ, sh,
                                                                 d_rows + row_off[4 + l]);
            if (path_len[4 + l] > 0) {
                int pc = NQ * path_len[4 + l];
                k_gather_paths<<<(pc + 127) / 128, 128, 0, st>>>(layers[l].tree.d_levels.p, path_len[4 + l], d_idx, NQ, sh,
                                                                 d_paths + path_off[4 + l]);
            }
            count_launch(c, 2);
        }
    }
    CUDA_CHECK(cudaGetLastError());
    std::vector<u64> h_rows(row_words);
    std::vector<digest_t> h_paths(path_digests + 1);
    CUDA_CHECK(cudaMemcpyAsync(h_rows.data(), d_rows, 8 * row_words, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(h_paths.data(), d_paths, sizeof(digest_t) * path_digests, cudaMemcpyDeviceToHost, st));
    stream_sync(C);
    if (C->world > 1) {
        // opened rows and paths of the four committed oracles come from the rank that owns the leaf
        const size_t rw = (size_t)NQ * total, pd = path_off[3] + (size_t)NQ * path_len[3];
        const size_t bytes = 8 * rw + sizeof(digest_t) * pd;
        std::vector<uint8_t> mine(bytes), all(bytes * C->world);
        memcpy(mine.data(), h_rows.data(), 8 * rw);
        memcpy(mine.data() + 8 * rw, h_paths.data(), sizeof(digest_t) * pd);
        shard_allgather(C, mine.data(), all.data(), bytes, false);
        for (int q = 0; q < NQ; q++) {
            const int owner = (int)(indices[q] / C->lde_l);
            const uint8_t* src = all.data() + (size_t)owner * bytes;
            for (int o = 0; o < 4; o++) {
                memcpy(h_rows.data() + row_off[o] + (size_t)q * widths[o], src + 8 * (row_off[o] + (size_t)q * widths[o]), 8 * (size_t)widths[o]);
                memcpy(h_paths.data() + path_off[o] + (size_t)q * path_len[o],
                       src + 8 * rw + sizeof(digest_t) * (path_off[o] + (size_t)q * path_len[o]), sizeof(digest_t) * (size_t)path_len[o]);
            }
        }
    }

    // serialise: Proof || public_inputs   (App. A.12, uncompressed)
    Writer wb(out, out ? *out_len : 0);
    for (auto& dg : wires_cap) wb.digest(dg, hs);
    for (auto& dg : zpp_cap) wb.digest(dg, hs);
    for (auto& dg : quot_cap) wb.digest(dg, hs);
    {
        // OpeningSet: constants | sigmas | wires | zs | zs_next | partial_products | quotient
        const int o_w = P, o_z = P + W, o_q = P + W + nzp;
        for (int i = 0; i < o_w; i++) wb.e2v(op[i]);
        for (int i = 0; i < W; i++) wb.e2v(op[o_w + i]);
        for (int i = 0; i < NC; i++) wb.e2v(op[o_z + i]);
        for (int i = 0; i < NC; i++) wb.e2v(zs_next[i]);
        for (int i = NC; i < nzp; i++) wb.e2v(op[o_z + i]);
        for (int i = 0; i < nq; i++) wb.e2v(op[o_q + i]);
    }
    for (auto& dg : C->fri_caps) wb.digest(dg, hs);
    if (!compressed) {
        for (int q = 0; q < NQ; q++) {
            for (int o = 0; o < 4 + nl; o++) {
                size_t wdt = o < 4 ? (size_t)widths[o] : (size_t)2 * (1 << layers[o - 4].ab);
                wb.put(h_rows.data() + row_off[o] + (size_t)q * wdt, 8 * wdt);
                wb.u8v((uint8_t)path_len[o]);
                for (int k = 0; k < path_len[o]; k++) wb.digest(h_paths[path_off[o] + (size_t)q * path_len[o] + k], hs);
            }
        }
    } else {
        // CompressedFriProof (plonky2 fri/proof.rs FriProof::compress + hash/path_compression.rs compress_merkle_proofs):
        // query indices, then per DISTINCT index (ascending) the opened rows with the Merkle siblings no other opened path
        // already determines, then per FRI layer per distinct coset index the evaluations minus the one the verifier infers.
        for (int q = 0; q < NQ; q++) {
            u32 ix = indices[q];
            wb.put(&ix, 4);
        }
        // keep[o][q][k]: sibling k of query q's path in tree o is written
        std::vector<std::vector<std::vector<char>>> keep(4 + nl);
        std::vector<std::vector<u64>> tree_idx(4 + nl, std::vector<u64>(NQ));
        {
            int sh = 0;
            for (int o = 0; o < 4 + nl; o++) {
                if (o >= 4) sh += layers[o - 4].ab;
                const int plen = path_len[o];
                // node ids as in plonky2: (leaf + 2^H) >> j over the full tree of height H
                std::map<u64, bool> kn;
                const u64 nlv = (u64)1 << (o < 4 ? C->loglde : layers[o - 4].logcur - layers[o - 4].ab);
                for (int q = 0; q < NQ; q++) {
                    tree_idx[o][q] = (u64)indices[q] >> sh;
                    for (int j = 0; j < plen; j++) kn[(tree_idx[o][q] + nlv) >> j] = true;
                }
                keep[o].assign(NQ, std::vector<char>(plen, 0));
                for (int q = 0; q < NQ; q++) {
                    u64 node = tree_idx[o][q] + nlv;
                    for (int k = 0; k < plen; k++) {
                        u64 sib = node ^ 1;
                        if (!kn.count(sib)) {
                            keep[o][q][k] = 1;
                            kn[sib] = true;
                        }
                        node >>= 1;
                    }
                }
            }
        }
        auto write_path = [&](int o, int q) {
            int cnt = 0;
            for (int k = 0; k < path_len[o]; k++) cnt += keep[o][q][k];
            wb.u8v((uint8_t)cnt);
            for (int k = 0; k < path_len[o]; k++)
                if (keep[o][q][k]) wb.digest(h_paths[path_off[o] + (size_t)q * path_len[o] + k], hs);
        };
        {   // initial trees: first query of every distinct index, ascending
            std::map<u64, int> first;
            for (int q = NQ - 1; q >= 0; q--) first[indices[q]] = q;
            for (auto& kv : first) {
                const int q = kv.second;
                for (int o = 0; o < 4; o++) {
                    wb.put(h_rows.data() + row_off[o] + (size_t)q * widths[o], 8 * (size_t)widths[o]);
                    write_path(o, q);
                }
            }
        }
        for (int l = 0; l < nl; l++) {
            const int o = 4 + l, arity = 1 << layers[l].ab;
            std::map<u64, int> first;
            for (int q = NQ - 1; q >= 0; q--) first[tree_idx[o][q]] = q;
            for (auto& kv : first) {
                const int q = kv.second;
                const u64 prev = l == 0 ? (u64)indices[q] : tree_idx[o - 1][q];
                const int within = (int)(prev & (u64)(arity - 1));
                const u64* ev = h_rows.data() + row_off[o] + (size_t)q * 2 * arity;
                for (int i = 0; i < arity; i++)
                    if (i != within) wb.put(ev + 2 * i, 16);
                write_path(o, q);
            }
        }
    }
    for (size_t i = 0; i < m; i++) {
        wb.u64v(C->final_poly[2 * i]);
        wb.u64v(C->final_poly[2 * i + 1]);
    }
    wb.u64v(pow_witness);
    for (size_t i = 0; i < n_pi; i++) wb.u64v(public_inputs[i]);
    CUDA_CHECK(cudaEventRecord(ev[6], st));
    stream_sync(C);
    tr.mark("queries + serialise");
    C->last_caps[0] = wires_cap;
    C->last_caps[1] = zpp_cap;
    C->last_caps[2] = quot_cap;
    C->proved = true;

    resolve_timers(c);
    if (tm) {
        float t;
        cudaEventElapsedTime(&t, ev[0], ev[1]); tm->wires_commit_ms = t;
        cudaEventElapsedTime(&t, ev[1], ev[2]); tm->zs_pp_ms = t;
        cudaEventElapsedTime(&t, ev[2], ev[3]); tm->quotient_ms = t;
        cudaEventElapsedTime(&t, ev[3], ev[4]); tm->openings_ms = t;
        cudaEventElapsedTime(&t, ev[4], ev[5]); tm->fri_ms = t;
        cudaEventElapsedTime(&t, ev[5], ev[6]); tm->d2h_ms = t;
        cudaEventElapsedTime(&t, ev_start ? ev_start : ev[0], ev[6]); tm->total_ms = t;
        if (ev_start && C->up.last) { cudaEventElapsedTime(&t, ev_start, C->up.last); tm->h2d_ms = t; } else tm->h2d_ms = 0;
        tm->quotient_kernel_ms = c->quot_ms;
        tm->ntt_ms = c->ntt_ms;
        tm->merkle_ms = c->merkle_ms;
        tm->ntt_bytes = c->ntt_bytes;
        tm->merkle_bytes = c->merkle_bytes;
        tm->kernel_launches = c->launches;
        tm->leaf_hash_launches = c->leaf_launches;
        tm->leaf_hash_ms = c->leaf_ms;
        tm->lde_ms = c->lde_ms;
        tm->leaf_hash_bytes = c->leaf_bytes;
        tm->lde_bytes = c->lde_bytes;
        tm->lde_launches = c->lde_launches;
        tm->h2d_bytes = C->up.active ? C->up.bytes : 0;
    }
    size_t need = wb.len;
    bool small = !out || need > wb.cap;
    *out_len = need;
    if (small) throw p2g_error(P2G_ESMALLBUF, "output buffer too small");
}

// wires: one [W][N] block (host or device), or cols: W pointers to N-element host columns (MatrixWitness.wire_values as plonky2
// holds it) -- exactly one of the two
static int prove_entry(p2g_circuit* C, const u64* wires, bool on_device, const u64* public_inputs, size_t n_pi,
                       const u64* forced_pow, uint8_t* out, size_t* out_len, p2g_timings* tm, bool compressed = false,
                       const u64* const* cols = nullptr, bool routed_only = false) {
    return guard([&] {
        if (!C || (!wires && !cols) || !out_len || (n_pi && !public_inputs)) throw p2g_error(P2G_EBADARG, "p2g_prove: null argument");
        // routed_only: cols holds the num_routed_wires routed columns; the advice columns are computed on the device
        const int cols_given = routed_only ? (int)C->d.num_routed_wires : (int)C->d.num_wires;
        if (routed_only && (!cols || C->world != 1))
            throw p2g_error(P2G_EBADARG, "p2g_prove_routed_columns: needs column pointers and a single-GPU handle (sharded: "
                                         "p2g_fill_advice_device + p2g_prove_device)");
        if (cols)
            for (int i = 0; i < cols_given; i++)
                if (!cols[i]) throw p2g_error(P2G_EBADARG, "p2g_prove_columns: null column pointer");
        if (n_pi != C->d.num_public_inputs) throw p2g_error(P2G_EBADARG, "p2g_prove: public input count");
        for (size_t i = 0; i < n_pi; i++)
            if (public_inputs[i] >= GL_P) throw p2g_error(P2G_EBADARG, "p2g_prove: non-canonical public input");
        if (!out) {   // size query: answered from the bound, nothing is proved
            *out_len = p2g_proof_size_bound(C);
            throw p2g_error(P2G_ESMALLBUF, "p2g_prove: out is NULL; *out_len holds a sufficient buffer size");
        }
        std::lock_guard<std::mutex> lk(C->mu);
        DevCtx* c = C->ctx;
        CUDA_CHECK(cudaSetDevice(c->device));
        resolve_timers(c);   // stage events left behind by a proof that failed half way must not leak into this one's totals
        c->launches = 0;
        c->ntt_bytes = c->merkle_bytes = c->leaf_bytes = c->lde_bytes = 0;
        c->ntt_ms = c->merkle_ms = c->leaf_ms = c->lde_ms = c->quot_ms = 0;
        c->leaf_launches = c->lde_launches = 0;
        c->timing = tm != nullptr;
        if (tm) memset(tm, 0, sizeof(*tm));
        struct TimingOff {
            DevCtx* c;
            ~TimingOff() { c->timing = false; }
        } toff{c};
        const u64* d_wires = wires;
        cudaEvent_t ev_start = nullptr;
        p2g_circuit::Upload& up = C->up;
        up.active = false;
        if (!on_device) {
            // Upload in column chunks on a copy stream; the inverse NTT of chunk k overlaps the transfer of chunk k+1.
            // A sharded rank uploads only what it reads: its inverse-NTT column block and the routed columns (Z computation).
            const int W = C->d.num_wires, R = C->d.num_routed_wires;
            const size_t n = C->n;
            size_t cnt = (size_t)W * n;
            if (C->wires_values.n != cnt) C->wires_values.alloc(cnt);
            if (!up.copy) CUDA_CHECK(cudaStreamCreateWithFlags(&up.copy, cudaStreamNonBlocking));
            size_t used = 0;
            auto next_event = [&]() {
                if (used == up.pool.size()) {
                    cudaEvent_t e;
                    CUDA_CHECK(cudaEventCreate(&e));
                    up.pool.push_back(e);
                }
                return up.pool[used++];
            };
            // is the caller's buffer page-locked?  (cudaHostAlloc / p2g_host_alloc / cudaHostRegister)
            bool pinned = true;
            for (int i = 0; i < (cols ? cols_given : 1); i++) {
                cudaPointerAttributes attr;
                if (cudaPointerGetAttributes(&attr, cols ? cols[i] : wires) == cudaSuccess) {
                    if (attr.type == cudaMemoryTypeDevice)
                        throw p2g_error(P2G_EBADARG, "p2g_prove: wires is device memory; use p2g_prove_device");
                    pinned = pinned && attr.type == cudaMemoryTypeHost;
                } else {
                    cudaGetLastError();
                    pinned = false;
                }
            }
            auto col_ptr = [&](int col) { return cols ? cols[col] : wires + (size_t)col * n; };
            int slot = 0;
            int step = std::max(1, std::min(32, (int)(((size_t)64 << 20) / (n * 8)) + 1));   // >= 64 MB per chunk ...
            // rows [r0, r1) of columns [a, e) -> the same place of the device staging matrix
            std::function<void(int, int, size_t, size_t)> copy_block = [&](int a, int e, size_t r0, size_t r1) {
                const size_t nr = r1 - r0;
                const int fit = (int)std::max<size_t>(1, ((size_t)step * n) / nr);   // columns of nr rows per staging buffer
                if (e - a > fit && !pinned) {   // staging buffers hold one chunk: split longer runs
                    for (int x = a; x < e; x += fit) copy_block(x, std::min(e, x + fit), r0, r1);
                    return;
                }
                const size_t words = (size_t)(e - a) * nr;
                u64* d_dst = C->wires_values.p + (size_t)a * n + r0;
                if (!pinned && words * 8 >= ((size_t)1 << 20)) {
                    // stage through pinned memory: wait for the slot's previous DMA, fill it with 8 threads, DMA from it
                    if (up.stage_words < words) {
                        CUDA_CHECK(cudaStreamSynchronize(up.copy));
                        for (int i = 0; i < p2g_circuit::Upload::NSTAGE; i++) {
                            if (up.stage[i]) cudaFreeHost(up.stage[i]);
                            up.stage[i] = nullptr;
                            CUDA_CHECK(cudaHostAlloc((void**)&up.stage[i], words * 8, cudaHostAllocDefault));
                            if (!up.stage_free[i]) CUDA_CHECK(cudaEventCreateWithFlags(&up.stage_free[i], cudaEventDisableTiming));
                            CUDA_CHECK(cudaEventRecord(up.stage_free[i], up.copy));
                        }
                        up.stage_words = words;
                    }
                    CUDA_CHECK(cudaEventSynchronize(up.stage_free[slot]));
                    u64* dst = up.stage[slot];
                    // gather threads: up to 16 (P2G_STAGE_THREADS overrides); the copy is bound by host memory bandwidth
                    static const int T = [] {
                        const char* e = getenv("P2G_STAGE_THREADS");
                        int t = e ? atoi(e) : (int)std::min(16u, std::max(1u, std::thread::hardware_concurrency()));
                        return std::max(1, std::min(t, 64));
                    }();
                    std::vector<std::thread> th(T);
                    for (int t = 0; t < T; t++) {
                        // words [lo, hi) of the chunk, column by column (the columns need not be adjacent in host memory)
                        size_t lo = words * t / T, hi = words * (t + 1) / T;
                        th[t] = std::thread([=] {
                            for (size_t x = lo; x < hi;) {
                                const size_t col = x / nr, r = x % nr, run = std::min(nr - r, hi - x);
                                memcpy(dst + x, col_ptr(a + (int)col) + r0 + r, run * 8);
                                x += run;
                            }
                        });
                    }
                    for (int t = 0; t < T; t++) th[t].join();
                    if (nr == n) CUDA_CHECK(cudaMemcpyAsync(d_dst, dst, words * 8, cudaMemcpyHostToDevice, up.copy));
                    else CUDA_CHECK(cudaMemcpy2DAsync(d_dst, n * 8, dst, nr * 8, nr * 8, e - a, cudaMemcpyHostToDevice, up.copy));
                    CUDA_CHECK(cudaEventRecord(up.stage_free[slot], up.copy));
                    slot = (slot + 1) % p2g_circuit::Upload::NSTAGE;
                } else if (!cols) {
                    if (nr == n) CUDA_CHECK(cudaMemcpyAsync(d_dst, col_ptr(a), words * 8, cudaMemcpyHostToDevice, up.copy));
                    else CUDA_CHECK(cudaMemcpy2DAsync(d_dst, n * 8, col_ptr(a) + r0, n * 8, nr * 8, e - a, cudaMemcpyHostToDevice, up.copy));
                } else {
                    for (int x = a; x < e; x++)
                        CUDA_CHECK(cudaMemcpyAsync(C->wires_values.p + (size_t)x * n + r0, col_ptr(x) + r0, nr * 8, cudaMemcpyHostToDevice, up.copy));
                }
                up.bytes += (double)words * 8;
            };
            auto copy_cols = [&](int a, int e) { copy_block(a, e, 0, n); };
            up.chunks.clear();
            up.canon = cols != nullptr;
            up.bytes = 0;
            up.routed = nullptr;
            // the copy stream must not overwrite the staging buffer while the previous proof's kernels still read it
            cudaEvent_t idle = next_event();
            CUDA_CHECK(cudaEventRecord(idle, c->stream));
            CUDA_CHECK(cudaStreamWaitEvent(up.copy, idle, 0));
            up.start = next_event();
            CUDA_CHECK(cudaEventRecord(up.start, up.copy));
            ev_start = up.start;
            int c0, c1;
            column_block(C, W, &c0, &c1);
            const int c_up = routed_only ? R : c1;   // routed_only (single GPU: c0 = 0, c1 = W): the upload stops at the routed columns
            step = std::max(step, (c_up - c0 + 7) / 8);                                       // ... and at most 8 chunks for small traces
            for (int a = c0; a < c_up; a += step) {
                int e = std::min(c_up, a + step);
                copy_cols(a, e);
                cudaEvent_t ev = next_event();
                CUDA_CHECK(cudaEventRecord(ev, up.copy));
                up.chunks.emplace_back(a, e, ev);
                up.last = ev;
            }
            up.fill_from = -1;
            if (routed_only) {
                up.fill_from = R;
                for (int a = R; a < W; a += step) up.chunks.emplace_back(a, std::min(W, a + step), (cudaEvent_t) nullptr);
            }
            if (c0 > 0 || c1 < R) {   // sharded: routed columns outside the block -- only the rows of this rank's Z / partial products
                if (c0 > 0) copy_block(0, std::min(c0, R), C->row0, C->row1);
                if (c1 < R) copy_block(c1, R, C->row0, C->row1);
                up.routed = next_event();
                CUDA_CHECK(cudaEventRecord(up.routed, up.copy));
                up.last = up.routed;
            }
            up.active = true;
            d_wires = C->wires_values.p;
        }
        prove_impl(C, d_wires, public_inputs, n_pi, forced_pow, out, out_len, tm, ev_start, compressed, cols != nullptr);
        up.active = false;
    });
}

extern "C" int p2g_prove(p2g_circuit* c, const uint64_t* wires, const uint64_t* public_inputs, size_t num_public_inputs,
                         const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len, p2g_timings* timings) {
    return prove_entry(c, wires, false, public_inputs, num_public_inputs, forced_pow_witness, out, out_len, timings);
}
// Same proof in plonky2's CompressedProofWithPublicInputs::to_bytes layout: exactly the bytes the reference CLI writes
// (prove_action.rs:75-78: proof.compress(..).to_bytes()); `wires_on_device` selects a host or device trace pointer.
extern "C" int p2g_prove_compressed(p2g_circuit* c, const uint64_t* wires, int wires_on_device, const uint64_t* public_inputs,
                                    size_t num_public_inputs, const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len,
                                    p2g_timings* timings) {
    return prove_entry(c, wires, wires_on_device != 0, public_inputs, num_public_inputs, forced_pow_witness, out, out_len, timings,
                       true);
}
// The witness as plonky2 holds it: MatrixWitness.wire_values is Vec<Vec<F>>, one heap allocation per wire column (SURVEY 8a row
// a3), so the shim passes the W column pointers and never flattens 1.96 GB on the host.  Columns in pageable memory are staged
// through the library's pinned ring by several host threads; page-locked columns (cudaHostRegister) are copied directly.
extern "C" int p2g_prove_columns(p2g_circuit* c, const uint64_t* const* wire_columns, const uint64_t* public_inputs,
                                 size_t num_public_inputs, const uint64_t* forced_pow_witness, int compressed, uint8_t* out,
                                 size_t* out_len, p2g_timings* timings) {
    return prove_entry(c, nullptr, false, public_inputs, num_public_inputs, forced_pow_witness, out, out_len, timings,
                       compressed != 0, wire_columns);
}
// p2g_prove_columns with two thirds of the trace left at home: only the num_routed_wires routed columns are passed and uploaded;
// the advice columns are computed on the device (k_fill_advice) as soon as the routed ones are in place, inside the same chunked
// upload / inverse NTT / LDE pipeline.
extern "C" int p2g_prove_routed_columns(p2g_circuit* c, const uint64_t* const* routed_columns, const uint64_t* public_inputs,
                                        size_t num_public_inputs, const uint64_t* forced_pow_witness, int compressed, uint8_t* out,
                                        size_t* out_len, p2g_timings* timings) {
    return prove_entry(c, nullptr, false, public_inputs, num_public_inputs, forced_pow_witness, out, out_len, timings,
                       compressed != 0, routed_columns, true);
}
extern "C" int p2g_prove_device(p2g_circuit* c, const uint64_t* d_wires, const uint64_t* public_inputs, size_t num_public_inputs,
                                const uint64_t* forced_pow_witness, uint8_t* out, size_t* out_len, p2g_timings* timings) {
    return prove_entry(c, d_wires, true, public_inputs, num_public_inputs, forced_pow_witness, out, out_len, timings);
}

// Completes the advice columns (>= num_routed_wires) of a device-resident trace in place from its routed columns (advice.cuh).
extern "C" int p2g_fill_advice_device(p2g_circuit* C, uint64_t* d_wires) {
    return guard([&] {
        if (!C || !d_wires) throw p2g_error(P2G_EBADARG, "p2g_fill_advice_device: null argument");
        std::lock_guard<std::mutex> lk(C->mu);
        DevCtx* c = C->ctx;
        CUDA_CHECK(cudaSetDevice(c->device));
        launch_fill_advice(C, d_wires);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

extern "C" int p2g_circuit_read(p2g_circuit* C, int what, void* out, size_t* len) {
    return guard([&] {
        if (!C || !len) throw p2g_error(P2G_EBADARG, "p2g_circuit_read: null argument");
        std::lock_guard<std::mutex> lk(C->mu);
        DevCtx* c = C->ctx;
        CUDA_CHECK(cudaSetDevice(c->device));
        const p2g_circuit_desc& d = C->d;
        const int NC = d.num_challenges, nzp = NC * (1 + d.num_partial_products), nq = NC * d.quotient_degree_factor;
        const void* dsrc = nullptr;   // device source
        const void* hsrc = nullptr;   // host source
        std::vector<uint8_t> packed;
        size_t sz = 0;
        auto cap_of = [&](const std::vector<digest_t>& cap) {
            packed.resize(cap.size() * C->hs);
            pack_digests(C->h, cap.data(), cap.size(), packed.data());
            hsrc = packed.data();
            sz = packed.size();
        };
        u64 info[6] = {(u64)C->rank, (u64)C->world, (u64)C->npeer, (u64)(C->npeer == C->world - 1 && C->world > 1),
                       (u64)(C->nccl != nullptr), (u64)C->collectives};
        if (what != P2G_BUF_CS_CAP && what != P2G_BUF_SHARD_INFO && !C->proved) throw p2g_error(P2G_EBADARG, "p2g_circuit_read: no proof has been produced yet");
        switch (what) {
        case P2G_BUF_WIRES_CAP: cap_of(C->last_caps[0]); break;
        case P2G_BUF_ZS_PP_CAP: cap_of(C->last_caps[1]); break;
        case P2G_BUF_QUOTIENT_CAP: cap_of(C->last_caps[2]); break;
        case P2G_BUF_CS_CAP: cap_of(C->cs_cap); break;
        case P2G_BUF_ZS_PP_VALUES: dsrc = C->zpp_values.p; sz = (size_t)nzp * C->n * 8; break;
        case P2G_BUF_QUOTIENT_CHUNKS: dsrc = C->quot.coeffs.p; sz = (size_t)nq * C->n * 8; break;
        case P2G_BUF_WIRES_COEFFS: dsrc = C->wires.coeffs.p; sz = (size_t)d.num_wires * C->n * 8; break;
        case P2G_BUF_CHALLENGES: hsrc = C->challenges.data(); sz = C->challenges.size() * 8; break;
        case P2G_BUF_FINAL_POLY: hsrc = C->final_poly.data(); sz = C->final_poly.size() * 8; break;
        case P2G_BUF_FRI_CAPS:
            packed.resize(C->fri_caps.size() * C->hs);
            pack_digests(C->h, C->fri_caps.data(), C->fri_caps.size(), packed.data());
            hsrc = packed.data();
            sz = packed.size();
            break;
        case P2G_BUF_SHARD_INFO: hsrc = info; sz = sizeof(info); break;
        case P2G_BUF_WIRES_LDE: dsrc = C->wires.lde.p; sz = (size_t)d.num_wires * C->lde_l * 8; break;   // this rank's leaves
        default: throw p2g_error(P2G_EBADARG, "p2g_circuit_read: unknown buffer");
        }
        if (!out || *len < sz) {
            *len = sz;
            throw p2g_error(P2G_ESMALLBUF, "p2g_circuit_read: buffer too small");
        }
        *len = sz;
        if (dsrc) {
            CUDA_CHECK(cudaMemcpyAsync(out, dsrc, sz, cudaMemcpyDeviceToHost, c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
        } else if (sz) {
            memcpy(out, hsrc, sz);
        }
    });
}

// filtered gate constraints at arbitrary points (evaluate_gate_constraints_base_batch), for gate-level parity tests
extern "C" int p2g_eval_gate_constraints(const p2g_circuit_desc* desc, const uint64_t* constants, const uint64_t* wires,
                                         const uint64_t* pi_hash, size_t npoints, uint64_t* out, int device) {
    return guard([&] {
        if (!desc || !constants || !wires || !out || !desc->gates) throw p2g_error(P2G_EBADARG, "p2g_eval_gate_constraints: null argument");
        if (desc->num_gates == 0 || desc->num_gates > P2G_MAX_GATES) throw p2g_error(P2G_EBADARG, "p2g_eval_gate_constraints: gate table");
        DevCtx* c = get_ctx(device);
        if (!npoints) return;
        static thread_local QuotientParams qp;
        memset(&qp, 0, sizeof(qp));
        qp.num_selectors = desc->num_selectors;
        qp.num_constants = desc->num_constants;
        qp.num_gates = desc->num_gates;
        for (u32 g = 0; g < desc->num_gates; g++) {
            const p2g_gate& s = desc->gates[g];
            if (s.kind >= P2G_GATE_KIND_COUNT || s.num_constraints > desc->num_gate_constraints)
                throw p2g_error(P2G_EBADARG, "p2g_eval_gate_constraints: bad gate");
            GateDev& o = qp.gates[g];
            o.kind = s.kind;
            for (int k = 0; k < 4; k++) o.params[k] = s.params[k];
            o.selector_index = s.selector_index;
            o.group_lo = s.group_lo;
            o.group_hi = s.group_hi;
            o.num_constraints = s.num_constraints;
        }
        if (pi_hash) memcpy(qp.pi_hash, pi_hash, 32);
        size_t nc = (size_t)desc->num_constants * npoints, nw = (size_t)desc->num_wires * npoints,
               no = (size_t)desc->num_gate_constraints * npoints;
        dbuf<u64> dc(nc), dw(nw), dout(std::max<size_t>(no, 1));
        CUDA_CHECK(cudaMemcpyAsync(dc.p, constants, nc * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(dw.p, wires, nw * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemsetAsync(dout.p, 0, no * 8, c->stream));
        gates_eval_standalone(c, qp, dc.p, dw.p, dout.p, npoints);
        CUDA_CHECK(cudaMemcpyAsync(out, dout.p, no * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
