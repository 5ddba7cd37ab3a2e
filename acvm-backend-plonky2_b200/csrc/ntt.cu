// Goldilocks NTT / LDE kernels for sm_100a.
//
// Replaces plonky2_field 0.2.2 fft.rs (fft_classic / ifft) and plonky2 fri/oracle.rs PolynomialBatch::{from_values,lde_values},
// reached from the reference at plonky2-backend/src/actions/prove_action.rs:96.
//
// Design (B200-first, not a translation of the radix-2 CPU loop):
//  * column-major batches, one grid covers every column (grid.y) and every LDE coset (grid.z) of a commitment;
//  * a transform of size 2^n is split into <= 3 passes; each pass moves a [A x Q] tile (A = sub-transform length, Q = 64-byte
//    runs of consecutive elements, so strided passes still read/write whole sectors) into shared memory, runs log2(A)
//    butterfly stages there, applies the inter-pass twist on the way out and writes the tile back -- HBM sees each element
//    once per pass;
//  * the per-stage twiddle tile (stage-major, conflict-free) is staged into shared memory by a TMA bulk copy
//    (cp.async.bulk + mbarrier) overlapped with the tile load;
//  * forward transforms are decimation-in-frequency (natural in -> bit-reversed out), which IS plonky2's leaf order, so the
//    reference's transpose + reverse_index_bits pass disappears; the rate-8 LDE is 8 independent size-N coset transforms
//    (leaves [r*N,(r+1)*N) = coset shift*omega_{8N}^{bitrev3(r)}), never a zero-padded size-8N transform;
//  * inverse transforms are decimation-in-time (bit-reversed in -> natural out); for natural-order input (the witness) the
//    first pass gathers bit-reversed 64-byte runs, so no separate permutation pass exists either.
#include <stdlib.h>

#include "internal.h"

namespace {

struct PassArgs {
    const u64* in;
    u64* out;
    size_t in_cs, out_cs;      // column strides (elements)
    size_t in_zs, out_zs;      // per-coset (grid.z) offsets (elements)
    int logn;                  // column transform length
    int logB;                  // strided pass: current block length (A * S)
    int loga;                  // sub-transform length A
    int logq;                  // strided: tile width Q; contiguous: log2(blocks per tile)
    const u64* tw;             // stage-major twiddles of size A: stage u at offset A - (A >> u), (A >> (u+1)) entries
    const u64* twist;          // full inter-pass twist table of the block: twist[(m << logS) + q] = omega_B^{+-q bitrev_a(m)}, or null
    int twist_split;           // (unused with full tables)
    int stab_full;             // stab is a full table of 2^logn entries per coset (forward LDE) instead of a [lo | hi] pair
    const u64* stab;           // [lo 2^split | hi] index-power table applied on load (forward) / store (inverse), or null
    int stab_split;
    size_t stab_zs;            // per-coset table stride
    u64 scale;                 // constant multiplier on store (1 = none)
    int gather;                // contiguous inverse pass: input is in natural order, gather bit-reversed runs
    u32 nz;                    // cosets per tile
    int col_fast;              // grid = (columns, tiles * nz): the columns of one (tile, coset) are adjacent in launch order
    // coset-sharded proofs: the final pass of the inverse transform also stores every result into the same place of each peer
    // GPU's coefficient buffer (P2P stores over NVLink), so the column blocks are exchanged by the kernel that computes them
    int npeer;
    u64* peer_out[P2G_MAX_PEERS];
};

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

// TMA bulk copy global -> shared of the twiddle tile, completion on an mbarrier (SASS: UBLKCP + SYNCS)
__device__ __forceinline__ void tma_load_tw(u64* dst, const u64* src, u32 bytes, u64* mbar) {
    u32 bar = smem_u32(mbar), d = smem_u32(dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_init(u64* mbar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(u64* mbar, u32 phase) {
    u32 bar = smem_u32(mbar), ok = 0;
    for (int spin = 0; spin < (1 << 26) && !ok; spin++) {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(bar), "r"(phase)
            : "memory");
    }
    if (!ok) __trap();
}

// stage the twiddle tile: TMA when it is at least 16 bytes, plain loads otherwise
__device__ __forceinline__ void stage_tw_begin(u64* s_tw, const u64* g_tw, int A, u64* mbar) {
    if (A >= 4) {
        if (threadIdx.x == 0) mbar_init(mbar);
        __syncthreads();
        if (threadIdx.x == 0) tma_load_tw(s_tw, g_tw, (u32)(A * 8), mbar);
    } else {
        if (threadIdx.x < A) s_tw[threadIdx.x] = g_tw[threadIdx.x];
    }
}
__device__ __forceinline__ void stage_tw_end(int A, u64* mbar) {
    if (A >= 4) mbar_wait(mbar, 0);
    __syncthreads();
}

__device__ __forceinline__ u64 tab_pow(const u64* tab, int split, u32 i) {
    u64 lo = __ldg(tab + (i & ((1u << split) - 1)));
    u64 hi = __ldg(tab + (1u << split) + (i >> split));
    return gl_mul(lo, hi);
}

// Shared-memory tile: logical index e = (((bb << a) + m) << logq) | qq  (bb: block of a contiguous pass, m: position along
// the sub-transform, qq: position inside a 2^logq-element run of a strided pass), stored at e + (e >> 4): one pad word per 16
// keeps every access pattern of the rounds below free of bank conflicts.
__device__ __forceinline__ int phys(int e) { return e + (e >> 4); }
__host__ __device__ inline size_t tile_words(size_t elems) { return (elems + (elems >> 4) + 3) & ~(size_t)1; }   // even: TMA dst 16 B aligned

// v * omega_{2h}^{+-kk} for the compile-time-unrolled (h, kk) of a LAST round: omega_{2h} = 2^(96 / h), so the factor is 2^S with
// S = 96 kk / h forward and -2^(96 - S) inverse (the caller folds the sign into its add / sub).  Returns v 2^S resp. v 2^(96-S).
template <int R>
__device__ __forceinline__ u64 shl_by_index(u64 v, int, int h, int kk, bool inv) {
    const int S = inv ? 96 - 96 * kk / h : 96 * kk / h;   // h in {1,2,4,8}, 0 < kk < h: multiples of 12 in (0, 96)
    switch (S) {
    case 12: return gl_shl<12>(v);
    case 24: return gl_shl<24>(v);
    case 36: return gl_shl<36>(v);
    case 48: return gl_shl<48>(v);
    case 60: return gl_shl<60>(v);
    case 72: return gl_shl<72>(v);
    case 84: return gl_shl<84>(v);
    default: return v;   // unreachable
    }
}

// One register round: R butterfly stages [s0, s0 + R) of the 2^a-point sub-transforms of the tile.  A thread owns the 2^R
// elements that differ in bits [a-s0-R, a-s0) of m, runs the R stages on them in registers (radix-2^R, 2^(R-1) R butterflies,
// 2^R - 1 twiddle loads) and writes them back: the tile makes one shared-memory round trip per R stages instead of one per
// stage.  Forward = decimation in frequency (stage order s0 .. s0+R-1), inverse = decimation in time (reverse order).
// LAST: the round that ends the sub-transform (s0 + R == a).  Its twiddles are omega_{2h}^kk with 2h <= 2^R <= 16, i.e. the
// compile-time powers of two 2^(96 kk / h): multiplication-free butterflies, no twiddle loads.
// The R butterfly stages [s0, s0 + R) on the 2^R registers of one task (`lo` = the task's position inside the half-blocks of the
// round's last stage, lo_bits its width): forward = decimation in frequency, inverse = decimation in time.
template <int R, bool INV, bool LAST>
__device__ __forceinline__ void reg_butterflies(u64 (&x)[1 << R], const u64* __restrict__ s_tw, int A, int s0, int lo_bits, int lo) {
#pragma unroll
    for (int ii = 0; ii < R; ii++) {
        const int i = INV ? R - 1 - ii : ii;
        const int h = 1 << (R - 1 - i);
        const u64* tw = s_tw + (A - (A >> (s0 + i)));
#pragma unroll
        for (int kk = 0; kk < h; kk++) {
            u64 w = 0;
            if (!LAST) w = tw[(kk << lo_bits) | lo];
#pragma unroll
            for (int g = 0; g < (1 << R); g += 2 * h) {
                u64 u = x[g + kk], v = x[g + kk + h];
                if (LAST) {
                    // forward: (u + v, (u - v) 2^S), S = 96 kk / h;  inverse: v 2^-S = -v 2^(96 - S)
                    if (kk == 0) {
                        x[g] = glf_add(u, v);
                        x[g + h] = gl_sub(u, v);
                    } else if (INV) {
                        v = shl_by_index<R>(v, 0, h, kk, true);
                        x[g + kk] = gl_sub(u, v);
                        x[g + kk + h] = glf_add(u, v);
                    } else {
                        x[g + kk] = glf_add(u, v);
                        x[g + kk + h] = shl_by_index<R>(gl_sub(u, v), 0, h, kk, false);
                    }
                } else if (INV) {
                    v = glf_mul(v, w);
                    x[g + kk] = glf_add(u, v);
                    x[g + kk + h] = gl_sub(u, v);
                } else {
                    x[g + kk] = glf_add(u, v);
                    x[g + kk + h] = glf_mul(gl_sub(u, v), w);
                }
            }
        }
    }
}

// CA / CS0 / CQ / CLE / CTH: the sub-transform length, first stage, run width, tile size and block size as compile-time constants (the hot
// plans of the 2^20-row LDE are instantiated that way: every shift, mask and padded address below folds; ncu showed a third of a
// round's instructions going into them).  A thread's 2^R elements sit `stride` apart in the tile; when stride is a multiple of 16
// their padded addresses are phys(first) + k * (stride + stride / 16).
template <int R, bool INV, bool LAST, int CA = 0, int CS0 = -1, int CQ = -1, int CLE = 0, int CTH = 0>
__device__ __forceinline__ void reg_round(u64* __restrict__ sm, const u64* __restrict__ s_tw, int a_, int s0_, int logq_, int log_tasks_) {
    const int a = CA > 0 ? CA : a_, s0 = CS0 >= 0 ? CS0 : s0_, logq = CQ >= 0 ? CQ : logq_;
    const int log_tasks = CLE > 0 ? CLE - R : log_tasks_;   // CLE: log2 of the tile's element count
    const int nthreads = CTH > 0 ? CTH : (int)blockDim.x;
    const int lo_bits = LAST ? 0 : a - s0 - R;
    const int A = 1 << a;
    const int lo_mask = (1 << lo_bits) - 1, q_mask = (1 << logq) - 1;
    const int stride = 1 << (lo_bits + logq);
    const bool linear = (stride & 15) == 0;
    const int pstride = stride + (stride >> 4);
    for (int task = threadIdx.x; task < (1 << log_tasks); task += nthreads) {
        const int qq = task & q_mask, t = task >> logq;
        const int lo = t & lo_mask;
        const int base_m = ((t >> lo_bits) << (lo_bits + R)) | lo;
        const int e0 = (base_m << logq) | qq, p0 = phys(e0);
        u64 x[1 << R];
#pragma unroll
        for (int k = 0; k < (1 << R); k++) x[k] = sm[linear ? p0 + k * pstride : phys(e0 + k * stride)];
        reg_butterflies<R, INV, LAST>(x, s_tw, A, s0, lo_bits, lo);
#pragma unroll
        for (int k = 0; k < (1 << R); k++) sm[linear ? p0 + k * pstride : phys(e0 + k * stride)] = x[k];
    }
    __syncthreads();
}

// all `a` stages of the tile: one round with the remainder a % RMAX, then rounds of RMAX stages -- the last round (the one with
// the multiplication-free twiddles) is a full one
template <bool INV, int RMAX>
__device__ __forceinline__ void tile_butterflies(u64* sm, const u64* s_tw, int a, int logq, int log_elems) {
    const int rem = a % RMAX, full = a / RMAX;
    if (!INV) {
        int s0 = 0;
        if (full == 0) {
            if (rem == 3) reg_round<3, false, true>(sm, s_tw, a, 0, logq, log_elems - 3);
            else if (rem == 2) reg_round<2, false, true>(sm, s_tw, a, 0, logq, log_elems - 2);
            else if (rem == 1) reg_round<1, false, true>(sm, s_tw, a, 0, logq, log_elems - 1);
            return;
        }
        if (rem == 3) reg_round<3, false, false>(sm, s_tw, a, 0, logq, log_elems - 3);
        else if (rem == 2) reg_round<2, false, false>(sm, s_tw, a, 0, logq, log_elems - 2);
        else if (rem == 1) reg_round<1, false, false>(sm, s_tw, a, 0, logq, log_elems - 1);
        s0 = rem;
        for (int r = 0; r + 1 < full; r++, s0 += RMAX) reg_round<RMAX, false, false>(sm, s_tw, a, s0, logq, log_elems - RMAX);
        reg_round<RMAX, false, true>(sm, s_tw, a, s0, logq, log_elems - RMAX);
    } else {
        if (full == 0) {
            if (rem == 3) reg_round<3, true, true>(sm, s_tw, a, 0, logq, log_elems - 3);
            else if (rem == 2) reg_round<2, true, true>(sm, s_tw, a, 0, logq, log_elems - 2);
            else if (rem == 1) reg_round<1, true, true>(sm, s_tw, a, 0, logq, log_elems - 1);
            return;
        }
        int s0 = a - RMAX;
        reg_round<RMAX, true, true>(sm, s_tw, a, s0, logq, log_elems - RMAX);
        for (int r = 0; r + 1 < full; r++) {
            s0 -= RMAX;
            reg_round<RMAX, true, false>(sm, s_tw, a, s0, logq, log_elems - RMAX);
        }
        if (rem == 3) reg_round<3, true, false>(sm, s_tw, a, 0, logq, log_elems - 3);
        else if (rem == 2) reg_round<2, true, false>(sm, s_tw, a, 0, logq, log_elems - 2);
        else if (rem == 1) reg_round<1, true, false>(sm, s_tw, a, 0, logq, log_elems - 1);
    }
}

// inverse order: the multiplication-free round first, then the generic rounds with descending stages, then the remainder round
template <int CA, int CQ, int CLE, int CTH>
__device__ __forceinline__ void tile_butterflies_fixed_inv(u64* sm, const u64* s_tw) {
    constexpr int rem = CA % 3, full = CA / 3;
    static_assert(full >= 1 && full <= 3, "shape");
    reg_round<3, true, true, CA, CA - 3, CQ, CLE, CTH>(sm, s_tw, CA, CA - 3, CQ, CLE - 3);
    if (full >= 2) reg_round<3, true, false, CA, CA - 6, CQ, CLE, CTH>(sm, s_tw, CA, CA - 6, CQ, CLE - 3);
    if (full >= 3) reg_round<3, true, false, CA, CA - 9, CQ, CLE, CTH>(sm, s_tw, CA, CA - 9, CQ, CLE - 3);
    if (rem == 2) reg_round<2, true, false, CA, 0, CQ, CLE, CTH>(sm, s_tw, CA, 0, CQ, CLE - 2);
    else if (rem == 1) reg_round<1, true, false, CA, 0, CQ, CLE, CTH>(sm, s_tw, CA, 0, CQ, CLE - 1);
}

// Strided pass: tile [A][Q], element (m, qq) at column index blk*B + m*S + q0 + qq.  grid.x = tile * nz + coset: the cosets
// of one coefficient tile are adjacent in launch order, so the tile is read from HBM once and from L2 seven times.
// CA / CQ / CTH > 0 (inverse passes of the hot plans): the butterflies run with compile-time shapes; fill and drain stay generic
template <bool INV, int RMAX, int CA = 0, int CQ = 0, int CTH = 0>
__global__ void __launch_bounds__(512, RMAX == 4 ? 1 : 2) k_pass_strided(PassArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const int A = 1 << a.loga, Q = 1 << a.logq;
    const int logS = a.logB - a.loga;
    const int total = A << a.logq;
    u64* s_tw = sm + tile_words(total);
    u64* mbar = s_tw + A;
    const u32 col = a.col_fast ? blockIdx.x : blockIdx.y, tz = a.col_fast ? blockIdx.y : blockIdx.x;
    const u32 z = tz % a.nz, tile = tz / a.nz;
    const u32 tiles_per_blk = 1u << (logS - a.logq);
    const u32 blk = tile / tiles_per_blk;
    const u32 q0 = (tile % tiles_per_blk) << a.logq;
    const u64* in = a.in + (size_t)col * a.in_cs + (size_t)z * a.in_zs;
    u64* out = a.out + (size_t)col * a.out_cs + (size_t)z * a.out_zs;
    const u64* stab = a.stab ? a.stab + (size_t)z * a.stab_zs : nullptr;
    const size_t base = ((size_t)blk << a.logB) + q0;

    stage_tw_begin(s_tw, a.tw, A, mbar);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int qq = e & (Q - 1), m = e >> a.logq;
        size_t idx = base + ((size_t)m << logS) + qq;
        u64 v = in[idx];
        if (!INV && stab) v = glf_mul(v, a.stab_full ? __ldg(stab + idx) : tab_pow(stab, a.stab_split, (u32)idx));
        if (INV && a.twist) v = glf_mul(v, __ldg(a.twist + (((size_t)m << logS) + q0 + qq)));
        sm[phys(e)] = v;
    }
    stage_tw_end(A, mbar);
    if (CA > 0 && INV) tile_butterflies_fixed_inv<CA ? CA : 3, CQ, CA + CQ, CTH>(sm, s_tw);
    else tile_butterflies<INV, RMAX>(sm, s_tw, a.loga, a.logq, a.loga + a.logq);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int qq = e & (Q - 1), m = e >> a.logq;
        size_t idx = base + ((size_t)m << logS) + qq;
        u64 v = sm[phys(e)];
        if (!INV && a.twist) v = glf_mul(v, __ldg(a.twist + (((size_t)m << logS) + q0 + qq)));
        if (INV && stab) v = glf_mul(v, tab_pow(stab, a.stab_split, (u32)idx));
        if (a.scale != 1) v = gl_mul(v, a.scale);
        out[idx] = v;
        if (INV)
            for (int p = 0; p < a.npeer; p++) a.peer_out[p][(size_t)col * a.out_cs + idx] = v;
    }
}

// Contiguous pass: tile = 2^logq consecutive blocks of A elements.
template <bool INV, int RMAX, int CA = 0, int CNB = 0, int CTH = 0>
__global__ void __launch_bounds__(512, RMAX == 4 ? 1 : 2) k_pass_contig(PassArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    const int A = 1 << a.loga, NB = 1 << a.logq;
    const int total = A << a.logq;
    u64* s_tw = sm + tile_words(total);
    u64* mbar = s_tw + A;
    const u32 col = a.col_fast ? blockIdx.x : blockIdx.y, tz = a.col_fast ? blockIdx.y : blockIdx.x;
    const u32 z = tz % a.nz, tile = tz / a.nz;
    const u64* in = a.in + (size_t)col * a.in_cs + (size_t)z * a.in_zs;
    u64* out = a.out + (size_t)col * a.out_cs + (size_t)z * a.out_zs;
    const u64* stab = a.stab ? a.stab + (size_t)z * a.stab_zs : nullptr;
    const int lognb = a.logn - a.loga;  // log2(#blocks in the column)
    const u32 c0 = tile << a.logq;

    stage_tw_begin(s_tw, a.tw, A, mbar);
    if (INV && a.gather) {
        // virtual bit-reversed array: block b = bitrev(c), element m' <- natural index bitrev_a(m') * (N/A) + c
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            int c = e & (NB - 1), m = e >> a.logq;
            size_t idx = ((size_t)bitrev32(m, a.loga) << lognb) + c0 + c;
            sm[phys((c << a.loga) + m)] = in[idx];
        }
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            size_t idx = ((size_t)c0 << a.loga) + e;
            u64 v = in[idx];
            if (!INV && stab) v = glf_mul(v, a.stab_full ? __ldg(stab + idx) : tab_pow(stab, a.stab_split, (u32)idx));
            sm[phys(e)] = v;
        }
    }
    stage_tw_end(A, mbar);
    if (CA > 0 && INV) tile_butterflies_fixed_inv<CA ? CA : 3, 0, CA + CNB, CTH>(sm, s_tw);
    else tile_butterflies<INV, RMAX>(sm, s_tw, a.loga, 0, a.loga + a.logq);
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        int bb = e >> a.loga, m = e & (A - 1);
        u32 blk = (INV && a.gather) ? bitrev32(c0 + bb, lognb) : (c0 + bb);
        size_t idx = ((size_t)blk << a.loga) + m;
        u64 v = sm[phys(e)];
        if (INV && stab) v = gl_mul(v, tab_pow(stab, a.stab_split, (u32)idx));
        if (a.scale != 1) v = gl_mul(v, a.scale);
        out[idx] = v;
        if (INV)
            for (int p = 0; p < a.npeer; p++) a.peer_out[p][(size_t)col * a.out_cs + idx] = v;
    }
}

// ---- forward passes with compile-time shapes (the plans of the hot LDE sizes) ---------------------------------------------------
// Same data movement and arithmetic as k_pass_strided<false> / k_pass_contig<false>; tile size, sub-transform length, run width and
// block size are template constants, so the fill / drain loops are fully unrolled with constant strides and the register rounds'
// index arithmetic folds (ncu source page of the generic kernels: ~35 % of a round's instructions were address arithmetic).
// The first register round runs on the elements a thread has just loaded (its EPT fill elements ARE whole tasks of that round) and
// the last one on the elements it is about to store: two of the tile's shared-memory round trips and their barriers disappear.
template <int CA, int CQ, int CTH>
__global__ void __launch_bounds__(CTH, 1024 / CTH) k_pass_strided_fwd(PassArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int A = 1 << CA, Q = 1 << CQ, total = A << CQ, EPT = total / CTH, MSTEP = CTH >> CQ, PSTEP = CTH + (CTH >> 4);
    constexpr int R0 = (CA % 3 == 0) ? 3 : CA % 3, NT = EPT >> R0, MID = (CA - R0 - 3) / 3;   // first round, its tasks per thread, middle rounds
    static_assert(CTH % Q == 0 && CTH % 16 == 0 && total % CTH == 0 && EPT >= 8 && CA >= R0 + 3 && (CA - R0) % 3 == 0 && MID <= 2, "shape");
    const int logS = a.logB - CA;
    u64* s_tw = sm + tile_words(total);
    u64* mbar = s_tw + A;
    const u32 col = a.col_fast ? blockIdx.x : blockIdx.y, tz = a.col_fast ? blockIdx.y : blockIdx.x;
    const u32 z = tz % a.nz, tile = tz / a.nz;
    const u32 tiles_per_blk = 1u << (logS - CQ);
    const u32 blk = tile / tiles_per_blk;
    const u32 q0 = (tile % tiles_per_blk) << CQ;
    const u64* in = a.in + (size_t)col * a.in_cs + (size_t)z * a.in_zs;
    u64* out = a.out + (size_t)col * a.out_cs + (size_t)z * a.out_zs;
    const u64* stab = a.stab ? a.stab + (size_t)z * a.stab_zs : nullptr;
    const int tid = threadIdx.x, qq = tid & (Q - 1), m0 = tid >> CQ, p0 = phys(tid);
    // fill element i of this thread: tile index e = tid + i CTH -> (m0 + i MSTEP, qq); column index idx0 + i * istep
    const size_t idx0 = ((size_t)blk << a.logB) + q0 + ((size_t)m0 << logS) + qq, istep = (size_t)MSTEP << logS;

    stage_tw_begin(s_tw, a.tw, A, mbar);
    // fill + first round: task t of this thread = elements i = t + k NT (m = m0 + t MSTEP + k A / 2^R0)
#pragma unroll
    for (int t = 0; t < NT; t++) {
        u64 x[1 << R0], sc[1 << R0];
#pragma unroll
        for (int k = 0; k < (1 << R0); k++) {
            x[k] = in[idx0 + (t + k * NT) * istep];
            sc[k] = stab ? __ldg(stab + idx0 + (t + k * NT) * istep) : 1;
        }
        if (stab) {
#pragma unroll
            for (int k = 0; k < (1 << R0); k++) x[k] = glf_mul(x[k], sc[k]);
        }
        if (t == 0) stage_tw_end(A, mbar);   // the twiddle tile (TMA) has landed; every thread passes here exactly once
        reg_butterflies<R0, false, false>(x, s_tw, A, 0, CA - R0, m0 + t * MSTEP);
#pragma unroll
        for (int k = 0; k < (1 << R0); k++) sm[p0 + (t + k * NT) * PSTEP] = x[k];
    }
    __syncthreads();
    if (MID >= 1) reg_round<3, false, false, CA, R0, CQ, CA + CQ, CTH>(sm, s_tw, CA, R0, CQ, CA + CQ - 3);
    if (MID >= 2) reg_round<3, false, false, CA, R0 + 3, CQ, CA + CQ, CTH>(sm, s_tw, CA, R0 + 3, CQ, CA + CQ - 3);
    // last round (multiplication-free twiddles) + twist + drain: task = 8 consecutive m at one qq
    const u64* twist = a.twist ? a.twist + q0 + qq : nullptr;
    u64* o = out + ((size_t)blk << a.logB) + q0 + qq;
#pragma unroll
    for (int j = 0; j < EPT / 8; j++) {
        const int base_m = (m0 + j * MSTEP) << 3;
        u64 x[8], tw[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x[k] = sm[phys(((base_m + k) << CQ) | qq)];
            tw[k] = twist ? __ldg(twist + ((size_t)(base_m + k) << logS)) : 1;
        }
        reg_butterflies<3, false, true>(x, s_tw, A, CA - 3, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; k++) o[(size_t)(base_m + k) << logS] = twist ? glf_mul(x[k], tw[k]) : x[k];
    }
}

template <int CA, int CNB, int CTH>
__global__ void __launch_bounds__(CTH, 1024 / CTH) k_pass_contig_fwd(PassArgs a) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int A = 1 << CA, total = A << CNB, EPT = total / CTH, EPB = EPT >> CNB, PSTEP = CTH + (CTH >> 4);
    constexpr int R0 = (CA % 3 == 0) ? 3 : CA % 3, NTB = EPB >> R0, MID = (CA - R0 - 3) / 3;
    static_assert(CTH % 16 == 0 && total % CTH == 0 && A % CTH == 0 && EPB >= (1 << R0) && (CA - R0) % 3 == 0 && MID >= 0 && MID <= 2, "shape");
    u64* s_tw = sm + tile_words(total);
    u64* mbar = s_tw + A;
    const u32 col = a.col_fast ? blockIdx.x : blockIdx.y, tz = a.col_fast ? blockIdx.y : blockIdx.x;
    const u32 z = tz % a.nz, tile = tz / a.nz;
    const u64* in = a.in + (size_t)col * a.in_cs + (size_t)z * a.in_zs;
    u64* out = a.out + (size_t)col * a.out_cs + (size_t)z * a.out_zs;
    const u64* stab = a.stab ? a.stab + (size_t)z * a.stab_zs : nullptr;
    const int tid = threadIdx.x, p0 = phys(tid);
    const size_t idx0 = ((size_t)tile << (CA + CNB)) + tid;

    stage_tw_begin(s_tw, a.tw, A, mbar);
    // fill + first round: fill element i = tid + i CTH lies in block bb = i / EPB at m = tid + (i % EPB) CTH; task (bb, t) = the
    // elements i = bb EPB + t + k NTB, i.e. m = (tid + t CTH) + k A / 2^R0
#pragma unroll
    for (int bt = 0; bt < (NTB << CNB); bt++) {
        const int bb = bt / NTB, t = bt % NTB;
        u64 x[1 << R0];
#pragma unroll
        for (int k = 0; k < (1 << R0); k++) {
            const int i = bb * EPB + t + k * NTB;
            x[k] = in[idx0 + (size_t)i * CTH];
            if (stab) x[k] = glf_mul(x[k], __ldg(stab + idx0 + (size_t)i * CTH));   // single-pass plans only
        }
        if (bt == 0) stage_tw_end(A, mbar);
        reg_butterflies<R0, false, false>(x, s_tw, A, 0, CA - R0, tid + t * CTH);
#pragma unroll
        for (int k = 0; k < (1 << R0); k++) sm[p0 + (bb * EPB + t + k * NTB) * PSTEP] = x[k];
    }
    __syncthreads();
    if (MID >= 1) reg_round<3, false, false, CA, R0, 0, CA + CNB, CTH>(sm, s_tw, CA, R0, 0, CA + CNB - 3);
    if (MID >= 2) reg_round<3, false, false, CA, R0 + 3, 0, CA + CNB, CTH>(sm, s_tw, CA, R0 + 3, 0, CA + CNB - 3);
    // last round + drain: task = 8 consecutive elements, written as one 64-byte run
#pragma unroll
    for (int j = 0; j < EPT / 8; j++) {
        const int e0 = (tid + j * CTH) << 3, pe = e0 + (e0 >> 4);
        u64 x[8];
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = sm[pe + k];
        reg_butterflies<3, false, true>(x, s_tw, A, CA - 3, 0, 0);
        ulonglong2* o = reinterpret_cast<ulonglong2*>(out + ((size_t)tile << (CA + CNB)) + e0);
#pragma unroll
        for (int k = 0; k < 8; k += 2) o[k >> 1] = make_ulonglong2(x[k], x[k + 1]);
    }
}

// out[i] = premul * base^(i * step) for i < n
__global__ void k_fill_pow(u64* out, u32 n, u64 base, u64 step, u64 premul) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = gl_mul(premul, gl_pow(base, (u64)i * step));
}
// full twist table of a strided pass: out[(m << logS) + q] = root^(q * bitrev_a(m)), root = omega_B^{+-1}
__global__ void k_fill_twist(u64* out, int logB, int loga, u64 root) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ((size_t)1 << logB)) return;
    const int logS = logB - loga;
    u32 m = (u32)(i >> logS), q = (u32)(i & (((size_t)1 << logS) - 1));
    out[i] = gl_pow(root, (u64)q * bitrev32(m, loga));
}
// stage-major twiddle tile of size A = 2^loga: stage u holds root^(j << u), j < A >> (u+1)
__global__ void k_fill_tw(u64* out, int loga, u64 root) {
    int A = 1 << loga;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    if (i == A - 1) {
        out[i] = 0;
        return;
    }
    // find stage u with A - (A >> u) <= i < A - (A >> (u+1))
    int u = 0;
    while (i >= A - (A >> (u + 1))) u++;
    int j = i - (A - (A >> u));
    out[i] = gl_pow(root, (u64)j << u);
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
// tuning knobs (defaults chosen on B200, see profiles/): log2 of the tile size in elements, log2 of the last (contiguous)
// pass, threads per block (0 = by tile size)
int ntt_tile_log() { static int v = env_int("P2G_NTT_TILE", 12); return v; }
int ntt_last_log() { static int v = env_int("P2G_NTT_LAST", 11); return v; }
int ntt_threads() { static int v = env_int("P2G_NTT_TH", 0); return v; }
int ntt_fixed() { static int v = env_int("P2G_NTT_FIXED", 1); return v; }   // 0: generic kernels only (A/B)
int ntt_rmax() {
    static int r = 0;
    if (!r) {
        const char* e = getenv("P2G_NTT_R");
        r = (e && atoi(e) == 4) ? 4 : 3;   // 8 elements per thread: 64 registers, two 512-thread blocks per SM
    }
    return r;
}
const int MAX_CONTIG_LOG = 11;
const int MAX_STRIDED_LOG = 10;

// forward-order pass plan: strided passes (largest block first) then one contiguous pass
std::vector<int> make_plan(int logn, bool inverse = false) {
    std::vector<int> plan;
    if (logn <= MAX_CONTIG_LOG) {
        plan.push_back(logn);
        return plan;
    }
    // 2^22 forward (configs[4]): two passes 2^11 x 2^11 -- a 16 K-element strided tile (139 KB of shared memory, one 1024-thread block
    // per SM) -- instead of three passes over HBM
    if (!inverse && logn == 22 && ntt_fixed() && ntt_rmax() == 3 && !getenv("P2G_NTT_NO_BIG_TILE")) {
        plan.push_back(11);
        plan.push_back(11);
        return plan;
    }
    // the inverse transform's contiguous pass gathers 8 blocks per tile (64-byte runs of the natural-order input), so its
    // sub-transform is kept at 2^10 to stay at two resident tiles per SM
    int last = inverse ? std::min(10, ntt_last_log()) : std::min(ntt_last_log(), MAX_CONTIG_LOG);
    int rem = logn - last;
    int cnt = (rem + MAX_STRIDED_LOG - 1) / MAX_STRIDED_LOG;
    for (int i = 0; i < cnt; i++) {
        int a = (rem + (cnt - i) - 1) / (cnt - i);
        plan.push_back(a);
        rem -= a;
    }
    plan.push_back(last);
    return plan;
}

size_t strided_smem(int loga, int logq) { return tile_words((size_t)1 << (loga + logq)) * 8 + (size_t)(1 << loga) * 8 + 16; }
size_t contig_smem(int loga, int logq) { return strided_smem(loga, logq); }

void set_smem_attrs() {
    static std::mutex m;
    static std::vector<int> devs;   // the attribute is per device
    std::lock_guard<std::mutex> lk(m);
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    if (std::find(devs.begin(), devs.end(), dev) != devs.end()) return;
    devs.push_back(dev);
    const int lim = 160 * 1024;
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<9, 3, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<11, 3, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<7, 5, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<5, 7, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<6, 6, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided_fwd<8, 4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig_fwd<11, 1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided<true, 3, 10, 3, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig<true, 3, 10, 3, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_strided<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    CUDA_CHECK(cudaFuncSetAttribute(k_pass_contig<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
}

// Launch order (x fastest): all columns of one (tile, coset) first, then the next coset of the same tile.  The per-coset scale
// slice and the twist slice of the tile stay in L2 across the columns, and a column's coefficient tile is re-read by its 8
// cosets only `ncols` blocks later (also from L2).  Falls back to (tiles * nz, columns) when grid.y would overflow.
dim3 make_grid(PassArgs& a, size_t tz, int ncols) {
    a.col_fast = tz <= 65535;
    return a.col_fast ? dim3((unsigned)ncols, (unsigned)tz, 1) : dim3((unsigned)tz, (unsigned)ncols, 1);
}
int pick_threads(size_t tile_elems) {
    if (ntt_threads() && tile_elems >= 4096) return ntt_threads();
    return tile_elems >= 8192 ? 512 : (tile_elems >= 4096 ? 256 : (tile_elems >= 1024 ? 64 : 32));
}

// run the passes of one direction.  fwd: in -> out (first pass), then in place on out.
struct XformDesc {
    const u64* in;
    size_t in_cs, in_zs;
    u64* out;
    size_t out_cs, out_zs;
    int logn, ncols, nz;
    const u64* stab;
    int stab_split;
    size_t stab_zs;
    int stab_full;
    u64 scale;
    bool natural_input;  // inverse only
    int npeer;           // inverse only: peer copies of the output written by the last pass
    u64* const* peer_out;
};

void run_forward(DevCtx* c, const XformDesc& d) {
    set_smem_attrs();
    std::vector<int> plan = make_plan(d.logn);
    int done = 0;
    for (size_t pi = 0; pi < plan.size(); pi++) {
        bool first = pi == 0, last = pi + 1 == plan.size();
        PassArgs a = {};
        a.in = first ? d.in : d.out;
        a.in_cs = first ? d.in_cs : d.out_cs;
        a.in_zs = first ? d.in_zs : d.out_zs;
        a.out = d.out;
        a.out_cs = d.out_cs;
        a.out_zs = d.out_zs;
        a.logn = d.logn;
        a.loga = plan[pi];
        a.tw = c->get_tw(a.loga, false);
        a.scale = 1;
        if (first) {
            a.stab = d.stab;
            a.stab_split = d.stab_split;
            a.stab_zs = d.stab_zs;
            a.stab_full = d.stab_full;
        }
        if (!last) {
            a.logB = d.logn - done;
            int logS = a.logB - a.loga;
            a.logq = std::min(logS, std::max(3, ntt_tile_log() - a.loga));
            const bool big_tile = a.loga == 11 && a.logq == 3 && ntt_fixed() && ntt_rmax() == 3;
            a.twist = c->get_twist_full(a.logB, a.loga, false);
            size_t tiles = (size_t)1 << (d.logn - a.loga - a.logq);
            a.nz = d.nz;
            dim3 grid = make_grid(a, tiles * d.nz, d.ncols);
            int th = pick_threads((size_t)1 << (a.loga + a.logq));
            const bool fixed = ntt_fixed() && ntt_rmax() == 3 && th == 256 && a.scale == 1;
            if (big_tile) k_pass_strided_fwd<11, 3, 1024><<<grid, 1024, strided_smem(11, 3), c->stream>>>(a);
            else if (fixed && a.loga == 9 && a.logq == 3) k_pass_strided_fwd<9, 3, 256><<<grid, 256, strided_smem(9, 3), c->stream>>>(a);
            else if (fixed && a.loga == 7 && a.logq == 5) k_pass_strided_fwd<7, 5, 256><<<grid, 256, strided_smem(7, 5), c->stream>>>(a);
            else if (fixed && a.loga == 5 && a.logq == 7) k_pass_strided_fwd<5, 7, 256><<<grid, 256, strided_smem(5, 7), c->stream>>>(a);
            else if (fixed && a.loga == 6 && a.logq == 6) k_pass_strided_fwd<6, 6, 256><<<grid, 256, strided_smem(6, 6), c->stream>>>(a);
            else if (fixed && a.loga == 8 && a.logq == 4) k_pass_strided_fwd<8, 4, 256><<<grid, 256, strided_smem(8, 4), c->stream>>>(a);
            else if (ntt_rmax() == 4) k_pass_strided<false, 4><<<grid, th, strided_smem(a.loga, a.logq), c->stream>>>(a);
            else k_pass_strided<false, 3><<<grid, th, strided_smem(a.loga, a.logq), c->stream>>>(a);
        } else {
            int lognb = d.logn - a.loga;
            a.logq = std::min(lognb, std::max(0, ntt_tile_log() - a.loga));
            size_t tiles = (size_t)1 << (lognb - a.logq);
            a.nz = d.nz;
            dim3 grid = make_grid(a, tiles * d.nz, d.ncols);
            int th = pick_threads((size_t)1 << (a.loga + a.logq));
            const bool fixed = ntt_fixed() && ntt_rmax() == 3 && th == 256 && a.scale == 1;
            if (fixed && a.loga == 11 && a.logq == 1) k_pass_contig_fwd<11, 1, 256><<<grid, 256, contig_smem(11, 1), c->stream>>>(a);
            else if (ntt_rmax() == 4) k_pass_contig<false, 4><<<grid, th, contig_smem(a.loga, a.logq), c->stream>>>(a);
            else k_pass_contig<false, 3><<<grid, th, contig_smem(a.loga, a.logq), c->stream>>>(a);
        }
        count_launch(c);
        done += plan[pi];
    }
    CUDA_CHECK(cudaGetLastError());
}

void run_inverse(DevCtx* c, const XformDesc& d) {
    set_smem_attrs();
    std::vector<int> plan = make_plan(d.logn, true);
    int np = (int)plan.size();
    // inverse order: contiguous pass first, then strided passes with growing blocks
    int done = 0;
    for (int pi = np - 1; pi >= 0; pi--) {
        bool first = pi == np - 1, last = pi == 0;
        PassArgs a = {};
        a.in = first ? d.in : d.out;
        a.in_cs = first ? d.in_cs : d.out_cs;
        a.in_zs = first ? d.in_zs : d.out_zs;
        a.out = d.out;
        a.out_cs = d.out_cs;
        a.out_zs = d.out_zs;
        a.logn = d.logn;
        a.loga = plan[pi];
        a.tw = c->get_tw(a.loga, true);
        a.scale = last ? d.scale : 1;
        if (last) {
            a.stab = d.stab;
            a.stab_split = d.stab_split;
            a.stab_zs = d.stab_zs;
            a.npeer = d.npeer;
            for (int p = 0; p < d.npeer; p++) a.peer_out[p] = d.peer_out[p];
        }
        if (first) {
            int lognb = d.logn - a.loga;
            a.gather = d.natural_input ? 1 : 0;
            int want = a.gather ? 3 : 0;
            a.logq = std::min(lognb, std::max(want, ntt_tile_log() - a.loga));
            size_t tiles = (size_t)1 << (lognb - a.logq);
            a.nz = d.nz;
            dim3 grid = make_grid(a, tiles * d.nz, d.ncols);
            int th = pick_threads((size_t)1 << (a.loga + a.logq));
            const bool fixed = ntt_fixed() && ntt_rmax() == 3 && th == 512;
            if (fixed && a.loga == 10 && a.logq == 3) k_pass_contig<true, 3, 10, 3, 512><<<grid, 512, contig_smem(10, 3), c->stream>>>(a);
            else if (ntt_rmax() == 4) k_pass_contig<true, 4><<<grid, th, contig_smem(a.loga, a.logq), c->stream>>>(a);
            else k_pass_contig<true, 3><<<grid, th, contig_smem(a.loga, a.logq), c->stream>>>(a);
        } else {
            a.logB = done + a.loga;
            int logS = a.logB - a.loga;
            a.logq = std::min(logS, std::max(3, ntt_tile_log() - a.loga));
            a.twist = c->get_twist_full(a.logB, a.loga, true);
            size_t tiles = (size_t)1 << (d.logn - a.loga - a.logq);
            a.nz = d.nz;
            dim3 grid = make_grid(a, tiles * d.nz, d.ncols);
            int th = pick_threads((size_t)1 << (a.loga + a.logq));
            const bool fixed = ntt_fixed() && ntt_rmax() == 3 && th == 512;
            if (fixed && a.loga == 10 && a.logq == 3) k_pass_strided<true, 3, 10, 3, 512><<<grid, 512, strided_smem(10, 3), c->stream>>>(a);
            else if (ntt_rmax() == 4) k_pass_strided<true, 4><<<grid, th, strided_smem(a.loga, a.logq), c->stream>>>(a);
            else k_pass_strided<true, 3><<<grid, th, strided_smem(a.loga, a.logq), c->stream>>>(a);
        }
        count_launch(c);
        done += plan[pi];
    }
    CUDA_CHECK(cudaGetLastError());
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// table caches
// ---------------------------------------------------------------------------------------------------------------------
// Tables are filled on the root's stream and the stream is drained before a new table is published, so any stream may use it.
const u64* DevCtx::get_tw(int log, bool inverse) {
    if (parent) return parent->get_tw(log, inverse);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(log, (int)inverse);
    auto it = tw.find(key);
    if (it != tw.end()) return it->second.p;
    int A = 1 << log;
    dbuf<u64> t((size_t)std::max(A, 2));
    u64 root = gl_root_of_unity(log);
    if (inverse) root = gl_inv(root);
    k_fill_tw<<<(A + 255) / 256, 256, 0, stream>>>(t.p, log, root);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream));
    const u64* p = t.p;
    tw.emplace(key, std::move(t));
    return p;
}

const u64* DevCtx::get_twist(int logB, bool inverse, int* split) {
    if (parent) return parent->get_twist(logB, inverse, split);
    std::lock_guard<std::mutex> lk(mu);
    int sp = (logB + 1) / 2;
    *split = sp;
    auto key = std::make_pair(logB, (int)inverse);
    auto it = twist.find(key);
    if (it != twist.end()) return it->second.p;
    u32 nlo = 1u << sp, nhi = 1u << (logB - sp);
    dbuf<u64> t((size_t)nlo + nhi);
    u64 root = gl_root_of_unity(logB);
    if (inverse) root = gl_inv(root);
    k_fill_pow<<<(nlo + 255) / 256, 256, 0, stream>>>(t.p, nlo, root, 1, 1);
    k_fill_pow<<<(nhi + 255) / 256, 256, 0, stream>>>(t.p + nlo, nhi, root, (u64)nlo, 1);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream));
    const u64* p = t.p;
    twist.emplace(key, std::move(t));
    return p;
}

const u64* DevCtx::get_twist_full(int logB, int loga, bool inverse) {
    if (parent) return parent->get_twist_full(logB, loga, inverse);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(logB * 64 + loga, (int)inverse + 2);   // shares the `twist` map with the two-level tables
    auto it = twist.find(key);
    if (it != twist.end()) return it->second.p;
    dbuf<u64> t((size_t)1 << logB);
    u64 root = gl_root_of_unity(logB);
    if (inverse) root = gl_inv(root);
    k_fill_twist<<<(unsigned)((((size_t)1 << logB) + 255) / 256), 256, 0, stream>>>(t.p, logB, loga, root);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream));
    const u64* p = t.p;
    twist.emplace(key, std::move(t));
    return p;
}

const u64* DevCtx::get_powtab(int logn, u64 base, u64 premul, int* split) {
    if (parent) return parent->get_powtab(logn, base, premul, split);
    std::lock_guard<std::mutex> lk(mu);
    int sp = (logn + 1) / 2;
    *split = sp;
    auto key = std::make_tuple(logn, base, premul);
    auto it = powtab.find(key);
    if (it != powtab.end()) return it->second.p;
    u32 nlo = 1u << sp, nhi = 1u << (logn - sp);
    dbuf<u64> t((size_t)nlo + nhi);
    k_fill_pow<<<(nlo + 255) / 256, 256, 0, stream>>>(t.p, nlo, base, 1, 1);
    k_fill_pow<<<(nhi + 255) / 256, 256, 0, stream>>>(t.p + nlo, nhi, base, (u64)nlo, premul);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(stream));
    const u64* p = t.p;
    powtab.emplace(key, std::move(t));
    return p;
}

const u64* DevCtx::get_coset_tabs(int logn, int rate_bits, u64 shift, int* split_out, size_t* tab_len_out) {
    if (parent) return parent->get_coset_tabs(logn, rate_bits, shift, split_out, tab_len_out);
    // coset z covers leaves [z*N, (z+1)*N) = shift * omega_{N*nz}^{bitrev(z)} * <omega_N>; table z holds s_z^i for every i < N
    // (one load + one multiply per element in the first LDE pass; 8 N words per LDE shape, 64 MB at N = 2^20)
    const int nz = 1 << rate_bits;
    const size_t tab_len = (size_t)1 << logn;
    *split_out = 0;
    *tab_len_out = tab_len;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(logn + 64 * (rate_bits + 1), shift, (u64)0);
    auto it = powtab.find(key);
    if (it == powtab.end()) {
        const u64 wl = gl_root_of_unity(logn + rate_bits);
        dbuf<u64> t(tab_len * nz);
        for (int z = 0; z < nz; z++) {
            u64 s = gl_mul(shift, gl_pow(wl, bitrev32((u32)z, rate_bits)));
            k_fill_pow<<<(unsigned)((tab_len + 255) / 256), 256, 0, stream>>>(t.p + z * tab_len, (u32)tab_len, s, 1, 1);
        }
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaStreamSynchronize(stream));
        it = powtab.emplace(key, std::move(t)).first;
    }
    return it->second.p;
}

// ---------------------------------------------------------------------------------------------------------------------
// public transforms
// ---------------------------------------------------------------------------------------------------------------------
void ntt_ifft(DevCtx* c, const u64* d_values, size_t in_cs, u64* d_coeffs, size_t out_cs, int logn, int ncols, int npeer,
              u64* const* peer_coeffs) {
    if (ncols <= 0) return;
    StageTimer tm(c, &c->ntt_ms);
    XformDesc d = {};
    d.npeer = logn > 0 ? npeer : 0;
    d.peer_out = peer_coeffs;
    d.in = d_values;
    d.in_cs = in_cs;
    d.out = d_coeffs;
    d.out_cs = out_cs;
    d.logn = logn;
    d.ncols = ncols;
    d.nz = 1;
    d.scale = gl_inv((u64)1 << logn);
    d.natural_input = true;
    if (logn == 0) {
        CUDA_CHECK(cudaMemcpy2DAsync(d_coeffs, out_cs * 8, d_values, in_cs * 8, 8, ncols, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    run_inverse(c, d);
    c->ntt_bytes += 16.0 * (double)((size_t)1 << logn) * ncols;
}

void ntt_lde(DevCtx* c, const u64* d_coeffs, size_t in_cs, u64* d_lde, size_t out_cs, int logn, int rate_bits, int ncols,
             u64 shift, int z0, int nzl) {
    if (ncols <= 0) return;
    StageTimer tm(c, &c->ntt_ms);
    const int nz = 1 << rate_bits;
    if (nzl <= 0) {   // all cosets
        z0 = 0;
        nzl = nz;
    }
    const size_t n = (size_t)1 << logn;
    int split;
    size_t tab_len;
    const u64* tabs = c->get_coset_tabs(logn, rate_bits, shift, &split, &tab_len);
    XformDesc d = {};
    d.in = d_coeffs;
    d.in_cs = in_cs;
    d.in_zs = 0;
    d.out = d_lde;
    d.out_cs = out_cs;
    d.out_zs = n;
    d.logn = logn;
    d.ncols = ncols;
    d.nz = nzl;                                   // this rank's cosets [z0, z0 + nzl): leaves [z0 N, (z0 + nzl) N)
    d.stab = tabs + (size_t)z0 * tab_len;
    d.stab_split = split;
    d.stab_zs = tab_len;
    d.stab_full = 1;
    d.scale = 1;
    if (logn == 0) {
        // constant polynomials: every coset value equals the coefficient
        for (int z = 0; z < nzl; z++)
            CUDA_CHECK(cudaMemcpy2DAsync(d_lde + z, out_cs * 8, d_coeffs, in_cs * 8, 8, ncols, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    {
        StageTimer tl(c, &c->lde_ms);
        unsigned before = c->launches;
        run_forward(c, d);
        c->lde_launches += c->launches - before;
    }
    c->lde_bytes += (8.0 + 8.0 * nzl) * (double)n * ncols;
    c->ntt_bytes += (8.0 + 8.0 * nzl) * (double)n * ncols;
}

void ntt_coset_ifft_leaforder(DevCtx* c, u64* d_data, size_t cs, int logn, int ncols, u64 shift) {
    if (ncols <= 0 || logn == 0) return;
    StageTimer tm(c, &c->ntt_ms);
    XformDesc d = {};
    d.in = d_data;
    d.in_cs = cs;
    d.out = d_data;
    d.out_cs = cs;
    d.logn = logn;
    d.ncols = ncols;
    d.nz = 1;
    d.natural_input = false;
    d.scale = 1;
    // coefficient i is multiplied by shift^-i / N
    d.stab = c->get_powtab(logn, gl_inv(shift), gl_inv((u64)1 << logn), &d.stab_split);
    run_inverse(c, d);
    c->ntt_bytes += 16.0 * (double)((size_t)1 << logn) * ncols;
}
