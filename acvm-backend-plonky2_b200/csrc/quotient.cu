// Vanishing-polynomial / quotient evaluation on the LDE coset, sm_100a: one thread per LDE point, every constraint folded
// into the two alpha-weighted sums as it is produced (no per-point constraint vector), then divided by Z_H.
//
// Replaces plonky2 0.2.2 plonk/prover.rs compute_quotient_polys + plonk/vanishing_poly.rs eval_vanishing_poly_base_batch +
// evaluate_gate_constraints_base_batch, which call back into the reference's custom gates
// (plonky2-backend/src/plonky2_ecdsa/biguint/gates/*.rs; see gates.cuh).  SURVEY.md App. A.8.
//
// Reads are column-major in leaf order: thread j touches [col][j], so every wire / constant / sigma / Z column load of a
// warp is one coalesced 256-byte run.  "next row" of Z is index +8 in natural order = same coset, bit-reversed neighbour.
#include "gates.cuh"
#include "internal.h"
#include "quotient.h"

#include <stdlib.h>
#include <string.h>

// The per-proof parameters (alpha powers, beta k_i, gate table: ~13 KB) travel as a __grid_constant__ kernel parameter: they
// sit in the constant bank (uniform, broadcast reads) and nothing is shared between proofs in flight on other streams.
namespace {

struct WeightedSink {  // folds constraint k into sum_c += v * alpha_c^(off + k); products accumulated unreduced (gl_acc)
    gl_acc a0, a1;
    int k, base;
    const u64 *ap0, *ap1;   // alpha_0^k, alpha_1^k tables inside the kernel's parameter block
    __device__ __forceinline__ void seek(int idx) { k = base + idx; }
    __device__ __forceinline__ void emit(u64 v) {   // v: class N
        a0.mac(v, ap0[k]);
        a1.mac(v, ap1[k]);
        k++;
    }
};

// Gate-major evaluation: one launch per gate of the circuit, each a small specialised kernel (template on the gate kind,
// so its SASS holds exactly one gate's code and stays in the instruction cache) that streams only the wire columns the
// gate reads -- thread j touches [col][j], fully coalesced -- and accumulates filter * sum_k alpha^k constraint_k into the
// two running sums.  A first kernel seeds the sums with the Z and permutation (partial-product) terms; the last gate
// kernel divides by Z_H.  (A single fused kernel holding all 12 gates was measured first: 307 ms at 2^20 rows, 30% of the
// warp time stalled on instruction fetch and 38% on a load-imbalance barrier; see profiles/.)
// Sharding: a rank evaluates the `npts` leaves [j0, j0 + npts) (whole cosets).  Input columns are local (stride L = npts,
// index j); xs / l0s are the full per-circuit tables (index j0 + j); `out` is the full-size [NC][OL] quotient-value buffer.
// CNC / CR / CCH > 0: num_challenges, num_routed and the chunk size as compile-time constants (2 / 80 / 8 = the reference's
// configuration): loop bounds fold, the boundary predicates of the 4-wide load batches disappear.
template <int CNC, int CR, int CCH>
__global__ void __launch_bounds__(256) k_quotient_perm(const __grid_constant__ QuotientParams P, const u64* __restrict__ cs,
                                                       const u64* __restrict__ wires,
                                                       const u64* __restrict__ zpp, const u64* __restrict__ xs,
                                                       const u64* __restrict__ l0s, u64* __restrict__ out_, int scale_now,
                                                       size_t L, size_t j0, size_t OL) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    u64* __restrict__ out = out_ + j0;
    xs += j0;
    l0s += j0;
    const int NC = CNC ? CNC : P.num_challenges, R = CR ? CR : P.num_routed, C = P.num_constants;
    const int NPP = (CR && CCH) ? (CR + CCH - 1) / CCH - 1 : P.num_partial_products;
    // next row of Z: natural index + 2^rate_bits  <=>  same coset, k -> k + 1 (k = bitrev_n(j mod N))
    const u32 nmask = (1u << P.logn) - 1;
    const u32 k = bitrev32((u32)j & nmask, P.logn);
    const size_t jn = (j & ~(size_t)nmask) | bitrev32((k + 1) & nmask, P.logn);
    const u64 x = xs[j], l0 = l0s[j];
    const int chunk = CCH ? CCH : P.qdf;
    gl_acc acc0, acc1;   // the two alpha-weighted sums, one reduction each at the end
    acc0.clear();
    acc1.clear();
    // Both challenges in one sweep: every wire / sigma column value is loaded once and feeds the running products of
    // challenge 0 and challenge 1 (ncu: the two-sweep version read 23 GB for 11.6 GB of columns).
    u64 prev[2], pn[2], pd[2];
    for (int c = 0; c < NC; c++) {
        prev[c] = zpp[(size_t)c * L + j];
        u64 tz = glz_mul(l0, gl_sub(prev[c], 1));   // L_0(x) (Z_c(x) - 1)
        acc0.mac(tz, P.apow[0][c]);
        acc1.mac(tz, P.apow[1][c]);
    }
    for (int m = 0; m <= NPP; m++) {
        pn[0] = pn[1] = pd[0] = pd[1] = 1;
        const int lo = m * chunk, hi = min((m + 1) * chunk, R);
        for (int r0 = lo; r0 < hi; r0 += 4) {   // batches of 4 + 4 independent loads
            u64 wv[4], sv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                wv[i] = (r0 + i < hi) ? wires[(size_t)(r0 + i) * L + j] : 0;
                sv[i] = (r0 + i < hi) ? cs[(size_t)(C + r0 + i) * L + j] : 0;
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = r0 + i;
                if (r < hi) {
                    for (int c = 0; c < NC; c++) {
                        u64 base = gl_add(wv[i], P.gammas[c]);
                        pn[c] = glz_mul(pn[c], gl_add(base, gl_mul(P.beta_k[c][r], x)));
                        pd[c] = glz_mul(pd[c], gl_add(base, gl_mul(P.betas[c], sv[i])));
                    }
                }
            }
        }
        for (int c = 0; c < NC; c++) {
            const int term = NC + c * (NPP + 1);
            u64 next = (m < NPP) ? zpp[(size_t)(NC + c * NPP + m) * L + j] : zpp[(size_t)c * L + jn];
            u64 tv = gl_sub(glz_mul(prev[c], pn[c]), gl_mul(next, pd[c]));
            acc0.mac(tv, P.apow[0][term + m]);
            acc1.mac(tv, P.apow[1][term + m]);
            prev[c] = next;
        }
    }
    out[j] = acc0.reduce();
    if (NC > 1) out[OL + j] = acc1.reduce();
    if (scale_now) {
        const u64 zhi = P.zh_inv[bitrev32((u32)((j0 + j) >> P.logn), P.rate_bits)];
        out[j] = gl_mul(out[j], zhi);
        if (NC > 1) out[OL + j] = gl_mul(out[OL + j], zhi);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_quotient_gate(const __grid_constant__ QuotientParams P, const u64* __restrict__ cs,
                                                       const u64* __restrict__ wires, u64* __restrict__ out_, int g, u32 op_lo,
                                                       u32 op_hi, int scale_now, size_t L, size_t j0, size_t OL) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    u64* __restrict__ out = out_ + j0;
    const GateDev& gd = P.gates[g];
    auto wire = [&](int i) -> u64 { return __ldg(wires + (size_t)i * L + j); };
    auto konst = [&](int i) -> u64 { return __ldg(cs + (size_t)(P.num_selectors + i) * L + j); };
    const int off = P.num_challenges * (2 + P.num_partial_products);
    u64 f = gate_filter(gd, g, cs[(size_t)gd.selector_index * L + j], P.num_selectors > 1);
    WeightedSink sink;
    sink.a0.clear();
    sink.a1.clear();
    sink.k = sink.base = off;
    sink.ap0 = P.apow[0];
    sink.ap1 = P.apow[1];
    eval_gate_kind<KIND>(gd, op_lo, op_hi, wire, konst, P.pi_hash, sink);
    u64 a0 = gl_add(out[j], gl_mul(f, sink.a0.reduce()));
    u64 a1 = P.num_challenges > 1 ? gl_add(out[OL + j], gl_mul(f, sink.a1.reduce())) : 0;
    if (scale_now) {
        const u64 zhi = P.zh_inv[bitrev32((u32)((j0 + j) >> P.logn), P.rate_bits)];
        a0 = gl_mul(a0, zhi);
        a1 = gl_mul(a1, zhi);
    }
    out[j] = a0;
    if (P.num_challenges > 1) out[OL + j] = a1;
}

// Limb sweep + the rest of the u32 gates, one pass over the wires (quotient.h LimbPlan).  Replaces the k_quotient_gate<7..10>
// launches (and BaseSum<4>): those evaluated l(l-1)(l-2)(l-3) once per (gate, limb) -- 1300 times per point for the ECDSA mix
// on 234 wires -- and re-read the limb columns once per gate.  Here a point's thread walks the wire axis once: check(w), up to
// four multiply-accumulates into the running prefixes, and at every window boundary one `filter * prefix * coef` term.  The
// gates' remaining constraints (recombinations, carries) follow in the same kernel while the block's columns are L2-resident.
#define SWEEP_B 4
template <int MINB, int NTH>
__global__ void __launch_bounds__(NTH, MINB) k_quotient_limb(const __grid_constant__ QuotientParams P, const __grid_constant__ LimbPlan LP,
                                                       const u64* __restrict__ cs, const u64* __restrict__ wires,
                                                       u64* __restrict__ out_, int scale_now, size_t L, size_t j0, size_t OL) {
    __shared__ u64 fsh[P2G_MAX_LIMB_GATES][NTH];
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    u64* __restrict__ out = out_ + j0;
    const int tid = threadIdx.x;
    const bool many = P.num_selectors > 1;
    for (int s = 0; s < LP.ngates; s++) {
        const int g = LP.gate[s];
        const GateDev& gd = P.gates[g];
        fsh[s][tid] = gate_filter(gd, g, __ldg(cs + (size_t)gd.selector_index * L + j), many);
    }
    gl_acc acc0, acc1, A0, A1, B0, B1;
    acc0.clear(); acc1.clear(); A0.clear(); A1.clear(); B0.clear(); B1.clear();
    int e = 0;
    const int ne = LP.nevents;
    auto events_at = [&](int w) {   // boundary events at wire position w (the prefixes cover wires < w)
        if (e < ne && LP.ev[e].pos == w) {
            int cur_dir = -1;
            u64 r0 = 0, r1 = 0;
            do {
                const LimbEvent& E = LP.ev[e];
                if ((int)E.dir != cur_dir) {
                    cur_dir = E.dir;
                    r0 = cur_dir ? B0.reduce() : A0.reduce();
                    r1 = cur_dir ? B1.reduce() : A1.reduce();
                }
                const u64 f = fsh[E.slot][tid];
                acc0.mac(glz_mul(f, r0), E.coef[0]);
                acc1.mac(glz_mul(f, r1), E.coef[1]);
                e++;
            } while (e < ne && LP.ev[e].pos == w);
        }
    };
    auto wire = [&](int i) -> u64 { return __ldg(wires + (size_t)i * L + j); };
    const int off = P.num_challenges * (2 + P.num_partial_products);
    for (int ph = 0; ph < LP.nphases; ph++) {
        // 1. sweep the strip, SWEEP_B independent loads at a time (one load per iteration leaves the thread waiting on DRAM)
        const int w_end = LP.cut[ph + 1];
        for (int w0 = LP.cut[ph]; w0 < w_end; w0 += SWEEP_B) {
            u64 v[SWEEP_B];
#pragma unroll
            for (int i = 0; i < SWEEP_B; i++) v[i] = (w0 + i < w_end && LP.need[w0 + i]) ? wire(w0 + i) : 0;
            // the SWEEP_B checks are independent of each other and of the event bookkeeping below: computing them up front gives
            // the scheduler four multiply chains to interleave
#pragma unroll
            for (int i = 0; i < SWEEP_B; i++) v[i] = limb4_check(v[i]);
#pragma unroll
            for (int i = 0; i < SWEEP_B; i++) {
                const int w = w0 + i;
                if (w >= w_end) break;
                events_at(w);
                const unsigned need = LP.need[w];
                if (need) {
                    const u64 ck = v[i];
                    if (need & 1) {
                        A0.mac(ck, P.apow[0][w]);
                        A1.mac(ck, P.apow[1][w]);
                    }
                    if (need & 2) {
                        B0.mac(ck, LP.bpow[0][w]);
                        B1.mac(ck, LP.bpow[1][w]);
                    }
                }
            }
        }
        if (ph + 1 == LP.nphases) events_at(LP.wmax);
        // 2. everything else the ops of this strip constrain (recombinations, carries, borrows)
        for (int s = 0; s < LP.ngates; s++) {
            const u32 op_lo = LP.oplo[s][ph], op_hi = LP.oplo[s][ph + 1];
            if (op_hi <= op_lo) continue;
            const GateDev& gd = P.gates[LP.gate[s]];
            WeightedSink sink;
            sink.a0.clear();
            sink.a1.clear();
            sink.k = sink.base = off;
            sink.ap0 = P.apow[0];
            sink.ap1 = P.apow[1];
            switch (gd.kind) {
            case P2G_GATE_BASE_SUM: eval_gate_nonlimb<P2G_GATE_BASE_SUM>(gd, op_lo, op_hi, wire, sink); break;
            case P2G_GATE_U32_ARITHMETIC: eval_gate_nonlimb<P2G_GATE_U32_ARITHMETIC>(gd, op_lo, op_hi, wire, sink); break;
            case P2G_GATE_U32_ADD_MANY: eval_gate_nonlimb<P2G_GATE_U32_ADD_MANY>(gd, op_lo, op_hi, wire, sink); break;
            case P2G_GATE_U32_SUBTRACTION: eval_gate_nonlimb<P2G_GATE_U32_SUBTRACTION>(gd, op_lo, op_hi, wire, sink); break;
            case P2G_GATE_U32_RANGE_CHECK: eval_gate_nonlimb<P2G_GATE_U32_RANGE_CHECK>(gd, op_lo, op_hi, wire, sink); break;
            default: break;
            }
            const u64 f = fsh[s][tid];
            acc0.mac(f, sink.a0.reduce());
            acc1.mac(f, sink.a1.reduce());
        }
    }
    u64 a0 = gl_add(out[j], acc0.reduce());
    u64 a1 = P.num_challenges > 1 ? gl_add(out[OL + j], acc1.reduce()) : 0;
    if (scale_now) {
        const u64 zhi = P.zh_inv[bitrev32((u32)((j0 + j) >> P.logn), P.rate_bits)];
        a0 = gl_mul(a0, zhi);
        a1 = gl_mul(a1, zhi);
    }
    out[j] = a0;
    if (P.num_challenges > 1) out[OL + j] = a1;
}

struct StoreSink {  // stand-alone entry point: out[k][pt] += filter * constraint_k
    u64* out;
    size_t np, pt;
    u64 filter;
    int k;
    __device__ __forceinline__ void seek(int idx) { k = idx; }
    __device__ __forceinline__ void emit(u64 v) {
        u64* o = out + (size_t)k * np + pt;
        *o = gl_add(*o, gl_mul(v, filter));
        k++;
    }
};

__global__ void __launch_bounds__(128) k_eval_gates(const __grid_constant__ QuotientParams P, const u64* __restrict__ consts,
                                                    const u64* __restrict__ wires, u64* out, size_t np) {
    size_t pt = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pt >= np) return;
    auto wire = [&](int i) -> u64 { return wires[(size_t)i * np + pt]; };
    auto konst = [&](int i) -> u64 { return consts[(size_t)(P.num_selectors + i) * np + pt]; };
    const bool many = P.num_selectors > 1;
    for (int g = 0; g < P.num_gates; g++) {
        const GateDev& gd = P.gates[g];
        StoreSink sink = {out, np, pt, gate_filter(gd, g, consts[(size_t)gd.selector_index * np + pt], many), 0};
        eval_gate_unfiltered(gd, 0, gate_num_ops(gd.kind, gd.params), wire, konst, P.pi_hash, sink);
    }
}

// x_j = shift * omega_{8N}^{bitrev(j)} and L_0(x_j) = (x_j^N - 1) / (N (x_j - 1)) for every leaf j, once per circuit
__global__ void k_points(u64* xs, u64* l0s, int logn, int rate_bits, u64 shift, u64 w_lde, u64 n_as_field, const u64* zh) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t L = (size_t)1 << (logn + rate_bits);
    if (j >= L) return;
    u64 x = gl_mul(shift, gl_pow(w_lde, bitrev32((u32)j, logn + rate_bits)));
    xs[j] = x;
    u64 z = zh[bitrev32((u32)(j >> logn), rate_bits)];
    l0s[j] = gl_mul(z, gl_inv(gl_mul(n_as_field, gl_sub(x, 1))));
}

}  // namespace


void quotient_points(DevCtx* c, u64* d_xs, u64* d_l0s, int logn, int rate_bits, const u64* h_zh) {
    size_t L = (size_t)1 << (logn + rate_bits);
    dbuf<u64> zh(1 << rate_bits);
    CUDA_CHECK(cudaMemcpyAsync(zh.p, h_zh, 8 << rate_bits, cudaMemcpyHostToDevice, c->stream));
    k_points<<<(unsigned)((L + 255) / 256), 256, 0, c->stream>>>(d_xs, d_l0s, logn, rate_bits, GL_GEN,
                                                                  gl_root_of_unity(logn + rate_bits), ((u64)1 << logn) % GL_P,
                                                                  zh.p);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    count_launch(c);
}

bool quotient_limb_plan(const QuotientParams& qp, int num_wires, const u64* alphas, LimbPlan* lp) {
    memset(lp, 0, sizeof *lp);
    if (getenv("P2G_QUOTIENT_GATE_MAJOR") || num_wires > P2G_MAX_WIRES) return false;
    const int NC = qp.num_challenges, off = NC * (2 + qp.num_partial_products);
    u64 ainv[2] = {0, 0};
    for (int c = 0; c < NC; c++) {
        if (alphas[c] == 0) return false;
        ainv[c] = gl_inv(alphas[c]);
    }
    auto apow = [&](int c, long e) { return e >= 0 ? gl_pow(alphas[c], (u64)e) : gl_pow(ainv[c], (u64)(-e)); };
    struct Win { int slot, lo, len, k0, dir; };   // limb j of the window sits on wire lo + j, constraint k0 + j (dir 0) or k0 - j (dir 1)
    std::vector<Win> wins;
    for (int g = 0; g < qp.num_gates; g++) {
        const GateDev& gd = qp.gates[g];
        const u32* p = gd.params;
        if (!gd.num_constraints || !gate_has_limb4_sweep(gd.kind, p)) continue;
        if (lp->ngates == P2G_MAX_LIMB_GATES) return false;
        const int slot = lp->ngates++;
        lp->gate[slot] = g;
        switch (gd.kind) {
        case P2G_GATE_BASE_SUM: wins.push_back({slot, 1, (int)p[1], 1, 0}); break;
        case P2G_GATE_U32_ARITHMETIC:
            for (u32 i = 0; i < p[0]; i++) wins.push_back({slot, (int)(6 * p[0] + 32 * i), 32, (int)(36 * i + 33), 1});
            break;
        case P2G_GATE_U32_ADD_MANY:
            for (u32 i = 0; i < p[1]; i++) wins.push_back({slot, (int)((p[0] + 3) * p[1] + 18 * i), 18, (int)(21 * i + 18), 1});
            break;
        case P2G_GATE_U32_SUBTRACTION:
            for (u32 i = 0; i < p[0]; i++) wins.push_back({slot, (int)(5 * p[0] + 16 * i), 16, (int)(19 * i + 16), 1});
            break;
        case P2G_GATE_U32_RANGE_CHECK:
            for (u32 i = 0; i < p[0]; i++) wins.push_back({slot, (int)(p[0] + 16 * i), 16, (int)(17 * i + 1), 0});
            break;
        default: break;
        }
    }
    if (wins.empty()) return false;
    // boundary events, merged per (position, direction, gate)
    std::map<std::tuple<int, int, int>, std::pair<u64, u64>> evs;
    lp->wmin = num_wires;
    lp->wmax = 0;
    for (const Win& wn : wins) {
        if (wn.lo + wn.len > num_wires) return false;
        u64 cf[2] = {0, 0};
        for (int c = 0; c < NC; c++) cf[c] = apow(c, wn.dir ? (long)off + wn.k0 + wn.lo : (long)off + wn.k0 - wn.lo);
        auto& hi = evs[std::make_tuple(wn.lo + wn.len, wn.dir, wn.slot)];
        auto& lo = evs[std::make_tuple(wn.lo, wn.dir, wn.slot)];
        hi.first = gl_add(hi.first, cf[0]);
        hi.second = gl_add(hi.second, cf[1]);
        lo.first = gl_sub(lo.first, cf[0]);
        lo.second = gl_sub(lo.second, cf[1]);
        for (int w = wn.lo; w < wn.lo + wn.len; w++) lp->need[w] |= wn.dir ? 2 : 1;
        lp->wmin = std::min(lp->wmin, wn.lo);
        lp->wmax = std::max(lp->wmax, wn.lo + wn.len);
    }
    // strips: equal shares of the swept range; an op belongs to the strip in which its limb window ends
    {
        int nph = P2G_LIMB_PHASES;
        if (const char* e = getenv("P2G_LIMB_PHASES")) nph = std::max(1, std::min(P2G_LIMB_PHASES, atoi(e)));
        const int span = lp->wmax - lp->wmin;
        nph = std::max(1, std::min(nph, span / 16));
        lp->nphases = nph;
        for (int p = 0; p <= nph; p++) lp->cut[p] = lp->wmin + (int)((long)span * p / nph);
        std::vector<std::vector<int>> cnt(lp->ngates, std::vector<int>(nph, 0));
        for (const Win& wn : wins) {
            int p = 0;
            while (p + 1 < nph && wn.lo + wn.len > lp->cut[p + 1]) p++;
            cnt[wn.slot][p]++;
        }
        for (int s = 0; s < lp->ngates; s++) {
            int acc = 0;
            for (int p = 0; p < nph; p++) {
                lp->oplo[s][p] = (unsigned char)acc;
                acc += cnt[s][p];
            }
            if (acc > 255) return false;
            lp->oplo[s][nph] = (unsigned char)acc;
        }
    }
    for (auto& kv : evs) {
        if (kv.second.first == 0 && kv.second.second == 0) continue;
        if (lp->nevents == P2G_MAX_LIMB_EVENTS) return false;
        LimbEvent& E = lp->ev[lp->nevents++];
        E.pos = (unsigned short)std::get<0>(kv.first);
        E.dir = (unsigned char)std::get<1>(kv.first);
        E.slot = (unsigned char)std::get<2>(kv.first);
        E.coef[0] = kv.second.first;
        E.coef[1] = kv.second.second;
    }
    for (int c = 0; c < NC; c++) {
        u64 x = 1;
        for (int w = 0; w < num_wires; w++) {
            lp->bpow[c][w] = x;
            x = gl_mul(x, ainv[c]);
        }
    }
    return true;
}

void quotient_eval(DevCtx* c, const QuotientParams& qp, const LimbPlan* lp, const u64* d_cs, const u64* d_wires, const u64* d_zpp,
                   const u64* d_xs, const u64* d_l0s, u64* d_out, size_t npts, size_t j0, size_t out_stride) {
    const int TH = 256;
    const unsigned grid = (unsigned)((npts + TH - 1) / TH);
    // launch list: gates with constraints that the limb sweep does not cover, then the sweep; the last launch divides by Z_H
    std::vector<int> solo;
    for (int g = 0; g < qp.num_gates; g++) {
        const GateDev& gd = qp.gates[g];
        if (!gd.num_constraints) continue;
        if (lp && gate_has_limb4_sweep(gd.kind, gd.params)) continue;
        solo.push_back(g);
    }
    const bool sweep = lp != nullptr;
    if (qp.num_challenges == 2 && qp.num_routed == 80 && qp.qdf == 8 && qp.num_partial_products == 9)
        k_quotient_perm<2, 80, 8><<<grid, TH, 0, c->stream>>>(qp, d_cs, d_wires, d_zpp, d_xs, d_l0s, d_out, solo.empty() && !sweep, npts, j0, out_stride);
    else
        k_quotient_perm<0, 0, 0><<<grid, TH, 0, c->stream>>>(qp, d_cs, d_wires, d_zpp, d_xs, d_l0s, d_out, solo.empty() && !sweep, npts, j0, out_stride);
    count_launch(c);
    for (size_t i = 0; i < solo.size(); i++) {
        const int g = solo[i];
        const GateDev& gd = qp.gates[g];
        const u32 nops = gate_num_ops(gd.kind, gd.params);
        const int fin = (i + 1 == solo.size()) && !sweep;
        switch (gd.kind) {
#define P2G_LAUNCH(KIND) case KIND: k_quotient_gate<KIND><<<grid, TH, 0, c->stream>>>(qp, d_cs, d_wires, d_out, g, 0, nops, fin, npts, j0, out_stride); break;
            P2G_LAUNCH(P2G_GATE_CONSTANT) P2G_LAUNCH(P2G_GATE_PUBLIC_INPUT) P2G_LAUNCH(P2G_GATE_ARITHMETIC)
            P2G_LAUNCH(P2G_GATE_BASE_SUM) P2G_LAUNCH(P2G_GATE_POSEIDON) P2G_LAUNCH(P2G_GATE_RANDOM_ACCESS)
            P2G_LAUNCH(P2G_GATE_U32_ARITHMETIC) P2G_LAUNCH(P2G_GATE_U32_ADD_MANY) P2G_LAUNCH(P2G_GATE_U32_SUBTRACTION)
            P2G_LAUNCH(P2G_GATE_U32_RANGE_CHECK) P2G_LAUNCH(P2G_GATE_COMPARISON)
#undef P2G_LAUNCH
        default: throw p2g_error(P2G_EBADARG, "quotient: unknown gate kind");
        }
        count_launch(c);
    }
    if (sweep) {
        static const int minb = getenv("P2G_LIMB_MINB") ? atoi(getenv("P2G_LIMB_MINB")) : 2;
        // two blocks of 256 threads per SM at 126 registers.  Measured alternatives on B200 (tools/limb_sweep.sh): 3 x 256 at 80
        // registers spills (50 ms), 3 x 192 and 5 x 128 at 96 registers re-materialise (48 / 45 ms) against 34.5 ms here
        if (minb == 3) k_quotient_limb<3, 256><<<grid, TH, 0, c->stream>>>(qp, *lp, d_cs, d_wires, d_out, 1, npts, j0, out_stride);
        else k_quotient_limb<2, 256><<<grid, TH, 0, c->stream>>>(qp, *lp, d_cs, d_wires, d_out, 1, npts, j0, out_stride);
        count_launch(c);
    }
    CUDA_CHECK(cudaGetLastError());
}

void gates_eval_standalone(DevCtx* c, const QuotientParams& qp, const u64* d_consts, const u64* d_wires, u64* d_out, size_t npoints) {
    const int TH = 128;
    k_eval_gates<<<(unsigned)((npoints + TH - 1) / TH), TH, 0, c->stream>>>(qp, d_consts, d_wires, d_out, npoints);
    CUDA_CHECK(cudaGetLastError());
    count_launch(c);
}
