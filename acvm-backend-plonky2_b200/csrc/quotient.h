// Parameters of the quotient / gate-evaluation kernels, passed to every launch as a __grid_constant__ kernel parameter
// (constant bank: uniform, broadcast reads of alpha powers, beta*k_i, gate table).
#pragma once
#include "gates.cuh"

#define P2G_MAX_TERMS 512
#define P2G_MAX_GATES 64
#define P2G_MAX_ROUTED 128
#define P2G_MAX_RATE 5

struct QuotientParams {
    int logn, rate_bits, num_challenges, num_partial_products, num_routed, num_constants, num_selectors, num_gates, qdf;
    int pad_;
    u64 betas[2], gammas[2];
    u64 beta_k[2][P2G_MAX_ROUTED];   // beta_c * k_i
    u64 zh_inv[1 << P2G_MAX_RATE];   // 1 / Z_H on coset r (natural coset number)
    u64 pi_hash[4];
    u64 apow[2][P2G_MAX_TERMS];      // alpha_c^k
    GateDev gates[P2G_MAX_GATES];
};


// ---- limb-check sweep (quotient.cu k_quotient_limb) ------------------------------------------------------------------
// The u32 gates of the reference (plonky2_ecdsa/biguint/gates/{arithmetic_u32,add_many_u32,subtraction_u32,range_check_u32}.rs)
// and BaseSumGate<4> all range-check their 2-bit limbs with the same polynomial l(l-1)(l-2)(l-3), on overlapping wire ranges.
// check(wire w) is computed ONCE per wire and point and folded into alpha-weighted prefix sums over the wire axis,
//     A_c(m) = sum_{w < m} alpha_c^w  check(w)        (gates whose constraint index grows with the limb's wire index)
//     B_c(m) = sum_{w < m} alpha_c^-w check(w)        (gates that push their limbs in reverse order)
// so the limb constraints of one op (a window of consecutive wires with consecutive constraint indices) are
// coef * (prefix(end) - prefix(start)): one multiply per window instead of one per limb.  An event = "at wire position pos,
// add filter_g * coef_c * prefix_dir,c(pos) to the sum of challenge c".
#define P2G_MAX_LIMB_EVENTS 224
#define P2G_MAX_LIMB_GATES 16
#define P2G_MAX_WIRES 256
struct LimbEvent {
    unsigned short pos;   // wire position (the prefix covers wires < pos)
    unsigned char slot;   // index into LimbPlan::gate
    unsigned char dir;    // 0: A (alpha^w), 1: B (alpha^-w)
    u32 pad_;
    u64 coef[2];
};
// The wire axis is walked in up to P2G_LIMB_PHASES strips [cut[p], cut[p+1]): after a strip's sweep, the other constraints of the
// ops whose limbs end inside the strip are evaluated, so the limb columns are re-read while the strip is still L2-resident
// (one strip of the resident threads' columns fits the 126 MB L2; all 234 columns do not).
#define P2G_LIMB_PHASES 4
struct LimbPlan {
    int ngates, nevents, wmin, wmax;
    int nphases;
    int cut[P2G_LIMB_PHASES + 1];
    unsigned char oplo[P2G_MAX_LIMB_GATES][P2G_LIMB_PHASES + 1];   // ops [oplo[s][p], oplo[s][p+1]) of gate slot s belong to strip p
    int gate[P2G_MAX_LIMB_GATES];
    unsigned char need[P2G_MAX_WIRES];   // bit 0: wire feeds A, bit 1: wire feeds B
    u64 bpow[2][P2G_MAX_WIRES];          // alpha_c^-w
    LimbEvent ev[P2G_MAX_LIMB_EVENTS];   // sorted by pos
};
// fills `lp` for the circuit's gates and this proof's alphas; returns false when the sweep does not apply (no limb gate, an
// alpha without inverse, too many windows): the caller then evaluates those gates one by one as before
bool quotient_limb_plan(const QuotientParams& qp, int num_wires, const u64* alphas, LimbPlan* lp);

struct DevCtx;
void quotient_points(DevCtx* c, u64* d_xs, u64* d_l0s, int logn, int rate_bits, const u64* h_zh);
// evaluates leaves [j0, j0 + npts) (whole cosets; column stride of the inputs = npts) into d_out[c * out_stride + j0 + j];
// d_xs / d_l0s are the full tables
// lp: nullptr = every gate by its own kernel
void quotient_eval(DevCtx* c, const QuotientParams& qp, const LimbPlan* lp, const u64* d_cs, const u64* d_wires, const u64* d_zpp, const u64* d_xs,
                   const u64* d_l0s, u64* d_out, size_t npts, size_t j0, size_t out_stride);
void gates_eval_standalone(DevCtx* c, const QuotientParams& qp, const u64* d_consts, const u64* d_wires, u64* d_out, size_t npoints);
