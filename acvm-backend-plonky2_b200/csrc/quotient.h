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


struct DevCtx;
void quotient_points(DevCtx* c, u64* d_xs, u64* d_l0s, int logn, int rate_bits, const u64* h_zh);
// evaluates leaves [j0, j0 + npts) (whole cosets; column stride of the inputs = npts) into d_out[c * out_stride + j0 + j];
// d_xs / d_l0s are the full tables
void quotient_eval(DevCtx* c, const QuotientParams& qp, const u64* d_cs, const u64* d_wires, const u64* d_zpp, const u64* d_xs,
                   const u64* d_l0s, u64* d_out, size_t npts, size_t j0, size_t out_stride);
void gates_eval_standalone(DevCtx* c, const QuotientParams& qp, const u64* d_consts, const u64* d_wires, u64* d_out, size_t npoints);
