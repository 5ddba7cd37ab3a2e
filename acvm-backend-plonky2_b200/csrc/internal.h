// Internal interfaces of libp2g (not part of the C ABI in include/p2g.h).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/p2g.h"
#include "hash.cuh"

struct p2g_error : std::runtime_error {
    int code;
    p2g_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define CUDA_CHECK(expr)                                                                                              \
    do {                                                                                                              \
        cudaError_t e__ = (expr);                                                                                     \
        if (e__ != cudaSuccess)                                                                                       \
            throw p2g_error(e__ == cudaErrorMemoryAllocation ? P2G_ENOMEM : P2G_ECUDA,                                \
                            std::string(#expr) + ": " + cudaGetErrorString(e__));                                     \
    } while (0)

// Owning device buffer
template <class T>
struct dbuf {
    T* p = nullptr;
    size_t n = 0;
    dbuf() {}
    explicit dbuf(size_t count) { alloc(count); }
    dbuf(const dbuf&) = delete;
    dbuf& operator=(const dbuf&) = delete;
    dbuf(dbuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    dbuf& operator=(dbuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~dbuf() { release(); }
    void alloc(size_t count) {
        release();
        if (count) CUDA_CHECK(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

// Execution context: one stream, launch counters and stage timers, and the cached twiddle / twist / coset-power tables.
// There is one root context per device (stand-alone entry points) and one child per circuit handle: a child has its own
// stream and accounting -- proofs on different handles run concurrently -- and shares the root's tables.
struct DevCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevCtx* parent = nullptr;   // table owner (null for the root)
    std::mutex mu;
    // key: (log size, inverse) -> omega_{2^log}^{+-j}, j < 2^(log-1)        (smem-staged butterfly table)
    std::map<std::pair<int, int>, dbuf<u64>> tw;
    // key: (log B, inverse) -> [lo table 2^split | hi table 2^(b-split)] of omega_B^{+-k}   (twist between passes)
    std::map<std::pair<int, int>, dbuf<u64>> twist;
    // key: (log n, base, premul) -> [lo 2^split | hi 2^(n-split)] of premul * base^i        (coset scaling)
    std::map<std::tuple<int, u64, u64>, dbuf<u64>> powtab;
    // accounting of the current prove call
    unsigned launches = 0;
    double ntt_bytes = 0, merkle_bytes = 0;
    float ntt_ms = 0, merkle_ms = 0, leaf_ms = 0, lde_ms = 0, quot_ms = 0;
    double leaf_bytes = 0, lde_bytes = 0;
    unsigned leaf_launches = 0, lde_launches = 0;
    bool timing = false;  // when set, NTT / Merkle entry points bracket themselves with events (recorded, not waited on)
    std::vector<cudaEvent_t> ev_pool;                                       // reusable timing events
    std::vector<std::tuple<cudaEvent_t, cudaEvent_t, float*>> ev_pending;   // recorded pairs, read by resolve_timers()

    const u64* get_tw(int log, bool inverse);
    const u64* get_twist(int logB, bool inverse, int* split);
    const u64* get_twist_full(int logB, int loga, bool inverse);   // omega_B^{+-q bitrev_a(m)} at [(m << logS) + q]
    const u64* get_powtab(int logn, u64 base, u64 premul, int* split);
    // the 2^rate_bits per-coset index-power tables of an LDE: table z = [lo | hi] powers of shift * omega_{N 2^rate}^{bitrev(z)}
    const u64* get_coset_tabs(int logn, int rate_bits, u64 shift, int* split, size_t* tab_len);
};
DevCtx* get_ctx(int device);          // the device's root context
DevCtx* new_child_ctx(int device);    // own stream, shares the root's tables
void free_child_ctx(DevCtx* c);

void resolve_timers(DevCtx* c);   // after a stream synchronisation: adds every recorded segment to its accumulator
struct StageTimer {  // RAII: accumulates elapsed ms of a stream segment into *acc when ctx->timing is on (no host sync)
    DevCtx* c;
    float* acc;
    cudaEvent_t a = nullptr, b = nullptr;
    StageTimer(DevCtx* ctx, float* accum);
    ~StageTimer();
};

// ---- ntt.cu: transforms over column-major batches (column c at base + c*col_stride) ----
// values on <omega_N> in natural order -> coefficients in natural order (PolynomialValues::ifft)
// peer_coeffs (npeer pointers, may be 0): the same position in other GPUs' coefficient buffers (same out_cs); the last pass stores
// its results there too (P2P over NVLink) -- the fused "inverse NTT + all-gather" of coset-sharded proofs
#define P2G_MAX_PEERS 7
void ntt_ifft(DevCtx* c, const u64* d_values, size_t in_cs, u64* d_coeffs, size_t out_cs, int logn, int ncols, int npeer = 0,
              u64* const* peer_coeffs = nullptr);
// coefficients (natural, N) -> values on shift*<omega_{N<<rate_bits}> in LEAF order (index j <-> point shift*omega^bitrev(j))
// (PolynomialBatch::lde_values followed by reverse_index_bits_in_place)
// z0 / nzl: only the cosets [z0, z0 + nzl) are produced (coset-sharded proofs), at d_lde[col * out_cs + (z - z0) * N + i];
// nzl = 0 means all 2^rate_bits cosets.
void ntt_lde(DevCtx* c, const u64* d_coeffs, size_t in_cs, u64* d_lde, size_t out_cs, int logn, int rate_bits, int ncols,
             u64 shift, int z0 = 0, int nzl = 0);
// leaf-order values on shift*<omega_N> -> natural coefficients, in place (coset_ifft)
void ntt_coset_ifft_leaforder(DevCtx* c, u64* d_data, size_t cs, int logn, int ncols, u64 shift);

// ---- merkle.cu ----
struct MerkleTree {
    int log_leaves = 0, cap_height = 0, hasher = 0;
    // levels[0] = leaf digests (2^log_leaves), levels[k] = 2^(log_leaves-k) digests, last level = cap
    std::vector<dbuf<digest_t>> levels;
    const digest_t* cap() const { return levels.back().p; }
    int ncap() const { return 1 << cap_height; }
};
// leaves: column-major [ncols][nleaves] (leaf j = {col_0[j], col_1[j], ...}); interleave2: leaf = `ncols/2` consecutive
// (c0,c1) pairs taken from two columns re/im at positions [j*ncols/2, (j+1)*ncols/2)  (FRI layer leaves)
void merkle_build(DevCtx* c, MerkleTree* t, const u64* d_leaves, size_t col_stride, int log_leaves, int ncols,
                  int cap_height, int hasher, bool fri_layout = false, bool allow_clamp = false);

// ---- launch helper ----
inline void count_launch(DevCtx* c, unsigned n = 1) { c->launches += n; }

// ---- ctx.cu ----
void set_last_error(const std::string& s);
int guard(const std::function<void()>& f);  // runs f, maps exceptions to P2G_* codes + p2g_last_error
void pack_digests(int hasher, const digest_t* src, size_t n, uint8_t* dst);
