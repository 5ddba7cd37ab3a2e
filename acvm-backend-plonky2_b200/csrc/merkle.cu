// Merkle-cap hashing of LDE columns: KeccakHash<25> (the CLI's hasher) and PoseidonHash, sm_100a.
//
// Replaces plonky2 0.2.2 hash/merkle_tree.rs MerkleTree::new / fill_digests_buf (+ the transpose that precedes it in
// fri/oracle.rs), reached from the reference at plonky2-backend/src/actions/prove_action.rs:96.
//
// The leaf kernel reads the LDE straight from its column-major home: thread j absorbs column c at [c][j], so every warp
// load is one fully-coalesced 256-byte run and the "row" the reference materialises by transposing never exists.  The
// sponge state (25 lanes Keccak / 12 lanes Poseidon) lives in registers for the whole leaf; the S-box + 12-wide MDS of
// Poseidon are fused per thread.  Digests are kept in 32-byte slots for every level (they are the Merkle paths of the
// query phase).
#include "internal.h"

namespace {

template <bool FRI>
__device__ __forceinline__ u64 leaf_elem(const u64* in, size_t cs, int ncols, size_t j, int k) {
    if (FRI) {
        int arity = ncols >> 1;
        return in[(size_t)(k & 1) * cs + j * arity + (k >> 1)];
    }
    return in[(size_t)k * cs + j];
}

template <bool FRI>
__global__ void __launch_bounds__(128) k_leaf_keccak(const u64* __restrict__ in, size_t cs, int ncols, size_t nleaves,
                                                     digest_t* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nleaves) return;
    digest_t o;
    if (ncols * 8 <= 25) {  // hash_or_noop: short inputs are copied, zero padded
        for (int i = 0; i < 4; i++) o.w[i] = i < ncols ? leaf_elem<FRI>(in, cs, ncols, j, i) : 0;
        out[j] = o;
        return;
    }
    u64 A[25];
#pragma unroll
    for (int i = 0; i < 25; i++) A[i] = 0;
    int base = 0;
    for (; base + 17 <= ncols; base += 17) {
        u64 v[17];
#pragma unroll
        for (int i = 0; i < 17; i++) v[i] = leaf_elem<FRI>(in, cs, ncols, j, base + i);
#pragma unroll
        for (int i = 0; i < 17; i++) A[i] ^= v[i];
        keccak_f1600(A);
    }
    int rem = ncols - base;
#pragma unroll
    for (int i = 0; i < 17; i++) {
        if (i < rem) A[i] ^= leaf_elem<FRI>(in, cs, ncols, j, base + i);
        if (i == rem) A[i] ^= 0x01;
    }
    A[16] ^= 0x8000000000000000ULL;
    keccak_f1600(A);
    o.w[0] = A[0];
    o.w[1] = A[1];
    o.w[2] = A[2];
    o.w[3] = A[3] & 0xff;
    out[j] = o;
}

template <bool FRI>
__global__ void __launch_bounds__(128) k_leaf_poseidon(const u64* __restrict__ in, size_t cs, int ncols, size_t nleaves,
                                                       digest_t* __restrict__ out) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nleaves) return;
    digest_t o;
    if (ncols <= 4) {
        for (int i = 0; i < 4; i++) o.w[i] = i < ncols ? leaf_elem<FRI>(in, cs, ncols, j, i) : 0;
        out[j] = o;
        return;
    }
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    // one instance of the permutation: the last (partial) chunk overwrites only the lanes it has (overwrite-mode sponge)
#pragma unroll 1
    for (int base = 0; base < ncols; base += 8) {
        const int have = ncols - base;
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < have) s[i] = leaf_elem<FRI>(in, cs, ncols, j, base + i);
        poseidon_permute(s);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) o.w[i] = s[i];
    out[j] = o;
}

__global__ void __launch_bounds__(128) k_nodes(int h, const digest_t* __restrict__ in, digest_t* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    digest_t l = in[2 * i], r = in[2 * i + 1];
    out[i] = h == P2G_H_KECCAK25 ? keccak25_two_to_one(l, r) : poseidon_two_to_one(l, r);
}

// The last levels of a tree (<= 1024 nodes) in ONE launch: a level per __syncthreads instead of a launch per level (each of
// those launches ran below one wave and cost its launch latency: ~10 per tree, 7 trees per proof).
#define TAIL_MAX 12
struct TailArgs {
    const digest_t* in;          // the level below the first tail level
    digest_t* out[TAIL_MAX];     // the tail levels, bottom up
    int nlev;
    unsigned cnt0;               // nodes of the first tail level
};
__global__ void __launch_bounds__(1024) k_nodes_tail(int h, TailArgs a) {
    const unsigned t = threadIdx.x;
    const digest_t* in = a.in;
    unsigned cnt = a.cnt0;
    for (int k = 0; k < a.nlev; k++, cnt >>= 1) {
        if (t < cnt) {
            digest_t l = in[2 * t], r = in[2 * t + 1];
            a.out[k][t] = h == P2G_H_KECCAK25 ? keccak25_two_to_one(l, r) : poseidon_two_to_one(l, r);
        }
        __syncthreads();   // block-wide: the level just written is read by the same block
        in = a.out[k];
    }
}

}  // namespace

void merkle_build(DevCtx* c, MerkleTree* t, const u64* d_leaves, size_t col_stride, int log_leaves, int ncols, int cap_height,
                  int hasher, bool fri_layout, bool allow_clamp) {
    StageTimer tm(c, &c->merkle_ms);
    if (cap_height > log_leaves) {
        // plonky2's MerkleTree::new asserts here; only the stand-alone test entry point (p2g_merkle_cap) may clamp
        if (!allow_clamp) throw p2g_error(P2G_EBADARG, "merkle_build: fewer than 2^cap_height leaves");
        cap_height = log_leaves;
    }
    t->log_leaves = log_leaves;
    t->cap_height = cap_height;
    t->hasher = hasher;
    int nlevels = log_leaves - cap_height + 1;
    if ((int)t->levels.size() != nlevels) {
        t->levels.clear();
        t->levels.resize(nlevels);
    }
    size_t nl = (size_t)1 << log_leaves;
    for (int k = 0; k < nlevels; k++)
        if (t->levels[k].n != (nl >> k)) t->levels[k].alloc(nl >> k);
    const int TH = 128;
    unsigned grid = (unsigned)((nl + TH - 1) / TH);
    {
    StageTimer tl(c, &c->leaf_ms);
    if (hasher == P2G_H_KECCAK25) {
        if (fri_layout) k_leaf_keccak<true><<<grid, TH, 0, c->stream>>>(d_leaves, col_stride, ncols, nl, t->levels[0].p);
        else k_leaf_keccak<false><<<grid, TH, 0, c->stream>>>(d_leaves, col_stride, ncols, nl, t->levels[0].p);
    } else {
        if (fri_layout) k_leaf_poseidon<true><<<grid, TH, 0, c->stream>>>(d_leaves, col_stride, ncols, nl, t->levels[0].p);
        else k_leaf_poseidon<false><<<grid, TH, 0, c->stream>>>(d_leaves, col_stride, ncols, nl, t->levels[0].p);
    }
    }
    c->leaf_launches++;
    c->leaf_bytes += 8.0 * (double)nl * ncols;
    count_launch(c);
    for (int k = 1; k < nlevels; k++) {
        size_t cnt = nl >> k;
        if (cnt <= 1024 && nlevels - k <= TAIL_MAX) {   // the rest of the tree in one block
            TailArgs a = {};
            a.in = t->levels[k - 1].p;
            a.nlev = nlevels - k;
            a.cnt0 = (unsigned)cnt;
            for (int q = 0; q < a.nlev; q++) a.out[q] = t->levels[k + q].p;
            k_nodes_tail<<<1, (unsigned)std::max<size_t>(32, cnt), 0, c->stream>>>(hasher, a);
            count_launch(c);
            break;
        }
        k_nodes<<<(unsigned)((cnt + TH - 1) / TH), TH, 0, c->stream>>>(hasher, t->levels[k - 1].p, t->levels[k].p, cnt);
        count_launch(c);
    }
    CUDA_CHECK(cudaGetLastError());
    c->merkle_bytes += 8.0 * (double)nl * ncols;
}
