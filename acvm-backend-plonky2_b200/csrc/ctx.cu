// Device contexts, error slot and the stand-alone C-ABI entry points of include/p2g.h (host buffers in/out).
#include <string.h>

#include "internal.h"

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }

extern "C" const char* p2g_last_error(void) { return g_last_error.c_str(); }
extern "C" int p2g_version(void) { return P2G_VERSION; }
extern "C" int p2g_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// Pinned host memory for the witness matrix: p2g_prove's upload then runs at full PCIe rate and overlaps the inverse NTT.
extern "C" void* p2g_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        set_last_error("p2g_host_alloc: cudaHostAlloc failed");
        return nullptr;
    }
    return p;
}
extern "C" void p2g_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

static std::mutex g_ctx_mu;
static std::map<int, DevCtx*> g_ctx;

DevCtx* get_ctx(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    auto it = g_ctx.find(device);
    if (it != g_ctx.end()) {
        CUDA_CHECK(cudaSetDevice(device));
        return it->second;
    }
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) throw p2g_error(P2G_EBADARG, "no such CUDA device");
    CUDA_CHECK(cudaSetDevice(device));
    DevCtx* c = new DevCtx();
    c->device = device;
    CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    g_ctx[device] = c;
    return c;
}

DevCtx* new_child_ctx(int device) {
    DevCtx* root = get_ctx(device);
    DevCtx* c = new DevCtx();
    c->device = device;
    c->parent = root;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        throw p2g_error(P2G_ECUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    return c;
}
void free_child_ctx(DevCtx* c) {
    if (!c || !c->parent) return;
    cudaStreamSynchronize(c->stream);
    resolve_timers(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

// Stage timers never block the host: both events are recorded on the stream and parked in the context; resolve_timers() reads
// them once, after the proof's final synchronisation.  (Round 1 synchronised in the destructor, which serialised the host
// against the device whenever a caller asked for timings.)
static cudaEvent_t timer_event(DevCtx* c) {
    if (!c->ev_pool.empty()) {
        cudaEvent_t e = c->ev_pool.back();
        c->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
StageTimer::StageTimer(DevCtx* ctx, float* accum) : c(ctx), acc(accum) {
    if (!c->timing) return;
    a = timer_event(c);
    b = timer_event(c);
    cudaEventRecord(a, c->stream);
}
StageTimer::~StageTimer() {
    if (!a) return;
    cudaEventRecord(b, c->stream);
    c->ev_pending.emplace_back(a, b, acc);
}
void resolve_timers(DevCtx* c) {
    for (auto& t : c->ev_pending) {
        cudaEvent_t a = std::get<0>(t), b = std::get<1>(t);
        float ms = 0;
        if (cudaEventSynchronize(b) == cudaSuccess && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) *std::get<2>(t) += ms;
        else cudaGetLastError();
        c->ev_pool.push_back(a);
        c->ev_pool.push_back(b);
    }
    c->ev_pending.clear();
}

int guard(const std::function<void()>& f) {
    try {
        f();
        return P2G_OK;
    } catch (const p2g_error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return P2G_ENOMEM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return P2G_ECUDA;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stand-alone kernels (parity tests, micro-benchmarks)
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int p2g_ifft(const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t ncols, int device) {
    return guard([&] {
        if (!values || !coeffs || log_n > 26) throw p2g_error(P2G_EBADARG, "p2g_ifft: bad argument");
        DevCtx* c = get_ctx(device);
        size_t n = (size_t)1 << log_n, tot = n * ncols;
        dbuf<u64> in(tot), out(tot);
        CUDA_CHECK(cudaMemcpyAsync(in.p, values, tot * 8, cudaMemcpyHostToDevice, c->stream));
        ntt_ifft(c, in.p, n, out.p, n, log_n, ncols);
        CUDA_CHECK(cudaMemcpyAsync(coeffs, out.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

extern "C" int p2g_lde(const uint64_t* coeffs, uint64_t* lde, uint32_t log_n, uint32_t rate_bits, uint32_t ncols, int device) {
    return guard([&] {
        if (!coeffs || !lde || log_n + rate_bits > 28 || rate_bits > 5) throw p2g_error(P2G_EBADARG, "p2g_lde: bad argument");
        DevCtx* c = get_ctx(device);
        size_t n = (size_t)1 << log_n, m = n << rate_bits;
        dbuf<u64> in(n * ncols), out(m * ncols);
        CUDA_CHECK(cudaMemcpyAsync(in.p, coeffs, n * ncols * 8, cudaMemcpyHostToDevice, c->stream));
        ntt_lde(c, in.p, n, out.p, m, log_n, rate_bits, ncols, GL_GEN);
        CUDA_CHECK(cudaMemcpyAsync(lde, out.p, m * ncols * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

extern "C" int p2g_coset_ifft_leaforder(const uint64_t* values, uint64_t* coeffs, uint32_t log_n, uint32_t ncols, int device) {
    return guard([&] {
        if (!values || !coeffs || log_n > 28) throw p2g_error(P2G_EBADARG, "p2g_coset_ifft_leaforder: bad argument");
        DevCtx* c = get_ctx(device);
        size_t n = (size_t)1 << log_n, tot = n * ncols;
        dbuf<u64> buf(tot);
        CUDA_CHECK(cudaMemcpyAsync(buf.p, values, tot * 8, cudaMemcpyHostToDevice, c->stream));
        if (log_n == 0) {
            // single point: the value is the constant coefficient
        } else {
            ntt_coset_ifft_leaforder(c, buf.p, n, log_n, ncols, GL_GEN);
        }
        CUDA_CHECK(cudaMemcpyAsync(coeffs, buf.p, tot * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

// digests leave the library in the reference's serialised width (25 or 32 bytes each)
void pack_digests(int hasher, const digest_t* src, size_t n, uint8_t* dst) {
    int hs = hasher_bytes(hasher);
    for (size_t i = 0; i < n; i++) memcpy(dst + i * hs, &src[i], hs);
}

extern "C" int p2g_merkle_cap(const uint64_t* leaves_colmajor, uint32_t log_leaves, uint32_t ncols, uint32_t cap_height,
                              uint32_t hasher, uint8_t* cap_out, uint8_t* digests_out, int device) {
    return guard([&] {
        if (!leaves_colmajor || !cap_out || log_leaves > 28 || hasher > 1 || ncols == 0)
            throw p2g_error(P2G_EBADARG, "p2g_merkle_cap: bad argument");
        DevCtx* c = get_ctx(device);
        size_t nl = (size_t)1 << log_leaves;
        dbuf<u64> in(nl * ncols);
        CUDA_CHECK(cudaMemcpyAsync(in.p, leaves_colmajor, nl * ncols * 8, cudaMemcpyHostToDevice, c->stream));
        MerkleTree t;
        merkle_build(c, &t, in.p, nl, (int)log_leaves, (int)ncols, (int)cap_height, (int)hasher, false, true);
        std::vector<digest_t> cap(t.ncap());
        CUDA_CHECK(cudaMemcpyAsync(cap.data(), t.cap(), sizeof(digest_t) * cap.size(), cudaMemcpyDeviceToHost, c->stream));
        std::vector<digest_t> lv;
        if (digests_out) {
            lv.resize(nl);
            CUDA_CHECK(cudaMemcpyAsync(lv.data(), t.levels[0].p, sizeof(digest_t) * nl, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        pack_digests((int)hasher, cap.data(), cap.size(), cap_out);
        if (digests_out) pack_digests((int)hasher, lv.data(), nl, digests_out);
    });
}

namespace {
__global__ void k_poseidon_permute(const u64* in, u64* out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
    poseidon_permute(s);
#pragma unroll
    for (int k = 0; k < 12; k++) out[12 * i + k] = s[k];
}
__global__ void k_keccak256(const u8* msgs, size_t msg_len, size_t n, u8* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u8* m = msgs + i * msg_len;
    keccak_sponge sp;
    sp.init();
    size_t full = msg_len / 8;
    for (size_t k = 0; k < full; k++) {
        u64 wd = 0;
        for (int b = 0; b < 8; b++) wd |= (u64)m[8 * k + b] << (8 * b);
        sp.absorb_word(wd);
    }
    int tail_n = (int)(msg_len - 8 * full);
    u64 tail = 0;
    for (int b = 0; b < tail_n; b++) tail |= (u64)m[8 * full + b] << (8 * b);
    sp.finish(tail, tail_n);
    for (int k = 0; k < 4; k++)
        for (int b = 0; b < 8; b++) out[32 * i + 8 * k + b] = (u8)(sp.A[k] >> (8 * b));
}
}  // namespace

extern "C" int p2g_poseidon_permute(const uint64_t* in, uint64_t* out, size_t n, int device) {
    return guard([&] {
        if (!in || !out) throw p2g_error(P2G_EBADARG, "p2g_poseidon_permute: null pointer");
        DevCtx* c = get_ctx(device);
        if (!n) return;
        dbuf<u64> a(12 * n), b(12 * n);
        CUDA_CHECK(cudaMemcpyAsync(a.p, in, 96 * n, cudaMemcpyHostToDevice, c->stream));
        k_poseidon_permute<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(a.p, b.p, n);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(out, b.p, 96 * n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

extern "C" int p2g_keccak256(const uint8_t* msgs, size_t msg_len, size_t n, uint8_t* out, int device) {
    return guard([&] {
        if (!msgs || !out) throw p2g_error(P2G_EBADARG, "p2g_keccak256: null pointer");
        DevCtx* c = get_ctx(device);
        if (!n) return;
        dbuf<u8> a(std::max<size_t>(1, msg_len * n)), b(32 * n);
        CUDA_CHECK(cudaMemcpyAsync(a.p, msgs, msg_len * n, cudaMemcpyHostToDevice, c->stream));
        k_keccak256<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(a.p, msg_len, n, b.p);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(out, b.p, 32 * n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

// ---- host-side twins of the device primitives (the transcript runs on the host); exported for CPU-only tests ----
extern "C" void p2g_host_keccak256(const uint8_t* msg, size_t msg_len, uint8_t* out32) {
    keccak_sponge sp;
    sp.init();
    size_t full = msg_len / 8;
    for (size_t k = 0; k < full; k++) {
        u64 wd;
        memcpy(&wd, msg + 8 * k, 8);
        sp.absorb_word(wd);
    }
    u64 tail = 0;
    memcpy(&tail, msg + 8 * full, msg_len - 8 * full);
    sp.finish(tail, (int)(msg_len - 8 * full));
    memcpy(out32, sp.A, 32);
}
extern "C" void p2g_host_poseidon_permute(const uint64_t* in, uint64_t* out) {
    u64 s[12];
    memcpy(s, in, 96);
    poseidon_permute(s);
    memcpy(out, s, 96);
}
// runs a challenger: observe `n_obs` elements, then squeeze `n_out` challenges
extern "C" void p2g_host_challenger(uint32_t hasher, const uint64_t* obs, size_t n_obs, uint64_t* out, size_t n_out) {
    challenger_t ch;
    ch.init((int)hasher);
    ch.observe_many(obs, n_obs);
    for (size_t i = 0; i < n_out; i++) out[i] = ch.get();
}
extern "C" void p2g_host_two_to_one(uint32_t hasher, const uint8_t* l, const uint8_t* r, uint8_t* out) {
    digest_t a = {}, b = {};
    int hs = hasher_bytes((int)hasher);
    memcpy(&a, l, hs);
    memcpy(&b, r, hs);
    digest_t o = two_to_one((int)hasher, a, b);
    memcpy(out, &o, hs);
}
extern "C" uint64_t p2g_host_gl_mul(uint64_t a, uint64_t b) { return gl_mul(a, b); }

// ---- field-arithmetic self-test: the PTX carry-chain forms of gl.cuh against big-integer arithmetic on the host ----
// op: 0 sub(a N, b<=p)  1 add(a N, b C)  2 mul  3 glz_mul  4 mul_small(a, (u32)b)  5 canon(a)  6 glz_reduce128(hi=a, lo=b)
//     7 gl_acc dot product over groups of 8: out[i] = sum_k a[8i+k] b[8i+k]   (out has n/8 entries)
//     8 gl_acc32: the same with 32-bit multipliers (u32)b
//     9 glf_add(a C, b C)  10 glf_mul  11 glf_canon      (carry fixes on the FMA pipe)
// Results of N-class ops are canonicalised before they are returned.
GL_HD u64 field_op(int op, const u64* a, const u64* b, size_t i) {
    switch (op) {
    case 0: return gl_canon(gl_sub(a[i], b[i]));
    case 1: return gl_canon(gl_add(a[i], b[i]));
    case 2: return gl_mul(a[i], b[i]);
    case 3: return gl_canon(glz_mul(a[i], b[i]));
    case 4: return gl_mul_small(a[i], (u32)b[i]);
    case 5: return gl_canon(a[i]);
    case 6: return gl_canon(glz_reduce128(a[i], b[i]));
    case 7: {
        gl_acc acc;
        acc.clear();
        for (int k = 0; k < 8; k++) acc.mac(a[8 * i + k], b[8 * i + k]);
        return acc.reduce();
    }
    case 8: {
        gl_acc32 acc;
        acc.clear();
        for (int k = 0; k < 8; k++) acc.mac(a[8 * i + k], (u32)b[8 * i + k]);
        return acc.reduce();
    }
    case 9: return glf_add(a[i], b[i]);
    case 10: return glf_mul(a[i], b[i]);
    case 11: return glf_canon(a[i]);
    default: return 0;
    }
}
namespace {
__global__ void k_field_ops(int op, const u64* a, const u64* b, u64* out, size_t nout) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nout) out[i] = field_op(op, a, b, i);
}
}  // namespace
extern "C" int p2g_test_field_ops(int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, int device) {
    return guard([&] {
        if (!a || !b || !out || op < 0 || op > 11) throw p2g_error(P2G_EBADARG, "p2g_test_field_ops: bad argument");
        size_t nout = (op == 7 || op == 8) ? n / 8 : n;
        if (!nout) return;
        if (device < 0) {  // host twin
            for (size_t i = 0; i < nout; i++) out[i] = field_op(op, a, b, i);
            return;
        }
        DevCtx* c = get_ctx(device);
        dbuf<u64> da(n), db(n), dout(nout);
        CUDA_CHECK(cudaMemcpyAsync(da.p, a, n * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(db.p, b, n * 8, cudaMemcpyHostToDevice, c->stream));
        k_field_ops<<<(unsigned)((nout + 127) / 128), 128, 0, c->stream>>>(op, da.p, db.p, dout.p, nout);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(out, dout.p, nout * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
