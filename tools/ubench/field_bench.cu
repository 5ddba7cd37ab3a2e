// Micro-benchmark: Goldilocks butterfly throughput for alternative instruction selections (ALU pipe vs FMA pipe balance).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o field_bench field_bench.cu && ./field_bench
#include <cstdio>
#include <cstdint>
#include "../../acvm-backend-plonky2_b200/csrc/gl.cuh"

// ---- variant 1: carry fixes moved to the FMA pipe.  eps / one come from constant memory so ptxas keeps the IMAD forms ----
__constant__ u32 c_eps = 0xFFFFFFFFu;
__constant__ u32 c_one = 1u;
__device__ __forceinline__ u64 v1_fix_add_eps(u64 x, u32 m /*0 or 1*/) {   // x + m * eps (no overflow by construction)
    u64 r;
    asm("{\n\t.reg .u32 x0, x1;\n\t"
        "mov.b64 {x0, x1}, %3;\n\t"
        "mad.lo.cc.u32 x0, %1, %2, x0;\n\t"
        "madc.hi.u32 x1, %1, %2, x1;\n\t"
        "mov.b64 %0, {x0, x1};\n\t}" : "=l"(r) : "r"(m), "r"(c_eps), "l"(x));
    return r;
}
__device__ __forceinline__ u64 v1_fix_sub_eps(u64 d, u32 br /*0 or 0xffffffff*/) {   // d - (br ? eps : 0)
    return d - (u64)br;
}
__device__ __forceinline__ u64 v1_canon(u64 x) {
    u32 m;
    asm("{\n\t.reg .u64 y;\n\tadd.cc.u64 y, %1, 0xffffffff;\n\taddc.u32 %0, 0, 0;\n\t}" : "=r"(m) : "l"(x));
    return v1_fix_add_eps(x, m);
}
__device__ __forceinline__ u64 v1_reduce(u64 lo, u32 r2, u32 r3) {
    u64 t, x;
    u32 b, m;
    asm("{\n\tsub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;\n\t}" : "=l"(t), "=r"(b) : "l"(lo), "l"((u64)r3));   // b = borrow ? -1 : 0
    t = v1_fix_sub_eps(t, b);
    asm("{\n\t.reg .u32 x0, x1, t0, t1;\n\t"
        "mov.b64 {t0, t1}, %2;\n\t"
        "mad.lo.cc.u32 x0, %3, %4, t0;\n\t"
        "madc.hi.cc.u32 x1, %3, %4, t1;\n\t"
        "addc.u32 %1, 0, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\t}" : "=l"(x), "=r"(m) : "l"(t), "r"(r2), "r"(c_eps));
    return v1_fix_add_eps(x, m);
}
__device__ __forceinline__ u64 v1_mulz(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %7;\n\t"
        "mul.hi.u32 %3, %5, %7;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return v1_reduce(((u64)r1 << 32) | r0, r2, r3);
}
__device__ __forceinline__ u64 v1_mul(u64 a, u64 b) { return v1_canon(v1_mulz(a, b)); }
// a, b: C -> C
__device__ __forceinline__ u64 v1_add(u64 a, u64 b) {
    u64 s; u32 c;
    asm("{\n\t.reg .u64 y;\n\t"
        "add.cc.u64 %0, %2, %3;\n\t"
        "addc.u32 %1, 0, 0;\n\t"          // carry of a + b
        "add.cc.u64 y, %0, 0xffffffff;\n\t"
        "addc.u32 %1, %1, 0;\n\t}"        // + carry of (a + b mod 2^64) + eps : at most one of the two is set
        : "=&l"(s), "=&r"(c) : "l"(a), "l"(b));
    return v1_fix_add_eps(s, c);
}
__device__ __forceinline__ u64 v1_sub(u64 a, u64 b) {
    u64 d; u32 br;
    asm("{\n\tsub.cc.u64 %0, %2, %3;\n\tsubc.u32 %1, 0, 0;\n\t}" : "=l"(d), "=r"(br) : "l"(a), "l"(b));   // br = borrow ? -1 : 0
    return v1_fix_sub_eps(d, br);
}


// ---- variant 2: carry fixes as compare + predicated/selected adds (no IMAD.WIDE fixes) ----
__device__ __forceinline__ u64 v2_canon(u64 x) { u64 y = x + 0xFFFFFFFFULL; return y < x ? y : x; }
__device__ __forceinline__ u64 v2_add(u64 a, u64 b) {   // C, C -> C
    u64 s = a + b;
    u64 y = s + 0xFFFFFFFFULL;
    return (s < a || y < s) ? y : s;
}
__device__ __forceinline__ u64 v2_mul(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1, r2, r3;
    asm("{\n\t"
        "mul.lo.u32 %0, %4, %6;\n\t"
        "mul.hi.u32 %1, %4, %6;\n\t"
        "mul.lo.u32 %2, %5, %7;\n\t"
        "mul.hi.u32 %3, %5, %7;\n\t"
        "mad.lo.cc.u32 %1, %4, %7, %1;\n\t"
        "madc.hi.cc.u32 %2, %4, %7, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 %1, %5, %6, %1;\n\t"
        "madc.hi.cc.u32 %2, %5, %6, %2;\n\t"
        "addc.u32 %3, %3, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    u64 lo = ((u64)r1 << 32) | r0;
    u64 t = gl_sub(lo, (u64)r3);
    u64 u = ((u64)r2 << 32) - r2;
    u64 x = t + u;
    if (x < u) x += 0xFFFFFFFFULL;
    return v2_canon(x);
}
template <int V>
__global__ void k_bfly(u64* data, int iters, u64 w0) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u64 x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = data[i * 8 + k];
    u64 w = w0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int h = 4; h >= 1; h >>= 1) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k & h) continue;
                u64 u = x[k], v = x[k + h];
                if (V == 0) {
                    x[k] = gl_add(u, v);
                    x[k + h] = gl_mul(gl_sub(u, v), w);
                } else if (V == 1) {
                    x[k] = v1_add(u, v);
                    x[k + h] = v1_mul(v1_sub(u, v), w);
                } else if (V == 2) {
                    x[k] = v2_add(u, v);
                    x[k + h] = v2_mul(gl_sub(u, v), w);
                } else {
                    x[k] = glf_add(u, v);
                    x[k + h] = glf_mul(gl_sub(u, v), w);
                }
            }
        }
        w = x[3] | 1;   // data-dependent twiddle (keeps the compiler honest); may be non-canonical: mul accepts N
    }
#pragma unroll
    for (int k = 0; k < 8; k++) data[i * 8 + k] = x[k];
}

int main() {
    const size_t nthr = 148 * 2048 * 4;
    u64* d;
    cudaMalloc(&d, nthr * 8 * 8);
    u64* h = (u64*)malloc(nthr * 64);
    u64* h0 = (u64*)malloc(nthr * 64);
    u64* h1 = (u64*)malloc(nthr * 64);
    u64 s = 88172645463325252ULL;
    for (size_t i = 0; i < nthr * 8; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s % GL_P; }
    // adversarial lanes
    h[0] = GL_P - 1; h[1] = GL_P - 1; h[2] = 0; h[3] = GL_P - 1; h[4] = 1; h[5] = 0xFFFFFFFFULL; h[6] = 0xFFFFFFFF00000000ULL; h[7] = 0x100000000ULL;
    const int iters = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int v = 0; v < 4; v++) {
        float best = 1e9;
        for (int rep = 0; rep < 4; rep++) {
            cudaMemcpy(d, h, nthr * 64, cudaMemcpyHostToDevice);
            cudaEventRecord(a);
            if (v == 0) k_bfly<0><<<nthr / 256, 256>>>(d, iters, 12345678901234567ULL);
            else if (v == 1) k_bfly<1><<<nthr / 256, 256>>>(d, iters, 12345678901234567ULL);
            else if (v == 2) k_bfly<2><<<nthr / 256, 256>>>(d, iters, 12345678901234567ULL);
            else k_bfly<3><<<nthr / 256, 256>>>(d, iters, 12345678901234567ULL);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
        }
        cudaMemcpy(v ? h1 : h0, d, nthr * 64, cudaMemcpyDeviceToHost);
        if (v) { size_t bad = 0; for (size_t i = 0; i < nthr * 8; i++) bad += h0[i] != h1[i]; printf("  variant %d vs 0 mismatches: %zu\n", v, bad); }
        double bf = (double)nthr * iters * 12;
        printf("variant %d: %.3f ms  %.1f G butterflies/s  (err %s)\n", v, best, bf / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    size_t bad = 0;
    for (size_t i = 0; i < nthr * 8; i++) if (h0[i] != h1[i]) { if (!bad) printf("mismatch at %zu: %llx vs %llx\n", i, (unsigned long long)h0[i], (unsigned long long)h1[i]); bad++; }
    printf("mismatches: %zu\n", bad);
    return bad != 0;
}
