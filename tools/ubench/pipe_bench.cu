// Instruction-rate probe for the integer pipes of sm_100a (per SM per clock), to choose the Goldilocks instruction selection.
#include <cstdio>
#include <cstdint>
typedef uint32_t u32; typedef uint64_t u64;
__constant__ u32 c_k = 0x9E3779B9u;
template <int MODE>
__global__ void k(u32* out, int iters) {
    u32 a[8], b[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 77 + i; b[i] = blockIdx.x + i * 3; }
    u32 kk = c_k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) {        // IMAD.WIDE.U32 with 64-bit addend
                asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]) : "r"(kk));
            } else if (MODE == 1) { // IMAD (32-bit)
                asm volatile("mad.lo.u32 %0, %0, %2, %1;" : "+r"(a[i]) : "r"(b[i]), "r"(kk));
            } else if (MODE == 2) { // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[i]));
            } else if (MODE == 3) { // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %2, %1;" : "+r"(a[i]) : "r"(b[i]), "r"(kk));
            } else if (MODE == 4) { // mix: 1 IMAD + 1 IADD3 independent
                asm volatile("mad.lo.u32 %0, %0, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(kk));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(kk));
            } else if (MODE == 5) { // mix: 1 IMAD.WIDE + 2 IADD3
                asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]) : "r"(kk));
            }
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; i++) { asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(kk)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(kk)); }
        }
    }
    u32 s = 0;
    for (int i = 0; i < 8; i++) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int per_iter) {
    u32* d; cudaMalloc(&d, 148 * 8 * 512 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int iters = 4096;
    k<MODE><<<148 * 8, 512>>>(d, 16);
    cudaEventRecord(a);
    k<MODE><<<148 * 8, 512>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double inst = (double)148 * 8 * 512 * iters * 8 * per_iter;
    printf("%-28s %.3f ms  %.1f thread-instr/clk/SM (at 1.965 GHz)\n", name, ms, inst / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    run<0>("IMAD.WIDE.U32 (64b addend)", 1);
    run<1>("IMAD 32", 1);
    run<2>("IADD3+LOP3", 2);
    run<3>("IMAD.HI", 1);
    run<4>("IMAD + IADD3", 2);
    run<5>("IMAD.WIDE + 2 IADD3", 3);
    return 0;
}
