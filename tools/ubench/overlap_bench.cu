// Do the integer ALU pipe (IADD3/LOP3) and the FMA pipe (IMAD.WIDE.U32) overlap on sm_100a?  Independent dependency chains of each
// kind in one thread, ratio R ALU instructions per IMAD.WIDE.  Prints thread-instructions per clock per SM for each mix.
#include <cstdio>
#include <cstdint>
typedef uint32_t u32; typedef uint64_t u64;
__constant__ u32 c_k = 0x9E3779B9u;
template <int NW, int NA>   // NW wide chains, NA alu chains per thread; every chain advances once per iteration
__global__ void k(u32* out, int iters) {
    u32 wl[NW > 0 ? NW : 1], wh[NW > 0 ? NW : 1], al[NA > 0 ? NA : 1];
    for (int i = 0; i < NW; i++) { wl[i] = threadIdx.x * 77 + i; wh[i] = blockIdx.x + i * 3; }
    for (int i = 0; i < NA; i++) al[i] = threadIdx.x * 13 + i;
    u32 kk = c_k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NW; i++)
            asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %0, %2, t; mov.b64 {%0, %1}, t; }" : "+r"(wl[i]), "+r"(wh[i]) : "r"(kk));
#pragma unroll
        for (int i = 0; i < NA; i++)
            asm volatile("add.u32 %0, %0, %1;" : "+r"(al[i]) : "r"(kk));
    }
    u32 s = 0;
    for (int i = 0; i < NW; i++) s += wl[i] ^ wh[i];
    for (int i = 0; i < NA; i++) s += al[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NW, int NA> void run() {
    u32* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int iters = 4096;
    k<NW, NA><<<148 * 4, 512>>>(d, 16);
    cudaEventRecord(a);
    k<NW, NA><<<148 * 4, 512>>>(d, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double per = (double)148 * 4 * 512 * iters / (ms * 1e-3) / 148 / 1.965e9;   // iterations per clk per SM (thread-level)
    printf("wide chains %d  alu chains %2d : %.3f ms   IMAD.WIDE %.1f /clk/SM   IADD %.1f /clk/SM   total %.1f\n", NW, NA, ms, per * NW, per * NA,
           per * (NW + NA));
    cudaFree(d);
}
int main() {
    run<8, 0>(); run<0, 8>(); run<0, 16>(); run<4, 4>(); run<4, 8>(); run<4, 12>(); run<4, 16>(); run<4, 24>(); run<2, 16>();
    return 0;
}
