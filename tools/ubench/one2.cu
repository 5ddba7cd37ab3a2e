#define main main_unused
#include "field_bench.cu"
#undef main
__global__ void k1(const u64* a, const u64* b, u64* o){ int i=threadIdx.x; u64 u=a[i], v=b[i]; o[i]=v1_add(u,v); }
__global__ void k2(const u64* a, const u64* b, u64* o){ int i=threadIdx.x; u64 u=a[i], v=b[i]; o[i]=v1_sub(u,v); }
__global__ void k3(const u64* a, const u64* b, u64* o){ int i=threadIdx.x; u64 u=a[i], v=b[i]; o[i]=v1_mul(u,v); }
