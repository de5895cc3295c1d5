// dependent-issue latencies on one warp (scratch)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *clk, double a, double b)
{
    double x = threadIdx.x * 1e-3 + 1.0;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); x = __dadd_rn(x, b); }
    long long t1 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { x = __dmul_rn(x, a); x = __dmul_rn(x, a); x = __dmul_rn(x, a); x = __dmul_rn(x, a); }
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); x = __shfl_up_sync(0xffffffffu, x, 1); }
    long long t3 = clock64();
    __shared__ double sm[64];
    sm[threadIdx.x] = x;
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) {
        // store -> neighbour load chain, as in the wavefront
        for (int u = 0; u < 4; ++u) { double v = sm[(threadIdx.x + 31) & 31]; __syncwarp(); sm[threadIdx.x] = v + b; __syncwarp(); }
    }
    long long t4 = clock64();
    x = sm[threadIdx.x];
#pragma unroll 1
    for (int i = 0; i < 1000; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
    long long t5 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; clk[4] = t5 - t4; }
}
int main()
{
    double *out; long long *clk, h[5];
    cudaMalloc(&out, 256); cudaMalloc(&clk, 64);
    k<<<1, 32>>>(out, clk, 1.0000001, 1e-9);
    k<<<1, 32>>>(out, clk, 1.0000001, 1e-9);
    cudaMemcpy(h, clk, 40, cudaMemcpyDeviceToHost);
    printf("DADD %.1f  DMUL %.1f  SHFL(64-bit) %.1f  LDS->DADD->STS(+2 syncwarp) %.1f  DFMA %.1f cycles per dependent op\n",
           h[0] / 4000.0, h[1] / 4000.0, h[2] / 4000.0, h[3] / 4000.0, h[4] / 4000.0);
    return 0;
}
