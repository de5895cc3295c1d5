// FP64 pipe microbenchmark: independent DFMA / DMUL+DADD chains (scratch)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) x[j] = fma(x[j], a, b);
            else if (MODE == 1) x[j] = __dadd_rn(__dmul_rn(x[j], a), b);
            else x[j] = __dadd_rn(x[j], b);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int warps)
{
    double *out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148, threads = warps * 32;
    k<MODE><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * 8;
    printf("%s warps/SM=%d: %.3f ms, %.2f Tinstr-lanes/s (%.1f lanes/clk/SM at 1.9 GHz)\n", name, warps, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(out);
}
int main()
{
    for (int w : {4, 8, 16, 32}) { run<0>("DFMA", w); run<1>("DMUL+DADD", w); run<2>("DADD", w); }
    return 0;
}
