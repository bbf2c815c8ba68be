// Dev probe: FP64 FMA / FP32 FMA / F2F.F64.F32 throughput per SM on this GPU (ops per clock per SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float *out, int iters, float seed) {
    float f0 = seed + threadIdx.x, f1 = seed * 2 + threadIdx.x, f2 = seed * 3, f3 = seed * 5;
    double d0 = f0, d1 = f1, d2 = f2, d3 = f3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { d0 = fma(d0, d1, d2); d1 = fma(d1, d2, d3); d2 = fma(d2, d3, d0); d3 = fma(d3, d0, d1); }
        if (MODE == 1) { f0 = fmaf(f0, f1, f2); f1 = fmaf(f1, f2, f3); f2 = fmaf(f2, f3, f0); f3 = fmaf(f3, f0, f1); }
        if (MODE == 2) { d0 += (double)f0; f0 = (float)d0 + f1; d1 += (double)f1; f1 = (float)d1 + f0; }   // 2 F2F.F64.F32 + 2 F2F.F32.F64 + 2 DADD + 2 FADD
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + (float)(d0 + d1 + d2 + d3);
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
}
int main() {
    float *out; cudaMalloc(&out, 1 << 24);
    const int iters = 20000, threads = 1024, blocks = 148;
    const char *names[3] = {"DFMA", "FFMA", "cvt mix (4 F2F + 2 DADD + 2 FADD)"};
    for (int m = 0; m < 3; ++m) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            if (m == 0) k<0><<<blocks, threads>>>(out, iters, 1.0001f);
            if (m == 1) k<1><<<blocks, threads>>>(out, iters, 1.0001f);
            if (m == 2) k<2><<<blocks, threads>>>(out, iters, 1.0001f);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        float clk; cudaMemcpy(&clk, out, 4, cudaMemcpyDeviceToHost);
        printf("%s: %.3f ms, %.0f clk per block; %.2f thread-iterations per clk per SM (4 ops per iteration for FMA modes)\n", names[m], ms, clk,
               (double)iters * threads / clk);
    }
    return 0;
}
