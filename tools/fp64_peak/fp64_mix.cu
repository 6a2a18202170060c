// Does a DFMA cost one issue slot or two?  Per loop iteration: 32 DFMAs (4 independent chains) plus K integer
// multiply-adds per DFMA (independent chains, other pipe).  If the FP64 pipe only needs the issue port every other
// cycle, the DFMA rate stays ~0.45/clk/SMSP up to K = 1; if a DFMA blocks the port for two cycles, the rate is
// 1/(2+K).   nvcc -arch=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__global__ void mix(double* out, int iters, double a, double b, int m)
{
    double x[4];
    int y[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = threadIdx.x * 1e-3 + c;
#pragma unroll
    for (int c = 0; c < 8; ++c) y[c] = threadIdx.x + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                x[c] = fma(x[c], a, b);
#pragma unroll
                for (int k = 0; k < K; ++k) y[(c * K + k) & 7] = y[(c * K + k) & 7] * m + u;
            }
        }
    }
    double s = 0; int t = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) s += x[c];
#pragma unroll
    for (int c = 0; c < 8; ++c) t += y[c];
    if (s == 12345.678 || t == 0x7fffffff) out[0] = s + t;
}
template <int K>
static void run(int warps, int nsm, double ghz)
{
    double* d; cudaMalloc(&d, 8);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<K><<<nsm, 32 * warps>>>(d, 100, 0.999, 1e-3, 3);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); mix<K><<<nsm, 32 * warps>>>(d, iters, 0.999, 1e-3, 3); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double dfma = (double)nsm * warps * iters * 32 / (best * 1e-3) / (nsm * 4.0 * ghz * 1e9);
    printf("K=%d int ops per DFMA, %2d warps/SM: %.3f DFMA/clk/SMSP, %.3f instr/clk/SMSP\n", K, warps, dfma, dfma * (1 + K));
    cudaFree(d);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    for (int w : {8, 16}) { run<0>(w, p.multiProcessorCount, ghz); run<1>(w, p.multiProcessorCount, ghz); run<2>(w, p.multiProcessorCount, ghz); run<3>(w, p.multiProcessorCount, ghz); }
    return 0;
}
