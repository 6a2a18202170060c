// DFMA issue-rate probe: NCH independent dependent-chains of DFMAs per thread, W warps per CTA, one CTA per SM.
// Prints warp-level DFMA instructions per cycle per SM sub-partition (peak of the FP64 pipe = what the
// roofline of the fused sweeps is measured against).   nvcc -arch=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__global__ void dfma_chains(double* out, int iters, double a, double b)
{
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < NCH; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;
}
template <int NCH>
static void run(int warps, int nsm, double clk_ghz)
{
    double* d; cudaMalloc(&d, 8);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dfma_chains<NCH><<<nsm, 32 * warps>>>(d, 100, 0.999, 1e-3);
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        dfma_chains<NCH><<<nsm, 32 * warps>>>(d, iters, 0.999, 1e-3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double winstr = (double)nsm * warps * iters * 8 * NCH;
    const double per_s = winstr / (best * 1e-3);
    printf("chains %d warps/SM %2d : %.4e warp-DFMA/s  = %.3f per clk per SMSP at %.3f GHz  (%.2f TFLOP/s)\n", NCH, warps,
           per_s, per_s / (nsm * 4.0 * clk_ghz * 1e9), clk_ghz, per_s * 64 / 1e12);
    cudaFree(d);
}
int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, clock %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
    for (int w : {4, 8, 16, 32}) { run<1>(w, p.multiProcessorCount, ghz); run<2>(w, p.multiProcessorCount, ghz); run<4>(w, p.multiProcessorCount, ghz); run<8>(w, p.multiProcessorCount, ghz); }
    return 0;
}
