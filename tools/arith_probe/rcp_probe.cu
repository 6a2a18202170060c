// Accuracy of MUFU.RCP64H (rcp.approx.ftz.f64) + Newton iterations against IEEE division.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o rcp_probe rcp_probe.cu
#include <cstdio>
#include <cmath>
#include <cstring>
__device__ __forceinline__ double seed(double x) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r; }
__global__ void probe(int n, double lo, double hi, unsigned long long* out, double* worst)
{
    unsigned long long mis2 = 0, mis3 = 0, off2 = 0; double wseed = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // low-discrepancy sample of [lo, hi)
        double u = fmod(0.5 + i * 0.6180339887498949, 1.0);
        double x = lo + (hi - lo) * u;
        double ref = 1.0 / x;
        double r = seed(x);
        wseed = fmax(wseed, fabs(r * x - 1.0));
        double e = fma(-x, r, 1.0); r = fma(r, e, r);
        e = fma(-x, r, 1.0); r = fma(r, e, r);
        if (r != ref) { ++mis2; long long d = __double_as_longlong(r) - __double_as_longlong(ref); if (d > 1 || d < -1) ++off2; }
        e = fma(-x, r, 1.0); r = fma(r, e, r);
        if (r != ref) ++mis3;
    }
    atomicAdd(out, mis2); atomicAdd(out + 1, mis3); atomicAdd(out + 2, off2);
    atomicMax((unsigned long long*)worst, (unsigned long long)__double_as_longlong(wseed));
}
int main()
{
    unsigned long long* d; double* w; cudaMalloc(&d, 24); cudaMalloc(&w, 8);
    const double ranges[3][2] = {{0.05, 2.0}, {0.5, 1.5}, {1e-3, 1e3}};
    for (auto& rg : ranges) {
        cudaMemset(d, 0, 24); cudaMemset(w, 0, 8);
        const int n = 1 << 26;
        probe<<<1184, 256>>>(n, rg[0], rg[1], d, w);
        unsigned long long h[3]; double hw; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost); cudaMemcpy(&hw, w, 8, cudaMemcpyDeviceToHost);
        printf("x in [%g, %g): n=%d  seed max rel err %.3e (2^%.1f)  2 iterations: %llu not correctly rounded (%llu off by >1 ulp)  3 iterations: %llu\n",
               rg[0], rg[1], n, hw, log2(hw), h[0], h[2], h[1]);
    }
    return 0;
}
