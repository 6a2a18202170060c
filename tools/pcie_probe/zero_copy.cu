// Can a kernel move the state over the host link itself?  Stores from a kernel straight into pinned host memory
// (the dense [4][nz+4][nx+4] layout of the reference's array: rows are NOT 128-byte aligned) and loads from it,
// alone and while a DMA copy runs in the other direction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o zero_copy zero_copy.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// one block per row: copies `w` doubles of row r (src pitch ps, dst pitch pd) with 16-byte accesses
__global__ void copy_rows(const double* __restrict__ src, size_t ps, double* __restrict__ dst, size_t pd, int w, int rows)
{
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const double2* s = reinterpret_cast<const double2*>(src + (size_t)r * ps);
        double2* d = reinterpret_cast<double2*>(dst + (size_t)r * pd);
        for (int i = threadIdx.x; i < w / 2; i += blockDim.x) d[i] = s[i];
    }
}

int main(int argc, char** argv)
{
    const int nx = argc > 1 ? atoi(argv[1]) : 2048, nz = argc > 2 ? atoi(argv[2]) : 1024;
    const int w = nx + 4, rows = 4 * (nz + 4), pitch = (14 + nx + 8 + 15) / 16 * 16;
    const size_t nb = (size_t)w * rows * 8;
    double *h1, *h2, *d1, *d2;
    CK(cudaHostAlloc(&h1, nb, cudaHostAllocMapped)); CK(cudaHostAlloc(&h2, nb, cudaHostAllocMapped));
    CK(cudaMalloc(&d1, (size_t)pitch * rows * 8)); CK(cudaMalloc(&d2, (size_t)pitch * rows * 8));
    CK(cudaMemset(d1, 0, (size_t)pitch * rows * 8)); CK(cudaMemset(d2, 0, (size_t)pitch * rows * 8));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto timeit = [&](auto fn, const char* what, double bytes) {
        for (int i = 0; i < 3; ++i) fn();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, 0));
        const int reps = 10;
        for (int i = 0; i < reps; ++i) { fn(); CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2)); }
        CK(cudaEventRecord(e1, 0)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
        printf("%-62s %7.3f ms  %6.1f GB/s\n", what, ms, bytes / ms / 1e6);
    };
    for (int grid : {592}) {
        printf("grid %d x 256 threads\n", grid);
        timeit([&] { copy_rows<<<grid, 256, 0, s1>>>(d1 + 14, pitch, h1, w, w, rows); }, "kernel stores to host", nb);
        timeit([&] { copy_rows<<<grid, 256, 0, s1>>>(h1, w, d1 + 14, pitch, w, rows); }, "kernel loads from host", nb);
        timeit([&] { copy_rows<<<grid, 256, 0, s1>>>(d1 + 14, pitch, h1, w, w, rows);
                     CK(cudaMemcpy2DAsync(d2 + 14, pitch * 8, h2, w * 8, w * 8, rows, cudaMemcpyHostToDevice, s2)); },
               "kernel stores to host + DMA upload (both directions)", 2.0 * nb);
        timeit([&] { copy_rows<<<grid, 256, 0, s1>>>(h1, w, d1 + 14, pitch, w, rows);
                     CK(cudaMemcpy2DAsync(h2, w * 8, d2 + 14, pitch * 8, w * 8, rows, cudaMemcpyDeviceToHost, s2)); },
               "kernel loads from host + DMA download (both directions)", 2.0 * nb);
        timeit([&] { copy_rows<<<grid, 256, 0, s1>>>(h1, w, d1 + 14, pitch, w, rows);
                     copy_rows<<<grid, 256, 0, s2>>>(d2 + 14, pitch, h2, w, w, rows); },
               "kernel loads + kernel stores (both directions)", 2.0 * nb);
    }
    timeit([&] { CK(cudaMemcpy2DAsync(d2 + 14, pitch * 8, h2, w * 8, w * 8, rows, cudaMemcpyHostToDevice, s2));
                 CK(cudaMemcpy2DAsync(h1, w * 8, d1 + 14, pitch * 8, w * 8, rows, cudaMemcpyDeviceToHost, s1)); },
           "DMA upload + DMA download (both directions)", 2.0 * nb);
    // the same two transfers cut into bands of rows, one 3-D copy per band and direction (what pmw_evolve_host issues),
    // the two directions independent of each other: the cost of band-sized DMA jobs by itself
    for (int nbands : {4, 8, 16, 32, 64}) {
        char what[96];
        snprintf(what, sizeof what, "DMA upload + DMA download in %d bands (3-D copies)", nbands);
        timeit([&] {
            for (int b = 0; b < nbands; ++b) {
                const int r0 = (nz + 4) * b / nbands, r1 = (nz + 4) * (b + 1) / nbands;
                cudaMemcpy3DParms q = {};
                const cudaPitchedPtr hu = make_cudaPitchedPtr(h2, w * 8, w * 8, nz + 4), du = make_cudaPitchedPtr(d2 + 14, pitch * 8, w * 8, nz + 4);
                q.srcPtr = hu; q.dstPtr = du; q.srcPos = q.dstPos = make_cudaPos(0, r0, 0);
                q.extent = make_cudaExtent(w * 8, r1 - r0, 4); q.kind = cudaMemcpyHostToDevice;
                CK(cudaMemcpy3DAsync(&q, s2));
                const cudaPitchedPtr hd = make_cudaPitchedPtr(h1, w * 8, w * 8, nz + 4), dd = make_cudaPitchedPtr(d1 + 14, pitch * 8, w * 8, nz + 4);
                q.srcPtr = dd; q.dstPtr = hd; q.kind = cudaMemcpyDeviceToHost;
                CK(cudaMemcpy3DAsync(&q, s1));
            }
        }, what, 2.0 * nb);
    }
    return 0;
}
