// Variant 0 ("direct"): one thread per cell, stencil operands read straight from global
// memory through L1/L2, both interface fluxes of the cell recomputed by the owning thread.
// No shared memory, no inter-thread exchange: it is the simplest correct GPU formulation of a
// stage and serves as the on-device cross-check of the TMA variant (and as the fallback for
// grids too small for a tile).  Same arithmetic (interface_flux) as the production kernels.
#pragma once
#include "pmw_common.cuh"

namespace pmw {

// Bounded spin on a neighbour's epoch flag (see wait_epoch in pmw_tma.cuh; no async-proxy fence
// needed here, the direct kernels read with ordinary loads -- volatile, not the read-only path,
// for the halo columns would be wrong; the taps below use ld.global.nc only when not waiting).
__device__ __forceinline__ void spin_epoch(unsigned long long* flags, int which, unsigned long long epoch)
{
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + which) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 4000000000LL) {
            flags[2] = 1ull;
            break;
        }
    } while (true);
}

// Stores one updated cell and, when asked, its periodic / slab-neighbour halo image
// (set_bc_x, bcs.py:35-39, folded into the producer of the next x stage's forcing state).
__device__ __forceinline__ void store_cell(const StageArgs& a, int v, int k, int i, double val)
{
    const long long o = idx(a.L, v, k + HS, i + HS);
    a.out[o] = val;
    if (a.write_xhalo) {
        if (i < HS) a.out_left[o + a.L.nx] = val;             // left neighbour's right halo
        if (i >= a.L.nx - HS) a.out_right[o - a.L.nx] = val;  // right neighbour's left halo
    }
}

__device__ __forceinline__ IfaceBg bg_x(const Hydro& hy, int krow /* array row */)
{
    IfaceBg bg;
    bg.dens = __ldg(hy.dens_cell + krow);
    bg.dens_theta = __ldg(hy.dens_theta_cell + krow);
    bg.inv_dens_theta = __ldg(hy.inv_dens_theta_cell + krow);
    bg.pressure = __ldg(hy.pressure_cell + krow);
    return bg;
}

__device__ __forceinline__ IfaceBg bg_z(const Hydro& hy, int k /* interface */)
{
    IfaceBg bg;
    bg.dens = __ldg(hy.dens_int + k);
    bg.dens_theta = __ldg(hy.dens_theta_int + k);
    bg.inv_dens_theta = __ldg(hy.inv_dens_theta_int + k);
    bg.pressure = __ldg(hy.pressure_int + k);
    return bg;
}

template <int POW_MODE>
__device__ __forceinline__ void stage_x_direct_cell(const StageArgs& a, int i, int k);
template <int POW_MODE>
__device__ __forceinline__ void stage_z_direct_cell(const StageArgs& a, int i, int k);

// x stage: interpolate_x + compute_flux_x + compute_tend_x + update (step.py:69-71,80-82).
template <int POW_MODE>
__global__ void __launch_bounds__(256) stage_x_direct(const StageArgs a)
{
    const int push_rows = a.push_epoch ? 1 : 0;  // slab ring: first row of blocks pushes the edge columns
    if (push_rows && blockIdx.y == 0) {
        push_halo_role(a);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                // interior column
    const int k = (blockIdx.y - push_rows) * blockDim.y + threadIdx.y;  // interior row
    if (i < a.L.nx && k < a.L.nz) stage_x_direct_cell<POW_MODE>(a, i, k);
}

template <int POW_MODE>
__device__ __forceinline__ void stage_x_direct_cell(const StageArgs& a, int i, int k)
{
    if (a.wait_epoch) {  // slab ring over peer memory: halo columns come from the neighbours
        if (i < HS) spin_epoch(a.flags, 0, a.wait_epoch);
        if (i >= a.L.nx - HS) spin_epoch(a.flags, 1, a.wait_epoch);
    }
    double s[5][4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const double* f = a.forcing + idx(a.L, v, k + HS, i);  // array column i = cell i-2
#pragma unroll
        for (int j = 0; j < 5; ++j) s[j][v] = a.wait_epoch ? f[j] : __ldg(f + j);
    }
    const IfaceBg bg = bg_x(a.hy, k + HS);
    double fl[4], fr[4];
    interface_flux<false, POW_MODE>(s[0], s[1], s[2], s[3], bg, a.hv_coeff, false, fl);
    interface_flux<false, POW_MODE>(s[1], s[2], s[3], s[4], bg, a.hv_coeff, false, fr);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const double ini = (a.init == a.forcing) ? s[2][v] : __ldg(a.init + idx(a.L, v, k + HS, i + HS));
        double x;
        if (v == WMOM && a.src_w)
            x = cell_update<false, true>(fl[v], fr[v], ini, a.cd, a.cg, 0.0, a.dt_stage,
                                         __ldg(a.src_w + (long long)k * a.L.nx + i));
        else
            x = cell_update<false, false>(fl[v], fr[v], ini, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
        store_cell(a, v, k, i, x);
    }
}

// z stage: interpolate_z + compute_flux_z + compute_tend_z + update (step.py:74-76,80-82).
template <int POW_MODE>
__global__ void __launch_bounds__(256) stage_z_direct(const StageArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < a.L.nx && k < a.L.nz) stage_z_direct_cell<POW_MODE>(a, i, k);
}

template <int POW_MODE>
__device__ __forceinline__ void stage_z_direct_cell(const StageArgs& a, int i, int k)
{
    const int nz = a.L.nz;
    double s[5][4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int kk = k + j;  // array row = cell row k-2+j
            double val;
            if (a.fuse_bc_z && kk < HS) {
                val = wall_value(v, __ldg(a.forcing + idx(a.L, v, HS, i + HS)),
                                 __ldg(a.hy.dens_cell + HS), __ldg(a.hy.dens_cell + kk));
            } else if (a.fuse_bc_z && kk >= nz + HS) {
                val = wall_value(v, __ldg(a.forcing + idx(a.L, v, nz + HS - 1, i + HS)),
                                 __ldg(a.hy.dens_cell + nz + HS - 1), __ldg(a.hy.dens_cell + kk));
            } else {
                val = __ldg(a.forcing + idx(a.L, v, kk, i + HS));
            }
            s[j][v] = val;
        }
    }
    double fb[4], ft[4];
    interface_flux<true, POW_MODE>(s[0], s[1], s[2], s[3], bg_z(a.hy, k), a.hv_coeff, k == 0, fb);
    interface_flux<true, POW_MODE>(s[1], s[2], s[3], s[4], bg_z(a.hy, k + 1), a.hv_coeff,
                                   k + 1 == nz, ft);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const double ini = (a.init == a.forcing) ? s[2][v] : __ldg(a.init + idx(a.L, v, k + HS, i + HS));
        double x;
        if (v == WMOM) {  // hydrostatic source on the cell's rho' (interpolate.py:248-250)
            if (a.src_w)
                x = cell_update<true, true>(fb[v], ft[v], ini, a.cd, a.cg, s[2][DENS], a.dt_stage,
                                            __ldg(a.src_w + (long long)k * a.L.nx + i));
            else
                x = cell_update<true, false>(fb[v], ft[v], ini, a.cd, a.cg, s[2][DENS], a.dt_stage, 0.0);
        } else {
            x = cell_update<false, false>(fb[v], ft[v], ini, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
        }
        store_cell(a, v, k, i, x);
    }
}

}  // namespace pmw
