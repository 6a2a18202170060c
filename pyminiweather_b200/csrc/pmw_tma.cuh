// Variant 1 ("tma"): the production stage kernels for sm_100a.
//
// One CTA = one tile.  A single elected thread issues one TMA tiled load
// (cp.async.bulk.tensor.3d, SASS UTMALDG) per operand -- the forcing tile with its 2-cell
// stencil halo for all four variables, and (RK stages 2 and 3) the matching tile of the
// stage's initial state -- completing on an mbarrier; every thread then computes from shared
// memory and stores the updated state straight from registers with fully coalesced 8-byte
// stores.  Interpolated values, fluxes and tendencies only ever live in registers, so a stage
// moves exactly: forcing once (+ halo re-reads that hit L2), init once, out once.
// Several CTAs are resident per SM, so while one tile computes the TMA loads of the others
// are in flight: occupancy, not a software pipeline, hides the HBM latency.
//
// x stage.  A warp owns one tile row and marches along it in passes of 32 interfaces; lane l
// of pass q evaluates the flux through interface 32q+l and finalises the cell to the LEFT of
// it with the neighbour lane's flux (one __shfl_up per variable; the flux of lane 31 is
// carried into the next pass).  A tile of P passes therefore has 32P interfaces and 32P-1
// cells: no divergent "extra interface" pass, no block barrier, no flux array.
//
// z stage.  Lanes run along x (conflict-free shared-memory rows, coalesced stores); a thread
// marches up RPT rows of one column with a 4-row register window per variable, reusing the
// previous interface flux, so only 1 in RPT+1 fluxes is recomputed by the thread above.  The
// solid-wall halo rows (set_bc_z, bcs.py:92-148) are rebuilt in shared memory by the tiles
// that touch a wall, so z stages never read halo rows from HBM.
#pragma once
#include <cuda.h>

#include "pmw_direct.cuh"

namespace pmw {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    // make the initialised barrier visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// 3-D tiled TMA load global -> shared, completing `bar` with the box byte count.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z,
                                            uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ------------------------------------------------------------------------------------------
// x stage
//   TR   rows per tile (= warps per CTA)
//   P    passes of 32 interfaces per row; the tile owns TC = 32P-1 cells per row
//   box  forcing [4][TR][32P+4], init [4][TR][32P].  TMA needs the box to start on a 16-byte
//        boundary in global memory, i.e. at an even column; tiles start at multiples of the odd
//        number 32P-1, so odd tiles load from one column further left (`off` = 1) and index
//        shared memory one column further right.  The 32P+3 columns a tile can touch fit the
//        32P+4 wide box (whose extent must be a multiple of 16 bytes anyway).
// ------------------------------------------------------------------------------------------
template <int TR, int P>
struct XTile {
    static constexpr int TC = 32 * P - 1;
    static constexpr int FW = 32 * P + 4;
    static constexpr int IW = 32 * P;
    static constexpr int F_ELEMS = NVAR * TR * FW;
    static constexpr int I_ELEMS = NVAR * TR * IW;
    static constexpr int THREADS = 32 * TR;
    static constexpr size_t smem_bytes(bool has_init)
    {
        return (size_t)(F_ELEMS + (has_init ? I_ELEMS : 0)) * sizeof(double) + 128;
    }
};

template <int TR, int P, bool HAS_INIT, int POW_MODE>
__global__ void __launch_bounds__(32 * TR)
stage_x_tma(const __grid_constant__ CUtensorMap tm_forcing, const __grid_constant__ CUtensorMap tm_init,
            const StageArgs a)
{
    using T = XTile<TR, P>;
    extern __shared__ unsigned char smem_raw[];
    // TMA destinations need 128-byte alignment
    double* sF = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    double* sI = sF + T::F_ELEMS;  // F_ELEMS*8 is a multiple of 128 (FW*8*4 = 128*(P+...)): see static_assert
    static_assert((T::F_ELEMS * 8) % 128 == 0, "init tile must stay 128-byte aligned");
    __shared__ __align__(8) uint64_t bar;

    const int c0 = blockIdx.x * T::TC;  // first interior column of the tile
    const int r0 = blockIdx.y * TR;     // first interior row
    const int off = c0 & 1;             // box starts at the even column c0 - off
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_forcing);
        if (HAS_INIT) tma_prefetch_desc(&tm_init);
        mbar_init(&bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)((T::F_ELEMS + (HAS_INIT ? T::I_ELEMS : 0)) * sizeof(double)));
        tma_load_3d(sF, &tm_forcing, c0 - off, r0 + HS, 0, &bar);
        if (HAS_INIT) tma_load_3d(sI, &tm_init, c0 - off + HS, r0 + HS, 0, &bar);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = r0 + warp;  // interior row of this warp
    const bool row_ok = k < a.L.nz;
    const IfaceBg bg = bg_x(a.hy, min(k, a.L.nz - 1) + HS);
    mbar_wait(&bar, 0);

    const double* rowF = sF + warp * T::FW + off;
    const double* rowI = sI + warp * T::IW + off;
    double carry[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const int li = 32 * q + lane;  // tile-local interface; its taps are tile columns li..li+3
        double s0[4], s1[4], s2[4], s3[4], flux[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double* p = rowF + v * (TR * T::FW) + li;
            s0[v] = p[0]; s1[v] = p[1]; s2[v] = p[2]; s3[v] = p[3];
        }
        interface_flux<false, POW_MODE>(s0, s1, s2, s3, bg, a.hv_coeff, false, flux);
        const int cell = li - 1;   // tile-local cell to the left of interface li
        const int i = c0 + cell;   // interior column
        const bool ok = row_ok && cell >= 0 && i < a.L.nx;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            double fl = __shfl_up_sync(0xffffffffu, flux[v], 1);
            if (lane == 0) fl = carry[v];
            carry[v] = __shfl_sync(0xffffffffu, flux[v], 31);
            if (ok) {
                // forcing value of cell `cell` is tile column cell+2 = li+1 = tap s1
                const double ini = HAS_INIT ? rowI[v * (TR * T::IW) + cell] : s1[v];
                const double tend = (fl - flux[v]) * a.inv_d;
                store_cell(a, v, k, i, fma(a.dt_stage, tend, ini));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// z stage
//   tile TR x TC cells; forcing box [4][TR+4][TC] (array rows r0 .. r0+TR+3), init [4][TR][TC]
//   thread = (column, row group); RPT rows per thread; threads = TC * TR/RPT
// ------------------------------------------------------------------------------------------
template <int TR, int TC, int RPT>
struct ZTile {
    static_assert(TR % RPT == 0 && TC % 32 == 0, "tile shape");
    static constexpr int FH = TR + 4;
    static constexpr int F_ELEMS = NVAR * FH * TC;
    static constexpr int I_ELEMS = NVAR * TR * TC;
    static constexpr int THREADS = TC * (TR / RPT);
    static constexpr size_t smem_bytes(bool has_init)
    {
        return (size_t)(F_ELEMS + (has_init ? I_ELEMS : 0)) * sizeof(double) + 128;
    }
};

template <int TR, int TC, int RPT, bool HAS_INIT, int POW_MODE>
__global__ void __launch_bounds__(TC * (TR / RPT))
stage_z_tma(const __grid_constant__ CUtensorMap tm_forcing, const __grid_constant__ CUtensorMap tm_init,
            const StageArgs a)
{
    using T = ZTile<TR, TC, RPT>;
    extern __shared__ unsigned char smem_raw[];
    double* sF = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    double* sI = sF + T::F_ELEMS;
    static_assert((T::F_ELEMS * 8) % 128 == 0, "init tile must stay 128-byte aligned");
    __shared__ __align__(8) uint64_t bar;

    const int nz = a.L.nz;
    const int c0 = blockIdx.x * TC;
    const int r0 = blockIdx.y * TR;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_forcing);
        if (HAS_INIT) tma_prefetch_desc(&tm_init);
        mbar_init(&bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)((T::F_ELEMS + (HAS_INIT ? T::I_ELEMS : 0)) * sizeof(double)));
        tma_load_3d(sF, &tm_forcing, c0 + HS, r0, 0, &bar);
        if (HAS_INIT) tma_load_3d(sI, &tm_init, c0 + HS, r0 + HS, 0, &bar);
    }
    const int col = threadIdx.x % TC;
    const int grp = threadIdx.x / TC;
    const int i = c0 + col;
    mbar_wait(&bar, 0);

    // set_bc_z folded in: tiles touching a wall rebuild the two halo rows in shared memory
    if (a.fuse_bc_z) {
        const bool bottom = (r0 == 0);
        const int top_lr = nz + HS - r0;  // tile-local row of array row nz+2
        const bool top = top_lr < T::FH;  // (array row nz+1 is then tile-local row top_lr-1 >= 2)
        if (bottom || top) {
            for (int e = threadIdx.x; e < NVAR * 2 * TC; e += T::THREADS) {
                const int c = e % TC, j = (e / TC) & 1, v = e / (2 * TC);
                if (bottom)
                    sF[(v * T::FH + j) * TC + c] = wall_value(v, sF[(v * T::FH + HS) * TC + c],
                                                              __ldg(a.hy.dens_cell + HS),
                                                              __ldg(a.hy.dens_cell + j));
                if (top && top_lr + j < T::FH)
                    sF[(v * T::FH + top_lr + j) * TC + c] =
                        wall_value(v, sF[(v * T::FH + top_lr - 1) * TC + c],
                                   __ldg(a.hy.dens_cell + nz + HS - 1),
                                   __ldg(a.hy.dens_cell + nz + HS + j));
            }
            __syncthreads();
        }
    }

    const int lr0 = grp * RPT;  // first tile-local cell row of this thread
    // interface r0+lr0+j (bottom face of cell row lr0+j) uses tile rows lr0+j .. lr0+j+3
    double w0[4], w1[4], w2[4], w3[4], fprev[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const double* p = sF + (v * T::FH + lr0) * TC + col;
        w0[v] = p[0]; w1[v] = p[TC]; w2[v] = p[2 * TC];
        fprev[v] = 0.0;
    }
#pragma unroll
    for (int j = 0; j <= RPT; ++j) {
        const int kf = r0 + lr0 + j;  // global interface index
        double flux[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) w3[v] = sF[(v * T::FH + lr0 + j + 3) * TC + col];
        const int kc = min(kf, nz);
        interface_flux<true, POW_MODE>(w0, w1, w2, w3, bg_z(a.hy, kc), a.hv_coeff, kc == 0 || kc == nz, flux);
        if (j > 0) {
            const int k = kf - 1;  // interior cell row below interface kf; its forcing value is tap w1
            if (k < nz && i < a.L.nx) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    double tend = (fprev[v] - flux[v]) * a.inv_d;
                    if (v == WMOM) tend = fma(-w1[DENS], GRAV, tend);
                    const double ini = HAS_INIT ? sI[(v * TR + lr0 + j - 1) * TC + col] : w1[v];
                    store_cell(a, v, k, i, fma(a.dt_stage, tend, ini));
                }
            }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            fprev[v] = flux[v];
            w0[v] = w1[v]; w1[v] = w2[v]; w2[v] = w3[v];
        }
    }
}

}  // namespace pmw
