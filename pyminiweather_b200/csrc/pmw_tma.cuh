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
// Every thread owns TWO adjacent cells in x: shared-memory reads and global stores are
// 16-byte (LDS.128 / STG.128), and the two interface-flux evaluations a thread makes per step
// are independent dependency chains (ILP 2 through the FP64 pipe).
//
// x stage.  A warp owns one tile row and walks it right to left in passes of 64 interfaces;
// lane l of pass q evaluates the fluxes through interfaces 64q+2l and 64q+2l+1 and finalises
// the two cells to the right of them; the third flux it needs comes from lane l+1 by one
// rotating shuffle per variable (lane 31 receives the flux lane 0 kept from the pass before).
// No block barrier, no flux array, no divergent "extra interface" pass: a tile of P passes
// evaluates 64P interfaces for its 64P-2 cells.
//
// z stage.  Lanes run along x (conflict-free shared-memory rows, coalesced stores); warps own
// interface rows; the flux through a cell's top face comes from the warp above through a small
// shared-memory exchange (one block barrier per pass of 4 rows).  The stage's initial state is
// read straight from global memory while the fluxes are evaluated (same thread, same cells as
// its stores: safe for the in-place third stage), which keeps a tile at ~55 KB so that four
// CTAs are resident per SM.  The solid-wall halo rows (set_bc_z, bcs.py:92-148) are rebuilt
// in shared memory by the tiles that touch a wall, so z stages never read halo rows from HBM.
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
                                            uint64_t* bar, unsigned long long policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// Programmatic dependent launch: stage n+1 is launched while the last wave of stage n drains; its
// CTAs run their prologue (barrier init, descriptor prefetch, profile loads) and then block here
// until stage n has completed and flushed its stores.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Slab ring: spin until the neighbour's epoch flag (written over NVLink by its signal kernel after
// the stage that stored our halo columns) reaches `epoch`.  Bounded: on timeout the watchdog word
// is set and the kernel carries on, so a lost peer can never hang the GPU.
__device__ __forceinline__ void wait_epoch(unsigned long long* flags, int which, unsigned long long epoch)
{
    const long long t0 = clock64();
    unsigned long long v;
    do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + which) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 4000000000LL) {  // ~2 s
            flags[2] = 1ull;
            break;
        }
    } while (true);
    asm volatile("fence.proxy.async;" ::: "memory");  // order the acquire before the TMA (async proxy) reads
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ------------------------------------------------------------------------------------------
// helpers shared by both stage kernels
// ------------------------------------------------------------------------------------------
struct Pair {
    double a, b;
};
__device__ __forceinline__ Pair lds2(const double* p)  // p is 16-byte aligned shared memory
{
    const double2 t = *reinterpret_cast<const double2*>(p);
    return {t.x, t.y};
}

// Stores the updated cell pair (i, i+1) of interior row k for variable v, plus the periodic /
// slab-neighbour halo images when the pair is one of the two edge pairs (set_bc_x, bcs.py:35-39,
// folded into the producer).  `po` points at out[(v=0, k, i)].
__device__ __forceinline__ void store_pair(const StageArgs& a, double* po, long long voff, bool edge, int i,
                                           double x, double y, unsigned long long policy)
{
    (void)policy;  // an eviction hint on the stores never helped (tools/l2_probe.py) and costs two registers
    *reinterpret_cast<double2*>(po + voff) = make_double2(x, y);
    if (edge) {
        const long long o = (po - a.out) + voff;
        if (i == 0) *reinterpret_cast<double2*>(a.out_left + o + a.L.nx) = make_double2(x, y);
        else *reinterpret_cast<double2*>(a.out_right + o - a.L.nx) = make_double2(x, y);
    }
}

// ------------------------------------------------------------------------------------------
// x stage
//   TR   rows per tile (= warps per CTA)
//   P    passes of 64 interfaces per row; the tile owns TC = 64P-2 cells per row
//   box  forcing [4][TR][64P+4] starting at array column c0 (even: TMA needs a 16-byte aligned
//        box origin), init [4][TR][64P] starting at array column c0+2
// ------------------------------------------------------------------------------------------
template <int TR, int P>
struct XTile {
    static constexpr int TC = 64 * P - 2;
    static constexpr int FW = 64 * P + 4;
    static constexpr int IW = 64 * P;
    static constexpr int F_ELEMS = NVAR * TR * FW;
    static constexpr int I_ELEMS = NVAR * TR * IW;
    static constexpr int THREADS = 32 * TR;
    static constexpr size_t smem_bytes(bool has_init)
    {
        return (size_t)(F_ELEMS + (has_init ? I_ELEMS : 0)) * sizeof(double) + 16;
    }
};

template <int TR, int P, bool HAS_INIT, int POW_MODE, bool HAS_SRC = false>
__global__ void __launch_bounds__(32 * TR, (TR == 4 && !HAS_SRC) ? 5 : 1)  // 4-warp tiles: 5 CTAs/SM (<= 96 registers)
stage_x_tma(const __grid_constant__ CUtensorMap tm_forcing, const __grid_constant__ CUtensorMap tm_init,
            const StageArgs a)
{
    using T = XTile<TR, P>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sF = reinterpret_cast<double*>(smem_raw);
    double* sI = sF + T::F_ELEMS;
    static_assert((T::F_ELEMS * 8) % 128 == 0 && (T::I_ELEMS * 8) % 128 == 0, "tiles stay 128-byte aligned");
    uint64_t* bar = reinterpret_cast<uint64_t*>(sI + (HAS_INIT ? T::I_ELEMS : 0));

    // Launch order is blockIdx.x-fastest.  Single slab: tile (blockIdx.x, blockIdx.y), a row of tiles
    // after the other.  Slab ring (edge_last): the grid is walked column by column, columns in the
    // order 1, 2, ..., ntx-2, 0, ntx-1, so that the tiles whose halo cells arrive from a neighbour
    // over NVLink are the last CTAs of the kernel and normally never have to wait.
    // In a slab ring the kernel has one extra row of CTAs in front (blockIdx.y == 0): they push this
    // slab's own edge columns to the neighbours instead of computing a tile.
    const int ntx = gridDim.x;
    const int push_rows = a.push_epoch ? 1 : 0;
    const int nty = gridDim.y - push_rows;
    if (push_rows && blockIdx.y == 0) {
        pdl_launch_dependents();
        pdl_wait();
        push_halo_role(a);
        return;
    }
    int tx = blockIdx.x, ty = blockIdx.y - push_rows;
    if (a.edge_last) {
        const int lin = ty * ntx + tx;
        const int cidx = lin / nty;
        ty = lin % nty;
        tx = (cidx + 2 < ntx) ? cidx + 1 : (cidx + 2 == ntx ? 0 : ntx - 1);
    }
    ty += a.tile_y0;            // chunked sweeps: this launch covers a band of tile rows
    const int c0 = tx * T::TC;  // first interior column of the tile (even)
    const int r0 = ty * TR;     // first interior row
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_forcing);
        if (HAS_INIT) tma_prefetch_desc(&tm_init);
        mbar_init(bar, 1);
    }
    __syncthreads();
    pdl_wait();  // everything below reads or overwrites state produced by the previous stage
    if (threadIdx.x == 0) {
        if (a.wait_epoch) {
            if (tx == 0) wait_epoch(a.flags, 0, a.wait_epoch);
            if (tx == ntx - 1) wait_epoch(a.flags, 1, a.wait_epoch);
        }
        mbar_arrive_expect_tx(bar, (uint32_t)((T::F_ELEMS + (HAS_INIT ? T::I_ELEMS : 0)) * sizeof(double)));
        tma_load_3d(sF, &tm_forcing, c0, r0 + HS, 0, bar, l2_policy(a.hint_forcing));
        if (HAS_INIT) tma_load_3d(sI, &tm_init, c0 + HS, r0 + HS, 0, bar, l2_policy(a.hint_init));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = r0 + warp;  // interior row of this warp
    const bool row_ok = k < a.L.nz;
    const IfaceBg bg = bg_x(a.hy, min(k, a.L.nz - 1) + HS);
    const int nx = a.L.nx;
    const double* rowF = sF + warp * T::FW + 2 * lane;
    const double* rowI = sI + warp * T::IW + 2 * lane;
    double* po = a.out + idx(a.L, 0, min(k, a.L.nz - 1) + HS, c0 + HS + 2 * lane);
    const int src_lane = (lane + 1) & 31;
    const unsigned long long pol_out = l2_policy(a.hint_out);
    mbar_wait(bar, 0);

    double keep[4] = {0.0, 0.0, 0.0, 0.0};  // lane 0: its first flux of the pass to the right
#pragma unroll
    for (int q = P - 1; q >= 0; --q) {
        // interfaces li, li+1 (tile-local); taps of interface j are tile columns j..j+3
        double t0[4], t1[4], t2[4], t3[4], t4[4], f0[4], f1[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double* p = rowF + v * (TR * T::FW) + 64 * q;
            const Pair u01 = lds2(p), u23 = lds2(p + 2);
            t0[v] = u01.a; t1[v] = u01.b; t2[v] = u23.a; t3[v] = u23.b; t4[v] = p[4];
        }
        interface_flux<false, POW_MODE>(t0, t1, t2, t3, bg, a.hv_coeff, false, f0);
        interface_flux<false, POW_MODE>(t1, t2, t3, t4, bg, a.hv_coeff, false, f1);
        const int i = c0 + 64 * q + 2 * lane;  // interior column of the left cell of the pair
        // the rightmost pair of the rightmost pass has no flux to its right: it belongs to the next tile
        const bool ok = row_ok && i < nx && !(q == P - 1 && lane == 31);
        const bool edge = a.write_xhalo && (i == 0 || i == nx - 2);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double give = (lane == 0) ? keep[v] : f0[v];
            const double fr = __shfl_sync(0xffffffffu, give, src_lane);  // flux through interface li+2
            keep[v] = f0[v];
            if (ok) {
                double ia, ib;  // cell li is tile column li+2
                if (HAS_INIT) {
                    const Pair in = lds2(rowI + v * (TR * T::IW) + 64 * q);
                    ia = in.a; ib = in.b;
                } else {
                    ia = t2[v]; ib = t3[v];
                }
                double xa, xb;
                if (HAS_SRC && v == WMOM) {  // gravity-wave forcing (source.py:43-50)
                    const double2 g = __ldg(reinterpret_cast<const double2*>(a.src_w + (long long)k * nx + i));
                    xa = cell_update<false, true>(f0[v], f1[v], ia, a.cd, a.cg, 0.0, a.dt_stage, g.x);
                    xb = cell_update<false, true>(f1[v], fr, ib, a.cd, a.cg, 0.0, a.dt_stage, g.y);
                } else {
                    xa = cell_update<false, false>(f0[v], f1[v], ia, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
                    xb = cell_update<false, false>(f1[v], fr, ib, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
                }
                store_pair(a, po + 64 * q, v * a.L.vstride, edge, i, xa, xb, pol_out);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// z stage
//   tile TR x 64 cells, TR = 4*NP - 1; forcing box [4][TR+4][64] (array rows r0 .. r0+TR+3)
//   4 warps; the tile's 4*NP interface rows are processed in NP passes of 4 rows, top pass
//   first; in a pass warp w evaluates interface row 4p+w for its lane's column pair, publishes
//   the flux pair in shared memory, and after one block barrier finalises the cell row above
//   that interface with the flux of the row above it (warp w+1 of this pass, or warp 0 of the
//   pass before).  Every pass has its own flux slot (dead tile rows), so one barrier per pass suffices.
//   No interface is evaluated twice inside a tile and none of the work is serial per thread:
//   the dependency chain is one flux long, as in the x stage.
// ------------------------------------------------------------------------------------------
template <int NP>
struct ZTile {
    static constexpr int TC = 64;
    static constexpr int W = 4;                 // warps = interface rows per pass
    static constexpr int TR = W * NP - 1;       // cell rows owned by the tile
    static constexpr int FH = TR + 4;
    static constexpr int F_ELEMS = NVAR * FH * TC;
    static constexpr int X_ELEMS = NVAR * W * TC;      // flux exchange slot of the top pass (the other
                                                       // passes reuse tile rows that are already dead)
    static constexpr int H_ELEMS = 4 * W * NP;         // hydrostatic interface profiles of the tile
    static constexpr int I_ELEMS = 2 * NVAR * W * TC;  // initial state of two passes (cp.async ring)
    static constexpr int THREADS = 32 * W;
    static constexpr size_t smem_bytes(bool has_init)
    {
        return (size_t)(F_ELEMS + X_ELEMS + H_ELEMS + (has_init ? I_ELEMS : 0)) * sizeof(double) + 16;
    }
};

template <int NP, bool HAS_INIT, int POW_MODE, bool HAS_SRC = false>
__global__ void __launch_bounds__(128, HAS_INIT ? 4 : 5)
stage_z_tma(const __grid_constant__ CUtensorMap tm_forcing, const StageArgs a)
{
    using T = ZTile<NP>;
    constexpr int TC = T::TC, W = T::W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sF = reinterpret_cast<double*>(smem_raw);
    double* sX = sF + T::F_ELEMS;
    double* sH = sX + T::X_ELEMS;  // [4][W*NP]: dens, dens_theta, 1/dens_theta, pressure per interface row
    double* sIn = sH + T::H_ELEMS;  // [2][4][W][TC]: initial state of the pass in flight and the next one
    uint64_t* bar = reinterpret_cast<uint64_t*>(sIn + (HAS_INIT ? T::I_ELEMS : 0));

    const int nz = a.L.nz, nx = a.L.nx;
    const int c0 = (blockIdx.x + a.tile_x0) * TC;  // chunked sweeps: a band of tile columns per launch
    const int r0 = blockIdx.y * T::TR;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_forcing);
        mbar_init(bar, 1);
    }
    if (threadIdx.x < W * NP) {  // the tile's interface profiles (constant data: before the PDL wait)
        const int kc = min(r0 + (int)threadIdx.x, nz);
        sH[threadIdx.x] = __ldg(a.hy.dens_int + kc);
        sH[W * NP + threadIdx.x] = __ldg(a.hy.dens_theta_int + kc);
        sH[2 * W * NP + threadIdx.x] = __ldg(a.hy.inv_dens_theta_int + kc);
        sH[3 * W * NP + threadIdx.x] = __ldg(a.hy.pressure_int + kc);
    }
    __syncthreads();
    pdl_wait();  // everything below reads or overwrites state produced by the previous stage
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, (uint32_t)(T::F_ELEMS * sizeof(double)));
        tma_load_3d(sF, &tm_forcing, c0 + HS, r0, 0, bar, l2_policy(a.hint_forcing));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = 2 * lane;
    const int i = c0 + col;  // interior column of the left cell of the pair (even)
    const bool col_ok = i < nx;
    const bool edge = a.write_xhalo && (i == 0 || i == nx - 2);
    const long long colbase = idx(a.L, 0, HS, min(i, nx - 2) + HS);  // (v=0, interior row 0, pair)
    const unsigned long long pol_out = l2_policy(a.hint_out), pol_init = l2_policy(a.hint_init);
    // Initial state of this thread's cell pair, pass by pass: 16-byte cp.async copies into a
    // two-pass ring in shared memory, issued two passes ahead of their use (no registers held, no
    // barrier needed: every thread reads back only what it copied itself).
    auto fetch_init = [&](int p) {
        const int kf = r0 + W * p + warp;
        if (col_ok && kf < nz && !(p == NP - 1 && warp == W - 1)) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const double* src = a.init + colbase + v * a.L.vstride + (long long)kf * a.L.pitch;
                const uint32_t dst = smem_u32(sIn + (((p & 1) * NVAR + v) * W + warp) * TC + col);
                asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
                             "l"(pol_init)
                             : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (HAS_INIT) {
        fetch_init(NP - 1);
        if (NP >= 2) fetch_init(NP - 2);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    mbar_wait(bar, 0);

    // set_bc_z folded in: tiles touching a wall rebuild the two halo rows in shared memory
    {
        const bool bottom = (r0 == 0);
        const int top_lr = nz + HS - r0;  // tile-local row of array row nz+2
        const bool top = top_lr < T::FH;  // (array row nz+1 is then tile-local row top_lr-1 >= 2)
        if (a.fuse_bc_z && (bottom || top)) {
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
        }
        __syncthreads();  // profiles (and rebuilt wall rows) visible to every warp
    }

#pragma unroll
    for (int p = NP - 1; p >= 0; --p) {
        const int lf = W * p + warp;  // tile-local interface row = tile-local cell row above it
        const int kf = r0 + lf;       // global interface index; bottom face of interior cell row kf
        const bool cell_ok = col_ok && kf < nz && !(p == NP - 1 && warp == W - 1);
        // taps: tile rows lf .. lf+3; A = left column of the pair, B = right column
        double a0[4], a1[4], a2[4], a3[4], b0[4], b1[4], b2[4], b3[4], fa[4], fb[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const double* q = sF + (v * T::FH + lf) * TC + col;
            const Pair r0v = lds2(q), r1v = lds2(q + TC), r2v = lds2(q + 2 * TC), r3v = lds2(q + 3 * TC);
            a0[v] = r0v.a; b0[v] = r0v.b; a1[v] = r1v.a; b1[v] = r1v.b;
            a2[v] = r2v.a; b2[v] = r2v.b; a3[v] = r3v.a; b3[v] = r3v.b;
        }
        const int kc = min(kf, nz);
        IfaceBg bg;
        bg.dens = sH[lf]; bg.dens_theta = sH[W * NP + lf];
        bg.inv_dens_theta = sH[2 * W * NP + lf]; bg.pressure = sH[3 * W * NP + lf];
        const bool wall = (kc == 0 || kc == nz);
        interface_flux<true, POW_MODE>(a0, a1, a2, a3, bg, a.hv_coeff, wall, fa);
        interface_flux<true, POW_MODE>(b0, b1, b2, b3, bg, a.hv_coeff, wall, fb);
        // Flux exchange slot of pass p: [v][warp][col].  The top pass has its own small buffer; pass
        // p < NP-1 uses tile rows 4p+7 .. 4p+10 of each variable plane, which only pass p+1 read and
        // which are dead once every warp is past that pass's barrier.  Slots of different passes
        // never overlap, so one barrier per pass is enough.
        double* const xbase = (p == NP - 1) ? sX : sF + (W * p + 7) * TC;
        constexpr int xstride_top = W * TC, xstride_f = T::FH * TC;
        const int xstride = (p == NP - 1) ? xstride_top : xstride_f;
        double* xw = xbase + warp * TC + col;
#pragma unroll
        for (int v = 0; v < 4; ++v)
            *reinterpret_cast<double2*>(xw + v * xstride) = make_double2(fa[v], fb[v]);
        __syncthreads();
        if (HAS_INIT) asm volatile("cp.async.wait_group 1;" ::: "memory");  // this pass's initial state has landed
        if (cell_ok) {
            // flux through the top face: interface row lf+1 = warp w+1 of this pass, or warp 0 of the
            // pass above (p+1)
            const double* const ubase = (p + 1 == NP - 1) ? sX : sF + (W * (p + 1) + 7) * TC;
            const int ustride = (p + 1 == NP - 1) ? xstride_top : xstride_f;
            const double* xr = (warp < W - 1) ? xw + TC : ubase + col;
            const int rstride = (warp < W - 1) ? xstride : ustride;
            double* po = a.out + colbase + (long long)kf * a.L.pitch;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const Pair up = lds2(xr + v * rstride);
                double ia = a2[v], ib = b2[v];
                if (HAS_INIT) {
                    const Pair in = lds2(sIn + (((p & 1) * NVAR + v) * W + warp) * TC + col);
                    ia = in.a; ib = in.b;
                }
                double xa, xb;
                if (v == WMOM) {  // hydrostatic source (interpolate.py:248-250); cell kf is tap 2
                    if (HAS_SRC) {  // gravity-wave forcing (source.py:43-50)
                        const double2 g = __ldg(reinterpret_cast<const double2*>(a.src_w + (long long)kf * nx + i));
                        xa = cell_update<true, true>(fa[v], up.a, ia, a.cd, a.cg, a2[DENS], a.dt_stage, g.x);
                        xb = cell_update<true, true>(fb[v], up.b, ib, a.cd, a.cg, b2[DENS], a.dt_stage, g.y);
                    } else {
                        xa = cell_update<true, false>(fa[v], up.a, ia, a.cd, a.cg, a2[DENS], a.dt_stage, 0.0);
                        xb = cell_update<true, false>(fb[v], up.b, ib, a.cd, a.cg, b2[DENS], a.dt_stage, 0.0);
                    }
                } else {
                    xa = cell_update<false, false>(fa[v], up.a, ia, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
                    xb = cell_update<false, false>(fb[v], up.b, ib, a.cd, a.cg, 0.0, a.dt_stage, 0.0);
                }
                store_pair(a, po, v * a.L.vstride, edge, i, xa, xb, pol_out);
            }
        }
        if (HAS_INIT) {  // refill the ring slot just consumed with the pass two below (or an empty group)
            if (p >= 2) fetch_init(p - 2);
            else asm volatile("cp.async.commit_group;" ::: "memory");
        }
    }
}

}  // namespace pmw
