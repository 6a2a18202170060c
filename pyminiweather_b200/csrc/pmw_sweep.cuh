// Fused directional sweeps: all three low-storage RK stages of one direction (step.py:112-141)
// in ONE kernel, with the two intermediate states kept on chip.
//
// The three stages of a sweep only couple cells along the sweep direction:
//     T1  = S + dt/3 * T(S)        T2 = S + dt/2 * T(T1)        S' = S + dt * T(T2)
// so a tile that carries a 6-cell halo along that direction (2 cells per stage) can run all of
// them without touching HBM in between: the state is read once and written once per sweep,
// 64 B/cell instead of the 256 B/cell the stage-by-stage kernels move.  At that traffic the
// sweep is no longer HBM-bound on B200 but bound by the FP64 pipe (57 FP64 instructions per
// interface + 8 per cell update) and, at the occupancy the register file allows, by its latency, so
// the kernels below are organised around instruction count, independent FP64 chains and FP64
// issue, not around bytes.  The arithmetic per cell-stage is interface_flux + the same update expressions as
// the stage kernels (pmw_tma.cuh): results are bit-identical to the stage-by-stage path, halo
// cells of the intermediate states are simply recomputed by the neighbouring tile.
//
// x sweep (sweep_x).  Persistent and warp-autonomous: a work item is one row x L = 64P-10 owned
// cells; a warp brings the item's state row with its 6-column halo (64P+4 columns, all four
// variables) into shared memory with one TMA box load (the next item's row is in flight
// meanwhile), runs stage 1 over 64P-2 cells into a shared-memory row T1, stage 2 over 64P-6 cells
// into T2, stage 3 over its 64P-10 owned cells straight to HBM -- each stage with the pass
// structure of stage_x_tma (two interfaces per lane, the third flux by a rotating shuffle),
// separated only by __syncwarp.  Warps never synchronise with each other.  Periodic x: the halo columns are the 6-wide image of the opposite edge, written
// by whichever sweep produced the state (set_bc_x, bcs.py:35-39, folded into the producer).
//
// z sweep (sweep_z).  A warp owns a strip of 32 columns (one per lane) and streams upwards through
// a segment of LZ rows: per iteration it advances THREE software-pipelined interface evaluations
// -- stage 1 at interface j, stage 2 at j-3, stage 3 at j-6 -- whose 4-row stencil windows live
// in registers and are fed by the stage below (T1 and T2 never leave the register file).  The
// three evaluations of an iteration are independent (ILP 3).  State rows arrive through a 16-slot
// ring of 1 KB TMA boxes (one mbarrier per slot, issued 6 rows ahead by lane 0); the ring also
// serves the initial-state reads of stages 2 and 3.  Solid-wall halo rows (set_bc_z,
// bcs.py:92-148) are rebuilt in the register windows, so halo rows are never read.  A segment
// recomputes 4 + 2 rows of T1 / T2 on either side (none at a wall).
#pragma once
#include <type_traits>

#include "pmw_tma.cuh"

namespace pmw {

constexpr int SWEEP_HALO = 6;  // x halo columns a fused x sweep reads (3 stages x 2 cells)

struct SweepArgs {
    Layout L;
    const double* state;  // S  (x sweeps: with a valid 6-wide x halo)
    double* out;          // S' (a different buffer: neighbouring tiles still read S)
    double* tmp;          // T2 of the owned cells when WRITE_TMP (the reference's state_tmp), else unused
    Hydro hy;
    double hv_coeff;  // -hv_beta*d/(16*dt_full)   (interpolate.py:101,149)
    double inv_d;     // 1/dx or 1/dz
    double dt1, dt2, dt3;  // dt/3, dt/2, dt
    int periodic;     // also store the 6-wide periodic images of the edge columns of S'
    int lz;           // z sweep: rows per segment
    int tile_x0, tile_y0;
    // slab ring (x sweeps): see StageArgs
    unsigned long long* flags;
    unsigned long long wait_epoch, push_epoch;
    int edge_last;
    double* nbr_state_left;
    double* nbr_state_right;
    unsigned long long* nbr_flags_left;
    unsigned long long* nbr_flags_right;
    unsigned int* push_counter;
    const double* src_w;  // HAS_SRC instantiations: [nz][nx] extra rho*w tendency of every stage (ic_type
                          // "gravity", source.py:43-50); single periodic slab only
    double cd1, cd2, cd3;  // dt_s / d    per stage (cell_update, pmw_common.cuh)
    double cg1, cg2, cg3;  // dt_s * grav per stage
    // x sweep, dynamic item assignment (slab ring; nullptr = static walk): a counter that is never reset; this
    // launch owns the values item_base .. item_base + nitems + warps - 1 (every warp draws once at the
    // start and once per item it processes)
    unsigned long long* item_counter;
    unsigned long long item_base;
    // interior rows [row0, row1) this launch produces (whole sweep: 0, nz).  A band of rows is what pmw_evolve_host
    // streams: the sweeps only couple rows through the 6-row halo a z sweep reads from the state, so a band's result
    // has the bits of the whole sweep's.  (At the end of the struct: the members above keep their alignment.)
    int row0, row1;
};

// Slab ring, fused x sweep: where the halo columns travel.  Every context owns a STAGING area behind its flag words
// (same allocation, so the neighbours reach it through the IPC mapping of the flags):
//     halo_stage[parity][side][variable][row][6 columns]   doubles, dense;   side 0: my left halo, 1: my right halo
// The neighbours WRITE it (push_halo6_role, contiguous 16-byte stores: full 128-byte lines over NVLink -- the
// strided 48-byte pieces of the state array's own halo columns cost 16 x as many packets and made the push of a
// tall slab last longer than the sweep), the edge tiles of the x sweep READ it (halo_patch_row) into their
// shared-memory row after the neighbour's epoch has arrived.  Two parities: a neighbour may already push sweep e+1
// while this slab still reads sweep e, never e+2 (it needs this slab's epoch e+1 to get there).
constexpr int HALO_STAGE_HEADER = 16;  // u64 words in front of the staging area (flags[0..3] + padding: 128 bytes)
__host__ __device__ inline size_t halo_stage_bytes(int nz)
{
    return HALO_STAGE_HEADER * sizeof(unsigned long long) + (size_t)2 * 2 * NVAR * nz * SWEEP_HALO * sizeof(double);
}
__device__ __forceinline__ double* halo_stage(unsigned long long* flags, int nz, unsigned long long epoch, int side)
{
    return reinterpret_cast<double*>(flags + HALO_STAGE_HEADER) +
           ((size_t)(epoch & 1) * 2 + side) * ((size_t)NVAR * nz * SWEEP_HALO);
}

// The first CTAs of an x sweep store this slab's own six edge columns of S into the neighbours' staging areas
// and publish push_epoch (cf. push_halo_role).
__device__ __forceinline__ void push_halo6_role(const SweepArgs& a, int npush)
{
    const int tid = threadIdx.x, nthr = blockDim.x;
    const Layout& L = a.L;
    const int per_row = SWEEP_HALO / 2;  // column pairs per side
    const int total = NVAR * L.nz * per_row, stride = npush * nthr;
    // four elements per thread and iteration, loads first: the loop is bound by the latency of its (cold, 48-byte)
    // reads, and the epoch cannot be published before the last store is acknowledged
    // our first columns are the left neighbour's RIGHT halo (side 1); our last columns the right neighbour's LEFT halo
    double2* const to_left = reinterpret_cast<double2*>(halo_stage(a.nbr_flags_left, L.nz, a.push_epoch, 1));
    double2* const to_right = reinterpret_cast<double2*>(halo_stage(a.nbr_flags_right, L.nz, a.push_epoch, 0));
    for (int t0 = blockIdx.x * nthr + tid; t0 < total; t0 += 4 * stride) {
        double2 first[4], last[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = min(t0 + u * stride, total - 1);  // t = (v * nz + k) * 3 + j: the staging index itself
            const int j = t % per_row, k = (t / per_row) % L.nz, v = t / (per_row * L.nz);
            first[u] = *reinterpret_cast<const double2*>(a.state + idx(L, v, k + HS, HS + 2 * j));
            last[u] = *reinterpret_cast<const double2*>(a.state + idx(L, v, k + HS, L.nx + HS - SWEEP_HALO + 2 * j));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * stride;
            if (t < total) {
                to_left[t] = first[u];
                to_right[t] = last[u];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && atomicAdd(a.push_counter, 1u) == (unsigned)npush - 1) {
        *a.push_counter = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.nbr_flags_left + 1), "l"(a.push_epoch) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.nbr_flags_right + 0), "l"(a.push_epoch) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// x sweep
//   Work item = one interior row k x one column tile of LC = 64P-10 owned cells (tile column t holds
//   interior column c0-6+t, FW = 64P+4 tile columns).  Persistent kernel: every WARP walks its own
//   list of items (grid-stride) and never synchronises with another warp.  Per warp, in shared
//   memory: two state rows S[2][4][FW] -- the row of the next item is in flight (one TMA box, own
//   mbarrier) while the current one is computed -- and the rows T1[4][FW], T2[4][FW] of the
//   intermediate states; stage 3 stores the owned cells to HBM.  (The transposing z sweep further
//   down keeps T1 and T2 in ONE row, in place; here the second row is affordable and removes the
//   warp-level ordering that needs.)
// ------------------------------------------------------------------------------------------
template <int P>
struct XSweepTile {
    static constexpr int FW = 64 * P + 4;   // tile columns (state box width)
    static constexpr int LC = 64 * P - 10;  // owned cells per row
#ifndef PMW_XSWEEP_WARPS
#define PMW_XSWEEP_WARPS 4
#endif
    static constexpr int WARPS = PMW_XSWEEP_WARPS;  // per CTA (no block-level synchronisation: any number works)
    static constexpr int S_ELEMS = NVAR * FW;
    static constexpr int WARP_ELEMS = 4 * S_ELEMS;  // S[2] (double-buffered state row), T1, T2
    // per warp: the four row buffers, two mbarriers, and the hydrostatic profiles of the two rows in flight (2 x 32 B)
    static constexpr size_t smem_bytes() { return (size_t)WARPS * (WARP_ELEMS * sizeof(double) + 16 + 64); }
};

// Both interface fluxes of a lane's pair with ONE warp-uniform fallback branch (see
// interface_flux_fast).
template <int POW_MODE>
__device__ __forceinline__ void xpair_fix(bool bad0, bool bad1, const double (&t0)[4], const double (&t1)[4],
                                          const double (&t2)[4], const double (&t3)[4], const double (&t4)[4],
                                          const IfaceBg& bg, double hv, double (&f0)[4], double (&f1)[4])
{
    Taps T;
    if (bad0) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { T.s[0][v] = t0[v]; T.s[1][v] = t1[v]; T.s[2][v] = t2[v]; T.s[3][v] = t3[v]; }
        const Flux4 g = interface_flux_slow<false, POW_MODE>(T, bg, hv, false);
#pragma unroll
        for (int v = 0; v < 4; ++v) f0[v] = g.f[v];
    }
    if (bad1) {
#pragma unroll
        for (int v = 0; v < 4; ++v) { T.s[0][v] = t1[v]; T.s[1][v] = t2[v]; T.s[2][v] = t3[v]; T.s[3][v] = t4[v]; }
        const Flux4 g = interface_flux_slow<false, POW_MODE>(T, bg, hv, false);
#pragma unroll
        for (int v = 0; v < 4; ++v) f1[v] = g.f[v];
    }
}
template <int POW_MODE>
__device__ __forceinline__ void xpair_flux(const double (&t0)[4], const double (&t1)[4], const double (&t2)[4],
                                           const double (&t3)[4], const double (&t4)[4], const IfaceBg& bg, double hv,
                                           double (&f0)[4], double (&f1)[4])
{
    const bool bad0 = interface_flux_fast<false, POW_MODE>(t0, t1, t2, t3, bg, hv, false, f0);
    const bool bad1 = interface_flux_fast<false, POW_MODE>(t1, t2, t3, t4, bg, hv, false, f1);
    if (__any_sync(0xffffffffu, bad0 || bad1)) {
        asm volatile("" ::: "memory");  // keep the argument copies of the cold path inside the branch
        xpair_fix<POW_MODE>(bad0, bad1, t0, t1, t2, t3, t4, bg, hv, f0, f1);
    }
}

struct XItem {
    int k, c0, tx;
};
// Items are numbered tile column by tile column (rows fastest); in a slab ring the two edge columns,
// whose halo cells arrive from the neighbours over NVLink, come last.
__device__ __forceinline__ XItem xsweep_item(int n, int nrows, int row0, int ntx, int lc, int edge_last)
{
    XItem it;
    const int cidx = n / nrows;
    it.k = row0 + n - cidx * nrows;
    it.tx = !edge_last ? cidx : ((cidx + 2 < ntx) ? cidx + 1 : (cidx + 2 == ntx ? 0 : ntx - 1));
    it.c0 = it.tx * lc;
    return it;
}

#ifndef PMW_XSWEEP_MINB
#define PMW_XSWEEP_MINB 3
#endif
template <int P, int POW_MODE, bool WRITE_TMP, bool HAS_SRC = false, bool DYNAMIC = false>
#ifdef PMW_XSWEEP_MAXNREG
__global__ void __maxnreg__(PMW_XSWEEP_MAXNREG)
#else
__global__ void __launch_bounds__(32 * XSweepTile<P>::WARPS, HAS_SRC ? 2 : PMW_XSWEEP_MINB)
#endif
sweep_x(const __grid_constant__ CUtensorMap tm_row, const SweepArgs a, const int ntx, const int npush)
{
    using T = XSweepTile<P>;
    constexpr int FW = T::FW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* const sS = reinterpret_cast<double*>(smem_raw) + warp * T::WARP_ELEMS;
    double* const sT = sS + 2 * T::S_ELEMS;  // T1, then T2
    uint64_t* const bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) + T::WARPS * T::WARP_ELEMS) + 2 * warp;
    double* const sBg = reinterpret_cast<double*>(smem_raw) + T::WARPS * (T::WARP_ELEMS + 2) + 8 * warp;  // [2][4]
    static_assert((T::S_ELEMS * 8) % 128 == 0, "state rows stay 128-byte aligned");

    const int nx = a.L.nx, nz = a.L.nz;
    const int row0 = a.row0, nrows = a.row1 - a.row0;
    const int nitems = nrows * ntx;
    const int nwarps = gridDim.x * T::WARPS;
    const int w = blockIdx.x * T::WARPS + warp;
    pdl_launch_dependents();
    if (lane == 0) {
        tma_prefetch_desc(&tm_row);
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
    }
    // columns of T1 / T2 no stage writes (zero: they only feed interfaces whose results are discarded):
    // T1 0, 1 and 64P .. 64P+3; T2 0 .. 3 and 64P-2 .. 64P+3
    for (int e = lane; e < NVAR * 16; e += 32) {
        const int v = e >> 4, j = e & 15;
        if (j < 6) sT[v * FW + (j < 2 ? j : 64 * P - 2 + j)] = 0.0;
        else sT[T::S_ELEMS + v * FW + (j < 10 ? j - 6 : 64 * P - 12 + j)] = 0.0;
    }
    __syncwarp();
    pdl_wait();  // everything below reads state produced by the previous kernel
    if (a.push_epoch && blockIdx.x < npush) push_halo6_role(a, npush);

    const unsigned long long pol = l2_policy(1);
    auto request = [&](const XItem& it, int buf) {  // lane 0: start the load of the item's state row
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the old row before the TMA write
        mbar_arrive_expect_tx(bars + buf, (uint32_t)((T::S_ELEMS + 4) * sizeof(double)));
        // map column 0 is interior column -6 (array column -4)
        tma_load_3d(sS + buf * T::S_ELEMS, &tm_row, it.c0, it.k + HS, 0, bars + buf, pol);
        // the row's hydrostatic profiles ride along (32 bytes of Hydro::cell_pack onto the same mbarrier)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                     ::"r"(smem_u32(sBg + 4 * buf)), "l"(a.hy.cell_pack + 4 * (long long)(it.k + HS)), "r"(smem_u32(bars + buf))
                     : "memory");
    };
    const int src_lane = (lane + 1) & 31;
    int buf = 0;
    unsigned phase = 0;  // bit b: parity of the next completion of bars[b]
    // Item assignment.  Single slab: static grid-stride walk (warp w takes items w, w + nwarps, ...).
    // Slab ring (DYNAMIC instantiations): after its first, static item a warp DRAWS the next item
    // number from a global counter (lane 0; the draw for the item after next is in flight while an item
    // is computed, and its raw result is only decoded one item later, so the L2 round trip never stalls
    // the in-order instruction stream).  The CTAs that first push the slab's edge columns to the
    // neighbours start their items several microseconds late; statically they would finish that much
    // later than everybody else, dynamically they simply draw fewer items (N=2: 126.8 -> 121.4 us/step).
    // Item order is preserved, so the edge tile columns still come last.  On a single slab the static
    // walk is 1.5 us faster per sweep and stays (compile-time switch: no cost there).  The counter is 32 bits wide, never reset and compared
    // modulo 2^32: draw number v of this launch is item nwarps + v.
    constexpr bool dynamic = DYNAMIC;
    unsigned int* const ctr = reinterpret_cast<unsigned int*>(a.item_counter);
    const unsigned int base32 = (unsigned int)a.item_base;
    auto draw = [&]() -> unsigned int { return (dynamic && lane == 0) ? atomicAdd(ctr, 1u) : 0u; };
    auto decode = [&](unsigned int raw_l0) -> int {
        const unsigned int v = __shfl_sync(0xffffffffu, raw_l0, 0) - base32 + (unsigned int)nwarps;
        return (int)min(v, (unsigned int)nitems);
    };
    unsigned int raw1 = draw();
    int n = min(w, nitems);
    if (n < nitems && lane == 0) request(xsweep_item(n, nrows, row0, ntx, T::LC, a.edge_last), 0);
#pragma unroll 1
    while (n < nitems) {
        const XItem it = xsweep_item(n, nrows, row0, ntx, T::LC, a.edge_last);
        int n_next;
        if (dynamic) {
            n_next = decode(raw1);
            raw1 = draw();  // decoded one item from now
        } else {
            n_next = min(n + nwarps, nitems);
        }
        if (n_next < nitems && lane == 0) request(xsweep_item(n_next, nrows, row0, ntx, T::LC, a.edge_last), buf ^ 1);
        n = n_next;
        // ragged last tile of a row: stage s only needs its output columns t < rem + 12 - 2s, and a pass
        // q only matters while 64q <= that limit (warp-uniform)
        const int rem = min(nx - it.c0, T::LC);
        const double* const rowS = sS + buf * T::S_ELEMS + 2 * lane;
        double* const rowT = sT + 2 * lane;
        double* const po = a.out + idx(a.L, 0, it.k + HS, it.c0 - SWEEP_HALO + HS + 2 * lane);
        const long long tmp_off = a.tmp - a.out;
        const int i0 = it.c0 - SWEEP_HALO + 2 * lane + 2;  // interior column of this lane's pair in pass 0
        mbar_wait(bars + buf, (phase >> buf) & 1);
        phase ^= 1u << buf;
        if (a.wait_epoch) {
            // slab ring: the TMA box brought whatever the state array holds in its halo columns; the real halo cells
            // of an edge tile come from the neighbours' pushes (staging area, L2-coherent loads: L1 may still hold the
            // lines of two sweeps ago), once their epoch is there
            const bool need_l = it.c0 < SWEEP_HALO, need_r = it.c0 + T::LC + SWEEP_HALO > nx;
            if (need_l || need_r) {
                if (lane == 0) {
                    if (need_l) wait_epoch(a.flags, 0, a.wait_epoch);
                    if (need_r) wait_epoch(a.flags, 1, a.wait_epoch);
                }
                __syncwarp();
                const int side = lane / 12, e = lane - 12 * side, v = e / 3, j = e - 3 * v;  // lanes 0-11: left, 12-23: right
                if (lane < 24 && (side == 0 ? need_l : need_r)) {
                    const double2* st = reinterpret_cast<const double2*>(halo_stage(a.flags, nz, a.wait_epoch, side)) +
                                        ((size_t)v * nz + it.k) * 3 + j;
                    double2 h;
                    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(h.x), "=d"(h.y) : "l"(st) : "memory");
                    const int tcol = (side == 0 ? 0 : nx - it.c0 + SWEEP_HALO) + 2 * j;  // tile column of the halo pair
                    *reinterpret_cast<double2*>(sS + buf * T::S_ELEMS + v * FW + tcol) = h;
                }
                __syncwarp();
            }
        }
        IfaceBg bg;
        {
            const double2 b01 = *reinterpret_cast<const double2*>(sBg + 4 * buf), b23 = *reinterpret_cast<const double2*>(sBg + 4 * buf + 2);
            bg.dens = b01.x; bg.dens_theta = b01.y; bg.inv_dens_theta = b23.x; bg.pressure = b23.y;
        }

        const double* src = rowS;  // forcing row of the stage (+ 2*lane)
        double dts = a.dt1, cds = a.cd1;
        int tlo = 2, thi = 64 * P;
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
            const int nq = min(P, (rem + 10 - 2 * s) / 64 + 1);
            double keep[4] = {0.0, 0.0, 0.0, 0.0};  // lane 0: its first flux of the pass to the right
            const uint32_t tdst32 = smem_u32(rowT) + (uint32_t)(s * T::S_ELEMS * sizeof(double));  // T1 / T2 row (+ 2*lane)
            // the five taps of a lane's interface pair in pass q
            auto load_taps = [&](int q, double (&t0)[4], double (&t1)[4], double (&t2)[4], double (&t3)[4], double (&t4)[4]) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const double* p = src + v * FW + 64 * q;
                    const Pair u01 = lds2(p), u23 = lds2(p + 2);
                    t0[v] = u01.a; t1[v] = u01.b; t2[v] = u23.a; t3[v] = u23.b; t4[v] = p[4];
                }
            };
            // the pair's two cells from the fluxes of pass q (t2, t3: the cells' forcing values)
            auto finish_pass = [&](auto qc, const double (&t2)[4], const double (&t3)[4], const double (&f0)[4],
                                   const double (&f1)[4]) {
                constexpr int q = decltype(qc)::value;
                const int t = 64 * q + 2 * lane + 2;  // tile column of the left cell of the pair
                const int i = i0 + 64 * q;
                bool ok = t >= tlo && t < thi;
                if (s == 2) ok = ok && i < nx;
                // All four variables' updates as ONE straight-line block (eight independent FP64 chains, the
                // four shuffles in flight together), then the stores; the rare extra stores (periodic images
                // of the edge columns, state_tmp) sit behind a single branch per pass instead of one per
                // variable, which used to split the block and serialise the variables.
                double2 xv[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const double give = (lane == 0) ? keep[v] : f0[v];
                    const double fr = __shfl_sync(0xffffffffu, give, src_lane);  // flux through the pair's right face
                    keep[v] = f0[v];
                    // the initial state of the cells: always from the state row (in stage 1 these are the taps t2, t3
                    // again -- re-reading them costs four loads in one stage of three, selecting between registers
                    // and loads cost sixteen moves in every stage)
                    const Pair in = lds2(rowS + v * FW + 64 * q + 2);
                    if (HAS_SRC && v == WMOM) {
                        // gravity-wave forcing (source.py:43-50) of the pair's cells; halo cells of T1 / T2
                        // are periodic images, so they take the forcing of the cell they mirror
                        int iw = i % nx;
                        iw += (iw < 0) ? nx : 0;
                        const double2 g = __ldg(reinterpret_cast<const double2*>(a.src_w + (long long)it.k * nx + iw));
                        xv[v] = make_double2(cell_update<false, true>(f0[v], f1[v], in.a, cds, 0.0, 0.0, dts, g.x),
                                             cell_update<false, true>(f1[v], fr, in.b, cds, 0.0, 0.0, dts, g.y));
                    } else {
                        xv[v] = make_double2(cell_update<false, false>(f0[v], f1[v], in.a, cds, 0.0, 0.0, dts, 0.0),
                                             cell_update<false, false>(f1[v], fr, in.b, cds, 0.0, 0.0, dts, 0.0));
                    }
                }
                // stage 1 -> T1, stage 2 -> T2 (column t): shared memory through a 32-bit address and immediate
                // offsets; stage 3 -> HBM
                if (s != 2) {
                    if (ok) {
                        constexpr int o = 8 * (64 * q + 2), vs = 8 * FW;
                        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(tdst32), "n"(o), "d"(xv[0].x), "d"(xv[0].y) : "memory");
                        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(tdst32), "n"(o + vs), "d"(xv[1].x), "d"(xv[1].y) : "memory");
                        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(tdst32), "n"(o + 2 * vs), "d"(xv[2].x), "d"(xv[2].y) : "memory");
                        asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(tdst32), "n"(o + 3 * vs), "d"(xv[3].x), "d"(xv[3].y) : "memory");
                    }
                    return;
                }
                double* const dst = po + 64 * q + 2;
                const long long dvs = a.L.vstride;
                if (ok) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) *reinterpret_cast<double2*>(dst + v * dvs) = xv[v];
                }
                if (s == 2 && ok) {
                    if (a.periodic && (i < SWEEP_HALO || i >= nx - SWEEP_HALO)) {
                        const long long img = (i < SWEEP_HALO) ? nx : -nx;
#pragma unroll
                        for (int v = 0; v < 4; ++v) *reinterpret_cast<double2*>(dst + v * dvs + img) = xv[v];
                        if (i < SWEEP_HALO && i >= nx - SWEEP_HALO) {  // a domain narrower than 12 columns: both images
#pragma unroll
                            for (int v = 0; v < 4; ++v) *reinterpret_cast<double2*>(dst + v * dvs - nx) = xv[v];
                        }
                    }
                    if (WRITE_TMP) {
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            *reinterpret_cast<double2*>(dst + v * dvs + tmp_off) = make_double2(t2[v], t3[v]);
                    }
                }
            };
            auto one_pass = [&](auto qc) {
                constexpr int q = decltype(qc)::value;
                if (q >= nq) return;
                double t0[4], t1[4], t2[4], t3[4], t4[4], f0[4], f1[4];
                load_taps(q, t0, t1, t2, t3, t4);
                xpair_flux<POW_MODE>(t0, t1, t2, t3, t4, bg, a.hv_coeff, f0, f1);
                finish_pass(qc, t2, t3, f0, f1);
            };
            {
                if constexpr (P == 3) one_pass(std::integral_constant<int, 2>{});
                one_pass(std::integral_constant<int, 1>{});
                one_pass(std::integral_constant<int, 0>{});
            }
            __syncwarp();
            src = rowT + s * T::S_ELEMS;
            dts = (s == 0) ? a.dt2 : a.dt3;
            cds = (s == 0) ? a.cd2 : a.cd3;
            tlo += 2;
            thi -= 2;
        }
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// z sweep, transposing variant (sweep_zt): the x-sweep organisation applied along z.
//   The state stays x-fastest in HBM.  Work item = a group of 4 adjacent columns x one z tile of
//   LC = 64P-10 owned rows (+6 halo rows either side).  A CTA of 4 warps owns the item: one TMA
//   box [4 vars][FW rows][4 columns] (32-byte row pieces) lands in a raw buffer; every warp copies
//   ITS column out of it into a z-contiguous shared-memory row (the transposition), and from there
//   on runs exactly the warp-autonomous three-stage pipeline of sweep_x along z: lanes own pairs
//   of vertically adjacent cells, T1 / T2 live in one in-place row, fluxes travel by shuffle.
//   Stage 3 deposits the new state in the (dead) raw buffer, [var][row][4 columns], and the CTA
//   stores it as 32-byte row pieces.  The raw buffer is double-buffered: the next item's box is in
//   flight during the current item.  Solid walls (set_bc_z, bcs.py:92-148): the two halo cells
//   beyond a wall are rebuilt in the warp's row before each stage; rows outside the array arrive
//   as zeros (TMA out-of-bounds fill) and only feed discarded interfaces.
// ------------------------------------------------------------------------------------------
template <int P>
struct ZTSweepTile {
    static constexpr int FW = 64 * P + 4;
    static constexpr int LC = 64 * P - 10;
    static constexpr int TW = FW + 2;
    static constexpr int WARPS = 4;  // = columns per group
    static constexpr int RAW_ELEMS = NVAR * FW * WARPS;
    static constexpr int S_ELEMS = NVAR * FW;
    static constexpr int T_ELEMS = (NVAR * TW + 15) / 16 * 16;
    static constexpr int H_ELEMS = 4 * FW;  // hydrostatic interface profiles of the tile
    static constexpr size_t smem_bytes()
    {
        return (size_t)(2 * RAW_ELEMS + WARPS * (S_ELEMS + T_ELEMS) + H_ELEMS) * sizeof(double) + 32;
    }
};

// Rebuild the two cells beyond a wall in a forcing row (`row` = column 0 of variable 0, plane stride vs).
__device__ __forceinline__ void zt_wall_fix(double* row, int vs, int r0, int nz, int fw, const double* hd, int lane)
{
    if (lane < 8) {
        const int v = lane >> 1, j = lane & 1;
        if (r0 == 0)  // cells -2, -1 are tile columns 4, 5; interior cell 0 is column 6
            row[v * vs + 4 + j] = wall_value(v, row[v * vs + 6], __ldg(hd + HS), __ldg(hd + j));
        const int tt = nz - r0 + SWEEP_HALO;  // tile column of cell nz
        if (tt + 1 < fw)
            row[v * vs + tt + j] = wall_value(v, row[v * vs + tt - 1], __ldg(hd + nz + HS - 1), __ldg(hd + nz + HS + j));
    }
    __syncwarp();
}

#ifndef PMW_ZTSWEEP_MINB
#define PMW_ZTSWEEP_MINB 3
#endif
template <int P, int POW_MODE, bool WRITE_TMP>
__global__ void __launch_bounds__(32 * ZTSweepTile<P>::WARPS, PMW_ZTSWEEP_MINB)
sweep_zt(const __grid_constant__ CUtensorMap tm_box, const SweepArgs a, const int ngroups, const int ntz)
{
    using T = ZTSweepTile<P>;
    constexpr int FW = T::FW, TW = T::TW, LC = T::LC, NW = T::WARPS;
    constexpr int WSTRIDE = T::S_ELEMS + T::T_ELEMS;  // doubles between the rows of consecutive warps
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* const raw = reinterpret_cast<double*>(smem_raw);  // [2][4][FW][NW]
    double* const rows = raw + 2 * T::RAW_ELEMS;              // [NW]{ S[4][FW], T[4][TW] }
    double* const sS = rows + warp * WSTRIDE;
    double* const sT = sS + T::S_ELEMS;
    double* const sH = rows + NW * WSTRIDE;  // [4][FW]: dens, dens_theta, 1/dens_theta, pressure per tile interface
    uint64_t* const bars = reinterpret_cast<uint64_t*>(sH + T::H_ELEMS);
    static_assert((T::RAW_ELEMS * 8) % 128 == 0, "raw boxes stay 128-byte aligned");

    const int nx = a.L.nx, nz = a.L.nz;
    const int nitems = ngroups * ntz;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_box);
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
    }
    for (int e = lane; e < NVAR * 8; e += 32) {  // columns of T no stage writes
        const int v = e >> 3, j = e & 7;
        sT[v * TW + (j < 2 ? j : 64 * P - 2 + j)] = 0.0;
    }
    __syncthreads();
    pdl_wait();
    const unsigned long long pol = l2_policy(1);
    auto request = [&](int n, int buf) {  // thread 0: start the load of item n's box
        const int g = n % ngroups, tz = n / ngroups;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the old box before the TMA write
        mbar_arrive_expect_tx(bars + buf, (uint32_t)(T::RAW_ELEMS * sizeof(double)));
        // map column = array column + 4 = interior column + 6; array row = cell row + 2 (may be negative: zero fill)
        tma_load_3d(raw + buf * T::RAW_ELEMS, &tm_box, NW * g + SWEEP_HALO, tz * LC - SWEEP_HALO + HS, 0, bars + buf, pol);
    };
    if (blockIdx.x < nitems && threadIdx.x == 0) request(blockIdx.x, 0);
    const int src_lane = (lane + 1) & 31;
    const double* const hd = a.hy.dens_cell;
    int buf = 0, tz_loaded = -1;
    unsigned phase = 0;
#pragma unroll 1
    for (int n = blockIdx.x; n < nitems; n += gridDim.x) {
        const int g = n % ngroups, tz = n / ngroups;
        if (threadIdx.x == 0 && n + (int)gridDim.x < nitems) request(n + gridDim.x, buf ^ 1);
        const int r0 = tz * LC;  // first owned cell row
        if (tz != tz_loaded) {   // interface profiles of this z tile: tile column j is interface r0 - 4 + j
            for (int j = threadIdx.x; j < FW; j += 32 * NW) {
                const int kc = min(max(r0 - 4 + j, 0), nz);
                sH[j] = __ldg(a.hy.dens_int + kc);
                sH[FW + j] = __ldg(a.hy.dens_theta_int + kc);
                sH[2 * FW + j] = __ldg(a.hy.inv_dens_theta_int + kc);
                sH[3 * FW + j] = __ldg(a.hy.pressure_int + kc);
            }
            tz_loaded = tz;
        }
        const double* const box = raw + buf * T::RAW_ELEMS;
        mbar_wait(bars + buf, (phase >> buf) & 1);
        phase ^= 1u << buf;
        // transposition: every thread takes whole 32-byte rows of the box and deals the four columns
        // out to the four warps' z-contiguous rows
        for (int e = threadIdx.x; e < NVAR * FW; e += 32 * NW) {
            const double2 c01 = *reinterpret_cast<const double2*>(box + e * NW);
            const double2 c23 = *reinterpret_cast<const double2*>(box + e * NW + 2);
            rows[e] = c01.x;
            rows[WSTRIDE + e] = c01.y;
            rows[2 * WSTRIDE + e] = c23.x;
            rows[3 * WSTRIDE + e] = c23.y;
        }
        __syncthreads();  // (A) rows and profiles complete; the box is dead
        zt_wall_fix(sS, FW, r0, nz, FW, hd, lane);

        const int rem = min(nz - r0, LC);
        double* const rowS = sS + 2 * lane;
        double* const rowT = sT + 2 * lane;
        const double* src = rowS;
        int vs = FW;
        double dts = a.dt1, cds = a.cd1, cgs = a.cg1;
        int tlo = 2, thi = 64 * P;
#pragma unroll 1
        for (int s = 0; s < 3; ++s) {
            const int nq = min(P, (rem + 10 - 2 * s) / 64 + 1);
            double keep[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int q = P - 1; q >= 0; --q) {
                if (q >= nq) continue;
                double t0[4], t1[4], t2[4], t3[4], t4[4], f0[4], f1[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const double* p = src + v * vs + 64 * q;
                    const Pair u01 = lds2(p), u23 = lds2(p + 2);
                    t0[v] = u01.a; t1[v] = u01.b; t2[v] = u23.a; t3[v] = u23.b; t4[v] = p[4];
                }
                if (s == 1) __syncwarp();  // stage 2 overwrites T in place: every lane holds its taps first
                // interface index of tile column j is r0 - 4 + j; the pair's left cell has the same index
                const int j0 = 64 * q + 2 * lane;
                const int k0 = r0 - 4 + j0;
                IfaceBg bg0, bg1;
                {
                    const Pair d = lds2(sH + j0), dt = lds2(sH + FW + j0), idt = lds2(sH + 2 * FW + j0),
                               pr = lds2(sH + 3 * FW + j0);
                    bg0.dens = d.a; bg0.dens_theta = dt.a; bg0.inv_dens_theta = idt.a; bg0.pressure = pr.a;
                    bg1.dens = d.b; bg1.dens_theta = dt.b; bg1.inv_dens_theta = idt.b; bg1.pressure = pr.b;
                }
                const bool w0 = (k0 == 0 || k0 == nz), w1 = (k0 + 1 == 0 || k0 + 1 == nz);
                const bool bad0 = interface_flux_fast<true, POW_MODE>(t0, t1, t2, t3, bg0, a.hv_coeff, w0, f0);
                const bool bad1 = interface_flux_fast<true, POW_MODE>(t1, t2, t3, t4, bg1, a.hv_coeff, w1, f1);
                if (__any_sync(0xffffffffu, bad0 || bad1)) {
                    asm volatile("" ::: "memory");
                    Taps TT;
                    if (bad0) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) { TT.s[0][v] = t0[v]; TT.s[1][v] = t1[v]; TT.s[2][v] = t2[v]; TT.s[3][v] = t3[v]; }
                        const Flux4 gg = interface_flux_slow<true, POW_MODE>(TT, bg0, a.hv_coeff, w0);
#pragma unroll
                        for (int v = 0; v < 4; ++v) f0[v] = gg.f[v];
                    }
                    if (bad1) {
#pragma unroll
                        for (int v = 0; v < 4; ++v) { TT.s[0][v] = t1[v]; TT.s[1][v] = t2[v]; TT.s[2][v] = t3[v]; TT.s[3][v] = t4[v]; }
                        const Flux4 gg = interface_flux_slow<true, POW_MODE>(TT, bg1, a.hv_coeff, w1);
#pragma unroll
                        for (int v = 0; v < 4; ++v) f1[v] = gg.f[v];
                    }
                }
                const int t = j0 + 2;  // tile column of the pair's left cell (cell row k0)
                const bool ok = t >= tlo && t < thi && k0 >= 0 && k0 < nz;
                // stage 1 -> T1 at column t, stage 2 -> T2 at column t+2 (in place), stage 3 -> over S at column
                // t: the lane has read its initial state there and stage 3 reads no other column of S
                double* const dst = (s == 2) ? rowS + 64 * q + 2 : rowT + 64 * q + 2 + 2 * s;
                const int dvs = (s == 2) ? FW : TW;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const double give = (lane == 0) ? keep[v] : f0[v];
                    const double fr = __shfl_sync(0xffffffffu, give, src_lane);  // flux through the pair's top face
                    keep[v] = f0[v];
                    double ia = t2[v], ib = t3[v];
                    if (s != 0) {
                        const Pair in = lds2(rowS + v * FW + 64 * q + 2);
                        ia = in.a; ib = in.b;
                    }
                    double2 x;
                    if (v == WMOM)  // hydrostatic source (interpolate.py:248-250): the cells are taps 2 and 3
                        x = make_double2(cell_update<true, false>(f0[v], f1[v], ia, cds, cgs, t2[DENS], dts, 0.0),
                                         cell_update<true, false>(f1[v], fr, ib, cds, cgs, t3[DENS], dts, 0.0));
                    else
                        x = make_double2(cell_update<false, false>(f0[v], f1[v], ia, cds, cgs, 0.0, dts, 0.0),
                                         cell_update<false, false>(f1[v], fr, ib, cds, cgs, 0.0, dts, 0.0));
                    if (ok) *reinterpret_cast<double2*>(dst + v * dvs) = x;
                }
            }
            __syncwarp();
            if (s < 2) zt_wall_fix(sT + 2 * s, TW, r0, nz, FW, hd, lane);
            src = rowT + 2 * s;
            vs = TW;
            dts = (s == 0) ? a.dt2 : a.dt3;
            cds = (s == 0) ? a.cd2 : a.cd3;
            cgs = (s == 0) ? a.cg2 : a.cg3;
            tlo += 2;
            thi -= 2;
        }
        __syncthreads();  // (B) the four columns of the new state are complete
        if ((int)threadIdx.x < rem) {  // one owned row per thread: four columns = 32 bytes per variable
            const int rr = threadIdx.x, i = NW * g;
            double* const o = a.out + idx(a.L, 0, r0 + rr + HS, i + HS);
            const long long tmp_off = a.tmp - a.out;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const double* c = rows + v * FW + SWEEP_HALO + rr;
                const double2 x01 = make_double2(c[0], c[WSTRIDE]), x23 = make_double2(c[2 * WSTRIDE], c[3 * WSTRIDE]);
                double* ov = o + v * a.L.vstride;
                *reinterpret_cast<double2*>(ov) = x01;
                if (i + 2 < nx) *reinterpret_cast<double2*>(ov + 2) = x23;
                if (a.periodic) {
                    if (i < SWEEP_HALO) *reinterpret_cast<double2*>(ov + nx) = x01;
                    if (i + 2 < SWEEP_HALO) *reinterpret_cast<double2*>(ov + 2 + nx) = x23;
                    if (i >= nx - SWEEP_HALO) *reinterpret_cast<double2*>(ov - nx) = x01;
                    if (i + 2 >= nx - SWEEP_HALO && i + 2 < nx) *reinterpret_cast<double2*>(ov + 2 - nx) = x23;
                }
                if (WRITE_TMP) {  // T2 of the owned cells sits in the T rows, shifted by two columns
                    const double* ct = rows + T::S_ELEMS + v * TW + SWEEP_HALO + 2 + rr;
                    *reinterpret_cast<double2*>(ov + tmp_off) = make_double2(ct[0], ct[WSTRIDE]);
                    if (i + 2 < nx) *reinterpret_cast<double2*>(ov + 2 + tmp_off) = make_double2(ct[2 * WSTRIDE], ct[3 * WSTRIDE]);
                }
            }
        }
        __syncthreads();  // (C) rows free for the next item
        buf ^= 1;
    }
}

// ------------------------------------------------------------------------------------------
// z sweep
// ------------------------------------------------------------------------------------------
constexpr int ZS_COLS = 32;   // columns per strip (one per lane)
constexpr int ZS_RING = 16;   // state rows resident per warp
constexpr int ZS_AHEAD = 6;   // rows requested ahead of the newest row in use
constexpr int ZS_ROW = NVAR * ZS_COLS;  // doubles per ring slot
// state ring | interface-profile ring [16][4] (the entry of interface m rides along with state row m) | mbarriers
constexpr size_t zsweep_smem_bytes() { return (size_t)ZS_RING * (ZS_ROW + 4) * sizeof(double) + ZS_RING * 8; }

// Register window of one stage: the forcing cells k-2 .. k+1 of interface k, [slot][variable].
// The generic path keeps tap t in slot t and shifts; the steady-state path rotates instead (tap t
// of a stage whose window is at rotation R lives in slot (R+t)&3; four iterations, unrolled, bring
// the rotation back to 0), so that no register is ever moved.
template <int POW_MODE>
struct ZStage {
    double W[4][4];
    double fprev[4];  // flux through interface k-1

    // Generic step (warp-uniform branches for walls): evaluate interface k from slots 0..3 and
    // finalise cell k-1 = init + dt*tendency.
    // HAS_SRC: `src` is the gravity-wave forcing (source.py:43-50) of cell k-1.
    template <bool HAS_SRC = false>
    __device__ __forceinline__ void step(const SweepArgs& a, int k, const IfaceBg& bg, double dt_stage, double cd,
                                         double cg, const double (&init)[4], double (&cell)[4], double src = 0.0)
    {
        const int nz = a.L.nz;
        const double* hd = a.hy.dens_cell;
        if (k == 0) {  // set_bc_z, bottom rows (bcs.py:92-148) from interior row 0 = tap 2
            const double h2 = __ldg(hd + HS), h0 = __ldg(hd), h1 = __ldg(hd + 1);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                W[0][v] = wall_value(v, W[2][v], h2, h0);
                W[1][v] = wall_value(v, W[2][v], h2, h1);
            }
        }
        if (k == nz - 1) {  // top halo row nz+2 from interior row nz-1 = tap 2
            const double hi = __ldg(hd + nz + HS - 1), h = __ldg(hd + nz + HS);
#pragma unroll
            for (int v = 0; v < 4; ++v) W[3][v] = wall_value(v, W[2][v], hi, h);
        }
        if (k == nz) {  // interior row nz-1 = tap 1
            const double hi = __ldg(hd + nz + HS - 1), h2 = __ldg(hd + nz + HS), h3 = __ldg(hd + nz + HS + 1);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                W[2][v] = wall_value(v, W[1][v], hi, h2);
                W[3][v] = wall_value(v, W[1][v], hi, h3);
            }
        }
        const bool wall = (k == 0 || k == nz);
        double f[4];
        interface_flux<true, POW_MODE>(W[0], W[1], W[2], W[3], bg, a.hv_coeff, wall, f);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v == WMOM)  // hydrostatic source on the cell's rho' (interpolate.py:248-250)
                cell[v] = cell_update<true, HAS_SRC>(fprev[v], f[v], init[v], cd, cg, W[1][DENS], dt_stage, src);
            else
                cell[v] = cell_update<false, false>(fprev[v], f[v], init[v], cd, cg, 0.0, dt_stage, 0.0);
            fprev[v] = f[v];
        }
    }
    __device__ __forceinline__ void push(const double (&row)[4])
    {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            W[0][v] = W[1][v]; W[1][v] = W[2][v]; W[2][v] = W[3][v]; W[3][v] = row[v];
        }
    }

    // Steady state, rotation R0 (tap t in slot (R0+t)&3): flux of interior interface k.
    template <int R0>
    __device__ __forceinline__ bool flux_fast(const SweepArgs& a, const IfaceBg& bg, double (&f)[4])
    {
        return interface_flux_fast<true, POW_MODE>(W[R0 & 3], W[(R0 + 1) & 3], W[(R0 + 2) & 3], W[(R0 + 3) & 3], bg,
                                                   a.hv_coeff, false, f);
    }
    template <int R0>
    __device__ __forceinline__ void flux_slow(const SweepArgs& a, const IfaceBg& bg, double (&f)[4])
    {
        Taps T;
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int v = 0; v < 4; ++v) T.s[t][v] = W[(R0 + t) & 3][v];
        const Flux4 g = interface_flux_slow<true, POW_MODE>(T, bg, a.hv_coeff, false);
#pragma unroll
        for (int v = 0; v < 4; ++v) f[v] = g.f[v];
    }
    // cell k-1 (tap 1) from the fluxes through its two faces
    template <int R0, bool HAS_SRC = false>
    __device__ __forceinline__ void finish(const SweepArgs& a, const double (&f)[4], double dt_stage, double cd,
                                           double cg, const double (&init)[4], double (&cell)[4], double src = 0.0)
    {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v == WMOM)
                cell[v] = cell_update<true, HAS_SRC>(fprev[v], f[v], init[v], cd, cg, W[(R0 + 1) & 3][DENS], dt_stage, src);
            else
                cell[v] = cell_update<false, false>(fprev[v], f[v], init[v], cd, cg, 0.0, dt_stage, 0.0);
            fprev[v] = f[v];
        }
    }
};

struct ZStream {  // per-warp constants of the state-row stream
    double* ring;
    double* bgring;  // [ZS_RING][4]: {dens, dens_theta, 1/dens_theta, pressure} of interface m in the slot of row m
    uint64_t* bars;
    const CUtensorMap* tm;
    const double* int_pack;  // Hydro::int_pack
    int f0, last_cell, c0, lane;
    unsigned long long pol;
    // lane 0: start the load of state cell row m, and of the hydrostatic profiles of interface m with it (a 32-byte
    // bulk copy onto the same mbarrier: the steady iterations then read their profiles from shared memory instead
    // of through L1, whose misses -- one line per 16 interfaces and profile -- were 12 % of the stall samples)
    __device__ __forceinline__ void request(int m) const
    {
        if (m <= last_cell) {
            const int s = (m - f0) & (ZS_RING - 1);
            mbar_arrive_expect_tx(bars + s, (uint32_t)((ZS_ROW + 4) * sizeof(double)));
            tma_load_3d(ring + s * ZS_ROW, tm, c0 + HS + 4, m + HS, 0, bars + s, pol);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
                         ::"r"(smem_u32(bgring + 4 * s)), "l"(int_pack + 4 * (long long)m), "r"(smem_u32(bars + s))
                         : "memory");
        }
    }
    __device__ __forceinline__ IfaceBg bg(int m) const  // profiles of interface m (row m has been waited for)
    {
        const double* p = bgring + 4 * ((m - f0) & (ZS_RING - 1));
        const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        IfaceBg r;
        r.dens = a.x; r.dens_theta = a.y; r.inv_dens_theta = b.x; r.pressure = b.y;
        return r;
    }
    __device__ __forceinline__ const double* row(int m) const
    {
        return ring + ((m - f0) & (ZS_RING - 1)) * ZS_ROW + lane;
    }
    __device__ __forceinline__ void wait(int m) const
    {
        mbar_wait(bars + ((m - f0) & (ZS_RING - 1)), ((m - f0) >> 4) & 1);
    }
};

// One steady-state iteration at window rotation R: stage 1 at interface j, stage 2 at j-3, stage 3
// at j-6; all interior, all cells valid.  Straight-line code: the three evaluations interleave.
template <int R, int POW_MODE, bool WRITE_TMP, bool HAS_SRC = false>
__device__ __forceinline__ void zsweep_steady(const SweepArgs& a, const ZStream& zs, ZStage<POW_MODE>& s1,
                                              ZStage<POW_MODE>& s2, ZStage<POW_MODE>& s3, int j, double* pout,
                                              double* ptmp, bool col_ok, bool img_r, bool img_l,
                                              const double* psrc = nullptr)
{
    __syncwarp();  // every lane is done with the rows of the previous iteration
    if (zs.lane == 0) zs.request(j + 1 + ZS_AHEAD);
    zs.wait(j + 1);
    const double* top = zs.row(j + 1);
    double f1[4], f2[4], f3[4], c1[4], c2[4], c3[4];
#pragma unroll
    for (int v = 0; v < 4; ++v)
        s1.W[R & 3][v] = top[v * ZS_COLS];  // newest state row replaces the oldest: taps now start at slot R+1
    const IfaceBg bg1 = zs.bg(j), bg2 = zs.bg(j - 3), bg3 = zs.bg(j - 6);
    const bool bad3 = s3.template flux_fast<R>(a, bg3, f3);
    const bool bad2 = s2.template flux_fast<R>(a, bg2, f2);
    const bool bad1 = s1.template flux_fast<R + 1>(a, bg1, f1);
    if (__any_sync(0xffffffffu, bad1 || bad2 || bad3)) {
        asm volatile("" ::: "memory");  // keep the argument copies of the cold path inside the branch
        if (bad3) s3.template flux_slow<R>(a, bg3, f3);
        if (bad2) s2.template flux_slow<R>(a, bg2, f2);
        if (bad1) s1.template flux_slow<R + 1>(a, bg1, f1);
    }
    double in2[4], in3[4];  // initial state of the cells stages 2 and 3 finish (read late: short live ranges)
    {
        const double* r2 = zs.row(j - 4);
        const double* r3 = zs.row(j - 7);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            in2[v] = r2[v * ZS_COLS];
            in3[v] = r3[v * ZS_COLS];
        }
    }
    if (WRITE_TMP && col_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) ptmp[v * a.L.vstride] = s3.W[(R + 1) & 3][v];  // T2 of the cell
    }
    double g1 = 0.0, g2 = 0.0, g3 = 0.0;
    if (HAS_SRC) {  // psrc: this lane's column of source row j-7 (the cell stage 3 finishes)
        g3 = __ldg(psrc);
        g2 = __ldg(psrc + 3 * (long long)a.L.nx);  // stage 2 finishes cell j-4
        g1 = __ldg(psrc + 6 * (long long)a.L.nx);  // stage 1 finishes cell j-1
    }
    s3.template finish<R, HAS_SRC>(a, f3, a.dt3, a.cd3, a.cg3, in3, c3, g3);
    s2.template finish<R, HAS_SRC>(a, f2, a.dt2, a.cd2, a.cg2, in2, c2, g2);
    s1.template finish<R + 1, HAS_SRC>(a, f1, a.dt1, a.cd1, a.cg1, s1.W[(R + 2) & 3], c1, g1);
    if (col_ok) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            double* q = pout + v * a.L.vstride;
            *q = c3[v];
            if (img_r) q[a.L.nx] = c3[v];
            if (img_l) q[-a.L.nx] = c3[v];
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {  // the new cells replace the oldest taps
        s3.W[R & 3][v] = c2[v];
        s2.W[R & 3][v] = c1[v];
    }
}

}  // namespace pmw
namespace pmw {

#ifndef PMW_ZSWEEP_MINB
#define PMW_ZSWEEP_MINB 8  // resident warps (= CTAs) per SM the register allocation aims at: 8 -> 255 registers
#endif
template <int POW_MODE, bool WRITE_TMP, bool HAS_SRC = false>
#ifdef PMW_ZSWEEP_MAXNREG
__global__ void __maxnreg__(PMW_ZSWEEP_MAXNREG)
#else
__global__ void __launch_bounds__(32)
#endif
sweep_z(const __grid_constant__ CUtensorMap tm_row, const SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ZStream zs;
    zs.ring = reinterpret_cast<double*>(smem_raw);
    zs.bgring = zs.ring + ZS_RING * ZS_ROW;
    zs.bars = reinterpret_cast<uint64_t*>(zs.bgring + ZS_RING * 4);
    zs.int_pack = a.hy.int_pack;
    zs.tm = &tm_row;

    const int lane = threadIdx.x;
    const int nx = a.L.nx, nz = a.L.nz;
    const int c0 = (blockIdx.x + a.tile_x0) * ZS_COLS;
    const int i = c0 + lane;
    const bool col_ok = i < nx;
    const int lo3 = a.row0 + blockIdx.y * a.lz, hi3 = min(lo3 + a.lz, a.row1);
    const int lo2 = max(lo3 - 2, 0), hi2 = min(hi3 + 2, nz);
    const int lo1 = max(lo3 - 4, 0), hi1 = min(hi3 + 4, nz);
    zs.f0 = lo1 - 2;          // first state cell row of the stream (array row f0 + 2)
    zs.last_cell = hi1 + 1;   // last one
    zs.c0 = c0;
    zs.lane = lane;
    pdl_launch_dependents();
    if (lane == 0) {
        tma_prefetch_desc(&tm_row);
        for (int s = 0; s < ZS_RING; ++s) mbar_init(zs.bars + s, 1);
    }
    __syncwarp();
    pdl_wait();
    zs.pol = l2_policy(1);
    if (lane == 0)
        for (int m = zs.f0; m <= zs.f0 + 2 + ZS_AHEAD; ++m) zs.request(m);

    ZStage<POW_MODE> s1, s2, s3;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int v = 0; v < 4; ++v) s1.W[t][v] = s2.W[t][v] = s3.W[t][v] = 0.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) s1.fprev[v] = s2.fprev[v] = s3.fprev[v] = 0.0;
    for (int m = zs.f0; m <= zs.f0 + 2; ++m) {  // the first three taps of stage 1
        zs.wait(m);
        double r[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) r[v] = zs.row(m)[v * ZS_COLS];
        s1.push(r);
    }
    double* const pout0 = a.out + idx(a.L, 0, HS, min(i, nx - 1) + HS);
    double* const ptmp0 = a.tmp + idx(a.L, 0, HS, min(i, nx - 1) + HS);
    const bool img_r = a.periodic && i < SWEEP_HALO, img_l = a.periodic && i >= nx - SWEEP_HALO;

    // gravity-wave forcing of cell row m in this lane's column (rows beyond the domain: the cell is discarded)
    auto zsrc = [&](int m) -> double {
        if (!HAS_SRC) return 0.0;
        return __ldg(a.src_w + (long long)min(max(m, 0), nz - 1) * nx + min(i, nx - 1));
    };
    // steady iterations (all three stages active, every cell valid, no wall): js <= j <= je
    const int js = max(lo3 + 7, 7), je = min(hi1, nz - 2);
    int j = lo1;
    while (j <= hi3 + 6) {
        if (j >= js && j + 3 <= je) {
            double* po = pout0 + (long long)(j - 7) * a.L.pitch;
            double* pt = ptmp0 + (long long)(j - 7) * a.L.pitch;
            const int p = a.L.pitch;
            if (HAS_SRC) {
                const double* ps = a.src_w + (long long)(j - 7) * nx + min(i, nx - 1);
                zsweep_steady<0, POW_MODE, WRITE_TMP, true>(a, zs, s1, s2, s3, j, po, pt, col_ok, img_r, img_l, ps);
                zsweep_steady<1, POW_MODE, WRITE_TMP, true>(a, zs, s1, s2, s3, j + 1, po + p, pt + p, col_ok, img_r, img_l, ps + nx);
                zsweep_steady<2, POW_MODE, WRITE_TMP, true>(a, zs, s1, s2, s3, j + 2, po + 2 * p, pt + 2 * p, col_ok, img_r, img_l, ps + 2 * nx);
                zsweep_steady<3, POW_MODE, WRITE_TMP, true>(a, zs, s1, s2, s3, j + 3, po + 3 * p, pt + 3 * p, col_ok, img_r, img_l, ps + 3 * nx);
                j += 4;
                continue;
            }
            zsweep_steady<0, POW_MODE, WRITE_TMP>(a, zs, s1, s2, s3, j, po, pt, col_ok, img_r, img_l);
            zsweep_steady<1, POW_MODE, WRITE_TMP>(a, zs, s1, s2, s3, j + 1, po + p, pt + p, col_ok, img_r, img_l);
            zsweep_steady<2, POW_MODE, WRITE_TMP>(a, zs, s1, s2, s3, j + 2, po + 2 * p, pt + 2 * p, col_ok, img_r, img_l);
            zsweep_steady<3, POW_MODE, WRITE_TMP>(a, zs, s1, s2, s3, j + 3, po + 3 * p, pt + 3 * p, col_ok, img_r, img_l);
            j += 4;
            continue;
        }
        // generic iteration: segment start / end and walls
        __syncwarp();
        if (lane == 0) zs.request(j + 1 + ZS_AHEAD);
        double top[4], in2[4], in3[4];
        double c1[4] = {0.0, 0.0, 0.0, 0.0}, c2[4] = {0.0, 0.0, 0.0, 0.0}, c3[4];
        {
            const int m = j + 1;  // newest state row: tap 3 of interface j
            if (m <= zs.last_cell) {
                zs.wait(m);
#pragma unroll
                for (int v = 0; v < 4; ++v) top[v] = zs.row(m)[v * ZS_COLS];
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) top[v] = 0.0;
            }
            s1.push(top);
        }
        const int k1 = j, k2 = j - 3, k3 = j - 6;
        if (k3 >= lo3 && k3 <= hi3) {
#pragma unroll
            for (int v = 0; v < 4; ++v) in3[v] = (k3 > lo3) ? zs.row(k3 - 1)[v * ZS_COLS] : 0.0;
            s3.template step<HAS_SRC>(a, k3, zs.bg(k3), a.dt3, a.cd3, a.cg3, in3, c3, zsrc(k3 - 1));
            if (k3 > lo3 && col_ok) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const long long o = v * a.L.vstride + (long long)(k3 - 1) * a.L.pitch;
                    pout0[o] = c3[v];
                    if (img_r) pout0[o + nx] = c3[v];
                    if (img_l) pout0[o - nx] = c3[v];
                    if (WRITE_TMP) ptmp0[o] = s3.W[1][v];
                }
            }
        }
        if (k2 >= lo2 && k2 <= hi2) {
#pragma unroll
            for (int v = 0; v < 4; ++v) in2[v] = (k2 > lo2) ? zs.row(k2 - 1)[v * ZS_COLS] : 0.0;
            s2.template step<HAS_SRC>(a, k2, zs.bg(k2), a.dt2, a.cd2, a.cg2, in2, c2, zsrc(k2 - 1));
        }
        if (k1 <= hi1) s1.template step<HAS_SRC>(a, k1, zs.bg(k1), a.dt1, a.cd1, a.cg1, s1.W[1], c1, zsrc(k1 - 1));
        s3.push(c2);
        s2.push(c1);
        ++j;
    }
}

}  // namespace pmw
