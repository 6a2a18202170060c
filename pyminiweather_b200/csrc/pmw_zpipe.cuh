// z sweep, stage-pipelined across the warps of a CTA (sweep_z3): the production z sweep.
//
// Same mathematics as sweep_z (pmw_sweep.cuh): the three low-storage RK stages of a z sweep
// (step.py:112-141) in one kernel, a strip of 32 columns (one per lane) streamed upwards through a
// segment of lz rows, 4-row stencil windows in registers, solid-wall halo rows (set_bc_z,
// bcs.py:92-148) rebuilt in the windows.  What differs is who holds the windows.  sweep_z keeps the
// windows of all three stages in ONE warp (232 registers: two warps per scheduler partition, and a
// kernel bound by the latency of its own FP64 chains).  Here a CTA is three warps and warp s runs
// stage s+1 only:
//
//     warp 0:  S  rows (TMA ring, 32 slots, loaded in groups of four rows)  ->  T1 rows   (shared-memory ring, 4 slots)
//     warp 1:  T1 rows                       ->  T2 rows   (shared-memory ring, 4 slots)
//     warp 2:  T2 rows                       ->  S' rows   (HBM)
//
// One window per thread is ~100 registers, so five CTAs = 15 warps are resident per SM instead of 8,
// and the segments are longer (one wave of 5 x 148 CTAs: 94 rows at 2048x1024 instead of 57), which
// cuts the recomputed rows from 25 % to 13 %.  The warps run in lockstep: iteration `it` of warp s
// evaluates interface lo1 - 3 + it - 4 s, and the CTA meets at a barrier every second iteration.
// A T row is consumed two iterations (= one barrier) after it was produced and its slot is rewritten
// two iterations after it was consumed, so four slots per ring suffice; the state ring keeps a row
// until stage 3 has read it as the initial state of its cell (stages 2 and 3 read `init` from the
// ring).  The hydrostatic interface profiles ride along with the state rows (bulk copies of the
// packed table Hydro::int_pack onto the same mbarriers).  Every interface goes through interface_flux_fast and
// every cell through cell_update with the operands the stage-by-stage kernels use: bit-identical
// results.
#pragma once
#include "pmw_sweep.cuh"

namespace pmw {

constexpr int Z3_COLS = 32;    // columns per strip (one per lane)
constexpr int Z3_SRING = 32;   // state rows resident per CTA
constexpr int Z3_TRING = 4;    // T1 / T2 rows resident per CTA
constexpr int Z3_LAG = 4;      // iterations between consecutive stages (a multiple of the unrolled block)
constexpr int Z3_AHEAD = 12;   // state rows requested ahead of the newest row in use (<= 16, see below)
constexpr int Z3_ROW = NVAR * Z3_COLS;  // doubles per ring row
// Ring layout = the TMA box [4 variables][4 rows][32 columns]: rows live in groups of four, row r of a group
// at r * Z3_RS, variable v at v * Z3_VS; the T rings are one such group each.
constexpr int Z3_RS = Z3_COLS, Z3_VS = 4 * Z3_COLS, Z3_GROUP = 4 * Z3_ROW;
__host__ __device__ constexpr int z3_slot(int s) { return (s >> 2) * Z3_GROUP + (s & 3) * Z3_RS; }  // doubles
constexpr int Z3_WARPS = 3;
#ifndef PMW_Z3_MINB
#define PMW_Z3_MINB 4
#endif
// shared memory: state ring | T1 | T2 | interface-profile ring [32][4] | one mbarrier per group of four state rows
constexpr int Z3_OFF_T1 = Z3_SRING * Z3_ROW, Z3_OFF_T2 = Z3_OFF_T1 + Z3_TRING * Z3_ROW,
              Z3_OFF_BG = Z3_OFF_T2 + Z3_TRING * Z3_ROW, Z3_OFF_BAR = Z3_OFF_BG + Z3_SRING * 4;
constexpr size_t z3_smem_bytes() { return (size_t)(Z3_OFF_BAR + Z3_SRING / 4) * sizeof(double); }
// State-ring residency: at the start of block it0 warp 0 requests rows f0+it0+AHEAD .. +3 into the slots of rows
// 32 below, the last of which stage 3 read (as the initial state of a cell) at iteration it0 + AHEAD - 19; that
// must lie before the barrier that ends the previous block: AHEAD <= 18, and a multiple of 4.
static_assert(Z3_AHEAD <= 16 && Z3_AHEAD % 4 == 0, "state ring too small for this look-ahead");

// The three warps meet here: a named barrier with an explicit thread count.
// `tok` orders the ring loads below (plain asm, free to be scheduled) after the barrier they follow.
__device__ __forceinline__ void z3_barrier(uint32_t& tok) { asm volatile("barrier.sync 1, 96;" : "+r"(tok) : : "memory"); }
// Ring accesses by 32-bit shared-window address (+ a compile-time byte offset): the compiler neither
// re-derives the window base from a generic pointer nor carries 64-bit addresses for them.
template <int OFF>
__device__ __forceinline__ double z3_lds(uint32_t addr, uint32_t tok)
{
    double x;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(x) : "r"(addr), "n"(OFF), "r"(tok));
    return x;
}
template <int OFF>
__device__ __forceinline__ void z3_sts(uint32_t addr, double x)
{
    asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "d"(x) : "memory");
}
template <int OFF>
__device__ __forceinline__ IfaceBg z3_lds_bg(uint32_t addr, uint32_t tok)  // one ring entry {dens, dens_theta, 1/dens_theta, pressure}
{
    IfaceBg bg;
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(bg.dens), "=d"(bg.dens_theta) : "r"(addr), "n"(OFF), "r"(tok));
    asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(bg.inv_dens_theta), "=d"(bg.pressure) : "r"(addr), "n"(OFF + 16), "r"(tok));
    return bg;
}

template <int POW_MODE, bool WRITE_TMP, bool HAS_SRC = false>
#ifdef PMW_Z3_MAXNREG
__global__ void __maxnreg__(PMW_Z3_MAXNREG)
#else
__global__ void __launch_bounds__(32 * Z3_WARPS, PMW_Z3_MINB)
#endif
sweep_z3(const __grid_constant__ CUtensorMap tm_rows, const SweepArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Z3_OFF_BAR);
    // ONE instruction stream for the three roles (`role` is warp-uniform run-time data): the hot loop is
    // ~0.5 k instructions that all 15 warps of an SM share in the instruction caches.  (Three specialised
    // copies, one per role, lost 19 % of the issue slots to instruction fetch: profiles/r2c.)
    const int role = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = (blockIdx.x + a.tile_x0) * Z3_COLS;
    const int nx = a.L.nx, nz = a.L.nz;
    const int lo3 = blockIdx.y * a.lz, hi3 = min(lo3 + a.lz, nz);
    const int lo2 = max(lo3 - 2, 0), hi2 = min(hi3 + 2, nz);
    const int lo1 = max(lo3 - 4, 0), hi1 = min(hi3 + 4, nz);
    const int f0 = lo1 - 2, last = hi1 + 1;  // state cell rows the CTA streams (array rows f0+2 .. last+2)
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_rows);
        for (int s = 0; s < Z3_SRING / 4; ++s) mbar_init(bars + s, 1);
    }
    // this role's cells are lo .. hi-1, its interfaces lo .. hi; the forcing rows it may pull are srclo .. srchi
    const int lo = role == 0 ? lo1 : (role == 1 ? lo2 : lo3);
    const int hi = role == 0 ? hi1 : (role == 1 ? hi2 : hi3);
    const int srclo = role == 0 ? f0 : (role == 1 ? lo1 : lo2);
    const int srchi = role == 0 ? last : (role == 1 ? hi1 - 1 : hi2 - 1);
    const double dts = role == 0 ? a.dt1 : (role == 1 ? a.dt2 : a.dt3);
    const double cds = role == 0 ? a.cd1 : (role == 1 ? a.cd2 : a.cd3);
    const double cgs = role == 0 ? a.cg1 : (role == 1 ? a.cg2 : a.cg3);
    const int nit = (hi3 - lo1 + 3 + 2 * Z3_LAG + 1 + 3) & ~3;  // iterations (stage 3's last interface), whole blocks
    const int jbase = lo1 - 3 - Z3_LAG * role;                  // interface of iteration 0
    // 32-bit shared-window addresses of this lane's column in the rings
    const uint32_t sm0 = smem_u32(smem);
    const uint32_t sring = sm0 + 8u * lane;
    const uint32_t tsrc = sm0 + 8u * ((role == 2 ? Z3_OFF_T2 : Z3_OFF_T1) + lane);  // ring this role reads
    const uint32_t tdst = sm0 + 8u * ((role == 1 ? Z3_OFF_T2 : Z3_OFF_T1) + lane);  // ring it writes
    const uint32_t bgring = sm0 + 8u * Z3_OFF_BG;

    const int i = c0 + lane;
    const int icl = min(i, nx - 1);
    const bool to_hbm = role == 2;
    const bool st_ok = i < nx;
    const bool img_r = a.periodic && i < SWEEP_HALO, img_l = a.periodic && i >= nx - SWEEP_HALO && i < nx;
    const bool img_any = img_r || img_l;
    double* const dhbm = a.out + idx(a.L, 0, HS, icl + HS);
    const long long tmp_off = a.tmp - a.out;
    const long long vst = a.L.vstride;
    const int pitch = a.L.pitch;

    __syncthreads();
    pdl_wait();  // everything below reads state produced by the previous kernel
    const unsigned long long pol = l2_policy(1);
    // lane 0 of warp 0: start the load of the four state cell rows m .. m+3 and of the profiles of interfaces m .. m+3
    auto request = [&](int m) {
        if (m <= last) {
            const int s = (m - f0) & (Z3_SRING - 1);
            mbar_arrive_expect_tx(bars + (s >> 2), (uint32_t)((4 * Z3_ROW + 16) * sizeof(double)));
            tma_load_3d(smem + z3_slot(s), &tm_rows, c0 + HS + 4, m + HS, 0, bars + (s >> 2), pol);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                         ::"r"(sm0 + 8u * (Z3_OFF_BG + 4 * s)), "l"(a.hy.int_pack + 4 * (long long)m), "r"(smem_u32(bars + (s >> 2)))
                         : "memory");
        }
    };
    auto zsrc = [&](int m) -> double {  // gravity-wave forcing of cell row m in this lane's column
        if (!HAS_SRC) return 0.0;
        return __ldg(a.src_w + (long long)min(max(m, 0), nz - 1) * nx + icl);
    };
    if (role == 0 && lane == 0)
        for (int m = f0; m < f0 + Z3_AHEAD; m += 4) request(m);

    uint32_t tok = 0;
    volatile double wsave[16];  // cold path only (see the steady block)
    ZStage<POW_MODE> st;
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int v = 0; v < 4; ++v) st.W[t][v] = 0.0;
#pragma unroll
    for (int v = 0; v < 4; ++v) st.fprev[v] = 0.0;

    // steady blocks of four iterations: every interface interior, every pulled row valid, every cell stored
    const int js = max(lo + 1, 1), je = min(hi, nz - 2);
#pragma unroll 1
    for (int it0 = 0; it0 < nit; it0 += 4) {
        const int j0 = jbase + it0;
        if (role == 0) {  // the group of four state rows this block pulls; request the group Z3_AHEAD rows further
            if (lane == 0) request(f0 + it0 + Z3_AHEAD);
            if (f0 + it0 <= last) mbar_wait(bars + ((it0 & (Z3_SRING - 1)) >> 2), ((unsigned)it0 >> 5) & 1u);
            asm volatile("" : "+r"(tok));  // the ring loads below stay behind the wait
        }
        int r0 = 0;          // first sub-iteration the generic loop below has to run
        bool pulled = false;  // ... and whether its forcing row is already in the window
        if (j0 >= js && j0 + 3 <= je) {
            // Ring slots of this block.  State ring, slot of cell row m = (m - f0) & 31: the rows stage 1 pulls are
            // sb .. sb+3 (one TMA group); the init rows j0-1 .. j0+2 of any role are sb-2, sb-1, sb, sb+1 and the
            // profile entries of interfaces j0 .. j0+3 are sb-1 .. sb+2, with sb = (it0 - 4 role) & 31.  T rings,
            // slot of row m = (m - lo1) & 3: pulled rows 2, 3, 0, 1; stored rows 0, 1, 2, 3.
            const int sb = (it0 - Z3_LAG * role) & (Z3_SRING - 1);
            const uint32_t pI2 = sring + 8u * z3_slot(sb);                            // init rows of R = 2, 3
            const uint32_t pI0 = sring + 8u * z3_slot((sb - 2) & (Z3_SRING - 1));     // init rows of R = 0, 1
            const uint32_t pA = role == 0 ? pI2 : tsrc + 8u * 2 * Z3_RS;              // pulled rows of R = 0, 1
            const uint32_t pB = role == 0 ? pI2 + 8u * 2 * Z3_RS : tsrc;              // pulled rows of R = 2, 3
            const uint32_t pG0 = bgring + 32u * ((sb - 1) & (Z3_SRING - 1));          // profiles of R = 0
            const uint32_t pG1 = bgring + 32u * sb;                                   // profiles of R = 1, 2, 3
            double* const dst = dhbm + (long long)(j0 - 1) * pitch;
            const double* const ps = HAS_SRC ? a.src_w + (long long)(j0 - 1) * nx + icl : nullptr;
#define PMW_Z3_SUB(R)                                                                                              \
    do {                                                                                                           \
        _Pragma("unroll") for (int v = 0; v < 4; ++v) st.W[(R) & 3][v] =                                           \
            (R) < 2 ? z3_lds<8 * ((R) * Z3_RS + 0)>(pA + 8u * v * Z3_VS, tok) : z3_lds<8 * (((R) & 1) * Z3_RS)>(pB + 8u * v * Z3_VS, tok); \
        const IfaceBg bg = (R) == 0 ? z3_lds_bg<0>(pG0, tok) : z3_lds_bg<32 * (((R) + 3) & 3)>(pG1, tok);                     \
        double f[4], c[4], in[4];                                                                                  \
        const bool bad = st.template flux_fast<(R) + 1>(a, bg, f);                                                 \
        if (__any_sync(0xffffffffu, bad)) { /* |e| > 1/8 somewhere: this iteration and the rest of the block run    \
                                               the generic path (pow).  The window travels there in canonical      \
                                               order through local memory, so that the join constrains no         \
                                               register of the hot path (through registers the rotation below      \
                                               degenerated into 27 moves per iteration) */                         \
            asm volatile("" ::: "memory"); /* keep the copies below inside the cold branch */                     \
            _Pragma("unroll") for (int t = 0; t < 4; ++t)                                                          \
                _Pragma("unroll") for (int v = 0; v < 4; ++v) wsave[4 * t + v] = st.W[((R) + 1 + t) & 3][v];       \
            r0 = (R);                                                                                              \
            pulled = true;                                                                                         \
            goto generic;                                                                                          \
        }                                                                                                          \
        _Pragma("unroll") for (int v = 0; v < 4; ++v) in[v] =                                                      \
            (R) < 2 ? z3_lds<8 * ((R) * Z3_RS)>(pI0 + 8u * v * Z3_VS, tok) : z3_lds<8 * (((R) & 1) * Z3_RS)>(pI2 + 8u * v * Z3_VS, tok); \
        const double g = HAS_SRC ? __ldg(ps + (long long)(R) * nx) : 0.0;                                          \
        st.template finish<(R) + 1, HAS_SRC>(a, f, dts, cds, cgs, in, c, g);                                       \
        if (to_hbm) {                                                                                              \
            double* q = dst + (R) * pitch;                                                                         \
            if (st_ok) {                                                                                           \
                _Pragma("unroll") for (int v = 0; v < 4; ++v) q[v * vst] = c[v];                                   \
                if (WRITE_TMP) {                                                                                   \
                    _Pragma("unroll") for (int v = 0; v < 4; ++v) q[v * vst + tmp_off] = st.W[((R) + 2) & 3][v];   \
                }                                                                                                  \
            }                                                                                                      \
            if (img_any) {                                                                                         \
                _Pragma("unroll") for (int v = 0; v < 4; ++v) q[v * vst + (img_r ? nx : -nx)] = c[v];              \
            }                                                                                                      \
        } else {                                                                                                   \
            z3_sts<8 * ((R) * Z3_RS + 0 * Z3_VS)>(tdst, c[0]);                                                     \
            z3_sts<8 * ((R) * Z3_RS + 1 * Z3_VS)>(tdst, c[1]);                                                     \
            z3_sts<8 * ((R) * Z3_RS + 2 * Z3_VS)>(tdst, c[2]);                                                     \
            z3_sts<8 * ((R) * Z3_RS + 3 * Z3_VS)>(tdst, c[3]);                                                     \
        }                                                                                                          \
    } while (0)
            PMW_Z3_SUB(0);
            PMW_Z3_SUB(1);
            z3_barrier(tok);
            PMW_Z3_SUB(2);
            PMW_Z3_SUB(3);
            z3_barrier(tok);
#undef PMW_Z3_SUB
            continue;
        }
    generic:
        // generic iterations: pipeline fill / drain, segment ends, walls, and |e| > 1/8 (interface_flux: pow)
        if (pulled) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int v = 0; v < 4; ++v) st.W[t][v] = wsave[4 * t + v];
        }
#pragma unroll 1
        for (int r = r0; r < 4; ++r) {
            const int it = it0 + r;
            const int j = jbase + it, m = j + 1;
            if (!pulled) {
                double top[4] = {0.0, 0.0, 0.0, 0.0};
                if (m >= srclo && m <= srchi) {
                    const uint32_t prow = role == 0 ? sring + 8u * z3_slot((m - f0) & (Z3_SRING - 1))
                                                    : tsrc + 8u * ((m - lo1) & (Z3_TRING - 1)) * Z3_RS;
#pragma unroll
                    for (int v = 0; v < 4; ++v) top[v] = z3_lds<0>(prow + 8u * v * Z3_VS, tok);
                }
                st.push(top);
            }
            pulled = false;
            if (j >= lo && j <= hi) {
                double in[4] = {0.0, 0.0, 0.0, 0.0}, c[4];
                if (j > lo) {
                    const uint32_t pin = sring + 8u * z3_slot((j - 1 - f0) & (Z3_SRING - 1));
#pragma unroll
                    for (int v = 0; v < 4; ++v) in[v] = z3_lds<0>(pin + 8u * v * Z3_VS, tok);
                }
                st.template step<HAS_SRC>(a, j, z3_lds_bg<0>(bgring + 32u * ((j - f0) & (Z3_SRING - 1)), tok), dts, cds, cgs, in, c,
                                           zsrc(j - 1));
                if (j > lo) {
                    if (to_hbm) {
                        double* q = dhbm + (long long)(j - 1) * pitch;
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            if (st_ok) q[v * vst] = c[v];
                            if (img_r) q[v * vst + nx] = c[v];
                            if (img_l) q[v * vst - nx] = c[v];
                            if (WRITE_TMP && st_ok) q[v * vst + tmp_off] = st.W[1][v];
                        }
                    } else {
                        const uint32_t q = tdst + 8u * ((j - 1 - lo1) & (Z3_TRING - 1)) * Z3_RS;
#pragma unroll
                        for (int v = 0; v < 4; ++v) z3_sts<0>(q + 8u * v * Z3_VS, c[v]);
                    }
                }
            }
            if (r & 1) z3_barrier(tok);
        }
    }
}

}  // namespace pmw
