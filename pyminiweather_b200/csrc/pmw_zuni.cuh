// "Universal" iteration of the streaming z sweep (sweep_z, pmw_sweep.cuh): ONE straight-line body for every
// iteration of a segment -- pipeline fill, steady state, drain and the rows next to a wall alike.
//
// Why.  The production loop runs the software-pipelined steady body (three interleaved interface
// evaluations) only where all three RK stages are active and no wall is near; every other iteration of a
// segment -- 16 of 67 at 2048x1024 -- goes through a generic path that runs the active stages one after
// the other inside warp-uniform branches.  ncu (profiles/r1k): those 24 % of the iterations take 35 % of
// the kernel.  Here the three evaluations are ALWAYS issued together; a stage that is not active yet (or
// any more) simply computes on whatever its window holds and its results are discarded:
//   * activity only gates the stores and the slow-path vote, never the arithmetic;
//   * the cells a stage hands to the next one outside its valid range are garbage, but the next stage's
//     valid evaluations never read them (same data flow as the generic path, which hands over zeros);
//   * the halo cells beyond a wall (set_bc_z, bcs.py:92-148) are rebuilt in the register windows behind
//     ONE rare warp-uniform branch at the top of the iteration; the wall flag of the flux (w = 0, no
//     density diffusion: interpolate.py:168-173) is a run-time select;
//   * rows beyond the segment's stream are clamped to its last row (never waited for twice: the ring
//     slot of the last row is not reused once the requests stop).
//
// The file is written against small policy types (Env: arithmetic and warp primitives; Stream: the ring of
// state rows; Out: the stores) so that the SAME control flow compiles for the device (DeviceEnv /
// ZStream in pmw_sweep.cuh) and for the host (tools/zuni_probe, which checks it against the NumPy oracle
// without a GPU).  Selected by PMW_ZSWEEP_UNIVERSAL (pmw_sweep.cuh).
#pragma once

#ifndef PMW_ZU_FN
#define PMW_ZU_FN __device__ __forceinline__
#endif

namespace pmw {

struct ZUStage {
    double W[4][4];   // [slot][variable]: tap t of a stage at rotation R0 lives in slot (R0 + t) & 3
    double fprev[4];  // flux through the previous interface
};

struct ZUBounds {
    int lo1, hi1, lo2, hi2, lo3, hi3;  // interface ranges of the three stages in this segment
    int nz;
};

// Rebuild the cells beyond a wall in the window of one stage whose current interface is k (taps are the
// cells k-2 .. k+1).  Same cases as ZStage::step.
template <int R0, class Env, class St>
PMW_ZU_FN void zu_wall(const Env& env, St& s, int k, int nz)
{
    if (k == 0) {  // cells -2, -1 from interior cell 0 = tap 2
        const double h2 = env.hd(2), h0 = env.hd(0), h1 = env.hd(1);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            s.W[R0 & 3][v] = env.wall_value(v, s.W[(R0 + 2) & 3][v], h2, h0);
            s.W[(R0 + 1) & 3][v] = env.wall_value(v, s.W[(R0 + 2) & 3][v], h2, h1);
        }
    }
    if (k == nz - 1) {  // cell nz from interior cell nz-1 = tap 2
        const double hi = env.hd(nz + 1), h = env.hd(nz + 2);
#pragma unroll
        for (int v = 0; v < 4; ++v) s.W[(R0 + 3) & 3][v] = env.wall_value(v, s.W[(R0 + 2) & 3][v], hi, h);
    }
    if (k == nz) {  // cells nz, nz+1 from interior cell nz-1 = tap 1
        const double hi = env.hd(nz + 1), h2 = env.hd(nz + 2), h3 = env.hd(nz + 3);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            s.W[(R0 + 2) & 3][v] = env.wall_value(v, s.W[(R0 + 1) & 3][v], hi, h2);
            s.W[(R0 + 3) & 3][v] = env.wall_value(v, s.W[(R0 + 1) & 3][v], hi, h3);
        }
    }
}

// One iteration at window rotation R: stage 1 at interface j, stage 2 at j-3, stage 3 at j-6.
// (St: any type with W[4][4] and fprev[4] -- ZUStage here, ZStage<POW_MODE> in the hybrid loop of sweep_z.)
template <int R, bool WRITE_TMP, bool HAS_SRC, class Env, class Stream, class Out, class St>
PMW_ZU_FN void zu_iter(const Env& env, const Stream& zs, St& s1, St& s2, St& s3, int j,
                       const ZUBounds& b, const Out& out)
{
    const int nz = b.nz;
    env.syncwarp();  // every lane is done with the rows of the previous iteration
    zs.request_ahead(j + 1);
    const int mt = (j + 1 < zs.last_cell) ? j + 1 : zs.last_cell;  // newest state row, clamped to the stream
    zs.wait(mt);
    {
        double top[4];
        zs.load(mt, top);
#pragma unroll
        for (int v = 0; v < 4; ++v) s1.W[R & 3][v] = top[v];  // replaces the oldest tap: stage 1's taps now start at slot R+1
    }
    const int k1 = j, k2 = j - 3, k3 = j - 6;
    if ((k3 <= 0 && k1 >= 0) || (k1 >= nz - 1 && k3 <= nz)) {  // rare, warp-uniform: a wall is in reach
        zu_wall<R + 1>(env, s1, k1, nz);
        zu_wall<R>(env, s2, k2, nz);
        zu_wall<R>(env, s3, k3, nz);
    }
    const bool w1 = (k1 == 0 || k1 == nz), w2 = (k2 == 0 || k2 == nz), w3 = (k3 == 0 || k3 == nz);
    const bool act1 = k1 >= b.lo1 && k1 <= b.hi1, act2 = k2 >= b.lo2 && k2 <= b.hi2, act3 = k3 >= b.lo3 && k3 <= b.hi3;
    const auto bg1 = env.bg(env.clampi(k1, 0, nz)), bg2 = env.bg(env.clampi(k2, 0, nz)), bg3 = env.bg(env.clampi(k3, 0, nz));
    double f1[4], f2[4], f3[4], c1[4], c2[4], c3[4];
    const bool bad3 = env.flux(s3.W[R & 3], s3.W[(R + 1) & 3], s3.W[(R + 2) & 3], s3.W[(R + 3) & 3], bg3, w3, f3);
    const bool bad2 = env.flux(s2.W[R & 3], s2.W[(R + 1) & 3], s2.W[(R + 2) & 3], s2.W[(R + 3) & 3], bg2, w2, f2);
    const bool bad1 = env.flux(s1.W[(R + 1) & 3], s1.W[(R + 2) & 3], s1.W[(R + 3) & 3], s1.W[R & 3], bg1, w1, f1);
    if (env.any((bad1 && act1) || (bad2 && act2) || (bad3 && act3))) {
        env.cold_path_fence();
        if (bad3 && act3) env.flux_slow(s3.W[R & 3], s3.W[(R + 1) & 3], s3.W[(R + 2) & 3], s3.W[(R + 3) & 3], bg3, w3, f3);
        if (bad2 && act2) env.flux_slow(s2.W[R & 3], s2.W[(R + 1) & 3], s2.W[(R + 2) & 3], s2.W[(R + 3) & 3], bg2, w2, f2);
        if (bad1 && act1) env.flux_slow(s1.W[(R + 1) & 3], s1.W[(R + 2) & 3], s1.W[(R + 3) & 3], s1.W[R & 3], bg1, w1, f1);
    }
    double in2[4], in3[4];  // initial state of the cells stages 2 and 3 finish (cells k2-1, k3-1)
    zs.load(env.clampi(j - 4, zs.f0, zs.last_cell), in2);
    zs.load(env.clampi(j - 7, zs.f0, zs.last_cell), in3);
    double g1 = 0.0, g2 = 0.0, g3 = 0.0;
    if (HAS_SRC) {
        g3 = env.src(env.clampi(k3 - 1, 0, nz - 1));
        g2 = env.src(env.clampi(k2 - 1, 0, nz - 1));
        g1 = env.src(env.clampi(k1 - 1, 0, nz - 1));
    }
    const bool st3 = act3 && k3 > b.lo3;  // cell k3-1 is owned and both its fluxes are valid
    if (WRITE_TMP && st3) out.store_tmp(k3 - 1, s3.W[(R + 1) & 3]);  // T2 of the cell (tap 1 of stage 3)
    // cell k-1 of every stage (tap 1 of its window) from the fluxes through its two faces
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        c3[v] = env.template update<HAS_SRC>(v, s3.fprev[v], f3[v], in3[v], 3, s3.W[(R + 1) & 3][0], g3);
        c2[v] = env.template update<HAS_SRC>(v, s2.fprev[v], f2[v], in2[v], 2, s2.W[(R + 1) & 3][0], g2);
        c1[v] = env.template update<HAS_SRC>(v, s1.fprev[v], f1[v], s1.W[(R + 2) & 3][v], 1, s1.W[(R + 2) & 3][0], g1);
        s3.fprev[v] = f3[v];
        s2.fprev[v] = f2[v];
        s1.fprev[v] = f1[v];
    }
    if (st3) out.store(k3 - 1, c3);
#pragma unroll
    for (int v = 0; v < 4; ++v) {  // the new cells replace the oldest taps of the next stage
        s3.W[R & 3][v] = c2[v];
        s2.W[R & 3][v] = c1[v];
    }
}

// All iterations of a segment.  On entry stage 1's window holds the first three state rows of the stream
// in slots 1..3 (taps 0..2 of interface lo1 at rotation 0), everything else is zero.
template <bool WRITE_TMP, bool HAS_SRC, class Env, class Stream, class Out>
PMW_ZU_FN void zu_segment(const Env& env, const Stream& zs, ZUStage& s1, ZUStage& s2, ZUStage& s3,
                          const ZUBounds& b, const Out& out)
{
    for (int j = b.lo1; j <= b.hi3 + 6; j += 4) {  // the last block may run up to three idle iterations
        zu_iter<0, WRITE_TMP, HAS_SRC>(env, zs, s1, s2, s3, j, b, out);
        zu_iter<1, WRITE_TMP, HAS_SRC>(env, zs, s1, s2, s3, j + 1, b, out);
        zu_iter<2, WRITE_TMP, HAS_SRC>(env, zs, s1, s2, s3, j + 2, b, out);
        zu_iter<3, WRITE_TMP, HAS_SRC>(env, zs, s1, s2, s3, j + 3, b, out);
    }
}

}  // namespace pmw
