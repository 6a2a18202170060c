// Shared device-side definitions for the PyMiniWeather hot path on sm_100a.
//
// Device layout of one state buffer (chosen for 128-byte aligned interior rows,
// TMA-legal strides and vector access; the reference's host layout is
// [4][nz+4][nx+4] dense, pyminiweather/data/fields.py:67-72):
//
//     element (v, k, i), i in [0, nx+4)  ->  base[v*vstride + k*pitch + i]
//
// where `base` = allocation + LPAD doubles, LPAD = 14, so that the first
// interior column (i = 2) sits on a 128-byte boundary; pitch is a multiple of
// 16 doubles (128 B) and vstride = pitch * (nz+4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmw {

constexpr int HS = 2;
constexpr int NVAR = 4;
constexpr int LPAD = 14;
enum { DENS = 0, UMOM = 1, WMOM = 2, RHOT = 3 };  // pyminiweather/__init__.py:14-18

// pyminiweather/data/constants.py:4-27
constexpr double HV_BETA = 0.05;
constexpr double P0 = 1.0e5;
constexpr double C0 = 27.5629410929725921310572974482;
constexpr double GAMMA = 1.40027894002789400278940027894;
constexpr double GRAV = 9.8;
constexpr double CP = 1004.0;
constexpr double CV = 717.0;
constexpr double RD = 287.0;

struct Layout {
    int nx, nz;       // interior size of this slab
    int pitch;        // doubles per row
    long long vstride;  // doubles per variable plane
};

__host__ __device__ inline long long idx(const Layout& L, int v, int k, int i)
{
    return (long long)v * L.vstride + (long long)k * L.pitch + i;
}

// Hydrostatic background profiles on the device (pyminiweather/ics/initial.py:84-105) plus
// derived per-row tables for the background-relative pressure evaluation.
struct Hydro {
    const double* dens_cell;        // [nz+4]
    const double* dens_theta_cell;  // [nz+4]
    const double* dens_int;         // [nz+1]
    const double* dens_theta_int;   // [nz+1]
    const double* pressure_int;     // [nz+1]
    const double* inv_dens_theta_cell;  // [nz+4]  1/dens_theta_cell
    const double* pressure_cell;        // [nz+4]  C0*dens_theta_cell^gamma
    const double* inv_dens_theta_int;   // [nz+1]  1/dens_theta_int
    // the four interface profiles interleaved, {dens, dens_theta, 1/dens_theta, pressure} per interface, for
    // interfaces -HY_PACK_PAD .. nz+HY_PACK_PAD (clamped copies beyond the domain): int_pack[4*k + 0..3];
    // 32-byte aligned entries, so that a z sweep can pull them into shared memory with bulk copies
    const double* int_pack;
    // the four cell-row profiles of the x sweeps interleaved the same way, {dens, dens_theta, 1/dens_theta,
    // C0*dens_theta^gamma} per array row: cell_pack[4*k + 0..3], k in [0, nz+4)
    const double* cell_pack;
    const double* pad_;  // keeps the members that follow a Hydro in the kernel argument structs on their 16-byte alignment
};
constexpr int HY_PACK_PAD = 8;

struct StageArgs {
    Layout L;
    const double* forcing;
    const double* init;
    double* out;
    double* out_left;   // same logical buffer on the left  slab neighbour (== out when periodic_x)
    double* out_right;  // same logical buffer on the right slab neighbour
    Hydro hy;
    const double* src_w;  // [nz][nx] extra rho*w tendency (ic_type "gravity", source.py:43-50) or nullptr;
                          // the TMA kernels read it only in their HAS_SRC instantiations
    double hv_coeff;  // -hv_beta*d/(16*dt_full)   (interpolate.py:101,149)
    double inv_d;     // 1/dx or 1/dz
    double dt_stage;
    int write_xhalo;  // also store the periodic/neighbour images of the edge columns
    int fuse_bc_z;    // z stages: rebuild the wall halo rows on the fly (set_bc_z folded in)
    // slab ring over peer memory: an x stage's edge tiles wait until the neighbour that owns the
    // halo columns has published epoch >= wait_epoch (flags[0]: left neighbour, flags[1]: right,
    // flags[2]: watchdog error).  wait_epoch == 0: nothing to wait for.
    unsigned long long* flags;
    unsigned long long wait_epoch;
    int edge_last;  // walk tile columns first so that the two edge columns are the last CTAs of the grid
    // ... while the FIRST row of CTAs of the same kernel (blockIdx.y == 0, present when push_epoch != 0)
    // stores this slab's own edge columns of `forcing` into the neighbours' halo columns and
    // publishes push_epoch to them (we are the left neighbour's RIGHT neighbour: its flags[1]).
    unsigned long long push_epoch;
    double* nbr_forcing_left;
    double* nbr_forcing_right;
    unsigned long long* nbr_flags_left;
    unsigned long long* nbr_flags_right;
    unsigned int* push_counter;
    // L2 eviction priority per operand: 0 normal, 1 evict_first, 2 evict_last (createpolicy)
    int hint_forcing, hint_init, hint_out;
    // Chunked sweeps: this launch covers only the tile columns (z stages) / tile rows (x stages)
    // starting at these offsets; the grid dimensions give the extent.
    int tile_x0, tile_y0;
    double cd, cg;  // dt_stage/d and dt_stage*grav (cell_update)
};

// The stage update of one cell, out = init + dt_stage * tend with
//     tend = -(F_hi - F_lo)/d  [- rho'*grav: z sweeps, rho*w]  [+ src: gravity-wave forcing, rho*w]
// (interpolate.py:208-215,238-250, source.py:43-50, step.py:80-82), folded into fused multiply-adds
// with the host-computed factors cd = dt_stage/d and cg = dt_stage*grav: one multiplication per cell
// and variable fewer than tend-then-update, same value to rounding.  EVERY kernel updates cells
// through this function, which is what keeps the kernel variants bit-identical to each other.
template <bool HYDRO_SRC, bool EXTRA_SRC>
__device__ __forceinline__ double cell_update(double f_lo, double f_hi, double init, double cd, double cg,
                                              double dens, double dt_stage, double src)
{
    double base = init;
    if (HYDRO_SRC) base = fma(-cg, dens, base);
    if (EXTRA_SRC) base = fma(dt_stage, src, base);
    return fma(cd, f_lo - f_hi, base);
}

__device__ __forceinline__ unsigned long long l2_policy(int kind)
{
    unsigned long long p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// The push role of an x stage (see StageArgs::push_epoch): every thread of the first row of CTAs.
__device__ __forceinline__ void push_halo_role(const StageArgs& a)
{
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    const Layout& L = a.L;
    for (int t = blockIdx.x * nthr + tid; t < NVAR * L.nz; t += gridDim.x * nthr) {
        const int k = t % L.nz, v = t / L.nz;
        const double2 first = *reinterpret_cast<const double2*>(a.forcing + idx(L, v, k + HS, HS));
        const double2 last = *reinterpret_cast<const double2*>(a.forcing + idx(L, v, k + HS, L.nx));
        *reinterpret_cast<double2*>(a.nbr_forcing_left + idx(L, v, k + HS, L.nx + HS)) = first;  // its right halo
        *reinterpret_cast<double2*>(a.nbr_forcing_right + idx(L, v, k + HS, 0)) = last;          // its left halo
    }
    __threadfence_system();  // peer stores (NVLink) ordered before the flag
    __syncthreads();
    if (tid == 0 && atomicAdd(a.push_counter, 1u) == gridDim.x - 1) {
        *a.push_counter = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.nbr_flags_left + 1), "l"(a.push_epoch) : "memory");
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.nbr_flags_right + 0), "l"(a.push_epoch) : "memory");
    }
}

// ---------------------------------------------------------------------------------
// (1+e)^gamma - 1 = e * g(e) on |e| <= 1/8, g the degree-10 interpolant of ((1+e)^gamma-1)/e at
// Chebyshev nodes (tools/pow_poly.py generates the coefficients and checks them against mpmath:
// 3.1e-17 absolute error in exact arithmetic with these double coefficients, 6.2e-17 as evaluated --
// below half an ulp of the pressure; degrees 11 and 12 are no better once their coefficients are
// rounded to double, degree 9 is 3.7e-16).  Two interleaved Horner chains in e^2.  PMW_POW_BACKGROUND.
// ---------------------------------------------------------------------------------
// The coefficients live in the constant bank so that they are direct operands of the DFMAs (as
// immediates every one of them costs two UMOVs per use: 17 % of the instructions of a fused sweep).
__constant__ double kPow1p[11] = {
    0x1.6678ae3cb2859p+0,    // c0
    0x1.1efa23f1c098cp-2,    // c1
    -0x1.caf32d76b9336p-5,   // c2
    0x1.6f188e364e3d0p-6,    // c3
    -0x1.7dbd21d5f3b8dp-7,   // c4
    0x1.ca0d2931020eap-8,    // c5
    -0x1.2cfcacfb49727p-8,   // c6
    0x1.a542670ff1b05p-9,    // c7
    -0x1.34e6f29c10c68p-9,   // c8
    0x1.e27f3dba93b93p-10,   // c9
    -0x1.79a6778274f33p-10,  // c10
};
__constant__ double kInterp[2] = {-1.0 / 12, 7.0 / 12};  // fields.py:94-96
__constant__ double kGrav = GRAV;

__device__ __forceinline__ double pow1p_gamma_m1(double e)
{
    const double e2 = e * e;
    double a = kPow1p[10];
    double b = kPow1p[9];
    a = fma(a, e2, kPow1p[8]);
    b = fma(b, e2, kPow1p[7]);
    a = fma(a, e2, kPow1p[6]);
    b = fma(b, e2, kPow1p[5]);
    a = fma(a, e2, kPow1p[4]);
    b = fma(b, e2, kPow1p[3]);
    a = fma(a, e2, kPow1p[2]);
    b = fma(b, e2, kPow1p[1]);
    a = fma(a, e2, kPow1p[0]);
    return fma(b, e, a) * e;
}

// 1/x for normal, positive x (densities): MUFU.RCP64H seed and ONE cubically convergent step,
//     r = r0 (1 + e + e^2),  e = 1 - x r0:
// the seed is good to 2^-19.9 (tools/arith_probe/rcp_probe.cu), so the result is good to 2^-59.7 before
// its final rounding -- within 0.51 ulp of 1/x.  Straight-line code: __drcp_rn carries a branch for special
// operands, which splits the basic block of every interface evaluation and keeps the scheduler from
// interleaving independent evaluations.  (Two Newton steps -- four dependent FMAs, correctly rounded on all
// 2e8 samples of the probe -- were one FMA more on the longest dependency chain of an interface: the
// sweeps are 1.6 % faster with the cubic step, profiles/r2ba_ab.log.)
__device__ __forceinline__ double rcp_pos(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// Everything one interface needs besides the 4x4 stencil values.
struct IfaceBg {
    double dens;       // hydrostatic density at the interface (x: cell row value)
    double dens_theta; // hydrostatic rho*theta
    double inv_dens_theta;
    double pressure;   // x: C0*dens_theta^gamma of the row;  z: hy_pressure_int[k]
};

// Interface flux from the four stencil taps of the four variables.
//   DIR_Z = false: compute_flux_x (interpolate.py:105-129)
//   DIR_Z = true : compute_flux_z (interpolate.py:153-186), `wall` = (k==0 || k==nz)
// Interpolation weights: fields.py:94-97 (4th-order value, flipped 3rd difference).
// Differences from the reference's rounding: fused multiply-adds, one reciprocal of rho instead
// of three divisions, optional polynomial pressure -- all validated at <= 1e-12 rel-L2
// (tools/arith_probe).  The library is compiled with -fmad=false and every FMA below is
// written out: ptxas may otherwise contract mul+add pairs differently in different kernels
// (PTX mul/add without a rounding modifier are contractible), and then the two kernel
// variants -- or two tiles of one kernel -- would disagree in the last bit for the same cell.
// Algebraically reduced form (default).  With m = val[UMOM or WMOM] the interpolated momentum along the
// sweep and X = val[RHOT] + (rho*theta)_hy the interpolated rho*theta, the reference's
//     u = m / rho;  t = X / rho;  rho*u -> m;  rho*t -> X;  rho*u*t -> u*X;  rho*u*u -> m*u
// only differ from the right-hand sides by roundings (1e-16 relative), so
//     F_D = m - hv d3_D,  F_m = m*u + p - hv d3_m,  F_other = m*w - hv d3_other,  F_T = u*X - hv d3_T,
//     p from e = val[RHOT] / (rho*theta)_hy  (no cancellation, no dependence on 1/rho).
// Four FP64 operations fewer per interface, and -- what matters on a latency-bound FP64 pipe -- the
// pressure polynomial no longer waits for the reciprocal: the critical path of an interface drops from
// ~27 to ~16 dependent operations.  Parity with the reference is unchanged (tests: <= 1e-11 / 1e-12).
template <bool DIR_Z, int POW_MODE, bool FAST>
__device__ __forceinline__ bool interface_flux_core(const double (&s0)[4], const double (&s1)[4],
                                                    const double (&s2)[4], const double (&s3)[4],
                                                    const IfaceBg& bg, double hv, bool wall,
                                                    double (&flux)[4])
{
    const double c0 = kInterp[0], c1 = kInterp[1];
    double val[4], d3[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        val[v] = fma(c0, s3[v], fma(c1, s2[v], fma(c1, s1[v], c0 * s0[v])));
        d3[v] = fma(-3.0, s2[v], fma(3.0, s1[v], -s0[v])) + s3[v];
    }
    const double rho = val[DENS] + bg.dens;
    const double r = rcp_pos(rho);
    const double X = val[RHOT] + bg.dens_theta;         // rho*theta at the interface
    const double e = val[RHOT] * bg.inv_dens_theta;     // its relative perturbation
    double p;  // x: full pressure; z: pressure perturbation p - hy_pressure_int
    bool bad = false;
    if (POW_MODE == 1) {
        if (FAST || fabs(e) <= 0.125) {
            const double f = pow1p_gamma_m1(e);
            p = DIR_Z ? bg.pressure * f : fma(bg.pressure, f, bg.pressure);
            bad = !(fabs(e) <= 0.125);
        } else {
            p = C0 * pow(X, GAMMA);
            if (DIR_Z) p -= bg.pressure;
        }
    } else {
        p = C0 * pow(X, GAMMA);
        if (DIR_Z) p -= bg.pressure;
    }
    const double u = val[UMOM] * r;
    double w = val[WMOM] * r;
    if (DIR_Z) {
        double m = val[WMOM];  // rho*w
        if (wall) { w = 0.0; m = 0.0; d3[DENS] = 0.0; }
        flux[DENS] = fma(-hv, d3[DENS], m);
        flux[UMOM] = fma(-hv, d3[UMOM], m * u);
        flux[WMOM] = fma(-hv, d3[WMOM], fma(m, w, p));
        flux[RHOT] = fma(-hv, d3[RHOT], w * X);
    } else {
        const double m = val[UMOM];  // rho*u
        flux[DENS] = fma(-hv, d3[DENS], m);
        flux[UMOM] = fma(-hv, d3[UMOM], fma(m, u, p));
        flux[WMOM] = fma(-hv, d3[WMOM], m * w);
        flux[RHOT] = fma(-hv, d3[RHOT], u * X);
    }
    return bad;
}

template <bool DIR_Z, int POW_MODE>
__device__ __forceinline__ void interface_flux(const double (&s0)[4], const double (&s1)[4],
                                               const double (&s2)[4], const double (&s3)[4],
                                               const IfaceBg& bg, double hv, bool wall,
                                               double (&flux)[4])
{
    interface_flux_core<DIR_Z, POW_MODE, false>(s0, s1, s2, s3, bg, hv, wall, flux);
}

// The same evaluation without the range branch of the background-relative pressure: always the
// polynomial; returns true when |e| > 1/8 (or NaN), in which case `flux` must be recomputed with
// interface_flux.  The fused sweeps evaluate several interfaces per iteration and test all their
// flags with one warp vote instead of one divergent branch per interface.  Identical operations in
// identical order: where it returns false the result has the same bits as interface_flux.
template <bool DIR_Z, int POW_MODE>
__device__ __forceinline__ bool interface_flux_fast(const double (&s0)[4], const double (&s1)[4],
                                                    const double (&s2)[4], const double (&s3)[4],
                                                    const IfaceBg& bg, double hv, bool wall,
                                                    double (&flux)[4])
{
    return interface_flux_core<DIR_Z, POW_MODE, true>(s0, s1, s2, s3, bg, hv, wall, flux);
}

// Out-of-line fallback of the fused sweeps (rare: |e| > 1/8 somewhere in the warp).
struct Taps {
    double s[4][4];
};
struct Flux4 {
    double f[4];
};
// Arguments and result by value: nothing of the caller's is address-taken, so its flux arrays stay
// in registers on the hot path.
template <bool DIR_Z, int POW_MODE>
__device__ __noinline__ Flux4 interface_flux_slow(Taps t, IfaceBg bg, double hv, bool wall)
{
    Flux4 r;
    interface_flux<DIR_Z, POW_MODE>(t.s[0], t.s[1], t.s[2], t.s[3], bg, hv, wall, r.f);
    return r;
}

// Wall halo value for set_bc_z folded into a z stage (bcs.py:92-148): `interior` is the
// value of the nearest interior row in the same column, hd_* the hydrostatic densities.
__device__ __forceinline__ double wall_value(int v, double interior, double hd_interior,
                                             double hd_halo)
{
    if (v == WMOM) return 0.0;
    if (v == UMOM) return interior / hd_interior * hd_halo;
    return interior;
}

}  // namespace pmw
