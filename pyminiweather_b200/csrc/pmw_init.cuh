// Device-side initial conditions: `init` of the reference (pyminiweather/ics/initial.py:57-80) for the
// 2-D state -- 3x3 Gauss-Legendre quadrature of the configuration's (rho', u, w, theta', rho_hy,
// theta_hy) over every cell of [nz+4][nx+4], halo cells included.  The reference materialises
// [nz+4, nx+4, 3, 3] temporaries for this (9x the state, README.md:156); here one thread owns one
// cell and nothing but the state is written.  The 1-D hydrostatic profiles stay on the host
// (initial.py:84-105: nz-sized, and they must be bit-identical to the reference's).
//
// Expression order follows the reference (utils/utils.py:40-51, initial_conditions.py:26-81,
// initial.py:62-77) and the library is built with -fmad=false, so coordinates and distances are
// bit-identical to NumPy's; pow / exp / cos come from the CUDA math library and may differ from
// NumPy's by an ulp (tests: <= 1e-13 relative L2 against the host init).
#pragma once
#include "pmw_common.cuh"

namespace pmw {

constexpr int IC_MAX_BUBBLES = 4;
constexpr double IC_PI = 3.14159265358979323846264338327;  // constants.py:15
constexpr double IC_THETA0 = 300.0, IC_EXNER0 = 1.0;       // constants.py:16-17

struct IcSpec {
    int nbubbles;                      // squared-cosine potential-temperature bubbles
    double amp[IC_MAX_BUBBLES], x0[IC_MAX_BUBBLES], z0[IC_MAX_BUBBLES], xrad[IC_MAX_BUBBLES], zrad[IC_MAX_BUBBLES];
    double wind;                       // uniform u (gravity: 15 m/s)
    int bvfreq;                        // 0: constant-theta background, 1: constant Brunt-Vaisala frequency
    double bv0;
    double dx, dz;
};

// initial_conditions.py:26-52 / :54-81
__host__ __device__ inline void ic_background(const IcSpec& s, double z, double& hr, double& ht)
{
    double exner;
    if (s.bvfreq) {
        ht = IC_THETA0 * exp(s.bv0 * s.bv0 / GRAV * z);
        exner = IC_EXNER0 - GRAV * GRAV / (CP * s.bv0 * s.bv0) * (ht - IC_THETA0) / (ht * IC_THETA0);
    } else {
        ht = IC_THETA0;
        exner = IC_EXNER0 - GRAV * z / (CP * IC_THETA0);
    }
    const double p = P0 * pow(exner, CP / RD);
    hr = pow(p / C0, 1.0 / GAMMA) / ht;
}

// utils/utils.py:40-51
__host__ __device__ inline double ic_bubble(double x, double z, double amp, double x0, double z0, double xrad,
                                            double zrad)
{
    const double ax = (x - x0) / xrad, az = (z - z0) / zrad;
    const double dist = sqrt(ax * ax + az * az) * IC_PI / 2.0;
    if (!(dist <= IC_PI / 2.0)) return 0.0;
    const double c = cos(dist);
    return amp * (c * c);
}

// One cell: the 3x3 quadrature sums of rho*u and (rho*theta)' (initial.py:62-77) for the cell whose
// lower-left corner is (xc, zc).  rho' and rho*w integrate to exactly zero in every configuration.
// (__host__ as well so that tools/init_probe can run the same code on the CPU.)
__host__ __device__ inline void ic_cell(const IcSpec& s, double xc, double zc, double& su, double& st)
{
    // quadrature.py:9-21: 3-point Gauss-Legendre on [0,1]
    const double qp[3] = {0.112701665379258311482073460022, 0.5, 0.887298334620741688517926539980};
    const double qw[3] = {0.277777777777777777777777777779, 0.444444444444444444444444444444,
                          0.277777777777777777777777777779};
    su = 0.0;
    st = 0.0;  // running sums over the quadrature rows (z index)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double z = zc + qp[a] * s.dz;
        double hr, ht;
        ic_background(s, z, hr, ht);
        double ru = 0.0, rt = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double x = xc + qp[b] * s.dx;
            double t = 0.0;
            for (int n = 0; n < s.nbubbles; ++n) t = t + ic_bubble(x, z, s.amp[n], s.x0[n], s.z0[n], s.xrad[n], s.zrad[n]);
            const double w = qw[a] * qw[b];  // qweights_outer
            const double vu = ((0.0 + hr) * s.wind) * w;
            const double vt = ((0.0 + hr) * (t + ht) - hr * ht) * w;
            ru = (b == 0) ? vu : ru + vu;  // .sum(axis=-1): left to right
            rt = (b == 0) ? vt : rt + vt;
        }
        su = (a == 0) ? ru : su + ru;      // second .sum(axis=-1)
        st = (a == 0) ? rt : st + rt;
    }
}

// x_axis[nx+4], z_axis[nz+4]: lower-left corner coordinates of the array's columns / rows, exactly
// as the reference's mesh produces them (mesh.py:22-40) -- for a slab, its own part of the x axis.
__global__ void __launch_bounds__(256) init_state_kernel(double* __restrict__ state, double* __restrict__ state_tmp,
                                                         const Layout L, const IcSpec s,
                                                         const double* __restrict__ x_axis,
                                                         const double* __restrict__ z_axis)
{
    const int NX = L.nx + 2 * HS, NZ = L.nz + 2 * HS;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= (long long)NX * NZ) return;
    const int k = (int)(c / NX), i = (int)(c % NX);
    double su, st;
    ic_cell(s, x_axis[i], z_axis[k], su, st);
    const double vals[NVAR] = {0.0, su, 0.0, st};
#pragma unroll
    for (int v = 0; v < NVAR; ++v) {
        state[idx(L, v, k, i)] = vals[v];
        state_tmp[idx(L, v, k, i)] = vals[v];  // initial.py:80
    }
}

}  // namespace pmw
