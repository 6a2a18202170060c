// Unfused operator shims: interpolate_x/z, compute_flux_x/z, compute_tend_x/z as separate kernels
// that materialise the reference's intermediate arrays (fields.vals_*, d3_vals_*, flux, tend;
// pyminiweather/data/fields.py:70-78).  The production path never runs these -- the fused stage
// kernels keep all of it in registers -- they exist so that code written against the reference's
// individual operators (e.g. its tests/unit/test_interpolate.py) keeps working.  Every expression
// is written in the reference's operation order and the library is compiled with -fmad=false, so
// interpolation and tendencies are bit-identical to NumPy; fluxes differ only through pow().
#pragma once
#include "pmw_common.cuh"

namespace pmw {

// interpolate.py:33-43 / 69-79.  vals, d3: dense [4][nz][nx+1] (x) or [4][nz+1][nx] (z).
template <bool DIR_Z>
__global__ void interpolate_kernel(const double* __restrict__ s, const Layout L, double* vals, double* d3)
{
    const int ni = DIR_Z ? L.nx : L.nx + 1, nk = DIR_Z ? L.nz + 1 : L.nz;
    const long long n = (long long)NVAR * nk * ni;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % ni), k = (int)((t / ni) % nk), v = (int)(t / ((long long)ni * nk));
    double s0, s1, s2, s3;
    if (DIR_Z) {
        const double* p = s + idx(L, v, k, i + HS);
        s0 = p[0]; s1 = p[L.pitch]; s2 = p[2 * L.pitch]; s3 = p[3 * L.pitch];
    } else {
        const double* p = s + idx(L, v, k + HS, i);
        s0 = p[0]; s1 = p[1]; s2 = p[2]; s3 = p[3];
    }
    const double c0 = -1.0 / 12, c1 = 7.0 / 12;  // fields.py:94-96
    vals[t] = ((c0 * s0 + c1 * s1) + c1 * s2) + c0 * s3;
    d3[t] = ((-1.0 * s0 + 3.0 * s1) + -3.0 * s2) + 1.0 * s3;  // fields.py:97, flipped by the convolution
}

// interpolate.py:95-129 / 144-186.  flux: dense [4][nz+1][nx+1]; x fills [:, :nz, :nx+1], z [:, :nz+1, :nx].
template <bool DIR_Z>
__global__ void flux_kernel(const double* __restrict__ vals, const double* __restrict__ d3, const Layout L,
                            const Hydro hy, double hv, double* flux)
{
    const int ni = DIR_Z ? L.nx : L.nx + 1, nk = DIR_Z ? L.nz + 1 : L.nz;
    const long long plane = (long long)nk * ni;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plane) return;
    const int i = (int)(t % ni), k = (int)(t / ni);
    const double hd = DIR_Z ? hy.dens_int[k] : hy.dens_cell[k + HS];
    const double hdt = DIR_Z ? hy.dens_theta_int[k] : hy.dens_theta_cell[k + HS];
    const double rho = vals[t] + hd;
    const double u = vals[plane + t] / rho;
    double w = vals[2 * plane + t] / rho;
    const double th = (vals[3 * plane + t] + hdt) / rho;
    double p = C0 * pow(rho * th, GAMMA);
    double d3d = d3[t];
    if (DIR_Z) {
        p = p - hy.pressure_int[k];
        if (k == 0 || k == L.nz) { w = 0.0; d3d = 0.0; }  // interpolate.py:168-173
    }
    const long long fplane = (long long)(L.nz + 1) * (L.nx + 1);
    double* f = flux + (long long)k * (L.nx + 1) + i;
    if (DIR_Z) {
        f[0] = rho * w - hv * d3d;
        f[fplane] = rho * w * u - hv * d3[plane + t];
        f[2 * fplane] = rho * (w * w) + p - hv * d3[2 * plane + t];
        f[3 * fplane] = rho * w * th - hv * d3[3 * plane + t];
    } else {
        f[0] = rho * u - hv * d3d;
        f[fplane] = rho * (u * u) + p - hv * d3[plane + t];
        f[2 * fplane] = rho * u * w - hv * d3[2 * plane + t];
        f[3 * fplane] = rho * u * th - hv * d3[3 * plane + t];
    }
}

// interpolate.py:208-215 / 238-250.  tend: dense [4][nz][nx].
template <bool DIR_Z>
__global__ void tend_kernel(const double* __restrict__ flux, const double* __restrict__ s, const Layout L, double d,
                            double* tend)
{
    const long long n = (long long)NVAR * L.nz * L.nx;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = (int)(t % L.nx), k = (int)((t / L.nx) % L.nz), v = (int)(t / ((long long)L.nx * L.nz));
    const long long fplane = (long long)(L.nz + 1) * (L.nx + 1);
    const double* f = flux + v * fplane + (long long)k * (L.nx + 1) + i;
    const double hi = DIR_Z ? f[L.nx + 1] : f[1];
    double td = -(hi - f[0]) / d;
    if (DIR_Z && v == WMOM) td -= s[idx(L, DENS, k + HS, i + HS)] * GRAV;
    tend[t] = td;
}

}  // namespace pmw
