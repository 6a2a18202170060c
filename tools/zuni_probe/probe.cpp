// Host run of the universal z iteration (csrc/pmw_zuni.cuh): the SAME control flow as the device kernel
// -- window rotation, activity masks, wall rebuilds, clamped row stream -- with host policies (plain C++
// arithmetic in place of the CUDA flux routine, the state array in place of the TMA ring).  Compared with
// the NumPy oracle by run_probe.py; a wrong row, mask or rotation shows up as an O(1) error.
#include <algorithm>
#include <cmath>
#include <cstddef>
#define PMW_ZU_FN static inline
#include "../../pyminiweather_b200/csrc/pmw_zuni.cuh"

namespace {
const double HV_BETA = 0.05, C0 = 27.5629410929725921310572974482, GAMMA = 1.40027894002789400278940027894, GRAV = 9.8;
enum { DENS = 0, UMOM = 1, WMOM = 2, RHOT = 3 };

struct Bg { double dens, dens_theta, pressure; };

struct HostEnv {
    int nx, nz, col;
    const double *hd_, *hy_dens_int, *hy_dens_theta_int, *hy_pressure_int, *src_w;
    double hv, inv_d, dt[4];
    double hd(int idx) const { return hd_[idx]; }
    double wall_value(int v, double interior, double h_in, double h_halo) const
    {
        if (v == WMOM) return 0.0;
        if (v == UMOM) return interior / h_in * h_halo;
        return interior;
    }
    Bg bg(int k) const { return Bg{hy_dens_int[k], hy_dens_theta_int[k], hy_pressure_int[k]}; }
    int clampi(int x, int lo, int hi) const { return std::min(std::max(x, lo), hi); }
    // compute_flux_z (interpolate.py:153-186) in the reference's own form
    bool flux(const double (&s0)[4], const double (&s1)[4], const double (&s2)[4], const double (&s3)[4], const Bg& bg,
              bool wall, double (&f)[4]) const
    {
        double val[4], d3[4];
        for (int v = 0; v < 4; ++v) {
            val[v] = -1.0 / 12 * s0[v] + 7.0 / 12 * s1[v] + 7.0 / 12 * s2[v] - 1.0 / 12 * s3[v];
            d3[v] = -s0[v] + 3.0 * s1[v] - 3.0 * s2[v] + s3[v];
        }
        const double rho = val[DENS] + bg.dens;
        const double u = val[UMOM] / rho;
        double w = val[WMOM] / rho;
        const double t = (val[RHOT] + bg.dens_theta) / rho;
        const double p = C0 * std::pow(rho * t, GAMMA) - bg.pressure;
        if (wall) { w = 0.0; d3[DENS] = 0.0; }
        f[DENS] = rho * w - hv * d3[DENS];
        f[UMOM] = rho * w * u - hv * d3[UMOM];
        f[WMOM] = rho * w * w + p - hv * d3[WMOM];
        f[RHOT] = rho * w * t - hv * d3[RHOT];
        return false;
    }
    void flux_slow(const double (&)[4], const double (&)[4], const double (&)[4], const double (&)[4], const Bg&, bool,
                   double (&)[4]) const {}
    bool any(bool p) const { return p; }
    void cold_path_fence() const {}
    void syncwarp() const {}
    double src(int m) const { return src_w[(size_t)m * nx + col]; }
    template <bool HAS_SRC>
    double update(int v, double f_lo, double f_hi, double init, int stage, double dens, double g) const
    {
        double tend = -(f_hi - f_lo) * inv_d;
        if (v == WMOM) {
            tend -= dens * GRAV;
            if (HAS_SRC) tend += g;
        }
        return init + dt[stage] * tend;
    }
};

struct HostStream {
    const double* S;  // [4][nz+4][nx+4]
    int nx, nz, col, f0, last_cell;
    void request_ahead(int) const {}
    void wait(int) const {}
    void load(int m, double (&r)[4]) const
    {
        const size_t NX = nx + 4, NZ = nz + 4;
        for (int v = 0; v < 4; ++v) r[v] = S[((size_t)v * NZ + (m + 2)) * NX + (col + 2)];
    }
};

struct HostOut {
    double *O, *T;
    int nx, nz, col;
    void store(int krow, const double (&c)[4]) const
    {
        const size_t NX = nx + 4, NZ = nz + 4;
        for (int v = 0; v < 4; ++v) O[((size_t)v * NZ + (krow + 2)) * NX + (col + 2)] = c[v];
    }
    void store_tmp(int krow, const double (&t)[4]) const
    {
        const size_t NX = nx + 4, NZ = nz + 4;
        for (int v = 0; v < 4; ++v) T[((size_t)v * NZ + (krow + 2)) * NX + (col + 2)] = t[v];
    }
};
}  // namespace

// One fused z sweep of the whole grid, segment by segment and column by column, exactly as sweep_z sets a
// warp up (pmw_sweep.cuh).  has_src: src_w [nz][nx].
extern "C" void zuni_sweep(int nx, int nz, int lz, double dz, double dt_full, double dt, const double* S, double* O,
                           double* T, const double* hd, const double* hy_dens_int, const double* hy_dens_theta_int,
                           const double* hy_pressure_int, const double* src_w)
{
    using namespace pmw;
    for (int lo3 = 0; lo3 < nz; lo3 += lz) {
        const int hi3 = std::min(lo3 + lz, nz);
        const int lo2 = std::max(lo3 - 2, 0), hi2 = std::min(hi3 + 2, nz);
        const int lo1 = std::max(lo3 - 4, 0), hi1 = std::min(hi3 + 4, nz);
        for (int col = 0; col < nx; ++col) {
            HostEnv env{nx, nz, col, hd, hy_dens_int, hy_dens_theta_int, hy_pressure_int, src_w,
                        -HV_BETA * dz / (16 * dt_full), 1.0 / dz, {0.0, dt / 3, dt / 2, dt}};
            HostStream zs{S, nx, nz, col, lo1 - 2, hi1 + 1};
            HostOut out{O, T, nx, nz, col};
            ZUStage s1, s2, s3;
            for (int t = 0; t < 4; ++t)
                for (int v = 0; v < 4; ++v) s1.W[t][v] = s2.W[t][v] = s3.W[t][v] = 0.0;
            for (int v = 0; v < 4; ++v) s1.fprev[v] = s2.fprev[v] = s3.fprev[v] = 0.0;
            for (int n = 0; n < 3; ++n) {  // the first three taps of stage 1: slots 1..3
                double r[4];
                zs.load(zs.f0 + n, r);
                for (int v = 0; v < 4; ++v) s1.W[n + 1][v] = r[v];
            }
            const ZUBounds b{lo1, hi1, lo2, hi2, lo3, hi3, nz};
            if (src_w) zu_segment<true, true>(env, zs, s1, s2, s3, b, out);
            else zu_segment<true, false>(env, zs, s1, s2, s3, b, out);
        }
    }
}
