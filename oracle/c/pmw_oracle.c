/*
 * Plain-C restatement of the PyMiniWeather hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Same algorithm, same operation order and the same materialised intermediates
 * (interface fluxes, tendencies) as the reference's NumPy backend; loops are
 * OpenMP-parallel over rows so that it can serve as the multi-threaded CPU
 * baseline for bench.py and as a fast checker at mid-size grids.  Compile with
 * -ffp-contract=off (see Makefile): with contraction disabled every result is
 * bit-identical to the NumPy oracle except through pow(), where libm and
 * NumPy's SIMD pow may differ by an ulp.
 *
 * Reference locations (relative to the reference repo root):
 *   set_bc_x            pyminiweather/ics/bcs.py:35-39 (periodic), :37,41-64 (injection inflow)
 *   set_bc_z            pyminiweather/ics/bcs.py:92-148
 *   interpolate_x/z     pyminiweather/solve/interpolate.py:33-43, 69-79
 *                       (stencil weights: pyminiweather/data/fields.py:94-97)
 *   compute_flux_x/z    pyminiweather/solve/interpolate.py:95-129, 144-186
 *   compute_tend_x/z    pyminiweather/solve/interpolate.py:208-215, 238-250
 *   discrete_step       pyminiweather/solve/step.py:63-82
 *   evolve              pyminiweather/solve/step.py:105-143
 *   compute_stats       pyminiweather/post/stats.py:16-33
 *   constants           pyminiweather/data/constants.py:4-27
 *
 * State layout: [4][nz+4][nx+4] doubles, C order; DENS=0 UMOM=1 WMOM=2 RHOT=3.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef PMWO_POW
#define PMWO_POW(x, y) pow((x), (y))
#endif
#ifndef PMWO_DIV
#define PMWO_DIV(a, b) ((a) / (b))
#endif
/* Hooks used only by tools/arith_probe (which re-includes this file with other
 * definitions to study rounding sensitivity); the defaults ARE the reference. */
#ifndef PMWO_PRESSURE_X
#define PMWO_PRESSURE_X(c, rt, row) (C0 * PMWO_POW((rt), GAMMA_))
#endif
#ifndef PMWO_PRESSURE_Z
#define PMWO_PRESSURE_Z(c, rt, k) (C0 * PMWO_POW((rt), GAMMA_) - (c)->hy_pressure_int[(k)])
#endif

#define HS 2
enum { DENS = 0, UMOM = 1, WMOM = 2, RHOT = 3 };
enum { DIR_X = 1, DIR_Z = 2 };

static const double HV_BETA = 0.05;
static const double P0 = 1.0e5;
static const double C0 = 27.5629410929725921310572974482;
static const double GAMMA_ = 1.40027894002789400278940027894;
static const double GRAV = 9.8;
static const double CP = 1004.0;
static const double CV = 717.0;
static const double RD = 287.0;

typedef struct {
    int nx, nz;
    double dx, dz, dt; /* dt = FULL step: hv_coeff uses it (interpolate.py:99-101) */
    const double *hy_dens_cell;       /* [nz+4] */
    const double *hy_dens_theta_cell; /* [nz+4] */
    const double *hy_dens_int;        /* [nz+1] */
    const double *hy_dens_theta_int;  /* [nz+1] */
    const double *hy_pressure_int;    /* [nz+1] */
    double *flux;                     /* scratch [4][nz+1][nx+1] */
    double *tend;                     /* scratch [4][nz][nx]     */
    const double *source_w;           /* [nz][nx] or NULL: gravity-wave forcing on rho*w in every
                                         stage (source.py:43-50, step.py:78) */
    const unsigned char *inflow_rows; /* [nz+4] or NULL: 1 on the array rows of the injection jet
                                         (bcs.py:43-48, evaluated by the caller); non-NULL selects
                                         the injection branch of set_bc_x */
} pmwo_case;

#define S(s, v, k, i) (s)[((size_t)(v) * NZ + (size_t)(k)) * NX + (size_t)(i)]

/* bcs.py:35-39; injection: the right halo is left alone (:37) and the jet rows of the left halo are
 * forced to u = 50 m/s, theta = 298 K (:50-64) */
void pmwo_set_bc_x(const pmwo_case *c, double *s)
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    for (int v = 0; v < 4; ++v)
        for (int k = HS; k < nz + HS; ++k) {
            S(s, v, k, 0) = S(s, v, k, nx);
            S(s, v, k, 1) = S(s, v, k, nx + 1);
            if (!c->inflow_rows) {
                S(s, v, k, nx + HS) = S(s, v, k, HS);
                S(s, v, k, nx + HS + 1) = S(s, v, k, HS + 1);
            }
        }
    if (c->inflow_rows)
        for (int k = HS; k < nz + HS; ++k) {
            if (!c->inflow_rows[k]) continue;
            for (int i = 0; i < 2; ++i) {
                const double rho = S(s, DENS, k, i) + c->hy_dens_cell[k];
                S(s, UMOM, k, i) = rho * 50.0;
                S(s, RHOT, k, i) = rho * 298.0 - c->hy_dens_theta_cell[k];
            }
        }
}

/* bcs.py:92-148 */
void pmwo_set_bc_z(const pmwo_case *c, double *s)
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    const double *hd = c->hy_dens_cell;
    const int top = nz + HS - 1;
    for (size_t i = 0; i < NX; ++i) {
        S(s, WMOM, 0, i) = 0.0;
        S(s, WMOM, 1, i) = 0.0;
        S(s, WMOM, nz + HS, i) = 0.0;
        S(s, WMOM, nz + HS + 1, i) = 0.0;
        S(s, UMOM, 0, i) = S(s, UMOM, HS, i) / hd[HS] * hd[0];
        S(s, UMOM, 1, i) = S(s, UMOM, HS, i) / hd[HS] * hd[1];
        S(s, UMOM, nz + HS, i) = S(s, UMOM, top, i) / hd[top] * hd[nz + HS];
        S(s, UMOM, nz + HS + 1, i) = S(s, UMOM, top, i) / hd[top] * hd[nz + HS + 1];
        S(s, DENS, 0, i) = S(s, DENS, HS, i);
        S(s, DENS, 1, i) = S(s, DENS, HS, i);
        S(s, DENS, nz + HS, i) = S(s, DENS, top, i);
        S(s, DENS, nz + HS + 1, i) = S(s, DENS, top, i);
        S(s, RHOT, 0, i) = S(s, RHOT, HS, i);
        S(s, RHOT, 1, i) = S(s, RHOT, HS, i);
        S(s, RHOT, nz + HS, i) = S(s, RHOT, top, i);
        S(s, RHOT, nz + HS + 1, i) = S(s, RHOT, top, i);
    }
}

/* the two 4-tap correlations, accumulated left to right as _correlateND does */
static inline void stencil4(double a, double b, double cc, double d, double *val, double *d3)
{
    const double c0 = -1.0 / 12, c1 = 7.0 / 12;
    *val = ((c0 * a + c1 * b) + c1 * cc) + c0 * d;
    *d3 = ((-1.0 * a + 3.0 * b) + -3.0 * cc) + 1.0 * d;
}

#define F(v, k, i) c->flux[((size_t)(v) * (nz + 1) + (size_t)(k)) * (nx + 1) + (size_t)(i)]
#define T(v, k, i) c->tend[((size_t)(v) * nz + (size_t)(k)) * nx + (size_t)(i)]

/* interpolate_x + compute_flux_x + compute_tend_x */
static void tend_x(const pmwo_case *c, const double *s)
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    const double hv = -HV_BETA * c->dx / (16 * c->dt);
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k) {
        const double hd = c->hy_dens_cell[k + HS], hdt = c->hy_dens_theta_cell[k + HS];
        for (int i = 0; i <= nx; ++i) {
            double val[4], d3[4];
            for (int v = 0; v < 4; ++v)
                stencil4(S(s, v, k + HS, i), S(s, v, k + HS, i + 1), S(s, v, k + HS, i + 2),
                         S(s, v, k + HS, i + 3), &val[v], &d3[v]);
            const double rho = val[DENS] + hd;
            const double u = PMWO_DIV(val[UMOM], rho);
            const double w = PMWO_DIV(val[WMOM], rho);
            const double t = PMWO_DIV(val[RHOT] + hdt, rho);
            const double p = PMWO_PRESSURE_X(c, rho * t, k + HS);
            F(DENS, k, i) = rho * u - hv * d3[DENS];
            F(UMOM, k, i) = rho * (u * u) + p - hv * d3[UMOM];
            F(WMOM, k, i) = rho * u * w - hv * d3[WMOM];
            F(RHOT, k, i) = rho * u * t - hv * d3[RHOT];
        }
        for (int v = 0; v < 4; ++v)
            for (int i = 0; i < nx; ++i)
                T(v, k, i) = PMWO_DIV(-(F(v, k, i + 1) - F(v, k, i)), c->dx);
    }
}

/* interpolate_z + compute_flux_z + compute_tend_z */
static void tend_z(const pmwo_case *c, const double *s)
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    const double hv = -HV_BETA * c->dz / (16 * c->dt);
#pragma omp parallel for schedule(static)
    for (int k = 0; k <= nz; ++k) {
        const double hd = c->hy_dens_int[k], hdt = c->hy_dens_theta_int[k];
        const int wall = (k == 0 || k == nz);
        for (int i = 0; i < nx; ++i) {
            double val[4], d3[4];
            for (int v = 0; v < 4; ++v)
                stencil4(S(s, v, k, i + HS), S(s, v, k + 1, i + HS), S(s, v, k + 2, i + HS),
                         S(s, v, k + 3, i + HS), &val[v], &d3[v]);
            const double rho = val[DENS] + hd;
            const double u = PMWO_DIV(val[UMOM], rho);
            double w = PMWO_DIV(val[WMOM], rho);
            const double t = PMWO_DIV(val[RHOT] + hdt, rho);
            const double p = PMWO_PRESSURE_Z(c, rho * t, k);
            if (wall) { w = 0.0; d3[DENS] = 0.0; }
            F(DENS, k, i) = rho * w - hv * d3[DENS];
            F(UMOM, k, i) = rho * w * u - hv * d3[UMOM];
            F(WMOM, k, i) = rho * (w * w) + p - hv * d3[WMOM];
            F(RHOT, k, i) = rho * w * t - hv * d3[RHOT];
        }
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int v = 0; v < 4; ++v)
            for (int i = 0; i < nx; ++i) {
                double td = PMWO_DIV(-(F(v, k + 1, i) - F(v, k, i)), c->dz);
                if (v == WMOM) td -= S(s, DENS, k + HS, i + HS) * GRAV;
                T(v, k, i) = td;
            }
}

/* step.py:63-82 */
void pmwo_discrete_step(const pmwo_case *c, const double *init, double *forcing, double *out,
                        double dt_stage, int direction)
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    if (direction == DIR_X) {
        pmwo_set_bc_x(c, forcing);
        tend_x(c, forcing);
    } else {
        pmwo_set_bc_z(c, forcing);
        tend_z(c, forcing);
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nz; ++k)
        for (int v = 0; v < 4; ++v)
            for (int i = 0; i < nx; ++i)
            {
                double td = T(v, k, i);
                if (v == WMOM && c->source_w) td += c->source_w[(size_t)k * nx + i];
                S(out, v, k + HS, i + HS) = S(init, v, k + HS, i + HS) + dt_stage * td;
            }
}

/* step.py:105-143; *reverse is the direction flag (module global in the reference) */
void pmwo_evolve(const pmwo_case *c, double *state, double *state_tmp, double dt, int nsteps,
                 int *reverse)
{
    for (int n = 0; n < nsteps; ++n) {
        const int dirs[2] = { *reverse ? DIR_X : DIR_Z, *reverse ? DIR_Z : DIR_X };
        for (int d = 0; d < 2; ++d) {
            pmwo_discrete_step(c, state, state, state_tmp, dt / 3, dirs[d]);
            pmwo_discrete_step(c, state, state_tmp, state_tmp, dt / 2, dirs[d]);
            pmwo_discrete_step(c, state, state_tmp, state, dt / 1, dirs[d]);
        }
        *reverse = !*reverse;
    }
}

/* stats.py:16-33; plain row-wise summation (NumPy sums pairwise: agreement ~1e-15) */
void pmwo_stats(const pmwo_case *c, const double *s, double out[2])
{
    const int nx = c->nx, nz = c->nz;
    const size_t NX = nx + 2 * HS, NZ = nz + 2 * HS;
    double mass = 0.0, energy = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : mass, energy)
    for (int k = 0; k < nz; ++k) {
        double m = 0.0, e = 0.0;
        for (int i = 0; i < nx; ++i) {
            const double rho = S(s, DENS, k + HS, i + HS) + c->hy_dens_cell[k + HS];
            const double u = S(s, UMOM, k + HS, i + HS) / rho;
            const double w = S(s, WMOM, k + HS, i + HS) / rho;
            const double th = (S(s, RHOT, k + HS, i + HS) + c->hy_dens_theta_cell[k + HS]) / rho;
            const double p = C0 * pow(rho * th, GAMMA_);
            const double t = th / pow(P0 / p, RD / CP);
            m += rho;
            e += rho * (u * u + w * w) + rho * CV * t;
        }
        mass += m;
        energy += e;
    }
    out[0] = mass * c->dx * c->dz;
    out[1] = energy * c->dx * c->dz;
}

size_t pmwo_flux_len(int nx, int nz) { return (size_t)4 * (nz + 1) * (nx + 1); }
size_t pmwo_tend_len(int nx, int nz) { return (size_t)4 * nz * nx; }

/* Thread control for bench.py's CPU arms: launchers such as torchrun export OMP_NUM_THREADS=1, so the
 * baseline sets its team size explicitly and reports what the runtime actually granted. */
#ifdef _OPENMP
#include <omp.h>
int pmwo_set_threads(int n)
{
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
}
int pmwo_team_size(void)
{
    int n = 1;
#pragma omp parallel
    {
#pragma omp master
        n = omp_get_num_threads();
    }
    return n;
}
#else
int pmwo_set_threads(int n) { (void)n; return 1; }
int pmwo_team_size(void) { return 1; }
#endif
