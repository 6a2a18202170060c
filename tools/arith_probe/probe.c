/*
 * Arithmetic-sensitivity probe (development tool, not product, not oracle).
 * Re-includes the C oracle with GPU-like arithmetic (FMA contraction, division
 * by reciprocal-multiply, background-relative fast pow) so that the parity
 * margin of those choices can be measured on the CPU before spending GPU time.
 * Build variants: see run_probe.py.
 */
#include <math.h>
#include <stddef.h>

static const double G_ = 1.40027894002789400278940027894;

/* (1+e)^gamma - 1 for |e| <= 1/8 : 2*atanh(e/(2+e)) series, then expm1 Taylor */
static inline double pow1p_gamma_m1(double e)
{
    const double s = e / (2.0 + e);
    const double s2 = s * s;
    double q = 1.0 / 15;
    q = fma(q, s2, 1.0 / 13);
    q = fma(q, s2, 1.0 / 11);
    q = fma(q, s2, 1.0 / 9);
    q = fma(q, s2, 1.0 / 7);
    q = fma(q, s2, 1.0 / 5);
    q = fma(q, s2, 1.0 / 3);
    const double twos = s + s;
    const double L = fma(twos * s2, q, twos); /* log1p(e) */
    const double y = G_ * L;
    double r = 1.0 / 479001600.0; /* 1/12! */
    r = fma(r, y, 1.0 / 39916800.0);
    r = fma(r, y, 1.0 / 3628800.0);
    r = fma(r, y, 1.0 / 362880.0);
    r = fma(r, y, 1.0 / 40320.0);
    r = fma(r, y, 1.0 / 5040.0);
    r = fma(r, y, 1.0 / 720.0);
    r = fma(r, y, 1.0 / 120.0);
    r = fma(r, y, 1.0 / 24.0);
    r = fma(r, y, 1.0 / 6.0);
    r = fma(r, y, 0.5);
    return fma(y * y, r, y); /* expm1(y) */
}

#if defined(VAR_RECIP) || defined(VAR_FASTPOW)
static inline double div_rcp(double a, double b) { return a * (1.0 / b); }
#define PMWO_DIV(a, b) div_rcp((a), (b))
#endif

#ifdef VAR_FASTPOW
/* per-row background pressure computed the way the host would (libm pow) */
#define PMWO_PRESSURE_X(c, rt, row)                                                     \
    ({ const double H_ = (c)->hy_dens_theta_cell[(row)];                                 \
       const double PH_ = 27.5629410929725921310572974482 * pow(H_, G_);                 \
       const double e_ = ((rt) - H_) * (1.0 / H_);                                       \
       fma(PH_, pow1p_gamma_m1(e_), PH_); })
#define PMWO_PRESSURE_Z(c, rt, k)                                                       \
    ({ const double H_ = (c)->hy_dens_theta_int[(k)];                                    \
       const double e_ = ((rt) - H_) * (1.0 / H_);                                       \
       (c)->hy_pressure_int[(k)] * pow1p_gamma_m1(e_); })
#endif

#include "../../oracle/c/pmw_oracle.c"

double probe_pow1p(double e) { return pow1p_gamma_m1(e); }
