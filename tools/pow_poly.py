"""Coefficients of the pressure polynomial (csrc/pmw_common.cuh: kPow1p): f(e) = (1+e)^gamma - 1 = e * g(e),
g interpolated at Chebyshev nodes on [-R, R] by a polynomial of degree N (so f has degree N+1).
usage: python tools/pow_poly.py [N [R [EXPONENT]]]   -- prints the double coefficients and the error vs mpmath.
EXPONENT defaults to gamma; "stats" selects 1 + gamma*R_d/c_p, the exponent of rho*theta in the internal
energy of compute_stats (csrc/pmw_aux.cuh: kPowStats)."""
import sys
import mpmath as mp
import numpy as np
mp.mp.dps = 60
GAMMA = mp.mpf("1.40027894002789400278940027894")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 11
R = mp.mpf(sys.argv[2]) if len(sys.argv) > 2 else mp.mpf(1) / 8
if len(sys.argv) > 3:
    GAMMA = 1 + GAMMA * mp.mpf(287) / mp.mpf(1004) if sys.argv[3] == "stats" else mp.mpf(sys.argv[3])


def g(e):
    return mp.expm1(GAMMA * mp.log1p(e)) / e if abs(e) > mp.mpf(10) ** -40 else GAMMA


nodes = [R * mp.cos(mp.pi * (2 * k + 1) / (2 * (N + 1))) for k in range(N + 1)]
A = mp.matrix(N + 1, N + 1)
b = mp.matrix(N + 1, 1)
for r, x in enumerate(nodes):
    for c in range(N + 1):
        A[r, c] = x ** c
    b[r] = g(x)
coef = mp.lu_solve(A, b)
cd = [float(c) for c in coef]          # rounded to double: what the kernel holds
print(f"degree of g: {N}  (f = e*g has degree {N + 1}), |e| <= {float(R)}")
for i, c in enumerate(cd):
    print(f"    {float.hex(c)},  // c{i}")


def kernel_eval(e):  # the kernel's scheme: two Horner chains in e^2 (even / odd coefficients), then fma(b,e,a)*e
    e = np.float64(e)
    e2 = e * e
    ev = [cd[i] for i in range(0, N + 1, 2)]
    od = [cd[i] for i in range(1, N + 1, 2)]
    a = np.float64(ev[-1])
    for c in reversed(ev[:-1]):
        a = a * e2 + np.float64(c)   # (numpy has no fma: one extra rounding per step, an upper bound)
    bb = np.float64(od[-1])
    for c in reversed(od[:-1]):
        bb = bb * e2 + np.float64(c)
    return (bb * e + a) * e


worst_exact = worst_eval = mp.mpf(0)
for k in range(-4000, 4001):
    e = R * k / 4000
    ef = float(e)
    exact = mp.expm1(GAMMA * mp.log1p(mp.mpf(ef)))
    poly = sum(mp.mpf(cd[i]) * mp.mpf(ef) ** (i + 1) for i in range(N + 1))
    worst_exact = max(worst_exact, abs(poly - exact))
    worst_eval = max(worst_eval, abs(mp.mpf(float(kernel_eval(ef))) - exact))
print(f"max |poly - f| in exact arithmetic: {float(worst_exact):.2e};  evaluated in double (no fma): {float(worst_eval):.2e}")
