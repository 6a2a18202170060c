// Halo / boundary-condition kernels, slab halo pack/unpack, and the mass/energy reduction.
#pragma once
#include "pmw_common.cuh"

namespace pmw {

// set_bc_x, periodic branch (bcs.py:35-39): interior rows only, all four variables.
__global__ void bc_x_kernel(double* s, const Layout L)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NVAR * L.nz) return;
    const int v = t / L.nz, k = t % L.nz + HS;
    double* row = s + idx(L, v, k, 0);
    const int nx = L.nx;
    row[0] = row[nx];
    row[1] = row[nx + 1];
    row[nx + HS] = row[HS];
    row[nx + HS + 1] = row[HS + 1];
}

// set_bc_x, injection branch (bcs.py:35-37,41-64): the left halo is the periodic image of the last two
// interior columns; the right halo is NOT touched (it keeps whatever the array held); on the rows of
// the jet (mask, evaluated on the host exactly as bcs.py:43-48 does) the left halo of rho*u and
// (rho*theta)' is then forced to u_in, theta_in using the halo's own rho'.  One thread per row.
__global__ void bc_x_inflow_kernel(double* s, const Layout L, const double* __restrict__ hd,
                                   const double* __restrict__ hdt, const unsigned char* __restrict__ jet,
                                   double u_in, double theta_in)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.nz) return;
    const int k = t + HS, nx = L.nx;
    double dens[2];
#pragma unroll
    for (int v = 0; v < NVAR; ++v) {
        double* row = s + idx(L, v, k, 0);
        const double a = row[nx], b = row[nx + 1];
        row[0] = a;
        row[1] = b;
        if (v == DENS) { dens[0] = a; dens[1] = b; }
    }
    if (jet[t]) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double rho = dens[j] + hd[k];
            s[idx(L, UMOM, k, j)] = rho * u_in;
            s[idx(L, RHOT, k, j)] = rho * theta_in - hdt[k];
        }
    }
}

// The same periodic wrap, six columns wide: what a fused x sweep (pmw_sweep.cuh) reads.  Array
// columns -4 .. 1 are the image of nx-4 .. nx+1, columns nx+2 .. nx+7 the image of 2 .. 7.
__global__ void bc_x6_kernel(double* s, const Layout L)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NVAR * L.nz * 6) return;
    const int j = t % 6, k = (t / 6) % L.nz + HS, v = t / (6 * L.nz);
    double* row = s + idx(L, v, k, 0);
    const int nx = L.nx;
    row[j - 4] = row[nx + j - 4];
    row[nx + HS + j] = row[HS + j];
}

// set_bc_z (bcs.py:92-148): all nx+4 columns.  True divisions, in the reference's order.
__global__ void bc_z_kernel(double* s, const Layout L, const double* __restrict__ hd)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.nx + 2 * HS) return;
    const int nz = L.nz, top = nz + HS - 1;
    const double hb = hd[HS], ht = hd[top];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        s[idx(L, WMOM, j, i)] = 0.0;
        s[idx(L, WMOM, nz + HS + j, i)] = 0.0;
        s[idx(L, UMOM, j, i)] = s[idx(L, UMOM, HS, i)] / hb * hd[j];
        s[idx(L, UMOM, nz + HS + j, i)] = s[idx(L, UMOM, top, i)] / ht * hd[nz + HS + j];
        s[idx(L, DENS, j, i)] = s[idx(L, DENS, HS, i)];
        s[idx(L, DENS, nz + HS + j, i)] = s[idx(L, DENS, top, i)];
        s[idx(L, RHOT, j, i)] = s[idx(L, RHOT, HS, i)];
        s[idx(L, RHOT, nz + HS + j, i)] = s[idx(L, RHOT, top, i)];
    }
}

// Copies the halo ring (2 rows top/bottom over all columns, 2 columns left/right over the
// interior rows) from src to dst.  Used by pmw_stage when `out` aliases `forcing`
// (step.py:122-131): the update is written to the spare buffer, which then takes over the
// role of `out`; its halo must be the one the reference's in-place array would still hold.
__global__ void copy_halo_ring_kernel(double* dst, const double* __restrict__ src, const Layout L)
{
    const int NX = L.nx + 2 * HS, NZ = L.nz + 2 * HS;
    const int per_var = 4 * NX + 4 * L.nz;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NVAR * per_var) return;
    const int v = t / per_var;
    int e = t % per_var, k, i;
    if (e < 4 * NX) {
        const int r = e / NX;
        k = (r < 2) ? r : NZ - 4 + r;
        i = e % NX;
    } else {
        e -= 4 * NX;
        const int c = e % 4;
        k = e / 4 + HS;
        i = (c < 2) ? c : NX - 4 + c;
    }
    dst[idx(L, v, k, i)] = src[idx(L, v, k, i)];
}

// Slab halo exchange messages: [4][nz][2] doubles each (set_bc_x generalised to a ring of
// slabs).  to_left = first two interior columns, to_right = last two.
__global__ void pack_halo_x_kernel(const double* __restrict__ s, const Layout L, double* to_left,
                                   double* to_right)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NVAR * L.nz * 2) return;
    const int j = t & 1, k = (t >> 1) % L.nz, v = (t >> 1) / L.nz;
    to_left[t] = s[idx(L, v, k + HS, HS + j)];
    to_right[t] = s[idx(L, v, k + HS, L.nx + j)];
}
__global__ void unpack_halo_x_kernel(double* s, const Layout L, const double* __restrict__ from_left,
                                     const double* __restrict__ from_right)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= NVAR * L.nz * 2) return;
    const int j = t & 1, k = (t >> 1) % L.nz, v = (t >> 1) / L.nz;
    s[idx(L, v, k + HS, j)] = from_left[t];               // left neighbour's last two columns
    s[idx(L, v, k + HS, L.nx + HS + j)] = from_right[t];  // right neighbour's first two
}

// ---- compute_stats (stats.py:16-33) ----------------------------------------------------------
__device__ __forceinline__ double warp_sum(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

// Energy density of one cell, stats.py:19-33: rho (u^2 + w^2) + rho cv T with T = theta (p/p0)^(R/cp) and
// p = C0 (rho theta)^gamma.  Because 1 + gamma R/cp = gamma (R = cp - cv), rho cv T = K p with the constant
// K = (cv/C0) (C0/p0)^(R/cp): the two pow() of the reference collapse into the pressure, which is evaluated
// relative to the row's hydrostatic value exactly as in the flux kernels (pow1p_gamma_m1; pow() itself
// beyond |e| > 1/8).  Agreement with the two-pow form: ~2e-16 per cell.
// the rare branch, out of line: pow() inlined costs the diagnostics kernel 40 registers
__device__ __noinline__ double stats_ie_pow(double dens_theta, double k_c0) { return k_c0 * pow(dens_theta, GAMMA); }

__device__ __forceinline__ double cell_energy(double dens, double mu, double mw, double rt, double hd, double hdt,
                                              double ihdt, double kp_row, double k_c0, double& rho_out)
{
    const double rho = dens + hd;
    rho_out = rho;
    const double e = rt * ihdt;
    double ie;
    if (fabs(e) <= 0.125) ie = fma(kp_row, pow1p_gamma_m1(e), kp_row);
    else ie = stats_ie_pow(rt + hdt, k_c0);
    return fma(fma(mu, mu, mw * mw), rcp_pos(rho), ie);  // no 1/2 on the kinetic term (stats.py:27)
}

// Pass 1: items = (row, chunk of 512 columns = one 16-byte pair per thread), grid-stride over a persistent grid; the
// four loads of the NEXT item are in flight while the current one is evaluated; row constants come with them.
// Per-thread sums, warp-shuffle tree, one partial per block: partial[2*b] = sum rho, [2*b+1] = sum(ke+ie).
// HBM-bound: 32 B per cell.  (Interior column 0 sits on a 128-byte line, pmw_common.cuh.)
struct StatsItem {
    double2 d, u, w, t;
    int k;  // row, or -1: nothing to do
};
#ifndef PMW_STATS_MINB
#define PMW_STATS_MINB 3  // resident blocks per SM = blocks of the persistent grid per SM (pmw_api.cu: stats_blocks)
#endif
__global__ void __launch_bounds__(256, PMW_STATS_MINB) stats_partial_kernel(const double* __restrict__ s, const Layout L,
                                                               const double* __restrict__ hd,
                                                               const double* __restrict__ hdt,
                                                               const double* __restrict__ ihdt,
                                                               const double* __restrict__ pcell, const double kconst,
                                                               double* partial)
{
    constexpr int CHUNK = 512;
    const int nchunks = (L.nx + CHUNK - 1) / CHUNK;
    const int nitems = L.nz * nchunks;
    const double k_c0 = kconst * C0;
    const bool vec = (L.nx & 1) == 0;
    double mass = 0.0, energy = 0.0;
    if (vec) {
        auto fetch = [&](int item) {
            StatsItem it;
            const int k = item / nchunks, i = (item - k * nchunks) * CHUNK + 2 * threadIdx.x;
            it.k = (item < nitems && i < L.nx) ? k : -1;
            if (it.k >= 0) {
                const double* p = s + idx(L, 0, k + HS, i + HS);
                it.d = *reinterpret_cast<const double2*>(p);
                it.u = *reinterpret_cast<const double2*>(p + L.vstride);
                it.w = *reinterpret_cast<const double2*>(p + 2 * L.vstride);
                it.t = *reinterpret_cast<const double2*>(p + 3 * L.vstride);
            }
            return it;
        };
        StatsItem cur = fetch(blockIdx.x);
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const StatsItem nxt = fetch(item + gridDim.x);
            if (cur.k >= 0) {
                // the row's constants: four broadcast loads that hit L1 / L2
                const double h = __ldg(hd + cur.k + HS), ht = __ldg(hdt + cur.k + HS), iht = __ldg(ihdt + cur.k + HS);
                const double kp = kconst * __ldg(pcell + cur.k + HS);
                double r0, r1;
                energy += cell_energy(cur.d.x, cur.u.x, cur.w.x, cur.t.x, h, ht, iht, kp, k_c0, r0);
                energy += cell_energy(cur.d.y, cur.u.y, cur.w.y, cur.t.y, h, ht, iht, kp, k_c0, r1);
                mass += r0 + r1;
            }
            cur = nxt;
        }
    } else {  // odd widths: one cell per thread and load
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int k = item / nchunks, i0 = (item - k * nchunks) * CHUNK;
            const double h = __ldg(hd + k + HS), ht = __ldg(hdt + k + HS), iht = __ldg(ihdt + k + HS);
            const double kp = kconst * __ldg(pcell + k + HS);
            for (int ii = i0 + threadIdx.x; ii < min(i0 + CHUNK, L.nx); ii += 256) {
                const double* q = s + idx(L, 0, k + HS, ii + HS);
                double r0;
                energy += cell_energy(q[0], q[L.vstride], q[2 * L.vstride], q[3 * L.vstride], h, ht, iht, kp, k_c0, r0);
                mass += r0;
            }
        }
    }
    __shared__ double sm[2][8];
    mass = warp_sum(mass);
    energy = warp_sum(energy);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][warp] = mass; sm[1][warp] = energy; }
    __syncthreads();
    if (warp == 0) {
        mass = lane < 8 ? sm[0][lane] : 0.0;
        energy = lane < 8 ? sm[1][lane] : 0.0;
        mass = warp_sum(mass);
        energy = warp_sum(energy);
        if (lane == 0) { partial[2 * blockIdx.x] = mass; partial[2 * blockIdx.x + 1] = energy; }
    }
}

// Pass 2: one block folds the per-block partials in a fixed order and scales by dx*dz.
__global__ void __launch_bounds__(256) stats_final_kernel(const double* __restrict__ partial, int nblocks,
                                                          double cell_area, double* out2)
{
    double mass = 0.0, energy = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
        mass += partial[2 * b];
        energy += partial[2 * b + 1];
    }
    __shared__ double sm[2][8];
    mass = warp_sum(mass);
    energy = warp_sum(energy);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sm[0][warp] = mass; sm[1][warp] = energy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0, e = 0.0;
        for (int wv = 0; wv < 8; ++wv) { m += sm[0][wv]; e += sm[1][wv]; }
        out2[0] = m * cell_area;
        out2[1] = e * cell_area;
    }
}

// compute_solution_variables (stats.py:38-69): dense [4][nz][nx] output.
__global__ void solution_variables_kernel(const double* __restrict__ s, const Layout L,
                                          const double* __restrict__ hd, const double* __restrict__ hdt,
                                          double* out)
{
    const long long n = (long long)L.nx * L.nz;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int k = (int)(c / L.nx), i = (int)(c % L.nx);
    const double d = s[idx(L, DENS, k + HS, i + HS)];
    const double rho = hd[k + HS] + d;
    out[c] = d;
    out[n + c] = s[idx(L, UMOM, k + HS, i + HS)] / rho;
    out[2 * n + c] = s[idx(L, WMOM, k + HS, i + HS)] / rho;
    out[3 * n + c] = (s[idx(L, RHOT, k + HS, i + HS)] + hdt[k + HS]) / rho - hdt[k + HS] / hd[k + HS];
}


// FP64 issue-rate probe (pmw_fp64_peak): NCH independent chains of dependent DFMAs per thread.
template <int NCH>
__global__ void dfma_probe_kernel(double* out, int iters, double a, double b)
{
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < NCH; ++c) x[c] = fma(x[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) s += x[c];
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace pmw
