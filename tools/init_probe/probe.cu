// Host-side run of the device init code (csrc/pmw_init.cuh: ic_cell is __host__ __device__), so that
// its logic can be compared with the NumPy init without a GPU.  Build: see run_probe.py.
#include <cstdio>
#include <cstdlib>
#include "../../pyminiweather_b200/csrc/pmw_init.cuh"
using namespace pmw;

extern "C" void probe_init(const IcSpec* s, int NX, int NZ, const double* xa, const double* za, double* out /*[4][NZ][NX]*/)
{
    for (int k = 0; k < NZ; ++k)
        for (int i = 0; i < NX; ++i) {
            double su, st;
            ic_cell(*s, xa[i], za[k], su, st);
            const size_t n = (size_t)NX * NZ, c = (size_t)k * NX + i;
            out[c] = 0.0;
            out[n + c] = su;
            out[2 * n + c] = 0.0;
            out[3 * n + c] = st;
        }
}
