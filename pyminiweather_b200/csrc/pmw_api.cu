// C ABI of libpmw.so (see include/pmw.h): context, data movement, kernel dispatch.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/pmw.h"
#include "pmw_aux.cuh"
#include "pmw_direct.cuh"
#include "pmw_tma.cuh"
#include "pmw_sweep.cuh"
#include "pmw_zpipe.cuh"
#include "pmw_unfused.cuh"
#include "pmw_init.cuh"

using namespace pmw;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail(PMW_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, \
                        __LINE__);                                                                \
    } while (0)

#define NEED(cond, ...)                                 \
    do {                                                \
        if (!(cond)) return fail(PMW_EINVAL, __VA_ARGS__); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

struct TmapKey {
    int buf, bw, bh;
    CUtensorMap map;
};

// z tiles: z_cfg = NP, passes of 4 interface rows; a tile owns 4*NP-1 cell rows x 64 columns
static const int kZPassesMin = 1, kZPassesMax = 8;

struct pmw_ctx {
    pmw_params p;
    Layout L;
    size_t buf_doubles;
    double* alloc[3];
    double* base[3];
    int l2p[2];  // logical (STATE, TMP) -> physical buffer
    int spare;
    bool xhalo_valid[3];  // x halo columns hold the periodic image of the interior
    bool xhalo6_valid[3];  // ... and so do the four further columns a fused x sweep reads (6-wide image)
    // fused sweeps (pmw_sweep.cuh): 1 = pmw_evolve runs one kernel per directional sweep
    int fuse, keep_tmp, sweep_lz, sweep_xp, sweep_zt, sweep_z3, dyn_items;
    long long push_all_cells;  // slab ring: from this many cells per slab on, every CTA of an x sweep takes part in the halo push
    double* hydro_blob;
    double* src_w;  // gravity-wave forcing field or nullptr
    unsigned char* jet_rows;  // injection: [nz] mask of the inflow rows, or nullptr (periodic x)
    double jet_u, jet_theta;
    Hydro hy;
    bool hydro_set;
    bool hydro_consistent;  // pressure_int == C0 * dens_theta_int^gamma (else: pow_mode falls back to libdevice)
    cudaStream_t stream;
    int reverse;
    // tuning
    int x_tr, x_p, z_cfg, pdl, peer_dbg;
    int l2_hints;  // decimal digits: forcing(S1) init out | forcing(S2,S3): see pmw_set_tuning
    // tensor maps
    EncodeTiledFn encode;
    std::vector<TmapKey> tmaps;
    // stats scratch
    double* stats_partial;
    double* stats_out;
    int stats_blocks;
    // slab ring over peer memory (pmw_connect_peers)
    bool peers;
    double* nbr_base[2][3];          // [left|right][physical buffer]: neighbour's base pointers
    unsigned long long* nbr_flags[2];
    unsigned long long* flags;       // mine: [0] left neighbour's epoch, [1] right's, [2] watchdog
    unsigned int* edge_counters;     // last-arriver counter of the push CTAs
    unsigned long long* xitem_counter;  // work counter of the x sweeps (never reset, see SweepArgs::item_base)
    unsigned long long xitem_base;
    unsigned long long epoch;        // number of x stages run since pmw_connect_peers
    std::vector<void*> ipc_opened;
    // chunked sweeps: the three stages of a sweep only couple cells along the sweep direction, so
    // bands of columns (z sweep) / rows (x sweep) run as independent kernel chains on their own
    // streams and the tail of one kernel overlaps the head of the next band's kernel
    int chunks;
    cudaStream_t cstream[4];
    cudaEvent_t ev_fork, ev_join[4];
    cudaStream_t launch_stream;  // stream the stage launch helpers use
    int cur_chunk, cur_nchunks;
    // lazy state_tmp (keep_tmp == 1): the last sweep of a pmw_evolve call did NOT write the reference's stage-2 array;
    // ensure_tmp re-runs that sweep with the store switched on when somebody asks for it.  Valid while the sweep's
    // input (the spare buffer) and output (the state buffer) are what they were: every entry point that touches a
    // buffer materialises or drops it first (BIND / BIND_KEEP).
    bool tmp_pending;
    int tp_dir, tp_src, tp_out;
    double tp_dt;
    // pmw_evolve_host: copy streams (H2D, D2H) and one event per band and direction
    cudaStream_t hs_stream[2];
    std::vector<cudaEvent_t> hs_ev;
    cudaEvent_t hs_start;
    // bookkeeping
    long long launches;
    bool timing;
    std::vector<cudaEvent_t> ev;  // pairs
    size_t ev_used;
};

static int bind(pmw_ctx* c)
{
    if (!c) return fail(PMW_EINVAL, "null context");
    CU_TRY(cudaSetDevice(c->p.device));
    return PMW_OK;
}
static int ensure_tmp(pmw_ctx* c);
// BIND: entry points that may read state_tmp or write any buffer -- a pending state_tmp is materialised first.
// BIND_KEEP: entry points that leave the buffers alone or only read the state (they call ensure_tmp themselves
// when handed PMW_BUF_TMP).
#define BIND_KEEP(c)              \
    do {                          \
        int rc_ = bind(c);        \
        if (rc_ != PMW_OK) return rc_; \
    } while (0)
#define BIND(c)                                  \
    do {                                         \
        int rc_ = bind(c);                       \
        if (rc_ == PMW_OK) rc_ = ensure_tmp(c);  \
        if (rc_ != PMW_OK) return rc_;           \
    } while (0)
#define ENSURE_TMP_IF(c, cond)                   \
    do {                                         \
        if (cond) {                              \
            int rc_ = ensure_tmp(c);             \
            if (rc_ != PMW_OK) return rc_;       \
        }                                        \
    } while (0)

extern "C" const char* pmw_last_error(void) { return g_err; }
extern "C" int pmw_version(void) { return 100; }

static int round_up(int x, int m) { return (x + m - 1) / m * m; }

// number of 64-interface passes per x tile (tile = 64p-2 cells per row): two passes measured best
// on B200 (tools/sweep_tiles.py); one pass for grids narrower than a two-pass tile
static int pick_x_passes(int nx) { return nx > 62 ? 2 : 1; }

extern "C" int pmw_create(const pmw_params* params, pmw_ctx** out)
{
    NEED(params && out, "pmw_create: null argument");
    NEED(params->hs == HS, "pmw_create: hs must be 2 (got %d)", params->hs);
    NEED(params->nx >= 4 && params->nz >= 4, "pmw_create: nx and nz must be >= 4 (got %d x %d)", params->nx,
         params->nz);
    NEED(params->dx > 0 && params->dz > 0 && params->dt > 0, "pmw_create: dx, dz, dt must be positive");
    NEED(params->variant == PMW_VARIANT_DIRECT || params->variant == PMW_VARIANT_TMA,
         "pmw_create: unknown variant %d", params->variant);
    NEED(params->pow_mode == PMW_POW_LIBDEVICE || params->pow_mode == PMW_POW_BACKGROUND,
         "pmw_create: unknown pow_mode %d", params->pow_mode);
    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));
    NEED(params->device >= 0 && params->device < ndev, "pmw_create: device %d out of range (%d visible)",
         params->device, ndev);
    CU_TRY(cudaSetDevice(params->device));

    pmw_ctx* c = new (std::nothrow) pmw_ctx();
    if (!c) return fail(PMW_EINVAL, "out of host memory");
    c->p = *params;
    c->L.nx = params->nx;
    c->L.nz = params->nz;
    // room for the 6-wide x halo of the fused sweeps: array columns -4 .. nx+7
    c->L.pitch = round_up(LPAD + params->nx + HS + SWEEP_HALO, 16);
    c->L.vstride = (long long)c->L.pitch * (params->nz + 2 * HS);
    c->buf_doubles = (size_t)NVAR * c->L.vstride + 32;
    for (int b = 0; b < 3; ++b) c->alloc[b] = nullptr;
    c->hydro_blob = nullptr;
    c->src_w = nullptr;
    c->jet_rows = nullptr;
    c->jet_u = c->jet_theta = 0.0;
    c->stats_partial = c->stats_out = nullptr;
    c->flags = nullptr;
    c->edge_counters = nullptr;
    c->xitem_counter = nullptr;
    c->xitem_base = 0;
    // two bands per sweep measured best at 2048x1024 (tools/chunk_probe.py); small grids stay whole
    c->chunks = ((long long)params->nx * params->nz >= (1ll << 20)) ? 2 : 1;
    for (int k = 0; k < 4; ++k) { c->cstream[k] = nullptr; c->ev_join[k] = nullptr; }
    c->ev_fork = nullptr;
    c->hs_stream[0] = c->hs_stream[1] = nullptr;
    c->hs_start = nullptr;
    c->tmp_pending = false;
    c->launch_stream = nullptr;
    c->cur_chunk = 0;
    c->cur_nchunks = 1;
    c->peers = false;
    c->epoch = 0;
    for (int b = 0; b < 3; ++b) {
        cudaError_t e = cudaMalloc(&c->alloc[b], c->buf_doubles * sizeof(double));
        if (e != cudaSuccess) {
            pmw_destroy(c);
            return fail(PMW_ECUDA, "cudaMalloc of %zu bytes failed: %s", c->buf_doubles * sizeof(double),
                        cudaGetErrorString(e));
        }
        cudaMemset(c->alloc[b], 0, c->buf_doubles * sizeof(double));
        c->base[b] = c->alloc[b] + LPAD;
        c->xhalo_valid[b] = false;
        c->xhalo6_valid[b] = false;
    }
    c->fuse = 1;
    c->keep_tmp = 1;
    c->sweep_lz = 0;  // 0 = choose from the grid (pick_sweep_lz)
    c->sweep_xp = 2;
    c->push_all_cells = 0;  // never (measured worse: every CTA then waits for its peer stores, profiles/r2ae)
    c->dyn_items = 1;  // x sweeps of a slab ring draw their items from a counter (0 never, 2 always)
    c->sweep_z3 = 0;  // z sweeps: 1 = stage-pipelined CTA of three warps (pmw_zpipe.cuh), 0 = one warp per strip (sweep_z)
    c->sweep_zt = -1;  // z sweeps: 0 = streaming kernel, 1 = transposing x-style kernel, -1 = by grid size (kZtMaxCells)
    c->l2p[PMW_BUF_STATE] = 0;
    c->l2p[PMW_BUF_TMP] = 1;
    c->spare = 2;
    c->hydro_set = false;
    c->hydro_consistent = true;
    c->stream = 0;
    c->reverse = 0;
    c->x_tr = 4;
    c->x_p = pick_x_passes(params->nx);
    c->z_cfg = 3;
    c->pdl = 1;
    c->peer_dbg = 0;
    c->l2_hints = 1100;  // forcing tiles evict_first, everything else normal (tools/l2_probe.py)
    c->encode = nullptr;
    c->launches = 0;
    c->timing = false;
    c->ev_used = 0;
    c->stats_blocks = 148 * PMW_STATS_MINB;  // persistent: the resident blocks of stats_partial_kernel
    {
        // 8 profile tables + the packed interface table Hydro::int_pack (32-byte aligned, hence the slack)
        const size_t nhy = (size_t)4 * (params->nz + 4) + (size_t)4 * (params->nz + 1) + 4 +
                           (size_t)4 * (params->nz + 1 + 2 * HY_PACK_PAD) + (size_t)4 * (params->nz + 4);
        if (cudaMalloc(&c->hydro_blob, nhy * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&c->stats_partial, (size_t)2 * c->stats_blocks * sizeof(double)) != cudaSuccess ||
            cudaMalloc(&c->stats_out, 2 * sizeof(double)) != cudaSuccess ||
            // slab ring: the flag words, and behind them (one allocation = one IPC handle) the staging area the
            // neighbours push their edge columns into (pmw_sweep.cuh: halo_stage)
            cudaMalloc(&c->flags, halo_stage_bytes(params->nz)) != cudaSuccess ||
            cudaMemset(c->flags, 0, halo_stage_bytes(params->nz)) != cudaSuccess ||
            cudaMalloc(&c->xitem_counter, sizeof(unsigned long long)) != cudaSuccess ||
            cudaMemset(c->xitem_counter, 0, sizeof(unsigned long long)) != cudaSuccess ||
            cudaMalloc(&c->edge_counters, 2 * sizeof(unsigned int)) != cudaSuccess ||
            cudaMemset(c->edge_counters, 0, 2 * sizeof(unsigned int)) != cudaSuccess) {
            pmw_destroy(c);
            return fail(PMW_ECUDA, "cudaMalloc of auxiliary buffers failed");
        }
    }
    if (params->variant == PMW_VARIANT_TMA) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            pmw_destroy(c);
            return fail(PMW_ECUDA, "cuTensorMapEncodeTiled not available from the driver");
        }
        c->encode = (EncodeTiledFn)fn;
    }
    *out = c;
    return PMW_OK;
}

extern "C" int pmw_destroy(pmw_ctx* c)
{
    if (!c) return PMW_OK;
    cudaSetDevice(c->p.device);
    for (int b = 0; b < 3; ++b)
        if (c->alloc[b]) cudaFree(c->alloc[b]);
    if (c->hydro_blob) cudaFree(c->hydro_blob);
    if (c->src_w) cudaFree(c->src_w);
    if (c->jet_rows) cudaFree(c->jet_rows);
    if (c->stats_partial) cudaFree(c->stats_partial);
    if (c->stats_out) cudaFree(c->stats_out);
    for (int k = 0; k < 4; ++k) {
        if (c->cstream[k]) { cudaStreamSynchronize(c->cstream[k]); cudaStreamDestroy(c->cstream[k]); }
        if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]);
    }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (int k = 0; k < 2; ++k)
        if (c->hs_stream[k]) { cudaStreamSynchronize(c->hs_stream[k]); cudaStreamDestroy(c->hs_stream[k]); }
    for (cudaEvent_t e : c->hs_ev) cudaEventDestroy(e);
    if (c->hs_start) cudaEventDestroy(c->hs_start);
    for (void* p : c->ipc_opened) cudaIpcCloseMemHandle(p);
    if (c->flags) cudaFree(c->flags);
    if (c->edge_counters) cudaFree(c->edge_counters);
    if (c->xitem_counter) cudaFree(c->xitem_counter);
    for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
    delete c;
    return PMW_OK;
}

extern "C" int pmw_set_stream(pmw_ctx* c, void* s)
{
    BIND(c);
    c->stream = (cudaStream_t)s;
    return PMW_OK;
}

// Slab ring: a wait on a neighbour's epoch flag is bounded (~2 s, wait_epoch in pmw_tma.cuh) so that a lost
// peer cannot hang the GPU; the kernel then carries on with stale halo columns and raises flags[2].  Every
// synchronising entry point turns that into an error -- results computed after a time-out are invalid.
static int check_watchdog(pmw_ctx* c)
{
    if (!c->peers) return PMW_OK;
    unsigned long long f = 0;
    CU_TRY(cudaMemcpyAsync(&f, c->flags + 2, sizeof(f), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (f != 0)
        return fail(PMW_ECUDA, "slab ring: a neighbour's halo columns did not arrive within the watchdog time; "
                               "the state of this context is invalid");
    return PMW_OK;
}

extern "C" int pmw_synchronize(pmw_ctx* c)
{
    BIND_KEEP(c);
    CU_TRY(cudaStreamSynchronize(c->stream));
    return check_watchdog(c);
}

extern "C" int pmw_set_tuning(pmw_ctx* c, const char* key, int value)
{
    BIND(c);
    NEED(key, "pmw_set_tuning: null key");
    if (!strcmp(key, "x_tr")) {
        NEED(value == 4 || value == 8, "x_tr must be 4 or 8");
        c->x_tr = value;
    } else if (!strcmp(key, "x_p")) {
        NEED(value >= 1 && value <= 3, "x_p must be in 1..3");
        c->x_p = value;
    } else if (!strcmp(key, "z_cfg")) {
        NEED(value >= kZPassesMin && value <= kZPassesMax, "z_cfg must be in %d..%d", kZPassesMin, kZPassesMax);
        c->z_cfg = value;
    } else if (!strcmp(key, "pdl")) {
        c->pdl = value ? 1 : 0;
    } else if (!strcmp(key, "peer_dbg")) {
#ifdef PMW_DEV
        c->peer_dbg = value;
#else
        return fail(PMW_EINVAL, "pmw_set_tuning: 'peer_dbg' exists in development builds only (-DPMW_DEV)");
#endif
    } else if (!strcmp(key, "l2_hints")) {
        c->l2_hints = value;
    } else if (!strcmp(key, "chunks")) {
        NEED(value >= 1 && value <= 4, "chunks must be in 1..4");
        c->chunks = value;
    } else if (!strcmp(key, "fuse")) {
        c->fuse = value ? 1 : 0;
    } else if (!strcmp(key, "keep_tmp")) {
        NEED(value >= 0 && value <= 2, "keep_tmp must be 0 (never), 1 (on demand) or 2 (with every call)");
        c->keep_tmp = value;
    } else if (!strcmp(key, "sweep_lz")) {
        NEED(value == 0 || value >= 8, "sweep_lz must be 0 (automatic) or >= 8");
        c->sweep_lz = value;
    } else if (!strcmp(key, "sweep_zt")) {
        NEED(value >= -1 && value <= 1, "sweep_zt must be -1 (by grid size), 0 (streaming) or 1 (transposing)");
        c->sweep_zt = value;
    } else if (!strcmp(key, "sweep_z3")) {
        c->sweep_z3 = value ? 1 : 0;
    } else if (!strcmp(key, "sweep_xp")) {
        NEED(value == 2 || value == 3, "sweep_xp must be 2 or 3");
        c->sweep_xp = value;
    } else if (!strcmp(key, "push_all_mcells")) {
        NEED(value >= 0, "push_all_mcells must be >= 0 (0: never)");
        c->push_all_cells = (long long)value << 20;
    } else if (!strcmp(key, "dyn_items")) {
        NEED(value >= 0 && value <= 2, "dyn_items must be 0, 1 or 2");
        c->dyn_items = value;
    } else {
        return fail(PMW_EINVAL, "pmw_set_tuning: unknown key '%s'", key);
    }
    return PMW_OK;
}

extern "C" int pmw_get_tuning(pmw_ctx* c, const char* key, int* value)
{
    BIND_KEEP(c);
    NEED(key && value, "pmw_get_tuning: null argument");
    if (!strcmp(key, "x_tr")) *value = c->x_tr;
    else if (!strcmp(key, "x_p")) *value = c->x_p;
    else if (!strcmp(key, "z_cfg")) *value = c->z_cfg;
    else if (!strcmp(key, "pdl")) *value = c->pdl;
    else if (!strcmp(key, "chunks")) *value = c->chunks;
    else if (!strcmp(key, "fuse")) *value = c->fuse;
    else if (!strcmp(key, "keep_tmp")) *value = c->keep_tmp;
    else if (!strcmp(key, "sweep_lz")) *value = c->sweep_lz;
    else if (!strcmp(key, "sweep_xp")) *value = c->sweep_xp;
    else if (!strcmp(key, "sweep_zt")) *value = c->sweep_zt;
    else if (!strcmp(key, "sweep_z3")) *value = c->sweep_z3;
    else if (!strcmp(key, "dyn_items")) *value = c->dyn_items;
    else if (!strcmp(key, "push_all_mcells")) *value = (int)(c->push_all_cells >> 20);
    else return fail(PMW_EINVAL, "pmw_get_tuning: unknown key '%s'", key);
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// data movement
// ---------------------------------------------------------------------------------------------
extern "C" int pmw_set_hydrostatic(pmw_ctx* c, const double* dens_cell, const double* dens_theta_cell,
                                   const double* dens_int, const double* dens_theta_int,
                                   const double* pressure_int)
{
    BIND(c);
    NEED(dens_cell && dens_theta_cell && dens_int && dens_theta_int && pressure_int,
         "pmw_set_hydrostatic: null profile");
    const int ncell = c->p.nz + 4, nint = c->p.nz + 1, npack = nint + 2 * HY_PACK_PAD;
    const size_t pack_off = ((size_t)4 * ncell + (size_t)4 * nint + 3) / 4 * 4;  // 32-byte aligned (cudaMalloc base is)
    const size_t cpack_off = pack_off + (size_t)4 * npack;
    std::vector<double> h(cpack_off + (size_t)4 * ncell);
    double* q = h.data();
    double* o_dc = q;            q += ncell;
    double* o_dtc = q;           q += ncell;
    double* o_idtc = q;          q += ncell;
    double* o_pc = q;            q += ncell;
    double* o_di = q;            q += nint;
    double* o_dti = q;           q += nint;
    double* o_pi = q;            q += nint;
    double* o_idti = q;
    double* o_pack = h.data() + pack_off;
    for (int k = 0; k < ncell; ++k) {
        NEED(dens_cell[k] > 0 && dens_theta_cell[k] > 0, "pmw_set_hydrostatic: non-positive cell profile at %d", k);
        o_dc[k] = dens_cell[k];
        o_dtc[k] = dens_theta_cell[k];
        o_idtc[k] = 1.0 / dens_theta_cell[k];
        o_pc[k] = C0 * std::pow(dens_theta_cell[k], GAMMA);
    }
    // The background-relative pressure (PMW_POW_BACKGROUND) evaluates the z perturbation pressure as
    // hy_pressure_int * ((1+e)^gamma - 1), which is the reference's C0*(rho*theta)^gamma - hy_pressure_int
    // (interpolate.py:160-165) only if the caller's pressure profile IS C0 * hy_dens_theta_int^gamma.  Profiles
    // from init() are (initial.py:84-105); anything else gets the reference's own formula.
    double worst = 0.0;
    for (int k = 0; k < nint; ++k) {
        NEED(dens_int[k] > 0 && dens_theta_int[k] > 0, "pmw_set_hydrostatic: non-positive interface profile at %d", k);
        o_di[k] = dens_int[k];
        o_dti[k] = dens_theta_int[k];
        o_pi[k] = pressure_int[k];
        o_idti[k] = 1.0 / dens_theta_int[k];
        const double want = C0 * std::pow(dens_theta_int[k], GAMMA);
        worst = std::max(worst, std::fabs(pressure_int[k] - want) / want);
    }
    c->hydro_consistent = worst <= 1e-12;
    for (int j = 0; j < npack; ++j) {  // entry j = interface j - HY_PACK_PAD, clamped to the domain
        const int k = std::min(std::max(j - HY_PACK_PAD, 0), nint - 1);
        o_pack[4 * j + 0] = o_di[k];
        o_pack[4 * j + 1] = o_dti[k];
        o_pack[4 * j + 2] = o_idti[k];
        o_pack[4 * j + 3] = o_pi[k];
    }
    for (int k = 0; k < ncell; ++k) {
        double* e = h.data() + cpack_off + 4 * (size_t)k;
        e[0] = o_dc[k]; e[1] = o_dtc[k]; e[2] = o_idtc[k]; e[3] = o_pc[k];
    }
    CU_TRY(cudaMemcpyAsync(c->hydro_blob, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    double* d = c->hydro_blob;
    c->hy.dens_cell = d;                d += ncell;
    c->hy.dens_theta_cell = d;          d += ncell;
    c->hy.inv_dens_theta_cell = d;      d += ncell;
    c->hy.pressure_cell = d;            d += ncell;
    c->hy.dens_int = d;                 d += nint;
    c->hy.dens_theta_int = d;           d += nint;
    c->hy.pressure_int = d;             d += nint;
    c->hy.inv_dens_theta_int = d;
    c->hy.int_pack = c->hydro_blob + pack_off + (size_t)4 * HY_PACK_PAD;
    c->hy.cell_pack = c->hydro_blob + cpack_off;
    c->hy.pad_ = nullptr;
    c->hydro_set = true;
    return PMW_OK;
}

extern "C" int pmw_set_source_w(pmw_ctx* c, const double* host)
{
    BIND(c);
    if (!host) {
        if (c->src_w) cudaFree(c->src_w);
        c->src_w = nullptr;
        return PMW_OK;
    }
    const size_t n = (size_t)c->p.nx * c->p.nz;
    if (!c->src_w) CU_TRY(cudaMalloc(&c->src_w, n * sizeof(double)));
    CU_TRY(cudaMemcpyAsync(c->src_w, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PMW_OK;
}

extern "C" int pmw_set_inflow(pmw_ctx* c, const unsigned char* host_rows, double u_in, double theta_in)
{
    BIND(c);
    if (!host_rows) {
        if (c->jet_rows) cudaFree(c->jet_rows);
        c->jet_rows = nullptr;
        return PMW_OK;
    }
    NEED(c->p.periodic_x && !c->peers, "pmw_set_inflow: the injection inflow needs the whole domain in one context "
                                       "(periodic_x=1, no slab ring)");
    if (!c->jet_rows) CU_TRY(cudaMalloc(&c->jet_rows, (size_t)c->p.nz));
    CU_TRY(cudaMemcpyAsync(c->jet_rows, host_rows, (size_t)c->p.nz, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->jet_u = u_in;
    c->jet_theta = theta_in;
    for (int b = 0; b < 3; ++b) c->xhalo_valid[b] = c->xhalo6_valid[b] = false;
    return PMW_OK;
}

static int check_buf(int buf)
{
    if (buf != PMW_BUF_STATE && buf != PMW_BUF_TMP) return fail(PMW_EINVAL, "unknown buffer id %d", buf);
    return PMW_OK;
}
#define CHECK_BUF(b)              \
    do {                          \
        int rc_ = check_buf(b);   \
        if (rc_ != PMW_OK) return rc_; \
    } while (0)

static int copy_state(pmw_ctx* c, int buf, double* host, bool to_device, bool sync)
{
    BIND_KEEP(c);
    ENSURE_TMP_IF(c, to_device || buf == PMW_BUF_TMP);
    CHECK_BUF(buf);
    NEED(host, "null host pointer");
    const size_t NX = c->p.nx + 4, rows = (size_t)NVAR * (c->p.nz + 4);
    double* dev = c->base[c->l2p[buf]];
    if (to_device) {
        CU_TRY(cudaMemcpy2DAsync(dev, c->L.pitch * sizeof(double), host, NX * sizeof(double), NX * sizeof(double),
                                 rows, cudaMemcpyHostToDevice, c->stream));
        c->xhalo_valid[c->l2p[buf]] = false;
        c->xhalo6_valid[c->l2p[buf]] = false;
    } else {
        CU_TRY(cudaMemcpy2DAsync(host, NX * sizeof(double), dev, c->L.pitch * sizeof(double), NX * sizeof(double),
                                 rows, cudaMemcpyDeviceToHost, c->stream));
    }
    if (sync) CU_TRY(cudaStreamSynchronize(c->stream));
    return PMW_OK;
}

extern "C" int pmw_upload_state(pmw_ctx* c, int buf, const double* host)
{
    return copy_state(c, buf, const_cast<double*>(host), true, true);
}
extern "C" int pmw_download_state(pmw_ctx* c, int buf, double* host)
{
    const int rc = copy_state(c, buf, host, false, true);
    return rc != PMW_OK ? rc : check_watchdog(c);
}
extern "C" int pmw_upload_state_async(pmw_ctx* c, int buf, const double* host)
{
    return copy_state(c, buf, const_cast<double*>(host), true, false);
}
extern "C" int pmw_download_state_async(pmw_ctx* c, int buf, double* host)
{
    return copy_state(c, buf, host, false, false);
}

extern "C" int pmw_buffer_info(pmw_ctx* c, int buf, void** base, size_t* pitch, size_t* vstride)
{
    BIND(c);
    CHECK_BUF(buf);
    if (base) *base = c->base[c->l2p[buf]];
    if (pitch) *pitch = (size_t)c->L.pitch;
    if (vstride) *vstride = (size_t)c->L.vstride;
    return PMW_OK;
}

extern "C" long long pmw_launch_count(pmw_ctx* c) { return c ? c->launches : -1; }

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
static int after_launch(pmw_ctx* c, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(PMW_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    c->launches++;
    return PMW_OK;
}
#define LAUNCHED(c, what)                  \
    do {                                   \
        int rc_ = after_launch(c, what);   \
        if (rc_ != PMW_OK) return rc_;     \
    } while (0)

struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(double)); }
};

extern "C" int pmw_init_state(pmw_ctx* c, const pmw_ic_spec* spec, const double* x_axis, const double* z_axis)
{
    BIND(c);
    NEED(spec && x_axis && z_axis, "pmw_init_state: null argument");
    NEED(spec->nbubbles >= 0 && spec->nbubbles <= PMW_IC_MAX_BUBBLES, "pmw_init_state: nbubbles must be in 0..%d",
         PMW_IC_MAX_BUBBLES);
    static_assert(PMW_IC_MAX_BUBBLES == IC_MAX_BUBBLES, "bubble capacity");
    IcSpec s;
    s.nbubbles = spec->nbubbles;
    for (int n = 0; n < IC_MAX_BUBBLES; ++n) {
        s.amp[n] = spec->amp[n]; s.x0[n] = spec->x0[n]; s.z0[n] = spec->z0[n];
        s.xrad[n] = spec->xrad[n]; s.zrad[n] = spec->zrad[n];
        if (n < spec->nbubbles) NEED(s.xrad[n] > 0 && s.zrad[n] > 0, "pmw_init_state: bubble %d has a non-positive radius", n);
    }
    s.wind = spec->wind;
    s.bvfreq = spec->bvfreq ? 1 : 0;
    s.bv0 = spec->bv0;
    NEED(!s.bvfreq || s.bv0 > 0, "pmw_init_state: bv0 must be positive");
    s.dx = c->p.dx;
    s.dz = c->p.dz;
    const int NX = c->p.nx + 2 * HS, NZ = c->p.nz + 2 * HS;
    DevBuf axes;
    CU_TRY(axes.alloc((size_t)NX + NZ));
    CU_TRY(cudaMemcpyAsync(axes.p, x_axis, (size_t)NX * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(axes.p + NX, z_axis, (size_t)NZ * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const long long n = (long long)NX * NZ;
    const int pS = c->l2p[PMW_BUF_STATE], pT = c->l2p[PMW_BUF_TMP];
    init_state_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->base[pS], c->base[pT], c->L, s, axes.p,
                                                                          axes.p + NX);
    LAUNCHED(c, "init_state_kernel");
    CU_TRY(cudaStreamSynchronize(c->stream));  // the axes buffer is freed on return
    c->xhalo_valid[pS] = c->xhalo_valid[pT] = false;
    c->xhalo6_valid[pS] = c->xhalo6_valid[pT] = false;
    return PMW_OK;
}

extern "C" int pmw_bc_x(pmw_ctx* c, int buf)
{
    BIND(c);
    CHECK_BUF(buf);
    if (c->jet_rows) {  // injection branch of set_bc_x (bcs.py:37,41-64)
        NEED(c->hydro_set, "pmw_bc_x: the inflow rows need the hydrostatic profiles (pmw_set_hydrostatic)");
        bc_x_inflow_kernel<<<(c->p.nz + 127) / 128, 128, 0, c->stream>>>(
            c->base[c->l2p[buf]], c->L, c->hy.dens_cell, c->hy.dens_theta_cell, c->jet_rows, c->jet_u, c->jet_theta);
        LAUNCHED(c, "bc_x_inflow_kernel");
        c->xhalo_valid[c->l2p[buf]] = c->xhalo6_valid[c->l2p[buf]] = false;  // not a periodic image
        return PMW_OK;
    }
    const int n = NVAR * c->p.nz;
    bc_x_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[c->l2p[buf]], c->L);
    LAUNCHED(c, "bc_x_kernel");
    c->xhalo_valid[c->l2p[buf]] = true;
    c->xhalo6_valid[c->l2p[buf]] = false;
    return PMW_OK;
}

extern "C" int pmw_bc_z(pmw_ctx* c, int buf)
{
    BIND(c);
    CHECK_BUF(buf);
    NEED(c->hydro_set, "pmw_bc_z: hydrostatic profiles not set");
    const int n = c->p.nx + 4;
    bc_z_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, c->hy.dens_cell);
    LAUNCHED(c, "bc_z_kernel");
    return PMW_OK;
}

// wide = false: the reference-shaped array [4][nz+4][nx+4] (map column 0 = array column 0);
// wide = true : the same rows with the 6-wide x halo of the fused sweeps, [4][nz+4][nx+12]
//               (map column 0 = array column -4 = interior column -6).
static const int kTmapCap = 256;
// tuning sweeps create many box shapes: start over while no pointer into the cache is live
static void trim_tmaps(pmw_ctx* c)
{
    if (c->tmaps.size() + 8 > (size_t)kTmapCap) c->tmaps.clear();
}

static int get_tmap(pmw_ctx* c, int pbuf, int bw, int bh, const CUtensorMap** out, bool wide = false)
{
    const int key_buf = pbuf + (wide ? 16 : 0);
    for (const TmapKey& k : c->tmaps)
        if (k.buf == key_buf && k.bw == bw && k.bh == bh) {
            *out = &k.map;
            return PMW_OK;
        }
    // returned pointers stay valid until the next trim_tmaps(): the vector never reallocates (reserve) and is
    // only emptied by trim_tmaps, which the launch helpers call BEFORE they fetch the maps of a launch
    if (c->tmaps.capacity() < kTmapCap) c->tmaps.reserve(kTmapCap);
    if (c->tmaps.size() >= (size_t)kTmapCap) return fail(PMW_EINVAL, "tensor map cache full");
    TmapKey k;
    k.buf = key_buf; k.bw = bw; k.bh = bh;
    const cuuint64_t gdim[3] = {(cuuint64_t)(c->p.nx + (wide ? 2 * SWEEP_HALO : 2 * HS)), (cuuint64_t)(c->p.nz + 4),
                                (cuuint64_t)NVAR};
    const cuuint64_t gstride[2] = {(cuuint64_t)c->L.pitch * sizeof(double), (cuuint64_t)c->L.vstride * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)NVAR};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = c->encode(&k.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3,
                           c->base[pbuf] - (wide ? SWEEP_HALO - HS : 0), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(PMW_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (box %dx%dx4)", (int)r, bw, bh);
    c->tmaps.push_back(k);
    *out = &c->tmaps.back().map;
    return PMW_OK;
}

// Launch with the programmatic-stream-serialization attribute when `pdl` is set (the kernel
// then synchronises with its predecessor through griddepcontrol.wait).
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             bool pdl, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename K>
static int set_smem(K kernel, size_t bytes)
{
    CU_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return PMW_OK;
}

template <int TR, int P>
static int launch_x_tma(pmw_ctx* c, bool has_init, const CUtensorMap& tf, const CUtensorMap& ti, const StageArgs& a_in)
{
    using T = XTile<TR, P>;
    const int nty_all = (c->p.nz + TR - 1) / TR;
    const int ty0 = (int)((long long)nty_all * c->cur_chunk / c->cur_nchunks);
    const int ty1 = (int)((long long)nty_all * (c->cur_chunk + 1) / c->cur_nchunks);
    StageArgs a = a_in;
    a.tile_y0 = ty0;
    const dim3 grid((c->p.nx + T::TC - 1) / T::TC, ty1 - ty0 + (a.push_epoch ? 1 : 0));
    if (ty1 == ty0) return PMW_OK;
    const size_t smem = T::smem_bytes(has_init);
    const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
#define GO(HI, PM)                                                                   \
    do {                                                                             \
        static unsigned long long attr_done = 0; /* one bit per device */             \
        if (!(attr_done >> c->p.device & 1ull)) {                                    \
            int rc_ = set_smem(stage_x_tma<TR, P, HI, PM>, T::smem_bytes(true));     \
            if (rc_ != PMW_OK) return rc_;                                           \
            attr_done |= 1ull << c->p.device;                                        \
        }                                                                            \
        launch_ex(stage_x_tma<TR, P, HI, PM>, grid, dim3(T::THREADS), smem, c->launch_stream, c->pdl && !c->timing, \
                  tf, ti, a);                                                 \
    } while (0)
    if (has_init) { if (fast) GO(true, 1); else GO(true, 0); }
    else          { if (fast) GO(false, 1); else GO(false, 0); }
#undef GO
    return PMW_OK;
}

// Gravity-wave configuration: the HAS_SRC instantiations exist for the default tile shapes only.
static int launch_x_tma_src(pmw_ctx* c, bool has_init, const CUtensorMap& tf, const CUtensorMap& ti,
                            const StageArgs& a_in)
{
    using T = XTile<4, 2>;
    const int nty_all = (c->p.nz + 3) / 4;
    const int ty0 = (int)((long long)nty_all * c->cur_chunk / c->cur_nchunks);
    const int ty1 = (int)((long long)nty_all * (c->cur_chunk + 1) / c->cur_nchunks);
    StageArgs a = a_in;
    a.tile_y0 = ty0;
    const dim3 grid((c->p.nx + T::TC - 1) / T::TC, ty1 - ty0 + (a.push_epoch ? 1 : 0));
    if (ty1 == ty0) return PMW_OK;
    const size_t smem = T::smem_bytes(has_init);
    const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
#define GO(HI, PM)                                                                             \
    do {                                                                                       \
        static unsigned long long attr_done = 0;                                               \
        if (!(attr_done >> c->p.device & 1ull)) {                                              \
            int rc_ = set_smem(stage_x_tma<4, 2, HI, PM, true>, T::smem_bytes(true));          \
            if (rc_ != PMW_OK) return rc_;                                                     \
            attr_done |= 1ull << c->p.device;                                                  \
        }                                                                                      \
        launch_ex(stage_x_tma<4, 2, HI, PM, true>, grid, dim3(T::THREADS), smem, c->launch_stream, \
                  c->pdl && !c->timing, tf, ti, a);                                            \
    } while (0)
    if (has_init) { if (fast) GO(true, 1); else GO(true, 0); }
    else          { if (fast) GO(false, 1); else GO(false, 0); }
#undef GO
    return PMW_OK;
}

static int launch_z_tma_src(pmw_ctx* c, bool has_init, const CUtensorMap& tf, const StageArgs& a_in)
{
    using T = ZTile<3>;
    const int ntx_all = (c->p.nx + T::TC - 1) / T::TC;
    const int tx0 = (int)((long long)ntx_all * c->cur_chunk / c->cur_nchunks);
    const int tx1 = (int)((long long)ntx_all * (c->cur_chunk + 1) / c->cur_nchunks);
    StageArgs a = a_in;
    a.tile_x0 = tx0;
    if (tx1 == tx0) return PMW_OK;
    const dim3 grid(tx1 - tx0, (c->p.nz + T::TR - 1) / T::TR);
    const size_t smem = T::smem_bytes(has_init);
    const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
#define GO(HI, PM)                                                                         \
    do {                                                                                   \
        static unsigned long long attr_done = 0;                                           \
        if (!(attr_done >> c->p.device & 1ull)) {                                          \
            int rc_ = set_smem(stage_z_tma<3, HI, PM, true>, T::smem_bytes(true));         \
            if (rc_ != PMW_OK) return rc_;                                                 \
            attr_done |= 1ull << c->p.device;                                              \
        }                                                                                  \
        launch_ex(stage_z_tma<3, HI, PM, true>, grid, dim3(T::THREADS), smem, c->launch_stream, \
                  c->pdl && !c->timing, tf, a);                                            \
    } while (0)
    if (has_init) { if (fast) GO(true, 1); else GO(true, 0); }
    else          { if (fast) GO(false, 1); else GO(false, 0); }
#undef GO
    return PMW_OK;
}

template <int NP>
static int launch_z_tma(pmw_ctx* c, bool has_init, const CUtensorMap& tf, const StageArgs& a_in)
{
    using T = ZTile<NP>;
    const int ntx_all = (c->p.nx + T::TC - 1) / T::TC;
    const int tx0 = (int)((long long)ntx_all * c->cur_chunk / c->cur_nchunks);
    const int tx1 = (int)((long long)ntx_all * (c->cur_chunk + 1) / c->cur_nchunks);
    StageArgs a = a_in;
    a.tile_x0 = tx0;
    if (tx1 == tx0) return PMW_OK;
    const dim3 grid(tx1 - tx0, (c->p.nz + T::TR - 1) / T::TR);
    const size_t smem = T::smem_bytes(has_init);
    const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
#define GO(HI, PM)                                                                 \
    do {                                                                           \
        static unsigned long long attr_done = 0; /* one bit per device */           \
        if (!(attr_done >> c->p.device & 1ull)) {                                  \
            int rc_ = set_smem(stage_z_tma<NP, HI, PM>, T::smem_bytes(true));      \
            if (rc_ != PMW_OK) return rc_;                                         \
            attr_done |= 1ull << c->p.device;                                      \
        }                                                                          \
        launch_ex(stage_z_tma<NP, HI, PM>, grid, dim3(T::THREADS), smem, c->launch_stream,  \
                  c->pdl && !c->timing, tf, a);                                    \
    } while (0)
    if (has_init) { if (fast) GO(true, 1); else GO(true, 0); }
    else          { if (fast) GO(false, 1); else GO(false, 0); }
#undef GO
    return PMW_OK;
}

template <int TR>
static int dispatch_x_p(pmw_ctx* c, bool hi, const CUtensorMap& tf, const CUtensorMap& ti, const StageArgs& a)
{
    switch (c->x_p) {
        case 1: return launch_x_tma<TR, 1>(c, hi, tf, ti, a);
        case 2: return launch_x_tma<TR, 2>(c, hi, tf, ti, a);
        case 3: return launch_x_tma<TR, 3>(c, hi, tf, ti, a);
    }
    return fail(PMW_EINVAL, "bad x_p %d", c->x_p);
}

// One fused stage on PHYSICAL buffers.
static int launch_stage(pmw_ctx* c, int direction, int p_init, int p_forcing, int p_out, double dt_stage,
                        bool write_xhalo, bool fuse_bc_z)
{
    NEED(c->hydro_set, "stage: hydrostatic profiles not set (pmw_set_hydrostatic)");
    NEED(p_out != p_forcing, "internal: stage output aliases the stencil input");
    trim_tmaps(c);
    StageArgs a;
    a.L = c->L;
    a.forcing = c->base[p_forcing];
    a.init = c->base[p_init];
    a.out = c->base[p_out];
    a.out_left = a.out;  // single slab: periodic wrap onto itself
    a.out_right = a.out;
    a.flags = c->flags;
    a.wait_epoch = 0;
    a.edge_last = 0;
    a.push_epoch = 0;
    a.nbr_forcing_left = a.nbr_forcing_right = nullptr;
    a.nbr_flags_left = a.nbr_flags_right = nullptr;
    a.push_counter = c->edge_counters;
    a.tile_x0 = a.tile_y0 = 0;
    if (c->cur_nchunks == 1) c->launch_stream = c->stream;
    {   // l2_hints = decimal "abcd": a = forcing when init==forcing (stage 1), b = forcing otherwise,
        // c = init, d = out; each 0 normal | 1 evict_first | 2 evict_last
        const int h = c->l2_hints;
        a.hint_forcing = (p_init == p_forcing) ? (h / 1000) % 10 : (h / 100) % 10;
        a.hint_init = (h / 10) % 10;
        a.hint_out = h % 10;
    }
    a.hy = c->hy;
    a.src_w = c->src_w;
    const double d = (direction == PMW_DIR_X) ? c->p.dx : c->p.dz;
    a.hv_coeff = -HV_BETA * d / (16 * c->p.dt);
    a.inv_d = 1.0 / d;
    a.dt_stage = dt_stage;
    a.cd = dt_stage * a.inv_d;
    a.cg = dt_stage * GRAV;
    a.write_xhalo = (write_xhalo && c->p.periodic_x) ? 1 : 0;
    if (c->peers && fuse_bc_z && direction == PMW_DIR_X) {
        // Slab ring, fused path: this x stage first pushes its own edge columns of `forcing` into the
        // neighbours' halo columns (first row of CTAs) and publishes epoch e; its edge tiles, the last
        // CTAs of the grid, wait until both neighbours have published e for OUR halo columns.
        a.push_epoch = a.wait_epoch = ++c->epoch;
        a.edge_last = 1;
        a.nbr_forcing_left = c->nbr_base[0][p_forcing];
        a.nbr_forcing_right = c->nbr_base[1][p_forcing];
        a.nbr_flags_left = c->nbr_flags[0];
        a.nbr_flags_right = c->nbr_flags[1];
    }
    const int push_rows = a.push_epoch ? 1 : 0;
    a.fuse_bc_z = fuse_bc_z ? 1 : 0;
    const bool has_init = (p_init != p_forcing);

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {
        if (c->ev_used + 2 > c->ev.size()) {
            for (int n = 0; n < 2; ++n) {
                cudaEvent_t e;
                CU_TRY(cudaEventCreate(&e));
                c->ev.push_back(e);
            }
        }
        e0 = c->ev[c->ev_used];
        e1 = c->ev[c->ev_used + 1];
        c->ev_used += 2;
        CU_TRY(cudaEventRecord(e0, c->stream));
    }

    if (c->p.variant == PMW_VARIANT_DIRECT || (c->p.nx & 1)) {  // the TMA kernels pair cells in x
        const dim3 block(64, 4);
        const dim3 grid((c->p.nx + 63) / 64, (c->p.nz + 3) / 4 + (direction == PMW_DIR_X ? push_rows : 0));
        const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
        if (direction == PMW_DIR_X) {
            if (fast) stage_x_direct<1><<<grid, block, 0, c->stream>>>(a);
            else stage_x_direct<0><<<grid, block, 0, c->stream>>>(a);
        } else {
            if (fast) stage_z_direct<1><<<grid, block, 0, c->stream>>>(a);
            else stage_z_direct<0><<<grid, block, 0, c->stream>>>(a);
        }
    } else {
        const CUtensorMap *tf = nullptr, *ti = nullptr;
        int rc;
        if (c->src_w) {
            // gravity-wave forcing: kernels instantiated with HAS_SRC, default tile shapes only
            if (direction == PMW_DIR_X) {
                if ((rc = get_tmap(c, p_forcing, 132, 4, &tf)) != PMW_OK) return rc;
                if ((rc = get_tmap(c, p_init, 128, 4, &ti)) != PMW_OK) return rc;
                rc = launch_x_tma_src(c, has_init, *tf, *ti, a);
            } else {
                if ((rc = get_tmap(c, p_forcing, 64, 15, &tf)) != PMW_OK) return rc;
                rc = launch_z_tma_src(c, has_init, *tf, a);
            }
            if (rc != PMW_OK) return rc;
        } else if (direction == PMW_DIR_X) {
            const int fw = 64 * c->x_p + 4, iw = 64 * c->x_p;
            if ((rc = get_tmap(c, p_forcing, fw, c->x_tr, &tf)) != PMW_OK) return rc;
            if ((rc = get_tmap(c, p_init, iw, c->x_tr, &ti)) != PMW_OK) return rc;
            rc = (c->x_tr == 4) ? dispatch_x_p<4>(c, has_init, *tf, *ti, a) : dispatch_x_p<8>(c, has_init, *tf, *ti, a);
            if (rc != PMW_OK) return rc;
        } else {
            const int np = c->z_cfg;
            if ((rc = get_tmap(c, p_forcing, 64, 4 * np + 3, &tf)) != PMW_OK) return rc;
            switch (np) {
                case 1: rc = launch_z_tma<1>(c, has_init, *tf, a); break;
                case 2: rc = launch_z_tma<2>(c, has_init, *tf, a); break;
                case 3: rc = launch_z_tma<3>(c, has_init, *tf, a); break;
                case 4: rc = launch_z_tma<4>(c, has_init, *tf, a); break;
                case 5: rc = launch_z_tma<5>(c, has_init, *tf, a); break;
                case 6: rc = launch_z_tma<6>(c, has_init, *tf, a); break;
                case 7: rc = launch_z_tma<7>(c, has_init, *tf, a); break;
                case 8: rc = launch_z_tma<8>(c, has_init, *tf, a); break;
                default: rc = fail(PMW_EINVAL, "bad z_cfg");
            }
            if (rc != PMW_OK) return rc;
        }
    }
    LAUNCHED(c, direction == PMW_DIR_X ? "x stage kernel" : "z stage kernel");
    if (c->timing) CU_TRY(cudaEventRecord(e1, c->stream));
    c->xhalo_valid[p_out] = a.write_xhalo != 0;
    c->xhalo6_valid[p_out] = false;
    return PMW_OK;
}

extern "C" int pmw_stage(pmw_ctx* c, int direction, int init_buf, int forcing_buf, int out_buf, double dt_stage)
{
    BIND(c);
    CHECK_BUF(init_buf);
    CHECK_BUF(forcing_buf);
    CHECK_BUF(out_buf);
    NEED(direction == PMW_DIR_X || direction == PMW_DIR_Z, "pmw_stage: bad direction %d", direction);
    const int p_init = c->l2p[init_buf], p_forcing = c->l2p[forcing_buf];
    if (out_buf == forcing_buf) {
        // step.py:122-131: the update overwrites the array the stencil reads.  NumPy materialises
        // the tendency first; here the result goes to the spare buffer, which then becomes `out`
        // (with the halo ring the in-place array would still carry).
        const int p_out = c->spare;
        int rc = launch_stage(c, direction, p_init, p_forcing, p_out, dt_stage, false, false);
        if (rc != PMW_OK) return rc;
        const int n = NVAR * (4 * (c->p.nx + 4) + 4 * c->p.nz);
        copy_halo_ring_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[p_out], c->base[p_forcing], c->L);
        LAUNCHED(c, "copy_halo_ring_kernel");
        c->xhalo_valid[p_out] = false;
        c->xhalo6_valid[p_out] = false;  // the ring is the image of the OLD interior, as in the reference
        c->l2p[out_buf] = p_out;
        c->spare = p_forcing;
        return PMW_OK;
    }
    // the halo of `out` is untouched by the update (step.py:80-82) while its interior changes,
    // so it stops being the periodic image (launch_stage records that)
    return launch_stage(c, direction, p_init, p_forcing, c->l2p[out_buf], dt_stage, false, false);
}

extern "C" int pmw_discrete_step(pmw_ctx* c, int direction, int init_buf, int forcing_buf, int out_buf,
                                 double dt_stage)
{
    int rc = (direction == PMW_DIR_X) ? pmw_bc_x(c, forcing_buf) : pmw_bc_z(c, forcing_buf);
    if (rc != PMW_OK) return rc;
    return pmw_stage(c, direction, init_buf, forcing_buf, out_buf, dt_stage);
}

// One RK stage of the fused time step, with the three-buffer rotation:
//   stage 1: TMP   <- STATE + dt/3 * T(STATE)
//   stage 2: spare <- STATE + dt/2 * T(TMP);  TMP := spare
//   stage 3: STATE <- STATE + dt   * T(TMP)   (in place: init is only read at the written cell)
extern "C" int pmw_evolve_stage(pmw_ctx* c, int direction, int rk_stage, double dt)
{
    BIND(c);
    NEED(direction == PMW_DIR_X || direction == PMW_DIR_Z, "pmw_evolve_stage: bad direction %d", direction);
    NEED(rk_stage >= 1 && rk_stage <= 3, "pmw_evolve_stage: rk_stage must be 1..3");
    if (dt <= 0) dt = c->p.dt;
    const int S = c->l2p[PMW_BUF_STATE], T = c->l2p[PMW_BUF_TMP];
    const int p_forcing = (rk_stage == 1) ? S : T;
    if (direction == PMW_DIR_X && !c->xhalo_valid[p_forcing]) {
        if (c->peers) {
            // nothing to do: in a connected slab every x stage pushes/receives its own halo columns
        } else {
            NEED(c->p.periodic_x, "pmw_evolve_stage: x halos of the forcing buffer are stale; a slab context "
                                  "needs pmw_unpack_halo_x (neighbour columns) before every x stage");
            const int n = NVAR * c->p.nz;
            bc_x_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[p_forcing], c->L);
            LAUNCHED(c, "bc_x_kernel");
        }
        c->xhalo_valid[p_forcing] = true;
        c->xhalo6_valid[p_forcing] = false;
    }
    // Halo images are stored by the stage whose output the next x stage reads: x stages 1 and 2,
    // and every stage 3 (the next sweep may be an x sweep).  z stages 1 and 2 feed z stages only.
    const bool images = (direction == PMW_DIR_X) || rk_stage == 3;
    int rc;
    if (rk_stage == 1) {
        rc = launch_stage(c, direction, S, S, T, dt / 3, images, true);
    } else if (rk_stage == 2) {
        rc = launch_stage(c, direction, S, T, c->spare, dt / 2, images, true);
        if (rc == PMW_OK) {
            c->l2p[PMW_BUF_TMP] = c->spare;
            c->spare = T;
        }
    } else {
        rc = launch_stage(c, direction, S, T, S, dt / 1, images, true);
    }
    return rc;
}

// One directional sweep (three RK stages) as `chunks` independent kernel chains: bands of tile columns
// for a z sweep, bands of tile rows for an x sweep.  Kernels are submitted stage-major
// (A1 B1 A2 B2 A3 B3), each band on its own stream: B1 does not depend on A1, so it fills the SMs
// while A1 drains, A2 (which waits for A1 through PDL) follows B1, and so on -- only the fork/join
// at the sweep boundaries serialises.
static int evolve_sweep_chunked(pmw_ctx* c, int direction, double dt)
{
    if (dt <= 0) dt = c->p.dt;
    const int K = c->chunks;
    if (!c->ev_fork) {
        CU_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        for (int k = 0; k < 4; ++k) {
            CU_TRY(cudaStreamCreateWithFlags(&c->cstream[k], cudaStreamNonBlocking));
            CU_TRY(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
        }
    }
    const int S = c->l2p[PMW_BUF_STATE], T0 = c->l2p[PMW_BUF_TMP], SP = c->spare;
    if (direction == PMW_DIR_X && !c->xhalo_valid[S]) {
        const int n = NVAR * c->p.nz;
        bc_x_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[S], c->L);
        LAUNCHED(c, "bc_x_kernel");
        c->xhalo_valid[S] = true;
        c->xhalo6_valid[S] = false;
    }
    CU_TRY(cudaEventRecord(c->ev_fork, c->stream));
    for (int k = 0; k < K; ++k) CU_TRY(cudaStreamWaitEvent(c->cstream[k], c->ev_fork, 0));
    int rc = PMW_OK;
    c->cur_nchunks = K;
    for (int rk = 1; rk <= 3 && rc == PMW_OK; ++rk) {
        for (int k = 0; k < K && rc == PMW_OK; ++k) {
            c->cur_chunk = k;
            c->launch_stream = c->cstream[k];
            if (rk == 1) rc = launch_stage(c, direction, S, S, T0, dt / 3, true, true);
            else if (rk == 2) rc = launch_stage(c, direction, S, T0, SP, dt / 2, true, true);
            else rc = launch_stage(c, direction, S, SP, S, dt / 1, true, true);
        }
    }
    c->cur_chunk = 0;
    c->cur_nchunks = 1;
    c->launch_stream = c->stream;
    if (rc != PMW_OK) return rc;
    c->l2p[PMW_BUF_TMP] = SP;
    c->spare = T0;
    for (int k = 0; k < K; ++k) {
        CU_TRY(cudaEventRecord(c->ev_join[k], c->cstream[k]));
        CU_TRY(cudaStreamWaitEvent(c->stream, c->ev_join[k], 0));
    }
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// fused sweeps (pmw_sweep.cuh): one kernel per directional sweep
// ---------------------------------------------------------------------------------------------
static bool fuse_ok(const pmw_ctx* c)
{
    // (the gravity-wave forcing is folded into the sweeps of a single periodic slab; a slab of a ring
    // would need its neighbours' forcing for the recomputed halo cells and runs stage by stage)
    return c->fuse && c->p.variant == PMW_VARIANT_TMA && !(c->p.nx & 1) && c->p.nx >= 16 && c->p.nz >= 8 &&
           (!c->src_w || (c->p.periodic_x && !c->peers));
}

// Rows per z-sweep segment: a warp (32 columns x lz rows) is the unit of work and every SM holds
// kZWarps of them; pick the segment count that fills whole waves with the least recomputation.
static const int kZWarpsPerSM = PMW_ZSWEEP_MINB;
// `units_per_sm` resident work units (warps of sweep_z, CTAs of sweep_z3) per SM; a unit costs
// `stages * lz + fill` interface evaluations.
static int pick_sweep_lz(const pmw_ctx* c, int units_per_sm = kZWarpsPerSM, double stages = 3.0, double fill = 26.0,
                         int nrows = 0)
{
    if (nrows <= 0) nrows = c->p.nz;  // rows of this launch (a band of pmw_evolve_host, else the whole grid)
    if (c->sweep_lz) return std::min(c->sweep_lz, nrows);
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->p.device);
    const long long strips = (c->p.nx + ZS_COLS - 1) / ZS_COLS;
    const long long slots = (long long)nsm * units_per_sm;
    double best = 1e300;
    int best_lz = nrows;
    for (int nseg = 1; nseg <= std::max(nrows / 8, 1); ++nseg) {
        const int lz = (nrows + nseg - 1) / nseg;
        if ((nrows + lz - 1) / lz != nseg) continue;
        const long long warps = strips * nseg;
        const long long waves = (warps + slots - 1) / slots;
        // time ~ waves x (work of one unit, with the recomputed rows) x (how full the SMs are)
        const double per_warp = stages * lz + fill;  // interface evaluations + pipeline fill
        const double occupancy = (double)warps / (waves * slots);
        const double t = waves * per_warp * (0.35 + 0.65 * std::max(occupancy, 0.5));
        if (t < best) { best = t; best_lz = lz; }
    }
    return best_lz;
}

// Which z sweep a grid gets.  The streaming kernel (sweep_z) cannot be shorter than one warp's chain of ~(8 + 26)
// iterations, ~20 us, however small the grid; the transposing kernel (sweep_zt) has no such floor but costs 1.6 x as
// much per cell once the GPU is full.  They cross at ~0.5 M cells (profiles/r2ak, r2al: 100x50 7.0 vs 19.8 us,
// 512x256 11.9 vs 23.1, 1024x512 29.2 vs 28.7, 2048x1024 90.8 vs 57.8).
static const long long kZtMaxCells = 400000;
static bool use_zt(const pmw_ctx* c)
{
    return c->sweep_zt == 1 || (c->sweep_zt < 0 && (long long)c->p.nx * c->p.nz <= kZtMaxCells);
}

// row0, row1: the interior rows the launch produces (SweepArgs::row0/row1; row1 < 0: the whole grid)
static int launch_sweep(pmw_ctx* c, int direction, int pS, int pO, int pT, bool write_tmp, double dt, int row0 = 0,
                        int row1 = -1)
{
    if (row1 < 0) row1 = c->p.nz;
    const bool whole = row0 == 0 && row1 == c->p.nz;
    NEED(row0 >= 0 && row0 < row1 && row1 <= c->p.nz, "sweep: bad row range");
    trim_tmaps(c);
    NEED(c->hydro_set, "sweep: hydrostatic profiles not set (pmw_set_hydrostatic)");
    SweepArgs a;
    a.L = c->L;
    a.state = c->base[pS];
    a.out = c->base[pO];
    a.tmp = c->base[pT];
    a.hy = c->hy;
    const double d = (direction == PMW_DIR_X) ? c->p.dx : c->p.dz;
    a.hv_coeff = -HV_BETA * d / (16 * c->p.dt);
    a.inv_d = 1.0 / d;
    a.dt1 = dt / 3;
    a.dt2 = dt / 2;
    a.dt3 = dt / 1;
    // the same products launch_stage forms for the stage-by-stage path (bit-identical results)
    a.cd1 = a.dt1 * a.inv_d; a.cd2 = a.dt2 * a.inv_d; a.cd3 = a.dt3 * a.inv_d;
    a.cg1 = a.dt1 * GRAV;    a.cg2 = a.dt2 * GRAV;    a.cg3 = a.dt3 * GRAV;
    a.periodic = c->p.periodic_x ? 1 : 0;
    a.lz = 0;
    a.tile_x0 = a.tile_y0 = 0;
    a.row0 = row0;
    a.row1 = row1;
    a.flags = c->flags;
    a.wait_epoch = a.push_epoch = 0;
    a.edge_last = 0;
    a.nbr_state_left = a.nbr_state_right = nullptr;
    a.nbr_flags_left = a.nbr_flags_right = nullptr;
    a.push_counter = c->edge_counters;
    a.src_w = c->src_w;
    const bool has_src = c->src_w != nullptr;
    const bool fast = c->p.pow_mode == PMW_POW_BACKGROUND && c->hydro_consistent;
    const CUtensorMap* tm = nullptr;
    int rc;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->timing) {  // pmw_stage_timing: an event pair around every sweep kernel
        while (c->ev_used + 2 > c->ev.size()) {
            cudaEvent_t e;
            CU_TRY(cudaEventCreate(&e));
            c->ev.push_back(e);
        }
        e0 = c->ev[c->ev_used];
        e1 = c->ev[c->ev_used + 1];
        c->ev_used += 2;
    }
    if (direction == PMW_DIR_X) {
        if (c->peers) {
            NEED(whole, "sweep: row bands are single-context only");
            a.push_epoch = a.wait_epoch = ++c->epoch;
            a.edge_last = 1;
            a.nbr_state_left = c->nbr_base[0][pS];
            a.nbr_state_right = c->nbr_base[1][pS];
            a.nbr_flags_left = c->nbr_flags[0];
            a.nbr_flags_right = c->nbr_flags[1];
        } else if (!c->xhalo6_valid[pS]) {
            NEED(c->p.periodic_x, "pmw_evolve: the x halo of the state is stale on a slab context without peers");
            Layout Lb = c->L;  // the rows of this launch only (idx() never uses nz: the offset pointer does it)
            Lb.nz = row1 - row0;
            const int n = NVAR * Lb.nz * 6;
            bc_x6_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[pS] + (long long)row0 * c->L.pitch, Lb);
            LAUNCHED(c, "bc_x6_kernel");
            if (whole) c->xhalo_valid[pS] = c->xhalo6_valid[pS] = true;
        }
        const int P = c->sweep_xp;
        const int LC = 64 * P - 10;
        if ((rc = get_tmap(c, pS, 64 * P + 4, 1, &tm, true)) != PMW_OK) return rc;
        const int ntx = (c->p.nx + LC - 1) / LC;
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->p.device);
        // persistent: every warp walks its own list of (row, tile) items; as many CTAs of 4 warps as
        // are resident at once
        const long long nitems = (long long)(row1 - row0) * ntx;
        // "dyn_items": 0 never, 1 slab ring, 2 always (not with the gravity-wave forcing: single slab only)
        const bool dyn_items = !has_src && (c->peers ? (c->dyn_items != 0) : (c->dyn_items == 2));
        if (e0) CU_TRY(cudaEventRecord(e0, c->stream));
#define GO(PP, PM, WT)                                                   \
    do {                                                                 \
        if (has_src) GO_S(PP, PM, WT, true, false);                      \
        else if (dyn_items) GO_S(PP, PM, WT, false, true);               \
        else GO_S(PP, PM, WT, false, false);                             \
    } while (0)
#define GO_S(PP, PM, WT, SRC, DYN)                                                                           \
    do {                                                                                                \
        using T = XSweepTile<PP>;                                                                       \
        static unsigned long long attr_done = 0;                                                        \
        static int per_sm[64];                                                                          \
        if (!(attr_done >> c->p.device & 1ull)) {                                                       \
            if ((rc = set_smem(sweep_x<PP, PM, WT, SRC, DYN>, T::smem_bytes())) != PMW_OK) return rc;        \
            int nb = 0;                                                                                 \
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sweep_x<PP, PM, WT, SRC, DYN>, 32 * T::WARPS, \
                                                                 T::smem_bytes()));                     \
            per_sm[c->p.device & 63] = std::max(nb, 1);                                                 \
            attr_done |= 1ull << c->p.device;                                                           \
        }                                                                                               \
        const int ncta = (int)std::min<long long>((long long)nsm * per_sm[c->p.device & 63],            \
                                                  (nitems + T::WARPS - 1) / T::WARPS);                  \
        /* CTAs that push the edge columns first: enough of them that the push of a tall slab ends early */  \
        /* big slabs (sweep of several 100 us): EVERY CTA pushes its share before its first item, while the memory */ \
        /* system is still idle -- a few us per CTA; pushes issued next to running tiles were served late (r2z)   */ \
        const bool push_all = (long long)c->p.nx * c->p.nz >= (c->push_all_cells > 0 ? c->push_all_cells : (1ll << 62));   \
        const int npush = !a.push_epoch ? 0 : push_all ? ncta : std::max(1, std::min(ncta / 2, std::max(8, c->p.nz / 128))); \
        const dim3 grid(ncta);                                                                          \
        a.item_counter = dyn_items ? c->xitem_counter : nullptr;                                        \
        a.item_base = c->xitem_base;                                                                    \
        if (dyn_items) c->xitem_base += (unsigned long long)nitems + 1ull * ncta * T::WARPS;            \
        launch_ex(sweep_x<PP, PM, WT, SRC, DYN>, grid, dim3(32 * T::WARPS), T::smem_bytes(), c->stream, c->pdl && !c->timing, *tm, a, \
                  ntx, npush);                                                                          \
    } while (0)
#define GO_P(PM, WT) do { if (P == 2) GO(2, PM, WT); else GO(3, PM, WT); } while (0)
        if (fast) { if (write_tmp) GO_P(1, true); else GO_P(1, false); }
        else      { if (write_tmp) GO_P(0, true); else GO_P(0, false); }
#undef GO_P
#undef GO
#undef GO_S
        LAUNCHED(c, "sweep_x");
    } else if (use_zt(c) && !has_src) {
        NEED(whole, "sweep: row bands need the streaming z sweep");
        // transposing z sweep: items = (group of 4 columns, z tile); as many CTAs as are resident at once
        const int P = 2, LC = 64 * P - 10;
        if ((rc = get_tmap(c, pS, 4, 64 * P + 4, &tm, true)) != PMW_OK) return rc;
        const int ngroups = (c->p.nx + 3) / 4, ntz = (c->p.nz + LC - 1) / LC;
        int nsm = 148;
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->p.device);
        if (e0) CU_TRY(cudaEventRecord(e0, c->stream));
#define GO(PM, WT)                                                                                      \
    do {                                                                                                \
        using T = ZTSweepTile<2>;                                                                       \
        static unsigned long long attr_done = 0;                                                        \
        static int per_sm[64];                                                                          \
        if (!(attr_done >> c->p.device & 1ull)) {                                                       \
            if ((rc = set_smem(sweep_zt<2, PM, WT>, T::smem_bytes())) != PMW_OK) return rc;             \
            int nb = 0;                                                                                 \
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sweep_zt<2, PM, WT>, 32 * T::WARPS, \
                                                                 T::smem_bytes()));                     \
            per_sm[c->p.device & 63] = std::max(nb, 1);                                                 \
            attr_done |= 1ull << c->p.device;                                                           \
        }                                                                                               \
        const int ncta = (int)std::min<long long>((long long)nsm * per_sm[c->p.device & 63],            \
                                                  (long long)ngroups * ntz);                            \
        launch_ex(sweep_zt<2, PM, WT>, dim3(ncta), dim3(32 * T::WARPS), T::smem_bytes(), c->stream,     \
                  c->pdl && !c->timing, *tm, a, ngroups, ntz);                                                   \
    } while (0)
        if (fast) { if (write_tmp) GO(1, true); else GO(1, false); }
        else      { if (write_tmp) GO(0, true); else GO(0, false); }
#undef GO
        LAUNCHED(c, "sweep_zt");
    } else if (c->sweep_z3) {
        NEED(whole, "sweep: row bands need the streaming z sweep");
        // stage-pipelined z sweep: CTA = 3 warps = the three stages of one strip segment
        if ((rc = get_tmap(c, pS, Z3_COLS, 4, &tm, true)) != PMW_OK) return rc;
#define GO_S(PM, WT, SRC)                                                                                  \
    do {                                                                                                   \
        static unsigned long long attr_done = 0;                                                           \
        static int per_sm[64];                                                                             \
        if (!(attr_done >> c->p.device & 1ull)) {                                                          \
            if ((rc = set_smem(sweep_z3<PM, WT, SRC>, z3_smem_bytes())) != PMW_OK) return rc;              \
            int nb = 0;                                                                                    \
            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sweep_z3<PM, WT, SRC>, 32 * Z3_WARPS, \
                                                                 z3_smem_bytes()));                        \
            per_sm[c->p.device & 63] = std::max(nb, 1);                                                    \
            attr_done |= 1ull << c->p.device;                                                              \
        }                                                                                                  \
        a.lz = pick_sweep_lz(c, per_sm[c->p.device & 63], 1.0, 14.0);                                      \
        const dim3 grid((c->p.nx + Z3_COLS - 1) / Z3_COLS, (c->p.nz + a.lz - 1) / a.lz);                   \
        if (e0) CU_TRY(cudaEventRecord(e0, c->stream));                                                    \
        launch_ex(sweep_z3<PM, WT, SRC>, grid, dim3(32 * Z3_WARPS), z3_smem_bytes(), c->stream,            \
                  c->pdl && !c->timing, *tm, a);                                                           \
    } while (0)
#define GO(PM, WT) do { if (has_src) GO_S(PM, WT, true); else GO_S(PM, WT, false); } while (0)
        if (fast) { if (write_tmp) GO(1, true); else GO(1, false); }
        else      { if (write_tmp) GO(0, true); else GO(0, false); }
#undef GO
#undef GO_S
        LAUNCHED(c, "sweep_z3");
    } else {
        a.lz = pick_sweep_lz(c, kZWarpsPerSM, 3.0, 26.0, row1 - row0);
        if ((rc = get_tmap(c, pS, ZS_COLS, 1, &tm, true)) != PMW_OK) return rc;
        const dim3 grid((c->p.nx + ZS_COLS - 1) / ZS_COLS, (row1 - row0 + a.lz - 1) / a.lz);
        if (e0) CU_TRY(cudaEventRecord(e0, c->stream));
#define GO_S(PM, WT, SRC)                                                                                        \
    do {                                                                                                         \
        static unsigned long long attr_done = 0; /* one bit per device: all of the SM's shared memory, so that */ \
        if (!(attr_done >> c->p.device & 1ull)) { /* kZWarpsPerSM one-warp CTAs fit */                            \
            CU_TRY(cudaFuncSetAttribute(sweep_z<PM, WT, SRC>, cudaFuncAttributePreferredSharedMemoryCarveout,    \
                                        cudaSharedmemCarveoutMaxShared));                                        \
            attr_done |= 1ull << c->p.device;                                                                    \
        }                                                                                                        \
        launch_ex(sweep_z<PM, WT, SRC>, grid, dim3(32), zsweep_smem_bytes(), c->stream, c->pdl && !c->timing,    \
                  *tm, a);                                                                                       \
    } while (0)
#define GO(PM, WT) do { if (has_src) GO_S(PM, WT, true); else GO_S(PM, WT, false); } while (0)
        if (fast) { if (write_tmp) GO(1, true); else GO(1, false); }
        else      { if (write_tmp) GO(0, true); else GO(0, false); }
#undef GO
#undef GO_S
        LAUNCHED(c, "sweep_z");
    }
    if (e1) CU_TRY(cudaEventRecord(e1, c->stream));
    // S' carries the periodic images of its own edge columns (single slab); in a ring they are pushed
    // by the next x sweep
    if (whole) c->xhalo_valid[pO] = c->xhalo6_valid[pO] = (c->p.periodic_x != 0);
    return PMW_OK;
}

// The reference's state_tmp after evolve is the stage-2 array of the step's last sweep (step.py:122-131).  A
// fused sweep only stores it on request, and that store (32 B per cell) is 11 % of a time step for a caller that
// steps once per call (the reference's driver loop, __main__.py:237) and never looks at it.  So the last sweep of a
// pmw_evolve call leaves it out, remembers itself, and is run again with the store switched on -- from its
// untouched input in the spare buffer, rewriting its own output with the same bits -- the first time an entry
// point is about to read state_tmp or to change a buffer.
static int ensure_tmp(pmw_ctx* c)
{
    if (!c->tmp_pending) return PMW_OK;
    c->tmp_pending = false;
    if (c->spare != c->tp_src || c->l2p[PMW_BUF_STATE] != c->tp_out)
        return fail(PMW_EINVAL, "internal error: the buffers of a pending state_tmp were rotated");
    const int T = c->l2p[PMW_BUF_TMP];
    const int rc = launch_sweep(c, c->tp_dir, c->tp_src, c->tp_out, T, true, c->tp_dt);
    c->xhalo_valid[T] = c->xhalo6_valid[T] = false;
    return rc;
}

static int evolve_sweep_fused(pmw_ctx* c, int direction, double dt, bool write_tmp)
{
    if (dt <= 0) dt = c->p.dt;
    const int S = c->l2p[PMW_BUF_STATE], T = c->l2p[PMW_BUF_TMP], O = c->spare;
    int rc = launch_sweep(c, direction, S, O, T, write_tmp, dt);
    if (rc != PMW_OK) return rc;
    c->l2p[PMW_BUF_STATE] = O;
    c->spare = S;
    if (write_tmp) c->xhalo_valid[T] = c->xhalo6_valid[T] = false;
    return PMW_OK;
}

extern "C" int pmw_evolve(pmw_ctx* c, int nsteps, double dt)
{
    BIND_KEEP(c);
    NEED(nsteps >= 0, "pmw_evolve: negative nsteps");
    NEED(c->p.periodic_x || c->peers,
         "pmw_evolve: a slab context (periodic_x=0) needs pmw_connect_peers, or must be stepped stage by "
         "stage with pmw_evolve_stage and a halo exchange before every x stage");
    if (c->jet_rows) {
        int rc0 = ensure_tmp(c);  // (this path goes through the operator entry points, which read state_tmp)
        if (rc0 != PMW_OK) return rc0;
        // Injection: x is not periodic, so neither the fused sweeps (6-wide periodic halo) nor the
        // stages that write the wrap for their successor apply.  Run the reference's own sequence
        // (step.py:105-143): explicit halo fill, then the fused stage kernel, per stage; arrays that
        // are updated in place keep their halo ring, so the right halo columns keep their initial
        // values exactly as in the reference.
        NEED(c->p.periodic_x && !c->peers, "pmw_evolve: the injection inflow is single-context only");
        if (dt <= 0) dt = c->p.dt;
        for (int n = 0; n < nsteps; ++n) {
            const int dirs[2] = {c->reverse ? PMW_DIR_X : PMW_DIR_Z, c->reverse ? PMW_DIR_Z : PMW_DIR_X};
            for (int d = 0; d < 2; ++d) {
                int rc = pmw_discrete_step(c, dirs[d], PMW_BUF_STATE, PMW_BUF_STATE, PMW_BUF_TMP, dt / 3);
                if (rc == PMW_OK) rc = pmw_discrete_step(c, dirs[d], PMW_BUF_STATE, PMW_BUF_TMP, PMW_BUF_TMP, dt / 2);
                if (rc == PMW_OK) rc = pmw_discrete_step(c, dirs[d], PMW_BUF_STATE, PMW_BUF_TMP, PMW_BUF_STATE, dt / 1);
                if (rc != PMW_OK) return rc;
            }
            c->reverse = !c->reverse;
        }
        return PMW_OK;
    }
    if (dt <= 0) dt = c->p.dt;
    if (nsteps > 0) c->tmp_pending = false;  // this call produces a new state_tmp (no path below reads the old one)
    // (x sweeps of a connected slab carry the halo push / epoch wait per stage: those stay whole)
    const bool chunk_ok = c->chunks > 1 && c->p.variant == PMW_VARIANT_TMA && !(c->p.nx & 1) && !c->timing;
    const bool fused = fuse_ok(c);
    for (int n = 0; n < nsteps; ++n) {
        const int dirs[2] = {c->reverse ? PMW_DIR_X : PMW_DIR_Z, c->reverse ? PMW_DIR_Z : PMW_DIR_X};
        for (int d = 0; d < 2; ++d) {
            int rc = PMW_OK;
#ifdef PMW_DEV  // development builds only (tools/build_variant.py): time one direction of the fused step
            if (fused && (c->peer_dbg & (dirs[d] == PMW_DIR_X ? 16 : 8))) continue;
#endif
            if (fused) {
                // state_tmp (the reference's stage-2 array) is only materialised by the last sweep of the call
                const bool last = n == nsteps - 1 && d == 1;
                const bool lazy = last && c->keep_tmp == 1 && !c->peers && !c->timing;
                rc = evolve_sweep_fused(c, dirs[d], dt, last && c->keep_tmp && !lazy);
                if (rc == PMW_OK && lazy) {
                    c->tmp_pending = true;
                    c->tp_dir = dirs[d];
                    c->tp_dt = dt;
                    c->tp_src = c->spare;
                    c->tp_out = c->l2p[PMW_BUF_STATE];
                }
            } else if (chunk_ok && (!c->peers || dirs[d] == PMW_DIR_Z)) {
                rc = evolve_sweep_chunked(c, dirs[d], dt);
            } else {
                for (int s = 1; s <= 3 && rc == PMW_OK; ++s) rc = pmw_evolve_stage(c, dirs[d], s, dt);
            }
            if (rc != PMW_OK) return rc;
        }
        c->reverse = !c->reverse;
    }
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// One time step on a HOST array, streamed in row bands
// ---------------------------------------------------------------------------------------------
// What the drop-in `evolve` on host arrays costs is PCIe: 2 x 67.5 MB at 2048 x 1024 against 0.1 ms of
// kernels, and upload -> step -> download one after the other uses one direction of the link at a time.
// But a step only couples rows through the z sweep's 6-row halo (three stages x two cells), and the x sweep
// not at all: the final rows [k0, k1) depend on the initial rows [k0-6, k1+6) alone.  So the step runs in
// bands of rows -- band b is uploaded on one copy stream (the copies run six rows ahead of the bands), swept (both
// directions, row-range launches of the same fused kernels: bit-identical to the whole-grid launches) on the
// context's stream as soon as its copy has landed, and downloaded on a second copy stream while later bands are still on their way up: H2D and
// D2H overlap and the call takes little more than ONE transfer of the state.  Buffers: the upload goes to the
// state buffer A, the first sweep writes the spare buffer C, the second the tmp buffer B (A cannot take it:
// later bands still read A's rows next to the band), B becomes the state.  The reference's state_tmp is
// not produced (it is scratch there, solve/step.py:112-141).  The host array is updated in place: a band's
// download only overwrites rows whose upload has completed.
static int evolve_host_plain(pmw_ctx* c, double* host, double dt)
{
    int rc = copy_state(c, PMW_BUF_STATE, host, true, false);
    if (rc == PMW_OK) rc = pmw_evolve(c, 1, dt);
    if (rc == PMW_OK) rc = copy_state(c, PMW_BUF_STATE, host, false, true);
    return rc != PMW_OK ? rc : check_watchdog(c);
}

extern "C" int pmw_evolve_host(pmw_ctx* c, double* host_state, double dt, int nbands)
{
    BIND_KEEP(c);
    NEED(host_state, "pmw_evolve_host: null host pointer");
    NEED(nbands >= 0, "pmw_evolve_host: negative band count");
    if (dt <= 0) dt = c->p.dt;
    c->tmp_pending = false;  // every buffer is overwritten; the tmp buffer holds the previous state afterwards
    const int nz = c->p.nz;
    const bool can_band = fuse_ok(c) && c->p.periodic_x && !c->peers && !c->jet_rows && !c->src_w && !use_zt(c) &&
                          !c->sweep_z3 && !c->timing;
    if (nbands == 0)  // bands of ~128 rows, 32 at most: a band-sized DMA job costs ~8 us over what its bytes cost,
        nbands = std::min(32, nz / 128);  // the pipeline's fill and drain one band each (measured: profiles/r2bx)
    nbands = std::min(nbands, nz / 16);
    if (!can_band || nbands < 2) return evolve_host_plain(c, host_state, dt);

    for (int k = 0; k < 2; ++k)
        if (!c->hs_stream[k]) CU_TRY(cudaStreamCreateWithFlags(&c->hs_stream[k], cudaStreamNonBlocking));
    if (!c->hs_start) CU_TRY(cudaEventCreateWithFlags(&c->hs_start, cudaEventDisableTiming));
    while ((int)c->hs_ev.size() < 2 * nbands) {
        cudaEvent_t e;
        CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->hs_ev.push_back(e);
    }
    cudaStream_t up = c->hs_stream[0], down = c->hs_stream[1];
    const int A = c->l2p[PMW_BUF_STATE], B = c->l2p[PMW_BUF_TMP], C = c->spare;
    const bool xfirst = c->reverse != 0;
    const size_t NX = c->p.nx + 4;
    // one 3-D copy per band and direction: [4 variables][rows of the band][nx+4]
    auto copy_band = [&](int buf, int ar0, int ar1, bool to_device, cudaStream_t st) -> cudaError_t {
        cudaMemcpy3DParms q = {};
        const cudaPitchedPtr h = make_cudaPitchedPtr(host_state, NX * sizeof(double), NX * sizeof(double), nz + 4);
        const cudaPitchedPtr d = make_cudaPitchedPtr(c->base[buf], c->L.pitch * sizeof(double), NX * sizeof(double), nz + 4);
        q.srcPtr = to_device ? h : d;
        q.dstPtr = to_device ? d : h;
        q.srcPos = q.dstPos = make_cudaPos(0, ar0, 0);
        q.extent = make_cudaExtent(NX * sizeof(double), ar1 - ar0, NVAR);
        q.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
        return cudaMemcpy3DAsync(&q, st);
    };
    auto k_lo = [&](int b) { return (int)((long long)nz * b / nbands); };          // first interior row of band b
    auto a_lo = [&](int b) { return b == 0 ? 0 : k_lo(b) + HS; };                   // ... array row (band 0: + halo rows)
    auto a_hi = [&](int b) { return b == nbands - 1 ? nz + 2 * HS : k_lo(b + 1) + HS; };
    // uploads run six rows ahead of the bands: band b's copy ends with the six rows of band b+1 its z sweep reads (and
    // starts after the six it shares with band b-1), so that band b can be swept as soon as its OWN copy has landed
    auto u_lo = [&](int b) { return b == 0 ? 0 : a_lo(b) + SWEEP_HALO; };
    auto u_hi = [&](int b) { return b == nbands - 1 ? nz + 2 * HS : a_hi(b) + SWEEP_HALO; };

    // earlier work of the context (kernels reading or writing the three buffers) before the first upload lands
    CU_TRY(cudaEventRecord(c->hs_start, c->stream));
    CU_TRY(cudaStreamWaitEvent(up, c->hs_start, 0));
    CU_TRY(cudaStreamWaitEvent(down, c->hs_start, 0));
    for (int b = 0; b < nbands; ++b) {
        CU_TRY(copy_band(A, u_lo(b), u_hi(b), true, up));
        CU_TRY(cudaEventRecord(c->hs_ev[b], up));
    }
    c->xhalo_valid[A] = c->xhalo6_valid[A] = false;  // every band's x sweep fills the 6-wide wrap of its rows
    int xdone = 0, rc = PMW_OK;
    for (int b = 0; b < nbands && rc == PMW_OK; ++b) {
        const int k0 = k_lo(b), k1 = k_lo(b + 1);
        CU_TRY(cudaStreamWaitEvent(c->stream, c->hs_ev[b], 0));  // rows up to k1 + 6 are there
        if (!xfirst) {
            rc = launch_sweep(c, PMW_DIR_Z, A, C, B, false, dt, k0, k1);
            c->xhalo_valid[C] = c->xhalo6_valid[C] = true;  // the z sweep stored the periodic images of its rows
            if (rc == PMW_OK) rc = launch_sweep(c, PMW_DIR_X, C, B, B, false, dt, k0, k1);
        } else {
            const int xe = (b == nbands - 1) ? nz : std::min(k1 + SWEEP_HALO, nz);  // the z sweep reads six rows beyond the band
            if (xe > xdone) rc = launch_sweep(c, PMW_DIR_X, A, C, B, false, dt, xdone, xe);
            xdone = std::max(xdone, xe);
            if (rc == PMW_OK) rc = launch_sweep(c, PMW_DIR_Z, C, B, B, false, dt, k0, k1);
        }
        if (rc != PMW_OK) break;
        CU_TRY(cudaEventRecord(c->hs_ev[nbands + b], c->stream));
        CU_TRY(cudaStreamWaitEvent(down, c->hs_ev[nbands + b], 0));
        CU_TRY(copy_band(B, a_lo(b), a_hi(b), false, down));
    }
    const cudaError_t e1 = cudaStreamSynchronize(down), e2 = cudaStreamSynchronize(up), e3 = cudaStreamSynchronize(c->stream);
    if (rc != PMW_OK) return rc;
    CU_TRY(e1);
    CU_TRY(e2);
    CU_TRY(e3);
    c->l2p[PMW_BUF_STATE] = B;  // the new state, with the periodic images of its edge columns
    c->l2p[PMW_BUF_TMP] = A;    // (holds the previous state, not the reference's stage-2 array)
    c->spare = C;
    c->xhalo_valid[B] = c->xhalo6_valid[B] = true;
    c->xhalo_valid[C] = c->xhalo6_valid[C] = false;
    c->reverse = !c->reverse;
    return PMW_OK;
}

extern "C" int pmw_get_reverse_direction(pmw_ctx* c, int* reverse)
{
    BIND_KEEP(c);
    NEED(reverse, "null pointer");
    *reverse = c->reverse;
    return PMW_OK;
}
extern "C" int pmw_set_reverse_direction(pmw_ctx* c, int reverse)
{
    BIND_KEEP(c);
    c->reverse = reverse ? 1 : 0;
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// diagnostics
// ---------------------------------------------------------------------------------------------
extern "C" int pmw_stats_device(pmw_ctx* c, int buf, double* dev_out2)
{
    BIND_KEEP(c);
    ENSURE_TMP_IF(c, buf == PMW_BUF_TMP);
    CHECK_BUF(buf);
    NEED(dev_out2, "pmw_stats_device: null output");
    NEED(c->hydro_set, "pmw_stats: hydrostatic profiles not set");
    // rho cv T = kconst * p (pmw_aux.cuh: cell_energy)
    static const double kconst = (CV / C0) * std::pow(C0 / P0, RD / CP);
    stats_partial_kernel<<<c->stats_blocks, 256, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, c->hy.dens_cell,
                                                                c->hy.dens_theta_cell, c->hy.inv_dens_theta_cell,
                                                                c->hy.pressure_cell, kconst, c->stats_partial);
    LAUNCHED(c, "stats_partial_kernel");
    stats_final_kernel<<<1, 256, 0, c->stream>>>(c->stats_partial, c->stats_blocks, c->p.dx * c->p.dz, dev_out2);
    LAUNCHED(c, "stats_final_kernel");
    return PMW_OK;
}

extern "C" int pmw_stats(pmw_ctx* c, int buf, double out[2])
{
    NEED(out, "pmw_stats: null output");
    int rc = pmw_stats_device(c, buf, c ? c->stats_out : nullptr);
    if (rc != PMW_OK) return rc;
    CU_TRY(cudaMemcpyAsync(out, c->stats_out, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return check_watchdog(c);
}

extern "C" int pmw_solution_variables(pmw_ctx* c, int buf, double* host_out)
{
    BIND_KEEP(c);
    ENSURE_TMP_IF(c, buf == PMW_BUF_TMP);
    CHECK_BUF(buf);
    NEED(host_out, "pmw_solution_variables: null output");
    NEED(c->hydro_set, "pmw_solution_variables: hydrostatic profiles not set");
    const size_t n = (size_t)c->p.nx * c->p.nz;
    double* dev = nullptr;
    CU_TRY(cudaMalloc(&dev, 4 * n * sizeof(double)));
    solution_variables_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        c->base[c->l2p[buf]], c->L, c->hy.dens_cell, c->hy.dens_theta_cell, dev);
    int rc = after_launch(c, "solution_variables_kernel");
    if (rc == PMW_OK) {
        cudaError_t e = cudaMemcpyAsync(host_out, dev, 4 * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(PMW_ECUDA, "solution variables copy failed: %s", cudaGetErrorString(e));
    }
    cudaFree(dev);
    return rc;
}

// ---------------------------------------------------------------------------------------------
// x-slab halo messages
// ---------------------------------------------------------------------------------------------
extern "C" size_t pmw_halo_len(pmw_ctx* c) { return c ? (size_t)NVAR * c->p.nz * 2 : 0; }

extern "C" int pmw_pack_halo_x(pmw_ctx* c, int buf, double* to_left, double* to_right)
{
    BIND(c);
    CHECK_BUF(buf);
    NEED(to_left && to_right, "pmw_pack_halo_x: null message buffer");
    const int n = NVAR * c->p.nz * 2;
    pack_halo_x_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, to_left, to_right);
    LAUNCHED(c, "pack_halo_x_kernel");
    return PMW_OK;
}

extern "C" int pmw_unpack_halo_x(pmw_ctx* c, int buf, const double* from_left, const double* from_right)
{
    BIND(c);
    CHECK_BUF(buf);
    NEED(from_left && from_right, "pmw_unpack_halo_x: null message buffer");
    const int n = NVAR * c->p.nz * 2;
    unpack_halo_x_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, from_left, from_right);
    LAUNCHED(c, "unpack_halo_x_kernel");
    c->xhalo_valid[c->l2p[buf]] = true;
    c->xhalo6_valid[c->l2p[buf]] = false;
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// per-launch timing of the stage kernels
// ---------------------------------------------------------------------------------------------
extern "C" int pmw_stage_timing(pmw_ctx* c, int enable)
{
    BIND(c);
    c->timing = enable != 0;
    c->ev_used = 0;
    return PMW_OK;
}

extern "C" int pmw_stage_timing_read(pmw_ctx* c, double* mean_ms, long long* count)
{
    BIND(c);
    NEED(mean_ms && count, "null pointer");
    CU_TRY(cudaStreamSynchronize(c->stream));
    double total = 0.0;
    long long n = 0;
    for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
        float ms = 0.f;
        CU_TRY(cudaEventElapsedTime(&ms, c->ev[i], c->ev[i + 1]));
        total += ms;
        ++n;
    }
    *mean_ms = n ? total / (double)n : 0.0;
    *count = n;
    c->ev_used = 0;
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// slab ring over peer memory (NVLink P2P stores from the stage kernels)
// ---------------------------------------------------------------------------------------------
extern "C" int pmw_ipc_export(pmw_ctx* c, void* blob)
{
    BIND(c);
    NEED(blob, "pmw_ipc_export: null blob");
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(blob);
    for (int b = 0; b < 3; ++b) CU_TRY(cudaIpcGetMemHandle(&h[b], c->alloc[b]));
    CU_TRY(cudaIpcGetMemHandle(&h[3], c->flags));
    return PMW_OK;
}

extern "C" int pmw_ipc_open(pmw_ctx* c, const void* blob, void* ptrs_out[4])
{
    BIND(c);
    NEED(blob && ptrs_out, "pmw_ipc_open: null argument");
    const cudaIpcMemHandle_t* h = reinterpret_cast<const cudaIpcMemHandle_t*>(blob);
    for (int b = 0; b < 4; ++b) {
        void* p = nullptr;
        CU_TRY(cudaIpcOpenMemHandle(&p, h[b], cudaIpcMemLazyEnablePeerAccess));
        c->ipc_opened.push_back(p);
        ptrs_out[b] = p;
    }
    return PMW_OK;
}

extern "C" int pmw_local_ptrs(pmw_ctx* c, void* ptrs_out[4])
{
    BIND(c);
    NEED(ptrs_out, "pmw_local_ptrs: null argument");
    for (int b = 0; b < 3; ++b) ptrs_out[b] = c->alloc[b];
    ptrs_out[3] = c->flags;
    return PMW_OK;
}

extern "C" int pmw_connect_peers(pmw_ctx* c, void* const left[4], void* const right[4])
{
    BIND(c);
    NEED(!c->p.periodic_x, "pmw_connect_peers: the context was created with periodic_x=1");
    NEED(!c->jet_rows, "pmw_connect_peers: the injection inflow is single-context only");
    NEED((c->p.nx & 1) == 0, "pmw_connect_peers: the slab width must be even");
    NEED(left && right, "pmw_connect_peers: null argument");
    for (int b = 0; b < 3; ++b) {
        NEED(left[b] && right[b], "pmw_connect_peers: null buffer pointer");
        c->nbr_base[0][b] = static_cast<double*>(left[b]) + LPAD;
        c->nbr_base[1][b] = static_cast<double*>(right[b]) + LPAD;
    }
    c->nbr_flags[0] = static_cast<unsigned long long*>(left[3]);
    c->nbr_flags[1] = static_cast<unsigned long long*>(right[3]);
    // same-process neighbours on another device: make their memory addressable from this one
    // (IPC-opened pointers already are)
    for (int side = 0; side < 2; ++side) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, side ? right[0] : left[0]) == cudaSuccess && at.device != c->p.device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(PMW_ECUDA, "cannot enable peer access %d -> %d: %s", c->p.device, at.device,
                            cudaGetErrorString(e));
        }
        cudaGetLastError();
    }
    c->peers = true;
    c->epoch = 0;
    CU_TRY(cudaMemsetAsync(c->flags, 0, 4 * sizeof(unsigned long long), c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    for (int b = 0; b < 3; ++b) c->xhalo_valid[b] = c->xhalo6_valid[b] = false;
    return PMW_OK;
}

// 0 = fine; 1 = an edge tile gave up waiting for a neighbour (results are then invalid).
extern "C" int pmw_peer_status(pmw_ctx* c, int* timed_out)
{
    BIND(c);
    NEED(timed_out, "null pointer");
    unsigned long long f[4];
    CU_TRY(cudaMemcpyAsync(f, c->flags, sizeof(f), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    *timed_out = f[2] != 0;
    return PMW_OK;
}

extern "C" int pmw_fp64_peak(pmw_ctx* c, double* warp_dfma_per_s, double* sm_clock_mhz)
{
    BIND(c);
    NEED(warp_dfma_per_s, "pmw_fp64_peak: null output");
    int nsm = 148, khz = 0;
    CU_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->p.device));
    CU_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->p.device));
    constexpr int NCH = 4, WARPS = 16, ITERS = 2000;
    cudaEvent_t e0, e1;
    CU_TRY(cudaEventCreate(&e0));
    CU_TRY(cudaEventCreate(&e1));
    dfma_probe_kernel<NCH><<<nsm, 32 * WARPS, 0, c->stream>>>(c->stats_out, 50, 0.999, 1e-3);  // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0, c->stream);
        dfma_probe_kernel<NCH><<<nsm, 32 * WARPS, 0, c->stream>>>(c->stats_out, ITERS, 0.999, 1e-3);
        cudaEventRecord(e1, c->stream);
        cudaError_t e = cudaEventSynchronize(e1);
        float ms = 0;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e != cudaSuccess) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return fail(PMW_ECUDA, "pmw_fp64_peak: %s", cudaGetErrorString(e));
        }
        best = std::min(best, ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *warp_dfma_per_s = (double)nsm * WARPS * ITERS * 8.0 * NCH / (best * 1e-3);
    if (sm_clock_mhz) *sm_clock_mhz = khz * 1e-3;
    return PMW_OK;
}

// ---------------------------------------------------------------------------------------------
// unfused operator shims (API parity with the reference's individual operators; not the hot path)
// ---------------------------------------------------------------------------------------------
extern "C" int pmw_interpolate(pmw_ctx* c, int direction, int buf, double* host_vals, double* host_d3)
{
    BIND_KEEP(c);
    ENSURE_TMP_IF(c, buf == PMW_BUF_TMP);
    CHECK_BUF(buf);
    NEED(direction == PMW_DIR_X || direction == PMW_DIR_Z, "pmw_interpolate: bad direction %d", direction);
    NEED(host_vals && host_d3, "pmw_interpolate: null output");
    const bool z = direction == PMW_DIR_Z;
    const size_t n = (size_t)NVAR * (z ? c->p.nz + 1 : c->p.nz) * (z ? c->p.nx : c->p.nx + 1);
    DevBuf vals, d3;
    CU_TRY(vals.alloc(n));
    CU_TRY(d3.alloc(n));
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (z) interpolate_kernel<true><<<blocks, 256, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, vals.p, d3.p);
    else interpolate_kernel<false><<<blocks, 256, 0, c->stream>>>(c->base[c->l2p[buf]], c->L, vals.p, d3.p);
    LAUNCHED(c, "interpolate_kernel");
    CU_TRY(cudaMemcpyAsync(host_vals, vals.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaMemcpyAsync(host_d3, d3.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PMW_OK;
}

extern "C" int pmw_compute_flux(pmw_ctx* c, int direction, const double* host_vals, const double* host_d3,
                                double* host_flux)
{
    BIND(c);
    NEED(direction == PMW_DIR_X || direction == PMW_DIR_Z, "pmw_compute_flux: bad direction %d", direction);
    NEED(host_vals && host_d3 && host_flux, "pmw_compute_flux: null argument");
    NEED(c->hydro_set, "pmw_compute_flux: hydrostatic profiles not set");
    const bool z = direction == PMW_DIR_Z;
    const size_t plane = (size_t)(z ? c->p.nz + 1 : c->p.nz) * (z ? c->p.nx : c->p.nx + 1);
    const size_t nf = (size_t)NVAR * (c->p.nz + 1) * (c->p.nx + 1);
    DevBuf vals, d3, flux;
    CU_TRY(vals.alloc(NVAR * plane));
    CU_TRY(d3.alloc(NVAR * plane));
    CU_TRY(flux.alloc(nf));
    CU_TRY(cudaMemcpyAsync(vals.p, host_vals, NVAR * plane * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(d3.p, host_d3, NVAR * plane * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaMemcpyAsync(flux.p, host_flux, nf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const double d = z ? c->p.dz : c->p.dx;
    const double hv = -HV_BETA * d / (16 * c->p.dt);
    const unsigned blocks = (unsigned)((plane + 255) / 256);
    if (z) flux_kernel<true><<<blocks, 256, 0, c->stream>>>(vals.p, d3.p, c->L, c->hy, hv, flux.p);
    else flux_kernel<false><<<blocks, 256, 0, c->stream>>>(vals.p, d3.p, c->L, c->hy, hv, flux.p);
    LAUNCHED(c, "flux_kernel");
    CU_TRY(cudaMemcpyAsync(host_flux, flux.p, nf * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PMW_OK;
}

extern "C" int pmw_compute_tend(pmw_ctx* c, int direction, const double* host_flux, int state_buf, double* host_tend)
{
    BIND_KEEP(c);
    ENSURE_TMP_IF(c, state_buf == PMW_BUF_TMP);
    CHECK_BUF(state_buf);
    NEED(direction == PMW_DIR_X || direction == PMW_DIR_Z, "pmw_compute_tend: bad direction %d", direction);
    NEED(host_flux && host_tend, "pmw_compute_tend: null argument");
    const bool z = direction == PMW_DIR_Z;
    const size_t nf = (size_t)NVAR * (c->p.nz + 1) * (c->p.nx + 1), nt = (size_t)NVAR * c->p.nz * c->p.nx;
    DevBuf flux, tend;
    CU_TRY(flux.alloc(nf));
    CU_TRY(tend.alloc(nt));
    CU_TRY(cudaMemcpyAsync(flux.p, host_flux, nf * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const unsigned blocks = (unsigned)((nt + 255) / 256);
    if (z) tend_kernel<true><<<blocks, 256, 0, c->stream>>>(flux.p, c->base[c->l2p[state_buf]], c->L, c->p.dz, tend.p);
    else tend_kernel<false><<<blocks, 256, 0, c->stream>>>(flux.p, c->base[c->l2p[state_buf]], c->L, c->p.dx, tend.p);
    LAUNCHED(c, "tend_kernel");
    CU_TRY(cudaMemcpyAsync(host_tend, tend.p, nt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    return PMW_OK;
}
