/*
 * pmw.h -- C ABI of the B200-native PyMiniWeather hot path (libpmw.so).
 *
 * The reference (shriram-jagan/pyminiweather) is pure Python and has no FFI of
 * its own; its only backend seam is the array-module alias in
 * pyminiweather/__init__.py:4-9.  The boundary this library plugs into is
 * therefore the set of Python operator functions the reference's driver calls
 * (pyminiweather/__main__.py:205,214,237,239).  Each entry point below names
 * the reference function (file:line, relative to the reference repo) whose work
 * it performs.  INTEGRATION.md shows the ctypes stub a reference maintainer
 * would add to bind them.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (PMW_OK) on success or a
 *     negative PMW_E* code, and pmw_last_error() then returns a thread-local
 *     human-readable message;
 *   - host state arrays use the reference layout: double [4][nz+4][nx+4], C
 *     order, variables DENS=0, UMOM=1, WMOM=2, RHOT=3
 *     (pyminiweather/__init__.py:14-18, pyminiweather/data/fields.py:67-72);
 *   - a context owns its device buffers; the caller owns every host pointer;
 *   - calls on one context are not thread-safe; contexts are independent;
 *   - all work is enqueued on the context's stream (pmw_set_stream); functions
 *     that return data to the host synchronise that stream.
 */
#ifndef PMW_H_
#define PMW_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMW_OK 0
#define PMW_EINVAL (-1)  /* bad argument */
#define PMW_ECUDA (-2)   /* CUDA runtime/driver error */
#define PMW_ESTATE (-3)  /* call not valid in the current state */

/* logical state buffers (pyminiweather/data/fields.py:20-21) */
#define PMW_BUF_STATE 0 /* fields.state      */
#define PMW_BUF_TMP 1   /* fields.state_tmp  */

/* pyminiweather/ics/directions.py:4-6 */
#define PMW_DIR_X 1
#define PMW_DIR_Z 2

/* kernel variants (pmw_params.variant) */
#define PMW_VARIANT_DIRECT 0 /* one thread per cell, operands through L1/L2        */
#define PMW_VARIANT_TMA 1    /* TMA-staged shared-memory tiles (the production path) */

/* pressure evaluation (pmw_params.pow_mode) */
#define PMW_POW_LIBDEVICE 0 /* p = C0 * pow(rho*theta, gamma), CUDA math library      */
#define PMW_POW_BACKGROUND 1 /* p = P_hy(z) * (1+eps)^gamma via a degree-12 polynomial in
                                eps = (rho*theta - (rho*theta)_hy)/(rho*theta)_hy, falling
                                back to pow() when |eps| > 1/8                         */

typedef struct pmw_ctx pmw_ctx;

typedef struct pmw_params {
    int nx;       /* interior columns held by THIS context (slab width), >= 4       */
    int nz;       /* interior rows, >= 4                                            */
    int hs;       /* halo width; must be 2 (pyminiweather/__main__.py:92-100)       */
    double dx;    /* params["dx"]                                                   */
    double dz;    /* params["dz"]                                                   */
    double dt;    /* params["dt"], the FULL step: the hyperviscosity coefficient uses
                     it in every stage (pyminiweather/solve/interpolate.py:99-101)  */
    int device;   /* CUDA device ordinal                                            */
    int variant;  /* PMW_VARIANT_*                                                  */
    int pow_mode; /* PMW_POW_*                                                      */
    int periodic_x; /* 1: this context wraps its own x halos (single slab, the
                       reference's set_bc_x); 0: x halos arrive from slab neighbours
                       through pmw_unpack_halo_x / peer stores                      */
} pmw_params;

const char *pmw_last_error(void);
int pmw_version(void);

/* -- lifetime ------------------------------------------------------------------ */
/* Allocates three state buffers + profiles on params->device.
 * Replaces the allocation half of initialize_fields (pyminiweather/data/fields.py:58-123);
 * the interpolation/flux/tendency scratch arrays of that function do not exist here. */
int pmw_create(const pmw_params *params, pmw_ctx **out);
int pmw_destroy(pmw_ctx *ctx);
/* cudaStream_t as void*; NULL = the legacy default stream. */
int pmw_set_stream(pmw_ctx *ctx, void *cuda_stream);
int pmw_synchronize(pmw_ctx *ctx);

/* -- data movement ------------------------------------------------------------- */
/* The five 1-D hydrostatic profiles produced by init (pyminiweather/ics/initial.py:84-105):
 * hy_dens_cell[nz+4], hy_dens_theta_cell[nz+4], hy_dens_int[nz+1],
 * hy_dens_theta_int[nz+1], hy_pressure_int[nz+1]. */
int pmw_set_hydrostatic(pmw_ctx *ctx, const double *hy_dens_cell, const double *hy_dens_theta_cell,
                        const double *hy_dens_int, const double *hy_dens_theta_int,
                        const double *hy_pressure_int);
/* ic_type "gravity" only: the constant forcing wpert(x,z)*hy_dens_cell[k] that add_source_terms
 * (pyminiweather/solve/source.py:43-75, called at solve/step.py:78) adds to the rho*w tendency in
 * every stage; host [nz][nx], NULL clears it. */
int pmw_set_source_w(pmw_ctx *ctx, const double *host_nz_nx);
/* ic_type "injection" only: switches pmw_bc_x (and the halo fill inside pmw_evolve /
 * pmw_discrete_step) to the inflow branch of set_bc_x (pyminiweather/ics/bcs.py:37,41-64): the left
 * halo is the periodic image, the right halo is left untouched, and on the jet rows the left halo of
 * rho*u / (rho*theta)' is forced to (rho'+rho_hy)*u_in and (rho'+rho_hy)*theta_in - (rho*theta)_hy.
 * host_rows[nz]: 1 on the interior rows of the jet -- the caller evaluates the reference's row
 * condition (bcs.py:43-48; the reference uses u_in = 50, theta_in = 298).  NULL restores periodic x.
 * Single-context only (periodic_x = 1, no slab ring). */
int pmw_set_inflow(pmw_ctx *ctx, const unsigned char *host_rows, double u_in, double theta_in);
/* Device-side `init` for the 2-D state (pyminiweather/ics/initial.py:57-80): fills BOTH logical buffers
 * (state and state_tmp, halo cells included) by 3x3 Gauss-Legendre quadrature of the configuration
 * described by `spec` -- squared-cosine potential-temperature bubbles (utils/utils.py:5-51) on a
 * constant-theta or constant-Brunt-Vaisala background (ics/initial_conditions.py:26-81) with a uniform
 * wind -- without the reference's 9x-the-state host temporaries.  x_axis[nx+4] / z_axis[nz+4] are the
 * lower-left corner coordinates of the array's columns / rows as the reference's mesh yields them
 * (mesh.py:22-40; a slab passes its own part of the x axis).  The 1-D hydrostatic profiles are still
 * the caller's (pmw_set_hydrostatic). */
#define PMW_IC_MAX_BUBBLES 4
typedef struct pmw_ic_spec {
    int nbubbles;                       /* 0..PMW_IC_MAX_BUBBLES */
    double amp[PMW_IC_MAX_BUBBLES];     /* amplitude [K] */
    double x0[PMW_IC_MAX_BUBBLES], z0[PMW_IC_MAX_BUBBLES];     /* centre */
    double xrad[PMW_IC_MAX_BUBBLES], zrad[PMW_IC_MAX_BUBBLES]; /* radii */
    double wind;                        /* uniform u [m/s] */
    int bvfreq;                         /* 0: theta = theta0; 1: constant Brunt-Vaisala frequency bv0 */
    double bv0;
} pmw_ic_spec;
int pmw_init_state(pmw_ctx *ctx, const pmw_ic_spec *spec, const double *x_axis, const double *z_axis);
/* host [4][nz+4][nx+4] <-> device buffer `buf` (PMW_BUF_*).  Synchronous. */
int pmw_upload_state(pmw_ctx *ctx, int buf, const double *host);
int pmw_download_state(pmw_ctx *ctx, int buf, double *host);
/* Asynchronous variants on the context stream (host memory should be pinned). */
int pmw_upload_state_async(pmw_ctx *ctx, int buf, const double *host);
int pmw_download_state_async(pmw_ctx *ctx, int buf, double *host);

/* -- operators (one call = one reference function) ------------------------------- */
/* set_bc_x (pyminiweather/ics/bcs.py:8-64): periodic branch (:35-39), or the injection branch
 * (:37,41-64) on a context with pmw_set_inflow rows. */
int pmw_bc_x(pmw_ctx *ctx, int buf);
/* set_bc_z (pyminiweather/ics/bcs.py:92-148). */
int pmw_bc_z(pmw_ctx *ctx, int buf);
/* discrete_step WITHOUT its leading set_bc_* call (pyminiweather/solve/step.py:67-82 minus
 * :68/:73): interpolate + flux + tendency (+ hydrostatic source in z) + stage update,
 * fused.  out may alias init and/or forcing exactly as in step.py:112-141. */
int pmw_stage(pmw_ctx *ctx, int direction, int init_buf, int forcing_buf, int out_buf,
              double dt_stage);
/* discrete_step as the reference runs it: set_bc_* on forcing, then pmw_stage
 * (pyminiweather/solve/step.py:21-82). */
int pmw_discrete_step(pmw_ctx *ctx, int direction, int init_buf, int forcing_buf, int out_buf,
                      double dt_stage);
/* nsteps x evolve (pyminiweather/solve/step.py:85-143): two directional sweeps of three RK
 * stages per step, Z first on the first call, order alternating per step; the direction
 * flag (a module global in the reference, step.py:18) lives in the context.  Halo fill is
 * folded into the kernels.  dt <= 0 means params.dt.
 * Default on the TMA variant: ONE kernel per directional sweep (the three RK stages fused, the
 * two intermediate states on chip, 6-cell halo recomputed; bit-identical to running the stages
 * one by one) -- tuning key "fuse".  The reference's state_tmp (its stage-2 array) is then
 * produced by the last sweep of the call only, and by default only when it is asked for ("keep_tmp").  The gravity-wave forcing
 * (pmw_set_source_w) is applied inside the sweeps of a single periodic slab; a context with
 * inflow rows (pmw_set_inflow: x is not periodic) runs the reference's own sequence instead --
 * halo-fill kernel, then fused stage kernel, per stage (pmw_discrete_step x 6 per step). */
int pmw_evolve(pmw_ctx *ctx, int nsteps, double dt);
/* One evolve (pyminiweather/solve/step.py:85-143) on a HOST array -- what the reference's driver does when
 * `fields.state` is a NumPy array (pyminiweather/__main__.py:237): host_state is the dense
 * [4][nz+4][nx+4] array, updated in place.  Same result as pmw_upload_state + pmw_evolve(1) +
 * pmw_download_state (bit for bit), but streamed in `nbands` bands of rows (0 = chosen from the grid): a
 * step only couples rows through the 6-row halo of the z sweep, so the upload of later bands, the two fused
 * sweeps of a band and the download of earlier bands overlap, and the call costs about ONE transfer of
 * the state over PCIe instead of two.  The reference's state_tmp is not produced (the tmp buffer holds
 * the previous state afterwards).  Contexts the banded path does not cover (slab rings, inflow rows,
 * gravity-wave forcing, grids run by the transposing z sweep, "fuse" off) take the plain sequence.
 * Pinned host memory gives the overlap; pageable memory works, more slowly. */
int pmw_evolve_host(pmw_ctx *ctx, double *host_state, double dt, int nbands);
/* One RK stage (rk_stage = 1,2,3) of the fused step on the context's rotating buffers, for
 * callers that interleave their own work between stages (slab halo exchange).  The caller
 * sequences directions/stages as evolve does and flips the direction flag itself. */
int pmw_evolve_stage(pmw_ctx *ctx, int direction, int rk_stage, double dt);
int pmw_get_reverse_direction(pmw_ctx *ctx, int *reverse);
int pmw_set_reverse_direction(pmw_ctx *ctx, int reverse);
/* compute_stats (pyminiweather/post/stats.py:8-35): out[0] = total mass, out[1] = total
 * energy of buffer `buf`, over this context's interior, already scaled by dx*dz. */
int pmw_stats(pmw_ctx *ctx, int buf, double out[2]);
/* Same sums left on the device (2 doubles, already scaled) for an NCCL all-reduce. */
int pmw_stats_device(pmw_ctx *ctx, int buf, double *dev_out2);
/* compute_solution_variables (pyminiweather/post/stats.py:38-69): host [4][nz][nx]. */
int pmw_solution_variables(pmw_ctx *ctx, int buf, double *host_out);

/* Unfused operators, for code written against the reference's individual functions (not the hot
 * path; the fused stage kernels never materialise these arrays).  Dense host arrays in the
 * reference's shapes (pyminiweather/data/fields.py:70-78):
 *   interpolate_x/z   (solve/interpolate.py:10-79):   vals, d3 = [4][nz][nx+1] (x) / [4][nz+1][nx] (z)
 *   compute_flux_x/z  (solve/interpolate.py:82-186):  flux [4][nz+1][nx+1], sub-block updated
 *   compute_tend_x/z  (solve/interpolate.py:189-250): tend [4][nz][nx]; z reads rho' of `state_buf` */
int pmw_interpolate(pmw_ctx *ctx, int direction, int buf, double *host_vals, double *host_d3);
int pmw_compute_flux(pmw_ctx *ctx, int direction, const double *host_vals, const double *host_d3,
                     double *host_flux);
int pmw_compute_tend(pmw_ctx *ctx, int direction, const double *host_flux, int state_buf, double *host_tend);

/* -- x-slab sharding (one context per GPU) ---------------------------------------- */
/* Edge columns of `buf` as two contiguous device messages of pmw_halo_len() doubles each
 * ([4][nz][2]): to_left = my first two interior columns, to_right = my last two.  This is
 * set_bc_x (bcs.py:35-39) generalised to a ring of slabs. */
size_t pmw_halo_len(pmw_ctx *ctx);
int pmw_pack_halo_x(pmw_ctx *ctx, int buf, double *dev_to_left, double *dev_to_right);
int pmw_unpack_halo_x(pmw_ctx *ctx, int buf, const double *dev_from_left,
                      const double *dev_from_right);

/* Peer-memory ring (the production multi-GPU path): instead of pack / send-recv / unpack, the
 * stage kernels store their edge cell pairs straight into the neighbours' halo columns over
 * NVLink and publish a per-stage epoch; the neighbours' x stages schedule their edge tiles last
 * and make them wait for that epoch.  pmw_evolve then works on slab contexts.
 *   same process : pmw_local_ptrs(other_ctx) -> pmw_connect_peers (peer access must be enabled)
 *   one process per GPU: pmw_ipc_export -> exchange the 256-byte blobs -> pmw_ipc_open ->
 *                        pmw_connect_peers.  All ranks must call the stepping functions in the
 *                        same order; after pmw_upload_state on a connected context a collective
 *                        barrier is required before stepping.  */
#define PMW_IPC_BLOB_BYTES 256
int pmw_ipc_export(pmw_ctx *ctx, void *blob);
int pmw_ipc_open(pmw_ctx *ctx, const void *blob, void *ptrs_out[4]);
int pmw_local_ptrs(pmw_ctx *ctx, void *ptrs_out[4]);
int pmw_connect_peers(pmw_ctx *ctx, void *const left[4], void *const right[4]);
int pmw_peer_status(pmw_ctx *ctx, int *timed_out);

/* -- tuning ------------------------------------------------------------------------------
 * Tile shapes of the TMA variant.  Keys: "x_tr" (rows per x tile: 4|8), "x_p" (passes of 64
 * interfaces per x tile row, 1..3; a tile owns 64*x_p-2 cells per row), "z_cfg" (passes of 4
 * interface rows per z tile, 1..8; a tile owns 4*z_cfg-1 cell rows x 64 columns), "pdl"
 * (0|1: programmatic dependent launch between consecutive stage kernels; default 1), "l2_hints"
 * (decimal abcd = L2 eviction priority of: forcing in stage 1, forcing in stages 2-3, init, out;
 * 0 normal, 1 evict_first, 2 evict_last; default 1100), "chunks" (1..4 independent bands per directional sweep, each a kernel chain on
 * its own stream so that kernel tails overlap; default 2 for grids of >= 2^20 cells), "peer_dbg" (development switches).
 * Fused sweeps: "fuse" (0|1, default 1: pmw_evolve launches one kernel per directional sweep), "keep_tmp"
 * (state_tmp after pmw_evolve: 0 never produced, 1 [default] on demand -- the call's last sweep leaves the store
 * out and is run again with it the first time state_tmp is read or a buffer is about to change, so a loop of
 * one-step calls that never looks at state_tmp does not pay for it; 2 written by every call), "sweep_zt" (z sweep
 * organisation: 0 streaming [default], 1 transposing), "sweep_lz" (rows per segment of the streaming z sweep;
 * 0 = chosen from the grid), "sweep_xp" (passes of 64 interfaces per x tile: 2), "dyn_items" (x sweeps draw their
 * work items from a global counter: 0 never, 1 on a slab ring [default], 2 always). */
int pmw_set_tuning(pmw_ctx *ctx, const char *key, int value);
int pmw_get_tuning(pmw_ctx *ctx, const char *key, int *value);

/* -- introspection ------------------------------------------------------------------ */
/* Device address / geometry of a logical buffer (for torch-owned views and tests):
 * element (v,k,i) of the reference layout lives at base[v*vstride + k*pitch + i]. */
int pmw_buffer_info(pmw_ctx *ctx, int buf, void **base, size_t *pitch, size_t *vstride);
/* Number of kernels this context has launched since creation (bench.py: gpu_launches). */
long long pmw_launch_count(pmw_ctx *ctx);
/* Per-launch device timing of the stage kernels: enable, run, then read the mean duration
 * in milliseconds and the number of launches timed.  Uses CUDA events on the context stream. */
int pmw_stage_timing(pmw_ctx *ctx, int enable);
int pmw_stage_timing_read(pmw_ctx *ctx, double *mean_ms, long long *count);
/* Roofline denominator of the fused sweeps (bench.py): the FP64 issue rate this GPU sustains, measured
 * with independent chains of dependent DFMAs (4 chains per thread, 16 warps per SM, every SM) on the
 * context's device.  *warp_dfma_per_s = warp-level DFMA instructions per second, whole chip (x 64 = FLOP/s);
 * *sm_clock_mhz = SM clock the rate was normalised with when reporting instructions per clock (may be NULL). */
int pmw_fp64_peak(pmw_ctx *ctx, double *warp_dfma_per_s, double *sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* PMW_H_ */
